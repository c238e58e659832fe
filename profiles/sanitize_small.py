#!/usr/bin/env python
"""Small forward passes through every kernel family (fused, tcgen05 search, SIMT, RVQ + EMA + backward) for a
compute-sanitizer run:  compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
for (K, D, L, N, train) in [(512, 64, 1, 4096, False), (640, 128, 1, 1024, False), (1024, 512, 2, 512, True),
                            (100, 48, 1, 300, True), (256, 64, 1, 4096 + 128, True)]:
    q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev)
    q.train(train)
    z = torch.randn(N // 4, 4, D, device=dev, generator=g, requires_grad=True)
    st, zq, idx, stats = q(z, do_ema_update=True)
    (st.sum() + q.beta * q.commitment_loss(zq, z)).backward()
    torch.cuda.synchronize()
    print("ok", K, D, L, N, train, float(stats[0]), int(idx.max()))
