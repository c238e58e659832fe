#!/usr/bin/env python
"""A few training forwards at the stage-2 shape (for ncu captures of rvq_fused_kernel<..., SCATTER> and the refresh passes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
q = vq.VectorQuantizerEMA(1024, 512, num_quantizers=4, print_init=False).to(dev).train()
q.ema_embedding.copy_(q.embedding)                     # the k-means initial state: a healthy codebook
q.ema_cluster_size.fill_(1.0)
for _ in range(6):
    z = torch.randn(128, 64, 512, device=dev, generator=g)
    with torch.no_grad():
        q(z, do_ema_update=True)
torch.cuda.synchronize()
print("ok")
