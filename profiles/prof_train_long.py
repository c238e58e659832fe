#!/usr/bin/env python
"""Long training run of the bare quantizer at the stage-2 shape (fresh random batch every step, no dead-code re-init):
step time and how many row-levels left the certified path, every 50 steps -- the regime where codes that stopped being
used shrink towards zero (models/vq_vae.py:85-88 with cs -> eps).   python profiles/prof_train_long.py [steps] [kmeans]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
kmeans_state = len(sys.argv) > 2 and sys.argv[2] == "kmeans"
dev = torch.device("cuda:0")
K, D, L, N = 1024, 512, 4, 8192
g = torch.Generator(device=dev).manual_seed(1)
q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).train()
if kmeans_state:                                      # the state init_codebook_from_centroids leaves (models/vq_vae.py:606-612)
    q.ema_embedding.copy_(q.embedding)
    q.ema_cluster_size.fill_(1.0)
t_acc, n_acc = 0.0, 0
for step in range(steps):
    z = torch.randn(N // 64, 64, D, device=dev, generator=g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        st, zq, idx, stats = q(z, do_ema_update=True)
    torch.cuda.synchronize()
    t_acc += time.perf_counter() - t0
    n_acc += 1
    if (step + 1) % 50 == 0:
        ws = vq.ops.last_rvq_workspace
        cnt = ws[:8].view(torch.int32).tolist() if ws is not None else [-1, -1]
        alive = int((q.embedding.abs().amax(1) > 0).sum())
        tiny = int(((q.embedding.abs().amax(1) > 0) & (q.embedding.norm(dim=1) < 1e-3 * q.embedding.norm(dim=1).max())).sum())
        print(f"step {step + 1:4d}: {t_acc / n_acc * 1e3:7.3f} ms/forward  exhaustive row-levels {cnt[0]:6d}  uncertified {cnt[1]:6d}  "
              f"perplexity {float(stats[0]):8.1f}  non-zero codes {alive}  of which tiny {tiny}", flush=True)
        t_acc, n_acc = 0.0, 0
