#!/usr/bin/env python
"""A few soft assignments at K=1024 D=512 N=8192 (for ncu captures of softmax_stats_kernel / softmax_emit_kernel)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
K, D, N = 1024, 512, 8192
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N, D, device=dev, generator=g)
for _ in range(4):
    vq.ops.soft_assign(z, E, 1.0)
torch.cuda.synchronize()
print("ok")
