#!/usr/bin/env python
"""Per-role timeline of one CTA of the fused kernel (VQB200_DEBUG=3): clock64 stamps of the TMA, MMA, first epilogue and
first converter/output warp for a few row tiles, printed by the library as cycles since the CTA's first stamp.

    python profiles/prof_fused_trace.py [K] [D] [N] [full|codes]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
full = (sys.argv[4] if len(sys.argv) > 4 else "full") == "full"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N, D, device=dev, generator=g)
q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
q.embedding.copy_(E)
cache = q._codebook_cache()
idx = torch.empty(N, dtype=torch.int64, device=dev)
zq = torch.empty(N, D, device=dev) if full else None
zst = torch.empty(N, D, device=dev) if full else None
scratch = torch.zeros(2 + K, dtype=torch.int32, device=dev)
sq = scratch[:2].view(torch.float64) if full else None
for _ in range(3):
    vq.ops.quantize_fused(z, q.embedding, cache, 0, idx, zq_out=zq, zq_st_out=zst, sqerr_sum=sq, hist=scratch[2:])
torch.cuda.synchronize()
os.environ["VQB200_DEBUG"] = "3"
vq.ops.quantize_fused(z, q.embedding, cache, 0, idx, zq_out=zq, zq_st_out=zst, sqerr_sum=sq, hist=scratch[2:])
torch.cuda.synchronize()
