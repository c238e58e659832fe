#!/usr/bin/env python
"""One shape, fp32 full forward through the fused kernel: median ms (env switches select the variant)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

K, D, N = (int(a) for a in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "fp32"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N, D, device=dev, generator=g)
q = vq.VectorQuantizerEMA(K, D, print_init=False, search_mode=mode).to(dev).eval()
q.embedding.copy_(E)
cache = q._codebook_cache()
m = vq.quantizer._MODES[mode]
idx = torch.empty(N, dtype=torch.int64, device=dev)
zq, zst = torch.empty(N, D, device=dev), torch.empty(N, D, device=dev)
scratch = torch.zeros(2 + K, dtype=torch.int32, device=dev)
sq, hist = scratch[:2].view(torch.float64), scratch[2:]
reps = 2 if os.environ.get("VQB200_DEBUG") == "2" else 10
for _ in range(2):
    vq.ops.quantize_fused(z, q.embedding, cache, m, idx, zq_out=zq, zq_st_out=zst, sqerr_sum=sq, hist=hist)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
for i in range(reps):
    ev[i].record()
    vq.ops.quantize_fused(z, q.embedding, cache, m, idx, zq_out=zq, zq_st_out=zst, sqerr_sum=sq, hist=hist)
ev[reps].record()
torch.cuda.synchronize()
ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
print(f"BM={os.environ.get('VQB200_FUSED_BM', '256')} grid={os.environ.get('VQB200_FUSED_GRID', 'auto')} "
      f"K={K} D={D} N={N} {mode}: {ms[len(ms) // 2]:.4f} ms")
