#!/usr/bin/env python
"""Smallest program that launches the search kernels once per call on a named shape (for ncu captures).

    python profiles/prof_search.py K D N [mode] [reps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

K, D, N = (int(a) for a in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "fp32"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N, D, device=dev, generator=g)
q = vq.VectorQuantizerEMA(K, D, print_init=False, search_mode=mode).to(dev).eval()
q.embedding.copy_(E)
cache = q._codebook_cache()
idx = torch.empty(N, dtype=torch.int64, device=dev)
m = vq.quantizer._MODES[mode]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
for i in range(reps):
    ev[i].record()
    vq.ops.search(z, q.embedding, cache, 0, m, idx)
ev[reps].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print(f"K={K} D={D} N={N} mode={mode} search ms per call: {['%.3f' % t for t in ms]}  "
      f"-> {2.0 * N * K * D / (min(ms) * 1e-3) / 1e12:.1f} TFLOP/s algorithmic, "
      f"{N * (4 * D + 8) / (min(ms) * 1e-3) / 1e9:.1f} GB/s algorithmic")
