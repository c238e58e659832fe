#!/usr/bin/env python
"""Smallest program that runs the full single-level quantizer forward a few times (for ncu captures).

    python profiles/prof_forward.py K D N [mode] [reps]
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
z = torch.randn(N // 64, 64, D, device=dev, generator=g)
q = vq.VectorQuantizerEMA(K, D, print_init=False, search_mode=mode).to(dev).eval()
q.embedding.copy_(E)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
with torch.no_grad():
    for i in range(reps):
        ev[i].record()
        q(z, do_ema_update=False)
    ev[reps].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print(f"K={K} D={D} N={N} mode={mode} forward ms: {['%.3f' % t for t in ms]} -> {N / (min(ms) * 1e-3) / 1e9:.3f} G latents/s, "
      f"{N * (12 * D + 8) / (min(ms) * 1e-3) / 1e9:.0f} GB/s algorithmic")
if vq.ops.last_fused_workspace is not None:
    print("rows handed back to the exact kernel in the last fused forward:",
          int(vq.ops.last_fused_workspace[:4].view(torch.int32)[0]))
