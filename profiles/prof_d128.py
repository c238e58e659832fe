#!/usr/bin/env python
"""D = 128: the fused one-kernel forward (BM = 128, one CTA per SM) against the multi-kernel pipeline."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

K, D, N = 512, 128, 1 << 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N // 64, 64, D, device=dev, generator=g)
for flag in ("1", "0"):
    os.environ["VQB200_FUSED_D128"] = flag
    q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
    q.embedding.copy_(E)
    with torch.no_grad():
        for _ in range(3):
            q(z, do_ema_update=False)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        for i in range(10):
            ev[i].record()
            q(z, do_ema_update=False)
        ev[10].record()
        torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
    print(f"K={K} D={D} N={N} fused={flag}: {ms:.4f} ms  {N / ms / 1e6:.3f} G latents/s  "
          f"{N * (12 * D + 8) / ms / 1e6:.0f} GB/s algorithmic (fused path taken: {bool(vq.ops.fused_supported(N, K, D, 0))})")
