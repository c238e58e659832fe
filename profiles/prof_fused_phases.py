#!/usr/bin/env python
"""Phase decomposition of the fused small-D forward by switching outputs / modes / input distributions.

    python profiles/prof_fused_phases.py [K] [D] [N]

Variants (same launch, different optional outputs): full forward, codes only (no z_q / z_q_st / loss),
fp32 vs bf16_input (bf16_input has a ~100x smaller admission margin, so almost no row needs the exact
re-rank: the difference to fp32 is the re-rank's cost), randn vs clustered latents.
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
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z_randn = torch.randn(N, D, device=dev, generator=g)
z_clu = E[torch.randint(0, K, (N,), device=dev, generator=g)] + 0.1 / np.sqrt(D) * torch.randn(N, D, device=dev, generator=g)


def run(z, mode, full, reps=10):
    q = vq.VectorQuantizerEMA(K, D, print_init=False, search_mode=mode).to(dev).eval()
    q.embedding.copy_(E)
    cache = q._codebook_cache()
    m = vq.quantizer._MODES[mode]
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    zq = torch.empty(N, D, device=dev) if full else None
    zst = torch.empty(N, D, device=dev) if full else None
    scratch = torch.zeros(2 + K, dtype=torch.int32, device=dev)
    sq = scratch[:2].view(torch.float64) if full else None
    hist = scratch[2:]
    for _ in range(3):
        vq.ops.quantize_fused(z, q.embedding, cache, m, idx, zq_out=zq, zq_st_out=zst, sqerr_sum=sq, hist=hist)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for i in range(reps):
        ev[i].record()
        vq.ops.quantize_fused(z, q.embedding, cache, m, idx, zq_out=zq, zq_st_out=zst, sqerr_sum=sq, hist=hist)
    ev[reps].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    fb = int(vq.ops.last_fused_workspace[:4].view(torch.int32)[0])
    return ms[len(ms) // 2], fb


if not vq.ops.fused_supported(N, K, D, 0):
    print("shape does not take the fused path")
    sys.exit(0)
for zname, z in (("randn", z_randn), ("clustered", z_clu)):
    for mode in ("fp32", "bf16_input"):
        for full in (True, False):
            ms, fb = run(z, mode, full)
            by = N * ((12 * D + 8) if full else (4 * D + 8))
            print(f"K={K} D={D} N={N} {zname:9s} {mode:10s} {'full ' if full else 'codes'}: {ms:.4f} ms  "
                  f"{N / ms / 1e6:.3f} G latents/s  {by / ms / 1e6:.0f} GB/s algorithmic  handed back {fb}")
