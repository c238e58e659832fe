#!/usr/bin/env python
"""Times the auxiliary kernels of the path alone (CUDA events, L2 flushed between launches), each against the roofline
that bounds it: soft assignment and the usage-entropy regulariser (SIMT fp32 FMA: 4 K D flop per row), indices ->
decoder memory (HBM / L2: Q 4H bytes gathered + 4H written per token), the codebook refresh passes of a training step.

    python profiles/bench_aux.py          # stage-2 shape: 4 x 1024 codes, D = 512, 8192 rows
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
FP32_PEAK_TF = 148 * 128 * 2 * 1.965e9 / 1e12          # 148 SMs x 128 FMA lanes x 2 flop x max clock = 74.4 TFLOP/s


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def line(name, ms, flop=None, byts=None):
    s = f"{name:58s} {ms * 1e3:9.1f} us"
    if flop:
        s += f"  {flop / ms / 1e9:8.1f} TFLOP/s fp32 = {flop / ms / 1e9 / FP32_PEAK_TF:5.1%} of the FMA peak"
    if byts:
        s += f"  {byts / ms / 1e6:8.1f} GB/s = {byts / ms / 1e6 / 6540.2:5.1%} of the measured HBM peak"
    print(s)


for (K, D, N) in [(512, 64, 8192), (1024, 512, 8192), (512, 64, 1 << 17)]:
    E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
    z = torch.randn(N, D, device=dev, generator=g)
    line(f"soft_assign K={K} D={D} N={N} (tiled softmax + GEMM)", timeit(lambda: vq.ops.soft_assign(z, E, 1.0)), flop=4.0 * K * D * N)
    line(f"  first version: one warp per row", timeit(lambda: vq.ops.soft_assign_warp(z, E, 1.0)), flop=4.0 * K * D * N)
    line(f"usage_probs K={K} D={D} N={N} (tiled softmax, column sums)", timeit(lambda: vq.ops.usage_probs(z, E)), flop=4.0 * K * D * N)
    p, P = vq.ops.usage_probs(z, E, keep_probs=True)
    gp = torch.randn(K, device=dev, generator=g)
    line(f"usage_probs backward from the saved probabilities", timeit(lambda: vq.ops.usage_probs_backward_from_probs(P, E, gp)), flop=2.0 * K * D * N)
    p2, rs = vq.ops.usage_probs_warp(z, E)
    line(f"  first version: forward", timeit(lambda: vq.ops.usage_probs_warp(z, E)), flop=4.0 * K * D * N)
    line(f"  first version: backward", timeit(lambda: vq.ops.usage_probs_backward(z, E, rs, gp)), flop=4.0 * K * D * N)

K_per, D, L, H, n_tok = 1024, 512, 4, 512, 8192
E = torch.randn(K_per * L, D, device=dev, generator=g) / np.sqrt(D)
W = torch.randn(H, D, device=dev, generator=g) / np.sqrt(D)
b = torch.randn(H, device=dev, generator=g)
lw, lb = torch.ones(H, device=dev), torch.zeros(H, device=dev)
P = (E @ W.t()).contiguous()
ids = (torch.randint(0, K_per, (n_tok, L), device=dev, generator=g) + torch.arange(L, device=dev) * K_per)
lin, ln = torch.nn.Linear(D, H).to(dev), torch.nn.LayerNorm(H).to(dev)
with torch.no_grad():
    lin.weight.copy_(W); lin.bias.copy_(b)
    line(f"indices_to_memory Q={L} H={H} tokens={n_tok}", timeit(lambda: vq.ops.indices_to_memory(ids, P, L, b, lw, lb)),
         byts=n_tok * (L * 4 * H + 4 * H + 8 * L))
    line("  torch chain: indices_to_latent -> Linear (fp32 cuBLAS) -> LayerNorm",
         timeit(lambda: ln(lin(vq.ops.indices_to_latent(ids, E, L)))), flop=2.0 * n_tok * D * H)
