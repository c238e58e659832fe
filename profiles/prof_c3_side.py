#!/usr/bin/env python
"""c3 (K=8192 D=256 N=2^22) step and tensor-kernel time for the variants of the chunk pipeline:

    python profiles/prof_c3_side.py            # runs itself once per variant (VQB200_TC2_SIDE = 0, i, p, g, 1)

0 = plain pipeline (separate pre-pass / gather kernels), i = sixteen-warp kernel with idle side warps,
p / g = side warps run only the pre-pass / only the gather, 1 = both.  With VQB200_DEBUG=4 the library also prints
the effective SM clock of every pair-kernel launch (clock64 / globaltimer), which separates "the kernel needs more
cycles" from "the clock dropped under the power cap".
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch

    import pytorch_vae_b200 as vq
    K, D, N = 8192, 256, 1 << 22
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
    q.embedding.copy_(torch.randn(K, D, device=dev, generator=g) / np.sqrt(D))
    z = torch.randn(N // 64, 64, D, device=dev, generator=g)
    lib = vq._cabi.lib
    with torch.no_grad():
        for _ in range(3):
            q(z, do_ema_update=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            q(z, do_ema_update=False)
        e1.record()
        torch.cuda.synchronize()
        step = e0.elapsed_time(e1) / 8
        lib.vqb200_timing_enable(1)
        for _ in range(4):
            q(z, do_ema_update=False)
        torch.cuda.synchronize()
        lib.vqb200_timing_enable(0)
    tot, nl = ctypes.c_float(0), ctypes.c_int(0)
    lib.vqb200_timing_collect(ctypes.byref(tot), ctypes.byref(nl))
    print(f"VQB200_TC2_SIDE={os.environ.get('VQB200_TC2_SIDE', '1')}: step {step:.3f} ms, tensor kernel "
          f"{tot.value / max(1, nl.value):.3f} ms x {nl.value // 4} per step", flush=True)
    if os.environ.get("PROF_CLK") == "1":
        os.environ["VQB200_DEBUG"] = "4"
        with torch.no_grad():
            q(z, do_ema_update=False)
        torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for v in (sys.argv[1:] or ["0", "i", "p", "g", "1"]):
            env = dict(os.environ, VQB200_TC2_SIDE=v, PROF_CLK="1")
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env)
