#!/usr/bin/env python
"""Times the row kernels alone (CUDA events) with outputs toggled, to attribute HBM time.

    python profiles/bench_rowops.py K D N
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

K, D, N = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N, D, device=dev, generator=g)
idx = torch.randint(0, K, (N,), device=dev, generator=g)
zq, st = torch.empty_like(z), torch.empty_like(z)
sq = torch.zeros(1, dtype=torch.float64, device=dev)
hist = torch.zeros(K, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, bytes_moved, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                       # evict L2 between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = float(np.median(ts))
    print(f"{name:38s} {t * 1e3:9.1f} us  {bytes_moved / t / 1e6:8.1f} GB/s")


row = D * 4
timeit("gather zq+st+sq+hist", lambda: vq.ops.gather(z, E, idx, zq_out=zq, zq_st_out=st, sqerr_sum=sq, hist=hist), N * (3 * row + 8))
timeit("gather zq+st+sq", lambda: vq.ops.gather(z, E, idx, zq_out=zq, zq_st_out=st, sqerr_sum=sq), N * (3 * row + 8))
timeit("gather zq+st", lambda: vq.ops.gather(z, E, idx, zq_out=zq, zq_st_out=st), N * (3 * row + 8))
timeit("gather zq only", lambda: vq.ops.gather(z, E, idx, zq_out=zq), N * (2 * row + 8))
timeit("gather hist only", lambda: vq.ops.gather(z, E, idx, hist=hist), N * (row + 8))
timeit("st_loss st+sq", lambda: vq.ops.st_loss(z, zq, st, sq), N * 3 * row)
timeit("torch copy z->zq (reference)", lambda: zq.copy_(z), N * 2 * row)
