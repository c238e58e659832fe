#!/usr/bin/env python
"""How fast can THIS box move 256 MiB of pinned host memory to the GPU?  One copy, chunked copies, two streams,
write-combined pinned memory.  (bench.py's e2e is PCIe-bound: this is its ceiling.)"""
import ctypes
import time

import torch

N = 256 << 20
dev = torch.device("cuda:0")
dst = torch.empty(N, dtype=torch.uint8, device=dev)
src = torch.empty(N, dtype=torch.uint8).pin_memory()
src.fill_(3)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: dst.copy_(src, non_blocking=True))
print(f"one 256 MiB copy:            {ms:.3f} ms  {N / ms / 1e6:.1f} GB/s")
for chunk in (4 << 20, 32 << 20, 128 << 20):
    def f():
        for o in range(0, N, chunk):
            dst[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)
    ms = timed(f)
    print(f"chunks of {chunk >> 20:4d} MiB:          {ms:.3f} ms  {N / ms / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def two():
    h = N // 2
    with torch.cuda.stream(s1):
        dst[:h].copy_(src[:h], non_blocking=True)
    with torch.cuda.stream(s2):
        dst[h:].copy_(src[h:], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


ms = timed(two)
print(f"two streams, halves:         {ms:.3f} ms  {N / ms / 1e6:.1f} GB/s")
rt = ctypes.CDLL("libcudart.so.12")
p = ctypes.c_void_p()
rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(0x04))      # cudaHostAllocWriteCombined
if rc == 0:
    ctypes.memset(p, 3, N)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms = timed(lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), p, ctypes.c_size_t(N), 1, stream))
    print(f"write-combined pinned:       {ms:.3f} ms  {N / ms / 1e6:.1f} GB/s")
else:
    print("cudaHostAlloc(WriteCombined) failed", rc)
back = torch.empty(8 << 20, dtype=torch.uint8).pin_memory()


def both():
    with torch.cuda.stream(s2):
        back.copy_(dst[: 8 << 20], non_blocking=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


ms = timed(both)
print(f"256 MiB H2D + 8 MiB D2H:     {ms:.3f} ms  {N / ms / 1e6:.1f} GB/s")
