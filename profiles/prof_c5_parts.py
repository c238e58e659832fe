#!/usr/bin/env python
"""Where the stage-2 training step goes: each library call of the step timed alone (CUDA events over 200 back-to-back
calls, so host launch time hides behind GPU time where the GPU is the bound)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

dev = torch.device("cuda:0")
K, D, L, N = 1024, 512, 4, 8192
g = torch.Generator(device=dev).manual_seed(1)
q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).train()
z = torch.randn(N // 64, 64, D, device=dev, generator=g, requires_grad=True)


def timed(name, fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:64s} {a.elapsed_time(b) / reps * 1e3:8.1f} us")


def step():
    st, zq, idx, stats = q(z, do_ema_update=True)
    loss = st.sum() * 1e-6 + q.beta * q.commitment_loss(zq, z)
    loss.backward()
    z.grad = None


def fwd_only():
    with torch.no_grad():
        q(z, do_ema_update=True)


timed("full step (forward + EMA + vq_loss + backward), eager", step)
timed("training forward only (no autograd)", fwd_only)
flat = z.detach().reshape(-1, D)
cache = q._codebook_cache()
idx = torch.empty(L * N, dtype=torch.int64, device=dev)
zq, st = torch.empty_like(flat), torch.empty_like(flat)
sq = torch.zeros(1, dtype=torch.float64, device=dev)
hist = torch.zeros(K * L, dtype=torch.int32, device=dev)
timed("ops.rvq_train_forward (refresh 1 + persistent kernel + refresh 2)",
      lambda: vq.ops.rvq_train_forward(flat, q.embedding, cache, 0, q.decay, q.eps, q.ema_cluster_size, q.ema_embedding, idx,
                                       zq, zq_st_out=st, sqerr_sum=sq, hist=hist))
timed("ops.rvq_forward (eval: persistent kernel)",
      lambda: vq.ops.rvq_forward(flat, q.embedding, cache, 0, idx, zq_out=zq, zq_st_out=st, sqerr_sum=sq, hist=hist))
stats3 = torch.empty(3, device=dev)
timed("ops.stats_finalize", lambda: vq.ops.stats_finalize(hist, float(L * N), sq, 1.0 / (N * D), q._ep_usage, q._ep_cnt, stats3))
out = torch.empty_like(flat)
gc = torch.ones((), device=dev)
timed("ops.commit_backward", lambda: vq.ops.commit_backward(st, gc, flat, zq, 2.0 / flat.numel(), out))
seg = torch.zeros(K * L * D + K * L, device=dev)
timed("ops.ema_finalize (one refresh pass over all codes)",
      lambda: vq.ops.ema_finalize(seg[:K * L * D], seg[K * L * D:], q.decay, q.eps, q.ema_cluster_size, q.ema_embedding, q.embedding, cache))
timed("torch zero_ of the segment sums (8 MB)", lambda: seg.zero_())
gt = vq.GraphedTrainStep(q, z)
gst = torch.full_like(z, 1e-6)
timed("full step as one CUDA graph (GraphedTrainStep)", lambda: gt(z, gst, q.beta))
