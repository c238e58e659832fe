#!/usr/bin/env python
"""Codebook-sharded search on the tensor path over G GPUs (north star: "a codebook-sharded min-loc reduction"): every rank
holds the same rows and scans K / G codes with the tcgen05 kernels + exact re-rank, packs (40-bit exact score | 24-bit id)
and ONE all-reduce(MIN) of N words picks the winner.  Prints one JSON line (rank 0) in bench.py's vocabulary.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 profiles/bench_codebook_sharded.py [K D N]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402
from pytorch_vae_b200 import sharding as S  # noqa: E402

K, D, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 256, 1 << 20)
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
g = torch.Generator(device=dev).manual_seed(1234)                 # identical rows and codebook on every rank
E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
z = torch.randn(N, D, device=dev, generator=g)
q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
q.embedding.copy_(E)
with torch.no_grad():
    full = q(z.view(N // 64, 64, D), do_ema_update=False)[2].view(-1)      # replicated codebook: the answer
    for _ in range(3):
        got = S.codebook_sharded_search(q, z)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        got = S.codebook_sharded_search(q, z)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    a.record()
    for _ in range(steps):
        q(z.view(N // 64, 64, D), do_ema_update=False)
    b.record()
    torch.cuda.synchronize()
    ms_full = a.elapsed_time(b) / steps
same = bool(torch.equal(got, full))
if rank == 0:
    print(json.dumps({
        "metric": "latents quantized/sec (codebook-sharded distance+argmin, indices only)", "value": N / (float(ms) * 1e-3),
        "unit": "latents/s", "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": float(ms), "higher_is_better": True,
        "scaling": "strong (codes)", "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"codebook-sharded search K={K} D={D} N={N}", "K": K, "D": D, "rows": N,
                   "parallelism": f"codebook sharded x{world} ({K // world} codes per GPU), rows replicated, one all-reduce(MIN) of N int64"},
        "identical_to_replicated_search": same,
        "replicated_full_forward_ms_one_gpu": ms_full}))
if world > 1:
    dist.destroy_process_group()
