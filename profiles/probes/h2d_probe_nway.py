#!/usr/bin/env python
"""N ranks copying pinned host memory to their GPUs AT THE SAME TIME: the box's aggregate H2D ceiling, with no code of
this repo on the path ("no measurement separates the box from the code").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/probes/h2d_probe_nway.py
"""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
N = 1 << 30
src = torch.empty(N, dtype=torch.uint8).pin_memory()
src.fill_(rank + 1)
dst = torch.empty(N, dtype=torch.uint8, device=dev)
dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
for mode in ("alone", "together"):
    res = torch.zeros(world, device=dev)
    for turn in range(world if mode == "alone" else 1):
        dist.barrier()
        if mode == "together" or turn == rank:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            res[rank] = 4 * N / e0.elapsed_time(e1) / 1e6
        dist.barrier()
    dist.all_reduce(res)
    if rank == 0:
        per = ", ".join(f"{v:.1f}" for v in res.tolist())
        print(f"{world} ranks, each copying 4 x 1 GiB, {mode:8s}: per rank GB/s [{per}]  "
              f"{'sum' if mode == 'together' else 'mean'} {res.sum().item() if mode == 'together' else res.mean().item():.1f} GB/s",
              flush=True)
dist.destroy_process_group()
