"""world_size=2 gloo tests of the multi-GPU host logic (the N>1 path) on CPU.

The per-rank quantities a GPU rank would get from the kernels are produced here by the oracle (test
infrastructure); what is under test is the sharding, packing and reduction logic in sharding.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import vq_oracle as O
        from pytorch_vae_b200 import sharding as S
        rs = np.random.RandomState(3)
        K, D, N = 96, 16, 1001
        E = (rs.standard_normal((K, D)) / 4).astype(np.float32)
        z = rs.standard_normal((N, D)).astype(np.float32)
        # (1) row sharding + ONE stats all-reduce == single-process statistics
        s, e = S.shard_rows(N, world, rank)
        idx = O.nearest_code(z[s:e], E)
        hist = torch.from_numpy(np.bincount(idx, minlength=K).astype(np.int32))
        sq = torch.tensor([float(((E[idx] - z[s:e]).astype(np.float64) ** 2).sum())], dtype=torch.float64)
        mean, ghist = S.allreduce_stats(sq, (e - s) * D, hist)
        full_idx = O.nearest_code(z, E)
        usage, ppl, dead = O.usage_stats(full_idx, K)
        ok = np.array_equal(ghist.numpy(), usage.astype(np.int32))
        ok &= abs(float(mean) - float(O.commitment_mse(E[full_idx], z))) < 1e-6
        u2, ppl2, dead2 = O.usage_stats(np.repeat(np.arange(K), ghist.numpy()), K)
        ok &= abs(float(ppl2) - float(ppl)) < 1e-4 and float(dead2) == float(dead)
        # (2) codebook sharding + packed min-loc all-reduce(MIN) == full argmin, ties to the lowest index
        E2 = E.copy()
        E2[70] = E2[5]                                     # a twin living in the other rank's shard
        cs, ce = S.shard_codes(K, world, rank)
        d_full = O.distances_fp32(z, E2)                    # one product, sliced: shards see identical values
        d = d_full[:, cs:ce]
        loc = d.argmin(1)
        dv = d[np.arange(N), loc].astype(np.float32)
        bits = dv.view(np.uint32).astype(np.uint64)
        key = np.where(bits & 0x80000000, ~bits & 0xFFFFFFFF, bits | 0x80000000)
        packed = ((key << np.uint64(32)) | (loc + cs).astype(np.uint64)).view(np.int64)
        red = S.allreduce_minloc(torch.from_numpy(packed.copy()))
        got = (red.numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
        ok &= np.array_equal(got, O.argmin_first(d_full))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
