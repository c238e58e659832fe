"""Multi-GPU checks run under torchrun (one process per GPU, NCCL); launched by the gpu tests when >= 2 GPUs exist,
or by hand:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_worker.py all
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sharded_search(dev, rank, world):
    import pytorch_vae_b200 as vq
    from pytorch_vae_b200 import sharding as S
    K, D, N = 8192, 256, 50000
    gen = torch.Generator(device=dev).manual_seed(11)             # identical tensors on every rank
    E = torch.randn(K, D, device=dev, generator=gen) / np.sqrt(D)
    E[K // 2 + 7] = E[3]                                          # a twin in the other rank's slice
    z = torch.randn(N, D, device=dev, generator=gen)
    z[:64] = E[3] + 0.01 * torch.randn(64, D, device=dev, generator=gen)
    q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
    q.embedding.copy_(E)
    with torch.no_grad():
        full = q(z.view(1, N, D), do_ema_update=False)[2].view(-1)
        got = S.codebook_sharded_search(q, z)
    torch.cuda.synchronize()
    assert torch.equal(got, full), f"rank {rank}: {(got != full).sum().item()} rows differ"
    assert bool((got[:64] == 3).all())
    gathered = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(gathered, got)
    assert all(torch.equal(g, got) for g in gathered)             # identical on every rank
    if rank == 0:
        print("sharded_search ok", flush=True)


def row_sharded_stats(dev, rank, world):
    """Rows sharded, codebook replicated, stats_sync: the all-reduced statistics equal the single-GPU statistics of the
    concatenated batch; the indices equal the replicated run's slice."""
    import pytorch_vae_b200 as vq
    from pytorch_vae_b200 import sharding as S
    K, D, N = 512, 64, 1 << 16
    gen = torch.Generator(device=dev).manual_seed(5)
    E = torch.randn(K, D, device=dev, generator=gen) / np.sqrt(D)
    z = torch.randn(N // 64, 64, D, device=dev, generator=gen)
    q1 = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
    q1.embedding.copy_(E)
    with torch.no_grad():
        st, zq, idx, stats = q1(z, do_ema_update=False)
    a, b = S.shard_rows(N // 64, world, rank)
    seen = {}
    for how in ("peer", "nccl"):                                  # one kernel over NVLink peer memory / pack -> NCCL -> finalize
        q2 = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
        q2.embedding.copy_(E)
        q2.stats_sync = True
        q2.stats_exchange = how
        for rep in range(3):                                      # epochs 1..3: both payload areas and their reuse
            with torch.no_grad():
                st2, zq2, idx2, stats2 = q2(z[a:b].contiguous(), do_ema_update=False)
            torch.cuda.synchronize()
            assert torch.equal(idx2, idx[a:b]) and torch.equal(zq2, zq[a:b])
            assert torch.allclose(stats2, stats, rtol=1e-5), (how, rep, stats2, stats)
            assert abs(float(q2.last_commit) - float(q1.last_commit)) < 1e-5 * float(q1.last_commit)
        assert torch.equal(q2._ep_usage, 3 * q1._ep_usage) and float(q2._ep_cnt) == 3 * float(q1._ep_cnt) == 3 * N
        seen[how] = stats2.clone()
        both = [torch.empty_like(stats2) for _ in range(world)]
        dist.all_gather(both, stats2)
        assert all(torch.equal(g, stats2) for g in both), how     # every rank holds the same statistics, bit for bit
    if S.PeerStatsExchange.get(dev, K) is None and rank == 0:
        print("row_sharded_stats: peer memory unavailable on this box, NCCL path only", flush=True)
    assert torch.allclose(seen["peer"], seen["nccl"], rtol=1e-6)
    if rank == 0:
        print("row_sharded_stats ok", flush=True)


def allreduced_ema_training(dev, rank, world):
    """Data-parallel training with the EMA segment sums all-reduced (ema_sync="allreduce", ONE exchange per step through
    vqb200_rvq_train_begin / _finish): every rank ends each step with the codebook a single GPU gets from the
    concatenated batch; indices of a rank's rows equal that run's slice (step 0 exactly; later steps except near-ties)."""
    import pytorch_vae_b200 as vq
    from pytorch_vae_b200 import sharding as S
    K_per, D, L, B, M = 1024, 512, 4, 128, 64
    gen = torch.Generator(device=dev).manual_seed(21)
    E = torch.randn(K_per * L, D, device=dev, generator=gen) / np.sqrt(D)
    for l in range(1, L):
        E[l * K_per:(l + 1) * K_per] *= 0.6 ** l
    zs = [torch.randn(B, M, D, device=dev, generator=gen) for _ in range(3)]
    assert vq.ops.rvq_train_fused_supported(B * M // world, K_per, D, L, 0)
    q1 = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, decay=0.98).to(dev).train()
    q1.embedding.copy_(E)
    q2 = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, decay=0.98).to(dev).train()
    q2.embedding.copy_(E)
    q2.ema_sync = "allreduce"
    a, b = S.shard_rows(B, world, rank)
    n1, n2 = B * M, (b - a) * M
    problems = []                                                 # collected, agreed on by all ranks, raised at the end:
    for step, z in enumerate(zs):                                 # an assert between collectives would hang the others
        idx1 = q1(z, do_ema_update=True)[2].view(L, n1)
        idx2 = q2(z[a:b].contiguous(), do_ema_update=True)[2].view(L, n2)
        same = (idx2 == idx1[:, a * M:b * M]).all(0).float().mean().item()
        # step 0: identical codebooks, identical rows -> identical codes.  Later steps search codebooks whose segment
        # sums were added in a different order (per-rank partial sums): near-ties may move
        if not (same == 1.0 if step == 0 else same > 0.98):
            problems.append(f"step {step}: only {same:.4f} of the rows agree with the single-GPU run")
        scale = float(q1.ema_embedding.abs().max())
        if step == 0:                                             # (later steps inherit the moved near-ties)
            if not torch.allclose(q2.ema_cluster_size, q1.ema_cluster_size, rtol=1e-5, atol=1e-6):
                problems.append(f"step {step}: ema_cluster_size differs")
            if not torch.allclose(q2.ema_embedding, q1.ema_embedding, rtol=1e-4, atol=2e-6 * scale + 1e-7):
                problems.append(f"step {step}: ema_embedding differs")
            if not torch.allclose(q2.embedding, q1.embedding, rtol=1e-4, atol=2e-6 * float(q1.embedding.abs().max()) + 1e-7):
                problems.append(f"step {step}: embedding differs")
    gathered = [torch.empty_like(q2.embedding) for _ in range(world)]
    dist.all_gather(gathered, q2.embedding)
    if not all(torch.equal(g, q2.embedding) for g in gathered):   # every rank holds the same codebook, bit for bit
        problems.append("ranks hold different codebooks")
    bad = torch.tensor([len(problems)], device=dev)
    dist.all_reduce(bad)
    assert int(bad) == 0, f"rank {rank}: {problems}"
    if rank == 0:
        print("allreduced_ema_training ok", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    try:
        if what in ("sharded_search", "all"):
            sharded_search(dev, rank, world)
        if what in ("row_sharded_stats", "all"):
            row_sharded_stats(dev, rank, world)
        if what in ("allreduced_ema_training", "all"):
            allreduced_ema_training(dev, rank, world)
    finally:
        dist.destroy_process_group()
