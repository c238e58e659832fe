"""The tcgen05 search path against the exact SIMT path and the fp64 oracle: same inputs, both
kernels (VQB200_FORCE_SIMT=1 switches the dispatcher), including the rows the tensor path must hand
back to the exact kernel (non-finite values, collapsed codebooks that overflow the candidate list)."""
import os

import numpy as np
import pytest
import torch
from synth import large_case_inputs

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vq():
    import pytorch_vae_b200 as m
    return m


def run_search(vq, z, E, K_per, mode="fp32", force_simt=False):
    dev = torch.device("cuda:0")
    D = z.shape[-1]
    L = E.shape[0] // K_per
    q = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, search_mode=mode).to(dev).eval()
    q.embedding.copy_(torch.from_numpy(E).to(dev))
    old = os.environ.get("VQB200_FORCE_SIMT")
    os.environ["VQB200_FORCE_SIMT"] = "1" if force_simt else "0"
    try:
        path = vq._cabi.lib.vqb200_search_path(z.shape[0] * z.shape[1], K_per, D, 0)
        out = q(torch.from_numpy(z).to(dev), do_ema_update=False)
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("VQB200_FORCE_SIMT", None)
        else:
            os.environ["VQB200_FORCE_SIMT"] = old
    return path, out[2].cpu().numpy().reshape(-1), q


SHAPES = [
    # K,   D,   N     (N as B*64 rows)
    (128, 64, 64),            # one code tile, one partial row tile
    (512, 64, 8192),          # C2 shape
    (520, 64, 300 * 64 // 64 * 64),   # K tail (not a multiple of 128)
    (2048, 128, 4096),
    (8192, 256, 2048),        # C3 shape, code splits in play (few row tiles)
    (8192, 256, 40000 // 64 * 64),    # C3 shape, persistent loop with >1 item per CTA
    (1024, 512, 4096),        # stage-2 level shape (BM=128 variant)
    (384, 192, 1024),         # D = 3 swizzle blocks
]


@pytest.mark.parametrize("K,D,N", SHAPES)
def test_tc_matches_exact(vq, K, D, N):
    E, z = large_case_inputs(700 + K + D, K, D, 1, N // 64, 64)
    path_tc, idx_tc, _ = run_search(vq, z, E, K)
    path_s, idx_s, _ = run_search(vq, z, E, K, force_simt=True)
    assert path_tc == 1 and path_s == 0
    ref = O.nearest_code64(z.reshape(-1, D), E)
    flat = z.reshape(-1, D)
    mm, outside = O.near_tie_rows(flat, E, idx_tc, ref)
    assert outside.size == 0, f"tensor path: {outside.size} rows outside the allowance (of {mm.size} mismatches)"
    assert mm.size <= 2
    mm2, outside2 = O.near_tie_rows(flat, E, idx_s, ref)
    assert outside2.size == 0 and mm2.size <= 2


@pytest.mark.parametrize("K,D,N", [(512, 64, 8192), (4096, 256, 2048)])
def test_tc_bf16_mode(vq, K, D, N):
    E, z = large_case_inputs(800 + K, K, D, 1, N // 64, 64)
    path, idx, _ = run_search(vq, z, E, K, mode="bf16_input")
    assert path == 1
    zb = torch.from_numpy(z).bfloat16().float().numpy().reshape(-1, D)
    Eb = torch.from_numpy(E).bfloat16().float().numpy()
    ref = O.nearest_code64(zb, Eb)
    mm, outside = O.near_tie_rows(zb, Eb, idx, ref)
    assert outside.size == 0 and mm.size <= 2


def test_tc_clustered_and_scaled(vq):
    # trained-model regime (z near a code) and badly scaled latents (|z| >> |e|): the margin scales with |z|
    for seed, scale, clustered in ((1, None, True), (2, 25.0, False), (3, 1e-3, False)):
        E, z = large_case_inputs(seed, 1024, 128, 1, 64, 64, scale, clustered)
        path, idx, _ = run_search(vq, z, E, 1024)
        assert path == 1
        ref = O.nearest_code64(z.reshape(-1, 128), E)
        mm, outside = O.near_tie_rows(z.reshape(-1, 128), E, idx, ref)
        assert outside.size == 0 and mm.size <= 2


def test_tc_hands_back_hard_rows(vq):
    K, D = 256, 64
    E, z = large_case_inputs(5, K, D, 1, 8, 64)
    z = z.copy()
    z[0, 3, 5] = np.nan
    z[1, 7, 0] = np.inf
    z[2, 1, 2] = -np.inf
    z[3, 0] = 0.0                                          # zero row: margin degenerates to ~0
    path, idx_tc, _ = run_search(vq, z, E, K)
    _, idx_s, _ = run_search(vq, z, E, K, force_simt=True)
    assert path == 1 and np.array_equal(idx_tc, idx_s)
    assert np.array_equal(idx_tc, O.nearest_code(z.reshape(-1, D), E))
    # NaN in the codebook: every row goes to the exact kernel; first NaN code wins everywhere
    E2 = E.copy()
    E2[77, 3] = np.nan
    _, idx2, _ = run_search(vq, z, E2, K)
    assert (idx2[np.isfinite(z.reshape(-1, D)).all(1)] == 77).all()
    # collapsed codebook: 200 identical zero codes overflow the 32-slot list -> exact kernel, lowest twin
    E3 = E.copy()
    E3[40:240] = 0.0
    _, idx3, _ = run_search(vq, z, E3, K)
    _, idx3s, _ = run_search(vq, z, E3, K, force_simt=True)
    assert np.array_equal(idx3, idx3s)
    fin = np.isfinite(z.reshape(-1, D)).all(1)
    assert np.array_equal(idx3[fin], O.nearest_code64(z.reshape(-1, D)[fin], E3))
    # exact duplicates of a WINNING code: lowest index
    E4 = E.copy()
    E4[200] = E4[17]
    z4 = z.copy()
    z4[4, :] = E4[17] * 1.01
    _, idx4, _ = run_search(vq, z4, E4, K)
    assert (idx4[4 * 64:5 * 64] == 17).all()


def run_with_env(vq, z, E, K, env, mode="fp32"):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return run_search(vq, z, E, K, mode=mode)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("K,D,N", [(1000, 128, 20480), (8192, 256, 39936), (2048, 512, 19200), (256, 192, 32768)])
def test_cta_pair_kernel_matches_single_cta(vq, K, D, N):
    """tcgen05.mma.cta_group::2 variant (VQB200_TC2=1, default for large N) == 1-CTA kernel == fp64 oracle."""
    E, z = large_case_inputs(900 + K + D, K, D, 1, N // 64, 64)
    _, idx2, _ = run_with_env(vq, z, E, K, {"VQB200_TC2": "1", "VQB200_NO_FUSED": "1"})
    _, idx1, _ = run_with_env(vq, z, E, K, {"VQB200_TC2": "0", "VQB200_NO_FUSED": "1"})
    assert np.array_equal(idx1, idx2)
    ref = O.nearest_code64(z.reshape(-1, D), E)
    mm, outside = O.near_tie_rows(z.reshape(-1, D), E, idx2, ref)
    assert outside.size == 0 and mm.size <= 2


def _bf16(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).bfloat16().float().numpy()


@pytest.mark.parametrize("fused", [True, False])
def test_margin_holds_when_rounding_errors_align(vq, fused):
    """Adversarial rounding: every element of the latent and of code a sits just BELOW a bf16 rounding
    midpoint on dims 0..31 (their product is under-estimated by 2u, u = 2^-8) and just ABOVE one on dims
    32..62 for code b (over-estimated by 2u).  The approximate scores prefer b by ~0.44 although a wins
    exactly by ~0.047 -- more than the worst-case-halved bound 2(2u'+u'^2)|z||e| with u' = 2^-9 admits, less
    than the bound from the actual rounding-error norms (common.cuh: admission_margin_fp32)."""
    D, K, N = 64, 256, 8192
    f64 = lambda v: v.astype(np.float64)
    x_dn, x_up = np.float32(1 + 2.0 ** -8 - 2.0 ** -16), np.float32(1 + 2.0 ** -8 + 2.0 ** -16)
    z1 = np.zeros(D, np.float32); z1[:32] = x_dn; z1[32:63] = x_up; z1[63] = 1.0
    ea = np.zeros(D, np.float32); ea[:32] = x_dn; ea[63] = np.float32(-49 / 128)
    eb = np.zeros(D, np.float32); eb[32:63] = x_up
    s = lambda zz, e, e0: float(f64(zz) @ f64(e) - 0.5 * (f64(e0) ** 2).sum())
    gap_exact = s(z1, ea, ea) - s(z1, eb, eb)
    gap_approx = s(_bf16(z1), _bf16(eb), eb) - s(_bf16(z1), _bf16(ea), ea)
    nz, emax = np.linalg.norm(f64(z1)), max(np.linalg.norm(f64(ea)), np.linalg.norm(f64(eb)))
    assert 0.03 < gap_exact < 0.06 and gap_approx > 2 * 0.00391007 * 1.02 * nz * emax        # the old margin loses a
    rz = np.linalg.norm(f64(z1 - _bf16(z1)))
    re = max(np.linalg.norm(f64(ea - _bf16(ea))), np.linalg.norm(f64(eb - _bf16(eb))))
    emb = max(np.linalg.norm(f64(_bf16(ea))), np.linalg.norm(f64(_bf16(eb))))
    assert gap_approx < 2.0 * (rz * emb + nz * re)                                              # the new one keeps it

    rs = np.random.RandomState(5)
    E = (rs.standard_normal((K, D)) * 0.25).astype(np.float32)        # |e| ~ 2: never wins, never sets max|e|
    ia, ib = 77, 5                                                    # b has the LOWER index: a tie-break cannot save a
    E[ia], E[ib] = ea, eb
    z = rs.standard_normal((N, D)).astype(np.float32)
    hot = np.arange(0, N, 3)
    z[hot] = z1 * np.float32(2.0) ** rs.randint(-3, 4, size=(hot.size, 1)).astype(np.float32)   # exact rescalings
    # a wins exactly only at scale 1 (the score is not scale-invariant); keep the scale-1 rows as the probe
    probe = hot[np.all(z[hot] == z1, axis=1)]
    assert probe.size > 100
    old = os.environ.get("VQB200_NO_FUSED")
    os.environ["VQB200_NO_FUSED"] = "0" if fused else "1"
    try:
        assert bool(vq.ops.fused_supported(N, K, D, 0)) == fused
        path, idx, _ = run_search(vq, z.reshape(N // 64, 64, D), E, K)
    finally:
        if old is None:
            os.environ.pop("VQB200_NO_FUSED", None)
        else:
            os.environ["VQB200_NO_FUSED"] = old
    assert path == 1
    assert (idx[probe] == ia).all(), f"{(idx[probe] != ia).sum()} of {probe.size} adversarial rows lost the exact winner"
    ref = O.nearest_code64(z, E)
    mm, outside = O.near_tie_rows(z, E, idx, ref)
    assert outside.size == 0


@pytest.mark.parametrize("D,K", [(64, 256), (128, 384)])
@pytest.mark.parametrize("case", ["huge_rows", "tiny_rows", "huge_codes", "tiny_codes", "mixed_scales"])
def test_fp16_operand_range_is_safe(vq, D, K, case):
    """fp32 mode feeds the tensor core fp16 copies.  Values outside fp16's range must cost speed, never
    correctness: an overflow makes the measured conversion error infinite (exact path for the row / the level),
    an underflow is measured like any other rounding error.  Indices must equal the fp64 arbiter's."""
    N = 8192
    E, z = large_case_inputs(700 + D, K, D, 1, N // 64, 64)
    z = z.reshape(N, D).copy()
    if case == "huge_rows":
        z[::5] *= np.float32(3e5)                       # > 65504 per element: fp16 overflow on those rows
        z[1::5] *= np.float32(2.0 ** 14)
    elif case == "tiny_rows":
        z[::3] *= np.float32(1e-7)                      # every element flushes to zero
        z[1::3] *= np.float32(2.0 ** -12)               # partly subnormal in fp16
    elif case == "huge_codes":
        E = (E * np.float32(1e6)).astype(np.float32)    # the whole level overflows fp16
        z = (z * np.float32(1e6)).astype(np.float32)
    elif case == "tiny_codes":
        E = (E * np.float32(1e-6)).astype(np.float32)
        z = (z * np.float32(1e-6)).astype(np.float32)
    else:
        E[::7] *= np.float32(1e-4)
        E[3::11] *= np.float32(50.0)
        z[::2] *= np.float32(30.0)
    path, idx, _ = run_search(vq, z.reshape(N // 64, 64, D), E, K)
    assert path == 1
    ref = O.nearest_code64(z, E)
    mm, outside = O.near_tie_rows(z, E, idx, ref)
    assert outside.size == 0, f"{case}: {outside.size} rows differ from the fp64 arbiter outside near-ties"
    assert mm.size <= 8


@pytest.mark.parametrize("K,D,N,mask", [(1024, 256, 2 * (1 << 20) + 20480, False), (512, 384, (1 << 20) + 19000, True)])
def test_side_job_pipeline_matches_plain_chunk_pipeline(K, D, N, mask, monkeypatch):
    """Several chunks on the CTA-pair kernel: the tensor kernel of chunk i carries the pre-pass of chunk i+1 and the
    gather / straight-through / loss / histogram pass of chunk i-1 on six side warps (SideJobs, vq_search_tc.cu).
    Every output must equal the default pipeline's (separate pre-pass and gather kernels), and the
    indices the oracle's on a sample of rows."""
    import pytorch_vae_b200 as vq
    dev = torch.device("cuda:0")
    lib = vq._cabi.lib
    gen = torch.Generator(device=dev).manual_seed(77)
    E = torch.randn(K, D, device=dev, generator=gen) / np.sqrt(D)
    z = torch.randn(1, N, D, device=dev, generator=gen)
    m = (torch.rand(1, N, device=dev, generator=gen) > 0.3) if mask else None

    def run():
        q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
        q.embedding.copy_(E)
        with torch.no_grad():
            st, zq, idx, stats = q(z, do_ema_update=False, mask=m)
        torch.cuda.synchronize()
        return st, zq, idx, stats, q._ep_usage.clone(), q.last_commit.clone()
    monkeypatch.delenv("VQB200_TC2_SIDE", raising=False)
    assert lib.vqb200_search_path(N, K, D, 0) == 1          # default: the three-stream chunk pipeline
    plain = run()
    monkeypatch.setenv("VQB200_TC2_SIDE", "1")              # opt-in switch (measured slower on power-capped B200s)
    assert lib.vqb200_search_path(N, K, D, 0) == 2, "this shape must take the side-job pipeline"
    side = run()
    assert torch.equal(side[2], plain[2])
    assert torch.equal(side[1], plain[1]) and torch.equal(side[0], plain[0])
    assert torch.equal(side[4], plain[4])
    assert torch.allclose(side[3], plain[3], rtol=1e-6) and torch.allclose(side[5], plain[5], rtol=1e-6)
    assert torch.equal(side[1].view(-1, D), E[side[2].view(-1)])
    assert float(side[4].sum()) == (float(m.sum()) if mask else N)
    rows = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:2048].to(dev)
    zs, En = z.view(-1, D)[rows].cpu().numpy(), E.cpu().numpy()
    mm, outside = O.near_tie_rows(zs, En, side[2].view(-1)[rows].cpu().numpy(), O.nearest_code64(zs, En))
    assert outside.size == 0 and mm.size <= 2


def test_codebook_sharded_search_on_the_tensor_path_simulated_ranks():
    """North star: "a codebook-sharded min-loc reduction is used when K.D exceeds the per-SM staging budget".  Four
    simulated ranks each search their slice of 2048 codes with the tcgen05 kernels (exact winner per slice), pack
    40 bits of the winner's exact fp64 score with its id, and the MIN over the ranks must reproduce the replicated
    search -- including a twin pair that straddles two slices (the lower id wins)."""
    import pytorch_vae_b200 as vq
    from pytorch_vae_b200 import sharding as S
    dev = torch.device("cuda:0")
    K, D, N, world = 8192, 256, 20000, 4
    gen = torch.Generator(device=dev).manual_seed(11)
    E = torch.randn(K, D, device=dev, generator=gen) / np.sqrt(D)
    E[K // 2 + 7] = E[3]                                        # twins in slices 0 and 2
    E[K - 1] = E[K // 4 + 1]                                    # twins in slices 1 and 3
    z = torch.randn(N, D, device=dev, generator=gen)
    z[:50] = E[3] + 0.01 * torch.randn(50, D, device=dev, generator=gen)
    z[50:90] = E[K - 1] + 0.01 * torch.randn(40, D, device=dev, generator=gen)
    q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
    q.embedding.copy_(E)
    with torch.no_grad():
        full = q(z.view(1, N, D), do_ema_update=False)[2].view(-1)
    assert bool((full[:50] == 3).all()) and bool((full[50:90] == K // 4 + 1).all())
    assert vq._cabi.lib.vqb200_search_path(N, K // world, D, 0) >= 1      # the slices take the tensor path
    packs = []
    for r in range(world):
        S.sharded_search_slice(q, z, world, r, lambda p: (packs.append(p.clone()), p)[1])
    sign = -(2 ** 63)
    best = torch.stack([p ^ sign for p in packs]).min(0).values ^ sign
    out = torch.empty_like(full)
    vq.ops.minloc_unpack24(best, out)
    assert torch.equal(out, full)
    # world = 1 degenerates to the plain search
    assert torch.equal(S.sharded_search_slice(q, z, 1, 0, lambda p: p), full)


def test_codebook_sharded_search_two_gpus_nccl():
    """The same over NCCL on two real GPUs (torchrun, one process per GPU): sharded == replicated, twins across the
    shards resolve to the lower id.  Skipped on a single-GPU box."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(root, "tests", "multi_gpu_worker.py"), "sharded_search"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded_search ok" in r.stdout


def test_row_sharded_statistics_two_gpus_nccl():
    """Rows sharded over two GPUs with stats_sync: indices equal the replicated run's slice, the all-reduced statistics
    and the epoch accumulators (global position count included) equal the single-GPU ones.  Skipped on one GPU."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29632",
                        os.path.join(root, "tests", "multi_gpu_worker.py"), "row_sharded_stats"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "row_sharded_stats ok" in r.stdout


def test_allreduced_ema_training_two_gpus_nccl():
    """Training at the stage-2 shape on two GPUs with the EMA segment sums all-reduced ONCE per step (the persistent
    kernel between vqb200_rvq_train_begin and _finish): both ranks end every step with the codebook one GPU gets from
    the whole batch.  Skipped on one GPU."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633",
                        os.path.join(root, "tests", "multi_gpu_worker.py"), "allreduced_ema_training"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "allreduced_ema_training ok" in r.stdout
