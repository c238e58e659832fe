"""Training-mode parity ON THE TENSOR-CORE PATH (BASELINE configs[4] and the configs[1] codebook).

Every golden EMA case of vq_golden.npz has D = 32 or 16 and so takes the exact SIMT search; these tests run the
shapes the product sends through tcgen05 -- ``vqb200_rvq_train_forward`` at the stage-2 shape (4 x 1024 codes,
D = 512, 8192 rows per step), the fused one-kernel forward at K = 512 / D = 64 and the chunked tensor search at
the same shape -- for three steps from the reference's ZERO EMA buffers, so the codebook collapses to exact
duplicates after the first update (models/vq_vae.py:52-53,85-88: ~3000 of 4096 codes are the zero vector) and
the dead-code de-duplication of codebook_refresh_kernel is on the path.

Checked against (a) the live reference's outputs (tests/golden/train_golden.npz, make_golden_train.py) and
(b) the numpy oracle through the teacher-forced replay of oracle/replay.py.  Tolerances: indices equal to
the oracle's except near-ties (fp64 gap < 1e-6 relative, none outside), z_q bit-exact from the pre-update
codebook of the first step, loss 1e-5, ema_cluster_size 1e-6, ema_embedding / embedding 1e-5 relative + 2e-6 of the
buffer's largest magnitude (fp32 summation order of segment sums with up to 8192 terms).
"""
import os

import numpy as np
import pytest
import torch
from synth import large_case_inputs, train_step_inputs
from oracle.replay import chain_alive, replay_step

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz")


@pytest.fixture(scope="module")
def vq():
    import pytorch_vae_b200 as m
    return m


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def npy(t):
    return t.detach().cpu().numpy()


def run_case(vq, dev, name, expect_fused, graph=False):
    g = np.load(GOLD)
    seed, K_per, D, L, B, M, steps = (int(v) for v in g[f"{name}/meta"])
    N = B * M
    decay, beta = float(g[f"{name}/decay"]), 0.0005
    lib = vq._cabi.lib
    assert lib.vqb200_search_path(N, K_per, D, 0) == 1, "this shape must take the tcgen05 search"
    assert bool(lib.vqb200_quantize_fused_supported(N, K_per, D, 0)) == (expect_fused and L == 1)
    E, _ = large_case_inputs(seed, K_per, D, L, 1, 1)
    q = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, decay=decay, beta=beta).to(dev).train()
    q.embedding.copy_(torch.from_numpy(E).to(dev))
    oq = O.OracleQuantizer(K_per, D, num_quantizers=L, embedding=E, decay=decay, beta=beta)
    oq.training = True
    rows = g[f"{name}/rows"]
    rs = np.random.RandomState(seed + 99)
    stepper = None
    same_as_reference = True
    for s in range(steps):
        z = train_step_inputs(seed, s, B, M, D)
        g_st = rs.standard_normal(z.shape).astype(np.float32)
        zt = torch.from_numpy(z).to(dev)
        gt = torch.from_numpy(g_st).to(dev)
        if graph:
            if stepper is None:
                stepper = vq.GraphedTrainStep(q, zt)
            st, zq, idx, stats, commit, grad = stepper(zt, gt, beta)
        else:
            ze = zt.clone().requires_grad_(True)
            st, zq, idx, stats = q(ze, do_ema_update=True)
            commit = q.last_commit
            torch.autograd.backward([st, commit], [gt, torch.full((), beta, device=dev)])
            grad = ze.grad
        torch.cuda.synchronize()
        idx_n = npy(idx).reshape(-1)
        p = f"{name}/step{s}"

        # (b) every level's decision against the oracle's pick from identical inputs, then the reference's EMA
        rep = replay_step(oq, z, idx_n)
        assert rep["outside"] == [0] * L, f"step {s}: indices outside the near-tie allowance per level {rep['outside']}"
        assert sum(rep["mismatch"]) <= 4
        if s == 0:
            assert np.array_equal(npy(zq), rep["zq"]), "z_q must be gathered bit-exactly from the pre-update codebook"
            assert np.array_equal(npy(st), O.straight_through(z, rep["zq"]))
        else:
            np.testing.assert_allclose(npy(zq), rep["zq"], rtol=1e-5, atol=1e-6)
        if commit is not None:
            np.testing.assert_allclose(float(commit), float(rep["commit"]), rtol=1e-5)
        np.testing.assert_allclose(npy(grad), O.commit_backward(g_st, z, npy(zq), beta), rtol=1e-5, atol=1e-7)
        usage, ppl, dead = O.usage_stats(idx_n, K_per * L)
        np.testing.assert_allclose(npy(stats), [ppl, dead], rtol=1e-5)
        np.testing.assert_allclose(npy(q.ema_cluster_size), oq.ema_cluster_size, rtol=1e-6, atol=1e-7)
        # segment sums of up to 8192 rows (a collapsed level sends every row to ONE code): 1e-5 relative plus the
        # fp32 summation-order noise of such a sum, 2e-6 of the buffer's largest magnitude (observed: 1.1e-6 on 2 of
        # 2 M elements; the reference's own one-hot GEMM carries the same noise)
        for got, want in ((npy(q.ema_embedding), oq.ema_embedding), (npy(q.embedding), oq.embedding)):
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6 * max(1.0, float(np.abs(want).max())))
        # collapsed codes are EXACTLY zero on both sides (0 / (0 + eps))
        assert np.array_equal(np.abs(npy(q.embedding)).sum(1) == 0, np.abs(oq.embedding).sum(1) == 0)

        # (a) the live reference's own run
        ref = g[f"{p}/idx"].astype(np.int64).reshape(-1)
        first, frac = chain_alive(idx_n, ref, L)
        same_as_reference &= np.array_equal(idx_n, ref)
        if s == 0:
            assert sum(f.size for f in first) <= 2 and frac > 0.999
        if same_as_reference:
            np.testing.assert_allclose(npy(stats), g[f"{p}/stats"], rtol=1e-5)
            np.testing.assert_allclose(npy(q.ema_cluster_size), g[f"{p}/ema_cluster_size"], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(npy(q.embedding)[rows], g[f"{p}/embedding_rows"], rtol=1e-5, atol=1e-6)
            nrm = np.sqrt((npy(q.embedding).astype(np.float64) ** 2).sum(1))
            np.testing.assert_allclose(nrm, g[f"{p}/embedding_norm"], rtol=1e-5, atol=1e-6)
            assert int((np.abs(npy(q.embedding)).sum(1) == 0).sum()) == int(g[f"{p}/n_zero_codes"])
    return same_as_reference


def test_config5_training_steps_on_tensor_path(vq, dev):
    """BASELINE configs[4]: K_per = 1024, D = 512, L = 4, N = 8192 through vqb200_rvq_train_forward -- three launches:
    the decay-only updates every level's codes receive from the earlier levels (known before the forward), ONE persistent
    kernel for all levels that also reduces the EMA segment sums (csrc/vq_rvq_fused.cu), then every level's own update
    and its trailing decay-only steps."""
    assert vq._cabi.lib.vqb200_rvq_train_launches(8192, 1024, 512, 4, 0) == 3
    run_case(vq, dev, "c5_train", expect_fused=False)


def test_config5_training_steps_level_by_level(vq, dev, monkeypatch):
    """The same three steps with the persistent kernel switched off: search -> gather -> scatter-add -> refresh per level."""
    monkeypatch.setenv("VQB200_NO_RVQ_FUSED_TRAIN", "1")
    assert vq._cabi.lib.vqb200_rvq_train_launches(8192, 1024, 512, 4, 0) > 4
    run_case(vq, dev, "c5_train", expect_fused=False)


@pytest.mark.parametrize("K_per,D,L,N,mode", [
    (1024, 512, 4, 8192, "fp32"),
    (1024, 512, 4, 9000, "fp32"),          # ragged last tile of 64
    (300, 256, 3, 12000, "fp32"),          # 128-row tiles, N = 256 code tiles, codes per level off the tile
    (256, 128, 2, 700, "bf16_input"),
    (1024, 512, 8, 2048, "fp32"),
])
def test_persistent_training_kernel_matches_level_pipeline(vq, dev, monkeypatch, K_per, D, L, N, mode):
    """Three training steps (from a random codebook with the reference's zero EMA buffers, so levels collapse and the
    duplicate-zero-code rule is on the path) through the persistent kernel and through the level-by-level pipeline:
    same indices (both re-rank exactly; the segment sums differ in summation order, which may move a later step's
    near-ties), EMA buffers and codebook within the summation-order tolerance of the header."""
    lib = vq._cabi.lib
    E, z = large_case_inputs(77 + K_per + L, K_per, D, L, 3, N)     # three batches of N rows
    z = z.reshape(3, N, D)

    def run():
        q = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, search_mode=mode).to(dev).train()
        q.embedding.copy_(torch.from_numpy(E).to(dev))
        outs = []
        for step in range(3):
            zt = torch.from_numpy(z[step]).to(dev).view(1, N, D)
            st, zq, idx, stats = q(zt, do_ema_update=True)
            outs.append((npy(idx), npy(zq).reshape(N, D), npy(st).reshape(N, D), npy(stats), float(q.last_commit)))
        torch.cuda.synchronize()
        return outs, npy(q.embedding), npy(q.ema_embedding), npy(q.ema_cluster_size), npy(q._ep_usage)
    assert lib.vqb200_rvq_train_launches(N, K_per, D, L, 0) == 3
    a = run()
    monkeypatch.setenv("VQB200_NO_RVQ_FUSED_TRAIN", "1")
    assert lib.vqb200_rvq_train_launches(N, K_per, D, L, 0) > 3
    b = run()
    for step, (x, y) in enumerate(zip(a[0], b[0])):
        same = (x[0].reshape(L, N) == y[0].reshape(L, N)).all(0)
        print(f"persistent vs level pipeline, {K_per}x{L} D={D} N={N} step {step}: {int((~same).sum())} of {N} rows differ")
        assert same.mean() > (0.999 if step == 0 else 0.99), f"step {step}: {(~same).sum()} rows differ"
        if step == 0:
            assert np.array_equal(x[1][same], y[1][same]) and np.array_equal(x[2][same], y[2][same])
        np.testing.assert_allclose(x[4], y[4], rtol=1e-4)
    np.testing.assert_allclose(a[3], b[3], rtol=1e-5, atol=1e-6)
    for u, v in ((a[1], b[1]), (a[2], b[2])):
        np.testing.assert_allclose(u, v, rtol=1e-4, atol=2e-6 * max(1e-30, float(np.abs(v).max())) + 1e-7)


def test_config5_training_steps_as_cuda_graph(vq, dev):
    run_case(vq, dev, "c5_train", expect_fused=False, graph=True)


def test_c2_codebook_training_steps_fused_kernel(vq, dev):
    """K = 512, D = 64, N = 8192: the fused one-kernel forward + scatter-add + EMA finalize."""
    run_case(vq, dev, "c2_train", expect_fused=True)


def test_c2_codebook_training_steps_chunked_tensor_search(vq, dev, monkeypatch):
    """The same with the fused kernel switched off: pre-pass -> tcgen05 search -> re-rank -> gather."""
    monkeypatch.setenv("VQB200_NO_FUSED", "1")
    run_case(vq, dev, "c2_train", expect_fused=False)


def test_full_size_c3_every_row(vq, dev):
    """BASELINE configs[2] at its FULL size (K = 8192, D = 256, N = 2^22): every row's index is compared with an
    fp32 argmin of the reference's distance expression evaluated on the device in row chunks
    (models/vq_vae.py:183-188); the rows that differ go to the oracle's fp64 near-tie rule."""
    K, D, N = 8192, 256, 1 << 22
    gen = torch.Generator(device=dev).manual_seed(4321)
    E = torch.randn(K, D, device=dev, generator=gen) / np.sqrt(D)
    z = torch.randn(N // 64, 64, D, device=dev, generator=gen)
    q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev).eval()
    q.embedding.copy_(E)
    with torch.no_grad():
        st, zq, idx, stats = q(z, do_ema_update=False)
    flat = z.view(-1, D)
    idx = idx.view(-1)
    assert torch.equal(zq.view(-1, D), E[idx])
    assert torch.equal(st, z + (zq - z))
    assert float(q._ep_usage.sum()) == N
    ee = E.pow(2).sum(1, keepdim=True).t()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    bad = []
    try:
        for r0 in range(0, N, 1 << 16):
            zc = flat[r0:r0 + (1 << 16)]
            d = (zc.pow(2).sum(1, keepdim=True) - 2.0 * torch.matmul(zc, E.t())) + ee
            diff = (d.argmin(1) != idx[r0:r0 + (1 << 16)]).nonzero().view(-1) + r0
            bad.append(diff)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    bad = torch.cat(bad)
    assert bad.numel() < N * 2e-4, f"{bad.numel()} rows differ from the fp32 argmin"
    if bad.numel():
        zs, En = npy(flat[bad]), npy(E)
        ref = O.nearest_code64(zs, En)
        mm, outside = O.near_tie_rows(zs, En, npy(idx[bad]), ref)
        assert outside.size == 0, f"{outside.size} of {N} rows outside the near-tie allowance"


def test_full_size_c4_every_row_chain(vq, dev):
    """BASELINE configs[3] per-GPU share (stage-2 RVQ, 2^21 rows): level by level, from the residual the product's
    own earlier levels imply, every row's index against the device fp32 argmin; differing rows to the oracle."""
    K_per, D, L, N = 1024, 512, 4, 1 << 21
    E_np, _ = large_case_inputs(77, K_per, D, L, 1, 1)
    E = torch.from_numpy(E_np).to(dev)
    gen = torch.Generator(device=dev).manual_seed(78)
    z = torch.randn(N // 64, 64, D, device=dev, generator=gen)
    q = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False).to(dev).eval()
    q.embedding.copy_(E)
    with torch.no_grad():
        st, zq, idx, stats = q(z, do_ema_update=False)
    idx = idx.view(L, N)
    torch.backends.cuda.matmul.allow_tf32 = False
    residual = z.view(-1, D).clone()
    acc = None
    for lvl in range(L):
        El = E[lvl * K_per:(lvl + 1) * K_per]
        ee = El.pow(2).sum(1, keepdim=True).t()
        bad = []
        for r0 in range(0, N, 1 << 17):
            rc = residual[r0:r0 + (1 << 17)]
            d = (rc.pow(2).sum(1, keepdim=True) - 2.0 * torch.matmul(rc, El.t())) + ee
            bad.append((d.argmin(1) + lvl * K_per != idx[lvl, r0:r0 + (1 << 17)]).nonzero().view(-1) + r0)
        bad = torch.cat(bad)
        assert bad.numel() < N * 2e-4
        if bad.numel():
            rs_, En = npy(residual[bad]), npy(El)
            mm, outside = O.near_tie_rows(rs_, En, npy(idx[lvl, bad]) - lvl * K_per, O.nearest_code64(rs_, En))
            assert outside.size == 0, f"level {lvl}: {outside.size} rows outside the near-tie allowance"
        zq_l = E[idx[lvl]]
        acc = zq_l if acc is None else acc + zq_l
        residual = residual - zq_l
    assert torch.equal(zq.view(-1, D), acc)                       # level-order sum, bit for bit
    assert torch.equal(st.view(-1, D), z.view(-1, D) + (acc - z.view(-1, D)))
