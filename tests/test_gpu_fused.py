"""The fully fused small-D forward (one kernel) against the multi-kernel path (VQB200_NO_FUSED=1) and the
oracle: identical indices, bit-identical z_q / z_q_st, identical histogram, loss within 1e-6."""
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


def forward(vq, z, E, K, mode="fp32", fused=True, mask=None):
    dev = torch.device("cuda:0")
    D = z.shape[-1]
    q = vq.VectorQuantizerEMA(K, D, print_init=False, search_mode=mode).to(dev).eval()
    q.embedding.copy_(torch.from_numpy(E).to(dev))
    old = os.environ.get("VQB200_NO_FUSED")
    os.environ["VQB200_NO_FUSED"] = "0" if fused else "1"
    try:
        used = vq.ops.fused_supported(z.shape[0] * z.shape[1], K, D, 0)
        st, zq, idx, stats = q(torch.from_numpy(z).to(dev), do_ema_update=False,
                               mask=None if mask is None else torch.from_numpy(mask).to(dev))
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("VQB200_NO_FUSED", None)
        else:
            os.environ["VQB200_NO_FUSED"] = old
    return used, st, zq, idx, stats, q


@pytest.mark.parametrize("K,N,mode,D", [(512, 8192, "fp32", 64), (128, 4096, "fp32", 64), (1000, 300 * 64, "fp32", 64),
                                        (2048, 16384, "fp32", 64), (4096, 8192, "fp32", 64), (512, 1 << 18, "fp32", 64),
                                        (512, 8192, "bf16_input", 64), (512, 8192, "fp32", 128),
                                        (1000, 300 * 64, "fp32", 128), (512, 1 << 17, "fp32", 128),
                                        (2048, 8192, "bf16_input", 128)])
def test_fused_equals_multikernel(vq, K, N, mode, D):
    E, z = large_case_inputs(300 + K + N % 977, K, D, 1, N // 64, 64)
    uf, st_f, zq_f, idx_f, stats_f, qf = forward(vq, z, E, K, mode, fused=True)
    um, st_m, zq_m, idx_m, stats_m, qm = forward(vq, z, E, K, mode, fused=False)
    assert uf and not um
    assert torch.equal(idx_f, idx_m)
    assert torch.equal(zq_f, zq_m) and torch.equal(st_f, st_m)
    assert torch.equal(qf._ep_usage, qm._ep_usage) and float(qf._ep_usage.sum()) == N
    np.testing.assert_allclose(stats_f.cpu().numpy(), stats_m.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(float(qf.last_commit), float(qm.last_commit), rtol=1e-6)
    if mode == "fp32":
        ref = O.nearest_code64(z.reshape(-1, D), E)
        mm, outside = O.near_tie_rows(z.reshape(-1, D), E, idx_f.cpu().numpy().reshape(-1), ref)
        assert outside.size == 0 and mm.size <= 2


def test_fused_hard_rows_and_mask(vq):
    K, D, N = 256, 64, 8192
    E, z = large_case_inputs(41, K, D, 1, N // 64, 64)
    z = z.copy()
    z[0, 3, 5] = np.nan
    z[1, 7, 0] = np.inf
    z[2, 1, 2] = -np.inf
    z[3, 0] = 0.0
    z[5, :, :] *= 40.0                                     # large norms: wide margins, many survivors
    E[40:200] = 0.0                                        # collapsed codebook: candidate slots overflow
    E[210] = E[17]
    z[4, :] = E[17] * 1.01                                 # duplicate of a winning code: lowest index
    mask = np.random.RandomState(3).rand(N // 64, 64) > 0.3
    uf, st_f, zq_f, idx_f, stats_f, qf = forward(vq, z, E, K, fused=True, mask=mask)
    um, st_m, zq_m, idx_m, stats_m, qm = forward(vq, z, E, K, fused=False, mask=mask)
    assert uf and not um
    assert torch.equal(idx_f, idx_m)
    fin = torch.isfinite(st_m).all(-1) & torch.isfinite(st_f).all(-1)
    assert torch.equal(zq_f, zq_m) and torch.equal(st_f[fin], st_m[fin])
    assert torch.equal(qf._ep_usage, qm._ep_usage)
    assert (idx_f[4] == 17).all()
    flat = z.reshape(-1, D)
    ok = np.isfinite(flat).all(1)
    assert np.array_equal(idx_f.cpu().numpy().reshape(-1)[ok], O.nearest_code64(flat[ok], E))


def test_graphed_forward_matches_eager(vq):
    """One CUDA graph for the whole residual forward (launch-bound stage-2 shape) == eager results."""
    dev = torch.device("cuda:0")
    E, z = large_case_inputs(77, 256, 128, 3, 16, 64)
    q = vq.VectorQuantizerEMA(256, 128, num_quantizers=3, print_init=False).to(dev).eval()
    q.embedding.copy_(torch.from_numpy(E).to(dev))
    zt = torch.from_numpy(z).to(dev)
    with torch.no_grad():
        st, zq, idx, stats = [t.clone() for t in q(zt, do_ema_update=False)]
    g = vq.GraphedForward(q, zt)
    for _ in range(2):
        st2, zq2, idx2, stats2 = g(zt)
    assert torch.equal(idx, idx2) and torch.equal(zq, zq2) and torch.equal(st, st2)
    assert torch.allclose(stats, stats2)
    # external codebook write is picked up by the captured cache refresh
    q.embedding.mul_(-1.0)
    with torch.no_grad():
        idx_neg = q(zt, do_ema_update=False)[2].clone()
    assert torch.equal(g(zt)[2], idx_neg)
    z_other = torch.randn_like(zt)
    with torch.no_grad():
        ref = q(z_other, do_ema_update=False)[2].clone()
    assert torch.equal(g(z_other)[2], ref)


@pytest.mark.parametrize("K,N", [(512, 1 << 16), (1000, 300 * 64), (128, 4096)])
def test_fused_tile_heights_agree(vq, K, N):
    """BM = 128 (two CTAs per SM, default) and BM = 256 (one CTA per SM) are the same function."""
    D = 64
    E, z = large_case_inputs(900 + K, K, D, 1, N // 64, 64)
    outs = {}
    for bm in ("128", "256"):
        os.environ["VQB200_FUSED_BM"] = bm
        try:
            outs[bm] = forward(vq, z, E, K, "fp32", fused=True)
        finally:
            os.environ.pop("VQB200_FUSED_BM", None)
    a, b = outs["128"], outs["256"]
    assert a[0] and b[0]
    assert torch.equal(a[3], b[3]) and torch.equal(a[2], b[2]) and torch.equal(a[1], b[1])
    assert torch.equal(a[5]._ep_usage, b[5]._ep_usage)
    np.testing.assert_allclose(float(a[5].last_commit), float(b[5].last_commit), rtol=1e-6)


@pytest.mark.parametrize("K,D,L,N,chunk", [(512, 64, 1, 40000 // 64 * 64, 8192), (300, 128, 1, 6400, 4096),
                                           (256, 128, 3, 6400, 2048), (512, 64, 1, 4096, None)])
def test_forward_host_equals_forward(vq, K, D, L, N, chunk):
    """The host-buffer entry (chunked H2D / kernels / D2H pipeline) returns what forward() returns."""
    dev = torch.device("cuda:0")
    E, z = large_case_inputs(500 + K + D, K, D, L, N // 64, 64)
    q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).eval()
    q.embedding.copy_(torch.from_numpy(E).to(dev))
    zt = torch.from_numpy(z)
    with torch.no_grad():
        st, zq, idx, stats = q(zt.to(dev), do_ema_update=False)
    usage1 = q._ep_usage.clone()
    q.reset_epoch_stats()
    st_h, zq_h, idx_h, stats_h = q.forward_host(zt.pin_memory(), chunk_rows=chunk)
    assert not idx_h.is_cuda and idx_h.dtype == torch.int64 and idx_h.shape == idx.shape
    assert torch.equal(idx_h, idx.cpu())
    assert torch.equal(zq_h, zq) and torch.equal(st_h, st)
    assert torch.equal(q._ep_usage, usage1)
    np.testing.assert_allclose(stats_h.numpy(), stats.cpu().numpy(), rtol=1e-6)
    # token-major int32 output == the re-layout of scripts/extract_code_indices.py:195-209 on forward()'s ids
    _, _, tok_h, _ = q.forward_host(zt, chunk_rows=chunk, token_major=torch.int32)      # pageable input is pinned inside
    B, M = z.shape[0], z.shape[1]
    want = O.rvq_indices_batch_first(idx.cpu().numpy().reshape(-1), B, L) if L > 1 else idx.cpu().numpy().reshape(B, M)
    assert tok_h.dtype == torch.int32 and tuple(tok_h.shape) == (B, M * L)
    assert np.array_equal(tok_h.numpy().astype(np.int64), want)
    # codes only
    none_st, none_zq, idx_c, _ = q.forward_host(zt, chunk_rows=chunk, outputs="indices")
    assert torch.equal(idx_c, idx.cpu())
    if L == 1:
        assert none_st is None and none_zq is None


def test_graphed_train_step_matches_eager(vq):
    """Forward + EMA update + backward as one CUDA graph == the same steps run eagerly (buffers, grads, outputs)."""
    dev = torch.device("cuda:0")
    E, z = large_case_inputs(78, 256, 128, 3, 16, 64)

    def make():
        q = vq.VectorQuantizerEMA(256, 128, num_quantizers=3, print_init=False, decay=0.9).to(dev).train()
        q.embedding.copy_(torch.from_numpy(E).to(dev))
        return q
    qe, qg = make(), make()
    g = vq.GraphedTrainStep(qg, torch.from_numpy(z).to(dev))
    assert torch.equal(qg.embedding, qe.embedding) and float(qg.ema_cluster_size.abs().sum()) == 0.0   # capture left no trace
    rs = np.random.RandomState(5)
    for step in range(3):
        zs = torch.from_numpy((z + 0.1 * rs.standard_normal(z.shape)).astype(np.float32)).to(dev)
        gs = torch.from_numpy(rs.standard_normal(z.shape).astype(np.float32)).to(dev)
        ze = zs.clone().requires_grad_(True)
        st, zq, idx, stats = qe(ze, do_ema_update=True)
        torch.autograd.backward([st, qe.last_commit], [gs, torch.full((), 0.3, device=dev)])
        st2, zq2, idx2, stats2, commit2, grad2 = g(zs, gs, 0.3)
        # the EMA update after every level moves the codebook by the atomics' summation noise, so the two runs'
        # codebooks (and with them z_q) agree to ~1e-6, not bit for bit
        assert torch.equal(idx, idx2)
        assert torch.allclose(zq, zq2, rtol=1e-4, atol=1e-5) and torch.allclose(st.detach(), st2.detach(), rtol=1e-4, atol=1e-5)
        assert torch.allclose(stats, stats2) and torch.allclose(qe.last_commit, commit2)
        assert torch.allclose(ze.grad, grad2, rtol=1e-6, atol=1e-8)
        # atomics order differs run to run: EMA buffers agree to summation noise
        assert torch.allclose(qe.ema_embedding, qg.ema_embedding, rtol=1e-4, atol=1e-5)
        assert torch.allclose(qe.embedding, qg.embedding, rtol=1e-3, atol=1e-5)


def test_forward_host_empty_and_ragged(vq):
    dev = torch.device("cuda:0")
    q = vq.VectorQuantizerEMA(512, 64, print_init=False).to(dev).eval()
    st, zq, idx, stats = q.forward_host(torch.empty(0, 64, 64))
    assert idx.numel() == 0 and zq.shape == (0, 64, 64)
    z = torch.randn(70, 64, 64)                                   # 4480 rows: one full chunk of 4096 + a ragged tail
    with torch.no_grad():
        ref = q(z.to(dev), do_ema_update=False)
    out = q.forward_host(z, chunk_rows=4096)
    assert torch.equal(out[2], ref[2].cpu()) and torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0])
