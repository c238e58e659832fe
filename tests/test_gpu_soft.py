"""Soft-VQ training branch: the online-softmax kernel and `soft_forward` against golden vectors of the live
reference's VQVAE.forward (tests/golden/soft_golden.npz) and against the float64 oracle.  Tolerance: the
logits are |z - e|^2 / tau evaluated in fp32 with a different summation order than ATen's, so the soft
mixture agrees to ~1e-5 relative at the reference's tau range (2.0 ... 0.3); indices and hard codes are exact."""
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "soft_golden.npz")


@pytest.fixture(scope="module")
def vq():
    import pytorch_vae_b200 as m
    return m


def test_soft_branch_matches_live_reference(vq):
    dev = torch.device("cuda:0")
    g = np.load(GOLD)
    E0 = g["step0/E_before"]
    q = vq.VectorQuantizerEMA(E0.shape[0], E0.shape[1], print_init=False, reinit_dead_codes=False).to(dev).train()
    q.embedding.copy_(torch.from_numpy(E0))
    for s in range(int(g["n_steps"])):
        np.testing.assert_allclose(q.embedding.cpu().numpy(), g[f"step{s}/E_before"], rtol=2e-5, atol=1e-6)
        q.decay = float(g[f"step{s}/decay"])
        z_e = torch.from_numpy(g[f"step{s}/z_e"]).to(dev)
        tau, alpha = float(g[f"step{s}/tau"]), float(g[f"step{s}/alpha"])
        E_pre = q.embedding.clone()
        z_soft, z_hard, idx, stats = q.soft_forward(z_e, tau, do_ema_update=True)
        assert np.array_equal(idx.cpu().numpy(), g[f"step{s}/idx"])
        assert torch.equal(z_hard, E_pre[idx])                        # gathered from the codebook BEFORE the update
        # (after step 0 our codebook carries the atomics' summation noise relative to the reference's GEMM)
        np.testing.assert_allclose(z_hard.cpu().numpy(), g[f"step{s}/zq_hard"], rtol=2e-5, atol=1e-6)
        z_mix = (1 - alpha) * z_soft + alpha * z_hard                  # the reference's own host expression (:851-852)
        z_dec = z_e + (z_mix - z_e)
        np.testing.assert_allclose(z_dec.cpu().numpy(), g[f"step{s}/z_dec"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(float(stats[0]), float(g[f"step{s}/ppl"]), rtol=1e-5)
        np.testing.assert_allclose(float(stats[1]), float(g[f"step{s}/dead"]), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(q.ema_cluster_size.cpu().numpy(), g[f"step{s}/ema_cluster_size"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(q.embedding.cpu().numpy(), g[f"step{s}/E_after"], rtol=2e-5, atol=1e-6)
    assert float(q._ep_cnt) == 0.0                                     # the soft branch never touches the epoch stats


@pytest.mark.parametrize("K,D,N,tau", [(512, 64, 4096, 2.0), (512, 64, 4096, 0.3), (100, 48, 333, 1.0),
                                        (1024, 512, 512, 0.7), (37, 20, 65, 5.0)])
def test_soft_assign_matches_oracle(vq, K, D, N, tau):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(K + N)
    E = (rs.standard_normal((K, D)) / np.sqrt(D)).astype(np.float32)
    z = (E[rs.randint(0, K, N)] + 0.3 / np.sqrt(D) * rs.standard_normal((N, D))).astype(np.float32)
    got = vq.ops.soft_assign(torch.from_numpy(z).to(dev), torch.from_numpy(E).to(dev), tau).cpu().numpy()
    want = O.soft_assign(z, E, tau)
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=3e-6)


def test_soft_assign_edges(vq):
    dev = torch.device("cuda:0")
    E = torch.randn(64, 32, device=dev)
    z = E[:8].clone()
    out = vq.ops.soft_assign(z, E, 1e-12)                              # tau clamps at 1e-8: a hard arg-min
    assert torch.allclose(out, E[:8], atol=1e-6)
    assert vq.ops.soft_assign(torch.empty(0, 32, device=dev), E, 1.0).shape == (0, 32)
    q = vq.VectorQuantizerEMA(64, 32, num_quantizers=2, print_init=False).to(dev)
    with pytest.raises(RuntimeError):
        q.soft_forward(torch.randn(2, 4, 32, device=dev), 1.0)
    with pytest.raises(Exception):
        vq.ops.soft_assign(torch.randn(4, 1024, device=dev), torch.randn(8, 1024, device=dev), 1.0)   # D > 512


USAGE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "usage_golden.npz")


@pytest.mark.parametrize("tag", ["single", "sharp", "rvq"])
def test_usage_entropy_matches_live_reference(vq, tag):
    """The usage-entropy regulariser (forward value and gradient to z_e) against the reference's own
    loss_function (tests/golden/usage_golden.npz) and the float64 oracle."""
    dev = torch.device("cuda:0")
    g = np.load(USAGE)
    E, z, lam = g[f"{tag}/E"], g[f"{tag}/z_e"], float(g[f"{tag}/lambda"])
    q = vq.VectorQuantizerEMA(E.shape[0], E.shape[1], print_init=False).to(dev)      # K_total codes, one flat table
    q.embedding.copy_(torch.from_numpy(E))
    ze = torch.from_numpy(z).to(dev).requires_grad_(True)
    p_code = q.usage_code_probs(ze)
    entropy = -(p_code * p_code.clamp_min(1e-12).log()).sum()                         # the reference's own expression
    reg = -lam * entropy
    reg.backward()
    np.testing.assert_allclose(float(reg.detach()), float(g[f"{tag}/usage_reg"]), rtol=1e-5)
    scale = np.abs(g[f"{tag}/grad"]).max()
    np.testing.assert_allclose(ze.grad.cpu().numpy(), g[f"{tag}/grad"], rtol=2e-3, atol=2e-4 * scale)
    oreg, op, ograd = O.usage_entropy(z, E, lam)
    np.testing.assert_allclose(p_code.detach().cpu().numpy(), op, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(ze.grad.cpu().numpy(), ograd, rtol=1e-3, atol=1e-4 * scale)


@pytest.mark.parametrize("K,D,N", [(512, 64, 2048), (100, 48, 333), (1024, 512, 256), (37, 20, 65)])
def test_usage_probs_matches_oracle(vq, K, D, N):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(K + N)
    E = (rs.standard_normal((K, D)) * (2.0 / np.sqrt(D))).astype(np.float32)
    z = (rs.standard_normal((N, D)) * 1.5).astype(np.float32)
    zt = torch.from_numpy(z).to(dev).requires_grad_(True)
    q = vq.VectorQuantizerEMA(K, D, print_init=False).to(dev)
    q.embedding.copy_(torch.from_numpy(E))
    p = q.usage_code_probs(zt.view(1, N, D))
    (-(p * p.clamp_min(1e-12).log()).sum() * -0.3).backward()
    oreg, op, ograd = O.usage_entropy(z, E, 0.3)
    np.testing.assert_allclose(p.detach().cpu().numpy(), op, rtol=2e-5, atol=1e-8)
    scale = np.abs(ograd).max()
    np.testing.assert_allclose(zt.grad.cpu().numpy(), ograd, rtol=1e-3, atol=1e-4 * scale)


@pytest.mark.parametrize("K,D,N,tau", [(512, 64, 4096, 0.5), (1000, 128, 1111, 2.0), (1024, 512, 700, 0.7), (37, 20, 65, 1.0),
                                        (130, 36, 129, 0.3)])
def test_tiled_softmax_paths_match_first_version(vq, K, D, N, tau):
    """The tiled two-sweep kernel (vqb200_softmax_rows, + a plain fp32 GEMM) against the first implementation (one warp per
    row, online softmax, nothing stored): soft assignment, usage probabilities and their gradient."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(K + N)
    E = torch.randn(K, D, device=dev, generator=g) / np.sqrt(D)
    z = E[torch.randint(0, K, (N,), device=dev, generator=g)] + 0.3 / np.sqrt(D) * torch.randn(N, D, device=dev, generator=g)
    a = vq.ops.soft_assign(z, E, tau)
    b = vq.ops.soft_assign_warp(z, E, tau)
    torch.testing.assert_close(a, b, rtol=3e-5, atol=3e-6)
    P = vq.ops.softmax_rows(z, E, 1.0)
    assert P.shape == (N, K)
    torch.testing.assert_close(P.sum(1), torch.ones(N, device=dev), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(P, torch.softmax((z.double() @ E.double().t()), dim=1).float(), rtol=2e-5, atol=1e-8)
    p1, Pk = vq.ops.usage_probs(z, E, keep_probs=True)
    p2, rs = vq.ops.usage_probs_warp(z, E)
    torch.testing.assert_close(p1, p2, rtol=2e-5, atol=1e-8)
    gp = torch.randn(K, device=dev, generator=g)
    g1 = vq.ops.usage_probs_backward_from_probs(Pk, E, gp)
    g2 = vq.ops.usage_probs_backward(z, E, rs, gp)
    torch.testing.assert_close(g1, g2, rtol=1e-3, atol=1e-4 * float(g2.abs().max()))
