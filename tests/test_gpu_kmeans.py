"""k-means codebook initialiser (search + scatter-add + mean on the library's kernels) against the numpy
restatement: same starting centroids, same iterations.  Assignments are exact; centroids differ only by the
fp32 atomics' summation order (1e-5 relative)."""
import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vq():
    import pytorch_vae_b200 as m
    return m


def planted(seed, K, D, per, spread=8.0, noise=0.05):
    rs = np.random.RandomState(seed)
    centers = (rs.standard_normal((K, D)) * spread).astype(np.float32)
    z = (np.repeat(centers, per, 0) + noise * rs.standard_normal((K * per, D))).astype(np.float32)
    rs.shuffle(z)
    init = (centers + 0.3 * rs.standard_normal((K, D))).astype(np.float32)
    return z, init, centers


@pytest.mark.parametrize("K,D,per", [(128, 64, 64), (256, 32, 40), (40, 16, 100)])
def test_kmeans_matches_oracle(vq, K, D, per):
    dev = torch.device("cuda:0")
    z, init, _ = planted(3 + K, K, D, per)
    E, idx, counts = vq.kmeans_fit(torch.from_numpy(z).to(dev), K, iters=4, init=torch.from_numpy(init),
                                   return_assignments=True)
    Eo, idxo, hist = O.kmeans_lloyd(z, init, 4)
    assert np.array_equal(idx.cpu().numpy(), idxo)
    np.testing.assert_allclose(E.cpu().numpy(), Eo, rtol=1e-5, atol=1e-5)
    assert int(counts.sum()) == z.shape[0] and np.array_equal(counts.cpu().numpy(), np.bincount(idxo, minlength=K))


def test_kmeans_random_start_lowers_inertia_and_keeps_empty_clusters(vq):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(9)
    N, K, D = 20000, 512, 64
    z = rs.standard_normal((N, D)).astype(np.float32)
    zt = torch.from_numpy(z).to(dev)

    def inertia(E):
        Eh = E.cpu().numpy()
        return float(((z - Eh[O.nearest_code64(z, Eh)]) ** 2).sum())

    e0 = vq.kmeans_fit(zt, K, iters=0, seed=4)
    e5 = vq.kmeans_fit(zt, K, iters=5, seed=4)
    assert torch.equal(e0, vq.kmeans_fit(zt, K, iters=0, seed=4))            # seeded draw is reproducible
    assert inertia(e5) < 0.9 * inertia(e0)
    init = e0.clone()
    init[7] = 1e4                                                            # nobody is assigned to it
    e1 = vq.kmeans_fit(zt, K, iters=2, init=init)
    assert torch.equal(e1[7], init[7])


def test_rvq_kmeans_matches_oracle_and_feeds_the_quantizer(vq):
    """Level 0 (planted clusters, starts near the centres) is well conditioned: exact comparison.  Deeper levels
    cluster noise, where Lloyd is chaotic in the last bits of the centroids: compare the inertia they reach."""
    dev = torch.device("cuda:0")
    K, D, L = 128, 32, 3
    z, init0, _ = planted(21, K, D, 48)
    rs = np.random.RandomState(8)
    inits = np.stack([init0] + [(0.05 * rs.standard_normal((K, D))).astype(np.float32) for _ in range(L - 1)])
    zt = torch.from_numpy(z).to(dev)
    cent = vq.rvq_kmeans_fit(zt, K, L, iters=3, inits=torch.from_numpy(inits))
    assert tuple(cent.shape) == (L, K, D)
    want = O.rvq_kmeans_lloyd(z, inits, 3)
    np.testing.assert_allclose(cent[0].cpu().numpy(), want[0], rtol=1e-5, atol=1e-5)

    def chain_inertia(levels):
        r = z.copy()
        out = []
        for E in levels:
            r = (r - E[O.nearest_code64(r, E)]).astype(np.float32)
            out.append(float((r.astype(np.float64) ** 2).sum()))
        return out
    got_i, want_i = chain_inertia(cent.cpu().numpy()), chain_inertia(want)
    assert got_i[0] == pytest.approx(want_i[0], rel=1e-5)
    assert all(b < a for a, b in zip(got_i, got_i[1:]))                      # every level removes energy
    np.testing.assert_allclose(got_i, want_i, rtol=0.02)
    # seeded row draws (no inits) give the same shapes and a working residual codebook
    cent2 = vq.rvq_kmeans_fit(zt, K, L, iters=3, seed=2)
    q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).eval()
    q.embedding.copy_(cent2.reshape(L * K, D))
    _, zq, _, _ = q(zt.view(-1, 64, D), do_ema_update=False)
    err_rvq = float(((zq.view(-1, D) - zt) ** 2).mean())
    q1 = vq.VectorQuantizerEMA(K, D, num_quantizers=1, print_init=False).to(dev).eval()
    q1.embedding.copy_(cent2[0])
    err_1 = float(((q1(zt.view(-1, 64, D), do_ema_update=False)[1].view(-1, D) - zt) ** 2).mean())
    assert err_rvq < err_1
