"""Parity of the CUDA path (through the C ABI) against the golden vectors of the live reference
and against the numpy oracle.  Tolerances (BASELINE.json north_star): indices bit-exact except
documented near-ties (fp64 distance gap < 1e-6 relative), z_q / z_q_st bit-exact given the
indices, loss within 1e-5 relative (fp32), statistics 1e-6 relative."""
import hashlib

import numpy as np
import pytest
import torch
from conftest import gsub
from synth import large_case_inputs

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
STAT_RTOL = 1e-5      # perplexity is exp() of an fp32-vs-fp64 entropy: 1e-6 typical, 1e-5 bound


@pytest.fixture(scope="module")
def vq():
    import pytorch_vae_b200 as m
    return m


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def sha(a):
    if torch.is_tensor(a):
        a = a.detach().cpu().contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def mk(vq, dev, K_per, D, L, E, **kw):
    q = vq.VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, **kw).to(dev)
    q.embedding.copy_(torch.as_tensor(E).to(dev))
    return q


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def npy(t):
    return t.detach().cpu().numpy()


# ----------------------------------------------------------------------------- small, full tensors
def test_small_single_eval(vq, dev, golden):
    g = gsub(golden, "small_single")
    q = mk(vq, dev, 64, 32, 1, g["E"]).eval()
    st, zq, idx, stats = q(T(g["z"], dev), do_ema_update=False)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == g["idx"].shape
    assert stats.shape == (2,) and stats.dtype == torch.float32
    assert np.array_equal(npy(idx), g["idx"])
    assert np.array_equal(npy(zq), g["zq"])
    assert np.array_equal(npy(st), g["zq_st"])
    np.testing.assert_allclose(npy(stats), g["stats"], rtol=STAT_RTOL)
    assert np.array_equal(npy(q._ep_usage), g["ep_usage"])
    assert np.array_equal(npy(q._ep_cnt), g["ep_cnt"])
    np.testing.assert_allclose(float(q.last_commit), float(g["commit"]), rtol=LOSS_RTOL)
    # the reference's own loss expression on our outputs
    np.testing.assert_allclose(float(torch.nn.functional.mse_loss(zq, T(g["z"], dev))), float(g["commit"]),
                               rtol=LOSS_RTOL)


def test_small_single_mask(vq, dev, golden):
    g0, g = gsub(golden, "small_single"), gsub(golden, "small_single_mask")
    q = mk(vq, dev, 64, 32, 1, g0["E"]).eval()
    _, _, idx, stats = q(T(g0["z"], dev), do_ema_update=False, mask=T(g["mask"], dev))
    assert np.array_equal(npy(idx), g["idx"])
    np.testing.assert_allclose(npy(stats), g["stats"], rtol=STAT_RTOL)
    assert np.array_equal(npy(q._ep_usage), g["ep_usage"])
    assert np.array_equal(npy(q._ep_cnt), g["ep_cnt"])
    # all-false mask: histogram empty -> perplexity 0, dead ratio 1 (models/vq_vae.py:204,212-217)
    q2 = mk(vq, dev, 64, 32, 1, g0["E"]).eval()
    _, _, _, s2 = q2(T(g0["z"], dev), do_ema_update=False, mask=torch.zeros(4, 16, dtype=torch.bool, device=dev))
    assert float(s2[0]) == 0.0 and float(s2[1]) == 1.0


def test_small_single_train(vq, dev, golden):
    g = gsub(golden, "small_single_train")
    q = mk(vq, dev, 64, 32, 1, g["E"], decay=float(g["decay"])).train()
    for step in range(3):
        s = gsub(golden, f"small_single_train/step{step}")
        _, zq, idx, stats = q(T(s["z"], dev), do_ema_update=True)
        assert np.array_equal(npy(idx), s["idx"])
        if step == 0:
            assert np.array_equal(npy(zq), s["zq"])          # gathered from the PRE-update codebook
        np.testing.assert_allclose(npy(q.ema_cluster_size), s["ema_cluster_size"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(npy(q.ema_embedding), s["ema_embedding"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(q.embedding), s["embedding"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(npy(q._ep_usage), g["ep_usage"])
    e = gsub(golden, "small_single_train/epoch")
    es = q.get_epoch_stats()
    assert es["n_positions"] == int(e["n_positions"])
    np.testing.assert_allclose(es["perplexity"], float(e["perplexity"]), rtol=1e-5)
    np.testing.assert_allclose(es["dead_ratio"], float(e["dead_ratio"]), rtol=1e-6)
    q.reset_epoch_stats()
    assert q.get_epoch_stats()["n_positions"] == 0


def test_small_single_train_mask_and_frozen(vq, dev, golden):
    g0, gm = gsub(golden, "small_single"), gsub(golden, "small_single_mask")
    g = gsub(golden, "small_single_train_mask")
    q = mk(vq, dev, 64, 32, 1, g0["E"], decay=0.95).train()
    _, _, idx, stats = q(T(g0["z"], dev), do_ema_update=True, mask=T(gm["mask"], dev))
    assert np.array_equal(npy(idx), g["idx"])
    np.testing.assert_allclose(npy(q.ema_cluster_size), g["ema_cluster_size"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(npy(q.embedding), g["embedding"], rtol=1e-5, atol=1e-6)
    # do_ema_update=False in training mode and eval mode with do_ema_update=True: buffers untouched
    for train, flag in ((True, False), (False, True)):
        q2 = mk(vq, dev, 64, 32, 1, g0["E"])
        q2.train(train)
        q2(T(g0["z"], dev), do_ema_update=flag)
        assert np.array_equal(npy(q2.embedding), g0["E"])
        assert float(q2.ema_cluster_size.abs().sum()) == 0.0
    # all-false mask in training: the reference skips the EMA entirely
    q3 = mk(vq, dev, 64, 32, 1, g0["E"]).train()
    q3(T(g0["z"], dev), do_ema_update=True, mask=torch.zeros(4, 16, dtype=torch.bool, device=dev))
    assert np.array_equal(npy(q3.embedding), g0["E"])


def test_small_grad(vq, dev, golden):
    g0, g = gsub(golden, "small_single"), gsub(golden, "small_single_grad")
    q = mk(vq, dev, 64, 32, 1, g0["E"], beta=float(g["beta"])).eval()
    z = T(g0["z"], dev).requires_grad_(True)
    st, zq, idx, stats = q(z, do_ema_update=False)
    assert st.requires_grad and not zq.requires_grad
    loss = (T(g["w"], dev) * st).sum() + q.beta * q.commitment_loss(zq, z)
    loss.backward()
    np.testing.assert_allclose(npy(z.grad), g["grad"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-5)
    # the reference's own loss expression (torch mse on our outputs) gives the same gradient
    z2 = T(g0["z"], dev).requires_grad_(True)
    st2, zq2, _, _ = q(z2, do_ema_update=False)
    ((T(g["w"], dev) * st2).sum() + q.beta * torch.nn.functional.mse_loss(zq2.detach(), z2)).backward()
    np.testing.assert_allclose(npy(z2.grad), g["grad"], rtol=1e-5, atol=1e-7)
    # straight-through alone: identity
    z3 = T(g0["z"], dev).requires_grad_(True)
    q(z3, do_ema_update=False)[0].sum().backward()
    assert bool((z3.grad == 1).all())
    # commitment_loss on a foreign pair takes the stand-alone kernel
    z4 = T(g0["z"], dev).requires_grad_(True)
    c = q.commitment_loss(zq.clone(), z4)
    c.backward()
    np.testing.assert_allclose(float(c), float(gsub(golden, "small_single")["commit"]), rtol=LOSS_RTOL)
    ref = 2.0 * (g0["z"] - npy(zq)) / g0["z"].size
    np.testing.assert_allclose(npy(z4.grad), ref, rtol=1e-5, atol=1e-9)


def test_small_rvq(vq, dev, golden):
    g = gsub(golden, "small_rvq")
    K_per, L = int(g["K_per"]), int(g["L"])
    q = mk(vq, dev, K_per, 16, L, g["E"]).eval()
    st, zq, idx, stats = q(T(g["z"], dev), do_ema_update=False)
    assert idx.dim() == 1 and idx.numel() == g["idx"].size          # [L*B*M] level-major, global ids
    assert np.array_equal(npy(idx), g["idx"])
    assert np.array_equal(npy(zq), g["zq"])                          # level-order sum, bitwise
    assert np.array_equal(npy(st), g["zq_st"])
    np.testing.assert_allclose(npy(stats), g["stats"], rtol=STAT_RTOL)
    assert np.array_equal(npy(q._ep_cnt), g["ep_cnt"])
    np.testing.assert_allclose(float(q.last_commit), float(g["commit"]), rtol=LOSS_RTOL)
    # wire formats: level-major -> [B, M*Q] (narrowed) -> latent
    B, M = g["z"].shape[:2]
    for dt in (torch.int64, torch.int32, torch.int16):
        bf = vq.ops.relayout_indices(idx, L, B, M, dtype=dt)
        assert bf.dtype == dt and np.array_equal(npy(bf).astype(np.int64), O.rvq_indices_batch_first(g["idx"], B, L))
        lat = vq.ops.indices_to_latent(bf[1], q.embedding, L)
        assert np.array_equal(npy(lat), g["zq"][1])


def test_small_rvq_train(vq, dev, golden):
    g, gt = gsub(golden, "small_rvq"), gsub(golden, "small_rvq_train")
    K_per, L = int(g["K_per"]), int(g["L"])
    q = mk(vq, dev, K_per, 16, L, g["E"], decay=float(gt["decay"])).train()
    for step in range(3):
        s = gsub(golden, f"small_rvq_train/step{step}")
        st, zq, idx, stats = q(T(s["z"], dev), do_ema_update=True)
        assert np.array_equal(npy(idx), s["idx"])
        np.testing.assert_allclose(npy(zq), s["zq"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(q.ema_cluster_size), s["ema_cluster_size"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(npy(q.ema_embedding), s["ema_embedding"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(q.embedding), s["embedding"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(npy(q._ep_usage), gt["ep_usage"])
    gm = gsub(golden, "small_rvq_train_mask")
    q = mk(vq, dev, K_per, 16, L, g["E"], decay=0.9).train()
    _, _, idx, stats = q(T(g["z"], dev), do_ema_update=True, mask=T(gm["mask"], dev))
    assert np.array_equal(npy(idx), gm["idx"])
    np.testing.assert_allclose(npy(q.embedding), gm["embedding"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(npy(q._ep_usage), gm["ep_usage"])


def test_semantics(vq, dev, golden):
    d = gsub(golden, "sem_dup")
    q = mk(vq, dev, 16, 8, 1, d["E"]).eval()
    assert np.array_equal(npy(q(T(d["z"], dev), do_ema_update=False)[2]), d["idx"])      # lowest twin wins
    n = gsub(golden, "sem_nan_row")
    assert np.array_equal(npy(q(T(n["z"], dev), do_ema_update=False)[2]), n["idx"])      # NaN / inf rows
    c = gsub(golden, "sem_nan_code")
    q2 = mk(vq, dev, 16, 8, 1, c["E"]).eval()
    assert np.array_equal(npy(q2(T(d["z"], dev), do_ema_update=False)[2]), c["idx"])     # first NaN code wins
    k = gsub(golden, "sem_collapsed")
    q3 = mk(vq, dev, 16, 8, 1, k["E"]).eval()
    assert np.array_equal(npy(q3(T(d["z"], dev), do_ema_update=False)[2]), k["idx"])     # collapsed codebook


# ----------------------------------------------------------------------------- larger, digests
def check_indices(z, E, got, ref, L, N, max_ties=2):
    """Near-tie rule; for RVQ chain-aware (a flipped row is excluded at deeper levels)."""
    D = z.shape[-1]
    got = got.reshape(L, N)
    ref = ref.reshape(L, N).astype(np.int64)
    if L == 1:
        mm, outside = O.near_tie_rows(z.reshape(-1, D), E, got[0], ref[0])
        assert outside.size == 0, f"{outside.size} rows differ outside the 1e-6 near-tie allowance"
        assert mm.size <= max_ties, f"{mm.size} near-tie rows"
        return mm.size
    alive = np.ones(N, bool)
    flips = 0
    for lvl in range(L):
        bad = alive & (got[lvl] != ref[lvl])
        flips += int(bad.sum())
        alive &= ~bad
    assert flips <= max_ties * L, f"{flips} chain flips"
    return flips


LARGE = ["c2_like", "c2_clustered", "ragged_k", "one_row", "c3_like", "c3_clustered", "stage2_rvq", "d128_scaled"]


@pytest.mark.parametrize("name", LARGE)
def test_large_cases(vq, dev, golden, name):
    g = gsub(golden, name)
    K_per, D, L, B, M = (int(g[k]) for k in ("K_per", "D", "L", "B", "M"))
    scale = None if float(g["scale"]) < 0 else float(g["scale"])
    E, z = large_case_inputs(int(g["seed"]), K_per, D, L, B, M, scale, bool(int(g["clustered"])))
    assert sha(E) == str(g["sha_E"]) and sha(z) == str(g["sha_z"])
    q = mk(vq, dev, K_per, D, L, E).eval()
    zt = T(z, dev)
    st, zq, idx, stats = q(zt, do_ema_update=False)
    got = npy(idx)
    flips = check_indices(z, E, got, g["idx"], L, B * M)
    # z_q is bit-exact GIVEN the indices, whatever they are
    Et = q.embedding
    if L == 1:
        assert torch.equal(zq.view(-1, D), Et[idx.view(-1)])
    else:
        acc = Et[idx[: B * M]]
        for lvl in range(1, L):
            acc = acc + Et[idx[lvl * B * M:(lvl + 1) * B * M]]
        assert torch.equal(zq.view(-1, D), acc)
    assert torch.equal(st, zt + (zq - zt))
    if flips == 0:
        assert sha(zq) == str(g["sha_zq"]) and sha(st) == str(g["sha_zq_st"]) and \
            sha(got.astype(np.int64)) == str(g["sha_idx"])
        np.testing.assert_allclose(npy(stats), g["stats"], rtol=STAT_RTOL)
    np.testing.assert_allclose(float(q.last_commit), float(g["commit"]), rtol=LOSS_RTOL)


KATS = ["kat_512_64", "kat_8192_256", "kat_rvq_stage2", "kat_512_64_big"]


@pytest.mark.parametrize("name", KATS)
def test_survey_known_answers(vq, dev, golden, name):
    """SURVEY.md section 8c rows, regenerated with the survey's torch-RNG recipe."""
    g = gsub(golden, name)
    K_per, D, L, B, M = (int(g[k]) for k in ("K_per", "D", "L", "B", "M"))
    torch.manual_seed(int(g["seed"]))
    E = torch.randn(K_per * L, D) * (1.0 / np.sqrt(D))
    z = torch.randn(B, M, D)
    if sha(E) != str(g["sha_E"]) or sha(z) != str(g["sha_z"]):
        pytest.skip("torch CPU RNG stream differs from the build container's")
    q = mk(vq, dev, K_per, D, L, E).eval()
    st, zq, idx, stats = q(z.to(dev), do_ema_update=False)
    flips = check_indices(z.numpy(), E.numpy(), npy(idx), g["idx"], L, B * M)
    if flips == 0:
        assert sha(idx) == str(g["sha_idx"]) and int(idx.sum()) == int(g["idx_sum"])
        assert sha(zq) == str(g["sha_zq"]) and sha(st) == str(g["sha_zq_st"])
        np.testing.assert_allclose(npy(stats), g["stats"], rtol=STAT_RTOL)
    np.testing.assert_allclose(float(q.last_commit), float(g["commit"]), rtol=LOSS_RTOL)


# ----------------------------------------------------------------------------- bf16-input mode
@pytest.mark.parametrize("K,D,N", [(512, 64, 4096), (1000, 48, 777), (4096, 256, 2048)])
def test_bf16_input_mode(vq, dev, K, D, N):
    """Oracle for bf16-input mode = the fp32 algorithm fed bf16-rounded inputs."""
    E, z = large_case_inputs(500 + K, K, D, 1, 1, N)
    zb = torch.from_numpy(z).bfloat16().float().numpy()
    Eb = torch.from_numpy(E).bfloat16().float().numpy()
    ref = O.nearest_code64(zb.reshape(-1, D), Eb)
    q = mk(vq, dev, K, D, 1, E, search_mode="bf16_input").eval()
    st, zq, idx, stats = q(T(z, dev), do_ema_update=False)
    mm, outside = O.near_tie_rows(zb.reshape(-1, D), Eb, npy(idx).reshape(-1), ref)
    assert outside.size == 0 and mm.size <= 2
    assert torch.equal(zq.view(-1, D), q.embedding[idx.view(-1)])       # z_q always comes from the fp32 codebook
    ref_loss = float(((E[npy(idx).reshape(-1)] - z.reshape(-1, D)).astype(np.float64) ** 2).mean())
    np.testing.assert_allclose(float(q.last_commit), ref_loss, rtol=1e-2)   # bf16-input tolerance of the north star


# ----------------------------------------------------------------------------- boundary behaviour
def test_boundary_behaviour(vq, dev, golden):
    g = gsub(golden, "small_single")
    q = mk(vq, dev, 64, 32, 1, g["E"]).eval()
    z = T(g["z"], dev)
    with pytest.raises(ValueError):
        q(z.view(-1, 32))                                               # not 3-D
    with pytest.raises(RuntimeError):
        q(z.cpu())                                                      # no CPU fallback
    with pytest.raises(RuntimeError):
        q(z.bfloat16())                                                 # dtype mismatch, like matmul in the reference
    # non-contiguous z_e (the reference reshapes, models/vq_vae.py:179)
    zt = z.transpose(0, 1).contiguous().transpose(0, 1)
    assert not zt.is_contiguous()
    assert np.array_equal(npy(q(zt, do_ema_update=False)[2]), g["idx"])
    # the codebook is the source of truth: mutate it from outside, the cache must follow
    q.embedding.mul_(-1.0)
    idx_neg = npy(q(z, do_ema_update=False)[2])
    assert np.array_equal(idx_neg.reshape(-1), O.nearest_code(g["z"].reshape(-1, 32), -g["E"]))
    sd = {k: v.clone() for k, v in q.state_dict().items()}
    sd["embedding"] = T(g["E"], dev)
    q.load_state_dict(sd, strict=True)
    assert np.array_equal(npy(q(z, do_ema_update=False)[2]), g["idx"])
    # empty batch
    st, zq, idx, stats = q(torch.empty(0, 16, 32, device=dev), do_ema_update=False)
    assert st.shape == (0, 16, 32) and idx.shape == (0, 16) and float(stats[1]) == 1.0
    # dead-code re-init keeps the three buffers consistent (models/vq_vae.py:91-107)
    q.train()
    usage = torch.ones(64, device=dev)
    usage[[3, 17, 40]] = 0
    flat = z.reshape(-1, 32)
    q._maybe_reinit_dead_codes(flat, usage)
    for k in (3, 17, 40):
        assert bool((flat == q.embedding[k]).all(1).any())
        assert torch.equal(q.embedding[k], q.ema_embedding[k]) and float(q.ema_cluster_size[k]) == 1.0
    assert np.array_equal(npy(q.embedding[0]), g["E"][0])


def test_cabi_error_codes(vq, dev):
    import ctypes as C
    lib = vq._cabi.lib
    z = torch.zeros(8, 32, device=dev)
    E = torch.zeros(16, 32, device=dev)
    cache = vq.ops.CodebookCache(16, 32, 16, dev)
    idx = torch.zeros(8, dtype=torch.int64, device=dev)
    ws = torch.zeros(1024, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def search(zp, D=32, mode=0, wsb=1024, K=16):
        return lib.vqb200_search(zp, 8, D, E.data_ptr(), cache.E_bf16.data_ptr(), cache.ee_half.data_ptr(),
                                 cache.ee_half.data_ptr() + 64, cache.level_meta.data_ptr(), K, mode, 0,
                                 idx.data_ptr(), ws.data_ptr(), wsb, s)
    assert search(z.data_ptr()) == 0
    assert search(None) == -1                   # VQB200_EINVAL
    assert search(z.data_ptr(), D=30) == -2     # VQB200_ESHAPE
    assert search(z.data_ptr() + 4) == -3       # VQB200_EALIGN
    assert search(z.data_ptr(), wsb=0) == -4    # VQB200_EWORKSPACE
    assert search(z.data_ptr(), mode=7) == -1
    assert search(z.data_ptr(), K=0) == -1
    assert lib.vqb200_codebook_prepare(E.data_ptr(), 16, 32, 5, cache.E_bf16.data_ptr(), cache.ee_half.data_ptr(),
                                       cache.level_meta.data_ptr(), s) == -2
    assert b"aligned" in lib.vqb200_status_string(-3)
    torch.cuda.synchronize()


def test_packed_minloc(vq, dev):
    """Codebook-sharded search: per-shard packed (key, idx) + elementwise min == full search."""
    E, z = large_case_inputs(900, 1024, 64, 1, 1, 3000)
    q = mk(vq, dev, 1024, 64, 1, E).eval()
    zt = T(z.reshape(-1, 64), dev)
    full = q(zt.view(1, -1, 64), do_ema_update=False)[2].view(-1)
    cache = q._codebook_cache()
    parts = []
    for s0 in range(0, 1024, 256):
        p = torch.empty(zt.shape[0], dtype=torch.int64, device=dev)
        vq.ops.search_packed(zt, q.embedding[s0:s0 + 256], cache.ee_half[0, s0:s0 + 256], s0, p)
        parts.append(p ^ (-2 ** 63))            # unsigned order -> signed order for a MIN all-reduce
    best = torch.stack(parts).min(0).values ^ (-2 ** 63)
    out = torch.empty_like(full)
    vq.ops.minloc_unpack(best, out)
    assert torch.equal(out, full)


# ----------------------------------------------------------------------------- BASELINE sizes: properties
@pytest.mark.parametrize("K,D,N,mode", [(512, 64, 1 << 20, "fp32"), (8192, 256, 1 << 19, "fp32")])
def test_full_size_properties(vq, dev, K, D, N, mode):
    gen = torch.Generator(device=dev).manual_seed(1234)
    E = torch.randn(K, D, device=dev, generator=gen) / np.sqrt(D)
    z = torch.randn(N // 64, 64, D, device=dev, generator=gen)
    q = vq.VectorQuantizerEMA(K, D, print_init=False, search_mode=mode).to(dev).eval()
    q.embedding.copy_(E)
    st, zq, idx, stats = q(z, do_ema_update=False)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    assert torch.equal(zq.view(-1, D), E[idx.view(-1)])                 # gather exactness
    assert torch.equal(st, z + (zq - z))                                # straight-through value
    assert float(q._ep_usage.sum()) == N and float(q._ep_cnt) == N      # checksum of the histogram
    assert torch.equal(torch.bincount(idx.view(-1), minlength=K).float(), q._ep_usage)
    np.testing.assert_allclose(float(q.last_commit), float(torch.nn.functional.mse_loss(zq, z)), rtol=LOSS_RTOL)
    # idempotence: a code's nearest code is itself (random codebook: no duplicates)
    idx2 = q(zq, do_ema_update=False)[2]
    assert torch.equal(idx2, idx)
    # a seeded sample of rows against the oracle
    rows = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:4096]
    zs = z.view(-1, D)[rows.to(dev)].cpu().numpy()
    ref = O.nearest_code64(zs, E.cpu().numpy())
    mm, outside = O.near_tie_rows(zs, E.cpu().numpy(), npy(idx.view(-1)[rows.to(dev)]), ref)
    assert outside.size == 0 and mm.size <= 2


@pytest.mark.parametrize("K_per,D,L,N,lstride_pad", [(64, 48, 5, 333, 0), (100, 20, 8, 1, 7), (1024, 512, 4, 2048, 0),
                                                     (32, 64, 1, 100, 0), (16, 36, 3, 4097, 100)])
def test_rvq_finalize_matches_level_order_sum(vq, dev, K_per, D, L, N, lstride_pad):
    """vqb200_rvq_finalize: z_q = ((E[i0] + E[i1]) + ...) in level order, bit for bit what stack().sum(0) gives
    (models/vq_vae.py:261); z_q_st, squared error and histogram as the per-level path produces them.  Odd sizes:
    D/4 not a power of two, one row, a padded level stride."""
    rs = np.random.RandomState(K_per + N)
    E = rs.standard_normal((K_per * L, D)).astype(np.float32)
    z = rs.standard_normal((N, D)).astype(np.float32)
    stride = N + lstride_pad
    ids = np.full((L, stride), -7, dtype=np.int64)
    for l in range(L):
        ids[l, :N] = rs.randint(0, K_per, N) + l * K_per
    want = E[ids[0, :N]].copy()
    for l in range(1, L):
        want = (want + E[ids[l, :N]]).astype(np.float32)
    want_st = (z + (want - z).astype(np.float32)).astype(np.float32)
    zq = torch.empty(N, D, device=dev)
    st = torch.empty(N, D, device=dev)
    scratch = torch.zeros(2 + K_per * L, dtype=torch.int32, device=dev)
    idt = T(ids, dev)
    vq.ops.rvq_finalize(T(z, dev), idt.view(-1), stride, L, T(E, dev), zq_out=zq, zq_st_out=st,
                        sqerr_sum=scratch[:2].view(torch.float64), hist=scratch[2:])
    assert np.array_equal(npy(zq), want) and np.array_equal(npy(st), want_st)
    assert np.array_equal(npy(scratch[2:]), np.bincount(ids[:, :N].reshape(-1), minlength=K_per * L))
    np.testing.assert_allclose(float(scratch[:2].view(torch.float64)[0]), float(((want.astype(np.float64) - z) ** 2).sum()),
                               rtol=1e-5)
    # optional outputs may be absent
    vq.ops.rvq_finalize(T(z, dev), idt.view(-1), stride, L, T(E, dev), zq_out=zq)
    assert np.array_equal(npy(zq), want)


def test_packed_statistics_equal_direct(vq, dev):
    """The multi-GPU statistics path (pack -> all-reduce -> finalize on the pack) on one rank == the direct path."""
    K = 777
    rs = np.random.RandomState(4)
    hist = T(rs.randint(0, 50, K).astype(np.int32) * (rs.rand(K) > 0.3), dev).to(torch.int32)
    sq = torch.tensor([123.456, 0.0], dtype=torch.float64, device=dev)
    n_elems = 4096 * 64
    a, b = torch.empty(3, device=dev), torch.empty(3, device=dev)
    ua, ub = torch.zeros(K, device=dev), torch.zeros(K, device=dev)
    ca, cb = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    vq.ops.stats_finalize(hist, 4096.0, sq, 1.0 / n_elems, ua, ca, a)
    pack = torch.empty(K + 2, dtype=torch.float64, device=dev)
    vq.ops.stats_pack(hist, sq, n_elems, pack)
    assert float(pack[0]) == 123.456 and float(pack[1]) == n_elems and torch.equal(pack[2:].to(torch.int32), hist)
    vq.ops.stats_finalize_packed(pack, K, 1, 64, ub, cb, b)          # 4096 * 64 elements, D = 64 -> 4096 positions
    assert torch.equal(a, b) and torch.equal(ua, ub) and torch.equal(ca, cb)
    vq.ops.stats_finalize_packed(pack * 2, K, 1, 64, ub, cb, b)       # two identical ranks: same statistics
    np.testing.assert_allclose(npy(a), npy(b), rtol=1e-6)


@pytest.mark.parametrize("K_per,D,one_exchange", [(256, 128, True), (96, 64, False)])
def test_allreduced_ema_training_path_equals_local_on_one_rank(vq, dev, monkeypatch, K_per, D, one_exchange):
    """The multi-GPU training paths -- ONE exchange per step around the persistent kernel (vqb200_rvq_train_begin /
    _finish) where it takes the shape, else per level: one library call up to the exchange point, all-reduce, EMA
    finalize -- with the all-reduce stubbed to the identity must reproduce the single-call local path."""
    import torch.distributed as dist
    from synth import large_case_inputs
    E, z = large_case_inputs(91, K_per, D, 3, 16, 64)
    assert vq.ops.rvq_train_fused_supported(16 * 64, K_per, D, 3, 0) == one_exchange

    def run(allreduce):
        q = vq.VectorQuantizerEMA(K_per, D, num_quantizers=3, print_init=False, decay=0.9).to(dev).train()
        q.embedding.copy_(T(E, dev))
        if allreduce:
            q.ema_sync = "allreduce"
            monkeypatch.setattr(vq.sharding, "dist_ready", lambda: True)
            monkeypatch.setattr(dist, "all_reduce", lambda t, *a, **k: None)
            q.stats_sync = False
        outs = []
        for step in range(2):
            zz = T(z, dev) * (1.0 + 0.1 * step)
            outs.append([t.clone() for t in q(zz, do_ema_update=True)])
        monkeypatch.undo()
        return q, outs
    ql, ol = run(False)
    qa, oa = run(True)
    for a, b in zip(ol, oa):
        assert torch.equal(a[2], b[2])
        assert torch.allclose(a[1], b[1], rtol=1e-4, atol=1e-5) and torch.allclose(a[3], b[3], rtol=1e-5)
    assert torch.allclose(ql.ema_cluster_size, qa.ema_cluster_size, rtol=1e-5, atol=1e-7)
    assert torch.allclose(ql.embedding, qa.embedding, rtol=1e-3, atol=1e-5)
