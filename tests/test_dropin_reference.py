"""Drop-in proof: the REFERENCE's own code runs on the B200 quantizer.

``pytorch_vae_b200.install()`` rebinds ``models.vq_vae.VectorQuantizerEMA`` (looked up as a module global at
models/vq_vae.py:506), after which the reference's own ``VQVAE(**yaml['model_params'])`` from
``configs/stage2_vq.yaml`` constructs this package's quantizer, loads a checkpoint written by the unpatched
reference with ``strict=True``, and the reference's caller helpers --
``scripts/extract_code_indices.py::tokenize_and_quantize`` / ``_ensure_batch_first_2d`` (:250-322, :195-209) and
``scripts/decode_with_vqvae.py::indices_to_latent`` (:89-130) -- execute against it unchanged.

The reference is found at ``/root/reference`` (build container) or ``baseline/_ref`` (staged by
``baseline/stage_reference.py``, travels to the GPU box); with neither the tests skip.  The CPU half builds,
loads and checks the attribute surface; the ``gpu`` half runs forward + loss_function on the device and compares
with the reference quantizer fed the same encoder latents on the CPU.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(cand, "models", "vq_vae.py")):
            return cand
    return None


REF = _ref_root()
needs_ref = pytest.mark.skipif(REF is None, reason="reference sources neither at /root/reference nor staged in baseline/_ref")


@pytest.fixture(scope="module")
def ref():
    """The reference's modules, imported from its own tree: models.vq_vae and the two scripts."""
    added = [REF, os.path.join(REF, "scripts")]
    for p in added:
        sys.path.insert(0, p)
    for name in ("models", "models.vq_vae", "extract_code_indices", "decode_with_vqvae"):
        sys.modules.pop(name, None)
    mods = {"vq": importlib.import_module("models.vq_vae"),
            "extract": importlib.import_module("extract_code_indices"),
            "decode": importlib.import_module("decode_with_vqvae")}
    yield mods
    import pytorch_vae_b200
    pytorch_vae_b200.uninstall("models.vq_vae")
    for p in added:
        sys.path.remove(p)


def stage2_params(small: bool):
    import yaml
    with open(os.path.join(REF, "configs", "stage2_vq.yaml")) as f:
        mp = dict(yaml.safe_load(f)["model_params"])
    mp["print_init"] = False
    mp["beta"] = 0.0005            # epoch-0 schedule value; the YAML's own 0.0 makes VQ_Loss identically 0
    if small:                      # same quantizer (4 x 1024, D = 512), thinner transformer stacks: CPU seconds
        mp.update(num_layers=1, tokenizer_layers=1)
    return mp


def synth_batch(B, L, seed):
    rs = np.random.RandomState(seed)
    x = np.zeros((B, L, 6), dtype=np.float32)
    x[..., :3] = rs.standard_normal((B, L, 3))
    x[np.arange(B)[:, None], np.arange(L)[None, :], 3 + rs.randint(0, 3, (B, L))] = 1.0
    mask = np.ones((B, L), dtype=bool)
    if B > 1:
        mask[1, L - 7:] = False
    return torch.from_numpy(x), torch.from_numpy(mask)


def build_pair(ref, small=True):
    """(reference model with the reference quantizer, the SAME reference VQVAE class built after install() and
    loaded with the first one's checkpoint, strict=True)."""
    import pytorch_vae_b200 as b200
    b200.uninstall("models.vq_vae")
    torch.manual_seed(5)
    m_ref = ref["vq"].VQVAE(**stage2_params(small)).eval()
    assert type(m_ref.quantizer).__module__ == "models.vq_vae"
    sd = {k: v.clone() for k, v in m_ref.state_dict().items()}
    b200.install("models.vq_vae")
    m_new = ref["vq"].VQVAE(**stage2_params(small)).eval()
    assert isinstance(m_new.quantizer, b200.VectorQuantizerEMA)
    missing, unexpected = m_new.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m_ref, m_new


@needs_ref
def test_reference_vqvae_builds_and_loads_on_the_b200_quantizer(ref):
    m_ref, m_new = build_pair(ref)
    q0, q1 = m_ref.quantizer, m_new.quantizer
    assert list(m_ref.state_dict().keys()) == list(m_new.state_dict().keys())
    for name in ("embedding", "ema_cluster_size", "ema_embedding", "_ep_usage", "_ep_top1_sum", "_ep_top2_sum",
                 "_ep_cnt", "_ep_qe_sum", "_ep_qe_hist"):
        a, b = getattr(q0, name), getattr(q1, name)
        assert torch.is_tensor(b) and a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b), name
    for name in ("K", "K_per", "D", "num_quantizers", "beta", "decay", "eps", "reinit_dead_codes", "reinit_prob",
                 "dead_usage_threshold"):
        assert getattr(q0, name) == getattr(q1, name), name
    assert (q1.K, q1.K_per, q1.D, q1.num_quantizers) == (4096, 1024, 512, 4)
    for meth in ("_ema_update", "_maybe_reinit_dead_codes", "reset_epoch_stats", "get_epoch_stats",
                 "get_embedding_snapshot", "forward"):
        assert callable(getattr(q1, meth))
    assert q1.get_epoch_stats().keys() == q0.get_epoch_stats().keys()
    # the reference model's own beta property writes through to the quantizer (models/vq_vae.py:555-563)
    m_new.beta = 0.003
    assert q1.beta == 0.003
    # there is no CPU path: the reference's forward reaches the quantizer and it refuses loudly
    x, mask = synth_batch(2, 40, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m_new(x, mask)


@needs_ref
def test_reference_script_helpers_on_cpu(ref):
    """The index re-layout and the indices -> latent gather of the reference's scripts are pure torch on
    ``quantizer.embedding`` / ``num_quantizers``: they run on the patched model as is and agree with the oracle."""
    _, m_new = build_pair(ref)
    q = m_new.quantizer
    B, M, Q = 3, 64, q.num_quantizers
    rs = np.random.RandomState(2)
    idx = np.concatenate([rs.randint(0, q.K_per, B * M) + l * q.K_per for l in range(Q)]).astype(np.int64)
    mask = torch.ones(B, 50, dtype=torch.bool)
    bf = ref["extract"]._ensure_batch_first_2d(torch.from_numpy(idx), mask, num_quantizers=Q, latent_tokens=M)
    assert np.array_equal(bf.numpy(), O.rvq_indices_batch_first(idx, B, Q))
    lat = ref["decode"].indices_to_latent(m_new, bf[0].numpy())
    np.testing.assert_allclose(lat.numpy(), O.indices_to_latent(bf[0].numpy(), q.embedding.numpy(), Q), rtol=1e-6,
                               atol=1e-7)
    assert ref["decode"].get_codebook(m_new) is q.embedding and ref["decode"].get_num_quantizers(m_new) == Q


@needs_ref
@pytest.mark.gpu
def test_reference_forward_loss_and_callers_on_gpu(ref):
    import pytorch_vae_b200 as b200
    dev = torch.device("cuda:0")
    m_ref, m_new = build_pair(ref, small=False)              # the full stage-2 model (43.6 M parameters)
    m_new.to(dev)
    L = m_new.quantizer.num_quantizers
    x, mask = synth_batch(8, 96, 3)
    with torch.no_grad():
        recons, target, vq_pack, _ = m_new(x.to(dev), mask.to(dev))
        out = m_new.loss_function(recons, target, vq_pack, mask.to(dev))
    zq, ze, idx, ppl, dead = vq_pack
    assert idx.dtype == torch.int64 and idx.dim() == 1 and idx.numel() == L * 8 * 64      # level-major global ids
    # the reference quantizer on the CPU, fed the SAME encoder latents
    ze_c = ze.detach().cpu()
    with torch.no_grad():
        st_r, zq_r, idx_r, stats_r = m_ref.quantizer(ze_c, do_ema_update=False)
    a, b = idx.cpu().numpy().reshape(L, -1), idx_r.numpy().reshape(L, -1)
    alive = np.ones(a.shape[1], bool)
    residual = ze_c.reshape(-1, ze_c.shape[-1]).numpy()
    E = m_ref.quantizer.embedding.numpy()
    for lvl in range(L):                                     # chain-aware near-tie rule
        bad = alive & (a[lvl] != b[lvl])
        if bad.any():
            El = E[lvl * 1024:(lvl + 1) * 1024]
            mm, outside = O.near_tie_rows(residual[bad], El, a[lvl][bad] - lvl * 1024, b[lvl][bad] - lvl * 1024)
            assert outside.size == 0
        alive &= ~bad
        residual = residual - E[b[lvl]]
    assert alive.mean() > 0.99
    if alive.all():
        assert torch.equal(zq.cpu(), zq_r)
        np.testing.assert_allclose(float(ppl), float(stats_r[0]), rtol=1e-5)
        np.testing.assert_allclose(float(dead), float(stats_r[1]), rtol=1e-6)
    commit_r = torch.nn.functional.mse_loss(zq_r, ze_c)
    np.testing.assert_allclose(float(out["VQ_Loss"]), float(m_ref.quantizer.beta * commit_r), rtol=1e-5 if alive.all() else 1e-3)
    assert float(out["VQ_Loss"]) > 0 and torch.isfinite(out["loss"])

    # scripts/extract_code_indices.py::tokenize_and_quantize, unchanged, against the installed quantizer
    indices, lengths, z_e2 = ref["extract"].tokenize_and_quantize(m_new, x, mask)
    assert tuple(indices.shape) == (8, 64 * L) and indices.is_cuda
    assert np.array_equal(lengths, mask.sum(1).numpy())
    with torch.no_grad():
        idx_r2 = m_ref.quantizer(z_e2.detach().cpu(), do_ema_update=False)[2]
    want = O.rvq_indices_batch_first(idx_r2.numpy(), 8, L)
    assert (indices.cpu().numpy() == want).mean() > 0.99
    # scripts/decode_with_vqvae.py::indices_to_latent + the model's own decode
    lat = ref["decode"].indices_to_latent(m_new, indices[0].cpu().numpy())
    mine = b200.ops.indices_to_latent(indices[0].contiguous(), m_new.quantizer.embedding, L)
    np.testing.assert_allclose(lat[0].cpu().numpy(), mine.cpu().numpy(), rtol=1e-6, atol=1e-6)
    rec = m_new.decode(lat, mask=torch.ones(1, 96, dtype=torch.bool, device=dev))
    assert torch.isfinite(rec).all()

    # one training-mode forward + backward through the reference model: EMA update on, gradient reaches the encoder
    m_new.train()
    recons, target, vq_pack, _ = m_new(x.to(dev), mask.to(dev))
    out = m_new.loss_function(recons, target, vq_pack, mask.to(dev))
    out["loss"].backward()
    assert m_new.to_code.weight.grad is not None and torch.isfinite(m_new.to_code.weight.grad).all()
    assert float(m_new.quantizer.ema_cluster_size.sum()) > 0
    torch.cuda.synchronize()
