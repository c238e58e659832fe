"""indices -> decoder memory (include/vq_b200.h: vqb200_indices_to_memory; SURVEY.md section 8f rank 1, the
"from_code + mem_ln after the gather" half).  Oracle: the reference's own chain
scripts/decode_with_vqvae.py:110-130 (ids -> sum of the levels' codes) -> models/vq_vae.py:749
(mem_ln(from_code(z_q))) evaluated with stock torch modules in fp64 and in fp32.  Tolerance: 2e-5 absolute on the
LayerNorm output (unit scale) -- fp32 summation order differs (K_total-row table vs a GEMM over D)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vq():
    import pytorch_vae_b200 as m
    return m


@pytest.mark.parametrize("K_per,D,L,H,n_tok,dtype", [
    (1024, 512, 4, 512, 8192, torch.int64),      # stage-2 shape
    (1024, 512, 4, 512, 77, torch.int32),
    (512, 64, 1, 256, 4096, torch.int16),
    (300, 128, 3, 1024, 513, torch.int64),
    (64, 32, 2, 36, 100, torch.int64),           # H not a multiple of 128
])
def test_indices_to_memory_matches_torch_chain(vq, K_per, D, L, H, n_tok, dtype):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5 + K_per + H)
    E = torch.randn(K_per * L, D, device=dev, generator=g) / np.sqrt(D)
    lin = torch.nn.Linear(D, H).to(dev)
    ln = torch.nn.LayerNorm(H).to(dev)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5, generator=g)
        ln.bias.uniform_(-0.5, 0.5, generator=g)
    local = torch.randint(0, K_per, (n_tok, L), device=dev, generator=g)
    ids = (local + torch.arange(L, device=dev) * K_per).to(dtype)              # token-major global ids
    P = (E.double() @ lin.weight.double().t()).float().contiguous()
    with torch.no_grad():
        got = vq.ops.indices_to_memory(ids, P, L, lin.bias, ln.weight, ln.bias, ln.eps)
        zq = vq.ops.indices_to_latent(ids, E, L)                                # the reference's latent (level-order sum)
        want32 = ln(lin(zq))
        want64 = torch.nn.functional.layer_norm(zq.double() @ lin.weight.double().t() + lin.bias.double(), (H,),
                                                ln.weight.double(), ln.bias.double(), ln.eps)
    assert got.shape == (n_tok, H)
    err = (got.double() - want64).abs().max().item()
    ref_err = (want32.double() - want64).abs().max().item()
    assert err < 2e-5, (err, ref_err)
    assert err < 4 * ref_err + 2e-6, (err, ref_err)       # as close to fp64 as the stock fp32 chain is


def test_mirror_vqvae_decode_indices_matches_decode(vq):
    """VQVAE.decode_indices(ids) against VQVAE.decode(indices_to_latent(ids)) on the stage-2 model shape (small)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    from pytorch_vae_b200.vqvae import VQVAE
    m = VQVAE(input_dim=6, hidden_dim=128, code_dim=128, codebook_size=256, residual_vq=True, num_quantizers=2,
                 latent_tokens=16, max_seq_len=32, num_layers=1, num_heads=4, tokenizer_layers=1, tokenizer_heads=4,
                 print_init=False).to(dev).eval()
    B, M, Q = 3, 16, 2
    ids = torch.stack([torch.randint(0, 256, (B, M), device=dev) + q * 256 for q in range(Q)], dim=-1)   # [B, M, Q]
    mask = torch.ones(B, 32, dtype=torch.bool, device=dev)
    with torch.no_grad():
        zq = vq.ops.indices_to_latent(ids.reshape(-1), m.quantizer.embedding, Q).view(B, M, -1)
        want = m.decode(zq, mask)
        got = m.decode_indices(ids, mask)
        mem = m.memory_from_indices(ids.reshape(B, M * Q))
    assert mem.shape == (B, M, 128)
    torch.testing.assert_close(mem, m.mem_ln(m.from_code(zq)), rtol=0, atol=2e-5)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    # the table follows the codebook: an in-place change of the embedding invalidates it
    with torch.no_grad():
        m.quantizer.embedding.mul_(0.5)
        mem2 = m.memory_from_indices(ids)
        zq2 = vq.ops.indices_to_latent(ids.reshape(-1), m.quantizer.embedding, Q).view(B, M, -1)
    torch.testing.assert_close(mem2, m.mem_ln(m.from_code(zq2)), rtol=0, atol=2e-5)
