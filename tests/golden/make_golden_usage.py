#!/usr/bin/env python
"""Golden vectors for the usage-entropy regulariser, from the LIVE reference's own `loss_function`
(build container only: needs /root/reference).  The latents enter `loss_function` as a leaf, beta is 0 and the
reconstruction is detached, so `loss.backward()` leaves exactly d(usage_reg)/d(z_e) in their `.grad`.

    python tests/golden/make_golden_usage.py        ->  tests/golden/usage_golden.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from models.vq_vae import VQVAE  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "usage_golden.npz")


def run(tag, G, num_quantizers, codebook_size, code_dim, lam, seed, scale):
    torch.manual_seed(seed)
    m = VQVAE(input_dim=6, hidden_dim=32, num_layers=1, num_heads=2, max_seq_len=40, codebook_size=codebook_size,
              code_dim=code_dim, beta=0.0, use_vq=True, num_quantizers=num_quantizers, label_smoothing=0.0,
              ss_tv_lambda=0.0, xyz_align_alpha=0.0, usage_entropy_lambda=lam, latent_tokens=8, tokenizer_heads=2,
              tokenizer_layers=1, tokenizer_dropout=0.0, reinit_dead_codes=False, print_init=False, name="tiny-usage").eval()
    m.quantizer.beta = 0.0
    m.quantizer.embedding.mul_(scale)                      # sharper / flatter softmax
    rs = np.random.RandomState(seed + 1)
    B, L = 5, 20
    x = np.zeros((B, L, 6), dtype=np.float32)
    x[..., :3] = rs.standard_normal((B, L, 3))
    x[np.arange(B)[:, None], np.arange(L)[None, :], 3 + rs.randint(0, 3, (B, L))] = 1.0
    mask = torch.ones(B, L, dtype=torch.bool)
    with torch.no_grad():
        recons, target, vq_pack, _ = m(torch.from_numpy(x), mask)
    zq, ze, idx, ppl, dead = vq_pack
    ze = (ze.detach() * scale).clone().requires_grad_(True)
    out = m.loss_function(recons.detach(), target, (zq.detach(), ze, idx, ppl, dead), mask)
    base = m.loss_function(recons.detach(), target, (zq.detach(), ze.detach(), idx, ppl, dead), mask)["loss"]
    out["loss"].backward()
    G[f"{tag}/E"] = m.quantizer.embedding.detach().numpy().copy()
    G[f"{tag}/z_e"] = ze.detach().numpy().copy()
    G[f"{tag}/lambda"] = np.asarray(lam)
    G[f"{tag}/usage_reg"] = np.asarray(float(out["Usage_Reg"]))
    G[f"{tag}/grad"] = ze.grad.numpy().copy()
    assert abs(float(out["loss"]) - float(base)) < 1e-6    # the loss value does not depend on the leaf trick


def main():
    torch.set_num_threads(4)
    G = {}
    run("single", G, 1, 48, 16, 0.2, 5, 1.0)
    run("sharp", G, 1, 64, 32, 0.05, 6, 3.0)
    run("rvq", G, 2, 32, 16, 0.1, 7, 1.5)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")
    for k in ("single", "sharp", "rvq"):
        print(k, float(G[f"{k}/usage_reg"]), float(np.abs(G[f"{k}/grad"]).max()))


if __name__ == "__main__":
    main()
