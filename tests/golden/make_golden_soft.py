#!/usr/bin/env python
"""Golden vectors for the soft-VQ training branch, from the LIVE reference (run in the build container only:
needs /root/reference).  Runs the reference's own VQVAE.forward in training mode with soft_vq_use=True on a tiny
configuration for a few steps and records, per step: the codebook before the step, the encoder latents
z_e_tokens (output of `to_code`), what the decoder received (`from_code`'s input = z_for_decode), the hard
indices / perplexity / dead ratio of vq_pack, the schedule values tau and alpha, and the EMA buffers after it.

    python tests/golden/make_golden_soft.py        ->  tests/golden/soft_golden.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from models.vq_vae import VQVAE  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "soft_golden.npz")
CFG = dict(input_dim=6, hidden_dim=32, num_layers=1, num_heads=2, max_seq_len=40, codebook_size=48, code_dim=16,
           beta=0.25, use_vq=True, num_quantizers=1, label_smoothing=0.0, ss_tv_lambda=0.0, xyz_align_alpha=0.0,
           latent_tokens=8, tokenizer_heads=2, tokenizer_layers=1, tokenizer_dropout=0.0, reinit_dead_codes=False,
           soft_vq_use=True, soft_vq_tau_start=2.0, soft_vq_tau_end=0.3, soft_vq_tau_warm_steps=6,
           soft_vq_alpha_warm_steps=4, print_init=False, name="tiny-soft")


def main():
    torch.manual_seed(91)
    torch.set_num_threads(4)
    m = VQVAE(**CFG).train()
    for mod in m.modules():                                  # deterministic forward: no dropout anywhere
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, (torch.nn.TransformerEncoderLayer, torch.nn.TransformerDecoderLayer)):
            for name in ("dropout", "dropout1", "dropout2", "dropout3"):
                if hasattr(mod, name):
                    getattr(mod, name).p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    grab = {}
    m.to_code.register_forward_hook(lambda mod, inp, out: grab.__setitem__("z_e", out.detach().clone()))
    m.from_code.register_forward_pre_hook(lambda mod, inp: grab.__setitem__("z_dec", inp[0].detach().clone()))
    rs = np.random.RandomState(92)
    G = {"n_steps": np.asarray(4)}
    for step in range(4):
        B, L = 4, 20
        x = np.zeros((B, L, 6), dtype=np.float32)
        x[..., :3] = rs.standard_normal((B, L, 3))
        x[np.arange(B)[:, None], np.arange(L)[None, :], 3 + rs.randint(0, 3, (B, L))] = 1.0
        mask = np.ones((B, L), dtype=bool)
        mask[1, 13:] = False
        G[f"step{step}/E_before"] = m.quantizer.embedding.detach().numpy().copy()
        recons, target, vq_pack, _ = m(torch.from_numpy(x), torch.from_numpy(mask))
        zq, ze, idx, ppl, dead = vq_pack
        ts = m.training_steps
        tau = m._interp_linear(m.soft_vq_tau_start, m.soft_vq_tau_end, ts, m.soft_vq_tau_warm_steps)
        alpha = m._linear_schedule(1.0, m.soft_vq_alpha_warm_steps)
        assert torch.equal(ze.detach(), grab["z_e"])
        G[f"step{step}/z_e"] = grab["z_e"].numpy()
        G[f"step{step}/z_dec"] = grab["z_dec"].numpy()
        G[f"step{step}/zq_hard"] = zq.detach().numpy()
        G[f"step{step}/idx"] = idx.detach().numpy()
        G[f"step{step}/ppl"] = np.asarray(float(ppl))
        G[f"step{step}/dead"] = np.asarray(float(dead))
        G[f"step{step}/tau"] = np.asarray(float(tau))
        G[f"step{step}/alpha"] = np.asarray(float(alpha))
        G[f"step{step}/E_after"] = m.quantizer.embedding.detach().numpy().copy()
        G[f"step{step}/ema_cluster_size"] = m.quantizer.ema_cluster_size.detach().numpy().copy()
        G[f"step{step}/decay"] = np.asarray(float(m.quantizer.decay))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
