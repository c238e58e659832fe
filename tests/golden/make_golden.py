"""Generate tests/golden/vq_golden.npz from the LIVE reference class.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as a
fixture.  Inputs come from ``np.random.RandomState`` (frozen stream) or, for the
SURVEY.md section 8c known-answer rows, from ``torch.manual_seed`` exactly as the
survey recipe states.  Small cases store full tensors; large cases store the
indices (narrowed), scalars and sha256 digests.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from models.vq_vae import VQVAE, VectorQuantizerEMA  # noqa: E402  (the oracle's authority)
from synth import large_case_inputs, rs_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vq_golden.npz")
G = {}


def sha(t) -> str:
    a = t.detach().cpu().contiguous().numpy() if torch.is_tensor(t) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def make_q(K_per, D, L, E, **kw):
    q = VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, **kw)
    q.embedding.copy_(torch.from_numpy(E))
    return q


def narrow(idx):
    a = idx.detach().cpu().numpy()
    return a.astype(np.int16) if a.max() < 32768 else a.astype(np.int32)


def put(prefix, **kv):
    for k, v in kv.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        G[f"{prefix}/{k}"] = np.array(v, copy=True)     # .numpy() aliases buffers that are updated in place


def d64_idx(z, E):
    z64 = torch.from_numpy(z.reshape(-1, z.shape[-1])).double()
    E64 = torch.from_numpy(E).double()
    d = z64.pow(2).sum(1, keepdim=True) - 2 * z64 @ E64.t() + E64.pow(2).sum(1)[None]
    return d.argmin(1)


# --------------------------------------------------------------------------- #
# small full-tensor cases                                                     #
# --------------------------------------------------------------------------- #
def case_small_single():
    K, D, B, M = 64, 32, 4, 16
    E, z = rs_inputs(11, K, D, B, M)
    q = make_q(K, D, 1, E).eval()
    zt = torch.from_numpy(z)
    st, zq, idx, stats = q(zt, do_ema_update=False)
    put("small_single", E=E, z=z, zq_st=st, zq=zq, idx=idx, stats=stats,
        ep_usage=q._ep_usage, ep_cnt=q._ep_cnt,
        commit=torch.nn.functional.mse_loss(zq, zt))
    # masked eval forward (mask restricts the histogram only)
    mask = torch.from_numpy(np.random.RandomState(12).rand(B, M) > 0.4)
    q2 = make_q(K, D, 1, E).eval()
    st, zq, idx, stats = q2(zt, do_ema_update=False, mask=mask)
    put("small_single_mask", mask=mask, idx=idx, stats=stats, ep_usage=q2._ep_usage, ep_cnt=q2._ep_cnt)
    # training: 3 EMA steps, decay 0.9, inputs drift
    q3 = make_q(K, D, 1, E, decay=0.9).train()
    rs = np.random.RandomState(13)
    for step in range(3):
        zs = rs.standard_normal((B, M, D)).astype(np.float32)
        st, zq, idx, stats = q3(torch.from_numpy(zs), do_ema_update=True)
        put(f"small_single_train/step{step}", z=zs, idx=idx, zq=zq, stats=stats,
            embedding=q3.embedding, ema_cluster_size=q3.ema_cluster_size, ema_embedding=q3.ema_embedding)
    put("small_single_train", E=E, decay=0.9, ep_usage=q3._ep_usage, ep_cnt=q3._ep_cnt)
    es = q3.get_epoch_stats()
    put("small_single_train/epoch", perplexity=es["perplexity"], dead_ratio=es["dead_ratio"],
        n_positions=es["n_positions"])
    # training with a mask: EMA sees valid rows only
    q4 = make_q(K, D, 1, E, decay=0.95).train()
    st, zq, idx, stats = q4(zt, do_ema_update=True, mask=mask)
    put("small_single_train_mask", idx=idx, stats=stats, embedding=q4.embedding,
        ema_cluster_size=q4.ema_cluster_size, ema_embedding=q4.ema_embedding)
    # training but do_ema_update=False: buffers untouched
    q5 = make_q(K, D, 1, E).train()
    q5(zt, do_ema_update=False)
    assert torch.equal(q5.embedding, torch.from_numpy(E))
    # autograd: d/dz [ sum(w * z_q_st) + beta * mse(z_q.detach(), z) ]
    w = torch.from_numpy(np.random.RandomState(14).standard_normal((B, M, D)).astype(np.float32))
    q6 = make_q(K, D, 1, E, beta=0.25).eval()
    zg = zt.clone().requires_grad_(True)
    st, zq, idx, stats = q6(zg, do_ema_update=False)
    loss = (w * st).sum() + q6.beta * torch.nn.functional.mse_loss(zq.detach(), zg)
    loss.backward()
    put("small_single_grad", w=w, beta=0.25, grad=zg.grad, loss=loss)


def case_small_rvq():
    K_per, D, L, B, M = 32, 16, 3, 2, 8
    E, z = rs_inputs(21, K_per * L, D, B, M)
    E[K_per:] *= 0.5          # deeper levels quantise smaller residuals
    q = make_q(K_per, D, L, E).eval()
    zt = torch.from_numpy(z)
    st, zq, idx, stats = q(zt, do_ema_update=False)
    put("small_rvq", E=E, z=z, zq_st=st, zq=zq, idx=idx, stats=stats, ep_usage=q._ep_usage,
        ep_cnt=q._ep_cnt, commit=torch.nn.functional.mse_loss(zq, zt), K_per=K_per, L=L)
    q2 = make_q(K_per, D, L, E, decay=0.9).train()
    rs = np.random.RandomState(22)
    for step in range(3):
        zs = rs.standard_normal((B, M, D)).astype(np.float32)
        st, zq, idx, stats = q2(torch.from_numpy(zs), do_ema_update=True)
        put(f"small_rvq_train/step{step}", z=zs, idx=idx, zq=zq, zq_st=st, stats=stats,
            embedding=q2.embedding, ema_cluster_size=q2.ema_cluster_size, ema_embedding=q2.ema_embedding)
    put("small_rvq_train", decay=0.9, ep_usage=q2._ep_usage, ep_cnt=q2._ep_cnt)
    mask = torch.from_numpy(np.random.RandomState(23).rand(B, M) > 0.3)
    q3 = make_q(K_per, D, L, E, decay=0.9).train()
    st, zq, idx, stats = q3(zt, do_ema_update=True, mask=mask)
    put("small_rvq_train_mask", mask=mask, idx=idx, stats=stats, embedding=q3.embedding,
        ema_cluster_size=q3.ema_cluster_size, ema_embedding=q3.ema_embedding, ep_usage=q3._ep_usage)


def case_semantics():
    # exact duplicates -> lowest index; NaN row -> torch.argmin's answer; NaN code -> wins everywhere
    K, D = 16, 8
    E, z = rs_inputs(31, K, D, 1, 12)
    E[9] = E[3]
    E[12] = E[3]
    z[0, 0] = E[3] + 1e-3
    z[0, 1] = E[9]
    q = make_q(K, D, 1, E).eval()
    idx = q(torch.from_numpy(z), do_ema_update=False)[2]
    put("sem_dup", E=E, z=z, idx=idx)
    z2 = z.copy()
    z2[0, 4, 2] = np.nan
    z2[0, 7, 0] = np.inf
    z2[0, 8, 1] = -np.inf
    idx = q(torch.from_numpy(z2), do_ema_update=False)[2]
    put("sem_nan_row", z=z2, idx=idx)
    E3 = E.copy()
    E3[5, 1] = np.nan
    E3[11, 0] = np.nan
    q3 = make_q(K, D, 1, E3).eval()
    idx = q3(torch.from_numpy(z), do_ema_update=False)[2]
    put("sem_nan_code", E=E3, idx=idx)
    # collapsed codebook: 90% of codes are the zero vector
    Ec = E.copy()
    Ec[2:] = 0.0
    qc = make_q(K, D, 1, Ec).eval()
    idx = qc(torch.from_numpy(z), do_ema_update=False)[2]
    put("sem_collapsed", E=Ec, idx=idx)


# --------------------------------------------------------------------------- #
# larger cases: indices + digests                                             #
# --------------------------------------------------------------------------- #
def case_large(name, seed, K_per, D, L, B, M, scale=None, clustered=False):
    E, z = large_case_inputs(seed, K_per, D, L, B, M, scale, clustered)
    q = make_q(K_per, D, L, E).eval()
    zt = torch.from_numpy(z)
    with torch.no_grad():
        st, zq, idx, stats = q(zt, do_ema_update=False)
    extra = {}
    if L == 1:
        i64 = d64_idx(z, E)
        extra["n_fp64_mismatch"] = int((i64 != idx.view(-1)).sum())
    put(name, seed=seed, K_per=K_per, D=D, L=L, B=B, M=M, clustered=int(clustered),
        scale=-1.0 if scale is None else scale,
        idx=narrow(idx), stats=stats, commit=torch.nn.functional.mse_loss(zq, zt),
        sha_E=sha(E), sha_z=sha(z), sha_zq=sha(zq), sha_zq_st=sha(st), sha_idx=sha(idx), **extra)


def case_survey_kat(name, seed, K_per, D, L, B, M):
    """SURVEY.md section 8c recipe, verbatim (torch RNG)."""
    torch.manual_seed(seed)
    q = VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False).eval()
    z = torch.randn(B, M, D)
    with torch.no_grad():
        st, zq, idx, stats = q(z, do_ema_update=False)
    put(name, seed=seed, K_per=K_per, D=D, L=L, B=B, M=M, idx=narrow(idx), stats=stats,
        commit=torch.nn.functional.mse_loss(zq, z), idx_sum=int(idx.sum()),
        sha_E=sha(q.embedding), sha_z=sha(z), sha_idx=sha(idx), sha_zq=sha(zq), sha_zq_st=sha(st))
    print(name, sha(q.embedding), sha(z), sha(idx), int(idx.sum()), float(stats[0]), float(stats[1]))


VQVAE_CFG = dict(input_dim=6, hidden_dim=32, num_layers=1, num_heads=2, max_seq_len=40, codebook_size=32, code_dim=16,
                 beta=0.25, use_vq=True, num_quantizers=2, label_smoothing=0.01, ss_tv_lambda=0.002, xyz_align_alpha=0.0,
                 latent_tokens=8, tokenizer_heads=2, tokenizer_layers=1, tokenizer_dropout=0.1, reinit_dead_codes=True,
                 print_init=False, name="tiny")


def case_vqvae():
    """Whole-model call contract on a tiny configuration, eval mode, CPU."""
    torch.manual_seed(77)
    m = VQVAE(**VQVAE_CFG).eval()
    rs = np.random.RandomState(78)
    B, L = 3, 24
    x = np.zeros((B, L, 6), dtype=np.float32)
    x[..., :3] = rs.standard_normal((B, L, 3))
    x[np.arange(B)[:, None], np.arange(L)[None, :], 3 + rs.randint(0, 3, (B, L))] = 1.0
    mask = np.ones((B, L), dtype=bool)
    mask[1, 17:] = False
    mask[2, 9:] = False
    xt, mt = torch.from_numpy(x), torch.from_numpy(mask)
    with torch.no_grad():
        recons, target, vq_pack, _ = m(xt, mt)
        out = m.loss_function(recons, target, vq_pack, mt, ss_weight=0.7, rmsd_weight=1.3)
    zq, ze, idx, ppl, dead = vq_pack
    for k, v in m.state_dict().items():
        G[f"vqvae/sd/{k}"] = v.detach().cpu().numpy().copy()
    put("vqvae", x=x, mask=mask, recons=recons, zq=zq, ze=ze, idx=idx, ppl=ppl, dead=dead)
    for k in ("loss", "Reconstruction_Loss_XYZ", "XYZ_MSE_Raw", "Reconstruction_Loss_SS", "SS_Accuracy", "VQ_Loss",
              "SS_TV", "VQ_Perplexity", "VQ_DeadRatio", "RMSD_Raw"):
        G[f"vqvae/loss/{k}"] = np.asarray(float(out[k]))
    G["vqvae/loss_keys"] = np.array(sorted(out.keys()))


def main():
    torch.set_num_threads(8)
    case_vqvae()
    case_small_single()
    case_small_rvq()
    case_semantics()
    case_large("c2_like", 101, 512, 64, 1, 128, 64)
    case_large("c2_clustered", 102, 512, 64, 1, 64, 64, clustered=True)
    case_large("ragged_k", 103, 1000, 48, 1, 37, 29)          # K, D, N off every tile size
    case_large("one_row", 104, 300, 80, 1, 1, 1)
    case_large("c3_like", 105, 8192, 256, 1, 32, 64)
    case_large("c3_clustered", 106, 8192, 256, 1, 16, 64, clustered=True)
    case_large("stage2_rvq", 107, 1024, 512, 4, 32, 64)
    case_large("d128_scaled", 108, 2048, 128, 1, 64, 64, scale=7.5)
    case_survey_kat("kat_512_64", 1234, 512, 64, 1, 64, 64)
    case_survey_kat("kat_8192_256", 1234, 8192, 256, 1, 16, 64)
    case_survey_kat("kat_rvq_stage2", 1265, 1024, 512, 4, 64, 64)
    case_survey_kat("kat_512_64_big", 7, 512, 64, 1, 1024, 64)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
