"""Generate tests/golden/train_golden.npz from the LIVE reference class: training-mode forwards (EMA update
on) at shapes that take the tensor-core search in the product -- the stage-2 / BASELINE configs[4] shape
(4 x 1024 codes, D = 512, 8192 rows) and the configs[1] codebook (K = 512, D = 64) -- starting from the
reference's default ZERO EMA buffers (models/vq_vae.py:52-53), so every code that receives no row in the first
update collapses to the zero vector (:85-88) and later steps search a codebook full of exact duplicates.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_train.py

Per step the file keeps the indices (int16), the statistics, the commitment mse, ``ema_cluster_size`` in full,
and of ``embedding`` / ``ema_embedding`` the per-row norms plus 32 sampled rows (the stage-2 buffers are 8 MB
each); the small codebook is kept in full.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from models.vq_vae import VectorQuantizerEMA  # noqa: E402
from synth import large_case_inputs, train_step_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_golden.npz")
G = {}

CASES = {   # name: (seed, K_per, D, L, B, M, decay, steps)
    "c5_train": (201, 1024, 512, 4, 128, 64, 0.98, 3),
    "c2_train": (202, 512, 64, 1, 128, 64, 0.98, 3),
}


def case(name, seed, K_per, D, L, B, M, decay, steps):
    E, _ = large_case_inputs(seed, K_per, D, L, 1, 1)
    q = VectorQuantizerEMA(K_per, D, num_quantizers=L, print_init=False, decay=decay, beta=0.0005).train()
    q.embedding.copy_(torch.from_numpy(E))
    rows = np.random.RandomState(seed + 5).choice(K_per * L, 32, replace=False)
    G[f"{name}/meta"] = np.array([seed, K_per, D, L, B, M, steps], dtype=np.int64)
    G[f"{name}/decay"] = np.array(decay)
    G[f"{name}/rows"] = rows
    for s in range(steps):
        z = train_step_inputs(seed, s, B, M, D)
        zt = torch.from_numpy(z)
        st, zq, idx, stats = q(zt, do_ema_update=True)
        p = f"{name}/step{s}"
        G[f"{p}/idx"] = idx.numpy().astype(np.int16)
        G[f"{p}/stats"] = stats.numpy().copy()
        G[f"{p}/commit"] = np.array(float(torch.nn.functional.mse_loss(zq, zt)))
        G[f"{p}/ema_cluster_size"] = q.ema_cluster_size.numpy().copy()
        for buf in ("embedding", "ema_embedding"):
            a = getattr(q, buf).numpy()
            G[f"{p}/{buf}_norm"] = np.sqrt((a.astype(np.float64) ** 2).sum(1)).astype(np.float32)
            G[f"{p}/{buf}_rows"] = a[rows].copy()
        if K_per * L * D <= 65536:
            G[f"{p}/embedding"] = q.embedding.numpy().copy()
        zero = int((q.embedding.abs().sum(1) == 0).sum())
        G[f"{p}/n_zero_codes"] = np.array(zero)
        print(name, s, "ppl", float(stats[0]), "dead", float(stats[1]), "zero codes", zero)


def main():
    torch.set_num_threads(8)
    for name, cfg in CASES.items():
        case(name, *cfg)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
