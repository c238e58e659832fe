"""Teacher-forced replay of training-mode forwards for the parity tests (test infrastructure).

A multi-step training run cannot be compared index by index with a free-running reference: one legitimate
near-tie flip (fp64 distance gap < 1e-6 relative) moves a row between two codes, shifts both codes' EMA means by
O(1/count) and so changes later decisions.  The replay therefore checks every decision GIVEN IDENTICAL INPUTS:
it walks the levels of a step with the indices the implementation under test returned, and at each level
  * recomputes the oracle's pick (oracle/vq_oracle.py, models/vq_vae.py:238-245) from the same residual and the
    same codebook state and applies the near-tie rule to the rows that differ (none may fall outside it);
  * applies the reference's EMA update (models/vq_vae.py:77-89) with the returned indices,
so that the caller can compare z_q, the loss and the three EMA buffers after the step with a tolerance that
only has to absorb summation order.
"""
import numpy as np

from oracle import vq_oracle as O


def replay_step(oq: "O.OracleQuantizer", z, idx_impl, do_ema=True):
    """Advance ``oq`` (an OracleQuantizer holding the state BEFORE the step) through one training forward using
    the implementation's indices.  Returns a dict: mismatches / outside-allowance counts per level against the
    oracle's own picks, z_q (level-order sum) and the commitment mse."""
    z = np.asarray(z, dtype=np.float32)
    D = z.shape[-1]
    flat = z.reshape(-1, D)
    N, L, K_per = flat.shape[0], oq.num_quantizers, oq.K_per
    idx_impl = np.asarray(idx_impl, dtype=np.int64).reshape(L, N)
    residual = flat
    zq_sum = None
    mism, outside = [], []
    for lvl in range(L):
        s = lvl * K_per
        emb_l = oq.embedding[s:s + K_per].copy()
        local = idx_impl[lvl] - s
        assert local.min() >= 0 and local.max() < K_per, "level-major GLOBAL ids expected (models/vq_vae.py:246)"
        pick = O.nearest_code(residual, emb_l)
        mm, out = O.near_tie_rows(residual, emb_l, local, pick)
        mism.append(int(mm.size))
        outside.append(int(out.size))
        zq_l = emb_l[local]                                   # gathered BEFORE this level's update (:248 -> :251)
        zq_sum = zq_l.copy() if zq_sum is None else zq_sum + zq_l
        if do_ema:
            oq.ema_update(residual, idx_impl[lvl])
        residual = residual - zq_l
    zq = zq_sum.reshape(z.shape)
    return {"mismatch": mism, "outside": outside, "zq": zq, "commit": O.commitment_mse(zq, z)}


def chain_alive(idx_a, idx_b, L):
    """Chain-aware comparison of two level-major id vectors: a row is dropped from deeper levels once it differs.
    Returns (rows that differ at the level where they first differ, per level; fraction of rows identical)."""
    a = np.asarray(idx_a, dtype=np.int64).reshape(L, -1)
    b = np.asarray(idx_b, dtype=np.int64).reshape(L, -1)
    alive = np.ones(a.shape[1], bool)
    first = []
    for lvl in range(L):
        bad = alive & (a[lvl] != b[lvl])
        first.append(np.nonzero(bad)[0])
        alive &= ~bad
    return first, float(alive.mean())
