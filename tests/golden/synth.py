"""Seeded synthetic inputs shared by make_golden.py and the parity tests.

``np.random.RandomState`` streams are frozen by numpy policy, so the same seed
gives the same bytes in the build container and on the GPU box.
"""
import numpy as np


def rs_inputs(seed, K_total, D, B, M, scale=None):
    rs = np.random.RandomState(seed)
    E = (rs.standard_normal((K_total, D)) / np.sqrt(D)).astype(np.float32)
    z = rs.standard_normal((B, M, D)).astype(np.float32)
    if scale is not None:
        z *= np.float32(scale)
    return E, z


def large_case_inputs(seed, K_per, D, L, B, M, scale=None, clustered=False):
    E, z = rs_inputs(seed, K_per * L, D, B, M, scale)
    if L > 1:
        for lvl in range(1, L):
            E[lvl * K_per:(lvl + 1) * K_per] *= np.float32(0.6 ** lvl)
    if clustered:
        rs = np.random.RandomState(seed + 1)
        src = rs.randint(0, K_per, size=(B, M))
        z = (E[src] + 0.1 / np.sqrt(D) * rs.standard_normal((B, M, D))).astype(np.float32)
    return E, z


def train_step_inputs(seed, step, B, M, D):
    """Latents of training step ``step`` of a train_golden.npz case (fresh rows every step)."""
    return np.random.RandomState(seed + 10 + step).standard_normal((B, M, D)).astype(np.float32)
