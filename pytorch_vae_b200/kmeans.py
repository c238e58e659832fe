"""k-means codebook initialiser on the quantizer's own kernels.

The reference starts stage 2 from centroids computed by an external script and copied in by
``VQVAE.init_codebook_from_centroids`` (``run.py:74-89,150-155``, ``models/vq_vae.py:577-613``): a ``[K, D]`` array
for a single codebook or ``[L, K_per, D]`` for the residual quantizer.  Lloyd's algorithm is the hot path in a
loop -- nearest-code search (tcgen05 kernel), segment sums (``vqb200_scatter_add``) and a mean
(``vqb200_kmeans_finalize``) -- so the same library produces those arrays on the box:

    cent = kmeans_fit(latents, K)                      # [K, D]
    cent = rvq_kmeans_fit(latents, K_per, levels)      # [L, K_per, D], each level fitted to the residual of the last
    model.init_codebook_from_centroids(cent)

Assignments are the exact ones of the quantizer (fp64-arbitrated nearest code, lowest index on ties); an empty
cluster keeps its centroid.  Multi-GPU: rows sharded, one all-reduce of the segment sums per iteration.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _cabi, ops, sharding

Tensor = torch.Tensor


def _check(z: Tensor, K: int):
    if not z.is_cuda:
        raise RuntimeError("kmeans runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if z.dim() != 2 or z.dtype != torch.float32:
        raise RuntimeError(f"expected float32 latents [N, D], got {tuple(z.shape)} {z.dtype}")
    if z.shape[1] % 4 != 0:
        raise ValueError("D must be a multiple of 4")
    if K < 1:
        raise ValueError("K must be positive")


@torch.no_grad()
def lloyd_step(z: Tensor, E: Tensor, cache: ops.CodebookCache, idx: Tensor, seg: Tensor,
               sync: bool = False) -> None:
    """One Lloyd iteration in place: assign rows of ``z`` to ``E`` (writes ``idx``), then move the centroids."""
    K, D = E.shape
    ops.search(z, E, cache, 0, _cabi.MODE_FP32_EXACT, idx)
    seg.zero_()
    seg_sum, seg_cnt = seg[: K * D], seg[K * D:]
    ops.scatter_add(z, idx, None, seg_sum, seg_cnt)
    if sync and sharding.dist_ready():
        torch.distributed.all_reduce(seg)
    ops.kmeans_finalize(seg_sum, seg_cnt, E, cache)


@torch.no_grad()
def kmeans_fit(z: Tensor, K: int, iters: int = 20, seed: int = 0, init: Optional[Tensor] = None,
               sync: bool = False, return_assignments: bool = False):
    """Lloyd's k-means of the rows of ``z`` [N, D].  ``init``: starting centroids [K, D] (default: K distinct rows
    drawn with ``torch.Generator().manual_seed(seed)`` on the host, so every rank draws the same ones when the
    same ``z`` prefix is visible; pass ``init`` explicitly for sharded data).  Returns centroids [K, D] fp32
    (and the final assignments + counts when asked)."""
    _check(z, K)
    z = z.contiguous()
    N, D = z.shape
    if init is None:
        if N < K:
            raise ValueError(f"need at least K={K} rows to draw the initial centroids, got {N}")
        g = torch.Generator().manual_seed(int(seed))
        rows = torch.randperm(N, generator=g)[:K].to(z.device)
        E = z.index_select(0, rows).clone()
    else:
        if tuple(init.shape) != (K, D):
            raise ValueError(f"init must be [{K}, {D}], got {tuple(init.shape)}")
        E = init.to(device=z.device, dtype=torch.float32).contiguous().clone()
    cache = ops.CodebookCache(K, D, K, z.device)
    cache.prepare(E)
    idx = torch.empty(N, dtype=torch.int64, device=z.device)
    seg = torch.empty(K * D + K, dtype=torch.float32, device=z.device)
    for _ in range(int(iters)):
        lloyd_step(z, E, cache, idx, seg, sync)
    if not return_assignments:
        return E
    ops.search(z, E, cache, 0, _cabi.MODE_FP32_EXACT, idx)       # assignments to the FINAL centroids
    counts = torch.bincount(idx, minlength=K)
    return E, idx, counts


@torch.no_grad()
def rvq_kmeans_fit(z: Tensor, K_per: int, levels: int, iters: int = 20, seed: int = 0,
                   sync: bool = False, inits: Optional[Tensor] = None) -> Tensor:
    """Residual k-means: level l is fitted to what levels < l left over.  Returns [levels, K_per, D], the layout
    ``init_codebook_from_centroids`` takes for the residual quantizer (models/vq_vae.py:590-601).  ``inits``
    (optional, [levels, K_per, D]) gives the starting centroids of every level instead of seeded row draws."""
    _check(z, K_per)
    if inits is not None and tuple(inits.shape) != (levels, K_per, z.shape[1]):
        raise ValueError(f"inits must be [{levels}, {K_per}, {z.shape[1]}], got {tuple(inits.shape)}")
    residual = z.contiguous()
    N, D = residual.shape
    out = torch.empty(levels, K_per, D, dtype=torch.float32, device=z.device)
    nxt = torch.empty_like(residual)
    for lvl in range(levels):
        E, idx, _ = kmeans_fit(residual, K_per, iters, seed + lvl, init=None if inits is None else inits[lvl],
                               sync=sync, return_assignments=True)
        out[lvl] = E
        if lvl + 1 < levels:                              # residual <- residual - E[idx] (the gather kernel's own pass)
            ops.gather(residual, E, idx, residual_out=nxt)
            residual, nxt = nxt, (torch.empty_like(residual) if lvl == 0 else residual)
    return out
