"""Multi-GPU layout of the path: one process per GPU, rows sharded, codebook replicated.

Rows are independent, so the forward needs NO data-path collective.  The only
exchanges are (1) one small all-reduce(SUM) of ``[sum sq err | element count | histogram]`` for
global statistics / loss, (2) in training one all-reduce(SUM) of the EMA segment sums, and
(3) for codebooks too large to replicate, an all-reduce(MIN) over packed ``(distance key, index)``
words.  Everything here is backend-agnostic host logic (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_SIGN = -(2 ** 63)


def shard_rows(n_rows: int, world: int, rank: int):
    """Contiguous balanced partition of ``n_rows``: the first ``n_rows % world`` ranks get one extra row.
    (The reference pads by duplicating samples, scripts/extract_code_indices.py:135-140; a contiguous
    split needs no padding and keeps every row exactly once.)"""
    base, extra = divmod(int(n_rows), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_codes(k_codes: int, world: int, rank: int, multiple: int = 8):
    """Contiguous code ranges for the codebook-sharded search, aligned to ``multiple`` codes."""
    per = -(-k_codes // world)
    per = -(-per // multiple) * multiple
    start = min(rank * per, k_codes)
    return start, min(start + per, k_codes)


def dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class PeerStatsExchange:
    """Symmetric buffers for ``ops.stats_exchange`` (one NVLink / NVSwitch domain): the statistics all-reduce of the
    forward as ONE kernel over peer memory instead of pack kernel -> NCCL -> finalize kernel.  Built lazily, once per
    (device, K_total); ``None`` from ``get`` when symmetric memory cannot be set up (other backend, no peer access),
    in which case the caller keeps the NCCL path."""

    _cache = {}

    def __init__(self, device, K_total: int):
        import torch.distributed._symmetric_memory as symm
        from ._cabi import lib
        world = dist.get_world_size()
        nbytes = int(lib.vqb200_stats_exchange_buffer_bytes(int(K_total), world))
        self.buf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.buf.zero_()
        torch.cuda.current_stream(device).synchronize()
        dist.barrier()                                          # every rank's flags and epoch counter are zero
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.peer_ptrs_dev = int(self.hdl.buffer_ptrs_dev)

    @classmethod
    def get(cls, device, K_total: int):
        key = (str(device), int(K_total))
        if key not in cls._cache:
            inst = None
            try:
                if dist.get_backend() != "nccl" or torch.device(device).type != "cuda":
                    raise RuntimeError("peer memory needs CUDA devices")
                inst = cls(device, K_total)
            except Exception as e:                              # noqa: BLE001 -- any set-up failure: keep NCCL
                import warnings
                warnings.warn(f"peer-memory statistics exchange unavailable ({type(e).__name__}: {e}); using NCCL")
            # every rank must take the same route: one rank falling back alone would leave the others waiting for its flag
            try:
                ok = torch.tensor([1 if inst is not None else 0], dtype=torch.int32, device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 0:
                    inst = None
            except Exception:                                   # noqa: BLE001
                inst = None
            cls._cache[key] = inst
        return cls._cache[key]


def allreduce_stats(sqerr_sum: torch.Tensor, n_elems: int, hist: torch.Tensor, group=None):
    """ONE all-reduce of ``[sum sq err | element count | histogram]`` (float64: counts stay exact to 2^53).

    Returns (global mean sq err as float64[1], global histogram as int32[K]).
    """
    # torch.full is a device-side fill: no pageable host->device copy, hence no host sync on the step path
    pack = torch.cat([sqerr_sum.reshape(1).to(torch.float64),
                      torch.full((1,), float(n_elems), dtype=torch.float64, device=hist.device),
                      hist.to(torch.float64)])
    dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    mean = (pack[:1] / pack[1:2].clamp_min(1.0)).contiguous()
    return mean, pack[2:].to(torch.int32)


def allreduce_minloc(packed: torch.Tensor, group=None) -> torch.Tensor:
    """Tie-stable min-loc across ranks.  ``packed`` holds uint64 words ``key(d) << 32 | idx`` stored in an
    int64 tensor; flipping the sign bit turns unsigned order into signed order so ReduceOp.MIN (NCCL has
    no MINLOC) picks the smallest distance and, on equal distances, the smallest index."""
    flipped = packed ^ _SIGN
    dist.all_reduce(flipped, op=dist.ReduceOp.MIN, group=group)
    return flipped ^ _SIGN


def codebook_sharded_search(q, flat: torch.Tensor, group=None) -> torch.Tensor:
    """Nearest code when the codebook is split over ranks: every rank holds all rows ``flat`` [N, D] and
    scans only its slice of the (single-level) codebook; the winners are combined with one
    all-reduce(MIN) of N packed words.  Returns int64 indices [N], identical on every rank."""
    from . import ops
    if q.num_quantizers != 1:
        raise NotImplementedError("codebook-sharded search is defined for single-level codebooks")
    world = dist.get_world_size(group) if dist_ready() else 1
    rank = dist.get_rank(group) if dist_ready() else 0
    return sharded_search_slice(q, flat, world, rank, lambda p: allreduce_minloc(p, group) if world > 1 else p)


def sharded_search_slice(q, flat: torch.Tensor, world: int, rank: int, reduce_min):
    """One rank's share of the codebook-sharded search, with the cross-rank MIN supplied by the caller
    (``reduce_min(packed int64) -> packed int64``; the tests simulate the ranks on one device).  On shapes the tensor
    path serves (slice of >= 128 codes, D a multiple of 64, q.K <= 2^24) the slice is searched by the tcgen05 kernels
    with the exact re-rank and the winner's exact fp64 score is packed (40 bits of score | 24 bits of id);
    otherwise by the exact SIMT kernel (32-bit fp32 distance key | 32 bits of id)."""
    from . import _cabi, ops
    from .quantizer import _MODES
    s, e = shard_codes(q.K, world, rank)
    cache = q._codebook_cache()
    flat = flat.contiguous()
    N, D = flat.shape
    mode = _MODES[q.search_mode]
    # every rank must take the same route: decide on the SMALLEST slice
    smallest = min(b - a for a, b in (shard_codes(q.K, world, r) for r in range(world)))
    tensor = smallest > 0 and q.K <= (1 << 24) and mode == _cabi.MODE_FP32_EXACT and \
        bool(_cabi.lib.vqb200_search_path(N, smallest, D, mode))
    packed = torch.full((N,), -1, dtype=torch.int64, device=flat.device)   # all ones = +inf key
    if tensor:
        idx = torch.empty(N, dtype=torch.int64, device=flat.device)
        ops.search_slice(flat, q.embedding, cache, s, e - s, mode, idx)
        ops.pack_exact(flat, q.embedding, idx, packed)
    elif e > s:
        ops.search_packed(flat, q.embedding[s:e], cache.ee_half[0, s:e], s, packed)
    packed = reduce_min(packed)
    out = torch.empty_like(packed)
    (ops.minloc_unpack24 if tensor else ops.minloc_unpack)(packed, out)
    return out
