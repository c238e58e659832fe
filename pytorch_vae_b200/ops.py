"""Tensor-level wrappers over the C ABI: validate, pass raw pointers, count kernel launches.

PyTorch is plumbing here (device memory, the current stream); every function
below enqueues hand-written sm_100a kernels from libvqb200.so and nothing else.
"""
from __future__ import annotations

import functools

import torch

from . import _cabi
from ._cabi import check, lib, ptr, stream_ptr

# kernels launched by this process through the library (bench.py reports the delta per step)
_launches = 0
last_fused_workspace = None
last_rvq_workspace = None


def launch_count() -> int:
    return _launches


def _count(n: int):
    global _launches
    _launches += n


def _on_device(fn):
    """Run ``fn`` with the CUDA device of its first CUDA tensor argument current: the library launches on the
    current device's current stream (``stream_ptr``), keeps its helper streams per device and sets kernel
    attributes per device, so a module moved with ``.to('cuda:1')`` must not launch on device 0."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for t in args:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if t.device.index != torch.cuda.current_device():
                    with torch.cuda.device(t.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapped


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pytorch-vae_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


class CodebookCache:
    """Per-code data derived from the codebook: bf16 copy, |e|^2/2 (fp32 and bf16-rounded), level norms."""

    def __init__(self, K_total: int, D: int, K_per: int, device):
        self.K_total, self.D, self.K_per = K_total, D, K_per
        self.levels = K_total // K_per
        # 16-bit operand planes: [0] bf16 (bf16_input mode), [1] fp16 bit patterns (fp32 mode)
        self.E_bf16 = torch.empty(2, K_total, D, dtype=torch.bfloat16, device=device)
        self.ee_half = torch.empty(2, K_total, dtype=torch.float32, device=device)
        self.level_meta = torch.empty(self.levels, _cabi.LEVEL_META_FLOATS, dtype=torch.float32, device=device)
        self.key = None

    def operand_ptr(self, mode: int, first_code: int) -> int:
        """Device pointer of the tensor-core operand copy for ``mode`` starting at code ``first_code``."""
        plane = 0 if mode == _cabi.MODE_BF16_INPUT else 1
        return self.E_bf16.data_ptr() + (plane * self.K_total + first_code) * self.D * 2

    @_on_device
    def prepare(self, E: torch.Tensor):
        _need_cuda(E)
        _f32c(E, "embedding")
        check(lib.vqb200_codebook_prepare(ptr(E), self.K_total, self.D, self.K_per, ptr(self.E_bf16),
                                          ptr(self.ee_half), ptr(self.level_meta), stream_ptr()),
              "vqb200_codebook_prepare")
        _count(1)


@_on_device
def search(z: torch.Tensor, E: torch.Tensor, cache: CodebookCache, level: int, mode: int,
           idx_out: torch.Tensor, idx_offset: int | None = None):
    """idx_out[n] = idx_offset + argmin_k |z_n - E[level*K_per + k]|^2 (one level)."""
    _need_cuda(z, E, idx_out)
    _f32c(z, "z")
    N, D = z.shape
    K = cache.K_per
    s = level * K
    if idx_offset is None:
        idx_offset = s
    ws_bytes = lib.vqb200_search_workspace_bytes(N, K, D, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    esz = E.element_size()
    check(lib.vqb200_search(ptr(z), N, D, E.data_ptr() + s * D * esz, cache.operand_ptr(mode, s),
                            cache.ee_half.data_ptr() + s * 4, cache.ee_half.data_ptr() + (cache.K_total + s) * 4,
                            cache.level_meta.data_ptr() + level * _cabi.LEVEL_META_FLOATS * 4, K, mode,
                            idx_offset, ptr(idx_out), ptr(ws), ws_bytes, stream_ptr()), "vqb200_search")
    _count(search_launches(N, K, D, mode))


@_on_device
def rvq_forward(z, E, cache: CodebookCache, mode: int, idx_out, zq_out=None, zq_st_out=None, sqerr_sum=None,
                hist=None, stats=None):
    """Eval-mode residual forward (all levels, finalize included) in one library call.  ``stats`` =
    ``(count_add, inv_elems, ep_usage, ep_cnt, stats_out)``: the statistics of ``stats_finalize`` in the same call (inside
    the persistent kernel where it runs)."""
    _need_cuda(z, E, idx_out)
    _f32c(z, "z")
    N, D = z.shape
    K, L = cache.K_per, cache.levels
    ws_bytes = lib.vqb200_rvq_forward_workspace_bytes(N, K, D, L, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    if stats is not None:
        count_add, inv_elems, ep_usage, ep_cnt, stats_out = stats
        check(lib.vqb200_rvq_forward_stats(ptr(z), N, D, ptr(E), cache.operand_ptr(mode, 0), cache.ee_half.data_ptr(),
                                           cache.ee_half.data_ptr() + cache.K_total * 4, ptr(cache.level_meta), K, L,
                                           mode, ptr(idx_out), ptr(zq_out), ptr(zq_st_out), ptr(sqerr_sum), ptr(hist),
                                           ptr(ws), ws_bytes, float(count_add), float(inv_elems), ptr(ep_usage),
                                           ptr(ep_cnt), ptr(stats_out), stream_ptr()), "vqb200_rvq_forward_stats")
        if not lib.vqb200_rvq_fused_supported(N, K, D, L, mode):
            _count(1)
    else:
        check(lib.vqb200_rvq_forward(ptr(z), N, D, ptr(E), cache.operand_ptr(mode, 0), cache.ee_half.data_ptr(),
                                     cache.ee_half.data_ptr() + cache.K_total * 4, ptr(cache.level_meta), K, L, mode,
                                     ptr(idx_out), ptr(zq_out), ptr(zq_st_out), ptr(sqerr_sum), ptr(hist), ptr(ws),
                                     ws_bytes, stream_ptr()), "vqb200_rvq_forward")
    global last_rvq_workspace
    last_rvq_workspace = ws           # persistent kernel: counters[0] = rows that took its exhaustive search (read lazily)
    _count(lib.vqb200_rvq_forward_launches(N, K, D, L, mode))


@_on_device
def rvq_train_forward(z, E, cache: CodebookCache, mode, decay, eps, ema_cluster_size, ema_embedding, idx_out, zq_out,
                      zq_st_out=None, sqerr_sum=None, hist=None, stats=None):
    """Training-mode residual forward with a local EMA update after every level, in one library call (``stats`` as in
    ``rvq_forward``)."""
    _need_cuda(z, E, idx_out, zq_out)
    _f32c(z, "z")
    N, D = z.shape
    K, L = cache.K_per, cache.levels
    ws_bytes = lib.vqb200_rvq_train_workspace_bytes(N, K, D, L, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    if stats is not None:
        count_add, inv_elems, ep_usage, ep_cnt, stats_out = stats
        check(lib.vqb200_rvq_train_forward_stats(ptr(z), N, D, ptr(E), ptr(cache.E_bf16), ptr(cache.ee_half),
                                                 ptr(cache.level_meta), K, L, mode, float(decay), float(1 - decay),
                                                 float(eps), ptr(ema_cluster_size), ptr(ema_embedding), ptr(idx_out),
                                                 ptr(zq_out), ptr(zq_st_out), ptr(sqerr_sum), ptr(hist), ptr(ws),
                                                 ws_bytes, float(count_add), float(inv_elems), ptr(ep_usage), ptr(ep_cnt),
                                                 ptr(stats_out), stream_ptr()), "vqb200_rvq_train_forward_stats")
        if not lib.vqb200_rvq_train_fused_supported(N, K, D, L, mode):
            _count(1)
    else:
        check(lib.vqb200_rvq_train_forward(ptr(z), N, D, ptr(E), ptr(cache.E_bf16), ptr(cache.ee_half),
                                           ptr(cache.level_meta), K, L, mode, float(decay), float(1 - decay), float(eps),
                                           ptr(ema_cluster_size), ptr(ema_embedding), ptr(idx_out), ptr(zq_out),
                                           ptr(zq_st_out), ptr(sqerr_sum), ptr(hist), ptr(ws), ws_bytes, stream_ptr()),
              "vqb200_rvq_train_forward")
    if lib.vqb200_rvq_train_fused_supported(N, K, D, L, mode):
        # persistent kernel: its counters follow the segment sums in the workspace (diagnostics, read lazily)
        global last_rvq_workspace
        seg_bytes = ((K * L * D * 4 + K * L * 4) + 255) // 256 * 256
        last_rvq_workspace = ws[seg_bytes:]
    _count(lib.vqb200_rvq_train_launches(N, K, D, L, mode))


def rvq_train_fused_supported(N, K_per, D, L, mode) -> bool:
    return bool(lib.vqb200_rvq_train_fused_supported(N, K_per, D, L, mode))


@_on_device
def rvq_train_begin(z, E, cache: CodebookCache, mode, decay, eps, ema_cluster_size, ema_embedding, idx_out, zq_out,
                    seg_sum, seg_cnt, zq_st_out=None, sqerr_sum=None, hist=None):
    """First half of the training forward (see include/vq_b200.h): every level searched, outputs written, this rank's
    EMA segment sums in ``seg_sum`` / ``seg_cnt``; ``rvq_train_finish`` applies the updates after the exchange."""
    _need_cuda(z, E, idx_out, zq_out, seg_sum, seg_cnt)
    _f32c(z, "z")
    N, D = z.shape
    K, L = cache.K_per, cache.levels
    ws_bytes = lib.vqb200_rvq_train_begin_workspace_bytes(N, K, D, L, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    check(lib.vqb200_rvq_train_begin(ptr(z), N, D, ptr(E), ptr(cache.E_bf16), ptr(cache.ee_half), ptr(cache.level_meta),
                                     K, L, mode, float(decay), float(1 - decay), float(eps), ptr(ema_cluster_size),
                                     ptr(ema_embedding), ptr(idx_out), ptr(zq_out), ptr(zq_st_out), ptr(sqerr_sum),
                                     ptr(hist), ptr(seg_sum), ptr(seg_cnt), ptr(ws), ws_bytes, stream_ptr()),
          "vqb200_rvq_train_begin")
    _count(4)


@_on_device
def rvq_train_finish(seg_sum, seg_cnt, E, cache: CodebookCache, decay, eps, ema_cluster_size, ema_embedding):
    D = E.shape[1]
    check(lib.vqb200_rvq_train_finish(ptr(seg_sum), ptr(seg_cnt), float(decay), float(1 - decay), float(eps),
                                      cache.K_per, cache.levels, D, ptr(ema_cluster_size), ptr(ema_embedding), ptr(E),
                                      ptr(cache.E_bf16), ptr(cache.ee_half), ptr(cache.level_meta), stream_ptr()),
          "vqb200_rvq_train_finish")
    _count(1)


@_on_device
def rvq_train_level(residual, E, cache: CodebookCache, level, mode, idx_out, zq_out, residual_out, hist, seg_sum,
                    seg_cnt):
    """search -> gather -> scatter-add of one training level (this rank's segment sums; the EMA finalize follows
    the caller's all-reduce)."""
    N, D = residual.shape
    K, L = cache.K_per, cache.levels
    ws_bytes = lib.vqb200_search_workspace_bytes(N, K, D, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=residual.device)
    check(lib.vqb200_rvq_train_level(ptr(residual), N, D, ptr(E), ptr(cache.E_bf16), ptr(cache.ee_half),
                                     ptr(cache.level_meta), K, L, level, mode, ptr(idx_out), ptr(zq_out),
                                     ptr(residual_out), ptr(hist), ptr(seg_sum), ptr(seg_cnt), ptr(ws), ws_bytes,
                                     stream_ptr()), "vqb200_rvq_train_level")
    _count(search_launches(N, K, D, mode) + 2)


@_on_device
def residual_prep(z, E, idx, cache: CodebookCache, next_level: int, mode: int, residual_out, z16_out, margin_out):
    """residual_out = z - E[idx] together with the next level's 16-bit operand copy and admission margins."""
    _need_cuda(z, E, idx, residual_out)
    _f32c(z, "z")
    N, D = z.shape
    check(lib.vqb200_residual_prep(ptr(z), ptr(E), ptr(idx), N, D, E.shape[0], mode,
                                   cache.level_meta.data_ptr() + next_level * _cabi.LEVEL_META_FLOATS * 4,
                                   ptr(residual_out), ptr(z16_out), ptr(margin_out), stream_ptr()),
          "vqb200_residual_prep")
    _count(1)


@_on_device
def search_prepped(z, z16, margin, E, cache: CodebookCache, level: int, mode: int, idx_out):
    """``search`` for rows whose operand copy and margins ``residual_prep`` already produced (tensor path only)."""
    _need_cuda(z, E, idx_out)
    _f32c(z, "z")
    N, D = z.shape
    K = cache.K_per
    s = level * K
    ws_bytes = lib.vqb200_search_workspace_bytes(N, K, D, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    check(lib.vqb200_search_prepped(ptr(z), ptr(z16), ptr(margin), N, D, E.data_ptr() + s * D * E.element_size(),
                                    cache.operand_ptr(mode, s), cache.ee_half.data_ptr() + s * 4,
                                    cache.ee_half.data_ptr() + (cache.K_total + s) * 4,
                                    cache.level_meta.data_ptr() + level * _cabi.LEVEL_META_FLOATS * 4, K, mode, s,
                                    ptr(idx_out), ptr(ws), ws_bytes, stream_ptr()), "vqb200_search_prepped")
    _count(4 * tc_chunks(N, K, D, mode))       # per chunk: tcgen05 search, re-rank, hand-back, unpack


def tc_chunks(N, K, D, mode) -> int:
    """Chunks the tensor-path search splits N rows into."""
    path = lib.vqb200_search_path(N, K, D, mode)
    if not path:
        return 0
    n = search_launches(N, K, D, mode)
    return max(1, (n - 1) // 4) if path == 2 else max(1, n // 5)


def search_launches(N, K, D, mode) -> int:
    return lib.vqb200_search_launches(N, K, D, mode)


def fused_supported(N, K, D, mode) -> bool:
    return bool(lib.vqb200_quantize_fused_supported(N, K, D, mode))


@_on_device
def quantize_fused(z, E, cache: CodebookCache, mode, idx_out, zq_out=None, zq_st_out=None, sqerr_sum=None,
                   hist=None, row_mask=None):
    """Single-level forward in one kernel: idx, z_q, z_q_st, sum (z_q - z)^2 and histogram from one read of z."""
    _need_cuda(z, E, idx_out)
    _f32c(z, "z")
    N, D = z.shape
    K = cache.K_per
    ws_bytes = lib.vqb200_quantize_fused_workspace_bytes(N, K, D, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    check(lib.vqb200_quantize_fused(ptr(z), N, D, ptr(E), cache.operand_ptr(mode, 0), cache.ee_half.data_ptr(),
                                    cache.ee_half.data_ptr() + cache.K_total * 4, ptr(cache.level_meta), K, mode, 0,
                                    ptr(idx_out), ptr(zq_out), ptr(zq_st_out), ptr(sqerr_sum), ptr(hist),
                                    ptr(row_mask), ptr(ws), ws_bytes, stream_ptr()), "vqb200_quantize_fused")
    global last_fused_workspace
    last_fused_workspace = ws          # counters[0] = rows handed to the exact kernel (read lazily: no sync here)
    _count(3)     # fused kernel, exact hand-back kernel, fix-up kernel


@_on_device
def quantize(z, E, cache: CodebookCache, mode, idx_out, zq_out=None, zq_st_out=None, sqerr_sum=None, hist=None,
             row_mask=None):
    """Single-level search + gather in one call; on the tensor path the gather of each chunk of rows overlaps the
    tensor kernel of the next chunk (three-stream pipeline inside the library)."""
    _need_cuda(z, E, idx_out)
    _f32c(z, "z")
    N, D = z.shape
    K = cache.K_per
    ws_bytes = lib.vqb200_search_workspace_bytes(N, K, D, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    check(lib.vqb200_quantize(ptr(z), N, D, ptr(E), cache.operand_ptr(mode, 0), cache.ee_half.data_ptr(),
                              cache.ee_half.data_ptr() + cache.K_total * 4, ptr(cache.level_meta), K, mode, 0,
                              ptr(idx_out), ptr(E), E.shape[0], ptr(zq_out), ptr(zq_st_out), ptr(sqerr_sum), ptr(hist),
                              ptr(row_mask), ptr(ws), ws_bytes, stream_ptr()), "vqb200_quantize")
    path = lib.vqb200_search_path(N, K, D, mode)
    # one gather per chunk -- or, on the side-job pipeline (path 2), only the last chunk's: the others ride on the
    # tensor kernels
    _count(search_launches(N, K, D, mode) + (1 if path == 2 else max(1, tc_chunks(N, K, D, mode))))


@_on_device
def gather(z, E, idx, zq_out=None, accumulate=False, zq_st_out=None, residual_out=None, sqerr_sum=None,
           hist=None, row_mask=None):
    _need_cuda(z, E, idx)
    _f32c(z, "z")
    N, D = z.shape
    check(lib.vqb200_gather(ptr(z), ptr(E), ptr(idx), N, D, E.shape[0], ptr(zq_out), int(accumulate),
                            ptr(zq_st_out), ptr(residual_out), ptr(sqerr_sum), ptr(hist), ptr(row_mask),
                            stream_ptr()), "vqb200_gather")
    _count(1)


@_on_device
def st_loss(z, zq, zq_st_out=None, sqerr_sum=None):
    _need_cuda(z, zq)
    check(lib.vqb200_st_loss(ptr(z), ptr(zq), z.numel(), ptr(zq_st_out), ptr(sqerr_sum), stream_ptr()),
          "vqb200_st_loss")
    _count(1)


@_on_device
def stats_finalize(hist, count_add, sqerr_sum, inv_elems, ep_usage, ep_cnt, stats_out):
    check(lib.vqb200_stats_finalize(ptr(hist), hist.numel(), float(count_add), ptr(sqerr_sum), float(inv_elems),
                                    ptr(ep_usage), ptr(ep_cnt), ptr(stats_out), stream_ptr()),
          "vqb200_stats_finalize")
    _count(1)


@_on_device
def stats_pack(hist, sqerr_sum, n_elems, out):
    check(lib.vqb200_stats_pack(ptr(hist), hist.numel(), ptr(sqerr_sum), float(n_elems), ptr(out), stream_ptr()),
          "vqb200_stats_pack")
    _count(1)


@_on_device
def stats_finalize_packed(packed, K_total, levels, D, ep_usage, ep_cnt, stats_out):
    check(lib.vqb200_stats_finalize_packed(ptr(packed), K_total, int(levels), int(D), ptr(ep_usage), ptr(ep_cnt),
                                           ptr(stats_out), stream_ptr()), "vqb200_stats_finalize_packed")
    _count(1)


@_on_device
def stats_exchange(hist, sqerr_sum, n_elems, K_total, levels, D, peer_ptrs_dev: int, rank: int, world: int, ep_usage,
                   ep_cnt, stats_out, spin_limit: int = 1 << 27):
    """pack -> exchange over peer memory -> reduce -> finalize in one kernel (include/vq_b200.h).  ``peer_ptrs_dev``
    is the device address of the array of the ranks' symmetric-buffer pointers."""
    check(lib.vqb200_stats_exchange(ptr(hist), int(K_total), ptr(sqerr_sum), float(n_elems), int(levels), int(D),
                                    int(peer_ptrs_dev), int(rank), int(world), int(spin_limit), ptr(ep_usage),
                                    ptr(ep_cnt), ptr(stats_out), stream_ptr()), "vqb200_stats_exchange")
    _count(1)


@_on_device
def scatter_add(z, idx, row_mask, seg_sum, seg_cnt):
    _need_cuda(z, idx, seg_sum)
    _f32c(z, "z")
    N, D = z.shape
    check(lib.vqb200_scatter_add(ptr(z), ptr(idx), ptr(row_mask), N, D, seg_cnt.numel(), ptr(seg_sum),
                                 ptr(seg_cnt), stream_ptr()), "vqb200_scatter_add")
    _count(1)


@_on_device
def ema_finalize(seg_sum, seg_cnt, decay, eps, ema_cluster_size, ema_embedding, E, cache: CodebookCache):
    # the reference forms (1 - decay) in Python double and the multiply rounds it to fp32
    check(lib.vqb200_ema_finalize(ptr(seg_sum), ptr(seg_cnt), float(decay), float(1 - decay), float(eps),
                                  cache.K_total, cache.D, cache.K_per, ptr(ema_cluster_size), ptr(ema_embedding),
                                  ptr(E), ptr(cache.E_bf16), ptr(cache.ee_half), ptr(cache.level_meta),
                                  stream_ptr()), "vqb200_ema_finalize")
    _count(1)


@_on_device
def kmeans_finalize(seg_sum, seg_cnt, E, cache: CodebookCache):
    """Lloyd step: E[k] <- mean of the rows assigned to k (empty clusters keep their centroid) + cache refresh."""
    check(lib.vqb200_kmeans_finalize(ptr(seg_sum), ptr(seg_cnt), cache.K_total, cache.D, cache.K_per, ptr(E),
                                     ptr(cache.E_bf16), ptr(cache.ee_half), ptr(cache.level_meta), stream_ptr()),
          "vqb200_kmeans_finalize")
    _count(1)


@_on_device
def rvq_finalize(z, idx_level0, level_stride, L, E, zq_out=None, zq_st_out=None, sqerr_sum=None, hist=None):
    """Residual-VQ tail in one pass: z_q (level-order sum), z_q_st, squared error and histogram from the indices.
    ``idx_level0`` is level 0's id vector; level l's starts ``l * level_stride`` elements further on."""
    _need_cuda(z, idx_level0, E)
    _f32c(z, "z")
    N, D = z.shape
    check(lib.vqb200_rvq_finalize(ptr(z), ptr(idx_level0), level_stride, N, D, L, ptr(E), E.shape[0], ptr(zq_out),
                                  ptr(zq_st_out), ptr(sqerr_sum), ptr(hist), stream_ptr()), "vqb200_rvq_finalize")
    _count(1)


@_on_device
def softmax_rows(z, E, alpha: float, beta=None, want_probs: bool = True, p_sum=None, row_stats=None):
    """probs[n, k] = softmax_k(alpha * z_n . e_k + beta[k]) through the tiled two-sweep kernel (include/vq_b200.h).
    Returns the [N, K] probabilities (``None`` with ``want_probs=False``); ``p_sum`` [K] receives their column sums."""
    _need_cuda(z, E)
    _f32c(z, "z")
    _f32c(E, "embedding")
    N, D = z.shape
    K = E.shape[0]
    P = torch.empty(N, K, dtype=torch.float32, device=z.device) if want_probs else None
    ws_bytes = lib.vqb200_softmax_rows_workspace_bytes(N, K)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=z.device)
    check(lib.vqb200_softmax_rows(ptr(z), N, D, ptr(E), ptr(beta), K, float(alpha), ptr(row_stats), ptr(P), ptr(p_sum),
                                  ptr(ws), ws_bytes, stream_ptr()), "vqb200_softmax_rows")
    _count(2)
    return P


def _matmul_fp32(a, b):
    """Plain library GEMM in full fp32 (never TF32, whatever the process-wide switch says)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return a @ b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@_on_device
def usage_probs(z, E, keep_probs: bool = False):
    """(p_code [K], probs [N, K] or None): p_code = mean_n softmax_k(z_n . e_k) (models/vq_vae.py:1305-1307).
    ``keep_probs`` also returns the probabilities (the backward needs them)."""
    N = z.shape[0]
    p_sum = torch.zeros(E.shape[0], dtype=torch.float32, device=z.device)
    P = softmax_rows(z, E, 1.0, None, want_probs=keep_probs, p_sum=p_sum)
    return p_sum / max(N, 1), P


def usage_probs_backward_from_probs(P, E, grad_p):
    """grad_z_n = (1/N) sum_j P_nj (g_j - sum_k P_nk g_k) e_j from the saved probabilities: elementwise passes over
    [N, K] and one plain fp32 GEMM."""
    N = P.shape[0]
    rowdot = _matmul_fp32(P, grad_p.unsqueeze(1)).squeeze(1)   # [N]
    dS = P * (grad_p.unsqueeze(0) - rowdot.unsqueeze(1))
    dS.mul_(1.0 / max(N, 1))
    return _matmul_fp32(dS, E)


@_on_device
def usage_probs_warp(z, E):
    """The first implementation (one warp per row, nothing stored): (p_code [K], row_stats [N, 2])."""
    _need_cuda(z, E)
    _f32c(z, "z")
    _f32c(E, "embedding")
    N, D = z.shape
    p_sum = torch.zeros(E.shape[0], dtype=torch.float32, device=z.device)
    row_stats = torch.empty(N, 2, dtype=torch.float32, device=z.device)
    check(lib.vqb200_usage_probs(ptr(z), N, D, ptr(E), E.shape[0], ptr(p_sum), ptr(row_stats), stream_ptr()),
          "vqb200_usage_probs")
    _count(1)
    return p_sum / max(N, 1), row_stats


@_on_device
def usage_probs_backward(z, E, row_stats, grad_p):
    N, D = z.shape
    out = torch.empty_like(z)
    check(lib.vqb200_usage_probs_backward(ptr(z), N, D, ptr(E), E.shape[0], ptr(row_stats), ptr(grad_p),
                                          1.0 / max(N, 1), ptr(out), stream_ptr()), "vqb200_usage_probs_backward")
    _count(1)
    return out


@_on_device
def soft_assign(z, E, tau: float, out=None):
    """z_soft = softmax(-|z - e|^2 / tau) @ E (models/vq_vae.py:838-843): the tiled row-softmax kernel
    (logits 2 z.e / tau - |e|^2 / tau: the row constant cancels) and one plain fp32 GEMM probs @ E.  The reference
    materialises the [N, K, D] differences; this stores the [N, K] probabilities only."""
    _need_cuda(z, E)
    _f32c(z, "z")
    _f32c(E, "embedding")
    if z.shape[1] > 512:
        raise RuntimeError("soft_assign: D <= 512")
    if z.shape[0] == 0:
        return torch.empty_like(z) if out is None else out
    t = max(float(tau), 1e-8)
    beta = (E * E).sum(1).mul_(-1.0 / t)
    P = softmax_rows(z, E, 2.0 / t, beta)
    res = _matmul_fp32(P, E)
    if out is not None:
        out.copy_(res)
        return out
    return res


@_on_device
def soft_assign_warp(z, E, tau: float, out=None):
    """The first implementation: one warp per row, online softmax, nothing materialised (vqb200_soft_assign)."""
    _need_cuda(z, E)
    _f32c(z, "z")
    _f32c(E, "embedding")
    N, D = z.shape
    if out is None:
        out = torch.empty_like(z)
    check(lib.vqb200_soft_assign(ptr(z), N, D, ptr(E), E.shape[0], float(tau), ptr(out), stream_ptr()),
          "vqb200_soft_assign")
    _count(1)
    return out


@_on_device
def commit_backward(grad_st, grad_commit, z, zq, scale, out):
    check(lib.vqb200_commit_backward(ptr(grad_st), ptr(grad_commit), ptr(z), ptr(zq), z.numel(), float(scale),
                                     ptr(out), stream_ptr()), "vqb200_commit_backward")
    _count(1)


_IDX_BYTES = {torch.int16: 2, torch.int32: 4, torch.int64: 8}


@_on_device
def relayout_indices(idx_level_major: torch.Tensor, Q: int, B: int, M: int, dtype=torch.int64,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """Level-major flat RVQ ids [Q*B*M] -> token-major [B, M*Q] (optionally narrowed), on the device."""
    _need_cuda(idx_level_major)
    if idx_level_major.dtype != torch.int64 or idx_level_major.numel() != Q * B * M:
        raise RuntimeError(f"expected {Q * B * M} int64 indices, got {tuple(idx_level_major.shape)} "
                           f"{idx_level_major.dtype}")
    if out is None:
        out = torch.empty(B, M * Q, dtype=dtype, device=idx_level_major.device)
    elif out.dtype != dtype or out.numel() != Q * B * M or not out.is_contiguous() or not out.is_cuda:
        raise RuntimeError(f"out must be a contiguous CUDA {dtype} tensor of {Q * B * M} elements")
    check(lib.vqb200_relayout_indices(ptr(idx_level_major.contiguous()), Q, B, M, ptr(out), _IDX_BYTES[dtype],
                                      stream_ptr()), "vqb200_relayout_indices")
    _count(1)
    return out


@_on_device
def indices_to_latent(idx: torch.Tensor, E: torch.Tensor, Q: int) -> torch.Tensor:
    """Token-major ids [n_tok*Q] -> z_q [n_tok, D], summing the Q levels in level order."""
    _need_cuda(idx, E)
    _f32c(E, "embedding")
    if idx.dtype not in _IDX_BYTES:
        raise RuntimeError(f"indices must be int16/int32/int64, got {idx.dtype}")
    idx = idx.contiguous().view(-1)
    if idx.numel() % Q != 0:
        raise ValueError(f"flattened indices length {idx.numel()} is not divisible by num_quantizers={Q}")
    n_tok = idx.numel() // Q
    out = torch.empty(n_tok, E.shape[1], dtype=torch.float32, device=E.device)
    check(lib.vqb200_indices_to_latent(ptr(idx), _IDX_BYTES[idx.dtype], n_tok, Q, ptr(E), E.shape[0], E.shape[1],
                                       ptr(out), stream_ptr()), "vqb200_indices_to_latent")
    _count(1)
    return out


@_on_device
def indices_to_memory(idx: torch.Tensor, P: torch.Tensor, Q: int, bias=None, ln_weight=None, ln_bias=None,
                      ln_eps: float = 1e-5) -> torch.Tensor:
    """Token-major ids [n_tok*Q] -> LayerNorm(sum_q P[id_q] + bias) [n_tok, H], P = E @ W_from_code^T (see
    include/vq_b200.h: the from_code Linear over the tokens becomes a gather, models/vq_vae.py:749)."""
    _need_cuda(idx, P)
    _f32c(P, "projected codebook")
    for name, t in (("bias", bias), ("ln_weight", ln_weight), ("ln_bias", ln_bias)):
        if t is not None:
            _f32c(t, name)
    if idx.dtype not in _IDX_BYTES:
        raise RuntimeError(f"indices must be int16/int32/int64, got {idx.dtype}")
    idx = idx.contiguous().view(-1)
    if idx.numel() % Q != 0:
        raise ValueError(f"flattened indices length {idx.numel()} is not divisible by num_quantizers={Q}")
    n_tok = idx.numel() // Q
    out = torch.empty(n_tok, P.shape[1], dtype=torch.float32, device=P.device)
    check(lib.vqb200_indices_to_memory(ptr(idx), _IDX_BYTES[idx.dtype], n_tok, Q, ptr(P), P.shape[0], P.shape[1],
                                       ptr(bias), ptr(ln_weight), ptr(ln_bias), float(ln_eps), ptr(out), stream_ptr()),
          "vqb200_indices_to_memory")
    _count(1)
    return out


@_on_device
def search_packed(z, E_slice, ee_half_slice, idx_offset, packed_out):
    """Codebook-sharded search: packed[n] = key(d) << 32 | (idx_offset + argmin) over this shard's codes."""
    _need_cuda(z, E_slice, packed_out)
    _f32c(z, "z")
    N, D = z.shape
    check(lib.vqb200_search_packed(ptr(z), N, D, ptr(E_slice), ptr(ee_half_slice), E_slice.shape[0], idx_offset,
                                   ptr(packed_out), stream_ptr()), "vqb200_search_packed")
    _count(1)


@_on_device
def search_slice(z, E, cache: CodebookCache, first_code: int, n_codes: int, mode: int, idx_out):
    """idx_out[n] = first_code + argmin over the codes [first_code, first_code + n_codes) of a single-level codebook
    (the slice a rank owns in the codebook-sharded search); same kernels as ``search``."""
    _need_cuda(z, E, idx_out)
    _f32c(z, "z")
    N, D = z.shape
    ws_bytes = lib.vqb200_search_workspace_bytes(N, n_codes, D, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    check(lib.vqb200_search(ptr(z), N, D, E.data_ptr() + first_code * D * 4, cache.operand_ptr(mode, first_code),
                            cache.ee_half.data_ptr() + first_code * 4,
                            cache.ee_half.data_ptr() + (cache.K_total + first_code) * 4, cache.level_meta.data_ptr(),
                            n_codes, mode, first_code, ptr(idx_out), ptr(ws), ws_bytes, stream_ptr()), "vqb200_search")
    _count(search_launches(N, n_codes, D, mode))


@_on_device
def pack_exact(z, E, idx, packed_out):
    """packed[n] = 40 bits of the exact fp64 score of z_n against E[idx[n]] | 24 bits of idx[n] (MIN-reducible)."""
    _need_cuda(z, E, idx, packed_out)
    _f32c(z, "z")
    N, D = z.shape
    check(lib.vqb200_pack_exact(ptr(z), N, D, ptr(E), E.shape[0], ptr(idx), ptr(packed_out), stream_ptr()),
          "vqb200_pack_exact")
    _count(1)


@_on_device
def minloc_unpack24(packed, idx_out):
    check(lib.vqb200_minloc_unpack24(ptr(packed), packed.numel(), ptr(idx_out), stream_ptr()),
          "vqb200_minloc_unpack24")
    _count(1)


@_on_device
def minloc_unpack(packed, idx_out):
    check(lib.vqb200_minloc_unpack(ptr(packed), packed.numel(), ptr(idx_out), stream_ptr()),
          "vqb200_minloc_unpack")
    _count(1)
