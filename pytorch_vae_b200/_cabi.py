"""ctypes binding of libvqb200.so (the C ABI declared in include/vq_b200.h).

The library is the product: if it is missing or fails to load, importing the
quantizer fails loudly -- there is no eager/PyTorch or CPU fallback anywhere in
this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvqb200.so")

MODE_FP32_EXACT = 0
MODE_BF16_INPUT = 1
MAX_LEVELS = 32
LEVEL_META_FLOATS = 8
ABI_VERSION = 18

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/vq_b200.h one to one
SIGNATURES = {
    "vqb200_abi_version": (_i, []),
    "vqb200_status_string": (C.c_char_p, [_i]),
    "vqb200_timing_enable": (_i, [_i]),
    "vqb200_timing_collect": (_i, [_p, _p]),
    "vqb200_search_path": (_i, [_i64, _i, _i, _i]),
    "vqb200_codebook_prepare": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "vqb200_search_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "vqb200_search_launches": (_i, [_i64, _i, _i, _i]),
    "vqb200_search": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _sz, _p]),
    "vqb200_rvq_forward_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_forward_launches": (_i, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_fused_supported": (_i, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_forward_stats": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz,
                                      C.c_float, C.c_double, _p, _p, _p, _p]),
    "vqb200_rvq_train_forward_stats": (_i, [_p, _i64, _i, _p, _p, _p, _p, _i, _i, _i, C.c_float, C.c_float, C.c_float,
                                            _p, _p, _p, _p, _p, _p, _p, _p, _sz, C.c_float, C.c_double, _p, _p, _p, _p]),
    "vqb200_rvq_train_fused_supported": (_i, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_train_begin_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_train_begin": (_i, [_p, _i64, _i, _p, _p, _p, _p, _i, _i, _i, C.c_float, C.c_float, C.c_float, _p, _p,
                                    _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "vqb200_rvq_train_finish": (_i, [_p, _p, C.c_float, C.c_float, C.c_float, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "vqb200_stats_exchange_buffer_bytes": (_sz, [_i, _i]),
    "vqb200_stats_exchange": (_i, [_p, _i, _p, C.c_double, _i, _i, _p, _i, _i, C.c_uint64, _p, _p, _p, _p]),
    "vqb200_softmax_rows_workspace_bytes": (_sz, [_i64, _i]),
    "vqb200_softmax_rows": (_i, [_p, _i64, _i, _p, _p, _i, C.c_float, _p, _p, _p, _p, _sz, _p]),
    "vqb200_indices_to_memory": (_i, [_p, _i, _i64, _i, _p, _i, _i, _p, _p, _p, C.c_float, _p, _p]),
    "vqb200_rvq_forward": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "vqb200_rvq_train_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_train_launches": (_i, [_i64, _i, _i, _i, _i]),
    "vqb200_rvq_train_forward": (_i, [_p, _i64, _i, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _p, _p, _p, _p, _p, _p, _p,
                                      _p, _sz, _p]),
    "vqb200_rvq_train_level": (_i, [_p, _i64, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "vqb200_residual_prep": (_i, [_p, _p, _p, _i64, _i, _i, _i, _p, _p, _p, _p, _p]),
    "vqb200_search_prepped": (_i, [_p, _p, _p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _sz, _p]),
    "vqb200_quantize_fused_supported": (_i, [_i64, _i, _i, _i]),
    "vqb200_quantize_fused_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "vqb200_quantize_fused": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "vqb200_quantize": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "vqb200_gather": (_i, [_p, _p, _p, _i64, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]),
    "vqb200_st_loss": (_i, [_p, _p, _i64, _p, _p, _p]),
    "vqb200_stats_finalize": (_i, [_p, _i, _f, _p, _d, _p, _p, _p, _p]),
    "vqb200_stats_pack": (_i, [_p, _i, _p, _d, _p, _p]),
    "vqb200_stats_finalize_packed": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "vqb200_scatter_add": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p, _p]),
    "vqb200_ema_finalize": (_i, [_p, _p, _f, _f, _f, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "vqb200_kmeans_finalize": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "vqb200_rvq_finalize": (_i, [_p, _p, _i64, _i64, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    "vqb200_usage_probs": (_i, [_p, _i64, _i, _p, _i, _p, _p, _p]),
    "vqb200_usage_probs_backward": (_i, [_p, _i64, _i, _p, _i, _p, _p, _f, _p, _p]),
    "vqb200_soft_assign": (_i, [_p, _i64, _i, _p, _i, _f, _p, _p]),
    "vqb200_commit_backward": (_i, [_p, _p, _p, _p, _i64, _f, _p, _p]),
    "vqb200_relayout_indices": (_i, [_p, _i, _i64, _i64, _p, _i, _p]),
    "vqb200_indices_to_latent": (_i, [_p, _i, _i64, _i, _p, _i, _i, _p, _p]),
    "vqb200_search_packed": (_i, [_p, _i64, _i, _p, _p, _i, _i64, _p, _p]),
    "vqb200_minloc_unpack": (_i, [_p, _i64, _p, _p]),
    "vqb200_pack_exact": (_i, [_p, _i64, _i, _p, _i, _p, _p, _p]),
    "vqb200_minloc_unpack24": (_i, [_p, _i64, _p, _p]),
}


class VQB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C pytorch_vae_b200/csrc`). This package has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    got = lib.vqb200_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libvqb200.so ABI {got} != binding ABI {ABI_VERSION}: rebuild the library")
    return lib


lib = _load()


def check(status: int, what: str):
    if status != 0:
        msg = lib.vqb200_status_string(status).decode()
        raise VQB200Error(f"{what} failed with status {status}: {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    """Raw handle of the current stream of the current device.  (torch.cuda.current_stream().cuda_stream builds a Stream
    object and resolves the device through several Python layers: 10 us per call, two calls per training step of a
    0.2 ms step; the raw query is one C call.)"""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return raw(torch._C._cuda_getDevice())
    return torch.cuda.current_stream().cuda_stream
