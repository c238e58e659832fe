"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python class mirrors the reference's surface, and nothing falls back to a CPU path."""
import inspect
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import pytorch_vae_b200 as vq  # noqa: E402


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vq_b200.h")).read()
    return sorted(set(re.findall(r"VQB200_API[^;]*?\b(vqb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 16
    assert sorted(vq._cabi.SIGNATURES) == syms            # binding and header agree one to one
    out = subprocess.run(["nm", "-D", "--defined-only", vq._cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (vqb200_\w+)", out))
    assert exported == set(syms)
    # no-compute calls work without a GPU
    assert vq._cabi.lib.vqb200_abi_version() == vq._cabi.ABI_VERSION
    assert vq._cabi.lib.vqb200_status_string(-2).startswith(b"unsupported shape")
    assert vq._cabi.lib.vqb200_search_workspace_bytes(1 << 20, 512, 64, 0) > 0


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", vq._cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


REF_CTOR = ["num_embeddings", "embedding_dim", "beta", "decay", "eps", "reinit_dead_codes", "reinit_prob",
            "dead_usage_threshold", "print_init", "diag_qe_cap", "diag_qe_bins", "num_quantizers"]
REF_BUFFERS = [("embedding", (96, 16)), ("ema_cluster_size", (96,)), ("ema_embedding", (96, 16)),
               ("_ep_usage", (96,)), ("_ep_top1_sum", (1,)), ("_ep_top2_sum", (1,)), ("_ep_cnt", (1,)),
               ("_ep_qe_sum", (1,)), ("_ep_qe_hist", (64,))]


def test_class_surface_matches_reference():
    params = list(inspect.signature(vq.VectorQuantizerEMA.__init__).parameters)[1:]
    assert params[:len(REF_CTOR)] == REF_CTOR             # same names, same order (models/vq_vae.py:20-34)
    fwd = list(inspect.signature(vq.VectorQuantizerEMA.forward).parameters)[1:]
    assert fwd == ["z_e", "do_ema_update", "allow_reinit", "mask"]
    q = vq.VectorQuantizerEMA(32, 16, num_quantizers=3, print_init=False)
    sd = q.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == REF_BUFFERS
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert (q.K, q.K_per, q.D, q.num_quantizers) == (96, 32, 16, 3)
    assert len(list(q.parameters())) == 0                 # buffers only: nothing for DDP to all-reduce
    assert abs(float(q.embedding.std()) - 0.25) < 0.03    # randn / sqrt(D)
    for name in ("_ema_update", "_maybe_reinit_dead_codes", "reset_epoch_stats", "get_epoch_stats",
                 "get_embedding_snapshot"):
        assert callable(getattr(q, name))
    assert vq.VectorQuantizer is vq.VectorQuantizerEMA
    es = q.get_epoch_stats()
    assert set(es) == {"usage_hist", "margin_mean", "qe_mean", "qe_p90", "n_positions", "perplexity", "dead_ratio"}
    if os.path.isdir("/root/reference"):                  # build container only: compare with the live class
        sys.path.insert(0, "/root/reference")
        from models.vq_vae import VectorQuantizerEMA as Ref
        r = Ref(32, 16, num_quantizers=3, print_init=False)
        assert [(k, tuple(v.shape)) for k, v in r.state_dict().items()] == REF_BUFFERS
        assert list(inspect.signature(Ref.__init__).parameters)[1:] == REF_CTOR
        q.load_state_dict(r.state_dict(), strict=True)
        assert torch.equal(q.embedding, r.embedding)


def test_no_cpu_fallback():
    q = vq.VectorQuantizerEMA(32, 16, print_init=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q(torch.randn(2, 4, 16))
    with pytest.raises(ValueError):
        q(torch.randn(8, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        vq.ops.gather(torch.zeros(4, 16), torch.zeros(8, 16), torch.zeros(4, dtype=torch.int64))
    with pytest.raises(ValueError):
        vq.VectorQuantizerEMA(32, 18, print_init=False)
    # the product never touches the oracle
    pkg = os.path.join(ROOT, "pytorch_vae_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, fn)).read().lower(), fn


def test_install_rebinds_reference_global():
    fake = types.ModuleType("fake_models_vq_vae")
    fake.VectorQuantizerEMA = object
    sys.modules[fake.__name__] = fake
    vq.install(fake.__name__)
    assert fake.VectorQuantizerEMA is vq.VectorQuantizerEMA
    vq.uninstall(fake.__name__)
    assert fake.VectorQuantizerEMA is object


def test_shard_partitions():
    from pytorch_vae_b200 import sharding as S
    for n, w in [(0, 4), (7, 8), (1 << 22, 8), (1000003, 3)]:
        spans = [S.shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    for k, w in [(8192, 8), (1000, 3), (16, 8)]:
        spans = [S.shard_codes(k, w, r) for r in range(w)]
        assert spans[0][0] == 0 and max(e for _, e in spans) == k
        assert all(a[1] == b[0] or b[0] == b[1] == k for a, b in zip(spans, spans[1:]))


def test_extra_entry_points_refuse_cpu_tensors():
    """forward_host / soft_forward / usage_code_probs / kmeans / graph replays: no CPU path behind any of them."""
    q = vq.VectorQuantizerEMA(32, 16, print_init=False)
    with pytest.raises(RuntimeError):
        q.forward_host(torch.randn(2, 4, 16))                 # module buffers are not on a CUDA device
    with pytest.raises(RuntimeError):
        q.forward_host(torch.randn(2, 4, 16).double())
    with pytest.raises(ValueError):
        q.forward_host(torch.randn(2, 4, 16), outputs="everything")
    with pytest.raises(ValueError):
        q.forward_host(torch.randn(8, 16))                    # non-3-D, like forward()
    with pytest.raises(RuntimeError):
        q.soft_forward(torch.randn(2, 4, 16), 1.0)
    with pytest.raises(RuntimeError):
        q.usage_code_probs(torch.randn(2, 4, 16))
    with pytest.raises(RuntimeError):
        vq.kmeans_fit(torch.randn(64, 16), 8)
    with pytest.raises(RuntimeError):
        vq.rvq_kmeans_fit(torch.randn(64, 16), 8, 2)
    with pytest.raises(RuntimeError):
        vq.GraphedForward(q.eval(), torch.randn(2, 4, 16))
    with pytest.raises(RuntimeError):
        vq.GraphedTrainStep(q.eval(), torch.randn(2, 4, 16))  # needs training mode (and CUDA)
    with pytest.raises(RuntimeError):
        vq.ops.rvq_finalize(torch.randn(4, 16), torch.zeros(4, dtype=torch.int64), 4, 1, torch.randn(8, 16))
    # round-2 entry points: tiled row softmax, indices -> decoder memory, the mirror model's decode_indices
    with pytest.raises(RuntimeError):
        vq.ops.softmax_rows(torch.randn(4, 16), torch.randn(8, 16), 1.0)
    with pytest.raises(RuntimeError):
        vq.ops.soft_assign(torch.randn(4, 16), torch.randn(8, 16), 1.0)
    with pytest.raises(RuntimeError):
        vq.ops.indices_to_memory(torch.zeros(4, dtype=torch.int64), torch.randn(8, 16), 1)
    from pytorch_vae_b200.vqvae import VQVAE
    m = VQVAE(hidden_dim=32, code_dim=16, codebook_size=8, latent_tokens=4, max_seq_len=8, num_layers=1, num_heads=2,
              tokenizer_layers=1, tokenizer_heads=2, print_init=False)
    assert m.projected_codebook().shape == (8, 32)              # the table itself is host-side algebra ...
    with pytest.raises(RuntimeError):
        m.memory_from_indices(torch.zeros(1, 4, dtype=torch.int64))   # ... the gather + LayerNorm kernel is not


def test_persistent_kernel_shape_rules_without_a_gpu():
    """Which residual shapes run as ONE persistent kernel (csrc/vq_rvq_fused.cu) and what a training forward launches."""
    lib = vq._cabi.lib
    assert lib.vqb200_rvq_fused_supported(8192, 1024, 512, 4, 0) == 1
    assert lib.vqb200_rvq_fused_supported(65536, 1024, 512, 4, 0) == 1 and lib.vqb200_rvq_fused_supported(65537, 1024, 512, 4, 0) == 0
    assert lib.vqb200_rvq_fused_supported(49152, 1024, 256, 4, 0) == 1 and lib.vqb200_rvq_fused_supported(65536, 1024, 256, 4, 0) == 0
    assert lib.vqb200_rvq_fused_supported(8192, 1024, 64, 4, 0) == 0            # D = 64: level pipeline
    assert lib.vqb200_rvq_fused_supported(8192, 64, 512, 4, 0) == 0             # fewer codes than one tile
    assert lib.vqb200_rvq_fused_supported(8192, 1024, 512, 1, 0) == 0 and lib.vqb200_rvq_fused_supported(8192, 1024, 512, 9, 0) == 0
    assert lib.vqb200_rvq_train_fused_supported(8192, 1024, 512, 4, 0) == 1
    assert lib.vqb200_rvq_train_launches(8192, 1024, 512, 4, 0) == 3            # refresh, persistent kernel, refresh
    assert lib.vqb200_rvq_train_workspace_bytes(8192, 1024, 512, 4, 0) >= 4096 * 512 * 4 + 4096 * 4 + 128 * 64 * 512 * 4
    assert lib.vqb200_stats_exchange_buffer_bytes(4096, 8) == 512 + 17 * 4098 * 8
    assert lib.vqb200_softmax_rows_workspace_bytes(8192, 1024) >= 8192 * 8


def test_launch_and_workspace_accounting_without_a_gpu():
    """The size / count queries of the one-call entry points are pure host functions."""
    lib = vq._cabi.lib
    N, K, D, L = 8192, 1024, 512, 4
    per = lib.vqb200_search_launches(N, K, D, 0)
    assert per == 5 and lib.vqb200_search_path(N, K, D, 0) == 1
    # the stage-2 shape runs as ONE persistent kernel; a shape it does not take (D = 64) goes level by level
    assert lib.vqb200_rvq_fused_supported(N, K, D, L, 0) == 1 and lib.vqb200_rvq_forward_launches(N, K, D, L, 0) == 1
    assert lib.vqb200_rvq_forward_workspace_bytes(N, K, D, L, 0) >= 64 * 128 * D * 4
    D = 64
    per = lib.vqb200_search_launches(N, K, D, 0)
    assert lib.vqb200_rvq_fused_supported(N, K, D, L, 0) == 0
    assert lib.vqb200_rvq_forward_launches(N, K, D, L, 0) == per + (L - 1) * (per - 1) + (L - 1) + 1
    assert lib.vqb200_rvq_train_launches(N, K, D, L, 0) == L * (per + 3) + 1
    ws = lib.vqb200_search_workspace_bytes(N, K, D, 0)
    assert lib.vqb200_rvq_forward_workspace_bytes(N, K, D, L, 0) >= ws + 2 * N * D * 4 + N * D * 2 + N * 4
    assert lib.vqb200_rvq_train_workspace_bytes(N, K, D, L, 0) >= ws + 2 * N * D * 4 + (K * L * D + K * L) * 4
    assert lib.vqb200_quantize_fused_supported(1 << 20, 512, 64, 0) == 1
    assert lib.vqb200_quantize_fused_supported(1 << 20, 512, 128, 0) == 1
    assert lib.vqb200_quantize_fused_supported(1 << 20, 512, 256, 0) == 0
    assert lib.vqb200_rvq_forward_launches(0, K, D, L, 0) == 0
