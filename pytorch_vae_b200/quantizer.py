"""B200-native drop-in for the reference's EMA vector quantizer.

Mirrors the public surface of ``models/vq_vae.py:19-282`` (class
``VectorQuantizerEMA``): constructor keywords, the nine registered buffers
(so reference checkpoints load with ``strict=True``), ``forward`` returning
``(z_q_st, z_q, indices, stats)``, ``_ema_update``, ``_maybe_reinit_dead_codes``,
``reset_epoch_stats`` / ``get_epoch_stats`` / ``get_embedding_snapshot`` and the
attributes callers poke (``K, K_per, D, num_quantizers, beta, decay, eps,
embedding``).  All arithmetic runs in libvqb200.so (hand-written sm_100a
kernels); this file is host-side sequencing only.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _cabi, ops, sharding

Tensor = torch.Tensor

_MODES = {"fp32": _cabi.MODE_FP32_EXACT, "bf16_input": _cabi.MODE_BF16_INPUT}


class _QuantizeFn(torch.autograd.Function):
    """Forward of the whole quantizer; backward of its two differentiable outputs.

    Outputs: z_q_st (grad -> identity to z_e), z_q, indices, stats
    (non-differentiable) and the commitment mse ``mean((z_q - z_e)^2)`` whose gradient
    w.r.t. z_e is ``2 (z_e - z_q) / (N D)`` (models/vq_vae.py:1293).
    """

    @staticmethod
    def forward(ctx, z_e: Tensor, q: "VectorQuantizerEMA", do_ema: bool, mask: Optional[Tensor]):
        z_q_st, z_q, indices, stats3 = q._run(z_e, do_ema, mask)
        stats = stats3[:2]
        commit = stats3[2]
        ctx.save_for_backward(z_e, z_q)
        ctx.set_materialize_grads(False)     # an unused output arrives as None, not as a zero tensor: the pure
                                             # straight-through backward (F.mse_loss on the reference side) launches nothing
        ctx.mark_non_differentiable(z_q, indices, stats)
        return z_q_st, z_q, indices, stats, commit

    @staticmethod
    def backward(ctx, g_st, _g_zq, _g_idx, _g_stats, g_commit):
        z_e, z_q = ctx.saved_tensors
        if g_st is None and g_commit is None:
            return None, None, None, None
        if g_commit is None:
            return g_st, None, None, None                     # pure straight-through: no kernel at all
        z = z_e.detach().contiguous()
        if z.numel() == 0:
            return torch.zeros_like(z_e), None, None, None
        out = torch.empty_like(z)
        gst = None if g_st is None else g_st.contiguous()
        gc = g_commit.to(torch.float32).contiguous()
        ops.commit_backward(gst, gc, z, z_q.contiguous(), 2.0 / z.numel(), out)
        return out.view_as(z_e), None, None, None


class VectorQuantizerEMA(nn.Module):
    """Nearest-code quantizer with EMA codebook; single level or residual (``num_quantizers > 1``).

    ``num_embeddings`` is per level, as in the reference (models/vq_vae.py:37-38).
    Extra keyword (not in the reference): ``search_mode`` = ``"fp32"`` (reference-exact
    indices) or ``"bf16_input"`` (inputs rounded to bf16, fp32 products and sums).
    """

    def __init__(
        self,
        num_embeddings: int,
        embedding_dim: int,
        beta: float = 0.25,
        decay: float = 0.98,
        eps: float = 1e-5,
        reinit_dead_codes: bool = True,
        reinit_prob: float = 1.0,
        dead_usage_threshold: int = 0,
        print_init: bool = True,
        diag_qe_cap: float = 10.0,
        diag_qe_bins: int = 64,
        num_quantizers: int = 1,
        search_mode: str = "fp32",
    ):
        super().__init__()
        self.num_quantizers = int(num_quantizers)
        self.K_per = int(num_embeddings)
        self.K = self.num_quantizers * self.K_per
        self.D = int(embedding_dim)
        if self.D % 4 != 0:
            raise ValueError(f"embedding_dim must be a multiple of 4 for the 128-bit row kernels, got {self.D}")
        if not 1 <= self.num_quantizers <= _cabi.MAX_LEVELS:
            raise ValueError(f"num_quantizers must be in [1, {_cabi.MAX_LEVELS}]")
        if search_mode not in _MODES:
            raise ValueError(f"search_mode must be one of {sorted(_MODES)}")
        self.search_mode = search_mode

        self.beta = float(beta)
        self.decay = float(decay)
        self.eps = float(eps)
        self.use_ema = True
        self.reinit_dead_codes = bool(reinit_dead_codes)
        self.reinit_prob = float(reinit_prob)
        self.dead_usage_threshold = int(dead_usage_threshold)
        self.diag_qe_cap = float(diag_qe_cap)
        self.diag_qe_bins = int(diag_qe_bins)
        # multi-GPU policy: "local" = reference behaviour (each rank
        # updates from its own shard), "allreduce" = segment sums summed over ranks before the EMA.
        self.ema_sync = "local"
        self.stats_sync = False
        # how the statistics of stats_sync travel: "peer" = one kernel over NVLink peer memory (symmetric buffers; falls
        # back to NCCL when they cannot be set up), "nccl" = pack kernel -> all-reduce -> finalize kernel
        self.stats_exchange = os.environ.get("VQB200_STATS_EXCHANGE", "peer")

        # same nine buffers, same order, same init law as models/vq_vae.py:50-62
        self.register_buffer("embedding", torch.randn(self.K, self.D) * (1.0 / math.sqrt(self.D)))
        self.register_buffer("ema_cluster_size", torch.zeros(self.K))
        self.register_buffer("ema_embedding", torch.zeros(self.K, self.D))
        self.register_buffer("_ep_usage", torch.zeros(self.K))
        self.register_buffer("_ep_top1_sum", torch.zeros(1))
        self.register_buffer("_ep_top2_sum", torch.zeros(1))
        self.register_buffer("_ep_cnt", torch.zeros(1))
        self.register_buffer("_ep_qe_sum", torch.zeros(1))
        self.register_buffer("_ep_qe_hist", torch.zeros(self.diag_qe_bins))

        self._cache: Optional[ops.CodebookCache] = None
        self.last_commit: Optional[Tensor] = None      # autograd-connected mean((z_q - z_e)^2) of the last forward
        self._last_pair = None

        if print_init:
            kind = (f"[RVQ] EMA (L2): L={self.num_quantizers}, K_per={self.K_per}, K_total={self.K}"
                    if self.num_quantizers > 1 else f"[VQ] EMA (L2): K={self.K}")
            print(f"{kind}, D={self.D}, beta={self.beta}, decay={self.decay} [libvqb200 sm_100a]")

    # ------------------------------------------------------------------ cache
    def _codebook_cache(self) -> ops.CodebookCache:
        """``embedding`` is the source of truth: it is mutated from outside at any time (load_state_dict,
        init_codebook_from_centroids, dead-code re-init), so the derived cache is keyed on its storage
        and version counter and rebuilt whenever either moves."""
        E = self.embedding
        if E.dtype != torch.float32 or not E.is_contiguous():
            raise RuntimeError("quantizer.embedding must be a contiguous float32 tensor")
        c = self._cache
        if c is None or c.E_bf16.device != E.device:
            c = self._cache = ops.CodebookCache(self.K, self.D, self.K_per, E.device)
        key = (E.data_ptr(), E._version)
        if c.key != key:
            c.prepare(E)
            c.key = key
        return c

    # ------------------------------------------------------------------ EMA
    @torch.no_grad()
    def _ema_update(self, flat_raw: Tensor, indices: Tensor, row_mask: Optional[Tensor] = None):
        """models/vq_vae.py:77-89 without the dense one-hot: segment sums by scatter-add, then the
        lerp / divide for ALL codes fused with the cache refresh."""
        if flat_raw.numel() == 0 or indices.numel() == 0:
            return
        cache = self._codebook_cache()
        flat = flat_raw.detach().reshape(-1, self.D).contiguous()
        idx = indices.detach().reshape(-1).contiguous()
        seg = torch.zeros(self.K * self.D + self.K, dtype=torch.float32, device=flat.device)
        seg_sum, seg_cnt = seg[: self.K * self.D], seg[self.K * self.D:]
        ops.scatter_add(flat, idx, row_mask, seg_sum, seg_cnt)
        if self.ema_sync == "allreduce" and sharding.dist_ready():
            # ONE call over the whole buffer: only this level's slice is non-zero, but slicing it costs a second
            # NCCL call (sums and counts are not adjacent) and this path is host-bound (measured at 2 GPUs,
            # stage-2 shape: 0.95 ms per step with one 8 MB call per level, 1.13 ms with two small ones)
            torch.distributed.all_reduce(seg)
        ops.ema_finalize(seg_sum, seg_cnt, self.decay, self.eps, self.ema_cluster_size, self.ema_embedding,
                         self.embedding, cache)
        cache.key = (self.embedding.data_ptr(), self.embedding._version)

    @torch.no_grad()
    def _maybe_reinit_dead_codes(self, flat_raw: Tensor, usage: Tensor):
        """models/vq_vae.py:91-107: rare (every 500 steps), host RNG + one host sync, stays in PyTorch."""
        if not self.reinit_dead_codes or self.reinit_prob <= 0.0:
            return
        dead = usage <= float(self.dead_usage_threshold)
        n_dead = int(dead.sum().item())
        if n_dead <= 0 or flat_raw.numel() == 0:
            return
        if torch.rand(()) > self.reinit_prob:
            return
        rows = torch.randint(0, flat_raw.size(0), (n_dead,), device=flat_raw.device)
        fresh = flat_raw[rows]
        self.embedding[dead] = fresh
        self.ema_embedding[dead] = fresh.clone()
        self.ema_cluster_size[dead] = 1.0

    # ------------------------------------------------------------------ epoch statistics
    @torch.no_grad()
    def reset_epoch_stats(self):
        for name in ("_ep_usage", "_ep_top1_sum", "_ep_top2_sum", "_ep_cnt", "_ep_qe_sum", "_ep_qe_hist"):
            getattr(self, name).zero_()

    @torch.no_grad()
    def get_epoch_stats(self) -> dict:
        """Same keys and formulas as models/vq_vae.py:118-164 (margin/qe fields are never written by
        the reference either, so they stay 0)."""
        usage = self._ep_usage.detach().cpu()
        cnt = float(self._ep_cnt.item())
        out = {"usage_hist": usage, "margin_mean": 0.0, "qe_mean": 0.0, "qe_p90": 0.0, "n_positions": 0,
               "perplexity": 0.0, "dead_ratio": 0.0}
        if cnt <= 0:
            return out
        out["n_positions"] = int(cnt)
        out["margin_mean"] = float(((self._ep_top1_sum - self._ep_top2_sum) / cnt).item())
        out["qe_mean"] = float((self._ep_qe_sum / cnt).item())
        total = float(usage.sum().item())
        if total > 0:
            p = (usage / max(total, 1e-12)).clamp_min(1e-12)
            out["perplexity"] = float(torch.exp(-(p * p.log()).sum()).item())
            out["dead_ratio"] = float((usage == 0).float().mean().item())
        hist = self._ep_qe_hist.detach().cpu()
        mass = float(hist.sum().item())
        if mass > 0:
            cdf = torch.cumsum(hist, 0) / max(mass, 1e-12)
            hit = (cdf >= 0.9).nonzero(as_tuple=True)[0]
            b = int(hit[0].item()) if hit.numel() else self.diag_qe_bins - 1
            out["qe_p90"] = float((b + 0.5) * self.diag_qe_cap / max(self.diag_qe_bins, 1))
        return out

    @torch.no_grad()
    def get_embedding_snapshot(self) -> Tensor:
        return self.embedding.detach().clone()

    # ------------------------------------------------------------------ forward
    def forward(self, z_e: Tensor, do_ema_update: bool = True, allow_reinit: bool = True,
                mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        B, M, D = z_e.shape                                    # non-3-D input -> ValueError, like the reference
        if not z_e.is_cuda:
            raise RuntimeError("VectorQuantizerEMA (libvqb200) needs CUDA tensors on an sm_100a device; "
                               "there is no CPU fallback")
        if z_e.dtype != torch.float32:
            raise RuntimeError(f"expected z_e of dtype float32 (the codebook's dtype), got {z_e.dtype}")
        if D != self.D:
            raise RuntimeError(f"z_e has last dim {D}, the codebook has D={self.D}")
        do_ema = bool(self.training and do_ema_update)
        z_q_st, z_q, indices, stats, commit = _QuantizeFn.apply(z_e, self, do_ema, mask)
        # plain attributes, set past nn.Module.__setattr__ (its parameter / buffer / module bookkeeping costs ~5 us per
        # assignment: a tenth of the host time of a stage-2 step)
        self.__dict__["last_commit"] = commit
        self.__dict__["_last_pair"] = (weakref.ref(z_q), weakref.ref(z_e))   # identity check only: keeps no graph alive
        return z_q_st, z_q, indices, stats

    # transient per-forward / per-device state: autograd-attached tensors, CUDA streams and events, the derived
    # codebook cache.  Never part of a copy or a pickle (copy.deepcopy(model) after a training forward would
    # otherwise fail on the non-leaf ``last_commit``); all of it is rebuilt on the next forward.
    _TRANSIENT = ("last_commit", "_last_pair", "_hpipe", "_cache")

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in self._TRANSIENT:
            if k in state:
                state[k] = None
        return state

    @torch.no_grad()
    def soft_forward(self, z_e: Tensor, tau: float, do_ema_update: bool = True):
        """The quantizer's share of the soft-VQ training branch (models/vq_vae.py:828-861, single level):
        ``z_soft = softmax(-|z - e|^2 / tau) @ E`` (one online-softmax kernel), the hard assignment and
        ``z_q_hard``, the EMA update from the hard assignment, and ``[perplexity, dead_ratio]`` of this batch.
        Everything is detached in the reference, so nothing here is differentiable; the epoch accumulators are
        left alone, as the reference's soft branch never calls ``forward``.  Returns
        ``(z_soft, z_q_hard, indices [B, M], stats)``; z_soft / z_q_hard use the codebook BEFORE the update."""
        if self.num_quantizers != 1:
            raise RuntimeError("soft VQ is defined for single-level codebooks (models/vq_vae.py:828)")
        B, M, D = z_e.shape
        if not z_e.is_cuda or z_e.dtype != torch.float32 or D != self.D:
            raise RuntimeError("soft_forward needs float32 CUDA latents [B, M, D] matching the codebook")
        dev = z_e.device
        flat = z_e.detach().reshape(-1, D).contiguous()
        N = flat.shape[0]
        z_soft = ops.soft_assign(flat, self.embedding, tau)
        scratch = torch.zeros(4 + self.K, dtype=torch.int32, device=dev)
        sqerr, hist = scratch[:4].view(torch.float64), scratch[4:]
        z_q = torch.empty(N, D, dtype=torch.float32, device=dev)
        idx = torch.empty(N, dtype=torch.int64, device=dev)
        do_ema = bool(self.training and do_ema_update and N > 0)
        if N > 0:
            self._quantize_rows(flat, [idx], z_q, None, sqerr, hist, None, do_ema, do_ema)
        stats3 = torch.empty(3, dtype=torch.float32, device=dev)
        spare = torch.zeros(self.K + 1, dtype=torch.float32, device=dev)      # stand-ins for the epoch accumulators
        ops.stats_finalize(hist, float(N), sqerr, 1.0 / max(N * D, 1), spare[: self.K], spare[self.K:], stats3)
        return z_soft.view(B, M, D), z_q.view(B, M, D), idx.view(B, M), stats3[:2]

    def usage_code_probs(self, z_e: Tensor) -> Tensor:
        """``softmax(z_e @ E^T).mean(0)`` [K_total] -- the soft code-usage distribution of the usage-entropy
        regulariser (models/vq_vae.py:1298-1309), differentiable w.r.t. ``z_e`` (the codebook is detached there);
        no [N, K] matrix is formed in either direction."""
        if not z_e.is_cuda or z_e.dtype != torch.float32 or z_e.shape[-1] != self.D:
            raise RuntimeError("usage_code_probs needs float32 CUDA latents whose last dim matches the codebook")
        return _UsageProbsFn.apply(z_e.reshape(-1, self.D), self.embedding)

    def commitment_loss(self, z_q: Tensor, z_e: Tensor) -> Tensor:
        """``F.mse_loss(z_q.detach(), z_e)`` (models/vq_vae.py:1293).  When called with the pair the
        last forward produced, returns the value the gather pass already accumulated (its backward is
        the fused commit_backward kernel); otherwise runs the fused st_loss kernel on the pair."""
        if self._last_pair is not None and z_q is self._last_pair[0]() and z_e is self._last_pair[1]():
            return self.last_commit
        return _CommitFn.apply(z_e, z_q.detach())

    # the actual work; runs under no_grad inside _QuantizeFn.forward
    def _run(self, z_e: Tensor, do_ema: bool, mask: Optional[Tensor]):
        B, M, D = z_e.shape
        dev = z_e.device
        flat = z_e.detach().reshape(-1, D)
        if not flat.is_contiguous():
            flat = flat.contiguous()
        N = flat.shape[0]
        L = self.num_quantizers

        # one zeroed scratch: [sqerr_sum (double) | hist int32[K]]
        scratch = torch.zeros(4 + self.K, dtype=torch.int32, device=dev)
        sqerr = scratch[:4].view(torch.float64)               # [sum sq err, element count (filled only when syncing)]
        hist = scratch[4:]
        stats3 = torch.empty(3, dtype=torch.float32, device=dev)
        z_q = torch.empty(N, D, dtype=torch.float32, device=dev)
        z_q_st = torch.empty(N, D, dtype=torch.float32, device=dev)

        valid_u8 = None
        if mask is not None:
            valid_u8 = mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
        # the reference skips the EMA entirely when no row is valid (:196,253): same host sync, mask only
        ema_ok = do_ema and N > 0 and (valid_u8 is None or bool(valid_u8.any()))

        idx_all = torch.empty(L * N, dtype=torch.int64, device=dev)   # RVQ: level-major, global ids (:260)
        # single-process statistics can ride in the library call that runs the levels (the persistent kernel's last CTA)
        stats_req = None
        if L > 1 and N > 0 and not (self.stats_sync and sharding.dist_ready()):
            stats_req = {"args": (float(L * N), 1.0 / max(N * D, 1), self._ep_usage, self._ep_cnt, stats3), "done": False}
        if N > 0:
            self._quantize_rows(flat, [idx_all[l * N:(l + 1) * N] for l in range(L)], z_q, z_q_st, sqerr, hist,
                                valid_u8, do_ema, ema_ok, stats_req)
        if stats_req is None or not stats_req["done"]:
            self._finalize_stats(hist, float(L * N), sqerr, N * D, stats3)
        return z_q_st.view(B, M, D), z_q.view(B, M, D), (idx_all.view(B, M) if L == 1 else idx_all), stats3

    def _quantize_rows(self, flat, idx_levels, z_q, z_q_st, sqerr, hist, valid_u8, do_ema=False, ema_ok=False,
                       stats_req=None):
        """Search + gather (+ EMA) of the rows ``flat`` [n, D] into caller-allocated outputs; ``sqerr`` and
        ``hist`` ACCUMULATE, so a batch may be fed in several calls (``forward_host``).  ``z_q`` / ``z_q_st``
        may be None for a single-level codebook (codes only)."""
        n, D = flat.shape
        mode = _MODES[self.search_mode]
        cache = self._codebook_cache()
        E = self.embedding
        L = self.num_quantizers
        if L == 1:
            idx = idx_levels[0]
            if ops.fused_supported(n, self.K, D, mode):      # one kernel: search + gather + loss + histogram
                ops.quantize_fused(flat, E, cache, mode, idx, zq_out=z_q, zq_st_out=z_q_st,
                                   sqerr_sum=sqerr if z_q is not None else None, hist=hist, row_mask=valid_u8)
            elif os.environ.get("VQB200_SPLIT_GATHER") == "1":   # measurement switch: gather as a separate pass
                ops.search(flat, E, cache, 0, mode, idx)
                ops.gather(flat, E, idx, zq_out=z_q, zq_st_out=z_q_st, sqerr_sum=sqerr, hist=hist, row_mask=valid_u8)
            else:                                            # search + gather, chunk-pipelined on the tensor path
                ops.quantize(flat, E, cache, mode, idx, zq_out=z_q, zq_st_out=z_q_st,
                             sqerr_sum=sqerr if z_q is not None else None,
                             hist=hist, row_mask=valid_u8)   # z_q is gathered BEFORE the EMA mutates E (:189 -> :193)
            if ema_ok:
                self._ema_update(flat, idx, valid_u8)
            return
        residual = flat
        class _Spare:                                       # residual ping-pong buffers of the level-by-level paths,
            bufs = None                                     # allocated only when one of them runs

            def __getitem__(self, i):
                if self.bufs is None:
                    self.bufs = [torch.empty(n, D, dtype=torch.float32, device=flat.device) for _ in range(min(2, L - 1))]
                return self.bufs[i]
        spare = _Spare()
        lstride = (idx_levels[1].data_ptr() - idx_levels[0].data_ptr()) // 8 if L > 1 else n
        even = all(idx_levels[l].data_ptr() == idx_levels[0].data_ptr() + 8 * l * lstride for l in range(L))
        if not do_ema and L <= 8 and even and lstride >= n:
            # the codebook does not move between levels: the levels only carry the residual forward, and ONE
            # pass at the end forms z_q (level-order sum), z_q_st, the loss partial sum and the histogram from
            # the indices -- z_q is not re-read and re-written on every level
            # ... and, on the tensor path, the residual update of a level and the pre-pass of the next (16-bit
            # operand copy + admission margins) are one kernel, one read of the rows.  With the ids packed
            # [L * n] the whole loop is ONE library call (at N = 8192 the separate calls cost more host time
            # than their kernels cost GPU time).
            if lstride == n:
                ops.rvq_forward(flat, E, cache, mode, idx_levels[0], zq_out=z_q, zq_st_out=z_q_st, sqerr_sum=sqerr,
                                hist=hist, stats=stats_req["args"] if stats_req else None)
                if stats_req:
                    stats_req["done"] = True
                return
            on_tc = bool(_cabi.lib.vqb200_search_path(n, self.K_per, D, mode)) and L > 1
            z16 = torch.empty(n, D, dtype=torch.bfloat16, device=flat.device) if on_tc else None
            mg = torch.empty(n, dtype=torch.float32, device=flat.device) if on_tc else None
            for level in range(L):
                if on_tc and level > 0:
                    ops.search_prepped(residual, z16, mg, E, cache, level, mode, idx_levels[level])
                else:
                    ops.search(residual, E, cache, level, mode, idx_levels[level])
                if level < L - 1:
                    if on_tc:
                        ops.residual_prep(residual, E, idx_levels[level], cache, level + 1, mode, spare[level % 2], z16, mg)
                    else:
                        ops.gather(residual, E, idx_levels[level], residual_out=spare[level % 2])
                    residual = spare[level % 2]
            ops.rvq_finalize(flat, idx_levels[0], lstride, L, E, zq_out=z_q, zq_st_out=z_q_st, sqerr_sum=sqerr,
                             hist=hist)
            return
        if ema_ok and valid_u8 is None and lstride == n and \
                not (self.ema_sync == "allreduce" and sharding.dist_ready()):
            # training, every rank updating from its own rows: all levels, their EMA updates and the
            # straight-through / loss pass in ONE library call
            ops.rvq_train_forward(flat, E, cache, mode, self.decay, self.eps, self.ema_cluster_size,
                                  self.ema_embedding, idx_levels[0], z_q, zq_st_out=z_q_st, sqerr_sum=sqerr, hist=hist,
                                  stats=stats_req["args"] if stats_req else None)
            if stats_req:
                stats_req["done"] = True
            cache.key = (self.embedding.data_ptr(), self.embedding._version)
            return
        if ema_ok and valid_u8 is None and n > 0 and self.ema_sync == "allreduce" and sharding.dist_ready():
            # training with the segment sums all-reduced over ranks: per level ONE library call up to the exchange
            # point (search, gather, scatter-add), the all-reduce, and the EMA finalize
            seg = torch.empty(self.K * D + self.K, dtype=torch.float32, device=flat.device)
            seg_sum, seg_cnt = seg[: self.K * D], seg[self.K * D:]
            if lstride == n and ops.rvq_train_fused_supported(n, self.K_per, D, L, mode):
                # ONE exchange per step: no level's search depends on this step's segment sums (include/vq_b200.h)
                ops.rvq_train_begin(flat, E, cache, mode, self.decay, self.eps, self.ema_cluster_size, self.ema_embedding,
                                    idx_levels[0], z_q, seg_sum, seg_cnt, zq_st_out=z_q_st, sqerr_sum=sqerr, hist=hist)
                torch.distributed.all_reduce(seg)
                ops.rvq_train_finish(seg_sum, seg_cnt, E, cache, self.decay, self.eps, self.ema_cluster_size,
                                     self.ema_embedding)
                cache.key = (self.embedding.data_ptr(), self.embedding._version)
                return
            for level in range(L):
                nxt = spare[level % 2] if level < L - 1 else None
                ops.rvq_train_level(residual, E, cache, level, mode, idx_levels[level], z_q, nxt, hist, seg_sum, seg_cnt)
                torch.distributed.all_reduce(seg)
                ops.ema_finalize(seg_sum, seg_cnt, self.decay, self.eps, self.ema_cluster_size, self.ema_embedding,
                                 self.embedding, cache)
                if nxt is not None:
                    residual = nxt
            cache.key = (self.embedding.data_ptr(), self.embedding._version)
            ops.st_loss(flat, z_q, zq_st_out=z_q_st, sqerr_sum=sqerr)
            return
        for level in range(L):
            idx_l = idx_levels[level]
            if level > 0 and do_ema:
                cache = self._codebook_cache()
            ops.search(residual, E, cache, level, mode, idx_l)
            nxt = spare[level % 2] if level < L - 1 else None
            # level sum in level order (:261); RVQ histogram ignores the mask (:266)
            ops.gather(residual, E, idx_l, zq_out=z_q, accumulate=level > 0, residual_out=nxt, hist=hist)
            if ema_ok:
                self._ema_update(residual, idx_l, valid_u8)
            if nxt is not None:
                residual = nxt
        ops.st_loss(flat, z_q, zq_st_out=z_q_st, sqerr_sum=sqerr)

    # ------------------------------------------------------------------ host-buffer entry (extraction path)
    @torch.no_grad()
    def forward_host(self, z_host: Tensor, chunk_rows: Optional[int] = None, outputs: str = "all",
                     out_indices: Optional[Tensor] = None, wait: bool = True, token_major=None):
        """Eval-mode forward over latents that live in (pinned) HOST memory -- the shape of
        ``scripts/extract_code_indices.py:296-320``, where every batch is copied to the device, quantized and
        its indices copied back.  The rows stream through a ring of device staging buffers: the H2D copy of
        chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1's indices run on three streams, so a
        call costs max(copy, compute) instead of their sum.

        Returns ``(z_q_st, z_q, indices_host, stats_host)``: ``z_q_st`` / ``z_q`` stay on the device
        (``None`` with ``outputs="indices"``, which also skips writing them), ``indices_host`` is a pinned
        int64 tensor laid out as ``forward`` lays indices out, ``stats_host`` a pinned float32 [2].  With
        ``token_major=torch.int32`` (or int16 / int64) the indices are re-laid out on the device to the
        token-major ``[B, M * Q]`` array of that dtype which ``scripts/extract_code_indices.py:195-209`` saves,
        chunk by chunk, and THAT is what is copied to the host.  With ``wait=False`` the host buffers are valid
        only after the current stream has been synchronised."""
        if outputs not in ("all", "indices"):
            raise ValueError("outputs must be 'all' or 'indices'")
        B, M, D = z_host.shape
        if z_host.is_cuda:
            raise RuntimeError("forward_host takes a host tensor; call forward() for device tensors")
        if z_host.dtype != torch.float32:
            raise RuntimeError(f"expected z_e of dtype float32 (the codebook's dtype), got {z_host.dtype}")
        if D != self.D:
            raise RuntimeError(f"z_e has last dim {D}, the codebook has D={self.D}")
        dev = self.embedding.device
        if dev.type != "cuda":
            raise RuntimeError("VectorQuantizerEMA (libvqb200) needs its buffers on an sm_100a device")
        flat_h = z_host.reshape(-1, D)
        if not flat_h.is_contiguous():
            flat_h = flat_h.contiguous()
        if not flat_h.is_pinned():
            flat_h = flat_h.pin_memory()                      # pageable memory would serialise the copies
        N, L = flat_h.shape[0], self.num_quantizers
        want_all = outputs == "all" or L > 1                  # the residual chain needs z_q anyway
        if chunk_rows is None:        # 2^17 rows: >= 32 MB per copy (full PCIe rate), several waves of the tensor
            chunk_rows = 1 << 17      # kernels per launch, and still dozens of chunks to pipeline at extraction sizes
        n_chunks = max(1, -(-N // chunk_rows))
        ring = min(3, n_chunks)

        main = torch.cuda.current_stream(dev)
        pipe = self._host_pipe(dev, ring)
        s_in, s_out = pipe["in"], pipe["out"]
        ev_in, ev_run = pipe["ev_in"], pipe["ev_run"]

        scratch = torch.zeros(4 + self.K, dtype=torch.int32, device=dev)
        sqerr, hist = scratch[:4].view(torch.float64), scratch[4:]
        stats3 = torch.empty(3, dtype=torch.float32, device=dev)
        # device staging buffers: persistent (one cudaMalloc per shape, not one block per call), and TWO sets when the
        # whole batch is one chunk, so that with wait=False the H2D copy of the next call overlaps the kernels of this one
        # (small-batch extraction: max(copy, kernels) per call instead of their sum)
        skey = (ring, min(chunk_rows, max(N, 1)), D, 2 if n_chunks == 1 else 1)
        if pipe.get("stage_key") != skey:
            pipe["stage_key"] = skey
            pipe["stage"] = [torch.empty(skey[0], skey[1], D, dtype=torch.float32, device=dev) for _ in range(skey[3])]
            pipe["ev_free"] = [torch.cuda.Event() for _ in range(skey[3])]
            pipe["calls"] = 0
        sset = pipe["calls"] % skey[3]
        pipe["calls"] += 1
        stage, ev_free = pipe["stage"][sset], pipe["ev_free"][sset]
        z_q = torch.empty(N, D, dtype=torch.float32, device=dev) if want_all else None
        z_q_st = torch.empty(N, D, dtype=torch.float32, device=dev) if want_all else None
        idt = torch.int64 if token_major is None else token_major
        idx_all = torch.empty(L * N, dtype=torch.int64, device=dev) if token_major is None else None
        tok_all = None if token_major is None else torch.empty(N * L, dtype=idt, device=dev)
        if out_indices is None:
            out_indices = torch.empty(L * N, dtype=idt).pin_memory()
        idx_h = out_indices.view(-1)
        if idx_h.numel() != L * N or idx_h.dtype != idt or idx_h.is_cuda:
            raise RuntimeError(f"out_indices must be a host {idt} tensor of {L * N} elements")
        stats_h = pipe["stats_h"]
        self._codebook_cache()                                # refresh on the caller's stream, before the fork
        fork = torch.cuda.Event()
        fork.record(main)
        s_in.wait_event(ev_free)                              # the call that last used this staging set has read it
        s_out.wait_event(fork)

        for c in range(n_chunks):
            r0, r1 = c * chunk_rows, min(N, (c + 1) * chunk_rows)
            n, b = r1 - r0, c % ring
            with torch.cuda.stream(s_in):
                if c >= ring:
                    s_in.wait_event(ev_run[b])                # the kernels of chunk c-ring are done with this buffer
                stage[b, :n].copy_(flat_h[r0:r1], non_blocking=True)
                ev_in[b].record(s_in)
            main.wait_event(ev_in[b])
            if tok_all is None:
                idx_levels = [idx_all[l * N + r0:l * N + r1] for l in range(L)]
            else:                                             # chunk-local level-major ids, re-laid out below
                idx_c = torch.empty(L * n, dtype=torch.int64, device=dev)
                idx_levels = [idx_c[l * n:(l + 1) * n] for l in range(L)]
            self._quantize_rows(stage[b, :n], idx_levels,
                                None if z_q is None else z_q[r0:r1], None if z_q_st is None else z_q_st[r0:r1],
                                sqerr, hist, None)
            if tok_all is not None:                           # rows r0:r1 of every level -> tokens [r0*L, r1*L)
                ops.relayout_indices(idx_c, L, 1, n, idt, out=tok_all[r0 * L:r1 * L])
            ev_run[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run[b])
                if tok_all is not None:
                    idx_h[r0 * L:r1 * L].copy_(tok_all[r0 * L:r1 * L], non_blocking=True)
                else:
                    for l in range(L):
                        idx_h[l * N + r0:l * N + r1].copy_(idx_all[l * N + r0:l * N + r1], non_blocking=True)
        ev_free.record(main)                                  # every kernel that reads the staging set is enqueued
        self._finalize_stats(hist, float(L * N), sqerr, N * D, stats3)
        stats_h.copy_(stats3[:2], non_blocking=True)
        done = torch.cuda.Event()
        done.record(s_out)
        main.wait_event(done)                                 # the caller's stream sees every copy finished
        if wait:
            main.synchronize()
        if token_major is not None:
            idx_ret = out_indices.view(B, M * L)
        else:
            idx_ret = out_indices.view(B, M) if L == 1 else out_indices.view(-1)
        if wait:
            stats_h = stats_h.clone()     # the pinned staging buffer is shared by every call on this device; with
                                          # wait=False the caller reads it after its own synchronise, before the next call
        if not want_all:
            return None, None, idx_ret, stats_h
        return z_q_st.view(B, M, D), z_q.view(B, M, D), idx_ret, stats_h

    def _host_pipe(self, dev, ring):
        p = getattr(self, "_hpipe", None)
        if p is None or p["dev"] != dev:
            p = self._hpipe = {"dev": dev, "in": torch.cuda.Stream(dev), "out": torch.cuda.Stream(dev),
                               "ev_in": [torch.cuda.Event() for _ in range(3)],
                               "ev_run": [torch.cuda.Event() for _ in range(3)],
                               "stats_h": torch.empty(2, dtype=torch.float32).pin_memory()}
        return p

    def _finalize_stats(self, hist, count_add, sqerr, n_elems, stats3):
        """``sqerr``: float64 tensor whose first word is this rank's sum of squared errors."""
        if self.stats_sync and sharding.dist_ready():
            # global statistics: ONE small all-reduce instead of the reference's per-rank
            # perplexities averaged by sync_dist.  pack kernel -> NCCL -> finalize on the reduced pack: three
            # launches on the step path (the 4 KB reduction sits inside a 0.4 ms step at the c2 shape)
            ex = sharding.PeerStatsExchange.get(hist.device, self.K) if self.stats_exchange == "peer" else None
            if ex is not None:                                  # ONE kernel over NVLink peer memory, no NCCL call
                ops.stats_exchange(hist, sqerr, n_elems, self.K, self.num_quantizers, self.D, ex.peer_ptrs_dev, ex.rank,
                                   ex.world, self._ep_usage, self._ep_cnt, stats3)
                return
            pack = torch.empty(self.K + 2, dtype=torch.float64, device=hist.device)
            ops.stats_pack(hist, sqerr, n_elems, pack)
            torch.distributed.all_reduce(pack)
            ops.stats_finalize_packed(pack, self.K, self.num_quantizers, self.D, self._ep_usage, self._ep_cnt, stats3)
            return
        ops.stats_finalize(hist, count_add, sqerr, 1.0 / max(n_elems, 1), self._ep_usage, self._ep_cnt, stats3)


class _UsageProbsFn(torch.autograd.Function):
    """p_code = softmax(z @ E^T).mean(0) with the codebook detached (models/vq_vae.py:1303-1307)."""

    @staticmethod
    def forward(ctx, z_flat, E):
        z = z_flat.detach().contiguous()
        Ed = E.detach().contiguous()
        need = bool(ctx.needs_input_grad[0])
        p_code, probs = ops.usage_probs(z, Ed, keep_probs=need)
        if need:
            ctx.save_for_backward(probs, Ed)
        return p_code

    @staticmethod
    def backward(ctx, g):
        probs, Ed = ctx.saved_tensors
        return ops.usage_probs_backward_from_probs(probs, Ed, g.to(torch.float32).contiguous()), None


class _CommitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_e, z_q):
        z = z_e.detach().contiguous()
        zq = z_q.contiguous()
        sq = torch.zeros(1, dtype=torch.float64, device=z.device)
        ops.st_loss(z, zq, None, sq)
        ctx.save_for_backward(z, zq)
        ctx.shape = z_e.shape
        return (sq / max(z.numel(), 1)).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        z, zq = ctx.saved_tensors
        out = torch.empty_like(z)
        ops.commit_backward(None, g.to(torch.float32).contiguous(), z, zq, 2.0 / z.numel(), out)
        return out.view(ctx.shape), None


# north-star name for the same class (BASELINE.json calls it VectorQuantizer)
VectorQuantizer = VectorQuantizerEMA
