"""CUDA-graph replay of the quantizer forward for launch-bound shapes.

The stage-2 residual shape (4 levels x K=1024, D=512, N=8192 rows) needs ~26 kernels of 5-20 us each per
forward: the step is bound by host launch overhead, not by the GPU.  Capturing the whole eval-mode forward
(cache refresh included, so external codebook writes stay visible) into one CUDA graph turns it into a single
launch.  Shapes are static: feed tensors of the captured shape; outputs are the graph's static buffers
(clone them if they must survive the next replay).
"""
from __future__ import annotations

import torch

from . import ops


class GraphedForward:
    """``g = GraphedForward(q, z_example); z_q_st, z_q, idx, stats = g(z)`` (eval mode, no EMA update)."""

    def __init__(self, quantizer, z_example: torch.Tensor, mask=None):
        if quantizer.training:
            raise RuntimeError("GraphedForward captures the eval-mode forward (no EMA update, no re-init)")
        if not z_example.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA tensor")
        self.q = quantizer
        self.z = z_example.detach().clone()
        self.mask = None if mask is None else mask.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                                  # warm-up: lazy init (attributes, helper streams)
                quantizer(self.z, do_ema_update=False, mask=self.mask)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            if quantizer._cache is not None:
                quantizer._cache.key = None                     # capture the cache refresh: the graph re-derives it
            self.out = quantizer(self.z, do_ema_update=False, mask=self.mask)
        self.kernels = ops.launch_count() - before             # kernels of this library inside one replay

    def __call__(self, z_e: torch.Tensor):
        if z_e.shape != self.z.shape:
            raise RuntimeError(f"captured shape {tuple(self.z.shape)}, got {tuple(z_e.shape)}")
        self.z.copy_(z_e)
        self.graph.replay()
        ops._count(self.kernels)
        return self.out


class GraphedTrainStep:
    """One CUDA graph for the quantizer's whole training step: forward with the EMA codebook update, and the
    backward of its two differentiable outputs (straight-through + commitment).

    The stage-2 training step (4 levels x K=1024, D=512, N=8192 rows per GPU) is ~35 kernels of 5-30 us behind as
    many Python / ctypes calls: host overhead bounds it, not the GPU.  ``step(z_e, grad_st, beta)`` copies the
    latents and the upstream gradient into the graph's static buffers, replays, and returns
    ``(z_q_st, z_q, indices, stats, commit, grad_z)`` (static buffers: clone what must survive the next replay).
    ``grad_st`` is dLoss/dz_q_st as the decoder's backward produces it; the loss term is ``beta * commit``
    (models/vq_vae.py:1292-1294).  The codebook cache refresh is captured too, so external codebook writes
    (dead-code re-init, ``load_state_dict``) stay visible.  Masks are not supported (the mask path syncs the host)."""

    def __init__(self, quantizer, z_example: torch.Tensor, do_ema_update: bool = True):
        if not quantizer.training:
            raise RuntimeError("GraphedTrainStep captures the training-mode step; call quantizer.train() first")
        if not z_example.is_cuda:
            raise RuntimeError("GraphedTrainStep needs a CUDA tensor")
        self.q = quantizer
        self.do_ema = bool(do_ema_update)
        self.z = z_example.detach().clone().requires_grad_(True)
        self.g_st = torch.zeros_like(self.z)
        self.beta = torch.zeros((), device=self.z.device)
        # the EMA buffers move during warm-up and capture: put them back afterwards
        saved = {k: v.detach().clone() for k, v in quantizer.state_dict().items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step_eager()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        self.z.grad = None
        with torch.cuda.graph(self.graph):
            if quantizer._cache is not None:
                quantizer._cache.key = None
            self.out = self._step_eager()
        self.kernels = ops.launch_count() - before
        quantizer.load_state_dict(saved)
        if quantizer._cache is not None:
            quantizer._cache.key = None

    def _step_eager(self):
        self.z.grad = None
        st, zq, idx, stats = self.q(self.z, do_ema_update=self.do_ema)
        commit = self.q.last_commit
        torch.autograd.backward([st, commit], [self.g_st, self.beta])
        return st, zq, idx, stats, commit, self.z.grad

    def __call__(self, z_e: torch.Tensor, grad_st: torch.Tensor, beta: float):
        if z_e.shape != self.z.shape or grad_st.shape != self.z.shape:
            raise RuntimeError(f"captured shape {tuple(self.z.shape)}, got {tuple(z_e.shape)} / {tuple(grad_st.shape)}")
        with torch.no_grad():
            self.z.copy_(z_e)
            self.g_st.copy_(grad_st)
            self.beta.fill_(float(beta))
        self.graph.replay()
        ops._count(self.kernels)
        return self.out
