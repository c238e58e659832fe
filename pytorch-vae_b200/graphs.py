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
