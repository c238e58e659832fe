"""Multi-threaded CPU port of the reference forward, used ONLY as the timed CPU baseline
(bench.py cpu_baseline / --impl reference) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's CPU path is torch ATen on all host threads (MKL sgemm + elementwise kernels).
The numpy oracle (vq_oracle.py) is the parity checker but its elementwise passes are single-threaded,
which would understate the reference; this port issues the same ATen op sequence as
models/vq_vae.py:183-189 / :238-258 (eval mode, no EMA), row-chunked because the reference
materialises N x K.  tests/test_oracle_golden.py pins it to the golden vectors too.
"""
import torch


@torch.no_grad()
def forward_eval(z: torch.Tensor, E: torch.Tensor, K_per: int, L: int, chunk: int = 65536):
    """z [N, D], E [L*K_per, D] (CPU fp32) -> (z_q [N, D], indices [L*N] level-major global ids, usage [L*K_per])."""
    N, D = z.shape
    idx_out = torch.empty(L, N, dtype=torch.int64)
    zq_out = torch.empty_like(z)
    for s in range(0, N, chunk):
        r = z[s:s + chunk]
        total = None
        for lvl in range(L):
            cb = E[lvl * K_per:(lvl + 1) * K_per]
            d = r.pow(2).sum(1, keepdim=True) - 2.0 * (r @ cb.t()) + cb.pow(2).sum(1, keepdim=True).t()
            pick = d.argmin(1)
            code = cb.index_select(0, pick)
            idx_out[lvl, s:s + chunk] = pick + lvl * K_per
            total = code if total is None else total + code
            r = r - code
        zq_out[s:s + chunk] = total
    idx = idx_out.reshape(-1)
    usage = torch.bincount(idx, minlength=L * K_per).float()
    return zq_out, idx, usage
