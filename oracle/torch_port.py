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


def train_step(z: torch.Tensor, state: dict, K_per: int, L: int, beta: float, decay: float, eps: float,
               g_st: torch.Tensor):
    """One training step of the path on the host cores, as the reference executes it (models/vq_vae.py:226-263
    with the EMA update of :77-89 after every level, the commitment term of :1292-1294, then autograd back to
    z): ``state`` holds E / ema_cluster_size / ema_embedding and is updated in place.  ``g_st`` is the
    gradient a decoder would feed into z_q_st.  Returns (indices [L*N], grad_z [N, D], commit)."""
    E, cs, es = state["E"], state["ema_cluster_size"], state["ema_embedding"]
    K_total = L * K_per
    z = z.detach().requires_grad_(True)
    residual, picks, codes = z, [], []
    for lvl in range(L):
        cb = E[lvl * K_per:(lvl + 1) * K_per]
        d = residual.pow(2).sum(1, keepdim=True) - 2.0 * (residual @ cb.t()) + cb.pow(2).sum(1, keepdim=True).t()
        pick = d.argmin(1)
        gid = pick + lvl * K_per
        picks.append(gid)
        code = torch.nn.functional.embedding(pick, cb)
        codes.append(code)
        with torch.no_grad():                                 # dense one-hot segment sums, as the reference forms them
            hot = torch.nn.functional.one_hot(gid, num_classes=K_total).float()
            cs.mul_(decay).add_(hot.sum(0) * (1 - decay))
            es.mul_(decay).add_(hot.t() @ residual.detach() * (1 - decay))
            E.copy_(es / (cs.unsqueeze(1) + eps))
        residual = residual - code
    zq = torch.stack(codes, 0).sum(0)
    st = z + (zq - z).detach()
    commit = torch.nn.functional.mse_loss(zq.detach(), z)
    torch.autograd.backward([st, beta * commit], [g_st, torch.ones(())])
    return torch.cat(picks, 0), z.grad, commit.detach()
