import os, sys, time, cProfile, pstats, io
import torch
sys.path.insert(0, os.getcwd())
import pytorch_vae_b200 as vq
dev = torch.device("cuda:0")
K, D, L, N = 1024, 512, 4, 8192
q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).train()
q.ema_embedding.copy_(q.embedding); q.ema_cluster_size.fill_(1.0)
z = torch.randn(N // 64, 64, D, device=dev)
g_st = torch.randn_like(z); beta_t = torch.full((), 0.0005, device=dev)
def step():
    ze = z.detach().requires_grad_(True)
    st, zq, idx, stats = q(ze, do_ema_update=True)
    torch.autograd.backward([st, q.last_commit], [g_st, beta_t])
for _ in range(20): step()
torch.cuda.synchronize()
tf = tb = 0.0
for _ in range(200):
    ze = z.detach().requires_grad_(True)
    t0 = time.perf_counter()
    st, zq, idx, stats = q(ze, do_ema_update=True)
    t1 = time.perf_counter()
    torch.autograd.backward([st, q.last_commit], [g_st, beta_t])
    t2 = time.perf_counter()
    tf += t1 - t0; tb += t2 - t1
    if _ % 20 == 19: torch.cuda.synchronize()
print(f"host time per step: forward {tf/200*1e6:.1f} us, backward {tb/200*1e6:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
