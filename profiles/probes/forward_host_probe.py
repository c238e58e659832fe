import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import pytorch_vae_b200 as vq
dev = torch.device("cuda:0")
K, D, L, N = 1024, 512, 4, 8192
q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).eval()
z_pin = torch.randn(N // 64, 64, D).pin_memory()
idx_host = torch.empty(L * N, dtype=torch.int64).pin_memory()
for rep in range(3):
    ts = []
    torch.cuda.synchronize()
    for i in range(20):
        t0 = time.perf_counter()
        q.forward_host(z_pin, out_indices=idx_host, wait=False)
        ts.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); torch.cuda.synchronize(); tail = (time.perf_counter() - t0) * 1e3
    print("host ms per call:", " ".join(f"{t:.2f}" for t in ts), "| final sync", f"{tail:.2f}", flush=True)
print(torch.cuda.memory_reserved() >> 20, "MiB reserved", torch.cuda.memory_stats()["num_alloc_retries"], "retries", torch.cuda.memory_stats().get("num_device_alloc", -1), "device allocs")
