#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent residual-VQ kernel (VQB200_DEBUG=5), stage-2 shape by default.

    python profiles/prof_rvq_trace.py [K_per] [D] [L] [N]

Events (cycles since the CTA's first stamp), per level:
  mma : e0 operand tile seen, e1 first code tile issued, e2 last MMA issued
  scan: e0 margins seen, e1 scan finished, e2 worker barrier passed, e3 rows decided (pass A), e4 next operand tile / outputs done
  help: e0 waiting at the worker barrier, e1 barrier passed, e2 pass A done, e3 pass B done
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_vae_b200 as vq  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
L = int(sys.argv[3]) if len(sys.argv) > 3 else 4
N = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False).to(dev).eval()
E = torch.randn(K * L, D, device=dev, generator=g) / np.sqrt(D)
for l in range(1, L):
    E[l * K:(l + 1) * K] *= 0.6 ** l
q.embedding.copy_(E)
z = torch.randn(N // 64, 64, D, device=dev, generator=g)
with torch.no_grad():
    for _ in range(3):
        q(z, do_ema_update=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        q(z, do_ema_update=False)
    torch.cuda.synchronize()
    print(f"K_per={K} D={D} L={L} N={N}: {(time.perf_counter() - t0) / 20 * 1e3:.4f} ms per forward; exhaustive rows "
          f"{int(vq.ops.last_rvq_workspace[:4].view(torch.int32)[0])}")
    os.environ["VQB200_DEBUG"] = "5"
    q(z, do_ema_update=False)
    torch.cuda.synchronize()
