#!/usr/bin/env python
"""Benchmark of the VQ hot path: latents quantized per second (distance + argmin + gather).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|rvq|c4|c5] [--scaling weak|strong]
                    [--impl reference] [--mode fp32|bf16_input] [--graph] [--also c2,rvq,c5|none]

One "step" = one full quantizer forward (search, gather, straight-through, commitment partial sums,
histogram, statistics) over one batch of synthetic latents resident in HBM.  Rank 0 prints ONE JSON
line on stdout.  Workloads (BASELINE.json configs):
  c3  = K=8192 D=256 N=2^22, fp32 (DEFAULT: configs[2], the shape BASELINE.json's "% tensor peak" metric and the
        north-star target are quoted on; it fits one GPU)
  c2  = K=512 D=64 N=2^20 (configs[1], HBM-bound)
  rvq = stage-2 shape 4x1024 D=512 N=8192 (eval);  c4 = index extraction (stage-2 RVQ over 2^21 latents per GPU,
        token-major int32 out, configs[3]);  c5 = training step of the path (forward with EMA scatter-add, vq_loss,
        backward) at the stage-2 shape (configs[4])
A default single-GPU run also measures c2, rvq and c5 AFTER the headline line is out and prints their JSON lines on
STDERR (evidence for the other BASELINE configs; stdout keeps exactly one line).
Multi-GPU: every rank quantizes its own rows against the replicated codebook; the only collective is one
all-reduce of [sq-err | count | histogram] per step.  --scaling weak (default): rows_per_gpu = the workload's N;
--scaling strong: the workload's N split over the ranks (2^22 / G rows per GPU for c3).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(K_per=512, D=64, L=1, N=1 << 20, desc="VectorQuantizer codebook search K=512 D=64, N=1M synthetic latents"),
    "c3": dict(K_per=8192, D=256, L=1, N=1 << 22, desc="large codebook K=8192 D=256 fp32, N=4M latents"),
    "rvq": dict(K_per=1024, D=512, L=4, N=8192, desc="stage2_vq RVQ 4x1024 D=512, N=8192 latents"),
    # configs[3]: what scripts/extract_code_indices.py executes per rank -- the stage-2 residual quantizer over
    # 2^21 latents per GPU (2^24 over 8 GPUs), indices re-laid out to token-major [B, M*Q] int32 on the device
    "c4": dict(K_per=1024, D=512, L=4, N=1 << 21, extract=True,
               desc="extract_code_indices, stage2 RVQ 4x1024 D=512, 2M latents per GPU (16M over 8), [B, M*Q] int32 out"),
    # configs[4]: one training step of the path at the stage-2 shape: forward with EMA update (scatter-add),
    # vq_loss, backward to z_e (straight-through + commitment)
    "c5": dict(K_per=1024, D=512, L=4, N=8192, train=True,
               desc="VQ training step (fwd + EMA scatter-add + vq_loss + backward), stage2 RVQ 4x1024 D=512, N=8192 per GPU"),
}
METRIC = "latents quantized/sec (distance+argmin+gather)"
UNIT = "latents/s"


def w_name(w):
    return [k for k, v in WORKLOADS.items() if v is w][0]


def rows_per_gpu(w, args, world):
    return w["N"] // world if args.scaling == "strong" else w["N"]


def config_of(w, args, world):
    """The SAME dictionary for the B200 arm and the reference arm (the driver compares them key by key)."""
    n = rows_per_gpu(w, args, world)
    cfg = {"workload": w["desc"], "K": w["K_per"], "D": w["D"], "levels": w["L"], "rows_per_gpu": n,
           "search_mode": args.mode, "scaling": args.scaling,
           "l2": "inputs+outputs per step exceed the 126 MB L2" if n * w["D"] * 12 > 126e6
                 else "L2-resident working set (no flush)",
           "parallelism": f"rows sharded x{world}, codebook replicated", "cuda_graph": bool(args.graph)}
    if w.get("train"):
        cfg["ema_sync"] = args.ema_sync if world > 1 else "local"
    return cfg


def synth(w, seed, n_rows=None):
    rs = np.random.RandomState(seed)
    K, D, L = w["K_per"] * w["L"], w["D"], w["L"]
    E = (rs.standard_normal((K, D)) / np.sqrt(D)).astype(np.float32)
    for lvl in range(1, L):
        E[lvl * w["K_per"]:(lvl + 1) * w["K_per"]] *= np.float32(0.6 ** lvl)
    n = w["N"] if n_rows is None else n_rows
    z = rs.standard_normal((n // 64, 64, D)).astype(np.float32)      # the first n rows of the full-size stream
    return E, z


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline on the host cores.  The reference is pure Python + torch ATen: when its sources
# are reachable (/root/reference in the build container, baseline/_ref staged by baseline/stage_reference.py on
# the GPU box) the LIVE class models/vq_vae.py::VectorQuantizerEMA is timed (kind "reference"); otherwise
# oracle/torch_port.py, the same ATen op sequence pinned to the same golden vectors (kind "port").
# ----------------------------------------------------------------------------------------------
def cpu_sample_rows(w):
    # bounded sample: ~10-30 s of CPU work on a few-dozen-core host
    return {"c2": 1 << 20, "c3": 1 << 16, "rvq": 8192, "c4": 8192, "c5": 8192}[w_name(w)]


def live_reference_class():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(cand, "models", "vq_vae.py")):
            sys.path.insert(0, cand)
            try:
                for name in ("models", "models.vq_vae"):
                    sys.modules.pop(name, None)
                from models.vq_vae import VectorQuantizerEMA
                return VectorQuantizerEMA, cand
            except Exception:
                pass
            finally:
                sys.path.remove(cand)
    return None, None


def time_cpu(w, steps, warmup, n_rows):
    """Times the reference's CPU implementation of the path on n_rows rows, all host threads.
    Returns (median seconds per pass, rows, kind, what, indices [L, n] of the FIRST pass -- the one that starts
    from the initial codebook, which matters for the training workload)."""
    import torch
    torch.set_num_threads(os.cpu_count())
    E, z = synth(w, 1234, n_rows)
    K_per, D, L = w["K_per"], w["D"], w["L"]
    Et, zt = torch.from_numpy(E), torch.from_numpy(z).reshape(-1, D)
    n = zt.shape[0]
    chunk = 65536 if K_per <= 1024 else 16384            # the reference materialises N x K: row chunks
    Ref, where = live_reference_class()
    g_st = torch.from_numpy(np.random.RandomState(7).standard_normal(zt.shape).astype(np.float32))

    if Ref is not None:
        kind, what = "reference", f"live models/vq_vae.py::VectorQuantizerEMA from {where}"
        q = Ref(K_per, D, num_quantizers=L, print_init=False, decay=0.98, beta=0.0005)
        q.embedding.copy_(Et)
        if w.get("train"):
            q.train()

            def one_pass():
                ze = zt.view(-1, 64, D).detach().requires_grad_(True)
                st, zq, idx, stats = q(ze, do_ema_update=True)
                loss_vq = q.beta * torch.nn.functional.mse_loss(zq.detach(), ze)      # models/vq_vae.py:1292-1294
                torch.autograd.backward([st, loss_vq], [g_st.view_as(st), torch.ones(())])
                return idx.view(L, -1)
        else:
            q.eval()

            def one_pass():
                out = torch.empty(L, n, dtype=torch.int64)
                with torch.no_grad():
                    for s in range(0, n, chunk):
                        idx = q(zt[s:s + chunk].view(-1, 64, D), do_ema_update=False)[2]
                        out[:, s:s + chunk] = idx.view(L, -1)
                return out
    else:
        from oracle import torch_port
        kind, what = "port", "torch ATen port (oracle/torch_port.py)"
        state = {"E": Et.clone(), "ema_cluster_size": torch.zeros(K_per * L), "ema_embedding": torch.zeros(K_per * L, D)}

        def one_pass():
            if w.get("train"):
                return torch_port.train_step(zt, state, K_per, L, 0.0005, 0.98, 1e-5, g_st)[0].view(L, -1)
            return torch_port.forward_eval(zt, Et, K_per, L, chunk)[1].view(L, -1)

    first = None
    for _ in range(warmup):
        idx = one_pass()
        first = idx if first is None else first
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        idx = one_pass()
        ts.append(time.perf_counter() - t0)
        first = idx if first is None else first
    return float(np.median(ts)), n, kind, what, first.numpy().copy()


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n_rows = cpu_sample_rows(w)
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    sec, n, kind, what, _ = time_cpu(w, steps, warm, n_rows)
    val = n / sec
    sample = (f"{n} of {rows_per_gpu(w, args, world)} rows per step ({w_name(w)}), median of {steps}, {what} on all "
              f"host threads, row-chunked")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(w, args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def parity_on_rows(z_rows, E, K_per, L, idx_gpu, idx_cpu):
    """Index parity of the B200 arm against the CPU arm ON THE ROWS THE CPU ARM QUANTIZED,
    chain-aware for residual levels: a row that differs at a level is judged there by the fp64 near-tie rule
    (gap < 1e-6 relative) and dropped from the deeper levels."""
    from oracle import vq_oracle as O
    a, b = np.asarray(idx_gpu).reshape(L, -1), np.asarray(idx_cpu).reshape(L, -1)
    n = a.shape[1]
    alive = np.ones(n, bool)
    residual = np.asarray(z_rows, dtype=np.float32).reshape(n, -1)
    mism = outside = 0
    for lvl in range(L):
        s = lvl * K_per
        El = E[s:s + K_per]
        bad = alive & (a[lvl] != b[lvl])
        if bad.any():
            mm, out = O.near_tie_rows(residual[bad], El, a[lvl][bad] - s, b[lvl][bad] - s)
            mism += int(bad.sum())
            outside += int(out.size)
        alive &= ~bad
        if lvl + 1 < L:
            residual = residual - El[b[lvl] - s]
    return {"rows": int(n), "levels": int(L), "mismatch": mism, "outside": outside}


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def summary(self, t0, t1):
        """Samples taken while the GPU was under this benchmark's load: the device-resident timed region, the
        end-to-end loop and the kernel-timing loop run back to back (same kernels, same load)."""
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(rows), "window_s": round(t1 - t0, 3)}
        try:
            sm = [float(r[0]) for r in rows]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][1]) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
            out["power_w_max"] = max(float(r[2]) for r in rows) if rows else None
        except Exception:
            pass
        return out


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained"), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, None, "fallback (B200_PROFILING.md)"


def run_b200(args, w, with_cpu=True):
    import torch
    import torch.distributed as dist

    import pytorch_vae_b200 as vq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    N = rows_per_gpu(w, args, world)
    E, z_host = synth(w, 1234 + rank, N)                  # rows differ per rank; codebook identical
    E, _ = synth(w, 1234, 64) if rank else (E, None)
    D, K, L = w["D"], w["K_per"], w["L"]
    train, extract = bool(w.get("train")), bool(w.get("extract"))
    q = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False, search_mode=args.mode).to(dev)
    q.embedding.copy_(torch.from_numpy(E))
    q.stats_sync = world > 1
    if train:                                            # stage-2 schedule values (configs/stage2_vq.yaml:117-123)
        q.train()
        q.beta, q.decay = 0.0005, 0.98
        # "local" (default): every rank updates its replicated codebook from its own rows -- the reference's behaviour
        # (under DDP rank 0's buffers are then broadcast) and BASELINE.json's north star ("only the scalar loss and
        # the code-usage histogram are NCCL-allreduced"); "allreduce": segment sums summed over ranks before every
        # level's EMA finalize, so the replicas stay identical without a broadcast
        q.ema_sync = args.ema_sync if world > 1 else "local"
    else:
        q.eval()
    z_pin = torch.from_numpy(z_host).pin_memory()
    z = z_pin.to(dev)
    Bz, Mz = z.shape[0], z.shape[1]
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    g_st = torch.randn(z.shape, device=dev, generator=gen) if train else None   # what a decoder would send back
    beta_t = torch.full((), q.beta, device=dev)

    # ---- index parity on the rows the CPU arm quantizes (rank 0: the first rows of its batch), taken from the
    # initial codebook BEFORE any training step moves it
    cpu_rows = min(cpu_sample_rows(w), N)
    idx_first = None
    if rank == 0 and with_cpu:
        q0 = vq.VectorQuantizerEMA(K, D, num_quantizers=L, print_init=False, search_mode=args.mode, decay=0.98).to(dev)
        q0.embedding.copy_(torch.from_numpy(E))
        q0.train(train)                                  # training workload: the FIRST training forward (EMA update on)
        with torch.no_grad():
            idx_first = q0(z[: cpu_rows // 64], do_ema_update=train)[2].reshape(L, -1).cpu().numpy()
        del q0

    graphed = vq.GraphedForward(q, z) if args.graph and not train else None
    graphed_train = vq.GraphedTrainStep(q, z) if args.graph and train else None

    def train_step(zd):
        if graphed_train is not None:                    # forward + EMA + backward replayed as ONE CUDA graph
            return graphed_train(zd, g_st, q.beta)
        ze = zd.detach().requires_grad_(True)
        st, zq, idx, stats = q(ze, do_ema_update=True)
        torch.autograd.backward([st, q.last_commit], [g_st, beta_t])    # d(recon)/d z_q_st + beta * commit (:1292-1294)
        return st, zq, idx, stats, ze.grad

    def step():
        if train:
            return train_step(z)
        if graphed is not None:
            return graphed(z)
        with torch.no_grad():
            o = q(z, do_ema_update=False)
            if extract:                                  # scripts/extract_code_indices.py:195-209 on the device
                return o + (vq.ops.relayout_indices(o[2], L, Bz, Mz, torch.int32),)
            return o

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = step()
    sync()
    clocks = Clocks(local) if rank == 0 else None
    time.sleep(0.3)
    l0 = vq.ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    sync()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    launches = vq.ops.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    ms_step = ms / args.steps
    value = world * N / (ms_step * 1e-3)
    del out

    # ---- end to end through the public API with HOST buffers, every step inside the timed region: the
    # latents start in pinned host memory and the step's result ends in host memory.  Inference workloads go
    # through VectorQuantizerEMA.forward_host (chunked: H2D, kernels and D2H overlap on three streams); the
    # training step copies its batch in, runs forward + backward and reads the loss back.
    idx_dtype = torch.int32 if extract else torch.int64
    idx_host = torch.empty(L * N, dtype=idx_dtype).pin_memory()
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    # training: every step's batch is copied from pinned host memory inside the timed region, on a copy stream,
    # double-buffered, so that the copy of step i+1 overlaps the kernels of step i (what a DataLoader with
    # pin_memory + non_blocking transfers does for the reference's training loop)
    copy_stream = torch.cuda.Stream(dev) if train else None
    stage = [torch.empty_like(z) for _ in range(2)] if train else None
    copied = [torch.cuda.Event() for _ in range(2)] if train else None
    used = [torch.cuda.Event() for _ in range(2)] if train else None

    def issue_copy(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(used[b])                      # the step that last read this buffer has finished
            stage[b].copy_(z_pin, non_blocking=True)
            copied[b].record(copy_stream)

    def e2e_loop(n_steps):
        main = torch.cuda.current_stream(dev)
        if not train:
            for _ in range(n_steps):
                q.forward_host(z_pin, out_indices=idx_host, wait=False, token_major=torch.int32 if extract else None)
            return
        for b in range(2):
            used[b].record(main)
        issue_copy(0)
        for i in range(n_steps):
            b = i & 1
            if i + 1 < n_steps:
                issue_copy(i + 1)
            main.wait_event(copied[b])
            o = train_step(stage[b])
            used[b].record(main)
            loss_host[:2].copy_(o[3], non_blocking=True)
            loss_host[2:].copy_(q.last_commit.detach().reshape(1), non_blocking=True)

    # warm-up of the end-to-end loop: two steps where a step moves >= 1 GiB (c3 / c4: 78 ms each); sixty for the small
    # workloads, whose back-to-back asynchronous calls otherwise spend the timed region growing the caching allocator's
    # pool (one cudaMalloc per call until ~40 calls are in flight: profiles/probes/forward_host_probe.py)
    e2e_loop(2 if N * D * 4 >= (1 << 30) else 60)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    sync()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t)
    e2e_val = world * N / (e2e_ms / args.steps * 1e-3)
    d2h_bytes = (12 if train else L * N * idx_host.element_size() + 8)

    # ---- roofline of the dominant kernel: the library brackets every launch of the fused distance+argmin
    # kernel (tcgen05 path; the SIMT kernel on shapes that take it) with CUDA events on the launching stream
    # while full steps run, so the duration is the kernel's own, measured inside a live step.
    import ctypes
    lib = vq._cabi.lib
    mode = vq.quantizer._MODES[args.mode]
    lib.vqb200_timing_enable(1)
    n_steps_t = max(3, min(args.steps, 10))
    for _ in range(n_steps_t):
        if train:
            gt, graphed_train = graphed_train, None              # eager even under --graph: the hooks live in the library calls
            train_step(z)
            graphed_train = gt
        else:
            with torch.no_grad():
                q(z, do_ema_update=False)
    torch.cuda.synchronize()
    lib.vqb200_timing_enable(0)
    tot, nl = ctypes.c_float(0), ctypes.c_int(0)
    lib.vqb200_timing_collect(ctypes.byref(tot), ctypes.byref(nl))
    launches_per_step = max(1, nl.value // n_steps_t)
    kern_ms = tot.value / max(1, nl.value)                       # average duration of ONE launch
    rows_per_launch = N * L / launches_per_step                  # rows one launch scans (chunks x levels per step)
    hbm, tf, tf_sus, peak_src = peaks()
    flops = 2.0 * rows_per_launch * K * D
    fused = L == 1 and bool(lib.vqb200_quantize_fused_supported(N, K, D, mode))
    # a kernel that does the full forward moves read z, write z_q, z_q_st, idx = 12 D + 8 bytes per
    # row; the stand-alone search kernel is "codes-only" (read z, write idx = 4 D + 8)
    bytes_alg = rows_per_launch * ((12.0 * D + 8.0) if fused else (4.0 * D + 8.0))
    tensor_bound = flops / (tf * 1e12) > bytes_alg / (hbm * 1e9)
    on_tc = bool(lib.vqb200_search_path(N, K, D, mode))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        key = (w_name(w) + ("_fused" if fused else "")) if args.mode == "fp32" else ""
        t = json.load(open(tpath)).get(key, None)
        if t:                                                    # dram bytes of one launch from one ncu --set full capture
            traffic = t["dram_bytes_per_row"] * rows_per_launch
    if tensor_bound:
        ach = flops / (kern_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf}
        if tf_sus:
            roof["frac_of_sustained_peak"] = ach / tf_sus
        step_tf = 2.0 * N * L * K * D / (ms_step * 1e-3) / 1e12  # the whole step (every side pass included)
        roof["step_achieved"] = step_tf
        roof["step_frac"] = step_tf / tf
    else:
        ach = bytes_alg / (kern_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm}
        step_gbs = N * (12.0 * D + 8.0 * L) / (ms_step * 1e-3) / 1e9
        roof["step_achieved"] = step_gbs
        roof["step_frac"] = step_gbs / hbm
    persistent = L > 1 and bool(lib.vqb200_rvq_fused_supported(N, K, D, L, mode))
    kname = ("quantize_fused_kernel (tcgen05 distance+argmin+gather, one pass)" if fused else
             "rvq_fused_kernel (persistent: every level's tcgen05 search, exact re-rank, residual update and the outputs "
             "in one launch" + ("; training: + EMA segment sums)" if train else ")") if persistent else
             "search_tc2_kernel / search_tc_kernel (tcgen05 distance+argmin; cta_group::2 pairs for large N)" if on_tc
             else "search_simt_kernel")
    roof.update({"traffic": traffic, "kernel": kname,
                 "kernel_ms": kern_ms, "launches_per_step": launches_per_step, "rows_per_launch": rows_per_launch,
                 "kernel_share_of_step": kern_ms * launches_per_step / ms_step, "peak_source": peak_src,
                 "algorithmic_unit": "2*K*D flop per row" if tensor_bound else
                 ("12*D+8 bytes per row" if fused else "4*D+8 bytes per row")})

    clk = clocks.summary(t0, time.perf_counter()) if clocks else None
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "bf16-in/f32-acc", "data": "synthetic",
            "config": config_of(w, args, world),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": N * D * 4,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps,
                    "api": "train step: H2D batch (double-buffered on a copy stream), forward+backward, D2H loss" if train else
                           "VectorQuantizerEMA.forward_host (chunked H2D / kernels / D2H on three streams)"},
            "gpu_launches": launches, "clocks": clk, "roofline": roof,
        }
        if with_cpu:
            sec, n, kind, what, idx_cpu = time_cpu(w, 1, 1, cpu_rows)
            line["cpu_baseline"] = {"value": n / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                                    "sample": f"{n} of {N} rows, one {'training step' if train else 'pass'}, {what} "
                                              f"on all host threads"}
            zr = z_host.reshape(-1, D)[:n]
            if not train:
                par = parity_on_rows(zr, E, K, L, idx_first, idx_cpu)
                par["what"] = "eval forward of the timed rows, B200 arm vs CPU arm, fp64 near-tie rule (1e-6), chain-aware"
            else:
                # the codebook moves between the levels of a training forward (EMA update after every level), so every
                # level's decision is judged by the oracle's teacher-forced replay from identical inputs; the CPU arm's
                # own first pass (same initial state) is compared chain-aware
                from oracle import vq_oracle as O
                from oracle.replay import chain_alive, replay_step
                oq = O.OracleQuantizer(K, D, num_quantizers=L, embedding=E, decay=0.98)
                oq.training = True
                rep = replay_step(oq, zr.reshape(-1, 64, D), idx_first)
                first, frac = chain_alive(idx_first, idx_cpu, L)
                par = {"rows": int(n), "levels": int(L), "mismatch": int(sum(rep["mismatch"])),
                       "outside": int(sum(rep["outside"])), "identical_to_cpu_arm": frac,
                       "what": "first training forward from the initial codebook (EMA update after every level): every "
                               "level against the oracle's pick from identical inputs; identical_to_cpu_arm = fraction "
                               "of rows whose whole index chain equals the CPU arm's"}
            line["parity_on_timed_rows"] = par
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=None,
                    help="default c3 (K=8192 D=256 N=4M: the configuration the metric is quoted on)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="multi-GPU: weak = the workload's N rows per GPU (default); strong = N / G rows per GPU")
    ap.add_argument("--mode", choices=["fp32", "bf16_input"], default="fp32")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--ema-sync", choices=["allreduce", "local"], default="local",
                    help="c5 on several GPUs: update the codebook per rank (default, as the reference) or all-reduce the "
                         "EMA segment sums")
    ap.add_argument("--graph", action="store_true", help="replay the forward as one CUDA graph (launch-bound shapes)")
    ap.add_argument("--also", default=None,
                    help="comma list of further workloads measured after the headline line, printed on STDERR "
                         "(default: c2,rvq,c5 for a single-GPU run of the default workload; 'none' to skip)")
    args = ap.parse_args()
    explicit = args.workload is not None
    w = WORKLOADS[args.workload or "c3"]
    if args.impl == "reference":
        run_reference(args, w)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = run_b200(args, w)
    if line is not None:
        print(json.dumps(line), flush=True)
    also = args.also if args.also is not None else ("none" if (explicit or world > 1) else "c2,rvq,c5")
    if also != "none" and world == 1:
        for name in [a for a in also.split(",") if a in WORKLOADS]:
            try:
                extra = run_b200(args, WORKLOADS[name])
                print("[also] " + json.dumps(extra), file=sys.stderr, flush=True)
            except Exception as e:                               # never let the extras break the headline run
                print(f"[also] {name} failed: {e!r}", file=sys.stderr, flush=True)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
