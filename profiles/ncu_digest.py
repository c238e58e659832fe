#!/usr/bin/env python
"""Digest an .ncu-rep (read offline with `ncu -i`): headline metrics, stall mix by code region, wait spins."""
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, vals = rows[0], rows[2]
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    for i, h in enumerate(hdr):
        if h in keys:
            print(f"{h:75s} {vals[i]} {rows[1][i]}")


def source(rep, top=14):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
    print("total samples", tot, "instructions", len(data))
    for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][ci["# Samples"]] or 0))[:top]:
        n = int(r[ci["# Samples"]])
        st = sorted(((int(r[ci[s]] or 0), s.replace("stall_", "")) for s in stalls), reverse=True)[:2]
        print(f"{i:5d} {n:7d} {n / tot:6.2%} ex={r[ci['Instructions Executed']]:>10s} t={r[ci['Avg. Threads Executed']]:>3s} "
              f"{r[ci['Source']][:60]:60s} {st}")
    print("-- waits (TRYWAIT executions = spins):")
    for i, r in enumerate(data):
        if "TRYWAIT" in r[ci["Source"]]:
            print(f"{i:5d} ex={r[ci['Instructions Executed']]:>10s} t={r[ci['Avg. Threads Executed']]:>3s} samples={r[ci['# Samples']]:>6s} {r[ci['Source']][:60]}")


if __name__ == "__main__":
    raw(sys.argv[1])
    source(sys.argv[1])
