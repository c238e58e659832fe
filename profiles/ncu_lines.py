#!/usr/bin/env python
"""Stall samples of an .ncu-rep per CUDA source line (needs -lineinfo and --import-source on).
    python profiles/ncu_lines.py rep.ncu-rep [min_samples]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, acc = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and r[0].isdigit() and len(r) > hdr["# Samples"]:
        if r[2] != "-":                    # a SASS row under its CUDA line
            continue
        iv = lambda x: int(x) if x.lstrip("-").isdigit() else 0
        n = iv(r[hdr["# Samples"]])
        stall = sorted(((iv(r[i]), h[6:]) for h, i in hdr.items() if h.startswith("stall_") and "Not" not in h),
                       reverse=True)[:2]
        acc.append((n, fname, int(r[0]), r[1].strip()[:100], iv(r[hdr["Instructions Executed"]]), stall))
tot = sum(a[0] for a in acc)
print("total samples", tot)
for n, f, ln, src, ex, stall in sorted(acc, reverse=True):
    if n < mins:
        break
    print(f"{n:6d} {n / tot:6.2%} ex={ex:9d} {f}:{ln:<5d} {src}  {stall}")
