#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0][-60:], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"{'kernel':62s} {'n':>4s} {'mean_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:62s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / total:7.1%}")


if __name__ == "__main__":
    main(sys.argv[1])
