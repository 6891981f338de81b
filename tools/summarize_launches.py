#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, min, share."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"][:72], []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {f}  (ns; ncu per-launch times are cold-cache and serialised: compare shares)")
    for k, v in agg.items():
        print(f"{k:72s} n={len(v):4d} mean={sum(v)/len(v):11.1f} min={min(v):11.1f} share={100*sum(v)/tot:5.1f}%")
