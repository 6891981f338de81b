#!/usr/bin/env python
"""(CPU) Sums dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum over the kernels of an
`ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` log and prints a JSON entry for
profiles/r02_lp_call_dram.json:  summarize_dram.py <key> <ncu csv> [<skip first n launches>]"""
import csv
import json
import sys

key, path = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for row in csv.DictReader(lines):
    rows.append(row)
per = {}
for row in rows:
    kid = int(row["ID"])
    d = per.setdefault(kid, dict(name=row["Kernel Name"], read=0.0, write=0.0, ns=0.0))
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if row["Metric Name"].startswith("dram__bytes"):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d["read" if "read" in row["Metric Name"] else "write"] += v * scale
    elif row["Metric Name"] == "gpu__time_duration.sum":
        d["ns"] += v * {"ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1, "msecond": 1e6}.get(unit, 1)
ids = sorted(per)[skip:]
tot = sum(per[i]["read"] + per[i]["write"] for i in ids)
out = dict(bytes=int(tot), read=int(sum(per[i]["read"] for i in ids)), write=int(sum(per[i]["write"] for i in ids)), capture=path,
           kernels=[dict(name=per[i]["name"][:60], bytes=int(per[i]["read"] + per[i]["write"]), us=round(per[i]["ns"] / 1e3, 1)) for i in ids])
print(json.dumps({key: out}, indent=1))
