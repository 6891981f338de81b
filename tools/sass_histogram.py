#!/usr/bin/env python
"""(CPU) Per-kernel SASS opcode histogram of radar_sounder_crw_b200/lib/libcrw_b200.so (cuobjdump -sass): the Blackwell opcodes
(UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor load, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, UTCCP =
tcgen05.cp, UTCBAR = tcgen05.commit), the legacy tensor opcode HMMA and the scalar fp32 work (FFMA) per kernel."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radar_sounder_crw_b200", "lib", "libcrw_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
ops = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "FFMA", "MUFU"]
cur, hist, order = None, collections.defaultdict(collections.Counter), []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur]["_total"] += 1
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + "."):
                hist[cur][o] += 1
print(f"{'kernel':78s} {'instrs':>7s} " + " ".join(f"{o:>7s}" for o in ops))
for fn, nm in zip(order, names):
    short = re.sub(r"\(.*", "", nm).replace("void ", "").replace("crw::", "")[:78]
    h = hist[fn]
    print(f"{short:78s} {h['_total']:7d} " + " ".join(f"{h[o]:7d}" for o in ops))
