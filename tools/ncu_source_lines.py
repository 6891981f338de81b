#!/usr/bin/env python
"""Attribute the per-instruction counters of an `ncu --page source --csv --print-source sass` dump to source lines, using the
line table of `nvdisasm -g -c <cubin>` of the same build (instruction order is identical).

    cuobjdump -xelf all obj.o && nvdisasm -g -c obj.sm_100a.cubin > disasm.txt
    python tools/ncu_source_lines.py disasm.txt source_sass.csv path/to/file.cu [top_n]
"""
import collections
import csv
import re
import sys

disasm, sass_csv, src_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 14
funcs, cur, line = collections.OrderedDict(), None, None
for l in open(disasm):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        cur, line = m.group(1), None
        funcs[cur] = []
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and cur:
        funcs[cur].append(line)
by_len = collections.defaultdict(list)
for k, v in funcs.items():
    by_len[len(v)].append(k)
src = open(src_path).read().splitlines()
src_name = src_path.split('/')[-1]
kern, data = None, collections.OrderedDict()
for r in csv.reader(open(sass_csv)):
    if r and r[0] == 'Kernel Name':
        kern = r[1]
        data[kern] = []
    elif r and r[0].startswith('0x'):
        data[kern].append(r)
for k, v in data.items():
    cands = by_len.get(len(v), [])
    short = re.sub(r'\(.*', '', k)
    if len(cands) != 1:
        print(f"== {short}: {len(cands)} disassembly candidates with {len(v)} instructions, skipped")
        continue
    agg, smp = collections.Counter(), collections.Counter()
    for ln, r in zip(funcs[cands[0]], v):
        agg[ln] += int(r[5])
        smp[ln] += int(r[4])
    tot, ts = sum(agg.values()), max(sum(smp.values()), 1)
    print(f"== {short}: {tot} warp instructions, {ts} stall samples")
    for ln, c in agg.most_common(top):
        txt = src[ln[1] - 1].strip()[:100] if ln and ln[0] == src_name else str(ln)
        print(f"  {100 * c / tot:5.1f}% instr {100 * smp[ln] / ts:5.1f}% stall  L{ln[1] if ln else '?'}: {txt}")
