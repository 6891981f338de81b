#!/usr/bin/env python
"""(CPU) Key metrics per kernel from `ncu -i <rep> --page raw --csv` output (stdin or file): duration, DRAM bytes, pipe
utilisation, issue activity, occupancy, top warp-stall reasons.  Used to write the profiles/*_key_metrics.txt summaries."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
ix = {h: i for i, h in enumerate(hdr)}
name_i = ix.get("Kernel Name", 4)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("=" * 100)
    print(r[name_i][:150])
    for w in want:
        if w in ix:
            print(f"  {w:75s} {r[ix[w]]:>16s} {rows[1][ix[w]]}")
    stalls = []
    for h, i in ix.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  warp stalls per issue-active cycle:", ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
