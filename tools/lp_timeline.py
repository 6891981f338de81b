#!/usr/bin/env python
"""Profiling aid: start/end of every kernel of one tensor-path label propagation call (config 3), from torch.profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

PREC = {"bf16x3": crw.ops.PREC_BF16X3, "tcx": crw.ops.PREC_TC_EXACT, "fp32": crw.ops.PREC_FP32}[os.environ.get("LP_PREC", "tcx")]

T, N, C, M = 1250, 49, 128, 4
torch.manual_seed(11)
feats = torch.randn(1, T, N, C, device="cuda")
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (1, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()


def run():
    return crw.ops.labelprop(feats, mask0, 20, 12.0, 0.07, 10, 0, PREC, True, False)


for _ in range(5):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{e.name[:60]:60s} start {e.time_range.start - t0:8.1f} us  dur {e.time_range.end - e.time_range.start:8.1f} us")
