#!/usr/bin/env python
"""Profiling aid: label propagation (tensor path) at BASELINE config 3 under the CRW_TC_DEBUG work-skipping flags, for the
single-CTA and the CTA-pair kernels; kernel time taken from torch.profiler (CUDA activity), not from a profiler-slowed clock."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

T, N, C, M = 1250, 49, 128, 4
torch.manual_seed(11)
feats = torch.randn(1, T, N, C, device="cuda")
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (1, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()


def run():
    return crw.ops.labelprop(feats, mask0, 20, 12.0, 0.07, 10, crw.ops.LP_REF_EXACT, crw.ops.PREC_BF16X3, True, False)


for pair in ("1", "0"):
    os.environ["CRW_LP_PAIR"] = pair
    for dbg in ("0", "1", "2", "4", "5", "6"):
        os.environ["CRW_TC_DEBUG"] = dbg
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                run()
            torch.cuda.synchronize()
        for e in prof.key_averages():
            if "lp_topk" in e.key:
                print(f"pair={pair} debug={dbg}  {e.key[:40]:40s} {e.device_time_total / e.count:8.1f} us")
