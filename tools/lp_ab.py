#!/usr/bin/env python
"""Profiling aid: same-box A/B of the tensor-path top-k kernel under different CRW_TC_DEBUG / CRW_LP_* settings.
usage: lp_ab.py "<env assignments>" "<env assignments>" ...   e.g.  lp_ab.py "" "CRW_TC_DEBUG=16"
Prints the mean duration of every kernel of one label-propagation call (config 3 unless LP_CFG=5) per setting."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

PREC = {"bf16x3": crw.ops.PREC_BF16X3, "tcx": crw.ops.PREC_TC_EXACT, "fp32": crw.ops.PREC_FP32}[os.environ.get("LP_PREC", "tcx")]

cfg5 = os.environ.get("LP_CFG") == "5"
R, T, N, C, M = (4, 3125, 49, 128, 4) if cfg5 else (1, 1250, 49, 128, 4)
k, r = (20, 24.0) if cfg5 else (10, 12.0)
torch.manual_seed(11)
feats = torch.randn(R, T, N, C, device="cuda")
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (R, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()
flush = torch.empty(64 * 1024 * 1024, device="cuda")


def run():
    return crw.ops.labelprop(feats, mask0, 20, r, 0.07, k, 0, PREC, True, False)


ref = None
for rep in range(2):
    for setting in sys.argv[1:] or [""]:
        keys = []
        for kv in setting.split():
            a, b = kv.split("=")
            os.environ[a] = b
            keys.append(a)
        for _ in range(3):
            out = run()
        torch.cuda.synchronize()
        if ref is None:
            ref = out[0].clone()
        same = bool((out[0] == ref).all())
        agg = collections.defaultdict(list)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                flush.add_(1.0)
                ev0.record()
                run()
                ev1.record()
                torch.cuda.synchronize()
                tot += ev0.elapsed_time(ev1)
        for e in prof.events():
            if e.device_type.name == "CUDA" and "elementwise" not in e.name:
                agg[e.name[:48]].append(e.time_range.end - e.time_range.start)
        line = "  ".join(f"{n.split('(')[0][-28:]} {sum(v) / 10:7.1f}" for n, v in agg.items() if sum(v) / 10 > 5.0)
        print(f"[{setting or 'default':24s}] call {tot / 10 * 1e3:7.1f} us  labels_same={same} | {line}", flush=True)
        for a in keys:
            del os.environ[a]
