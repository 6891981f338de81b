#!/usr/bin/env python
"""Profiling aid: per-kernel durations of the walk fwd+bwd at config 2 (B=32, T=10, N=47) for both precisions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

emb = torch.randn(32, 10, 47, 128, device="cuda", requires_grad=True)
for prec, name in [(crw.ops.PREC_FP32, "fp32"), (crw.ops.PREC_BF16X3, "bf16x3")]:
    for _ in range(3):
        loss, _, _ = crw.ops.walk_loss(emb, 0.07, False, prec)
        loss.backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            loss, _, _ = crw.ops.walk_loss(emb, 0.07, False, prec)
            loss.backward()
        torch.cuda.synchronize()
    tot = 0.0
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if "crw::" in e.key:
            print(f"{name:7s} {e.key[:50]:50s} {e.device_time_total / e.count:7.1f} us")
            tot += e.device_time_total / e.count
    print(f"{name:7s} sum {tot:.1f} us")
