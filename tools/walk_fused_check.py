#!/usr/bin/env python
"""Development check of the fused tcgen05 walk (walk_fused.cu) against the fp64 numpy oracle: loss, A, dx at a few geometries,
and its time next to the shared-memory fp32 kernels.  Runs on a B200 through gpurun."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402
from oracle import walk_oracle  # noqa: E402

BWD = os.environ.get("CHECK_BWD", "1") == "1"


def check(B, T, N, tau, seed=0, spread=1.0):
    torch.manual_seed(seed)
    x = (torch.randn(B, T, N, 128, device="cuda") * spread + torch.randn(1, 1, 1, 128, device="cuda") * (1.0 - spread)).contiguous()
    xr = x.clone().requires_grad_(True)
    loss, A, _ = crw.ops.walk_loss(xr, tau, True, crw.ops.PREC_BF16X3)
    torch.cuda.synchronize()
    x64 = x.double().cpu().numpy()
    ref_loss, dA, dE, dx = walk_oracle.walk_backward_chain(x64, tau)
    _, refA = walk_oracle.crw_forward(x64, tau)
    eA = np.abs(A.detach().cpu().numpy() - refA).max()
    el = abs(float(loss) - ref_loss) / max(abs(ref_loss), 1e-12)
    msg = f"B={B} T={T} N={N} tau={tau} spread={spread}: loss {float(loss):.7f} ref {ref_loss:.7f} rel {el:.2e}  max|A - ref| {eA:.2e}"
    if BWD:
        loss.backward()
        torch.cuda.synchronize()
        g = xr.grad.double().cpu().numpy()
        eg = np.abs(g - dx).max() / max(np.abs(dx).max(), 1e-30)
        msg += f"  dx rel(max-norm) {eg:.2e}"
    print(msg, flush=True)


for args in [(2, 3, 47, 0.07), (2, 4, 47, 0.07), (3, 10, 47, 0.07), (2, 10, 49, 0.07), (2, 20, 47, 0.07), (2, 6, 64, 0.07), (2, 6, 16, 0.07),
             (2, 10, 47, 0.01), (2, 10, 47, 0.07, 0, 0.3)]:
    check(*args)

# timing at BASELINE config 2
B, T, N = 32, 10, 47
x = torch.randn(B, T, N, 128, device="cuda", requires_grad=True)
for prec, name in [(crw.ops.PREC_BF16X3, "bf16x3"), (crw.ops.PREC_FP32, "fp32")]:
    for _ in range(5):
        loss, _, _ = crw.ops.walk_loss(x, 0.07, False, prec)
        if BWD:
            loss.backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        loss, _, _ = crw.ops.walk_loss(x, 0.07, False, prec)
        if BWD:
            loss.backward()
    e1.record()
    torch.cuda.synchronize()
    print(f"config 2 walk {'fwd+bwd' if BWD else 'fwd'} {name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call (eager dispatch)")
    with torch.no_grad():
        e0.record()
        for _ in range(n):
            crw.ops.walk_loss(x, 0.07, False, prec)
        e1.record()
        torch.cuda.synchronize()
    print(f"config 2 walk fwd only {name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call (eager dispatch)")

from torch.profiler import profile, ProfilerActivity  # noqa: E402
import collections  # noqa: E402
for prec, name in [(crw.ops.PREC_BF16X3, "bf16x3"), (crw.ops.PREC_FP32, "fp32")]:
    agg = collections.defaultdict(list)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            loss, _, _ = crw.ops.walk_loss(x, 0.07, False, prec)
            if BWD:
                loss.backward()
        torch.cuda.synchronize()
    for e in prof.events():
        if e.device_type.name == "CUDA":
            agg[e.name[:60]].append(e.time_range.end - e.time_range.start)
    print(name, "kernel times (us, mean of 10):", {k: round(sum(v) / 10, 1) for k, v in agg.items()}, "sum", round(sum(sum(v) for v in agg.values()) / 10, 1))
