#!/usr/bin/env python
"""Profiling aid: per-phase cycle counters of the exact tensor path's filter epilogue and refine kernel (CRW_TC_DEBUG=8) at
BASELINE config 3, serialised launches (CRW_LP_NO_FORK=1) so that the kernels do not compete for SMs."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CRW_TC_DEBUG"] = str(8 | int(os.environ.get("EXTRA_DEBUG", "0")))
os.environ.setdefault("CRW_LP_NO_FORK", "1")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

T, N, C, M = 1250, 49, 128, 4
torch.manual_seed(11)
feats = torch.randn(1, T, N, C, device="cuda")
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (1, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()
L = crw._lib.lib()
buf = np.zeros(2 * 160 * 18 * 8, dtype=np.uint64)


def run():
    crw.ops.labelprop(feats, mask0, 20, 12.0, 0.07, 10, 0, crw.ops.PREC_TC_EXACT, True, False)
    torch.cuda.synchronize()


run()
L.crw_debug_lp_x_profile(None, 1)
run()
L.crw_debug_lp_x_profile(buf.ctypes.data_as(ctypes.c_void_p), 1)
p = buf.reshape(2, 160, 18, 8).astype(np.float64)[:, :148]
fn = ["wait acc_full", "validity mask", "tmem ld + scan", "flushes", "final -> global", "item total"]
print("filter epilogue, cycles per warp over the launch (mean / max over 148 x 16 warps; 1 us ~ 1900 cycles):")
for i, nm in enumerate(fn):
    print(f"  {nm:16s} mean {p[0, :, :16, i].mean():9.0f}   max {p[0, :, :16, i].max():9.0f}")
rn = ["metadata staging", "wait", "issue / chain", "rank + finish", "rescans", "chunk total"]
print("refine, consumers (warps 0-7):")
for i, nm in enumerate(rn):
    print(f"  {nm:16s} mean {p[1, :, :8, i].mean():9.0f}   max {p[1, :, :8, i].max():9.0f}")
print("refine, producer (warp 8):")
for i, nm in enumerate(rn):
    print(f"  {nm:16s} mean {p[1, :, 8, i].mean():9.0f}   max {p[1, :, 8, i].max():9.0f}")
