#!/usr/bin/env python
"""Profiling aid: per-phase cycles of the fused walk kernels (CTA 0, thread 0) at BASELINE config 2."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CRW_WALK_FUSED"] = "1"
os.environ["CRW_WALK_PROF"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

B, T, N = 32, 10, 47
x = torch.randn(B, T, N, 128, device="cuda", requires_grad=True)
L = crw._lib.lib()
buf = np.zeros(32, dtype=np.uint64)
for _ in range(3):
    loss, _, _ = crw.ops.walk_loss(x, 0.07, False, crw.ops.PREC_BF16X3)
    loss.backward()
torch.cuda.synchronize()
L.crw_debug_walk_fused_profile(None, 1)
loss, _, _ = crw.ops.walk_loss(x, 0.07, False, crw.ops.PREC_BF16X3)
loss.backward()
torch.cuda.synchronize()
L.crw_debug_walk_fused_profile(buf.ctypes.data_as(ctypes.c_void_p), 1)
fw = ["wait frame", "convert", "issue A", "wait A", "epilogue A (+pub)", "issue X", "wait X", "epilogue X (+pub)", "issue M", "wait M", "epilogue M + save"]
bw = ["issue (a)", "wait (a)", "epilogue (b)", "issue (c,d)", "wait (c,d)", "epilogue (e)", "convert next", "wait block"]
if os.environ.get("CRW_WALK_ROLES", "") != "0":
    print("forward, role-split kernel, element 0: cycles [total, waiting for loads / hand-over, waiting for MMAs]")
    for i, n in enumerate(["producer 0", "producer 1", "chain", "cycle"]):
        print(f"  {n:12s} {int(buf[4 * i]):8d} {int(buf[4 * i + 1]):8d} {int(buf[4 * i + 2]):8d}")
print("forward, cycles over the kernel (thread 0 of CTA 0; 1 us ~ 1900 cycles):")
for i, n in enumerate(fw):
    print(f"  {n:20s} {int(buf[i]):8d}")
print("  total", int(buf[:16].sum()))
if os.environ.get("CRW_WALK_ROLES", "") != "0":
    print("backward, role-split kernel, element 0: cycles [total, waiting for loads / hand-over, waiting for MMAs]")
    for i, n in enumerate(["chain", "dA", "dE 0", "dE 1"]):
        print(f"  {n:12s} {int(buf[16 + 4 * i]):8d} {int(buf[16 + 4 * i + 1]):8d} {int(buf[16 + 4 * i + 2]):8d}")
print("backward:")
for i, n in enumerate(bw):
    print(f"  {n:20s} {int(buf[16 + i]):8d}")
print("  total", int(buf[16:].sum()))
