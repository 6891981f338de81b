#!/usr/bin/env python
"""Profiling aid: per-phase cycle counters of the tensor-path top-k epilogue (CRW_TC_DEBUG=8) at BASELINE config 3."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

T, N, C, M = 1250, 49, 128, 4
torch.manual_seed(11)
feats = torch.randn(1, T, N, C, device="cuda")
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (1, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()
os.environ["CRW_LP_PAIR"] = "0"
os.environ["CRW_TC_DEBUG"] = "8"
L = crw._lib.lib()
buf = np.zeros(160 * 8 * 10, dtype=np.uint64)
crw.ops.labelprop(feats, mask0, 20, 12.0, 0.07, 10, 0, crw.ops.PREC_BF16X3, True, False)
torch.cuda.synchronize()
L.crw_debug_lp_profile(None, 1)
crw.ops.labelprop(feats, mask0, 20, 12.0, 0.07, 10, 0, crw.ops.PREC_BF16X3, True, False)
torch.cuda.synchronize()
L.crw_debug_lp_profile(buf.ctypes.data_as(ctypes.c_void_p), 1)
p = buf.reshape(160, 8, 10).astype(np.float64)[:148]
names = ["wait acc_full", "ld+park+mask", "validity", "insert loop", "merge+finish", "iterations",
         "  m: partner", "  m: merge", "  m: finish", "  m: last bar"]
print("per-warp totals over the launch (cycles; mean / max over the 148 x 8 epilogue warps):")
for i, nm in enumerate(names):
    print(f"  {nm:14s} mean {p[:, :, i].mean():10.0f}   max {p[:, :, i].max():10.0f}   part0 {p[:, :4, i].mean():10.0f}   part1 {p[:, 4:, i].mean():10.0f}")
tot = p[:, :, :5].sum(2)
print(f"  sum of phases  mean {tot.mean():10.0f}   max {tot.max():10.0f}   (1 us = ~1965 cycles)")
print("  part 0 vs part 1 insert-loop cycles:", p[:, :4, 3].mean(), p[:, 4:, 3].mean())
print("  cycles per insertion-loop iteration:", p[:, :, 3].sum() / max(p[:, :, 5].sum(), 1))
# load balance: per-CTA totals (slowest warp of the CTA bounds it), 4-tile vs 3-tile CTAs of the bulk launch (grid 139, 470 tiles)
cta_max = tot.max(1)
cta_mean = tot.mean(1)
n4 = 470 - 3 * 139
print(f"  CTAs with 4 tiles (0..{n4 - 1}): slowest warp mean {cta_max[9:n4].mean():9.0f}  warp mean {cta_mean[9:n4].mean():9.0f}")
print(f"  CTAs with 3 tiles ({n4}..138): slowest warp mean {cta_max[n4:139].mean():9.0f}  warp mean {cta_mean[n4:139].mean():9.0f}")
print(f"  slowest CTA {cta_max[:139].max():9.0f}; within-CTA (slowest warp / mean warp) {np.mean(cta_max[9:139] / cta_mean[9:139]):.3f}")
by_g = tot[9:139].reshape(-1, 2, 4)          # [cta, part, quadrant]
print("  mean by quadrant (part 0):", np.round(by_g[:, 0, :].mean(0)), " (part 1):", np.round(by_g[:, 1, :].mean(0)))
# tail split (3 whole tiles + a half on CTAs 0..105, 3 whole tiles on the rest)
print(f"  split view: CTAs 9..52 {cta_max[9:53].mean():9.0f}   53..105 {cta_max[53:106].mean():9.0f}   106..138 {cta_max[106:139].mean():9.0f}")
for i, nm in enumerate(names):
    print(f"    {nm:14s} 9..52 {p[9:53, :, i].mean():9.0f}   53..105 {p[53:106, :, i].mean():9.0f}   106..138 {p[106:139, :, i].mean():9.0f}")
