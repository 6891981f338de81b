#!/usr/bin/env python
"""Profiling target: warm-up + ONE label-propagation call at BASELINE config 3 (LP_CFG=5: one GPU's share of config 5) with
LP_PREC = tcx | bf16x3 | fp32, bracketed by an NVTX range ("lp_call") so that ncu can be restricted to it:
    ncu --nvtx --nvtx-include "lp_call/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

PREC = {"bf16x3": crw.ops.PREC_BF16X3, "tcx": crw.ops.PREC_TC_EXACT, "fp32": crw.ops.PREC_FP32}[os.environ.get("LP_PREC", "tcx")]
cfg5 = os.environ.get("LP_CFG") == "5"
R, T, N, C, M = (8, 3125, 49, 128, 4) if cfg5 else (1, 1250, 49, 128, 4)
k, r = (20, 24.0) if cfg5 else (10, 12.0)
torch.manual_seed(11)
feats = torch.randn(R, T, N, C, device="cuda")
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (R, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for _ in range(2):
    crw.ops.labelprop(feats, mask0, 20, r, 0.07, k, 0, PREC, True, False)
flush.add_(1.0)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("lp_call")
out = crw.ops.labelprop(feats, mask0, 20, r, 0.07, k, 0, PREC, True, False)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("labels checksum", int(out[0].sum()))
