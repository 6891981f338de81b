#!/usr/bin/env python
"""Profiling aid: one tensor-path walk fwd+bwd at a scaled geometry under torch.profiler (which torch ops launch what)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw
from torch.profiler import profile, ProfilerActivity

N = int(sys.argv[1]) if len(sys.argv) > 1 else 369
emb = torch.randn(32, 20, N, 128, device="cuda", requires_grad=True)
for _ in range(2):
    loss, _, _ = crw.ops.walk_loss(emb, 0.07, False, crw.ops.PREC_BF16X3)
    loss.backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    loss, _, _ = crw.ops.walk_loss(emb, 0.07, False, crw.ops.PREC_BF16X3)
    loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
