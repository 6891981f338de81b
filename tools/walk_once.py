#!/usr/bin/env python
"""Profiling aid: a few walk fwd+bwd calls at config 2 (target for `ncu -k regex:walk_s_ ...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

prec = crw.ops.PREC_BF16X3 if (len(sys.argv) > 1 and sys.argv[1] == "bf16x3") else crw.ops.PREC_FP32
emb = torch.randn(32, 10, 47, 128, device="cuda", requires_grad=True)
for _ in range(3):
    loss, _, _ = crw.ops.walk_loss(emb, 0.07, False, prec)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
