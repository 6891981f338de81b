#!/usr/bin/env python
"""One-off numerical check of the large-N tile engine at a bench geometry (default B=32 T=20 N=369: 68 tiles per persistent
CTA in DsProb) against the fp32 FMA engine: loss, returned A and dx."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

B, T, N = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 20, 369)))
torch.manual_seed(0)
x = torch.randn(B, T, N, 128, device="cuda") + torch.randn(B, 1, 1, 128, device="cuda")
out = {}
for name, prec in (("tiles", crw.ops.PREC_BF16X3), ("fp32", crw.ops.PREC_FP32)):
    xt = x.clone().requires_grad_(True)
    loss, A, _ = crw.ops.walk_loss(xt, 0.07, True, prec)
    loss.backward()
    torch.cuda.synchronize()
    out[name] = (loss.item(), A.detach(), xt.grad.detach())
    del loss, A, xt
    torch.cuda.empty_cache()


def rel(a, b):
    return ((a - b).double().norm() / b.double().norm()).item()


print(f"B={B} T={T} N={N}: loss tiles {out['tiles'][0]:.7f} fp32 {out['fp32'][0]:.7f}  rel {abs(out['tiles'][0] - out['fp32'][0]) / abs(out['fp32'][0]):.2e}")
print(f"A rel err {rel(out['tiles'][1], out['fp32'][1]):.2e}   dx rel err {rel(out['tiles'][2], out['fp32'][2]):.2e}   finite {bool(torch.isfinite(out['tiles'][2]).all())}")
