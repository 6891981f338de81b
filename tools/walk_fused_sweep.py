import os, sys
sys.path.insert(0, "/root/repo")
import torch
import radar_sounder_crw_b200 as crw
torch.manual_seed(0)
worst = 0.0
for B in [1, 5, 36, 37, 38, 64, 96, 100, 160]:
    for T, N in [(3, 8), (7, 47), (10, 64), (4, 33)]:
        x = torch.randn(B, T, N, 128, device="cuda")
        outs = []
        for prec in (crw.ops.PREC_BF16X3, crw.ops.PREC_FP32):
            xr = x.clone().requires_grad_(True)
            loss, A, _ = crw.ops.walk_loss(xr, 0.07, True, prec)
            (loss + 0.01 * (A * A).mean()).backward()
            outs.append((loss.item(), A.detach(), xr.grad))
        torch.cuda.synchronize()
        el = abs(outs[0][0] - outs[1][0]) / abs(outs[1][0])
        eA = (outs[0][1] - outs[1][1]).abs().max().item() / outs[1][1].abs().max().item()
        eg = (outs[0][2] - outs[1][2]).abs().max().item() / outs[1][2].abs().max().item()
        worst = max(worst, el, eA, eg)
        if max(el, eA, eg) > 1e-4 or not torch.isfinite(outs[0][2]).all():
            print("MISMATCH", B, T, N, el, eA, eg)
print("sweep done, worst relative difference", worst)
