import os, sys
sys.path.insert(0, "/root/repo")
os.environ["CRW_WALK_FUSED"] = "1"
import numpy as np, torch
import radar_sounder_crw_b200 as crw
from oracle import walk_oracle
T = int(os.environ.get("TT", "6"))
x = torch.randn(2, T, 47, 128, device="cuda")
xr = x.clone().requires_grad_(True)
loss, A, _ = crw.ops.walk_loss(xr, 0.07, False, crw.ops.PREC_BF16X3)
loss.backward()
torch.cuda.synchronize()
l, dA, dE, dx = walk_oracle.walk_backward_chain(x.double().cpu().numpy(), 0.07)
print("T", T, "loss", float(loss), l, "dx err", np.abs(xr.grad.double().cpu().numpy() - dx).max() / np.abs(dx).max())
