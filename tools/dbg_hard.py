import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, numpy as np
import radar_sounder_crw_b200 as crw
from radar_sounder_crw_b200 import _lib
from test_gpu_hard_cases import _features
def align(v, a=256): return (v + a - 1) // a * a
for kind in ["white_noise", "layered"]:
  for k, radius in [(10, 12.0), (20, 24.0)]:
    emb, N = _features(crw, kind, 160, (16, 16), (8, 0))
    R, T, _, C = emb.shape; M = 4
    g = torch.Generator(device="cuda").manual_seed(3)
    label0 = torch.randint(0, M, (1, N), device="cuda", generator=g)
    mask0 = torch.nn.functional.one_hot(label0, M).permute(0, 2, 1).float().contiguous()
    l32, m32, W32, I32 = crw.ops.labelprop(emb, mask0, 20, radius, 0.07, k, 0, crw.ops.PREC_FP32, True, True)
    L = _lib.lib()
    labels = torch.empty((R, T, N), device="cuda", dtype=torch.int32); masks = torch.empty((R, T, M, N), device="cuda")
    W = torch.zeros((R, T, k, N), device="cuda"); I = torch.zeros((R, T, k, N), device="cuda", dtype=torch.int32)
    nb = L.crw_labelprop_scratch_bytes(R, T, N, C, k, crw.ops.PREC_TC_EXACT, 1, 1)
    scratch = torch.zeros(nb, device="cuda", dtype=torch.uint8)
    rc = L.crw_labelprop_forward(emb.data_ptr(), mask0.data_ptr(), R, T, N, C, M, 20, float(radius), 0.07, k, 0, crw.ops.PREC_TC_EXACT, 1,
                                 labels.data_ptr(), masks.data_ptr(), W.data_ptr(), I.data_ptr(), scratch.data_ptr(), nb, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    kl = 16 if k <= 10 else (24 if k <= 16 else 32)
    rows = R * T * N
    base = (-scratch.data_ptr()) % 256
    off = base + align(512 + R * 512) + 2 * align(rows * C * 2) + align(rows * C * 4) + align(rows * kl * 4)
    cnt = scratch[off:off + rows * 4].view(torch.int32).view(T, N)
    ovf = (cnt >> 30) & 1
    badI = (I[0, 1:] != I32[0, 1:]).any(dim=1)   # [T-1, N]
    badW = (W[0, 1:] != W32[0, 1:]).any(dim=1)
    print(kind, k, "bad I", int(badI.sum()), "bad W", int(badW.sum()), "of", badI.numel(), "ovf total", int(ovf.sum()),
          "bad&ovf", int((badI & (ovf[1:] == 1)).sum()), "bad&!ovf", int((badI & (ovf[1:] == 0)).sum()))
    if badI.any():
        idx = badI.nonzero()[:5]
        for t, q in idx.tolist():
            n = t + 1
            print("  n", n, "q", q, "cnt", int(cnt[n, q] & 0xffff), "ovf", int(ovf[n, q]))
            print("   I   ", I[0, n, :, q].tolist()); print("   I32 ", I32[0, n, :, q].tolist())
            print("   W   ", [f"{x:.6g}" for x in W[0, n, :, q].tolist()]); print("   W32 ", [f"{x:.6g}" for x in W32[0, n, :, q].tolist()])
