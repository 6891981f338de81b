#!/usr/bin/env python
"""Profiling aid: survivor statistics of the exact tensor path's filter (how many candidates per query reach the fp32 refine,
how many lists overflowed) for a few feature distributions.  Calls the C ABI directly with its own scratch so that the
filter's survivor counts can be read back (layout: lp_x_scratch_bytes in labelprop_x.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402
from radar_sounder_crw_b200 import _lib  # noqa: E402


def align(v, a=256):
    return (v + a - 1) // a * a


def stats(name, feats, k, radius):
    R, T, N, C = feats.shape
    M = 4
    L = _lib.lib()
    mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (R, N), device="cuda"), M).permute(0, 2, 1).float().contiguous()
    labels = torch.empty((R, T, N), device="cuda", dtype=torch.int32)
    masks = torch.empty((R, T, M, N), device="cuda")
    nb = L.crw_labelprop_scratch_bytes(R, T, N, C, k, crw.ops.PREC_TC_EXACT, 1, 0)
    scratch = torch.zeros(nb, device="cuda", dtype=torch.uint8)
    rc = L.crw_labelprop_forward(feats.data_ptr(), mask0.data_ptr(), R, T, N, C, M, 20, float(radius), 0.07, k, 0, crw.ops.PREC_TC_EXACT, 1,
                                 labels.data_ptr(), masks.data_ptr(), None, None, scratch.data_ptr(), nb, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    kl = 16 if k <= 10 else (24 if k <= 16 else 32)
    rows = R * T * N
    base = (-scratch.data_ptr()) % 256
    off = base + align(512 + 64 + R * 512) + 2 * align(rows * C * 2) + align(rows * C * 4) + align(rows * kl * 4)
    cnt = scratch[off:off + rows * 4].view(torch.int32).view(R, T, N)[:, 1:]
    st = scratch[base:base + 512].view(torch.float32).view(32, 4).max(dim=0).values
    ovf = (cnt >> 30) & 1
    c = (cnt & 0xffff).float()
    print(f"{name:28s} k={k} r={radius}: survivors/query mean {c.mean():.2f} max {int(c.max())}  overflowed {int(ovf.sum())} of {cnt.numel()}"
          f"  max|x| {st[0].sqrt():.4f} res {st[1].sqrt():.2e}; centred keys: max|k - mu| {st[2].sqrt():.4f} res {st[3].sqrt():.2e}")


torch.manual_seed(3)
T, N, C = 1250, 49, 128
g = torch.randn(1, T, N, C, device="cuda")
stats("white noise features", g, 10, 12)
stats("white noise features", g, 20, 24)
stats("collinear (cos ~0.92)", 0.3 * g + torch.randn(1, 1, 1, C, device="cuda"), 10, 12)
stats("collinear (cos ~0.99)", 0.1 * g + torch.randn(1, 1, 1, C, device="cuda"), 10, 12)
stats("collinear (cos ~0.99)", 0.1 * g + torch.randn(1, 1, 1, C, device="cuda"), 20, 24)
# encoder-derived: white-noise radargram through a random-init Resnet
torch.manual_seed(11)
enc = crw.Resnet(pos_embed=False).cuda().eval()
with torch.no_grad():
    rg = torch.randn(400, 300 * 16, device="cuda")
    ds_item = rg[: 49 * 16 - 8 * 48].unfold(0, 16, 8).unfold(1, 16, 16).permute(1, 0, 2, 3).contiguous()    # [T,N,16,16]
    emb = enc(ds_item.reshape(-1, 1, 16, 16)).view(1, 300, 49, 128)
stats("Resnet(random init) noise rg", emb.contiguous(), 10, 12)
stats("Resnet(random init) noise rg", emb.contiguous(), 20, 24)
