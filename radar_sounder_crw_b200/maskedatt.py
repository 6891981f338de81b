"""Drop-in ``batched_affinity`` / ``MaskedAttention`` (reference: src/imported/maskedatt.py:151-175, 210-265).

The reference materialises a ``[1,1,hw,hw]`` 0/-1e10 bias tensor and adds it to a dense affinity;
here the radius is a band predicate inside the kernel.  Only the node grid the reference's call
sites use is supported: ``h = N, w = 1`` (src/utils.py:148,153).
"""
from __future__ import annotations

import torch

from . import ops


class MaskedAttention(torch.nn.Module):
    """Spatial-radius mask on an (H, W) node grid (reference maskedatt.py:210-265, flat=False path)."""

    def __init__(self, radius, flat=True):
        super().__init__()
        self.radius = radius
        self.flat = flat
        self.masks = {}

    def make(self, H, W):
        if self.flat:
            H, W = int(H ** 0.5), int(W ** 0.5)
        gy, gx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        d2 = (gy[None, None] - gy[:, :, None, None]) ** 2 + (gx[None, None] - gx[:, :, None, None]) ** 2
        D = (d2.float().sqrt() < self.radius)[None].float()
        if self.flat:
            D = D.flatten(1, 2).flatten(-2, -1)
        self.masks[f"{H}-{W}"] = D
        return D

    def mask(self, H, W):
        key = f"{H}-{W}"
        if key not in self.masks:
            self.make(H, W)
        return self.masks[key]

    def forward(self, x):
        H, W = x.shape[-2:]
        return x * self.mask(H, W)[0].to(x.device)


def _band_radius_from_bias(mask: torch.Tensor) -> int:
    """Recover r from a [.., hw, hw] 0/-1e10 band bias (|i-j| < r <=> 0); raises if it is not a band."""
    m = mask.reshape(mask.shape[-2], mask.shape[-1])
    hw = m.shape[0]
    r = int((m[0] == 0).sum().item())
    i = torch.arange(hw, device=m.device)
    band = (i[:, None] - i[None, :]).abs() < r
    if not torch.equal(m == 0, band):
        raise NotImplementedError("crw_b200.batched_affinity supports the 1-D radius band mask (h=N, w=1) only")
    return r


def batched_affinity(query, keys, mask, temperature, topk, long_mem, ctx, device=None):
    """Same call contract as the reference (maskedatt.py:151):

    query [1,C,1,hw], keys [1,C,1,n,hw], mask [1,1,hw,hw] bias (0 / -1e10) or an int/float radius,
    returns (Ws, Is): one-element lists of [k,hw] float32 weights and int64 ids into the trimmed key set
    (frame 0 + last ``ctx`` frames once n > ctx+1).
    """
    radius = mask if isinstance(mask, (int, float)) else _band_radius_from_bias(mask)
    q = query[0, :, 0].t().contiguous()[None]                       # [1,hw,C]
    n = keys.shape[3]
    kk = keys[0, :, 0].permute(1, 2, 0)                             # [n,hw,C]
    if n > ctx + 1:                                                  # trim before any arithmetic (:166-167)
        kk = torch.cat([kk[:1], kk[n - ctx:]], 0)
    kk = kk.contiguous()
    F = kk.shape[0]
    W, I = ops.affinity_topk(kk, q, F, max(F, 1), float(radius), float(temperature), int(topk), ops.PREC_FP32)
    return [W[0]], [I[0].long()]
