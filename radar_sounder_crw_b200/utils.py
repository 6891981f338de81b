"""Drop-in ``propagate`` and helpers (reference: src/utils.py:15-23, 76-175).

``propagate`` keeps the reference's signature and return triple.  The per-frame Python loop of the
reference (utils.py:152-160, ~15 launches and an O(n) re-stack per frame) is replaced by one
``crw_b200::labelprop`` call; ``propagate_batch`` is the same for R radargrams at once (the shard
unit of the multi-GPU path).
"""
from __future__ import annotations

import torch

from . import ops


def create_model(id, pos_embed):
    """reference utils.py:15-23"""
    from .encoder import CNN, Resnet
    if id == 0:
        return CNN(pos_embed)
    if id == 1:
        return Resnet(pos_embed)
    raise ValueError(f"unknown model id {id}")


def pos_embed(seq):
    """Vertical-ramp positional channel (reference utils.py:76-90); device follows ``seq``."""
    BT, _, H, W = seq.shape
    pe = (torch.arange(0, H, device=seq.device, dtype=seq.dtype) / H - 0.5).view(1, 1, H, 1).expand(BT, 1, H, W)
    return torch.cat([pe, seq], dim=1)


def ndiag_matrix(size, n=1):
    """k-diagonal row-normalised matrix (reference utils.py:164-175)."""
    m = torch.zeros(size, size)
    m.diagonal(0).fill_(1)
    for i in range(0, n - 1):
        m.diagonal(i).fill_(1)
        m.diagonal(-i).fill_(1)
    return m / m.sum(dim=1, keepdim=True)


def first_column_labels(seg_ref: torch.Tensor, N: int) -> torch.Tensor:
    """``Resize((N,1), NEAREST)(seg_ref)[:,0]`` (reference utils.py:139-142) as index arithmetic.

    torch's nearest rule: src = min(floor(dst * float32(H/N)), H-1); output width 1 selects column 0.
    """
    H = seg_ref.shape[0]
    scale = torch.tensor(H, dtype=torch.float32) / torch.tensor(N, dtype=torch.float32)
    src = torch.clamp(torch.floor(torch.arange(N, dtype=torch.float32) * scale).long(), max=H - 1)
    return seg_ref[src.to(seg_ref.device), 0]


def one_hot_mask(label0: torch.Tensor, nclasses: int) -> torch.Tensor:
    """[N] class ids -> [M,N] float mask (reference utils.py:143-147)."""
    cls = torch.arange(nclasses, device=label0.device).view(-1, 1)
    return (label0.view(1, -1) == cls).float()


def seed_labels(seg: torch.Tensor, rows: int, col_start: int, col_stride: int, R: int, N: int, nclasses: int):
    """Device-side label seeding for R radargrams at once (reference utils.py:139-147 per radargram, with
    ``seg_ref = seg[:rows, col:col+W]`` of scripts/test/test_all.py:94): -> (label0 [R,N] i32, mask0 [R,M,N] f32)."""
    return ops.seed_labels(seg.float(), rows, col_start, col_stride, R, N, nclasses)


def fuse_reversed(final_pred: torch.Tensor, pred_rev: torch.Tensor, rg_len: int, dataset: int) -> torch.Tensor:
    """Reversed-pass fusion (scripts/test/test_all.py:146-158): ``pred_rev`` is the concatenated map of the
    ``use_last=True`` pass in its own (flipped) column order; returns the fused [H,W] map."""
    return ops.fuse_reversed(final_pred.float(), pred_rev.float(), rg_len, dataset)


def _change_point(xent: torch.Tensor):
    """PELT change point on the horizontality metric (reference utils.py:125-132); None without ruptures."""
    try:
        import ruptures as rpt
        diffs = (xent[:, :-1] - xent[:, 1:]).abs().sum(0)
        result = rpt.Pelt(model="rbf").fit(diffs.numpy()).predict(pen=5)
        return max(0, int(result[-2] + 5))
    except Exception:
        return None


@torch.no_grad()
def propagate_batch(emb_raw, label0, nclasses, lp, mode=None, precision=None, return_masks=False):
    """emb_raw [R,T,N,C] raw encoder output, label0 [R,N] class ids -> labels [R,N,T] (float, like the reference)."""
    R = emb_raw.shape[0]
    mask0 = torch.stack([one_hot_mask(label0[r], nclasses) for r in range(R)]).to(emb_raw.device)
    labels, masks, _, _ = ops.labelprop(
        emb_raw.float(), mask0, int(lp.cxt_size), float(lp.radius), float(lp.temperature), int(lp.topk),
        int(lp.mode if mode is None else mode), int(lp.precision if precision is None else precision), True, False)
    out = labels.transpose(1, 2).float()
    return (out, masks) if return_masks else out


@torch.no_grad()
def propagate(seq, seg_ref, model, lp, nclasses, do_pos_embed, use_last):
    """KNN label propagation over one radargram (reference utils.py:93-161).

    seq [T,N,H,W]; seg_ref [rg_h, W'] class ids; returns (final_prediction [N,T] float on the
    encoder's device, xent [N,T-1] on cpu, change_idx).
    """
    T, N, H, W = seq.shape
    if use_last:
        seq = torch.flip(seq, (0,))                                       # utils.py:108
    x = seq.reshape(-1, H, W).unsqueeze(1)
    if do_pos_embed:
        x = pos_embed(x)
    emb = ops.l2_normalize(model(x).view(T, N, -1).float())               # utils.py:114-115
    xent = ops.horizontality_xent(emb).cpu()                              # utils.py:118-123
    change_idx = _change_point(xent)
    if seg_ref.is_cuda:                                                   # utils.py:139-147 on the device
        label0, mask0 = ops.seed_labels(seg_ref.float(), seg_ref.shape[0], 0, 0, 1, N, nclasses)
        label0 = label0[0].long()
    else:
        label0 = first_column_labels(seg_ref, N).to(emb.device).long()
        mask0 = one_hot_mask(label0, nclasses)[None]
    labels, _, _, _ = ops.labelprop(emb[None], mask0, int(lp.cxt_size), float(lp.radius), float(lp.temperature),
                                    int(lp.topk), int(lp.mode), int(lp.precision), False, False)
    final_prediction = labels[0].t().float()                              # [N,T]
    final_prediction[:, 0] = label0.float()                               # utils.py:142 (raw label, not argmax)
    return final_prediction, xent, change_idx
