"""Patch geometry of the reference dataset (src/dataset.py:19-39) as pure tensor views.

No file I/O: callers hand in the radargram tensor ``[H, W]``; used by the synthetic-input
generators of tests and bench so frames have exactly the reference's layout ``[T, N, h, w]``.
"""
from __future__ import annotations

import torch


def nodes_per_frame(H: int, h: int, oh: int) -> int:
    """``N = (H - oh) // (h - oh)`` (dataset.py:22)."""
    return (H - oh) // (h - oh)


def item_width(length: int, w: int, ow: int) -> int:
    """pixels spanned by ``length`` frames (dataset.py:28)."""
    return length * w - ow * (length - 1)


def radargram_to_frames(rg: torch.Tensor, index: int, length: int, dim=(16, 16), overlap=(8, 0)) -> torch.Tensor:
    """``RGDataset.__getitem__`` (dataset.py:34-39): [H,W] -> [T=length, N, h, w] float."""
    h, w = dim
    oh, ow = overlap
    H = rg.shape[0]
    N = nodes_per_frame(H, h, oh)
    pxh = N * h - oh * (N - 1)
    pxw = item_width(length, w, ow)
    start = (w - ow) * index
    item = rg[:pxh, start:start + pxw]
    item = item.unfold(0, h, h - oh).unfold(1, w, w - ow)      # [N, T, h, w]
    return item.permute(1, 0, 2, 3).float()
