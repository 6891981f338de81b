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


def trim_miguel(T: torch.Tensor, length: int, dim) -> torch.Tensor:
    """Trim each of the seven concatenated MCoRDS1 ("Miguel") flight lines to a whole number of items (dataset.py:66-80)."""
    splits = [9984, 6656, 9984, 20000, 16640, 32864, 8992]
    out, start = [], 0
    for L in splits:
        item = dim[1] * length
        out.append(T[:, start:start + (L // item) * item])
        start += L
    return torch.cat(out, dim=1)


class RGDataset(torch.utils.data.Dataset):
    """Drop-in for the reference ``RGDataset`` (src/dataset.py:5-47) with the radargram resident in HBM.

    ``filepath`` may be a path (``torch.load``-ed, as the reference does) or the ``[H,W]`` tensor itself.  The
    radargram is copied to ``device`` once; ``__getitem__`` / ``get_smaller_item`` / ``items`` then cut the
    ``[T,N,h,w]`` frame sequences with the CUDA unfold kernel (``crw_patch_unfold``) instead of host-side views.
    """

    def __init__(self, filepath='/data/MCoRDS1_2010_DC8/RG2_MCoRDS1_2010_DC8.pt', length=10, dim=(24, 24), overlap=(0, 0),
                 flip=False, device="cuda"):
        from . import ops
        self._ops = ops
        self.filepath = filepath
        self.l = length
        T = torch.load(filepath) if isinstance(filepath, str) else filepath
        if isinstance(filepath, str) and filepath.endswith('rg2.pt'):
            T = trim_miguel(T, length, dim)                                # dataset.py:12-14
        if flip:
            T = torch.flip(T, dims=(1,))                                   # dataset.py:16-17
        self.T = T.to(device=device, dtype=torch.float32).contiguous()
        H, W = self.T.shape
        h, w = dim
        oh, ow = overlap
        self.nh = (H - oh) // (h - oh)
        self.nw = (W - item_width(length, w, ow)) // (w - ow) + 1
        self.oh, self.ow, self.h, self.w = oh, ow, h, w
        self.pxh = self.nh * h - oh * (self.nh - 1)
        self.pxw = item_width(length, w, ow)

    def __len__(self):
        return self.nw

    def items(self, first: int, count: int, stride_items: int = 1, length: int = None, reverse: bool = False) -> torch.Tensor:
        """``count`` items starting at dataset index ``first``, ``stride_items`` indices apart -> [count,T,N,h,w]."""
        length = self.l if length is None else length
        step = self.w - self.ow
        return self._ops.patch_unfold(self.T, step * first, step * stride_items, count, length, self.h, self.w, self.oh,
                                      self.ow, reverse)

    def __getitem__(self, index):
        if index < 0 or index >= self.nw:
            raise IndexError(index)
        return self.items(index, 1)[0]

    def get_smaller_item(self, index, small_length):
        self.small_pxw = self.pxw = item_width(small_length, self.w, self.ow)    # dataset.py:42 (the reference overwrites pxw)
        return self.items(index, 1, length=small_length)[0]
