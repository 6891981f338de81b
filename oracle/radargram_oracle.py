"""CPU restatement (numpy) of the data formats either side of the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(radar_sounder_crw_b200/) never does.

Pinned against the live reference (tests/golden/make_golden.py -> io_*.npz):
  * patch_unfold   -- ``RGDataset.__init__/__getitem__/get_smaller_item`` executed through oracle/ref_shim.py
                      (src/dataset.py:19-47), plus the ``use_last`` flip of src/utils.py:108;
  * seed_labels    -- torchvision ``Resize((N,1), NEAREST)`` and the one-hot loop exactly as src/utils.py:139-147;
  * fuse_reversed  -- scripts/test/test_all.py:146-158.  That code sits inside the script's ``main()`` (argparse, file
                      I/O) and cannot be imported; the generator re-executes those statements verbatim in meaning with
                      the same torch calls (unfold / flip / view / logical_and / all / repeat), so this one is pinned by
                      a torch restatement, not by the running script.
"""
from __future__ import annotations

import numpy as np

from .labelprop_oracle import first_column_labels, one_hot_mask


def nodes_per_frame(H: int, h: int, oh: int) -> int:
    """dataset.py:22"""
    return (H - oh) // (h - oh)


def dataset_len(W: int, length: int, w: int, ow: int) -> int:
    """dataset.py:23-24"""
    return (W - (length * (w - ow) + ow)) // (w - ow) + 1


def patch_unfold(rg: np.ndarray, index: int, length: int, dim, overlap, reverse: bool = False) -> np.ndarray:
    """dataset.py:34-39: [H,W] -> [T=length, N, h, w] float32 (frames flipped when ``reverse``, utils.py:108)."""
    h, w = dim
    oh, ow = overlap
    H = rg.shape[0]
    N = nodes_per_frame(H, h, oh)
    pxw = length * w - ow * (length - 1)
    start = (w - ow) * index
    assert start + pxw <= rg.shape[1]
    out = np.empty((length, N, h, w), dtype=np.float32)
    for t in range(length):
        for n in range(N):
            r0, c0 = n * (h - oh), start + t * (w - ow)
            out[t, n] = rg[r0:r0 + h, c0:c0 + w]
    return out[::-1].copy() if reverse else out


def seed_labels(seg: np.ndarray, rows: int, col: int, N: int, M: int):
    """utils.py:139-147 with seg_ref = seg[:rows, col:col+W] (test_all.py:94): (label0 [N] int, mask0 [M,N] f32)."""
    label0 = first_column_labels(seg[:rows, col:col + 1], N).astype(np.int64)
    return label0, one_hot_mask(label0, M, dtype=np.float32)


def fuse_reversed(fwd: np.ndarray, rev: np.ndarray, rg_len: int, rule: int) -> np.ndarray:
    """test_all.py:146-158.  fwd, rev [H,W]; rev in the reversed pass's column order."""
    H, W = fwd.shape
    assert W % rg_len == 0
    prev = rev.reshape(H, W // rg_len, rg_len)[:, :, ::-1].reshape(H, W)        # unfold / flip / view
    mask = prev == 2
    if rule == 1:
        mask = mask & (fwd != 3) & np.all(prev != 4, axis=0)[None, :]
    elif rule == 3:
        flat = mask.reshape(-1).copy()
        flat[:flat.size // 2] = False
        mask = flat.reshape(H, W)
    elif rule != 0:
        raise ValueError(rule)
    out = fwd.copy()
    out[mask] = 2
    return out
