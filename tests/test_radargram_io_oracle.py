"""oracle/radargram_oracle.py against the fixtures generated from the live reference (CPU only)."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import radargram_oracle as ro


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_patch_unfold_matches_reference_dataset(tag):
    g = load_golden("io_unfold.npz")
    rg = g[f"{tag}_rg"].astype(np.float32)
    length, h, w, oh, ow, flip = (int(v) for v in g[f"{tag}_geom"])
    src = rg[:, ::-1] if flip else rg                                   # dataset.py:16-17
    assert ro.dataset_len(rg.shape[1], length, w, ow) == int(g[f"{tag}_len"])
    for i, idx in enumerate(g[f"{tag}_idx"]):
        got = ro.patch_unfold(src, int(idx), length, (h, w), (oh, ow))
        assert np.array_equal(got, g[f"{tag}_items"][i].astype(np.float32))
    assert np.array_equal(ro.patch_unfold(src, 1, 2, (h, w), (oh, ow)), g[f"{tag}_small"].astype(np.float32))


def test_patch_unfold_reverse_is_frame_flip():
    rs = np.random.RandomState(0)
    rg = rs.randn(40, 200).astype(np.float32)
    a = ro.patch_unfold(rg, 3, 6, (16, 16), (8, 0))
    assert np.array_equal(ro.patch_unfold(rg, 3, 6, (16, 16), (8, 0), reverse=True), a[::-1])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_seed_labels_matches_reference_resize(tag):
    g = load_golden("io_seed.npz")
    rows, N, M, W = (int(v) for v in g[f"{tag}_geom"])
    seg = g[f"{tag}_seg"].astype(np.float32)
    for r, col in enumerate((0, W, 2 * W)):
        label0, mask0 = ro.seed_labels(seg, rows, col, N, M)
        assert np.array_equal(label0, g[f"{tag}_label0"][r].astype(np.int64))
        assert np.array_equal(mask0, g[f"{tag}_mask0"][r])


@pytest.mark.parametrize("rule", [0, 1, 3])
def test_fuse_reversed_matches_reference_statements(rule):
    g = load_golden("io_fuse.npz")
    out = ro.fuse_reversed(g[f"r{rule}_fwd"].astype(np.float32), g[f"r{rule}_rev"].astype(np.float32), int(g[f"r{rule}_rg_len"]), rule)
    assert np.array_equal(out, g[f"r{rule}_out"].astype(np.float32))
    assert (out != g[f"r{rule}_fwd"]).any()                              # the fixture exercises the overwrite
