"""CUDA radargram I/O kernels (patch unfold, label seeding, reversed-pass fusion) through the custom ops / C ABI:
bit-exact against oracle/radargram_oracle.py and the live-reference fixtures."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import radargram_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import radar_sounder_crw_b200 as crw
    return crw


def _dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_rgdataset_golden(pkg, tag):
    g = load_golden("io_unfold.npz")
    rg = torch.from_numpy(g[f"{tag}_rg"].astype(np.float32))
    length, h, w, oh, ow, flip = (int(v) for v in g[f"{tag}_geom"])
    ds = pkg.RGDataset(rg, length=length, dim=(h, w), overlap=(oh, ow), flip=bool(flip))
    assert len(ds) == int(g[f"{tag}_len"])
    for i, idx in enumerate(g[f"{tag}_idx"]):
        assert np.array_equal(ds[int(idx)].cpu().numpy(), g[f"{tag}_items"][i].astype(np.float32))
    assert np.array_equal(ds.get_smaller_item(1, 2).cpu().numpy(), g[f"{tag}_small"].astype(np.float32))
    with pytest.raises(IndexError):
        ds[len(ds)]


@pytest.mark.parametrize("geom", [
    # H, W, R, T, (h,w), (oh,ow), col_start, col_stride
    (400, 20000, 1, 1250, (16, 16), (8, 0), 0, 0),          # config 3: the whole radargram as one item
    (400, 3300, 3, 10, (32, 32), (24, 0), 32, 320),         # config-2 style training items
    (61, 517, 4, 7, (12, 10), (4, 6), 3, 29),               # unaligned: scalar path
    (40, 64, 1, 1, (16, 16), (8, 0), 48, 0),                # single frame at the right edge
])
def test_patch_unfold_vs_oracle(pkg, geom):
    H, W, R, T, (h, w), (oh, ow), c0, cs = geom
    rs = np.random.RandomState(H + W)
    rg = rs.randn(H, W).astype(np.float32)
    for reverse in (False, True):
        out = pkg.ops.patch_unfold(_dev(rg), c0, cs, R, T, h, w, oh, ow, reverse).cpu().numpy()
        for r in range(R):
            exp = ro.patch_unfold(rg[:, c0 + r * cs:], 0, T, (h, w), (oh, ow), reverse=reverse)
            assert np.array_equal(out[r], exp)


def test_patch_unfold_rejects_out_of_range(pkg):
    rg = torch.zeros(40, 100, device="cuda")
    with pytest.raises(RuntimeError):
        pkg.ops.patch_unfold(rg, 0, 0, 1, 7, 16, 16, 8, 0, False)        # 7*16 = 112 > 100 columns
    with pytest.raises(RuntimeError):
        pkg.ops.patch_unfold(rg, 90, 0, 1, 1, 16, 16, 8, 0, False)
    assert pkg.ops.patch_unfold(rg, 0, 0, 0, 5, 16, 16, 8, 0, False).shape == (0, 5, 4, 16, 16)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_seed_labels_golden(pkg, tag):
    g = load_golden("io_seed.npz")
    rows, N, M, W = (int(v) for v in g[f"{tag}_geom"])
    seg = _dev(g[f"{tag}_seg"])
    label0, mask0 = pkg.seed_labels(seg, rows, 0, W, 3, N, M)
    assert np.array_equal(label0.cpu().numpy(), g[f"{tag}_label0"].astype(np.int32))
    assert np.array_equal(mask0.cpu().numpy(), g[f"{tag}_mask0"])


def test_seed_labels_many_radargrams_vs_oracle(pkg):
    rs = np.random.RandomState(5)
    seg = rs.randint(0, 5, size=(410, 64 * 50)).astype(np.float32)
    label0, mask0 = pkg.seed_labels(_dev(seg), 400, 7, 50, 64, 49, 5)
    for r in range(64):
        l, m = ro.seed_labels(seg, 400, 7 + 50 * r, 49, 5)
        assert np.array_equal(label0[r].cpu().numpy(), l)
        assert np.array_equal(mask0[r].cpu().numpy(), m)


@pytest.mark.parametrize("rule", [0, 1, 3])
def test_fuse_reversed_golden(pkg, rule):
    g = load_golden("io_fuse.npz")
    out = pkg.fuse_reversed(_dev(g[f"r{rule}_fwd"]), _dev(g[f"r{rule}_rev"]), int(g[f"r{rule}_rg_len"]), rule)
    assert np.array_equal(out.cpu().numpy(), g[f"r{rule}_out"].astype(np.float32))


@pytest.mark.parametrize("rule", [0, 1, 3])
def test_fuse_reversed_full_size_vs_oracle(pkg, rule):
    rs = np.random.RandomState(rule)
    H, rg_len, tot = 400, 5000, 4
    fwd = rs.randint(0, 6, size=(H, rg_len * tot)).astype(np.float32)
    rev = rs.randint(0, 4, size=(H, rg_len * tot)).astype(np.float32)
    rev[:, ::7] = np.where(rs.rand(H, rev[:, ::7].shape[1]) < 0.01, 4, rev[:, ::7])
    out = pkg.fuse_reversed(_dev(fwd), _dev(rev), rg_len, rule).cpu().numpy()
    assert np.array_equal(out, ro.fuse_reversed(fwd, rev, rg_len, rule))
    with pytest.raises(RuntimeError):
        pkg.fuse_reversed(_dev(fwd), _dev(rev), 4999, rule)               # W not a whole number of radargrams
    with pytest.raises(RuntimeError):
        pkg.fuse_reversed(_dev(fwd), _dev(rev), rg_len, 2)                # no fusion rule for dataset 2


def test_propagate_with_device_seg_ref_matches_host_seg_ref(pkg):
    """`propagate` seeds labels on the device when seg_ref is a CUDA tensor: same prediction as the host path."""
    torch.manual_seed(3)
    T, N, M = 12, 25, 4
    enc = pkg.Resnet(pos_embed=False).cuda().eval()
    seq = torch.randn(T, N, 16, 16, device="cuda")
    seg_ref = torch.randint(0, M, (N * 8 + 8, 16)).float()
    lp = pkg.LabelPropVOS_CRW({'CXT_SIZE': 5, 'RADIUS': 6, 'TEMP': 0.07, 'KNN': 5})
    a, _, _ = pkg.propagate(seq, seg_ref, enc, lp, M, False, False)
    b, _, _ = pkg.propagate(seq, seg_ref.cuda(), enc, lp, M, False, False)
    assert torch.equal(a, b)
