"""CPU-side checks: the C-ABI library loads and exports every symbol include/crw_b200.h declares,
the torch ops are registered with fake (meta) kernels, and the host-side mirrors behave."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import radar_sounder_crw_b200 as p
    return p


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "crw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    import ctypes
    L = ctypes.CDLL(pkg._lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/crw_b200.h but not exported"
    # and the Python binding table covers the header one to one
    assert sorted(pkg._lib.SIGNATURES) == names


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "crw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # declarations only
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src


def test_version_and_error_strings(pkg):
    L = pkg._lib.lib()
    assert L.crw_version() >= 100 and L.crw_built_arch() == 100
    assert pkg._lib.error_string(0) == "ok"
    assert "workspace" in pkg._lib.error_string(-4)
    # size queries are pure host functions
    assert L.crw_walk_saved_bytes(32, 10, 47, 128, 0) > 32 * 9 * 47 * 47 * 4 * 6
    assert L.crw_walk_saved_bytes(32, 10, 47, 128, 1) >= L.crw_walk_saved_bytes(32, 10, 47, 128, 0)
    assert L.crw_walk_saved_bytes(0, 10, 47, 128, 0) == 0
    assert L.crw_labelprop_scratch_bytes(1, 1250, 49, 128, 10, 0, 1, 0) > 1250 * 49 * 128 * 4


def test_ops_refuse_cpu_tensors(pkg):
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.ops.l2_normalize(torch.randn(3, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.ops.walk_loss(torch.randn(1, 4, 5, 8), 0.07, False, 0)


def test_fake_kernels_give_shapes(pkg):
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(2, 6, 9, 16, device="cuda")
        loss, A, _ = torch.ops.crw_b200.walk_loss(x, 0.07, True, 0)
        assert loss.shape == () and A.shape == (2, 5, 9, 9)
        f = torch.empty(3, 20, 49, 128, device="cuda")
        m0 = torch.empty(3, 4, 49, device="cuda")
        labels, masks, W, I = torch.ops.crw_b200.labelprop(f, m0, 20, 12.0, 0.07, 10, 0, 0, True, True)
        assert labels.shape == (3, 20, 49) and labels.dtype == torch.int32
        assert masks.shape == (3, 20, 4, 49) and W.shape == (3, 20, 10, 49) and I.dtype == torch.int32


def test_host_helpers_match_oracle(pkg):
    from oracle import labelprop_oracle as lo
    from radar_sounder_crw_b200 import utils
    for H, N in [(400, 49), (400, 47), (912, 113), (1000, 485), (50, 7)]:
        seg = torch.randint(0, 5, (H, 3))
        assert np.array_equal(utils.first_column_labels(seg, N).numpy(), lo.first_column_labels(seg.numpy(), N))
    l0 = torch.tensor([0, 2, 1, 2])
    assert np.array_equal(utils.one_hot_mask(l0, 3).numpy(), lo.one_hot_mask(l0.numpy(), 3, np.float32))
    assert torch.equal(utils.ndiag_matrix(5, 1), torch.eye(5))
    pe = utils.pos_embed(torch.zeros(2, 1, 4, 3))
    assert pe.shape == (2, 2, 4, 3) and torch.allclose(pe[0, 0, :, 0], torch.arange(4) / 4 - 0.5)


def test_dataset_geometry_matches_reference_formula(pkg):
    from radar_sounder_crw_b200.dataset import nodes_per_frame, radargram_to_frames
    assert nodes_per_frame(400, 32, 24) == 47 and nodes_per_frame(400, 16, 8) == 49
    assert nodes_per_frame(912, 16, 8) == 113 and nodes_per_frame(1000, 32, 30) == 485
    rg = torch.arange(400 * 400, dtype=torch.float32).view(400, 400)
    fr = radargram_to_frames(rg, 2, 5, (16, 16), (8, 0))
    assert fr.shape == (5, 49, 16, 16)
    assert torch.equal(fr[1, 3], rg[3 * 8:3 * 8 + 16, 2 * 16 + 16:2 * 16 + 32])


def test_band_radius_recovered_from_bias_tensor(pkg):
    from oracle import labelprop_oracle as lo
    from radar_sounder_crw_b200.maskedatt import MaskedAttention, _band_radius_from_bias
    assert _band_radius_from_bias(torch.tensor(lo.radius_bias(49, 12, np.float32))[None, None]) == 12
    D = MaskedAttention(12, flat=False).mask(49, 1)[None].flatten(-4, -3).flatten(-2)   # as labelprop.py:92-93
    assert torch.equal(D[0, 0] == 1, torch.tensor(lo.radius_bias(49, 12)) == 0)
    with pytest.raises(NotImplementedError):
        _band_radius_from_bias((torch.eye(6) * -1e10)[None, None])


def test_encoder_state_dict_keys_are_reference_compatible(pkg):
    keys = set(pkg.Resnet(pos_embed=False).state_dict())
    for k in ["fc0.weight", "fc0.bias", "bn0.weight", "model.conv1.weight", "model.layer4.0.downsample.0.weight",
              "model.fc.weight"]:
        assert k in keys
    out = pkg.Resnet(pos_embed=False).eval()(torch.randn(2, 1, 16, 16))
    assert out.shape == (2, 128)
    assert pkg.CNN(False)(torch.randn(2, 1, 16, 16)).shape == (2, 128)


def test_generated_toplist_insert_is_current():
    """csrc/toplist_insert.inc is generated by tools/gen_toplist_insert.py: the committed file must be the generator's output."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_toplist_insert.py")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(root, "radar_sounder_crw_b200", "csrc", "toplist_insert.inc")) as fh:
        assert fh.read() == out


@pytest.mark.parametrize("n_tiles,grid", [(470, 139), (479, 148), (145, 139), (9, 9), (100, 148), (444, 148), (498, 130),
                                          (9696, 148), (424, 130), (1, 1), (297, 148), (0, 148)])
def test_lp_tail_split_schedule_covers_every_tile_once(pkg, n_tiles, grid):
    """Host mirror of the device scheduler of the tensor-path top-k: every tile is one whole item or exactly two halves (0 and 1) on
    different CTAs; halves come first, one per CTA; a CTA never owns more than ceil(items / grid) items."""
    import ctypes
    L = pkg._lib.lib()
    for split_ok in (0, 1):
        n = L.crw_debug_lp_schedule(n_tiles, grid, split_ok, None, 0)
        assert n >= n_tiles
        buf = (ctypes.c_int * (3 * max(n, 1)))()
        assert L.crw_debug_lp_schedule(n_tiles, grid, split_ok, ctypes.cast(buf, ctypes.c_void_p), n) == n
        items = np.frombuffer(buf, dtype=np.int32)[: 3 * n].reshape(n, 3)
        whole = items[items[:, 1] < 0]
        halves = items[items[:, 1] >= 0]
        assert sorted(whole[:, 0].tolist() + sorted(set(halves[:, 0].tolist()))) == list(range(n_tiles))
        if split_ok == 0:
            assert len(halves) == 0 and n == n_tiles
        for tile in set(halves[:, 0].tolist()):
            rows = np.nonzero((items[:, 0] == tile) & (items[:, 1] >= 0))[0]
            assert sorted(items[rows, 1].tolist()) == [0, 1]
            assert len({int(r) % grid for r in rows}) == 2                 # the two halves run on different CTAs
            assert len(set(items[rows, 2].tolist())) == 1 and 0 <= items[rows[0], 2] < 74
        if len(halves):
            assert n_tiles % grid != 0 and 2 * (n_tiles % grid) <= grid and n_tiles >= grid
            assert (items[: len(halves), 1] >= 0).all() and len(halves) <= grid  # first items, at most one per CTA
            assert len(set(halves[:, 2].tolist())) == len(halves) // 2
