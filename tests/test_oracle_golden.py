"""The oracle (numpy + pinned-order C) against the golden vectors produced by the live reference.

Runs on CPU (`-m "not gpu"`).  This is the pin that lets the GPU parity tests trust the oracle.
"""
import numpy as np
import pytest

from helpers import load_golden, lp_case, order_mismatches_are_ties, rel_err, topk_sets_equal
from oracle import c_oracle, labelprop_oracle as lo, walk_oracle as wo

WALK = ["walk_small_f64.npz", "walk_t3_f64.npz", "walk_cfg1_f32.npz", "walk_tau001_f32.npz", "walk_cfg4_f32.npz"]
LP = ["lp_quirk.npz", "lp_cfg3_short.npz", "lp_masked_ties.npz", "lp_clustered.npz", "lp_cfg5_short.npz"]


@pytest.mark.parametrize("name", WALK)
def test_walk_oracle_matches_reference(name):
    g = load_golden(name)
    dt = np.float64 if str(g["dtype"]) == "float64" else np.float32
    x = g["x"].astype(dt)
    tau = float(g["tau"])
    E = wo.l2_normalize(x)
    A = wo.affinities(E, tau)
    tol = 1e-12 if dt == np.float64 else 2e-5
    assert rel_err(A, g["A"]) < tol
    l_ref_order = wo.walk_loss_reference_order(A)
    l_chain, dA, dE, dx = wo.walk_backward_chain(x, tau)
    assert abs(l_ref_order - float(g["loss"])) <= tol * abs(float(g["loss"]))
    assert abs(l_chain - float(g["loss"])) <= tol * abs(float(g["loss"]))
    # gradient parity is the real test; fp32 fixtures carry the reference's own fp32 noise
    assert rel_err(dx, g["dx"]) < (1e-10 if dt == np.float64 else 2e-3)


def test_walk_oracle_f64_tightens_f32_reference():
    """The fp64 oracle on the fp32 fixture: shows the fixture's own error is ~1e-4, not the oracle's."""
    g = load_golden("walk_cfg1_f32.npz")
    l, _, _, dx = wo.walk_backward_chain(g["x"].astype(np.float64), float(g["tau"]))
    assert abs(l - float(g["loss"])) < 1e-5
    assert rel_err(dx, g["dx"]) < 2e-3


def test_walk_unused_affinities_have_zero_grad():
    """S_0 and A_{T-2} never enter the loss."""
    rs = np.random.RandomState(0)
    x = rs.randn(2, 6, 5, 8)
    _, dA, _, _ = wo.walk_backward_chain(x, 0.07)
    assert np.all(dA[:, -1] == 0)


def test_walk_t2_is_zero():
    x = np.random.RandomState(1).randn(1, 2, 5, 8)
    l, A = wo.crw_forward(x, 0.07)
    assert l == 0.0 and A.shape == (1, 1, 5, 5)


@pytest.mark.parametrize("name", LP)
@pytest.mark.parametrize("impl", ["numpy64", "c_f32"])
def test_lp_oracle_matches_reference(name, impl):
    g = lp_case(name)
    T, N, C = g["feats"].shape
    if impl == "numpy64":
        emb = wo.l2_normalize(g["feats"].astype(np.float64))
        labels, masks, W, I = lo.propagate_features(emb, g["label0"], g["M"], g["ctx"], g["radius"], g["temp"],
                                                    g["k"], return_topk=True)
        labels = labels  # [N,T]
    else:
        out = c_oracle.labelprop(g["feats"][None], g["label0"][None], g["M"], g["ctx"], g["radius"], g["temp"], g["k"])
        labels, masks, W, I = out["labels"][0].T, out["masks"][0], out["W"][0], out["I"][0]
    assert np.array_equal(labels, g["labels"].astype(labels.dtype)), "propagated labels must be identical"
    frac, _ = topk_sets_equal(I[1:], W[1:], g["I"], g["W"])
    assert frac == 1.0
    n_mis, ok = order_mismatches_are_ties(I[1:], g["I"], g["W"])
    assert ok, "sorted order of live ids may differ only between near-equal weights"
    assert n_mis <= 0.002 * g["I"].size
    assert np.abs(W[1:] - g["W"]).max() < 2e-6
    assert np.abs(masks[1:] - g["masks"]).max() < 5e-6


def test_lp_quirk_is_needed():
    """mode='fixed' (the 'intended' gather) does NOT reproduce the reference once n > ctx+1."""
    g = lp_case("lp_quirk.npz")
    out = c_oracle.labelprop(g["feats"][None], g["label0"][None], g["M"], g["ctx"], g["radius"], g["temp"], g["k"],
                             mode="fixed")
    agree = (out["labels"][0].T == g["labels"]).mean()
    assert agree < 0.999
    # ... but the two modes agree on the first ctx+1 frames
    assert np.array_equal(out["labels"][0].T[:, : g["ctx"] + 2], g["labels"][:, : g["ctx"] + 2])


def test_xent_oracle_matches_reference():
    g = lp_case("lp_cfg3_short.npz")
    emb = wo.l2_normalize(g["feats"])
    x = lo.horizontality_xent(emb)
    assert x.shape == g["xent"].shape
    assert np.abs(x - g["xent"]).max() < 1e-4


def test_first_column_labels_matches_torch_nearest():
    import torch
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode
    for H, N in [(400, 49), (400, 47), (410, 48), (912, 113), (1000, 485), (64, 64), (50, 7)]:
        seg = torch.randint(0, 5, (H, 9))
        down = transforms.Resize((N, 1), interpolation=InterpolationMode.NEAREST)
        ref = down(seg.unsqueeze(0)).squeeze(0)[:, 0].numpy()
        assert np.array_equal(lo.first_column_labels(seg.numpy(), N), ref), (H, N)


def test_pinned_expf_accuracy():
    x = -np.abs(np.random.RandomState(3).randn(5000).astype(np.float32)) * 15
    x = x[x > -80]
    e = c_oracle.expf(x)
    ref = np.exp(x.astype(np.float64))
    assert (np.abs(e - ref) / ref).max() < 2.5e-7
    assert c_oracle.expf(np.array([0.0], np.float32))[0] == 1.0
    assert c_oracle.expf(np.array([-1e10 / 0.07], np.float32))[0] == 0.0


@pytest.mark.parametrize("name", ["walk_small_f64.npz", "walk_cfg1_f32.npz", "walk_tau001_f32.npz", "walk_cfg4_f32.npz"])
def test_torch_port_of_the_walk_matches_the_live_reference(name):
    """oracle/walk_torch_port.crw_loss_reference_order (the restatement whose autograd the GPU tests compare against and the
    port leg of bench.py's CPU arm) against the outputs of the LIVE reference: loss, returned affinities and d loss / d x."""
    import torch
    from oracle.walk_torch_port import crw_loss_reference_order
    g = load_golden(name)
    dt = torch.float64 if str(g["dtype"]) == "float64" else torch.float32
    x = torch.tensor(g["x"].astype(np.float64 if dt == torch.float64 else np.float32), dtype=dt, requires_grad=True)
    loss, A = crw_loss_reference_order(x, float(g["tau"]))
    loss.backward()
    tol = 1e-12 if dt == torch.float64 else 2e-6
    assert abs(float(loss) - float(g["loss"])) <= tol * max(abs(float(g["loss"])), 1.0)
    assert np.abs(A.detach().numpy() - g["A"]).max() <= tol * max(np.abs(g["A"]).max(), 1.0)
    assert np.abs(x.grad.numpy() - g["dx"]).max() <= (1e-10 if dt == torch.float64 else 2e-6) * max(np.abs(g["dx"]).max(), 1e-30)
