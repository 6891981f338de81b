"""Property tests (hypothesis) of the oracle itself on CPU: the two walk formulations agree, the analytic backward
matches finite differences, and the pinned-order C port agrees with the fp64 numpy restatement over random shapes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from helpers import topk_sets_equal
from oracle import c_oracle, labelprop_oracle as lo, walk_oracle as wo

SET = settings(max_examples=12, deadline=None)


@SET
@given(B=st.integers(1, 3), T=st.integers(3, 8), N=st.integers(2, 12), C=st.integers(2, 9), seed=st.integers(0, 10_000),
       tau=st.sampled_from([0.01, 0.07, 0.5]))
def test_walk_chain_equals_reference_order(B, T, N, C, seed, tau):
    x = np.random.RandomState(seed).randn(B, T, N, C)
    A = wo.affinities(wo.l2_normalize(x), tau)
    a, b = wo.walk_loss_reference_order(A), wo.walk_loss_chain(A)
    assert abs(a - b) <= 1e-11 * max(1.0, abs(a))


@SET
@given(T=st.integers(3, 6), N=st.integers(2, 6), C=st.integers(2, 5), seed=st.integers(0, 10_000))
def test_walk_backward_matches_finite_differences(T, N, C, seed):
    rs = np.random.RandomState(seed)
    x = rs.randn(1, T, N, C)
    _, _, _, dx = wo.walk_backward_chain(x, 0.3)
    for _ in range(4):
        idx = tuple(rs.randint(0, s) for s in x.shape)
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[idx] += h
        xm[idx] -= h
        fd = (wo.crw_forward(xp, 0.3)[0] - wo.crw_forward(xm, 0.3)[0]) / (2 * h)
        assert abs(fd - dx[idx]) <= 1e-6 + 1e-4 * abs(fd)


@SET
@given(T=st.integers(1, 14), N=st.integers(8, 30), M=st.integers(2, 6), ctx=st.integers(1, 6), k=st.integers(1, 8),
       radius=st.integers(1, 12), mode=st.sampled_from(["ref_exact", "fixed"]), seed=st.integers(0, 10_000))
def test_c_port_matches_numpy_oracle(T, N, M, ctx, k, radius, mode, seed):
    rs = np.random.RandomState(seed)
    feats = rs.randn(1, T, N, 16).astype(np.float32)
    label0 = rs.randint(0, M, (1, N)).astype(np.int32)
    out = c_oracle.labelprop(feats, label0, M, ctx, radius, 0.07, k, mode=mode)
    labels, masks, W, I = lo.propagate_features(wo.l2_normalize(feats[0].astype(np.float64)), label0[0], M, ctx, radius, 0.07,
                                                k, mode=mode, return_topk=True)
    assert np.array_equal(out["labels"][0, 0], label0[0])     # frame 0 carries the reference labels
    if T == 1:
        return
    frac, _ = topk_sets_equal(out["I"][0, 1:], out["W"][0, 1:], I[1:], W[1:])
    assert frac >= 0.99                                  # fp32 vs fp64 may swap exact near-ties
    if frac == 1.0:
        assert np.abs(out["masks"][0] - masks).max() < 1e-4
        assert (out["labels"][0].T == labels).mean() >= 0.99


@SET
@given(n=st.integers(1, 40), ctx=st.integers(1, 25))
def test_key_and_label_frames(n, ctx):
    kf, lf = lo.key_frames(n, ctx), lo.label_frames(n, ctx, "ref_exact")
    assert len(kf) == min(n, ctx + 1) and kf[0] == 0 and kf == sorted(set(kf)) and all(f < n for f in kf)
    assert len(lf) == len(kf)                            # ids index equally long lists in both modes
    if n > ctx + 1:
        assert lf == list(range(ctx + 1)) and kf[1:] == list(range(n - ctx, n))   # the F5 quirk
    assert lo.label_frames(n, ctx, "fixed") == kf
