"""Shared comparison helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def lp_case(name):
    """Golden LP fixture -> dict with float32 feats [T,N,C], label0 [N] and the reference outputs."""
    g = load_golden(name)
    g["feats"] = g["feats"].astype(np.float32)
    N = g["feats"].shape[1]
    H = g["seg_col0"].shape[0]
    scale = np.float32(H) / np.float32(N)
    src = np.minimum(np.floor(np.arange(N, dtype=np.float32) * scale).astype(np.int64), H - 1)
    g["label0"] = g["seg_col0"][src].astype(np.int32)
    for key in ("M", "ctx", "k"):
        g[key] = int(g[key])
    g["radius"] = float(g["radius"])
    g["temp"] = float(g["temp"])
    return g


def topk_sets_equal(I_a, W_a, I_b, W_b, w_eps=1e-30):
    """Tie-aware top-k comparison.

    I_* [..., k, N] ids, W_* [..., k, N] softmax weights.  Ids whose weight is
    exactly 0 on both sides are masked-out candidates (logit -1e10/temp): their
    identity is unspecified by torch.topk, so only ids with weight > w_eps are
    compared, as sets per query.  Returns the fraction of queries whose sets agree.
    """
    I_a, I_b = np.asarray(I_a), np.asarray(I_b)
    W_a, W_b = np.asarray(W_a), np.asarray(W_b)
    k = I_a.shape[-2]
    a = np.where(W_a > w_eps, I_a, -1)
    b = np.where(W_b > w_eps, I_b, -1)
    a = np.sort(np.moveaxis(a, -2, -1), axis=-1)
    b = np.sort(np.moveaxis(b, -2, -1), axis=-1)
    same = (a == b).all(-1)
    assert a.shape[-1] == k
    return float(same.mean()), same


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def order_mismatches_are_ties(I_a, I_b, W_b, rtol=1e-4, w_eps=1e-30):
    """Where the sorted id lists differ, the reference weights of the swapped entries must be near-equal.

    I_a/I_b/W_b [..., k, N].  For every position p with I_a != I_b (and a live weight) find where I_a[p]
    sits in I_b's column and require |W_b[p] - W_b[there]| <= rtol * W_b[p].  Returns (n_mismatch, ok).
    """
    I_a = np.moveaxis(np.asarray(I_a), -2, -1).reshape(-1, I_a.shape[-2])
    I_b = np.moveaxis(np.asarray(I_b), -2, -1).reshape(-1, I_b.shape[-2])
    W_b = np.moveaxis(np.asarray(W_b), -2, -1).reshape(-1, W_b.shape[-2])
    bad = 0
    rows, cols = np.nonzero((I_a != I_b) & (W_b > w_eps))
    for r, p in zip(rows, cols):
        there = np.nonzero(I_b[r] == I_a[r, p])[0]
        if len(there) == 0 or abs(W_b[r, p] - W_b[r, there[0]]) > rtol * W_b[r, p]:
            bad += 1
    return len(rows), bad == 0
