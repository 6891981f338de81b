"""numpy oracle for k-NN label propagation -- TEST INFRASTRUCTURE ONLY.

Restates, for the ``h = N, w = 1`` node grid the reference always uses
(``src/utils.py:148,153`` builds ``[1,C,N,1]`` features):

* the radius bias          ``src/imported/maskedatt.py:232-245`` + ``labelprop.py:89-96``
* ``batched_affinity``      ``src/imported/maskedatt.py:151-175``
* ``LabelPropVOS_CRW.predict``  ``src/imported/labelprop.py:67-116``
* the frame loop of ``propagate``  ``src/utils.py:134-161``

including the context-trim gather quirk: once ``n > ctx+1`` the
top-k ids index the *trimmed* key set (frame 0 + last ``ctx`` frames) but the
labels are gathered from the *untrimmed* mask list, i.e. from frames ``0..ctx``.
``mode="ref_exact"`` reproduces that; ``mode="fixed"`` gathers from the frames
the keys came from.

The work per frame is restricted to the frames that survive the trim, so this
oracle is O(T*ctx) where the reference is O(T^2); the discarded rows never
influence the result (``maskedatt.py:166-167`` drops them before ``topk``).

Tie rule (pinned here, unspecified in ``torch.topk``): equal logits are ordered
by ascending candidate index.

Parity pin: ``tests/golden/lp_*.npz`` generated from the live reference.
"""
from __future__ import annotations

import numpy as np

MASK_BIAS = -1e10  # labelprop.py:94


def key_frames(n: int, ctx: int) -> list[int]:
    """Frames whose nodes are top-k candidates for query frame ``n`` (maskedatt.py:166-167)."""
    if n <= ctx + 1:
        return list(range(n))
    return [0] + list(range(n - ctx, n))


def label_frames(n: int, ctx: int, mode: str = "ref_exact") -> list[int]:
    """Frames whose soft masks the ids are gathered from (labelprop.py:82,106,109)."""
    if n <= ctx + 1 or mode == "fixed":
        return key_frames(n, ctx)
    if mode != "ref_exact":
        raise ValueError(mode)
    return list(range(ctx + 1))


def radius_bias(N: int, radius: float, dtype=np.float64) -> np.ndarray:
    """``[N,N]`` additive bias: 0 where ``|i-j| < radius`` else -1e10 (maskedatt.py:237-239)."""
    i = np.arange(N)
    d = np.sqrt(((i[:, None] - i[None, :]) ** 2).astype(np.float32))
    return np.where(d < radius, 0.0, MASK_BIAS).astype(dtype)


def affinity_topk(query: np.ndarray, keys: np.ndarray, radius: float, temp: float, k: int):
    """Top-k neighbours of every query node among ``keys`` (already trimmed).

    query [N,C], keys [F,N,C].  Returns (W [k,N] softmax weights, I [k,N] ids into F*N),
    sorted by descending logit, ties by ascending id.
    """
    F, N, C = keys.shape
    dt = query.dtype
    logits = keys.reshape(F * N, C) @ query.T                      # [F*N, N]  maskedatt.py:157
    logits = logits.reshape(F, N, N) + radius_bias(N, radius, dt)[None]  # :160 (every key frame)
    logits = logits.reshape(F * N, N) / dt.type(temp)              # :163-164
    order = np.argsort(-logits, axis=0, kind="stable")[:k]         # :169
    top = np.take_along_axis(logits, order, axis=0)
    e = np.exp(top - top[:1])
    W = e / e.sum(0, keepdims=True)                                # :170
    return W, order.astype(np.int64)


def first_column_labels(seg_ref: np.ndarray, N: int) -> np.ndarray:
    """``Resize((N,1), NEAREST)(seg_ref)[:,0]`` -- utils.py:139-142.

    torch's 'nearest' picks ``src = min(floor(dst * float32(in/out)), in-1)``; with
    output width 1 the column is always 0.
    """
    H = seg_ref.shape[0]
    scale = np.float32(H) / np.float32(N)
    src = np.minimum(np.floor(np.arange(N, dtype=np.float32) * scale).astype(np.int64), H - 1)
    return seg_ref[src, 0]


def one_hot_mask(label0: np.ndarray, M: int, dtype=np.float64) -> np.ndarray:
    """``mask[m,i] = 1[label0[i]==m]`` -- utils.py:143-147.  Returns [M,N]."""
    return (label0[None, :] == np.arange(M)[:, None]).astype(dtype)


def predict(feats: np.ndarray, masks: np.ndarray, cur: np.ndarray, ctx: int, radius: float,
            temp: float, k: int, mode: str = "ref_exact"):
    """One ``LabelPropVOS_CRW.predict`` call.

    feats [n,N,C] all previous frames, masks [n,M,N] all previous soft masks, cur [N,C].
    Returns (mask_n [M,N], W [k,N], I [k,N]).
    """
    n, N, _ = feats.shape
    Kf = key_frames(n, ctx)
    Lf = label_frames(n, ctx, mode)
    W, I = affinity_topk(cur, feats[Kf], radius, temp, k)
    lbl = masks[Lf]                                                # [F,M,N]
    F, M, _ = lbl.shape
    flat = lbl.transpose(1, 0, 2).reshape(M, F * N)                # labelprop.py:106
    pred = np.zeros((M, N), feats.dtype)
    for j in range(k):                                             # :109, summed in top-k order
        pred = pred + flat[:, I[j]] * W[j][None]
    return pred, W, I


def propagate_features(emb: np.ndarray, label0: np.ndarray, M: int, ctx: int, radius: float,
                       temp: float, k: int, mode: str = "ref_exact", return_topk: bool = False):
    """Frame loop of ``propagate`` (utils.py:134-161) on normalised features ``emb [T,N,C]``.

    Returns labels [N,T] (int64 class ids) and masks [T,M,N]; with ``return_topk`` also
    W [T,k,N] and I [T,k,N] (frame 0 rows are zero).
    """
    T, N, _ = emb.shape
    masks = np.zeros((T, M, N), emb.dtype)
    masks[0] = one_hot_mask(label0, M, emb.dtype)
    labels = np.zeros((N, T), np.int64)
    labels[:, 0] = label0
    Ws = np.zeros((T, k, N), emb.dtype)
    Is = np.zeros((T, k, N), np.int64)
    for n in range(1, T):
        m, W, I = predict(emb[:n], masks[:n], emb[n], ctx, radius, temp, k, mode)
        masks[n] = m
        labels[:, n] = m.argmax(0)                                 # utils.py:160
        Ws[n], Is[n] = W, I
    if return_topk:
        return labels, masks, Ws, Is
    return labels, masks


def horizontality_xent(emb: np.ndarray) -> np.ndarray:
    """The "horizontality" metric of utils.py:117-123.  emb [T,N,C] normalised -> [N,T-1].

    Note the reference slices the *last* axis of a [T,N,C] tensor (channels, not
    frames): ``A[t] = emb[t,:,:-1] emb[t,:,1:]^T / 0.1`` is an intra-frame,
    channel-shifted similarity.  The target is ``ndiag_matrix(N,1)`` = identity
    (utils.py:164-175), so ``xent[n,t] = lse_c A[t][c,n] - A[t][n,n]``.
    """
    emb = emb[:-1]                                                 # only frames 0..T-2 are used (:121)
    A = np.einsum("tnc,tmc->tnm", emb[:, :, :-1], emb[:, :, 1:]) / emb.dtype.type(0.1)
    mx = A.max(1, keepdims=True)
    lse = np.log(np.exp(A - mx).sum(1)) + mx[:, 0]                 # over c (rows of A[t])
    diag = np.einsum("tnn->tn", A)
    return (lse - diag).T
