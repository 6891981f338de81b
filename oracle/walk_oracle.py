"""numpy oracle for the CRW training walk -- TEST INFRASTRUCTURE ONLY.

Restates ``/root/reference/src/model.py:22-46`` (everything after the encoder
call).  Two forward formulations are given and must agree:

* :func:`walk_loss_reference_order` follows the reference's nested loop
  literally (``model.py:35-45``): for every ``k`` rebuild the 2k-long
  palindrome, skip its index 0, left-multiply row-softmaxes.
* :func:`walk_loss_chain` is the O(T) factorisation
  ``M_k = L_k R_k`` that the CUDA kernel runs.

:func:`walk_backward_chain` is the analytic reverse pass (Appendix A.3) giving
``d loss / d x`` for the *un-normalised* encoder output ``x``.

Parity pin: checked against the live reference (forward) and against torch
autograd through the live reference (backward) by
``tests/test_oracle_vs_reference.py`` and the committed
``tests/golden/walk_*.npz`` fixtures.
"""
from __future__ import annotations

import numpy as np

EPS_NORM = 1e-12  # torch.nn.functional.normalize default eps (model.py:22)


def l2_normalize(x: np.ndarray) -> np.ndarray:
    """``F.normalize(x, dim=-1)`` -- model.py:22, utils.py:115."""
    nrm = np.sqrt((x * x).sum(-1, keepdims=True))
    return x / np.maximum(nrm, x.dtype.type(EPS_NORM))


def row_softmax(a: np.ndarray) -> np.ndarray:
    """``softmax(a, dim=-1)`` -- model.py:44."""
    m = a.max(-1, keepdims=True)
    e = np.exp(a - m)
    return e / e.sum(-1, keepdims=True)


def affinities(emb: np.ndarray, tau: float) -> np.ndarray:
    """Stride-1 affinities ``A[b,t] = E[b,t] E[b,t+1]^T / tau`` -- model.py:26.

    emb: [B,T,N,C] already L2-normalised.  Returns [B,T-1,N,N].
    """
    return np.einsum("btnc,btmc->btnm", emb[:, :-1], emb[:, 1:]) / emb.dtype.type(tau)


def _cycle_ce(M: np.ndarray) -> float:
    """``cross_entropy(input=M^T, target=I)`` -- model.py:45.

    The class axis of the input is axis 1 of ``M^T`` i.e. the *column* index c of
    ``M[b,d,c]``; the probabilities are used as logits (no log).  Mean over b,d.
    """
    mx = M.max(-1, keepdims=True)
    lse = np.log(np.exp(M - mx).sum(-1)) + mx[..., 0]
    diag = np.einsum("bdd->bd", M)
    return float((lse - diag).mean())


def walk_loss_reference_order(A: np.ndarray) -> float:
    """Literal restatement of the loop at model.py:31-46 (returns ``loss/N``)."""
    B, Tm1, N, _ = A.shape
    T = Tm1 + 1
    loss = 0.0
    eye = np.broadcast_to(np.eye(N, dtype=A.dtype), (B, N, N))
    for k in range(1, T - 1):
        # model.py:41 -- first k forward affinities then the last k of the
        # flipped+transposed half, i.e. A_{k-1}^T ... A_0^T.
        pal = [A[:, t] for t in range(k)] + [A[:, k - 1 - t].transpose(0, 2, 1) for t in range(k)]
        M = eye.copy()
        for t in range(1, 2 * k):  # index 0 is skipped (model.py:42)
            M = row_softmax(pal[t]) @ M
        loss += _cycle_ce(M)
    return loss / N


def chain_factors(A: np.ndarray):
    """Row-softmaxes used by the chain form.

    Returns (S, Sp): ``S[:,t] = rowsoftmax(A_t)``, ``Sp[:,t] = rowsoftmax(A_t^T)``.
    """
    S = row_softmax(A)
    Sp = row_softmax(A.transpose(0, 1, 3, 2))
    return S, Sp


def walk_loss_chain(A: np.ndarray, return_state: bool = False):
    """O(T) form: ``L_k = L_{k-1} S'_{k-1}``, ``R_k = S_{k-1} R_{k-1}``, ``M_k = L_k R_k``."""
    B, Tm1, N, _ = A.shape
    T = Tm1 + 1
    S, Sp = chain_factors(A)
    eye = np.broadcast_to(np.eye(N, dtype=A.dtype), (B, N, N)).copy()
    Ls = [eye]          # L_0
    Rs = [None, eye]    # R_1 = I (R_0 unused)
    loss = 0.0
    Ms = [None]
    for k in range(1, T - 1):
        Ls.append(Ls[k - 1] @ Sp[:, k - 1])
        if k > 1:
            Rs.append(S[:, k - 1] @ Rs[k - 1])
        M = Ls[k] @ Rs[k]
        Ms.append(M)
        loss += _cycle_ce(M)
    loss = loss / N
    if return_state:
        return loss, dict(S=S, Sp=Sp, L=Ls, R=Rs, M=Ms)
    return loss


def _softmax_bwd(P: np.ndarray, dP: np.ndarray) -> np.ndarray:
    return P * (dP - (P * dP).sum(-1, keepdims=True))


def walk_backward_chain(x: np.ndarray, tau: float, dloss: float = 1.0):
    """Analytic gradient of ``loss/N`` w.r.t. the un-normalised embeddings ``x [B,T,N,C]``.

    Returns (loss, dA [B,T-1,N,N], demb [B,T,N,C], dx [B,T,N,C]).
    """
    dt = x.dtype.type
    B, T, N, C = x.shape
    nrm = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), dt(EPS_NORM))
    E = x / nrm
    A = affinities(E, tau)
    dA = np.zeros_like(A)
    dE = np.zeros_like(E)
    if T < 3:
        return 0.0, dA, dE, np.zeros_like(x)
    loss, st = walk_loss_chain(A, return_state=True)
    S, Sp, Ls, Rs, Ms = st["S"], st["Sp"], st["L"], st["R"], st["M"]
    eye = np.eye(N, dtype=x.dtype)
    K = T - 2
    scale = dt(dloss) / dt(B * N * N)
    dL = [np.zeros((B, N, N), x.dtype) for _ in range(K + 1)]
    dR = [np.zeros((B, N, N), x.dtype) for _ in range(K + 1)]
    for k in range(1, K + 1):
        dM = (row_softmax(Ms[k]) - eye) * scale
        dL[k] += dM @ Rs[k].transpose(0, 2, 1)
        dR[k] += Ls[k].transpose(0, 2, 1) @ dM
    dS = np.zeros_like(A)
    dSp = np.zeros_like(A)
    for k in range(K, 0, -1):
        dSp[:, k - 1] = Ls[k - 1].transpose(0, 2, 1) @ dL[k]
        dL[k - 1] += dL[k] @ Sp[:, k - 1].transpose(0, 2, 1)
        if k > 1:
            dS[:, k - 1] = dR[k] @ Rs[k - 1].transpose(0, 2, 1)
            dR[k - 1] += S[:, k - 1].transpose(0, 2, 1) @ dR[k]
    dA = _softmax_bwd(S, dS) + _softmax_bwd(Sp, dSp).transpose(0, 1, 3, 2)
    inv_tau = dt(1.0 / tau)
    dE[:, :-1] += np.einsum("btnm,btmc->btnc", dA, E[:, 1:]) * inv_tau
    dE[:, 1:] += np.einsum("btnm,btnc->btmc", dA, E[:, :-1]) * inv_tau
    dx = (dE - E * (E * dE).sum(-1, keepdims=True)) / nrm
    return loss, dA, dE, dx


def crw_forward(x: np.ndarray, tau: float):
    """``CRW.forward`` tail (model.py:22-46): returns (loss/N, A)."""
    E = l2_normalize(x)
    A = affinities(E, tau)
    if x.shape[1] < 3:
        return 0.0, A
    return walk_loss_chain(A), A
