"""``torch.library`` custom ops (namespace ``crw_b200::``) over the C ABI of libcrw_b200.so.

Each op validates on the Python side only what the C side cannot see (device, dtype,
contiguity), passes raw device pointers + the current CUDA stream across the ABI, and
raises ``RuntimeError`` on any non-zero return.  No op has a CPU implementation.
"""
from __future__ import annotations

import collections
from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import PREC_FP32, PREC_BF16X3, PREC_TC_EXACT, LP_REF_EXACT, LP_FIXED  # noqa: F401


PREC_AUTO = -1     # label propagation: the fastest path whose results are the fp32 path's (resolved by lp_precision())


def lp_precision(precision: int, N: int, C: int, k: int) -> int:
    """PREC_AUTO -> PREC_TC_EXACT where the exact tensor path serves the shape (C = 128, 8 <= N <= 128, k <= 24), else PREC_FP32.
    Both give bit-identical results (same pinned fp32 arithmetic decides every id and weight); the choice is speed only."""
    if int(precision) != PREC_AUTO:
        return int(precision)
    return PREC_TC_EXACT if (C == 128 and 8 <= N <= 128 and k <= 24) else PREC_FP32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Tensor, name: str, dtype=torch.float32) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"crw_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"crw_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _p(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# l2_normalize  (model.py:22, utils.py:115)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::l2_normalize", mutates_args=())
def l2_normalize(x: Tensor) -> Tensor:
    x = _chk(x, "x")
    out = torch.empty_like(x)
    C = x.shape[-1]
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().crw_l2_normalize(_p(x), x.numel() // max(C, 1), C, _p(out), _stream()), "crw_l2_normalize")
    return out


@l2_normalize.register_fake
def _(x):
    return torch.empty_like(x)


# ------------------------------------------------------------------------------------------------
# walk_loss / walk_loss_backward  (model.py:22-46 and its autograd)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::walk_loss", mutates_args=())
def walk_loss(x: Tensor, tau: float, need_A: bool, precision: int) -> Tuple[Tensor, Tensor, Tensor]:
    """x [B,T,N,C] raw encoder output -> (loss/N scalar, A [B,T-1,N,N] or empty, saved workspace)."""
    x = _chk(x, "x")
    B, T, N, C = x.shape
    L = _lib.lib()
    loss = torch.empty((), device=x.device, dtype=torch.float32)
    A = torch.empty((B, T - 1, N, N) if need_A else (0,), device=x.device, dtype=torch.float32)
    nbytes = L.crw_walk_saved_bytes(B, T, N, C, int(precision))
    saved = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        _lib.check(L.crw_walk_forward(_p(x), B, T, N, C, float(tau), int(precision), _p(loss), _p(A), _p(saved),
                                      nbytes, _stream()), "crw_walk_forward")
    return loss, A, saved


@walk_loss.register_fake
def _(x, tau, need_A, precision):
    B, T, N, C = x.shape
    nbytes = _lib.lib().crw_walk_saved_bytes(B, T, N, C, int(precision))     # host-only arithmetic: safe under FakeTensor
    return (x.new_empty(()), x.new_empty((B, T - 1, N, N) if need_A else (0,)),
            torch.empty(nbytes, device=x.device, dtype=torch.uint8))


@torch.library.custom_op("crw_b200::walk_loss_backward", mutates_args=())
def walk_loss_backward(x: Tensor, saved: Tensor, dloss: Tensor, dA: Tensor, tau: float, precision: int) -> Tensor:
    x = _chk(x, "x")
    B, T, N, C = x.shape
    L = _lib.lib()
    dloss = _chk(dloss.reshape(1), "dloss")
    dA_c = _chk(dA, "dA") if dA.numel() else None
    dx = torch.empty_like(x)
    sbytes = L.crw_walk_backward_scratch_bytes(B, T, N, C, int(precision))
    scratch = torch.empty(sbytes, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        _lib.check(L.crw_walk_backward(_p(x), _p(saved), saved.numel(), _p(dloss), _p(dA_c), B, T, N, C, float(tau),
                                       int(precision), _p(dx), _p(scratch), sbytes, _stream()), "crw_walk_backward")
    return dx


@walk_loss_backward.register_fake
def _(x, saved, dloss, dA, tau, precision):
    return torch.empty_like(x)


def _walk_setup(ctx, inputs, output):
    x, tau, need_A, precision = inputs
    _, _, saved = output
    ctx.save_for_backward(x, saved)
    ctx.tau, ctx.need_A, ctx.precision = tau, need_A, precision
    # no zero-filled stand-ins for the gradients of `A` (when unused) and of the `saved` workspace: at scaled
    # geometries those are GB-sized fills (measured: 2 x 0.48 ms per step at N=369)
    ctx.set_materialize_grads(False)


def _walk_bwd(ctx, g_loss, g_A, _g_saved):
    x, saved = ctx.saved_tensors
    if g_loss is None:
        g_loss = torch.zeros((), device=x.device, dtype=torch.float32)
    if g_A is None or not ctx.need_A:
        g_A = torch.empty(0, device=x.device, dtype=torch.float32)
    dx = walk_loss_backward(x, saved, g_loss.to(torch.float32), g_A, ctx.tau, ctx.precision)
    return dx, None, None, None


walk_loss.register_autograd(_walk_bwd, setup_context=_walk_setup)


# ------------------------------------------------------------------------------------------------
# affinity_topk  (maskedatt.py:151-175)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::affinity_topk", mutates_args=())
def affinity_topk(keys: Tensor, queries: Tensor, n_first: int, ctx: int, radius: float, temp: float, k: int,
                  precision: int) -> Tuple[Tensor, Tensor]:
    """keys [n_keys,N,C], queries [Q,N,C] (normalised) -> W [Q,k,N] f32, I [Q,k,N] i32."""
    keys, queries = _chk(keys, "keys"), _chk(queries, "queries")
    Q, N, C = queries.shape
    if keys.shape[0] < n_first + Q - 1:
        raise RuntimeError("crw_b200::affinity_topk: query n needs keys 0..n-1")
    W = torch.empty((Q, k, N), device=keys.device, dtype=torch.float32)
    I = torch.empty((Q, k, N), device=keys.device, dtype=torch.int32)
    with torch.cuda.device(keys.device):
        _lib.check(_lib.lib().crw_affinity_topk(_p(keys), _p(queries), n_first, Q, N, C, ctx, float(radius), float(temp),
                                                k, int(precision), _p(W), _p(I), _stream()), "crw_affinity_topk")
    return W, I


@affinity_topk.register_fake
def _(keys, queries, n_first, ctx, radius, temp, k, precision):
    Q, N, _ = queries.shape
    return keys.new_empty((Q, k, N)), torch.empty((Q, k, N), device=keys.device, dtype=torch.int32)


# ------------------------------------------------------------------------------------------------
# label_gather_step  (labelprop.py:106-116)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::label_gather_step", mutates_args=())
def label_gather_step(W: Tensor, I: Tensor, lbl: Tensor) -> Tensor:
    """W,I [k,N]; lbl [F,M,N] -> soft mask [M,N]."""
    W, I, lbl = _chk(W, "W"), _chk(I, "I", torch.int32), _chk(lbl, "lbl")
    k, N = W.shape
    F, M, _ = lbl.shape
    out = torch.empty((M, N), device=W.device, dtype=torch.float32)
    with torch.cuda.device(W.device):
        _lib.check(_lib.lib().crw_label_gather_step(_p(W), _p(I), _p(lbl), F, N, M, k, _p(out), None, _stream()),
                   "crw_label_gather_step")
    return out


@label_gather_step.register_fake
def _(W, I, lbl):
    return W.new_empty((lbl.shape[1], W.shape[1]))


# ------------------------------------------------------------------------------------------------
# labelprop  (utils.py:115,134-161 + labelprop.py:67-116 + maskedatt.py:151-175)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::labelprop", mutates_args=())
def labelprop(feats: Tensor, mask0: Tensor, ctx: int, radius: float, temp: float, k: int, mode: int, precision: int,
              normalize: bool, return_topk: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """feats [R,T,N,C], mask0 [R,M,N] -> labels [R,T,N] i32, masks [R,T,M,N], W/I [R,T,k,N] (or empty)."""
    feats, mask0 = _chk(feats, "feats"), _chk(mask0, "mask0")
    R, T, N, C = feats.shape
    M = mask0.shape[1]
    precision = lp_precision(precision, N, C, k)
    L = _lib.lib()
    dev = feats.device
    labels = torch.empty((R, T, N), device=dev, dtype=torch.int32)
    masks = torch.empty((R, T, M, N), device=dev, dtype=torch.float32)
    if return_topk:
        W = torch.zeros((R, T, k, N), device=dev, dtype=torch.float32)
        I = torch.zeros((R, T, k, N), device=dev, dtype=torch.int32)
    else:
        W = torch.empty(0, device=dev, dtype=torch.float32)
        I = torch.empty(0, device=dev, dtype=torch.int32)
    sbytes = L.crw_labelprop_scratch_bytes(R, T, N, C, k, int(precision), int(normalize), int(return_topk))
    scratch = torch.empty(sbytes, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.crw_labelprop_forward(_p(feats), _p(mask0), R, T, N, C, M, ctx, float(radius), float(temp), k,
                                           int(mode), int(precision), int(normalize), _p(labels), _p(masks), _p(W),
                                           _p(I), _p(scratch), sbytes, _stream()), "crw_labelprop_forward")
    return labels, masks, W, I


@labelprop.register_fake
def _(feats, mask0, ctx, radius, temp, k, mode, precision, normalize, return_topk):
    R, T, N, _ = feats.shape
    M = mask0.shape[1]
    dev = feats.device
    tk = (R, T, k, N) if return_topk else (0,)
    return (torch.empty((R, T, N), device=dev, dtype=torch.int32), feats.new_empty((R, T, M, N)),
            feats.new_empty(tk), torch.empty(tk, device=dev, dtype=torch.int32))


# labelprop_host returns while its H2D copies are still queued on the library's copy stream, which PyTorch's pinned-memory
# allocator knows nothing about: a temporary such as ``x.pin_memory()`` would go back to the pool (and could be handed out and
# overwritten) while the DMA is still reading it.  Every call therefore parks (event, source) here until the event -- recorded
# on the caller's stream, which has waited for every copy -- has completed.
_pinned_in_flight = collections.deque()


def _park_pinned(feats: Tensor) -> None:
    while _pinned_in_flight and _pinned_in_flight[0][0].query():
        _pinned_in_flight.popleft()
    ev = torch.cuda.Event()
    ev.record()
    _pinned_in_flight.append((ev, feats))


@torch.library.custom_op("crw_b200::labelprop_host", mutates_args=())
def labelprop_host(feats: Tensor, mask0: Tensor, ctx: int, radius: float, temp: float, k: int, mode: int,
                   normalize: bool, return_topk: bool, precision: int = PREC_BF16X3) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """``labelprop`` (tensor paths) for features in PINNED HOST memory: the H2D copy is cut into pieces and overlapped with the
    search of the previous piece.  ``precision``: PREC_BF16X3 (the approximate kernel; chunks of query tiles) or PREC_TC_EXACT /
    PREC_AUTO (the exact tensor path; segments of frames -- results bit-identical to ``labelprop`` with PREC_TC_EXACT).
    ``feats`` must not be modified until the current stream has caught up."""
    if feats.is_cuda or not feats.is_pinned():
        raise RuntimeError("crw_b200::labelprop_host: `feats` must be a pinned host tensor (use labelprop for CUDA tensors)")
    if feats.dtype != torch.float32:
        raise RuntimeError("crw_b200::labelprop_host: `feats` must be float32")
    feats = feats.contiguous()
    mask0 = _chk(mask0, "mask0")
    R, T, N, C = feats.shape
    M = mask0.shape[1]
    L = _lib.lib()
    dev = mask0.device
    prec = lp_precision(precision, N, C, k) if int(precision) == PREC_AUTO else int(precision)
    if prec not in (PREC_BF16X3, PREC_TC_EXACT):
        raise RuntimeError("crw_b200::labelprop_host: precision must be PREC_BF16X3 or PREC_TC_EXACT (the fp32 kernel has no host-streamed entry)")
    labels = torch.empty((R, T, N), device=dev, dtype=torch.int32)
    masks = torch.empty((R, T, M, N), device=dev, dtype=torch.float32)
    if return_topk:
        W = torch.zeros((R, T, k, N), device=dev, dtype=torch.float32)
        I = torch.zeros((R, T, k, N), device=dev, dtype=torch.int32)
    else:
        W = torch.empty(0, device=dev, dtype=torch.float32)
        I = torch.empty(0, device=dev, dtype=torch.int32)
    exact = prec == PREC_TC_EXACT
    sbytes = (L.crw_labelprop_host_exact_scratch_bytes if exact else L.crw_labelprop_host_scratch_bytes)(R, T, N, C, k, int(return_topk))
    scratch = torch.empty(sbytes, device=dev, dtype=torch.uint8)
    fn, name = ((L.crw_labelprop_forward_host_exact, "crw_labelprop_forward_host_exact") if exact
                else (L.crw_labelprop_forward_host, "crw_labelprop_forward_host"))
    with torch.cuda.device(dev):
        _lib.check(fn(feats.data_ptr(), _p(mask0), R, T, N, C, M, ctx, float(radius), float(temp), k, int(mode), int(normalize),
                      _p(labels), _p(masks), _p(W), _p(I), _p(scratch), sbytes, _stream()), name)
        _park_pinned(feats)
    return labels, masks, W, I


@labelprop_host.register_fake
def _(feats, mask0, ctx, radius, temp, k, mode, normalize, return_topk, precision=PREC_BF16X3):
    R, T, N, _ = feats.shape
    M = mask0.shape[1]
    dev = mask0.device
    tk = (R, T, k, N) if return_topk else (0,)
    return (torch.empty((R, T, N), device=dev, dtype=torch.int32), torch.empty((R, T, M, N), device=dev),
            torch.empty(tk, device=dev), torch.empty(tk, device=dev, dtype=torch.int32))


# ------------------------------------------------------------------------------------------------
# horizontality_xent  (utils.py:118-123)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::horizontality_xent", mutates_args=())
def horizontality_xent(emb: Tensor) -> Tensor:
    """emb [T,N,C] normalised -> xent [N,T-1]."""
    emb = _chk(emb, "emb")
    T, N, C = emb.shape
    out = torch.empty((N, max(T - 1, 0)), device=emb.device, dtype=torch.float32)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.lib().crw_horizontality_xent(_p(emb), T, N, C, _p(out), _stream()), "crw_horizontality_xent")
    return out


@horizontality_xent.register_fake
def _(emb):
    T, N, _ = emb.shape
    return emb.new_empty((N, max(T - 1, 0)))


# ------------------------------------------------------------------------------------------------
# labels_upsample  (scripts/test/test_all.py:79,96: Resize((seg_h, rg_len), NEAREST) of final_prediction)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::labels_upsample", mutates_args=())
def labels_upsample(labels: Tensor, H: int, W: int) -> Tensor:
    """labels [R,T,N] i32 -> [R,H,W] f32 nearest-neighbour upsample (rows <- nodes, columns <- frames)."""
    labels = _chk(labels, "labels", torch.int32)
    R, T, N = labels.shape
    out = torch.empty((R, H, W), device=labels.device, dtype=torch.float32)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().crw_labels_upsample(_p(labels), R, T, N, H, W, _p(out), _stream()), "crw_labels_upsample")
    return out


@labels_upsample.register_fake
def _(labels, H, W):
    return labels.new_empty((labels.shape[0], H, W), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------
# patch_unfold  (dataset.py:34-47 RGDataset.__getitem__ + the use_last flip of utils.py:108)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::patch_unfold", mutates_args=())
def patch_unfold(rg: Tensor, col_start: int, col_stride: int, R: int, T: int, h: int, w: int, oh: int, ow: int,
                 reverse: bool) -> Tensor:
    """rg [H,W] f32 (device) -> frames [R,T,N,h,w]; item r starts at column col_start + r*col_stride."""
    rg = _chk(rg, "rg")
    H, ld = rg.shape
    N = (H - oh) // (h - oh)                                            # dataset.py:22
    out = torch.empty((R, T, N, h, w), device=rg.device, dtype=torch.float32)
    with torch.cuda.device(rg.device):
        _lib.check(_lib.lib().crw_patch_unfold(_p(rg), H, ld, col_start, col_stride, R, T, N, h, w, oh, ow, int(reverse),
                                               _p(out), _stream()), "crw_patch_unfold")
    return out


@patch_unfold.register_fake
def _(rg, col_start, col_stride, R, T, h, w, oh, ow, reverse):
    return rg.new_empty((R, T, (rg.shape[0] - oh) // (h - oh), h, w))


# ------------------------------------------------------------------------------------------------
# seed_labels  (utils.py:139-147)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::seed_labels", mutates_args=())
def seed_labels(seg: Tensor, rows: int, col_start: int, col_stride: int, R: int, N: int, M: int) -> Tuple[Tensor, Tensor]:
    """seg [H,W] f32 class ids (device) -> (label0 [R,N] i32, mask0 [R,M,N] f32) from rows [0,rows) of one column each."""
    seg = _chk(seg, "seg")
    if rows > seg.shape[0]:
        raise RuntimeError("crw_b200::seed_labels: rows exceeds the segmentation height")
    label0 = torch.empty((R, N), device=seg.device, dtype=torch.int32)
    mask0 = torch.empty((R, M, N), device=seg.device, dtype=torch.float32)
    with torch.cuda.device(seg.device):
        _lib.check(_lib.lib().crw_seed_labels(_p(seg), rows, seg.shape[1], col_start, col_stride, R, N, M, _p(label0),
                                              _p(mask0), _stream()), "crw_seed_labels")
    return label0, mask0


@seed_labels.register_fake
def _(seg, rows, col_start, col_stride, R, N, M):
    return (torch.empty((R, N), device=seg.device, dtype=torch.int32), seg.new_empty((R, M, N)))


# ------------------------------------------------------------------------------------------------
# fuse_reversed  (scripts/test/test_all.py:146-158)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::fuse_reversed", mutates_args=())
def fuse_reversed(fwd: Tensor, rev: Tensor, rg_len: int, rule: int) -> Tensor:
    """fwd, rev [H,W] f32 class maps (rev in the reversed pass's column order) -> fused map [H,W]."""
    fwd, rev = _chk(fwd, "fwd"), _chk(rev, "rev")
    if fwd.shape != rev.shape or fwd.dim() != 2:
        raise RuntimeError("crw_b200::fuse_reversed: fwd and rev must both be [H,W]")
    H, W = fwd.shape
    L = _lib.lib()
    out = torch.empty_like(fwd)
    nb = L.crw_fuse_reversed_scratch_bytes(W)
    scratch = torch.empty(max(nb, 1), device=fwd.device, dtype=torch.uint8)
    with torch.cuda.device(fwd.device):
        _lib.check(L.crw_fuse_reversed(_p(fwd), _p(rev), H, W, rg_len, rule, _p(out), _p(scratch), nb, _stream()),
                   "crw_fuse_reversed")
    return out


@fuse_reversed.register_fake
def _(fwd, rev, rg_len, rule):
    return torch.empty_like(fwd)


# ------------------------------------------------------------------------------------------------
# adam_step  (scripts/train.py:56,72: torch.optim.Adam over flat buffers, one launch)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("crw_b200::adam_step", mutates_args=("params", "exp_avg", "exp_avg_sq"))
def adam_step(params: Tensor, grads: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, lr: float, beta1: float, beta2: float,
              eps: float, weight_decay: float, step: int, grad_scale: float) -> None:
    """In place: one torch.optim.Adam update of the flat f32 buffer ``params`` from ``grad_scale * grads``; ``step`` counts from 1."""
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 1 or t.numel() != params.numel():
            raise RuntimeError(f"crw_b200::adam_step: `{name}` must be a contiguous 1-D f32 CUDA tensor of the parameters' length")
    with torch.cuda.device(params.device):
        _lib.check(_lib.lib().crw_adam_step(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), lr, beta1, beta2, eps,
                                            weight_decay, step, grad_scale, _stream()), "crw_adam_step")


@adam_step.register_fake
def _(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale):
    return None
