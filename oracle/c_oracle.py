"""ctypes loader for ``oracle/libcrw_oracle.so`` (crw_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcrw_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "crw_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcrw_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.crw_oracle_expf.restype = ctypes.c_float
        L.crw_oracle_expf.argtypes = [ctypes.c_float]
        L.crw_oracle_num_threads.restype = ctypes.c_int
        L.crw_oracle_l2_normalize.restype = None
        L.crw_oracle_l2_normalize.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p]
        L.crw_oracle_lp_topk.restype = ctypes.c_int
        L.crw_oracle_lp_topk.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, _f32p, _i32p]
        L.crw_oracle_lp_gather.restype = ctypes.c_int
        L.crw_oracle_lp_gather.argtypes = [_f32p, _i32p, _f32p, _i32p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _i32p]
        L.crw_oracle_labelprop.restype = ctypes.c_int
        L.crw_oracle_labelprop.argtypes = [_f32p, _f32p, _i32p] + [ctypes.c_int] * 6 + \
            [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, _f32p, _f32p, _i32p]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def num_threads() -> int:
    return int(lib().crw_oracle_num_threads())


def expf(x: np.ndarray) -> np.ndarray:
    L = lib()
    x = np.asarray(x, np.float32)
    return np.array([L.crw_oracle_expf(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def l2_normalize(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().crw_oracle_l2_normalize(_fp(x), x.size // x.shape[-1], x.shape[-1], _fp(out))
    return out


def labelprop(feats: np.ndarray, label0: np.ndarray, M: int, ctx: int, radius: float, temp: float, k: int,
              mode: str = "ref_exact", normalize: bool = True, want_masks: bool = True, want_topk: bool = True):
    """feats [R,T,N,C] f32, label0 [R,N] int -> dict(labels [R,T,N] i32, masks [R,T,M,N], W/I [R,T,k,N])."""
    feats = np.ascontiguousarray(feats, np.float32)
    R, T, N, C = feats.shape
    label0 = np.ascontiguousarray(label0, np.int32).reshape(R, N)
    mask0 = (label0[:, None, :] == np.arange(M, dtype=np.int32)[None, :, None]).astype(np.float32)
    mask0 = np.ascontiguousarray(mask0)
    labels = np.zeros((R, T, N), np.int32)
    masks = np.zeros((R, T, M, N), np.float32) if want_masks else None
    W = np.zeros((R, T, k, N), np.float32) if want_topk else None
    I = np.zeros((R, T, k, N), np.int32) if want_topk else None
    rc = lib().crw_oracle_labelprop(_fp(feats), _fp(mask0), _ip(label0), R, T, N, C, M, ctx, float(radius),
                                    float(temp), k, 1 if mode == "fixed" else 0, 1 if normalize else 0,
                                    _ip(labels), _fp(masks) if want_masks else None,
                                    _fp(W) if want_topk else None, _ip(I) if want_topk else None)
    if rc != 0:
        raise RuntimeError(f"crw_oracle_labelprop failed rc={rc}")
    return dict(labels=labels, masks=masks, W=W, I=I)
