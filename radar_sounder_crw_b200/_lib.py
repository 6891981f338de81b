"""ctypes binding of lib/libcrw_b200.so (C ABI: include/crw_b200.h).  Fails loudly when missing."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcrw_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK = 0
PREC_FP32, PREC_BF16X3, PREC_TC_EXACT = 0, 1, 3
LP_REF_EXACT, LP_FIXED = 0, 1

_c_int, _c_f, _c_sz, _vp = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
_c_i64 = ctypes.c_int64

# name -> (restype, argtypes); mirrors include/crw_b200.h one to one
SIGNATURES = {
    "crw_version": (_c_int, []),
    "crw_built_arch": (_c_int, []),
    "crw_error_string": (ctypes.c_char_p, [_c_int]),
    "crw_l2_normalize": (_c_int, [_vp, ctypes.c_int64, _c_int, _vp, _vp]),
    "crw_walk_saved_bytes": (_c_sz, [_c_int] * 5),
    "crw_walk_forward": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_f, _c_int, _vp, _vp, _vp, _c_sz, _vp]),
    "crw_walk_backward_scratch_bytes": (_c_sz, [_c_int] * 5),
    "crw_walk_backward": (_c_int, [_vp, _vp, _c_sz, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_f, _c_int, _vp, _vp,
                                   _c_sz, _vp]),
    "crw_affinity_topk": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_int, _c_int, _vp,
                                   _vp, _vp]),
    "crw_label_gather": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "crw_label_gather_step": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "crw_labelprop_scratch_bytes": (_c_sz, [_c_int] * 8),
    "crw_labelprop_forward": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_int,
                                       _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_sz, _vp]),
    "crw_labelprop_host_scratch_bytes": (_c_sz, [_c_int] * 6),
    "crw_labelprop_forward_host": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_int,
                                            _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_sz, _vp]),
    "crw_labelprop_host_exact_scratch_bytes": (_c_sz, [_c_int] * 6),
    "crw_labelprop_forward_host_exact": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_int,
                                            _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_sz, _vp]),
    "crw_horizontality_xent": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "crw_labels_upsample": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "crw_patch_unfold": (_c_int, [_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  _c_int, _c_int, _vp, _vp]),
    "crw_adam_step": (_c_int, [_vp, _vp, _vp, _vp, _c_i64] + [ctypes.c_double] * 5 + [_c_i64, ctypes.c_double, _vp]),
    "crw_seed_labels": (_c_int, [_vp, _c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "crw_fuse_reversed_scratch_bytes": (_c_sz, [_c_i64]),
    "crw_fuse_reversed": (_c_int, [_vp, _vp, _c_int, _c_i64, _c_int, _c_int, _vp, _vp, _c_sz, _vp]),
    "crw_debug_umma_gemm": (_c_int, [_vp, _vp, _c_int, _vp, _vp]),
    "crw_debug_umma_ts_gemm": (_c_int, [_vp, _vp, _c_int, _vp, _vp]),
    "crw_debug_umma_tscp_gemm": (_c_int, [_vp, _vp, _c_int, _vp, _vp]),
    "crw_debug_umma_mn_gemm": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "crw_debug_umma_pair_gemm": (_c_int, [_vp, _vp, _c_int, _vp, _vp]),
    "crw_debug_lp_profile": (_c_int, [_vp, _c_int]),
    "crw_debug_lp_x_profile": (_c_int, [_vp, _c_int]),
    "crw_debug_walk_fused_profile": (_c_int, [_vp, _c_int]),
    "crw_debug_lp_schedule": (_c_int, [_c_int, _c_int, _c_int, _vp, _c_int]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into lib/libcrw_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libcrw_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CRW hot path has no fallback. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C radar_sounder_crw_b200/csrc`.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the header and the library drift apart
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def error_string(code: int) -> str:
    return lib().crw_error_string(code).decode()


def check(code: int, what: str) -> None:
    if code != OK:
        raise RuntimeError(f"{what} failed: {error_string(code)} (code {code})")
