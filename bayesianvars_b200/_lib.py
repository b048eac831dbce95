"""ctypes binding of libbvb.so - the same C ABI (include/bvb.h) the Rcpp shim binds.

There is deliberately no fallback: if the shared library is missing or no CUDA
device can be opened, importing/handle creation raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(__import__("os").environ.get("BVB_LIB", PKG / "libbvb.so"))   # BVB_LIB: A/B builds of the same ABI

BVB_OK = 0
ERR_NAMES = {1: "BVB_ERR_CUDA", 2: "BVB_ERR_ARG", 3: "BVB_ERR_NUMERIC", 4: "BVB_ERR_UNSUPPORTED", 5: "BVB_ERR_COMM"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/bvb.h declares.
SIGNATURES = {
    "bvb_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_uint64]),
    "bvb_destroy": (None, [_vp]),
    "bvb_version": (C.c_char_p, []),
    "bvb_set_seed": (C.c_int, [_vp, C.c_uint64]),
    "bvb_last_error": (C.c_int, [_vp, _ip, _ip, _ip, C.c_char_p, C.c_int]),
    "bvb_phi_draw_factor": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp]),
    "bvb_phi_draw_cholesky": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "bvb_sample_u": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "bvb_gig": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_uint64, _dp]),
    "bvb_irf": (C.c_int, [_vp] + [C.c_int] * 9 + [_dp] * 7),
    "bvb_predict": (C.c_int, [_vp] + [C.c_int] * 7 + [_dp] * 5 + [_ip, C.c_int, _dp, _dp, _dp, _ip, C.c_int, C.c_int, _dp,
                                              C.c_int, _ip, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "bvb_predh": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "bvb_vcov": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]),
    "bvb_insample": (C.c_int, [_vp] + [C.c_int] * 7 + [_dp] * 7),
    "bvb_sv_latent_draw": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.POINTER(C.c_ubyte)] + [C.c_double] * 4 + [_dp]),
    "bvb_measure_fp64_peak": (C.c_int, [_vp, _dp]),
    "bvb_last_timing": (C.c_int, [_vp, _dp]),
}

_lib = None


class BvbError(RuntimeError):
    def __init__(self, code, msg, step=0, sweep=0, equation=0):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code, self.step, self.sweep, self.equation = code, step, sweep, equation


def load():
    """Load libbvb.so (must have been built: python -m bayesianvars_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} not built - run `python -m bayesianvars_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def fptr(a):
    """Column-major float64 view -> double* (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.ndim == 1)
    return a.ctypes.data_as(_dp)


def fmat(a):
    """Copy-if-needed to a float64 Fortran-ordered array (what R / Armadillo hold)."""
    if a is None:
        return None
    return np.asfortranarray(a, dtype=np.float64)
