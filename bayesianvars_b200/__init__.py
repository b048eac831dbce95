"""bayesianvars_b200 - B200 (sm_100a) implementation of the bayesianVARs Gibbs
coefficient-draw hot path behind the reference's Rcpp entry points.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``libbvb.so``) and ``api`` (the host-side mirror of the reference interface that
tests and bench drive through the same C ABI the Rcpp shim binds).
"""
from ._lib import BvbError, load  # noqa: F401
from .api import Handle  # noqa: F401
from . import bvar  # noqa: F401  (registers the Gibbs entry points with the ctypes loader)

__all__ = ["Handle", "BvbError", "load"]
