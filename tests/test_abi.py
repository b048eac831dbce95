"""The C-ABI library loads and exports every symbol include/bvb.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "bvb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bvb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ["bvb_create", "bvb_destroy", "bvb_last_error", "bvb_phi_draw_factor",
                 "bvb_phi_draw_cholesky", "bvb_gig"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(str(lib_built))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/bvb.h but not exported"


def test_python_binding_covers_every_declared_symbol():
    from bayesianvars_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_device(lib_built):
    """Without a CUDA device handle creation must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    from bayesianvars_b200 import BvbError, Handle
    try:
        Handle(0)
    except BvbError as e:
        assert e.code == 1
    else:
        raise AssertionError("Handle() succeeded without a GPU")


def test_multi_handle_fails_loudly_without_devices(lib_built):
    """bvb_create_multi has no CPU fallback either: without CUDA devices it must return an error, never a handle."""
    import torch
    if torch.cuda.is_available():
        return
    from bayesianvars_b200 import BvbError
    from bayesianvars_b200 import bvar as B
    try:
        B.MultiHandle([0, 1])
    except BvbError as e:
        assert e.code != 0
    else:
        raise AssertionError("MultiHandle() succeeded without GPUs")
