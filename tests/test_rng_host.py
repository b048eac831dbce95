"""Distributional checks of the Philox / normal / gamma / GIG sampler LOGIC, run
on the host through a tiny harness around the same __host__ __device__ code the
kernels use (no GPU needed).  GIGrvg parity is unpinned: the
target is the GIG law itself, via scipy.stats.geninvgauss:
GIG(lambda, chi, psi) = sqrt(chi/psi) * geninvgauss(p=lambda, b=sqrt(chi*psi))."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
from scipy import stats

ROOT = Path(__file__).resolve().parent.parent
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hostrng():
    src = ROOT / "tests" / "hostrng" / "hostrng.cu"
    out = ROOT / "tests" / "hostrng" / "libhostrng.so"
    deps = [src, ROOT / "bayesianvars_b200/csrc/gig.cuh", ROOT / "bayesianvars_b200/csrc/common.cuh"]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                        "--expt-relaxed-constexpr", "-o", str(out), str(src)], check=True)
    return C.CDLL(str(out))


def gig_cdf_numeric(lam, chi, psi, npts=400001):
    """CDF of GIG(lam, chi, psi) by trapezoid integration in u = log x (robust for
    the extreme parameters where scipy's quad-based geninvgauss.cdf loses accuracy)."""
    # mode region in log space and generous tails
    lo = np.log(min(chi / max(2 * abs(lam) + 2, 1) / 50.0, 1e-3 * np.sqrt(chi / psi))) - 60
    hi = np.log(max((2 * abs(lam) + 2) / psi * 50.0, 1e3 * np.sqrt(chi / psi))) + 5
    u = np.linspace(lo, hi, npts)
    logf = lam * u - 0.5 * (chi * np.exp(-u) + psi * np.exp(u))   # includes the Jacobian x
    logf -= logf.max()
    f = np.exp(logf)
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (f[1:] + f[:-1]) * np.diff(u))])
    cdf /= cdf[-1]
    return lambda x: np.interp(np.log(x), u, cdf)


def gig(lib, n, lam, chi, psi, seed=1):
    out = np.empty(n)
    lib.hostrng_gig(n, C.c_double(lam), C.c_double(chi), C.c_double(psi), C.c_ulonglong(seed), out.ctypes.data_as(_dp))
    return out


def test_uniform_and_normal(hostrng):
    n = 200000
    u = np.empty(n); hostrng.hostrng_uniform(n, C.c_ulonglong(3), u.ctypes.data_as(_dp))
    assert 0 < u.min() and u.max() < 1
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    z = np.empty(n); hostrng.hostrng_normal(n, C.c_ulonglong(4), z.ctypes.data_as(_dp))
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 0.01


@pytest.mark.parametrize("shape", [0.0025, 0.3, 0.5, 1.0, 2.5, 50.0, 20000.5])
def test_gamma(hostrng, shape):
    n = 100000
    g = np.empty(n); hostrng.hostrng_gamma(n, C.c_double(shape), C.c_ulonglong(5), g.ctypes.data_as(_dp))
    assert np.isfinite(g).all() and (g >= 0).all()
    if shape >= 0.3:
        assert stats.kstest(g, "gamma", args=(shape,)).pvalue > 1e-3
    else:  # most mass underflows towards 0: compare on the log scale via the cdf at a few points
        for qv in [1e-200, 1e-50, 1e-5, 0.1]:
            assert abs((g <= qv).mean() - stats.gamma.cdf(qv, shape)) < 0.01


# (lambda, chi, psi): every branch of the sampler and the regimes the priors hit
GIG_CASES = [
    (0.5, 1.0, 1.0),           # |lambda| = 1/2: closed-form (reciprocal) inverse Gaussian
    (-0.5, 1.0, 3.0),          # R2D2 psi draw
    (0.5, 1e-12, 1.0),         # ... across the dynamic range the DL psi draw sees
    (0.5, 1e6, 1e-3),
    (-0.5, 1e-8, 2.0),
    (0.4999, 1.0, 1.0),        # next to 1/2: ROU no shift
    (0.6, 1e-6, 1.0),          # small omega, generic lambda: new approach, all three parts
    (0.3, 2e-3, 5e-2),
    (3.0, 2.0, 5.0),           # lambda > 2: ROU with shift
    (0.7, 40.0, 2.0),          # omega > 3: ROU with shift
    (0.9975, 2e-3, 1.0),       # DL lambda draw reflected (|lambda| ~ 1, small omega): new approach
    (-0.9975, 2e-4, 1.0),      # DL lambda draw as called: GIG(a-1, 2|phi|, 2 xi)
    (0.5, 1e-6, 1.0),          # DL psi draw, tiny chi
    (0.0, 0.01, 0.02),         # lambda == 0 branch of the new approach
    (0.1, 1e-4, 1e-3),
    (-0.4, 0.5, 0.5),          # R2D2 lambda draw  GIG(a - 1/2, ...) with a = 0.1
    (-20.0, 3.0, 0.7),         # HMP-like large |lambda|
    (1.5, 0.3, 0.2),
]


@pytest.mark.parametrize("lam,chi,psi", GIG_CASES)
def test_gig_matches_scipy(hostrng, lam, chi, psi):
    n = 60000
    x = gig(hostrng, n, lam, chi, psi, seed=11)
    assert np.isfinite(x).all() and (x > 0).all()
    assert stats.kstest(x, gig_cdf_numeric(lam, chi, psi)).pvalue > 1e-3


def test_numeric_cdf_agrees_with_scipy():
    lam, chi, psi = 0.5, 1.0, 3.0
    dist = stats.geninvgauss(lam, np.sqrt(chi * psi), scale=np.sqrt(chi / psi))
    xs = np.array([0.05, 0.2, 0.5, 1.0, 3.0])
    np.testing.assert_allclose(gig_cdf_numeric(lam, chi, psi)(xs), dist.cdf(xs), atol=1e-6)


def test_gig_limits(hostrng):
    n = 60000
    # chi -> 0, lambda > 0: Gamma(lambda, scale 2/psi)
    x = gig(hostrng, n, 0.5, 1e-17, 1.0)
    assert stats.kstest(x, "gamma", args=(0.5, 0, 2.0)).pvalue > 1e-3
    # psi -> 0, lambda < 0: 1/Gamma(-lambda, scale 2/chi)
    x = gig(hostrng, n, -2.0, 3.0, 1e-17)
    assert stats.kstest(1 / x, "gamma", args=(2.0, 0, 2.0 / 3.0)).pvalue > 1e-3
    # invalid parameters -> NaN (the reference raises)
    assert np.isnan(gig(hostrng, 1, 1.0, -1.0, 1.0)).all()
    assert np.isnan(gig(hostrng, 1, -1.0, 0.0, 1.0)).all()


def test_gig_extreme_dl_regime_stays_finite(hostrng):
    """DL psi draw chi = phi^2/lambda^2 spans hundreds of orders of magnitude."""
    for chi in [1e-300, 1e-100, 1e-30, 1e-14, 1e5, 1e30, 1e200]:
        x = gig(hostrng, 2000, 0.5, chi, 1.0, seed=13)
        assert np.isfinite(x).all() and (x > 0).all(), chi
    for chi in [1e-300, 2e-18, 1e-12, 1e3]:
        x = gig(hostrng, 2000, -0.9975, chi, 1.0, seed=14)
        assert np.isfinite(x).all() and (x > 0).all(), chi
