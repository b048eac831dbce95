"""GPU parity of sample_PHI_factor through the C ABI (host buffers), against the
oracle on the same seeded inputs, against the committed golden fixtures, and -
at BASELINE.json's full sizes - through size-independent properties.

Tolerance: north_star asks 1e-9 relative per conditional draw in FP64; metric =
max_i |x_i - xhat_i| / max_i |xhat_i| per equation."""
import numpy as np
import pytest

from oracle import regression as orc
from tools.synth import phi_draw_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-9


def per_eq_relerr(a, b):
    return (np.max(np.abs(a - b), axis=0) / np.maximum(np.max(np.abs(b), axis=0), 1e-300)).max()


def run(handle, d, **kw):
    return handle.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"],
                                    z=d["z"], **kw)


@pytest.mark.parametrize("M,p,T,r,prior", [
    (2, 0, 30, 0, "flat"),     # lags = 0: intercept only (Kp = 1)
    (3, 1, 40, 0, "flat"),
    (10, 2, 235, 1, "dl"),     # config 1 shape
    (14, 3, 100, 2, "dl"),     # reference smoke-grid shape (setup-all.R:3-5)
    (7, 4, 20, 1, "dl"),       # T < Kp: precision is prior-dominated
    (5, 13, 150, 1, "dl"),     # Kp = 66: 3 column blocks, last one partial (2 cols)
    (16, 4, 120, 3, "dl"),     # Kp = 65
    (50, 4, 300, 6, "dl"),     # config 2
    (26, 16, 450, 1, "dl"),    # Kp = 417: too tall for the look-ahead kernel (v1, one row tile)
    (26, 20, 560, 1, "dl"),    # Kp = 521: the panel no longer fits - two row tiles per block column
    (12, 70, 900, 1, "dl"),    # Kp = 841: three row tiles, last block column partial
])
def test_matches_oracle(handle, M, p, T, r, prior):
    d = phi_draw_inputs(M, p, T, r, seed=100 + M + p, prior=prior)
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], False, d["z"])
    out = run(handle, d)
    assert per_eq_relerr(out, ref) < TOL


def test_golden_fixtures(handle, request):
    g = np.load(request.config.rootpath / "tests" / "golden" / "phi_factor.npz")
    names = sorted({k.split("/")[0] for k in g.files})
    assert names
    for name in names:
        M, p, T, r = (int(g[f"{name}/{k}"]) for k in "MpTr")
        d = phi_draw_inputs(M, p, T, r, seed=int(g[f"{name}/seed"]), prior=str(g[f"{name}/prior"]))
        out = run(handle, d)
        assert per_eq_relerr(out, g[f"{name}/PHI"]) < TOL, name


def test_extreme_prior_dynamic_range(handle):
    """DL/R2D2 drive 1/V_prior over ~1e-3..1e36."""
    d = phi_draw_inputs(12, 3, 80, 1, seed=9)
    rng = np.random.default_rng(9)
    d["V"] = np.asfortranarray(10.0 ** rng.uniform(-36, 3, size=d["V"].shape))
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], False, d["z"])
    out = run(handle, d)
    assert per_eq_relerr(out, ref) < TOL


def test_nonzero_prior_mean_and_device_rng(handle):
    d = phi_draw_inputs(9, 2, 60, 2, seed=21)
    d["PHI0"] = np.asfortranarray(np.random.default_rng(3).standard_normal(d["PHI0"].shape))
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], False, d["z"])
    assert per_eq_relerr(run(handle, d), ref) < TOL
    # z = NULL -> Philox normals on device: two calls differ, both finite
    a = handle.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"])
    b = handle.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"])
    assert np.isfinite(a).all() and np.isfinite(b).all() and not np.allclose(a, b)


def test_full_size_normal_equations(handle):
    """Config 3 (M=100, p=4, T=500): with z = 0 the draw is the posterior mean and
    must satisfy X'W(y - X phi) = (phi - phi0)/v for every equation; linearity
    in z: draw(z) - draw(0) = L^-T z has covariance Omega^-1 => check
    (draw(z)-draw(0))' Omega (draw(z)-draw(0)) = z'z."""
    d = phi_draw_inputs(100, 4, 500, 1, seed=123)
    z = d["z"].copy()
    d["z"] = np.zeros_like(z, order="F")
    mean = run(handle, d)
    d["z"] = z
    draw = run(handle, d)
    Yh = d["Y"] - (d["facload"] @ d["fac"]).T
    X = d["X"]
    for j in range(0, 100, 7):
        w = np.exp(-d["logvar"][:, j])
        g = X.T @ (w * (Yh[:, j] - X @ mean[:, j])) - (mean[:, j] - d["PHI0"][:, j]) / d["V"][:, j]
        scale = np.abs(X.T @ (w * Yh[:, j])).max() + np.abs(mean[:, j] / d["V"][:, j]).max()
        assert np.abs(g).max() / scale < 1e-9
        dl = draw[:, j] - mean[:, j]
        q = dl @ ((X * w[:, None]).T @ (X @ dl)) + np.sum(dl * dl / d["V"][:, j])
        assert abs(q - z[:, j] @ z[:, j]) / (z[:, j] @ z[:, j]) < 1e-8


def test_cholesky_failure_is_reported(handle):
    from bayesianvars_b200 import BvbError
    d = phi_draw_inputs(4, 1, 30, 0, seed=5, prior="flat")
    d["V"][1, 2] = -1e-30   # huge negative precision on one diagonal -> pivot <= 0
    with pytest.raises(BvbError) as ei:
        run(handle, d)
    assert ei.value.code == 3 and ei.value.equation == 3 and "eqation" in str(ei.value)


# ------------------------------------------------------------------ expert_huge branch
@pytest.mark.parametrize("M,p,T,r", [(3, 2, 5, 0), (8, 6, 30, 1), (12, 10, 60, 2), (20, 8, 100, 1), (30, 14, 150, 1),
                                     (20, 32, 530, 1)])   # T x T = 530: row-tiled factorisation
def test_huge_branch_matches_oracle(handle, M, p, T, r):
    """sample_coefficients.cpp:40-52 (Bhattacharya et al.): Kp normals for u, then T normals for
    delta, per equation.  Kp = p*M + 1 > T in most of these shapes."""
    d = phi_draw_inputs(M, p, T, r, seed=300 + M)
    rng = np.random.default_rng(M)
    delta = np.asfortranarray(rng.standard_normal((T, M)))
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], True, d["z"], delta)
    out = handle.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"],
                                   huge=True, z=d["z"], delta=delta)
    # 1e-9 as everywhere, except where the T x T system X_norm V X_norm' + I is itself too ill-conditioned for two
    # correct factorisations (LAPACK's dgesv there, blocked Cholesky here) to agree that closely: then 10 cond eps.
    # (T = 530 with the heavy-tailed prior variances: cond = 3.4e6, measured difference 1.8e-9.)
    Xn = d["X"] * np.exp(-0.5 * d["logvar"][:, :1])
    cond = max(np.linalg.cond((Xn * d["V"][:, j]) @ Xn.T + np.eye(T)) for j in range(min(M, 3)))
    assert per_eq_relerr(out, ref) < max(TOL, 10 * cond * np.finfo(float).eps), cond


def test_huge_and_standard_agree_in_mean(handle):
    """With all injected normals zero both branches return the exact posterior mean (Woodbury)."""
    d = phi_draw_inputs(10, 4, 70, 1, seed=77)
    z0 = np.zeros_like(d["z"], order="F")
    a = handle.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], z=z0)
    b = handle.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], huge=True,
                                 z=z0, delta=np.zeros((70, 10), order="F"))
    assert per_eq_relerr(b, a) < 1e-8


@pytest.mark.parametrize("M,p,T", [(13, 31, 450), (25, 16, 430), (6, 9, 120)])
def test_cluster_form_matches_oracle(handle, M, p, T):
    """Few equations of a tall system (the shard a rank owns under 4-8 GPU equation sharding): K2 runs one equation
    per thread-block cluster (8 or 4 CTAs).  (6, 9, 120) is too short for clusters and checks the dispatch."""
    d = phi_draw_inputs(M, p, T, 1, seed=500 + M)
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], False, d["z"])
    assert per_eq_relerr(run(handle, d), ref) < TOL
