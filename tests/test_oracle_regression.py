"""CPU checks that pin the linear-algebra oracle (no GPU):
algebraic identities + the reference's own statistical acceptance tests
(tests/testthat/test-bvar.R) re-expressed against the restatement."""
import numpy as np
import pytest

from oracle import regression as orc
from tests.stat_helpers import geweke_phi_cholesky
from tools.synth import phi_draw_inputs


def test_standard_branch_solves_normal_equations():
    rng = np.random.default_rng(1)
    T, K = 60, 7
    X = rng.standard_normal((T, K)); y = rng.standard_normal(T)
    v = np.exp(rng.normal(0, 2, K)); m0 = rng.standard_normal(K)
    mean = orc.univariate_regression_update(m0, v, y, X, False, False, np.zeros(K))
    lhs = (X.T @ X + np.diag(1 / v)) @ mean
    rhs = X.T @ y + m0 / v
    np.testing.assert_allclose(lhs, rhs, rtol=1e-10, atol=1e-10)


def test_draw_has_posterior_covariance():
    rng = np.random.default_rng(2)
    T, K = 30, 3
    X = rng.standard_normal((T, K)); y = rng.standard_normal(T)
    v = np.array([0.5, 2.0, 10.0]); m0 = np.zeros(K)
    Z = rng.standard_normal((20000, K))
    draws = np.array([orc.univariate_regression_update(m0, v, y, X, False, False, z) for z in Z])
    cov = np.linalg.inv(X.T @ X + np.diag(1 / v))
    np.testing.assert_allclose(np.cov(draws.T), cov, atol=4 * cov.max() / np.sqrt(20000) * 3)


def test_huge_branch_same_mean_as_standard():
    """Bhattacharya et al. sampler (sample_coefficients.cpp:40-52): with all
    injected normals zero it returns the exact posterior mean (Woodbury)."""
    rng = np.random.default_rng(3)
    T, K = 25, 40
    X = rng.standard_normal((T, K)); y = rng.standard_normal(T)
    v = np.exp(rng.normal(0, 1, K)); m0 = rng.standard_normal(K)
    a = orc.univariate_regression_update(m0, v, y, X, False, False, np.zeros(K))
    b = orc.univariate_regression_update(m0, v, y, X, False, True, np.zeros(K), np.zeros(T))
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-11)


def test_huge_branch_covariance():
    rng = np.random.default_rng(4)
    T, K = 6, 4
    X = rng.standard_normal((T, K)); y = rng.standard_normal(T)
    v = np.array([0.5, 2.0, 1.0, 3.0]); m0 = np.zeros(K)
    n = 20000
    draws = np.array([orc.univariate_regression_update(m0, v, y, X, False, True,
                                                       rng.standard_normal(K), rng.standard_normal(T))
                      for _ in range(n)])
    cov = np.linalg.inv(X.T @ X + np.diag(1 / v))
    np.testing.assert_allclose(np.cov(draws.T), cov, atol=0.03)


def test_factor_draw_matches_direct_formula():
    d = phi_draw_inputs(6, 2, 50, r=2, seed=5)
    out = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], False, d["z"])
    Yh = d["Y"] - (d["facload"] @ d["fac"]).T
    for j in range(6):
        w = np.exp(-d["logvar"][:, j])
        Om = (d["X"] * w[:, None]).T @ d["X"] + np.diag(1 / d["V"][:, j])
        L = np.linalg.cholesky(Om)
        b = d["X"].T @ (w * Yh[:, j])
        ref = np.linalg.solve(L.T, np.linalg.solve(L, b) + d["z"][:, j])
        np.testing.assert_allclose(out[:, j], ref, rtol=1e-9, atol=1e-12)


def test_kronecker_design_collapses_to_weighted_syrk():
    """the T(M-j) x Kp Kronecker design of sample_PHI equals a
    weighted SYRK on X - the identity the CUDA path relies on."""
    rng = np.random.default_rng(6)
    T, M, K = 40, 5, 7
    X = rng.standard_normal((T, K)); Y = rng.standard_normal((T, M))
    U = np.triu(rng.standard_normal((M, M)) * 0.3, 1) + np.eye(M)
    d_sqrt = np.exp(rng.normal(0, 0.5, (T, M)))
    V = np.exp(rng.normal(0, 2, (K, M))); PHI0 = rng.standard_normal((K, M)) * 0.1
    PHI = rng.standard_normal((K, M)); z = rng.standard_normal((K, M))
    a = orc.sample_PHI(PHI, PHI0, Y, X, U, d_sqrt, V, z)
    b = orc.sample_PHI_collapsed(PHI, PHI0, Y, X, U, d_sqrt, V, z)
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def test_sample_U_structure_and_values():
    rng = np.random.default_rng(7)
    T, M = 50, 4
    E = rng.standard_normal((T, M)); d_sqrt = np.exp(rng.normal(0, 0.3, (T, M)))
    V = np.exp(rng.normal(0, 1, M * (M - 1) // 2)); z = rng.standard_normal(M * (M - 1) // 2)
    U = orc.sample_U(np.eye(M), E, V, d_sqrt, z)
    assert np.allclose(np.diag(U), 1) and np.allclose(np.tril(U, -1), 0)
    i, ind = 2, 1
    Zm = -E[:, :i] / d_sqrt[:, [i]]; c = E[:, i] / d_sqrt[:, i]
    Om = Zm.T @ Zm + np.diag(1 / V[ind:ind + i])
    L = np.linalg.cholesky(Om)
    ref = np.linalg.solve(L.T, np.linalg.solve(L, Zm.T @ c) + z[ind:ind + i])
    np.testing.assert_allclose(U[:i, i], ref, rtol=1e-10)


def test_geweke_sample_phi_cholesky_oracle():
    """tests/testthat/test-bvar.R:21-92 against the restated sample_PHI."""
    rng = np.random.default_rng(123)
    p1, p2 = geweke_phi_cholesky(orc.sample_PHI, 3000, rng)
    assert p1 > 0.01 and p2 > 0.01
