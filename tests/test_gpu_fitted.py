"""GPU parity of predh and vcov_cpp (src/utilities_cpp.cpp:181-212, 436-472) through the C ABI against the oracle
restatement (oracle/fitted.py): pure element-wise FP64 arithmetic, tolerance 1e-12 relative."""
import numpy as np
import pytest

from oracle import fitted as of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,M,r,R,factor", [(7, 5, 2, 4, True), (130, 6, 1, 3, True), (9, 4, 0, 2, True), (11, 5, 0, 3, False),
                                            (200, 13, 0, 2, False), (3, 1, 0, 2, False),
                                            (64, 9, 8, 2, True), (33, 7, 9, 2, True), (250, 23, 3, 2, True)])   # (r <= 8: the column-block kernel; r = 9: element-wise)
def test_vcov_matches_oracle(handle, T, M, r, R, factor):
    rng = np.random.default_rng(T + M)
    fl = rng.standard_normal((M, r, R)) if factor else np.zeros((0, 0, 0))
    U = 0.3 * rng.standard_normal((M * (M - 1) // 2, R))
    lv = rng.normal(-1.0, 0.5, (T, M + (r if factor else 0), R))
    ref = of.vcov_cpp(factor, fl, lv, U, M, r if factor else 0)
    got = handle.vcov_cpp(factor, fl, lv, U, M, r if factor else 0)
    assert got.shape == (T * M, M, R)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)
    cube = got.reshape(T, M, M, R, order="F")                     # utility_functions.R:226
    assert np.allclose(cube, np.transpose(cube, (0, 2, 1, 3)), rtol=1e-13)


def test_predh_matches_oracle_with_injected_normals(handle):
    rng = np.random.default_rng(5)
    M, R, each = 7, 9, 3
    lvT = rng.normal(-1, 0.5, (M, R)); mu = rng.normal(-1, 0.2, (M, R))
    phi = rng.uniform(0.7, 0.99, (M, R)); sig = rng.uniform(0.05, 0.4, (M, R))
    ahead = np.array([1, 2, 5, 12])
    z = rng.standard_normal((ahead.size, each, M, R))
    ref = of.predh(lvT, ahead, each, mu, phi, sig, z)
    got = handle.predh(lvT, ahead, each, mu, phi, sig, z)
    assert got.shape == (ahead.size, M, R * each)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)


def test_predh_device_normals_have_the_forecast_law(handle):
    rng = np.random.default_rng(6)
    M, R, each = 2, 3, 4000
    lvT = rng.normal(-1, 0.5, (M, R)); mu = rng.normal(-1, 0.2, (M, R))
    phi = rng.uniform(0.8, 0.95, (M, R)); sig = rng.uniform(0.1, 0.3, (M, R))
    ahead = np.array([2, 6])
    got = handle.predh(lvT, ahead, each, mu, phi, sig).reshape(ahead.size, M, each, R)   # slice = j*R + r
    for c, hh in enumerate(ahead):
        m_ref = mu + phi ** hh * (lvT - mu)
        v_ref = sig ** 2 * (1 - phi ** (2 * hh)) / (1 - phi ** 2)
        assert np.abs(got[c].mean(axis=1) - m_ref).max() < 5 * np.sqrt(v_ref.max() / each)
        assert np.abs(got[c].var(axis=1) / v_ref - 1).max() < 0.15


def test_argument_errors(handle):
    from bayesianvars_b200._lib import BvbError
    with pytest.raises(BvbError):
        handle.predh(np.zeros((2, 2)), np.array([2, 1]), 1, np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((2, 2)))


# ------------------------------------------------------------------ insample (src/utilities_cpp.cpp:396-433)
@pytest.mark.parametrize("T,M,p,r,R,factor", [(9, 4, 1, 0, 3, True), (40, 6, 2, 2, 4, True), (130, 11, 2, 1, 3, True),
                                              (25, 5, 1, 6, 2, True), (30, 7, 2, 0, 3, False), (12, 1, 1, 0, 2, False),
                                              (60, 20, 3, 16, 2, True)])
def test_insample_matches_oracle(handle, T, M, p, r, R, factor):
    """x_t' PHI_r and, with prediction, the error term z' chol_upper(Sigma_tr) with the reference's normals injected:
    the device forms it without any M x M factorisation (structured Cholesky of D + G G' / one substitution with U);
    1e-9 relative against the oracle's dense chol (measured ~1e-14)."""
    rng = np.random.default_rng(T * 7 + M)
    Kp = M * p + 1
    X = rng.standard_normal((T, Kp)); X[:, -1] = 1.0
    PHI = 0.2 * rng.standard_normal((Kp, M, R))
    nf = r if factor else 0
    fl = rng.standard_normal((M, nf, R))
    U = 0.3 * rng.standard_normal((M * (M - 1) // 2, R))
    lv = rng.normal(-1.0, 0.7, (T, M + nf, R))
    z = rng.standard_normal((T, R, M))
    ref0 = of.insample(X, PHI, U, fl, lv, False, factor)
    got0 = handle.insample(X, PHI, U, fl, lv, False, factor)
    assert got0.shape == (T, M, R)
    np.testing.assert_allclose(got0, ref0, rtol=1e-12, atol=1e-13)
    ref = of.insample(X, PHI, U, fl, lv, True, factor, z)
    got = handle.insample(X, PHI, U, fl, lv, True, factor, z)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-10)
    assert np.abs(got - got0).max() > 1e-3            # the error term is really there


def test_insample_device_normals_have_the_error_law(handle):
    """Without injected normals: residuals y_pred - x_t' PHI have covariance Sigma_t (factor structure, one draw)."""
    rng = np.random.default_rng(3)
    T, M, r, R = 3, 4, 2, 6000
    X = np.ones((T, 1)); PHI = np.zeros((1, M, R))
    fl1 = rng.standard_normal((M, r)); lv1 = rng.normal(-0.5, 0.4, (T, M + r))
    fl = np.repeat(fl1[:, :, None], R, axis=2); lv = np.repeat(lv1[:, :, None], R, axis=2)
    got = handle.insample(X, PHI, None, fl, lv, True, True)
    for t in range(T):
        S = (fl1 * np.exp(lv1[t, M:])) @ fl1.T + np.diag(np.exp(lv1[t, :M]))
        emp = np.cov(got[t])
        assert np.abs(emp - S).max() < 0.12 * np.abs(S).max(), (t, np.abs(emp - S).max())
