"""CPU pins of the oracle restatement of insample / vcov_cpp / predh (src/utilities_cpp.cpp:181-212, 396-472):
identities the reference's own definitions imply (no golden vectors exist in its test-suite for these paths)."""
import numpy as np

from oracle import fitted as of
from oracle.predict import build_sigma


def _draws(rng, T, M, p, r, nsave, factor):
    Kp = M * p + 1
    X = rng.standard_normal((T, Kp)); X[:, -1] = 1.0
    PHI = 0.2 * rng.standard_normal((Kp, M, nsave))
    fl = rng.standard_normal((M, r, nsave)) if factor else np.zeros((0, 0, 0))
    U = 0.3 * rng.standard_normal((M * (M - 1) // 2, nsave))
    lv = rng.normal(-1.0, 0.4, (T, M + (r if factor else 0), nsave))
    return X, PHI, U, fl, lv


def test_insample_without_error_term_is_the_fitted_value():
    rng = np.random.default_rng(0)
    X, PHI, U, fl, lv = _draws(rng, 9, 4, 2, 2, 3, True)
    y = of.insample(X, PHI, U, fl, lv, False, True)
    assert np.allclose(y, np.einsum("tk,kmr->tmr", X, PHI), rtol=1e-14)


def test_insample_error_term_has_the_model_covariance():
    rng = np.random.default_rng(1)
    for factor in (True, False):
        X, PHI, U, fl, lv = _draws(rng, 3, 3, 1, 1, 2, factor)
        n = 4000
        acc = np.zeros((3, 2, 3, 3))
        for _ in range(n):
            z = rng.standard_normal((3, 2, 3))
            e = of.insample(X, PHI, U, fl, lv, True, factor, z) - of.insample(X, PHI, U, fl, lv, False, factor)
            acc += np.einsum("tar,tbr->trab", e, e)
        acc /= n
        for t in range(3):
            for r in range(2):
                S = build_sigma(factor, fl[:, :, r] if factor else None, lv[t, :, r], lv.shape[1] - 3, 3,
                                None if factor else U[:, r])
                assert np.abs(acc[t, r] - S).max() < 6 * np.abs(S).max() / np.sqrt(n) + 0.05 * np.abs(S).max()


def test_vcov_layout_and_symmetry():
    rng = np.random.default_rng(2)
    T, M, r, nsave = 5, 4, 2, 3
    for factor in (True, False):
        X, PHI, U, fl, lv = _draws(rng, T, M, 1, r, nsave, factor)
        out = of.vcov_cpp(factor, fl, lv, U, M, r if factor else 0)
        assert out.shape == (T * M, M, nsave)
        cube = out.reshape(T, M, M, nsave, order="F")          # what vcov.bayesianVARs_bvar does (utility_functions.R:226)
        for t in range(T):
            for i in range(nsave):
                S = build_sigma(factor, fl[:, :, i] if factor else None, lv[t, :, i], r if factor else 0, M,
                                None if factor else U[:, i])
                assert np.allclose(cube[t, :, :, i], S, rtol=1e-14)
                assert np.allclose(S, S.T, rtol=1e-13) and np.all(np.linalg.eigvalsh(S) > 0)


def test_predh_is_the_ar1_forecast_law():
    rng = np.random.default_rng(3)
    M, R, each = 3, 4, 2
    lvT = rng.normal(-1, 0.5, (M, R)); mu = rng.normal(-1, 0.2, (M, R))
    phi = rng.uniform(0.8, 0.98, (M, R)); sig = rng.uniform(0.1, 0.3, (M, R))
    ahead = np.array([1, 3, 6])
    z0 = np.zeros((ahead.size, each, M, R))
    mean = of.predh(lvT, ahead, each, mu, phi, sig, z0)
    z1 = np.ones((ahead.size, each, M, R))
    sd = of.predh(lvT, ahead, each, mu, phi, sig, z1) - mean
    for c, hh in enumerate(ahead):
        m_ref = mu + phi ** hh * (lvT - mu)                       # E[h_{T+h} | h_T]
        v_ref = sig ** 2 * (1 - phi ** (2 * hh)) / (1 - phi ** 2)  # Var[h_{T+h} | h_T]
        for j in range(each):
            assert np.allclose(mean[c, :, j * R:(j + 1) * R], m_ref, rtol=1e-13)
            assert np.allclose(sd[c, :, j * R:(j + 1) * R] ** 2, v_ref, rtol=1e-12)
