"""GPU parity of predh and vcov_cpp (src/utilities_cpp.cpp:181-212, 436-472) through the C ABI against the oracle
restatement (oracle/fitted.py): pure element-wise FP64 arithmetic, tolerance 1e-12 relative."""
import numpy as np
import pytest

from oracle import fitted as of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,M,r,R,factor", [(7, 5, 2, 4, True), (130, 6, 1, 3, True), (9, 4, 0, 2, True), (11, 5, 0, 3, False),
                                            (200, 13, 0, 2, False), (3, 1, 0, 2, False)])
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
