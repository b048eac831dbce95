"""Distributional checks of the SV sampler: (i) the oracle (NumPy restatement of
the published algorithm) recovers the parameters of simulated SV data, (ii) the
device code's logic - exercised on the host through the harness around the same
__host__ __device__ functions - agrees with the oracle's posterior within
Monte-Carlo error.  stochvol parity itself is unpinned."""
import ctypes as C
import numpy as np
import pytest

from oracle import sv as osv
from tests.test_rng_host import hostrng  # noqa: F401  (fixture)
from tests.stat_helpers import batch_means_var

_dp = C.POINTER(C.c_double)


def simulate_sv(T, mu, phi, sigma, rng):
    h = np.empty(T)
    hp = mu + sigma / np.sqrt(1 - phi ** 2) * rng.standard_normal()
    for t in range(T):
        hp = mu + phi * (hp - mu) + sigma * rng.standard_normal()
        h[t] = hp
    return np.exp(h / 2) * rng.standard_normal(T), h


def oracle_chain(y, n, burn, rng, mu_const=False):
    T = y.size
    ystar = np.log(y ** 2)[:, None]
    pr = osv.make_prior(1, mu_const=mu_const)
    h = np.full((T, 1), -1.0); h0 = np.array([-1.0]); mu = np.array([0.0 if mu_const else -1.0])
    phi = np.array([0.8]); sigma = np.array([0.3])
    out = np.empty((n, 3 + 1))
    for it in range(n + burn):
        h, h0, mu, phi, sigma, _ = osv.update_fast_sv(ystar, h, h0, mu, phi, sigma, pr, rng)
        if it >= burn:
            out[it - burn] = [mu[0], phi[0], sigma[0], h.mean()]
    return out


def host_chain(lib, y, n, burn, mu_const=False, seed=5):
    T = y.size
    prior = np.array([0.0, 10.0, 10.0, 3.0, 0.5, -1.0, 1.0 if mu_const else 0.0])
    para = np.array([0.0 if mu_const else -1.0, 0.8, 0.3, -1.0])
    h = np.full(T, -1.0)
    draws = np.empty((n + burn, 4 + T))
    lib.hostrng_sv_chain(T, y.ctypes.data_as(_dp), n + burn, prior.ctypes.data_as(_dp), para.ctypes.data_as(_dp),
                         h.ctypes.data_as(_dp), C.c_ulonglong(seed), draws.ctypes.data_as(_dp))
    d = draws[burn:]
    return np.column_stack([d[:, 0], d[:, 1], d[:, 2], d[:, 4:].mean(axis=1)])


def test_oracle_recovers_sv_parameters():
    rng = np.random.default_rng(42)
    y, h = simulate_sv(1500, -1.0, 0.95, 0.25, rng)
    ch = oracle_chain(y, 1500, 500, rng)
    m = ch.mean(0)
    assert abs(m[0] - (-1.0)) < 0.35 and abs(m[1] - 0.95) < 0.04 and abs(m[2] - 0.25) < 0.1
    assert abs(m[3] - h.mean()) < 0.15


@pytest.mark.parametrize("mu_const", [False, True])
def test_device_logic_matches_oracle_posterior(hostrng, mu_const):
    rng = np.random.default_rng(7)
    y, h = simulate_sv(400, 0.0 if mu_const else -0.5, 0.9, 0.3, rng)
    a = oracle_chain(y, 3000, 500, rng, mu_const)
    b = host_chain(hostrng, y, 3000, 500, mu_const)
    assert np.isfinite(b).all()
    for k in range(4):
        if mu_const and k == 0:
            assert np.all(b[:, 0] == 0.0)
            continue
        se = np.sqrt(batch_means_var(a[:, k]) + batch_means_var(b[:, k]))
        assert abs(a[:, k].mean() - b[:, k].mean()) < 4.5 * se + 1e-3, (k, a[:, k].mean(), b[:, k].mean(), se)
        # spread of the posterior agrees too
        assert 0.6 < a[:, k].std() / b[:, k].std() < 1.6
