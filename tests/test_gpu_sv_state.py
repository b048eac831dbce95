"""The latent-state draw of the SV update on the GPU ("perturb, then solve" + parallel cyclic reduction, k8_fsv.cu) has
the law N(Omega^-1 c, Omega^-1) the all-without-a-loop sampler of stochvol draws from (oracle/sv.py::draw_latent, the
step the reference takes from stochvol::update_fast_sv at src/bvar_cpp.cpp:761-777).  The GPU does not reproduce the
oracle's variates (different algorithm for the same distribution), so parity is in distribution: with n_rep draws the
sample mean is within 5 standard errors of the exact conditional mean at every t, and variances / lag-1 covariances are
within 5 standard errors (normal theory) of the exact Omega^-1 entries.  The oracle's own draws pass the same check."""
import numpy as np
import pytest

from oracle import sv as osv

pytestmark = pytest.mark.gpu


def exact_moments(ystar, r, mu, phi, sigma, h0_var):
    T = ystar.size
    s2inv, phi2 = 1.0 / sigma ** 2, phi ** 2
    B0inv = (1.0 - phi2) if h0_var <= 0 else 1.0 / h0_var
    prec = 1.0 / osv.MIX_V2[r]
    Om = np.zeros((T + 1, T + 1))
    c = np.zeros(T + 1)
    Om[0, 0] = (B0inv + phi2) * s2inv
    c[0] = mu * (B0inv - phi * (1 - phi)) * s2inv
    for t in range(1, T + 1):
        Om[t, t] = prec[t - 1] + ((1 + phi2) if t < T else 1.0) * s2inv
        c[t] = (ystar[t - 1] - osv.MIX_M[r[t - 1]]) * prec[t - 1] + mu * ((1 - phi) ** 2 if t < T else (1 - phi)) * s2inv
        Om[t, t - 1] = Om[t - 1, t] = -phi * s2inv
    cov = np.linalg.inv(Om)
    return cov @ c, cov


def check(draws, mean, cov):
    n = draws.shape[0]
    sd = np.sqrt(np.diag(cov))
    assert (np.abs(draws.mean(axis=0) - mean) < 5.0 * sd / np.sqrt(n)).all()
    dc = draws - mean
    var = (dc ** 2).mean(axis=0)
    assert (np.abs(var - sd ** 2) < 5.0 * sd ** 2 * np.sqrt(2.0 / n)).all()
    c1 = (dc[:, 1:] * dc[:, :-1]).mean(axis=0)
    e1 = np.diag(cov, 1)
    se1 = np.sqrt((sd[1:] ** 2 * sd[:-1] ** 2 + e1 ** 2) / n)
    assert (np.abs(c1 - e1) < 5.0 * se1).all()
    # a long-range entry: first against last state
    e = cov[0, -1]
    assert abs((dc[:, 0] * dc[:, -1]).mean() - e) < 5.0 * np.sqrt((sd[0] ** 2 * sd[-1] ** 2 + e ** 2) / n)


@pytest.mark.parametrize("T,mu,phi,sigma,h0_var", [(40, -1.0, 0.95, 0.3, -1.0), (235, -9.0, 0.8, 0.8, -1.0), (500, 0.0, 0.99, 0.05, 2.0),
                                                   (513, 1.5, -0.5, 1.5, -1.0), (1, 0.3, 0.5, 1.0, -1.0), (2, 0.3, 0.5, 1.0, 4.0)])
def test_state_draw_has_the_exact_conditional_law(handle, T, mu, phi, sigma, h0_var):
    rng = np.random.default_rng(T)
    h = mu + sigma / np.sqrt(1 - phi ** 2) * rng.standard_normal(T)
    ystar = h + np.log(rng.standard_normal(T) ** 2 + 1e-6)
    r = rng.integers(0, 10, T)
    mean, cov = exact_moments(ystar, r, mu, phi, sigma, h0_var)
    n = 20000
    handle.reseed(99 + T)
    g = handle.sv_latent_draw(ystar, r, mu, phi, sigma, h0_var, n_rep=n)
    assert g.shape == (n, T + 1) and np.isfinite(g).all()
    check(g, mean, cov)
    # the oracle's sampler against the same moments (so the check itself is pinned)
    S = 4000
    h0, hh = osv.draw_latent(np.repeat(ystar[:, None], S, 1), np.repeat(r[:, None], S, 1), mu, phi, sigma, np.full(S, h0_var), rng)
    check(np.vstack([h0[None], hh]).T, mean, cov)


def test_state_draw_replicas_are_independent_and_deterministic(handle):
    rng = np.random.default_rng(3)
    T = 100
    ystar, r = rng.standard_normal(T) - 1.0, rng.integers(0, 10, T)
    handle.reseed(7)
    a = handle.sv_latent_draw(ystar, r, 0.0, 0.9, 0.4, -1.0, n_rep=64)
    handle.reseed(7)
    b = handle.sv_latent_draw(ystar, r, 0.0, 0.9, 0.4, -1.0, n_rep=64)
    assert np.array_equal(a, b)
    handle.reseed(8)
    c = handle.sv_latent_draw(ystar, r, 0.0, 0.9, 0.4, -1.0, n_rep=64)
    assert not np.array_equal(a, c)
    assert np.abs(np.corrcoef(a[:, 50], np.roll(a[:, 50], 1))[0, 1]) < 0.5


def test_state_draw_rejects_bad_arguments(handle):
    with pytest.raises(Exception):
        handle.sv_latent_draw(np.zeros(4), np.array([0, 1, 2, 11]), 0.0, 0.5, 1.0)
    with pytest.raises(Exception):
        handle.sv_latent_draw(np.zeros(4), np.zeros(4, int), 0.0, 1.0, 1.0)
