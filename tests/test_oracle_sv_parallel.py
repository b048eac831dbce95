"""CPU: the two samplers of the SV latent states - the sequential all-without-a-loop recursion (oracle/sv.py::draw_latent,
what stochvol::update_fast_sv does, called at src/bvar_cpp.cpp:742) and the parallel perturb-then-solve form the CUDA series
kernel runs (draw_latent_parallel, restating k8_fsv.cu::sv_state_draw_parallel) - draw from the same Gaussian: sample mean,
variances, lag-1 and corner covariances of both within 5 standard errors of the exact Omega^-1 c and Omega^-1; and the cyclic
reduction solves the tridiagonal system to rounding error."""
import numpy as np
import pytest

from oracle import sv as osv


def exact_moments(ystar, r, mu, phi, sigma, h0_var):
    T = ystar.size
    s2inv, phi2 = 1.0 / sigma ** 2, phi ** 2
    B0inv = (1.0 - phi2) if h0_var <= 0 else 1.0 / h0_var
    prec = 1.0 / osv.MIX_V2[r]
    Om = np.zeros((T + 1, T + 1)); c = np.zeros(T + 1)
    Om[0, 0] = (B0inv + phi2) * s2inv
    c[0] = mu * (B0inv - phi * (1 - phi)) * s2inv
    for t in range(1, T + 1):
        Om[t, t] = prec[t - 1] + ((1 + phi2) if t < T else 1.0) * s2inv
        c[t] = (ystar[t - 1] - osv.MIX_M[r[t - 1]]) * prec[t - 1] + mu * ((1 - phi) ** 2 if t < T else (1 - phi)) * s2inv
        Om[t, t - 1] = Om[t - 1, t] = -phi * s2inv
    cov = np.linalg.inv(Om)
    return cov @ c, cov, Om, c


def check(draws, mean, cov):
    n = draws.shape[0]
    sd = np.sqrt(np.diag(cov))
    assert (np.abs(draws.mean(axis=0) - mean) < 5.0 * sd / np.sqrt(n)).all()
    dc = draws - mean
    assert (np.abs((dc ** 2).mean(axis=0) - sd ** 2) < 5.0 * sd ** 2 * np.sqrt(2.0 / n)).all()
    e1 = np.diag(cov, 1)
    assert (np.abs((dc[:, 1:] * dc[:, :-1]).mean(axis=0) - e1) < 5.0 * np.sqrt((sd[1:] ** 2 * sd[:-1] ** 2 + e1 ** 2) / n)).all()


@pytest.mark.parametrize("T,mu,phi,sigma,h0_var", [(12, -1.0, 0.95, 0.3, -1.0), (37, -9.0, 0.8, 0.8, -1.0), (64, 0.0, 0.99, 0.05, 2.0),
                                                   (1, 0.3, 0.5, 1.0, -1.0), (2, 0.3, -0.5, 1.5, 4.0)])
def test_both_samplers_have_the_exact_conditional_law(T, mu, phi, sigma, h0_var):
    rng = np.random.default_rng(100 + T)
    ystar = mu + rng.standard_normal(T) + np.log(rng.standard_normal(T) ** 2 + 1e-6)
    r = rng.integers(0, 10, T)
    mean, cov, Om, c = exact_moments(ystar, r, mu, phi, sigma, h0_var)
    n = 4000
    par = np.array([np.concatenate([[h0], h]) for h0, h in (osv.draw_latent_parallel(ystar, r, mu, phi, sigma, h0_var, rng) for _ in range(n))])
    check(par, mean, cov)
    h0, hh = osv.draw_latent(np.repeat(ystar[:, None], n, 1), np.repeat(r[:, None], n, 1), mu, phi, sigma, np.full(n, h0_var), rng)
    check(np.vstack([h0[None], hh]).T, mean, cov)


def test_cyclic_reduction_solves_the_system():
    """With the noise switched off the parallel form returns Omega^-1 c."""
    class Zero:
        def standard_normal(self, n):
            return np.zeros(n)
    rng = np.random.default_rng(3)
    T = 100
    ystar, r = rng.standard_normal(T) - 1.0, rng.integers(0, 10, T)
    mean, cov, Om, c = exact_moments(ystar, r, -0.5, 0.97, 0.2, -1.0)
    h0, h = osv.draw_latent_parallel(ystar, r, -0.5, 0.97, 0.2, -1.0, Zero())
    got = np.concatenate([[h0], h])
    assert np.abs(got - mean).max() <= 1e-10 * np.abs(mean).max()
