"""Oracle: one sweep of the univariate SV sampler, vectorised over series
(TEST INFRASTRUCTURE ONLY).

The reference calls `stochvol::update_fast_sv` (src/bvar_cpp.cpp:742, and inside
`factorstochvol::update_fsv`, :617) - third-party code whose source is NOT under
/root/reference (stochvol >= 3.0.3, DESCRIPTION:34).  This restates the published
algorithm (Kastner & Fruehwirth-Schnatter 2014; Omori et al. 2007; McCausland et
al. 2011) with the expert settings of src/bvar_cpp.cpp:375-398.  PARITY UNPINNED:
checks against it are distributional.

Independent of the CUDA code: written against the papers with NumPy, loops over t
for the tridiagonal recursions, vectorised over the S series.
"""
from __future__ import annotations

import numpy as np

MIX_P = np.array([0.00609, 0.04775, 0.13057, 0.20674, 0.22715, 0.18842, 0.12047, 0.05591, 0.01575, 0.00115])
MIX_M = np.array([1.92677, 1.34744, 0.73504, 0.02266, -0.85173, -1.97278, -3.46788, -5.55246, -8.68384, -14.65000])
MIX_V2 = np.array([0.11265, 0.17788, 0.26768, 0.40611, 0.62699, 0.98583, 1.57469, 2.54498, 4.16591, 7.33342])
B011INV, B022INV = 1e-8, 1e-12


def _lognorm(x, mean, sd):
    return -np.log(sd) - 0.5 * ((x - mean) / sd) ** 2


def _logbeta(phi, a, b):
    return (a - 1.0) * np.log(0.5 * (1.0 + phi)) + (b - 1.0) * np.log(0.5 * (1.0 - phi))


def draw_indicators(ystar, h, rng):
    """P(r_t = k) ~ p_k N(y*_t - h_t; m_k, v_k^2); (T, S) int array."""
    d = (ystar - h)[..., None] - MIX_M
    lw = np.log(MIX_P) - 0.5 * np.log(MIX_V2) - 0.5 * d * d / MIX_V2
    lw -= lw.max(axis=-1, keepdims=True)
    w = np.exp(lw)
    cdf = np.cumsum(w, axis=-1)
    u = rng.uniform(size=ystar.shape)[..., None] * cdf[..., -1:]
    return np.minimum((u >= cdf).sum(axis=-1), 9)


def draw_latent(ystar, r, mu, phi, sigma, h0_var, rng):
    """AWOL draw of h_{0:T} | r, theta.  Returns (h0 (S,), h (T, S))."""
    T, S = ystar.shape
    s2inv = 1.0 / sigma ** 2
    phi2 = phi ** 2
    B0inv = np.where(h0_var <= 0, 1.0 - phi2, 1.0 / np.where(h0_var <= 0, 1.0, h0_var))
    prec = 1.0 / MIX_V2[r]
    od = np.empty((T + 1, S)); c = np.empty((T + 1, S))
    od[0] = (B0inv + phi2) * s2inv
    c[0] = mu * (B0inv - phi * (1 - phi)) * s2inv
    od[1:T] = prec[:T - 1] + (1 + phi2) * s2inv
    od[T] = prec[T - 1] + s2inv
    c[1:] = (ystar - MIX_M[r]) * prec
    c[1:T] += mu * (1 - phi) ** 2 * s2inv
    c[T] += mu * (1 - phi) * s2inv
    off = -phi * s2inv
    Ld = np.empty((T + 1, S)); Lo = np.empty((T + 1, S)); a = np.empty((T + 1, S))
    Ld[0] = np.sqrt(od[0]); Lo[0] = off / Ld[0]; a[0] = c[0] / Ld[0]
    for t in range(1, T + 1):
        Ld[t] = np.sqrt(od[t] - Lo[t - 1] ** 2)
        Lo[t] = off / Ld[t]
        a[t] = (c[t] - Lo[t - 1] * a[t - 1]) / Ld[t]
    eps = rng.standard_normal((T + 1, S))
    hh = np.empty((T + 1, S))
    hh[T] = (a[T] + eps[T]) / Ld[T]
    for t in range(T - 1, -1, -1):
        hh[t] = (a[t] + eps[t] - Lo[t] * hh[t + 1]) / Ld[t]
    return hh[0], hh[1:]


def draw_latent_parallel(ystar, r, mu, phi, sigma, h0_var, rng):
    """The SAME law as draw_latent - h_{0:T} | r, theta ~ N(Omega^-1 c, Omega^-1) - drawn the way the CUDA series kernel
    does it (k8_fsv.cu::sv_state_draw_parallel): perturb, then solve.  Omega = B'B + D with B the bidiagonal AR(1)
    innovation map (row 0: sqrt(B0inv)/sigma, row t: (h_t - phi h_{t-1})/sigma) and D = diag(0, 1/v^2_{r_1}, ..), so
    eta = B'xi + sqrt(D) zeta (xi ~ N(0, I_{T+1}), zeta ~ N(0, I_T)) has covariance Omega and h = Omega^-1 (c + eta) has
    the required law; the tridiagonal solve is a parallel cyclic reduction (log2(T+1) sweeps, every row eliminates its
    neighbours at distance 1, 2, 4, ..).  Written for one series (ystar, r: (T,)); returns (h0, h (T,)).
    Test infrastructure: tests/test_oracle_sv_parallel.py checks both samplers against the exact moments on the CPU."""
    ystar = np.asarray(ystar, float); r = np.asarray(r, int)
    T = ystar.size
    s2inv, phi2 = 1.0 / sigma ** 2, phi ** 2
    B0inv = (1.0 - phi2) if h0_var <= 0 else 1.0 / h0_var
    prec = 1.0 / MIX_V2[r]
    b = np.empty(T + 1); c = np.empty(T + 1)
    b[0] = (B0inv + phi2) * s2inv
    c[0] = mu * (B0inv - phi * (1 - phi)) * s2inv
    b[1:T] = prec[:T - 1] + (1 + phi2) * s2inv
    b[T] = prec[T - 1] + s2inv
    c[1:] = (ystar - MIX_M[r]) * prec
    c[1:T] += mu * (1 - phi) ** 2 * s2inv
    c[T] += mu * (1 - phi) * s2inv
    off = -phi * s2inv
    xi, zeta = rng.standard_normal(T + 1), rng.standard_normal(T)
    sinv = 1.0 / sigma
    eta = np.empty(T + 1)
    eta[0] = np.sqrt(B0inv) * sinv * xi[0] - phi * sinv * xi[1]
    eta[1:T] = sinv * xi[1:T] - phi * sinv * xi[2:T + 1] + zeta[:T - 1] * np.sqrt(prec[:T - 1])
    eta[T] = sinv * xi[T] + zeta[T - 1] * np.sqrt(prec[T - 1])
    rhs = c + eta
    a = np.full(T + 1, off); a[0] = 0.0          # sub-diagonal
    g = np.full(T + 1, off); g[T] = 0.0          # super-diagonal
    st = 1
    idx = np.arange(T + 1)
    while st <= T:
        im, ip = idx - st, idx + st
        k1 = np.where(im >= 0, a / b[np.maximum(im, 0)], 0.0)
        k2 = np.where(ip <= T, g / b[np.minimum(ip, T)], 0.0)
        bn = b - np.where(im >= 0, g[np.maximum(im, 0)], 0.0) * k1 - np.where(ip <= T, a[np.minimum(ip, T)], 0.0) * k2
        rn = rhs - np.where(im >= 0, rhs[np.maximum(im, 0)], 0.0) * k1 - np.where(ip <= T, rhs[np.minimum(ip, T)], 0.0) * k2
        an = -np.where(im >= 0, a[np.maximum(im, 0)], 0.0) * k1
        gn = -np.where(ip <= T, g[np.minimum(ip, T)], 0.0) * k2
        a, b, g, rhs = an, bn, gn, rn
        st *= 2
    hh = rhs / b
    return hh[0], hh[1:]


def draw_theta_centered(h0, h, mu, phi, sigma, prior, rng):
    T, S = h.shape
    hp = np.vstack([h0[None, :], h[:-1]])
    s_prev, s_cur = hp.sum(0), h.sum(0)
    s_cross, s_prev2, s_cur2 = (hp * h).sum(0), (hp * hp).sum(0), (h * h).sum(0)
    mu, phi, sigma = mu.copy(), phi.copy(), sigma.copy()
    rate, h0v = prior["s2_rate"], prior["h0_var"]
    free = ~prior["mu_const"]
    stat = h0v <= 0
    # ---- free mu: 2-block
    x00, x01, x11 = T + B011INV, s_prev, s_prev2 + B022INV
    det = x00 * x11 - x01 ** 2
    B00, B01, B11 = x11 / det, -x01 / det, x00 / det
    b0, b1 = B00 * s_cur + B01 * s_cross, B01 * s_cur + B11 * s_cross
    # y'y - b'(X'X + B0^-1) b = |y - X b|^2 + b' B0^-1 b, from the residuals: the raw-sum form cancels
    # catastrophically for a nearly deterministic series and can come out <= 0
    rss_free = ((h - b0[None, :] - b1[None, :] * hp) ** 2).sum(0)
    CT_free = np.maximum(0.5 * (rss_free + B011INV * b0 ** 2 + B022INV * b1 ** 2), 1e-300)
    stat0 = np.where(stat, (1 - phi ** 2) * h0 ** 2, h0 ** 2 / np.where(stat, 1.0, h0v))
    CT_fix = np.maximum(0.5 * (((h - phi[None, :] * hp) ** 2).sum(0) + stat0), 1e-300)
    cT = np.where(free, 0.5 * (T - 1.0), 0.5 * T)
    CT = np.where(free, CT_free, CT_fix)
    s2_prop = CT / rng.gamma(cT)
    acc = np.log(rng.uniform(size=S)) < (sigma ** 2 - s2_prop) * rate
    sigma = np.where(acc, np.sqrt(s2_prop), sigma)
    # phi (and gamma) proposals
    z1, z2, u = rng.standard_normal(S), rng.standard_normal(S), rng.uniform(size=S)
    den = s_prev2 + B022INV
    phi_prop = np.where(free, b1 + sigma * np.sqrt(B11) * z1, s_cross / den + sigma / np.sqrt(den) * z1)
    cv = np.maximum(B00 - B01 ** 2 / B11, 0.0)
    gam_prop = b0 + B01 / B11 * (phi_prop - b1) + sigma * np.sqrt(cv) * z2
    ok = (phi_prop > -1) & (phi_prop < 1)
    pp = np.where(ok, phi_prop, 0.0)
    gam_old = mu * (1 - phi)
    mu_prop = np.where(free, gam_prop / (1 - pp), 0.0)
    sd0_new = np.where(stat, sigma / np.sqrt(1 - pp ** 2), sigma * np.sqrt(np.where(stat, 1.0, h0v)))
    sd0_old = np.where(stat, sigma / np.sqrt(1 - phi ** 2), sigma * np.sqrt(np.where(stat, 1.0, h0v)))
    lr = _lognorm(h0, mu_prop, sd0_new) - _lognorm(h0, mu, sd0_old)
    lr = np.where(~free & ~stat, 0.0, lr)
    mm, ms = prior["mu_mean"], prior["mu_sd"]
    lr += np.where(free, _lognorm(gam_prop, mm * (1 - pp), ms * (1 - pp)) - _lognorm(gam_old, mm * (1 - phi), ms * (1 - phi)), 0.0)
    lr += _logbeta(pp, prior["phi_a"], prior["phi_b"]) - _logbeta(phi, prior["phi_a"], prior["phi_b"])
    lr += np.where(free, _lognorm(gam_old, 0, sigma / np.sqrt(B011INV)) - _lognorm(gam_prop, 0, sigma / np.sqrt(B011INV)), 0.0)
    lr += _lognorm(phi, 0, sigma / np.sqrt(B022INV)) - _lognorm(pp, 0, sigma / np.sqrt(B022INV))
    acc = ok & (np.log(u) < lr)
    mu = np.where(acc & free, mu_prop, mu)
    phi = np.where(acc, phi_prop, phi)
    return mu, phi, sigma


def draw_theta_noncentered(ystar, r, h0, h, mu, phi, sigma, prior, rng):
    T, S = h.shape
    free = ~prior["mu_const"]
    x = (h - mu) / sigma
    x0 = (h0 - mu) / sigma
    xp = np.vstack([x0[None, :], x[:-1]])
    w = 1.0 / MIX_V2[r]
    yv = ystar - MIX_M[r]
    Bsig_inv = 2.0 * prior["s2_rate"]
    p00 = w.sum(0) + 1.0 / prior["mu_sd"] ** 2
    p01 = (w * x).sum(0)
    p11 = (w * x * x).sum(0) + Bsig_inv
    l0 = (w * yv).sum(0) + prior["mu_mean"] / prior["mu_sd"] ** 2
    l1 = (w * x * yv).sum(0)
    det = p00 * p11 - p01 ** 2
    C00, C01, C11 = p11 / det, -p01 / det, p00 / det
    m0, m1 = C00 * l0 + C01 * l1, C01 * l0 + C11 * l1
    z1, z0 = rng.standard_normal(S), rng.standard_normal(S)
    sg_free = m1 + np.sqrt(C11) * z1
    mu_free = m0 + C01 / C11 * (sg_free - m1) + np.sqrt(np.maximum(C00 - C01 ** 2 / C11, 0)) * z0
    sg_fix = l1 / p11 + z1 / np.sqrt(p11)
    sg_n = np.where(free, sg_free, sg_fix)
    mu_n = np.where(free, mu_free, 0.0)
    den = (xp * xp).sum(0) + B022INV
    phi_prop = (xp * x).sum(0) / den + rng.standard_normal(S) / np.sqrt(den)
    u = rng.uniform(size=S)
    ok = (phi_prop > -1) & (phi_prop < 1)
    pp = np.where(ok, phi_prop, 0.0)
    stat = prior["h0_var"] <= 0
    lr = np.where(stat, _lognorm(x0, 0, 1 / np.sqrt(1 - pp ** 2)) - _lognorm(x0, 0, 1 / np.sqrt(1 - phi ** 2)), 0.0)
    lr += _logbeta(pp, prior["phi_a"], prior["phi_b"]) - _logbeta(phi, prior["phi_a"], prior["phi_b"])
    lr += _lognorm(phi, 0, 1 / np.sqrt(B022INV)) - _lognorm(pp, 0, 1 / np.sqrt(B022INV))
    phi = np.where(ok & (np.log(u) < lr), phi_prop, phi)
    h_new = mu_n + sg_n * x
    h0_new = mu_n + sg_n * x0
    return mu_n, phi, np.abs(sg_n), h0_new, h_new


def make_prior(S, mu_mean=0.0, mu_sd=10.0, phi_a=10.0, phi_b=3.0, s2_rate=0.5, h0_var=-1.0, mu_const=False):
    f = lambda v, dt=float: np.broadcast_to(np.asarray(v, dt), (S,)).copy()
    return dict(mu_mean=f(mu_mean), mu_sd=f(mu_sd), phi_a=f(phi_a), phi_b=f(phi_b), s2_rate=f(s2_rate),
                h0_var=f(h0_var), mu_const=f(mu_const, bool))


def update_fast_sv(ystar, h, h0, mu, phi, sigma, prior, rng):
    """One sweep for S series at once.  ystar, h: (T, S).  Returns new (h, h0, mu, phi, sigma, r)."""
    mu = np.where(prior["mu_const"], 0.0, mu)
    r = draw_indicators(ystar, h, rng)                                  # step (c)
    h0, h = draw_latent(ystar, r, mu, phi, sigma, prior["h0_var"], rng)  # step (a)
    mu, phi, sigma = draw_theta_centered(h0, h, mu, phi, sigma, prior, rng)  # step (b), centered
    mu, phi, sigma, h0, h = draw_theta_noncentered(ystar, r, h0, h, mu, phi, sigma, prior, rng)  # ASIS
    return h, h0, mu, phi, sigma, r
