"""Oracle: hyper-parameter updates of the shrinkage priors (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/sample_coefficients.cpp:190-377 with NumPy; RNG
draws come from a NumPy Generator (R's stream cannot be reproduced) and GIG
variates from oracle.gig (GIGrvg is third-party: parity unpinned).  Every
function updates the `state` dict in place, exactly as the reference mutates its
by-reference arguments, for the coefficients `ind` of one group.
"""
from __future__ import annotations

import numpy as np

from .gig import rgig


def _grid_update(state, g, n, cfg, rng):
    """Discrete update of `a` (and c) over the grid: :227-244 / :297-314."""
    a_vec, a_weight, norm_consts = cfg["a_vec"], cfg["a_weight"], cfg["norm_consts"]
    ind = state["groups_idx"][g]
    xi = state["xi"][g]
    logprobs = np.log(a_weight) + n * (a_vec * np.log(xi) - norm_consts) + (a_vec - 1.0) * np.log(state["lambda"][ind]).sum()
    if not cfg["c_rel_a"]:
        logprobs += -2.0 * xi * state["c"][g] / a_vec - state["b"][g] * np.log(a_vec)
    w = np.exp(logprobs - logprobs.max())
    i = rng.choice(a_vec.size, p=w / w.sum())
    state["a"][g] = a_vec[i]
    if cfg["c_rel_a"]:
        state["c"][g] = cfg["c_vec"][i]


def sample_V_i_DL(state, coefs, g, cfg, rng):
    """:263-316."""
    ind = state["groups_idx"][g]
    n = ind.size
    a, xi = state["a"][g], state["xi"][g]
    lam = rgig(a - 1.0, 2.0 * np.abs(coefs[ind]), 2.0 * xi, rng)          # :279
    psi = rgig(0.5, coefs[ind] ** 2 / lam ** 2, 1.0, rng)                  # :280-281
    zeta = lam.sum()
    if cfg["DL_plus"]:
        state["xi"][g] = rng.gamma(state["b"][g] + n * a) / (2.0 * state["c"][g] / a + zeta)   # :285
    if cfg["GL_tol"] > 0:
        theta = np.maximum(lam / zeta, cfg["GL_tol"])
        lam = theta * zeta
    state["lambda"][ind], state["psi"][ind] = lam, psi
    state["V_i"][ind] = psi * lam ** 2
    if cfg["DL_hyper"]:
        _grid_update(state, g, n, cfg, rng)


def sample_V_i_GT(state, coefs, g, cfg, rng):
    """:190-245 (R2D2: kernel exponential, vs 1/2; NG: kernel normal, vs 1)."""
    ind = state["groups_idx"][g]
    n = ind.size
    a, xi, vs = state["a"][g], state["xi"][g], cfg["GT_vs"]
    c2 = coefs[ind] ** 2
    psi = state["psi"][ind]
    if cfg["GT_priorkernel"] == "exponential":
        psi = 1.0 / rgig(-0.5, 1.0, c2 / (state["lambda"][ind] * vs), rng)  # :206
    lam = rgig(a - 0.5, c2 / (psi * vs), 2.0 * xi, rng)                     # :209
    zeta = lam.sum()
    if cfg["GL_tol"] > 0:
        lam = np.maximum(lam / zeta, cfg["GL_tol"]) * zeta
    state["xi"][g] = rng.gamma(state["b"][g] + n * a) / (2.0 * state["c"][g] / a + zeta)   # :219
    state["lambda"][ind], state["psi"][ind] = lam, psi
    state["V_i"][ind] = psi * lam * vs if cfg["GT_priorkernel"] == "exponential" else lam * vs
    if cfg["GT_hyper"]:
        _grid_update(state, g, n, cfg, rng)


def sample_V_i_HS(state, coefs, g, rng):
    """:247-261 (R::rgamma(shape, scale))."""
    ind = state["groups_idx"][g]
    n = ind.size
    c2 = coefs[ind] ** 2
    zeta = state["zeta"][g]
    theta = 1.0 / (rng.gamma(1.0, size=n) / (1.0 / state["nu"][ind] + c2 / (2.0 * zeta)))
    nu = 1.0 / (rng.gamma(1.0, size=n) / (1.0 + 1.0 / theta))
    zeta = 1.0 / (rng.gamma((n + 1) / 2.0) / (1.0 / state["varpi"][g] + 0.5 * (c2 / theta).sum()))
    state["varpi"][g] = 1.0 / (rng.gamma(1.0) / (1.0 + 1.0 / zeta))
    state["zeta"][g] = zeta
    state["theta"][ind], state["nu"][ind] = theta, nu
    state["V_i"][ind] = theta * zeta


def sample_V_i_SSVS_beta(state, coefs, g, cfg, rng):
    """:318-353."""
    ind = state["groups_idx"][g]
    n = ind.size
    t0, t1, p = cfg["SSVS_tau0"][ind], cfg["SSVS_tau1"][ind], state["p_i"][ind]
    u1 = -np.log(t1) - 0.5 * (coefs[ind] / t1) ** 2 + np.log(p)
    u2 = -np.log(t0) - 0.5 * (coefs[ind] / t0) ** 2 + np.log(1.0 - p)
    gst = 1.0 / (1.0 + np.exp(u2 - u1))
    gam = (rng.uniform(size=n) < gst).astype(float)
    state["gammas"][ind] = gam
    state["V_i"][ind] = np.where(gam == 1.0, t1 ** 2, t0 ** 2)
    if cfg["SSVS_hyper"]:
        incl = gam.sum()
        state["p_i"][ind] = rng.beta(cfg["SSVS_s_a"] + incl, cfg["SSVS_s_b"] + n - incl)


def sample_V_i_HMP(state, PHI_diff, V_i_prep, i_ol, i_cl, cfg, rng):
    """:355-368 (note the INTEGER division n_ol/2, n_cl/2)."""
    s1, r1 = cfg["lambda_1"]
    s2, r2 = cfg["lambda_2"]
    state["lambda_1"] = float(rgig(s1 - i_ol.size // 2, (PHI_diff[i_ol] ** 2 / V_i_prep[i_ol]).sum(), 2.0 * r1, rng))
    state["lambda_2"] = float(rgig(s2 - i_cl.size // 2, (PHI_diff[i_cl] ** 2 / V_i_prep[i_cl]).sum(), 2.0 * r2, rng))
    state["V_i"][i_ol] = state["lambda_1"] * V_i_prep[i_ol]
    state["V_i"][i_cl] = state["lambda_2"] * V_i_prep[i_cl]
