"""Oracle: one sweep of the factor stochastic volatility block given the VAR
residuals (TEST INFRASTRUCTURE ONLY).

The reference calls `factorstochvol::update_fsv` (src/bvar_cpp.cpp:617-647) -
third-party code whose source is NOT under /root/reference (factorstochvol >=
1.1.0, DESCRIPTION:26): PARITY UNPINNED, checks are distributional.  This is an
independent NumPy restatement of the published sampler (Kastner,
Fruehwirth-Schnatter & Lopes 2017; Kastner 2019 for the NG prior on the
loadings): SV updates of the M idiosyncratic and r factor series, NG prior
hyper-parameters, row-wise loadings regressions, shallow / deep interweaving,
factor regressions.
"""
from __future__ import annotations

import numpy as np

from . import sv as osv
from .gig import rgig


def update_fsv(st, y, cfg, rng):
    """st: dict(facload MxR, fac RxT, logvar Tx(M+R), logvar0, sv_para 3x(M+R), tau2 MxR, lambda2);
    y: M x T residuals.  Updated in place."""
    M, T = y.shape
    R = st["facload"].shape[1]
    S = M + R
    het = cfg["sv_heteroscedastic"].astype(bool)
    # ---- STEP 1: SV of the idiosyncratic and the factor series
    e = y - st["facload"] @ st["fac"] if R else y
    with np.errstate(divide="ignore"):
        ystar = np.log(np.vstack([e, st["fac"]]) ** 2).T if R else np.log(e ** 2).T     # T x S
    pr = cfg["_svprior"]
    h, h0, mu, phi, sigma, _ = osv.update_fast_sv(ystar, st["logvar"], st["logvar0"], st["sv_para"][0], st["sv_para"][1],
                                                  st["sv_para"][2], pr, rng)
    for s in range(S):
        if het[s]:
            st["logvar"][:, s] = h[:, s]
            st["logvar0"][s] = h0[s]
            st["sv_para"][:, s] = [mu[s], phi[s], sigma[s]]
        elif s < M:
            a0, b0 = cfg["factor_priorhomoskedastic"][s]
            v = (b0 + 0.5 * (e[s] ** 2).sum()) / rng.gamma(a0 + 0.5 * T)
            st["logvar"][:, s] = np.log(v)
    if R == 0:
        return
    restr = cfg["factor_restrinv"].astype(bool)
    # ---- STEP 2: normal-gamma prior on the loadings
    if cfg["factor_ngprior"]:
        sp = cfg["factor_shrinkagepriors"]
        units = R if cfg["factor_columnwise"] else M
        for u in range(units):
            mask = restr[:, u] if cfg["factor_columnwise"] else restr[u, :]
            tau = st["tau2"][:, u] if cfg["factor_columnwise"] else st["tau2"][u, :]
            lo = st["facload"][:, u] if cfg["factor_columnwise"] else st["facload"][u, :]
            a = sp["a"][u]
            lam2 = rng.gamma(sp["c"][u] + a * mask.sum()) / (sp["d"][u] + 0.5 * a * tau[mask].sum())
            st["lambda2"][u] = lam2
            tau[mask] = rgig(a - 0.5, lo[mask] ** 2, a * lam2, rng)
    # ---- STEP 2b: loadings, row by row
    w = np.exp(-st["logvar"][:, :M])            # T x M
    for i in range(M):
        idx = np.flatnonzero(restr[i])
        if idx.size == 0:
            continue
        F = st["fac"][idx].T                     # T x ri
        A = (F * w[:, [i]]).T @ F + np.diag(1.0 / st["tau2"][i, idx])
        b = (F * w[:, [i]]).T @ y[i]
        L = np.linalg.cholesky(A)
        draw = np.linalg.solve(L.T, np.linalg.solve(L, b) + rng.standard_normal(idx.size))
        tol = cfg["factor_facloadtol"]
        draw = np.where(np.abs(draw) < tol, np.where(draw < 0, -tol, tol), draw)
        st["facload"][i, idx] = draw
    # ---- STEP 2c: interweaving
    iw = min(int(cfg["factor_interweaving"]), 4)
    if iw:
        for j in range(R):
            col = st["facload"][:, j]
            lead = j if iw in (1, 2) else int(np.argmax(np.where(restr[:, j], np.abs(col), -1.0)))
            Lold = col[lead]
            if Lold == 0:
                continue
            m = restr[:, j]
            psis = ((col[m] / Lold) ** 2 / st["tau2"][m, j]).sum()
            nj = m.sum()
            hf = st["logvar"][:, M + j]
            if iw in (1, 3):   # shallow
                chi = (st["fac"][j] ** 2 * np.exp(-hf)).sum() * Lold ** 2
                x = float(rgig(0.5 * (nj - T), chi, psis, rng))
                scale = np.sqrt(x) / abs(Lold)
                st["facload"][:, j] *= scale
                st["fac"][j] /= scale
            elif het[M + j]:   # deep
                phi_j, sig_j = st["sv_para"][1, M + j], st["sv_para"][2, M + j]
                mu_old = np.log(Lold ** 2)
                hop = np.concatenate([[st["logvar0"][M + j]], hf]) + mu_old
                tmph = (hop[1:] - phi_j * hop[:-1]).sum()
                den = T + osv.B011INV
                gam_old = (1 - phi_j) * mu_old
                gam_prop = tmph / den + sig_j / np.sqrt(den) * rng.standard_normal()
                mu_prop = gam_prop / (1 - phi_j)
                h0v = cfg["sv_priorh0"][M + j]
                sd0 = sig_j / np.sqrt(1 - phi_j ** 2) if h0v <= 0 else sig_j * np.sqrt(h0v)
                lr = osv._lognorm(hop[0], mu_prop, sd0) - osv._lognorm(hop[0], mu_old, sd0)
                lr += 0.5 * nj * (mu_prop - mu_old) - 0.5 * psis * (np.exp(mu_prop) - np.exp(mu_old))
                lr += osv._lognorm(gam_old, 0, sig_j / np.sqrt(osv.B011INV)) - osv._lognorm(gam_prop, 0, sig_j / np.sqrt(osv.B011INV))
                if np.log(rng.uniform()) < lr:
                    scale = np.exp(0.5 * mu_prop) / Lold
                    st["facload"][:, j] *= scale
                    st["fac"][j] /= scale
                    st["logvar"][:, M + j] += mu_old - mu_prop
                    st["logvar0"][M + j] += mu_old - mu_prop
    # ---- STEP 3: factors, period by period
    Lam = st["facload"]
    hfac = st["logvar"][:, M:]
    for t in range(T):
        Lw = Lam * w[t][:, None]
        A = Lw.T @ Lam + np.diag(np.exp(-hfac[t]))
        b = Lw.T @ y[:, t]
        L = np.linalg.cholesky(A)
        st["fac"][:, t] = np.linalg.solve(L.T, np.linalg.solve(L, b) + rng.standard_normal(R))


def make_sv_prior(cfg, M, R):
    S = M + R
    ps2 = np.asarray(cfg["sv_priorsigma2"]).reshape(S, 2, order="F") if np.ndim(cfg["sv_priorsigma2"]) == 1 else np.asarray(cfg["sv_priorsigma2"])
    phi = np.asarray(cfg["sv_priorphi"], float)
    return osv.make_prior(S, mu_mean=cfg["sv_priormu"][0], mu_sd=cfg["sv_priormu"][1],
                          phi_a=np.concatenate([np.repeat(phi[0], M), np.repeat(phi[2], R)]),
                          phi_b=np.concatenate([np.repeat(phi[1], M), np.repeat(phi[3], R)]),
                          s2_rate=ps2[:, 1], h0_var=np.asarray(cfg["sv_priorh0"], float),
                          mu_const=np.concatenate([np.zeros(M, bool), np.ones(R, bool)]))
