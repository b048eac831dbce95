"""Oracle: predictive simulation / log predictive likelihoods per stored draw
(TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/utilities_cpp.cpp: `dmvnrm_arma_fast` (:19-42),
`build_sigma` (:45-67), `Sigma_pred_uncond` (:144-178) and `out_of_sample`
(:216-393) literally with NumPy, taking the standard normals the reference draws
with R::norm_rand as injected arrays:
  z_logvar[r, i, j, :]  - the |sv_indicator| normals of horizon step i, replicate j (:325-327)
  z_pred[r, c, j, :]    - the M normals of the predictive draw at the c-th requested horizon (:371-374)
Slice index of replicate j of draw r is r + j*R (:352, :364, :376).
The predictive mean is computed as in the reference through PHI_power0 (:125-142).
"""
from __future__ import annotations

import numpy as np

LOG2PI = np.log(2.0 * np.pi)


def dmvnrm_log(x, mean, sigma):
    """:19-42 with logd = true, one row."""
    R = np.linalg.cholesky(sigma).T
    rooti = np.linalg.inv(R)
    z = (x - mean) @ rooti
    return np.log(np.diag(rooti)).sum() - 0.5 * x.size * LOG2PI - 0.5 * z @ z


def build_sigma(factor, facload, logvar_t, factors, m, u):
    """:45-67 with return_chol = false."""
    if factor:
        S = facload * np.exp(logvar_t[m:m + factors] / 2.0)[None, :]
        Sigma = S @ S.T
        Sigma[np.diag_indices(m)] += np.exp(logvar_t[:m])
        return Sigma
    Um = np.eye(m)
    iu = np.triu_indices(m, 1)
    order = np.lexsort((iu[0], iu[1]))
    Um[iu[0][order], iu[1][order]] = u
    Sc = np.linalg.inv(Um) * np.exp(logvar_t / 2.0)[:, None]
    return Sc.T @ Sc


def PHI_power0(PHI_power, PHI):
    """:125-142 (in place)."""
    M = PHI.shape[1]
    Kp = PHI.shape[0]
    p = Kp // M
    K = p * M
    tmp = PHI_power.copy()
    for i in range(p):
        PHI_power[i * M:(i + 1) * M] = PHI[i * M:(i + 1) * M] @ tmp[:M]
        if i < p - 1:
            PHI_power[i * M:(i + 1) * M] += tmp[(i + 1) * M:(i + 2) * M]
    if Kp > K:
        PHI_power[Kp - 1] = PHI[Kp - 1] @ tmp[:M] + tmp[Kp - 1]


def Sigma_pred_uncond(SL, n_ahead, PHI, Sigma, M, p, Kp):
    """:144-178 (in place on SL)."""
    if p > 1:
        if n_ahead > 0:
            tmp = np.zeros_like(SL)
            mn = min(n_ahead, p)
            for j in range(mn):
                tmp[:M, j * M:(j + 1) * M] = PHI[:mn * M].T @ SL[:mn * M, j * M:(j + 1) * M]
            tmp[M:p * M] = SL[:(p - 1) * M]
            SL[:M, :M] = tmp[:M, :mn * M] @ PHI[:mn * M]
            SL[:, M:p * M] = tmp[:, :(p - 1) * M]
            il = np.tril_indices(Kp, -1)
            SL[il] = SL.T[il]
    elif p == 1:
        if n_ahead > 0:
            SL[:M, :M] = PHI[:p * M].T @ SL[:M, :M] @ PHI[:p * M]
    SL[:M, :M] += Sigma


def out_of_sample(each, X_T_plus_1, PHI, U, facload, logvar_T, ahead, sv_mu, sv_phi, sv_sigma, sv_indicator, factor,
                  LPL, Y_obs, LPL_subset, VoI, simulate_predictive, z_logvar, z_pred):
    """:216-393.  PHI (Kp, M, R); U (nU, R); facload (M, r, R); logvar_T (M+r, R); sv_* (n_sv, R)."""
    Kp, M, R = PHI.shape
    nsave = R * each
    ahead = np.asarray(ahead, int)
    na, max_ahead = ahead.size, int(ahead.max())
    lags = Kp // M
    anySV = np.size(sv_indicator) > 0
    factors = logvar_T.shape[0] - M
    pred = np.zeros((na, M, nsave)) if simulate_predictive else np.zeros((0, 0, 0))
    LPLd = np.zeros((na, nsave)) if LPL else np.zeros((0, 0))
    PLu = np.zeros((na, M, nsave)) if LPL else np.zeros((0, 0, 0))
    LPLs = np.zeros((na, nsave)) if LPL_subset else np.zeros((0, 0))
    for r in range(R):
        Sigma_comp = np.zeros((Kp, Kp, each))
        lv_var = np.zeros(np.size(sv_indicator))
        PHI_power = PHI[:, :, r].copy()
        counter = 0
        for i in range(max_ahead):
            if anySV:
                lv_mean = sv_mu[:, r] + sv_phi[:, r] ** (i + 1) * (logvar_T[sv_indicator, r] - sv_mu[:, r])
                lv_var = lv_var + (sv_sigma[:, r] * sv_phi[:, r] ** i) ** 2
            if i > 0:
                PHI_power0(PHI_power, PHI[:, :, r])
            lv_pred = logvar_T[:, r].copy()
            for j in range(each):
                if anySV:
                    lv_pred[sv_indicator] = lv_mean + z_logvar[r, i, j] * np.sqrt(lv_var)
                Sigma = build_sigma(factor, facload[:, :, r] if factor else None, lv_pred, factors, M,
                                    None if factor else U[:, r])
                SL = Sigma_comp[:, :, j].copy()
                Sigma_pred_uncond(SL, i, PHI[:, :, r], Sigma, M, lags, Kp)
                Sigma_comp[:, :, j] = SL
                if counter < na and i == ahead[counter] - 1:
                    y_pred = X_T_plus_1 @ PHI_power
                    sl = r + j * R
                    if LPL:
                        LPLd[counter, sl] = dmvnrm_log(Y_obs[counter], y_pred, SL[:M, :M])
                        if LPL_subset:
                            LPLs[counter, sl] = dmvnrm_log(Y_obs[counter, VoI], y_pred[VoI], SL[np.ix_(VoI, VoI)])
                        sd = np.sqrt(np.diag(SL)[:M])
                        PLu[counter, :, sl] = np.exp(-0.5 * ((Y_obs[counter] - y_pred) / sd) ** 2) / (sd * np.sqrt(2 * np.pi))
                    if simulate_predictive:
                        Rch = np.linalg.cholesky(SL[:M, :M]).T
                        pred[counter, :, sl] = y_pred + z_pred[r, counter, j] @ Rch
                    if j == each - 1:
                        counter += 1
    return dict(LPL_draws=LPLd, PL_univariate_draws=PLu, LPL_sub_draws=LPLs, predictions=pred)
