"""Oracle: in-sample fit, covariance draws and log-variance forecasts (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/utilities_cpp.cpp: `predh` (:181-212), `predict_y` (:69-93), `insample` (:396-433) and
`vcov_cpp` (:436-472) with NumPy, taking the standard normals the reference draws with R::norm_rand as injected arrays.
These are SURVEY.md section 8(f) rank 2 - the next widening step after the hot path; the CUDA side is not built yet,
the restatement and its CPU tests (tests/test_oracle_fitted.py) are the groundwork it will be checked against.
`build_sigma` is shared with oracle/predict.py.
"""
from __future__ import annotations

import numpy as np

from oracle.predict import build_sigma


def _sigma_chol_upper(factor, facload, logvar_t, factors, m, u):
    """build_sigma(..., return_chol = true) (:45-67): upper-triangular R with R'R = Sigma."""
    if factor:
        return np.linalg.cholesky(build_sigma(True, facload, logvar_t, factors, m, None)).T
    Um = np.eye(m)
    iu = np.triu_indices(m, 1)
    order = np.lexsort((iu[0], iu[1]))
    Um[iu[0][order], iu[1][order]] = u
    return np.linalg.inv(Um) * np.exp(logvar_t / 2.0)[:, None]


def predh(logvar_T, ahead, each, sv_mu, sv_phi, sv_sigma, z):
    """:181-212.  logvar_T, sv_* (M, R); z[c, j] (M, R) normals of the c-th requested horizon, replicate j.
    Returns h_pred (len(ahead), M, R * each); replicate j fills slices j*R .. (j+1)*R - 1 (:205)."""
    ahead = np.asarray(ahead, int)
    M, R = logvar_T.shape
    out = np.zeros((ahead.size, M, R * each))
    sigma2 = np.zeros_like(sv_mu)
    counter = 0
    for f in range(int(ahead.max())):
        mu_pred = sv_mu + sv_phi ** (f + 1) * (logvar_T - sv_mu)
        sigma2 = sigma2 + (sv_sigma * sv_phi ** f) ** 2
        if counter < ahead.size and f == ahead[counter] - 1:
            for j in range(each):
                out[counter, :, j * R:(j + 1) * R] = mu_pred + z[counter, j] * np.sqrt(sigma2)
            counter += 1
    return out


def insample(X, PHI, U, facload, logvar, prediction, factor, z=None):
    """:396-433.  X (T, Kp); PHI (Kp, m, nsave); U (nU, nsave); facload (m, r, nsave); logvar (T, m[+r], nsave);
    z[t, r] (m,) normals in the reference's loop order (t outer, draw inner).  Returns y_pred (T, m, nsave)."""
    T = X.shape[0]
    Kp, m, nsave = PHI.shape
    y = np.zeros((T, m, nsave))
    for t in range(T):
        for r in range(nsave):
            yp = X[t] @ PHI[:, :, r]
            if prediction:
                lv = logvar[t, :, r]
                factors = lv.size - m
                Rch = _sigma_chol_upper(factor, facload[:, :, r] if factor else None, lv, factors, m,
                                        None if factor else U[:, r])
                yp = yp + z[t, r] @ Rch
            y[t, :, r] = yp
    return y


def vcov_cpp(factor, facload, logvar, U, M, factors):
    """:436-472.  logvar (T, M[+r], nsave).  Returns (T*M, M, nsave): slice i is the cube (T, M, M) of Sigma_t
    flattened column-major, element (t, a, b) at row t + T*a, column b (:448-449, :468)."""
    T, _, nsave = logvar.shape
    out = np.zeros((T * M, M, nsave))
    for i in range(nsave):
        arr = np.zeros((T, M, M))
        for t in range(T):
            arr[t] = build_sigma(factor, facload[:, :, i] if factor else None, logvar[t, :, i], factors, M,
                                 None if factor else U[:, i])
        out[:, :, i] = arr.reshape(T * M, M, order="F")
    return out
