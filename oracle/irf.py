"""Oracle: impulse responses per posterior draw (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/irf.cpp:457-524 (`shift_and_insert`, `irf_cpp`) with
NumPy, including the two quirks SURVEY.md flags: Q1 - factor shocks are scaled by
exp(logvar_t[0:n_factors]/2), i.e. the first idiosyncratic log-variances
(irf.cpp:495); Q2 - `shift_and_insert` starts at column n_cols-2 (irf.cpp:461), so
the last column of the predictor row is never refreshed.
Output layout = what R/irf.R:275-276 builds: (variables, shocks, 1+ahead, draws).
"""
from __future__ import annotations

import numpy as np


def shift_and_insert(Xrow, new_y):
    """irf.cpp:457-465."""
    n_cols, m = Xrow.shape[1], new_y.shape[1]
    i = n_cols - 2
    while m <= i:
        Xrow[:, i] = Xrow[:, i - m]
        i -= 1
    Xrow[:, :m] = new_y


def irf_cpp(coefficients, factor_loadings, U_vecs, logvar_t, shocks, ahead, rotation=None):
    """irf.cpp:468-524.  coefficients (Kp, M, R); factor_loadings (M, r, R) or empty;
    U_vecs (M(M-1)/2, R) or empty; logvar_t (., R); shocks (dim, n_shocks); rotation (dim, dim, R)|None."""
    Kp, M, R = coefficients.shape
    n_shocks = shocks.shape[1]
    is_factor = factor_loadings is not None and factor_loadings.ndim == 3 and factor_loadings.shape[1] > 0
    is_chol = (not is_factor) and U_vecs is not None and U_vecs.ndim == 2 and U_vecs.shape[1] > 0
    iu = np.triu_indices(M, 1)
    order = np.lexsort((iu[0], iu[1]))
    ri, ci = iu[0][order], iu[1][order]
    out = np.zeros((M, n_shocks, ahead + 1, R))
    for r in range(R):
        rs = rotation[:, :, r] @ shocks if rotation is not None else shocks.copy()
        if is_factor:
            nf = factor_loadings.shape[1]
            rs = rs * np.exp(logvar_t[:nf, r] / 2.0)[:, None]          # :495 (quirk Q1)
            ir0 = factor_loadings[:, :, r] @ rs
        elif is_chol:
            U = np.eye(M)
            U[ri, ci] = U_vecs[:, r]
            rs = rs * np.exp(logvar_t[:, r] / 2.0)[:, None]
            ir0 = np.linalg.solve(np.tril(U.T), rs)                     # :505
        else:
            ir0 = rs
        out[:, :, 0, r] = ir0
        pred = np.zeros((n_shocks, Kp))
        pred[:, :M] = ir0.T
        for t in range(1, ahead + 1):
            new = pred @ coefficients[:, :, r]
            out[:, :, t, r] = new.T
            shift_and_insert(pred, new)
    return out
