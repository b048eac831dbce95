"""Synthetic inputs for the hot path, following the reference's own benchmark
protocol (vignettes/scalability-vignette.qmd:88-125): iid N(0,1) data,
X = embed(Yraw, p+1)[, -(1:M)] plus a ones column (R/bvar_wrapper.R:228-235)."""
from __future__ import annotations

import numpy as np


def embed_lags(Yraw: np.ndarray, p: int, intercept: bool = True):
    """R's embed(Yraw, p+1)[, -(1:M)] (+ ones): row t = [y_{t-1}, ..., y_{t-p}, 1]."""
    Traw, M = Yraw.shape
    T = Traw - p
    cols = [Yraw[p - l:Traw - l, :] for l in range(1, p + 1)]
    X = np.concatenate(cols, axis=1) if p > 0 else np.zeros((T, 0))
    if intercept:
        X = np.concatenate([X, np.ones((T, 1))], axis=1)
    Y = Yraw[p:, :]
    return np.asfortranarray(Y), np.asfortranarray(X)


def phi_draw_inputs(M, p, T, r=1, seed=123, prior="dl"):
    """Inputs of one sample_PHI_factor call at a plausible sampler state."""
    rng = np.random.default_rng(seed)
    Yraw = rng.standard_normal((T + p, M))
    Y, X = embed_lags(Yraw, p)
    Kp = X.shape[1]
    logvar = -1.0 + 0.5 * rng.standard_normal((T, M))
    if prior == "dl":  # heavy-tailed prior variances as under DL/R2D2
        V = np.exp(rng.normal(-6.0, 4.0, size=(Kp, M)))
    else:
        V = np.full((Kp, M), 10.0)
    V[-1, :] = 100.0  # intercept: prior sd 10 (R/bvar_wrapper.R:183,279)
    PHI0 = np.zeros((Kp, M))
    facload = 0.5 * rng.standard_normal((M, r)) if r else None
    fac = rng.standard_normal((r, T)) if r else None
    z = rng.standard_normal((Kp, M))
    return dict(Y=Y, X=X, logvar=np.asfortranarray(logvar), V=np.asfortranarray(V),
                PHI0=np.asfortranarray(PHI0), facload=facload, fac=fac, z=np.asfortranarray(z))
