"""Oracle: vectorised GIG sampler in NumPy (TEST INFRASTRUCTURE ONLY).

The reference draws GIG variates through the third-party GIGrvg package
(src/sample_coefficients.cpp:16-22; source not under /root/reference) - PARITY
UNPINNED.  This is an independent NumPy restatement of the algorithm GIGrvg
publishes (Hoermann & Leydold 2014): ratio-of-uniforms with / without mode shift
and the three-part hat for 0 <= lambda < 1 with small omega, plus the gamma /
inverse-gamma limits.  density: x^(lambda-1) exp(-(chi/x + psi x)/2).
"""
from __future__ import annotations

import numpy as np

ZTOL = 2.220446049250313e-16 * 10.0


def _mode(lam, om):
    return np.where(lam >= 1.0,
                    (np.sqrt((lam - 1.0) ** 2 + om ** 2) + (lam - 1.0)) / om,
                    om / (np.sqrt((1.0 - lam) ** 2 + om ** 2) + (1.0 - lam)))


def _rejection(n, propose, rng, maxit=10000):
    """propose(idx) -> (X, accept_mask) for the still-open indices."""
    out = np.empty(n)
    todo = np.arange(n)
    for _ in range(maxit):
        if todo.size == 0:
            break
        X, acc = propose(todo)
        out[todo[acc]] = X[acc]
        todo = todo[~acc]
    if todo.size:
        raise RuntimeError("GIG rejection sampler did not terminate")
    return out


def _rou_noshift(lam, om, rng):
    t, s = 0.5 * (lam - 1.0), 0.25 * om
    xm = _mode(lam, om)
    nc = t * np.log(xm) - s * (xm + 1.0 / xm)
    ym = ((lam + 1.0) + np.sqrt((lam + 1.0) ** 2 + om ** 2)) / om
    um = np.exp(0.5 * (lam + 1.0) * np.log(ym) - s * (ym + 1.0 / ym) - nc)

    def prop(i):
        U = um[i] * rng.uniform(size=i.size)
        V = rng.uniform(size=i.size)
        X = U / V
        return X, np.log(V) <= t[i] * np.log(X) - s[i] * (X + 1.0 / X) - nc[i]
    return _rejection(lam.size, prop, rng)


def _rou_shift(lam, om, rng):
    t, s = 0.5 * (lam - 1.0), 0.25 * om
    xm = _mode(lam, om)
    nc = t * np.log(xm) - s * (xm + 1.0 / xm)
    a = -(2.0 * (lam + 1.0) / om + xm)
    b = 2.0 * (lam - 1.0) * xm / om - 1.0
    c = xm
    p = b - a * a / 3.0
    q = 2.0 * a ** 3 / 27.0 - a * b / 3.0 + c
    fi = np.arccos(np.clip(-q / (2.0 * np.sqrt(-(p ** 3) / 27.0)), -1.0, 1.0))
    fak = 2.0 * np.sqrt(-p / 3.0)
    y1 = fak * np.cos(fi / 3.0) - a / 3.0
    y2 = fak * np.cos(fi / 3.0 + 4.0 / 3.0 * np.pi) - a / 3.0
    up = (y1 - xm) * np.exp(t * np.log(y1) - s * (y1 + 1.0 / y1) - nc)
    um = (y2 - xm) * np.exp(t * np.log(y2) - s * (y2 + 1.0 / y2) - nc)

    def prop(i):
        U = um[i] + rng.uniform(size=i.size) * (up[i] - um[i])
        V = rng.uniform(size=i.size)
        X = U / V + xm[i]
        ok = X > 0
        Xs = np.where(ok, X, 1.0)
        return Xs, ok & (np.log(V) <= t[i] * np.log(Xs) - s[i] * (Xs + 1.0 / Xs) - nc[i])
    return _rejection(lam.size, prop, rng)


def _newapproach(lam, om, rng):
    xm = _mode(lam, om)
    x0 = om / (1.0 - lam)
    k0 = np.exp((lam - 1.0) * np.log(xm) - 0.5 * om * (xm + 1.0 / xm))
    A0 = k0 * x0
    far = x0 >= 2.0 / om
    lam_s = np.where(lam == 0.0, 1.0, lam)
    k1 = np.where(far, 0.0, np.exp(-om))
    A1 = np.where(far, 0.0, np.where(lam == 0.0, k1 * np.log(2.0 / om ** 2),
                                     k1 / lam_s * ((2.0 / om) ** lam - x0 ** lam)))
    k2 = np.where(far, x0 ** (lam - 1.0), (2.0 / om) ** (lam - 1.0))
    A2 = np.where(far, k2 * 2.0 * np.exp(-om * x0 / 2.0) / om, k2 * 2.0 * np.exp(-1.0) / om)
    Atot = A0 + A1 + A2

    def prop(i):
        V = Atot[i] * rng.uniform(size=i.size)
        l, o = lam[i], om[i]
        in0 = V <= A0[i]
        in1 = ~in0 & (V - A0[i] <= A1[i])
        in2 = ~in0 & ~in1
        X = np.empty(i.size); hx = np.empty(i.size)
        X[in0] = x0[i][in0] * V[in0] / A0[i][in0]
        hx[in0] = k0[i][in0]
        V1 = V - A0[i]
        z = in1 & (l == 0.0)
        X[z] = o[z] * np.exp(np.exp(o[z]) * V1[z])
        hx[z] = k1[i][z] / X[z]
        nz = in1 & (l != 0.0)
        X[nz] = (x0[i][nz] ** l[nz] + l[nz] / k1[i][nz] * V1[nz]) ** (1.0 / l[nz])
        hx[nz] = k1[i][nz] * X[nz] ** (l[nz] - 1.0)
        V2 = V1 - A1[i]
        a = np.maximum(x0[i], 2.0 / o)
        X[in2] = -2.0 / o[in2] * np.log(np.exp(-o[in2] / 2.0 * a[in2]) - o[in2] / (2.0 * k2[i][in2]) * V2[in2])
        hx[in2] = k2[i][in2] * np.exp(-o[in2] / 2.0 * X[in2])
        U = rng.uniform(size=i.size) * hx
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.log(U) <= (l - 1.0) * np.log(X) - o / 2.0 * (X + 1.0 / X)
        return X, acc & np.isfinite(X) & (X > 0)
    return _rejection(lam.size, prop, rng)


def rgig(lam, chi, psi, rng):
    """Element-wise GIG(lam, chi, psi) draws (broadcasting)."""
    lam, chi, psi = np.broadcast_arrays(np.asarray(lam, float), np.asarray(chi, float), np.asarray(psi, float))
    shape = lam.shape
    lam, chi, psi = lam.ravel().copy(), chi.ravel().copy(), psi.ravel().copy()
    out = np.empty(lam.size)
    small_chi = chi < ZTOL
    small_psi = ~small_chi & (psi < ZTOL)
    i = np.flatnonzero(small_chi)
    if i.size:
        pos = lam[i] > 0
        g = rng.gamma(np.abs(lam[i])) * (2.0 / psi[i])
        out[i] = np.where(pos, g, 1.0 / g)
    i = np.flatnonzero(small_psi)
    if i.size:
        out[i] = 1.0 / (rng.gamma(np.abs(lam[i])) * (2.0 / chi[i]))
    i = np.flatnonzero(~small_chi & ~small_psi)
    if i.size:
        l = np.abs(lam[i])
        alpha, om = np.sqrt(chi[i] / psi[i]), np.sqrt(chi[i] * psi[i])
        X = np.empty(i.size)
        m1 = (l > 2.0) | (om > 3.0)
        m2 = ~m1 & ((l >= 1.0 - 2.25 * om ** 2) | (om > 0.2))
        m3 = ~m1 & ~m2
        if m1.any():
            X[m1] = _rou_shift(l[m1], om[m1], rng)
        if m2.any():
            X[m2] = _rou_noshift(l[m2], om[m2], rng)
        if m3.any():
            X[m3] = _newapproach(l[m3], om[m3], rng)
        out[i] = np.where(lam[i] < 0, alpha / X, alpha * X)
    return out.reshape(shape)
