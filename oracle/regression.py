"""Oracle: equation-wise conditional coefficient draws (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/sample_coefficients.cpp:27-188 with *injected*
standard normals (the reference draws them from R's RNG via ``R::norm_rand``;
that stream cannot be reproduced, so every function takes the normals as an
argument, consumed in the reference's order).

All matrices are float64; column-major (Fortran) order is used where it matters
for LAPACK so the same routines RcppArmadillo dispatches to are hit:
``X.t()*X`` -> dsyrk, ``chol(.,"upper")`` -> dpotrf('U'), ``inv(trimatu(.))``
-> dtrtri('U'), ``solve`` -> dgesv.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import blas, lapack


def _syrk_t(A: np.ndarray) -> np.ndarray:
    """A.t()*A the way Armadillo does it: dsyrk (upper) + symmetrise."""
    A = np.asfortranarray(A)
    if A.shape[1] == 0:
        return np.zeros((0, 0))
    C = blas.dsyrk(1.0, A, trans=1, lower=0)
    iu = np.triu_indices(C.shape[0], 1)
    C[(iu[1], iu[0])] = C[iu]
    return C


def univariate_regression_update(prior_mean, prior_variance, y_norm, X_norm,
                                 typeU: bool, huge: bool, z, delta=None,
                                 out_len: int | None = None):
    """sample_coefficients.cpp:27-81.

    ``z``: K standard normals (`rnormvec`, :67-68; in the huge branch the `u`
    normals, :42).  ``delta``: T standard normals (huge branch only, :47).
    Returns the draw (length K, or ``out_len`` zero-padded with a 1 at index K
    when ``typeU``, :73-77).
    """
    prior_variance = np.asarray(prior_variance, dtype=np.float64)
    K = prior_variance.size
    T = y_norm.size
    z = np.asarray(z, dtype=np.float64)
    assert z.size == K
    if huge:
        # :40-52  Bhattacharya, Chakraborty & Mallick (2016)
        u = z * np.sqrt(prior_variance)
        if not typeU:
            u = u + prior_mean
        assert delta is not None and delta.size == T
        v = X_norm @ u + delta
        X_pv = X_norm * prior_variance[None, :]
        A = X_pv @ X_norm.T + np.eye(T)
        _, _, w, info = lapack.dgesv(np.asfortranarray(A), (y_norm - v).reshape(-1, 1))
        if info != 0:
            raise RuntimeError("solve(): solution not found")
        draw = u + X_pv.T @ w[:, 0]
    else:
        # :54-70
        Omega = _syrk_t(X_norm)
        Omega[np.diag_indices(K)] += 1.0 / prior_variance
        R, info = lapack.dpotrf(np.asfortranarray(Omega), lower=0, clean=1)
        if info != 0:
            raise RuntimeError("chol(): decomposition failed")
        Rinv, info = lapack.dtrtri(R, lower=0)
        if info != 0:
            raise RuntimeError("inv(): matrix is singular")
        Rinv = np.triu(Rinv)
        coef_post = X_norm.T @ y_norm
        if not typeU:
            coef_post = coef_post + prior_mean / prior_variance
        # Armadillo evaluates A*B*c as A*(B*c)
        coef_post = Rinv @ (Rinv.T @ coef_post)
        draw = coef_post + Rinv @ z
    if typeU:
        out = np.zeros(out_len)
        out[:K] = draw
        out[K] = 1.0
        return out
    return draw


def sample_PHI_factor(PHI, PHI_prior, Y, X, logvaridi, V_prior, facload, fac,
                      huge: bool, z, delta=None):
    """sample_coefficients.cpp:84-114.

    ``z``: (Kp, M) normals - column j is consumed by equation j (the reference
    draws Kp normals per equation, equations in order).  ``delta``: (T, M), huge
    branch only (drawn after the Kp normals of the same equation).
    Returns the new PHI (Kp, M); the input is not modified.
    """
    M = Y.shape[1]
    r = fac.shape[0] if fac is not None and fac.size else 0
    Y_hat = Y - (facload @ fac).T if r > 0 else Y          # :93
    normalizer = np.exp(-logvaridi / 2.0)                   # :95
    out = np.array(PHI, dtype=np.float64, order="F", copy=True)
    for j in range(M):                                       # :97
        X_norm = X * normalizer[:, [j]]                      # :99
        y_norm = Y_hat[:, j] * normalizer[:, j]              # :100
        out[:, j] = univariate_regression_update(
            PHI_prior[:, j], V_prior[:, j], y_norm, X_norm, False, huge,
            z[:, j], None if delta is None else delta[:, j])
    return out


def sample_PHI(PHI, PHI_prior, Y, X, U, d_sqrt, V_prior, z):
    """sample_coefficients.cpp:118-146 - corrected triangular algorithm, as
    written (materialises the T(M-j) x Kp Kronecker design).  ``z``: (Kp, M)."""
    M = Y.shape[1]
    out = np.array(PHI, dtype=np.float64, order="F", copy=True)
    for j in range(M):                                       # :122
        PHI_0 = out.copy()                                   # :124
        PHI_0[:, j] = 0.0                                    # :125
        normalizer = d_sqrt[:, j:].reshape(-1, order="F")    # :126-127
        Y_new = ((Y - X @ PHI_0) @ U[:, j:]).reshape(-1, order="F") / normalizer  # :128-130
        X_new = np.kron(U[j:j + 1, j:].T, X) / normalizer[:, None]                # :131-132
        out[:, j] = univariate_regression_update(
            PHI_prior[:, j], V_prior[:, j], Y_new, X_new, False, False, z[:, j])
    return out


def sample_PHI_collapsed(PHI, PHI_prior, Y, X, U, d_sqrt, V_prior, z):
    """Algebraically collapsed form of ``sample_PHI``:
    X_new'X_new = X' diag(w_j) X with w_j[t] = sum_{i>=j} U[j,i]^2 / d_sqrt[t,i]^2.
    Used only to cross-check the identity the CUDA path relies on."""
    M = Y.shape[1]
    out = np.array(PHI, dtype=np.float64, order="F", copy=True)
    for j in range(M):
        PHI_0 = out.copy()
        PHI_0[:, j] = 0.0
        E = (Y - X @ PHI_0) @ U[:, j:]
        d2 = d_sqrt[:, j:] ** 2
        w = (U[j, j:] ** 2 / d2).sum(axis=1)
        s = (U[j, j:] * E / d2).sum(axis=1)
        Omega = (X * w[:, None]).T @ X
        Omega[np.diag_indices_from(Omega)] += 1.0 / V_prior[:, j]
        L = np.linalg.cholesky(Omega)
        b = X.T @ s + PHI_prior[:, j] / V_prior[:, j]
        y = np.linalg.solve(L, b)
        out[:, j] = np.linalg.solve(L.T, y + z[:, j])
    return out


def sample_U(U, Ytilde, V_i, d_sqrt, z):
    """sample_coefficients.cpp:160-188.  ``z``: flat array of M(M-1)/2 normals in
    the reference's consumption order (equation i=1..M-1 takes i normals).
    Returns the new U (M, M), upper unitriangular."""
    M = Ytilde.shape[1]
    out = np.array(U, dtype=np.float64, order="F", copy=True)
    ind = 0
    for i in range(1, M):                                    # :166
        Z = -1.0 * (Ytilde[:, :i] / d_sqrt[:, [i]])          # :169-170
        c = Ytilde[:, i] / d_sqrt[:, i]                      # :172
        out[:, i] = univariate_regression_update(
            None, V_i[ind:ind + i], c, Z, True, False, z[ind:ind + i], out_len=M)
        ind += i
    return out
