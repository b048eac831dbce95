"""Oracle for the SCHEDULE of the tiled K2 (bayesianvars_b200/csrc/k2_tiled.cu) - TEST INFRASTRUCTURE ONLY.

The tiled kernel factorises Omega = L L' (univariate_regression_update, src/sample_coefficients.cpp:58-70: chol, then the
triangular solves) block column by block column with 32 x 32 tiles, but the tile below the diagonal and the diagonal tile
itself never walk their own history: a far-tile task forms two look-ahead products on the side,

    P(I, I-1) = sum_{b <= I-3} L(I, b) L(I-1, b)'          (the old history of the chain's tile (I, I-1))
    D(I)      = - sum_{b <= I-2} L(I, b) L(I, b)'          (the history of the diagonal tile I, all but the newest term)

so that the chain of an equation only adds the NEWEST block column:

    diag_k   = A(k,k) + D(k) - L(k,k-1) L(k,k-1)'                        -> Cholesky -> Linv_k
    C(k+1,k) = A(k+1,k) - P(k+1,k) - L(k+1,k-1) L(k,k-1)',   L(k+1,k) = C Linv_k'
    far tile (I >= k+2):  L(I,k) = (A(I,k) - sum_{b<k} L(I,b) L(k,b)') Linv_k'

This module replays exactly that dependency structure in NumPy (which product is formed when, from which finished tiles)
and returns L; tests/test_oracle_tiled_schedule.py compares it with numpy.linalg.cholesky.  It follows the kernel's
bookkeeping (P published by the task of block column k for row block k+2, D for row block k+2 likewise), not its
parallelism."""
from __future__ import annotations

import numpy as np

NB = 32


def tiled_cholesky(A: np.ndarray, nb: int = NB) -> np.ndarray:
    n = A.shape[0]
    nbk = (n + nb - 1) // nb
    blk = lambda i: slice(i * nb, min((i + 1) * nb, n))
    L = np.zeros_like(A)
    P = {}          # P[I]: old history of tile (I, I-1), block columns <= I-3
    D = {}          # D[I]: minus the history of diagonal tile I, block columns <= I-2
    Linv = {}
    for k in range(nbk):
        # ---- chain step k
        dk = A[blk(k), blk(k)].copy()
        if k >= 2:
            dk += D[k]
        if k >= 1:
            dk -= L[blk(k), blk(k - 1)] @ L[blk(k), blk(k - 1)].T
        Lkk = np.linalg.cholesky(dk)
        L[blk(k), blk(k)] = Lkk
        Linv[k] = np.linalg.inv(Lkk)
        if k + 1 < nbk:
            C = A[blk(k + 1), blk(k)].copy()
            if k >= 2:
                C -= P[k + 1]
            if k >= 1:
                C -= L[blk(k + 1), blk(k - 1)] @ L[blk(k), blk(k - 1)].T
            L[blk(k + 1), blk(k)] = C @ Linv[k].T
        # ---- far tiles of block column k (row blocks >= k+2); the nearest one also forms the look-ahead products
        for I in range(k + 2, nbk):
            hist = np.zeros((blk(I).stop - blk(I).start, blk(k).stop - blk(k).start))
            for b in range(k):
                hist += L[blk(I), blk(b)] @ L[blk(k), blk(b)].T
            if I == k + 2:
                # P(k+2, k+1): block columns < k, from the rows this task accumulates anyway
                Pn = np.zeros((blk(I).stop - blk(I).start, blk(k + 1).stop - blk(k + 1).start))
                Dn = np.zeros((blk(I).stop - blk(I).start,) * 2)
                for b in range(k):
                    Pn += L[blk(I), blk(b)] @ L[blk(k + 1), blk(b)].T
                    Dn += L[blk(I), blk(b)] @ L[blk(I), blk(b)].T
                P[I] = Pn
            L[blk(I), blk(k)] = (A[blk(I), blk(k)] - hist) @ Linv[k].T
            if I == k + 2:
                # ... and with its own block column added, D(k+2) is complete (written by this one task)
                D[I] = -(Dn + L[blk(I), blk(k)] @ L[blk(I), blk(k)].T)
    return L
