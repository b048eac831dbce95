"""CPU: the look-ahead bookkeeping of the tiled K2 (oracle/tiled_cholesky.py restates which product is formed from which
finished tiles, k2_tiled.cu) reproduces the Cholesky factor of univariate_regression_update (sample_coefficients.cpp:58):
same L as numpy.linalg.cholesky to rounding error, for full and partial last blocks."""
import numpy as np
import pytest

from oracle.tiled_cholesky import tiled_cholesky


@pytest.mark.parametrize("n", [17, 32, 33, 64, 100, 161, 401])
def test_tiled_schedule_reproduces_cholesky(n):
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n + 40, n))
    A = X.T @ (X * rng.uniform(0.2, 3.0, n + 40)[:, None]) + np.diag(rng.uniform(0.01, 100.0, n))   # X'WX + diag(1/v)
    L = tiled_cholesky(A)
    ref = np.linalg.cholesky(A)
    assert np.abs(L - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.allclose(np.triu(L, 1), 0.0)
