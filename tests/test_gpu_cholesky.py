"""GPU parity of the cholesky-structure entry points through the C ABI:
sample_PHI_cholesky (triangular algorithm, collapsed form) against the oracle's
AS-WRITTEN Kronecker restatement, sample_U against the oracle, and the reference's
own Geweke joint-distribution test (tests/testthat/test-bvar.R:21-92).
Tolerance: 1e-9 relative per equation (north_star), metric as in test_gpu_phi_factor."""
import numpy as np
import pytest

from oracle import regression as orc
from tests.stat_helpers import geweke_phi_cholesky

pytestmark = pytest.mark.gpu
TOL = 1e-9


def per_eq_relerr(a, b):
    return (np.max(np.abs(a - b), axis=0) / np.maximum(np.max(np.abs(b), axis=0), 1e-300)).max()


def chol_inputs(M, Kp, T, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, Kp)); X[:, -1] = 1.0
    Y = rng.standard_normal((T, M))
    U = np.triu(0.3 * rng.standard_normal((M, M)), 1) + np.eye(M)
    d_sqrt = np.exp(rng.normal(0, 0.5, (T, M)))
    V = np.exp(rng.normal(-2, 3, (Kp, M)))
    PHI0 = 0.1 * rng.standard_normal((Kp, M))
    PHI = rng.standard_normal((Kp, M))
    z = rng.standard_normal((Kp, M))
    return X, Y, U, d_sqrt, V, PHI0, PHI, z


@pytest.mark.parametrize("M,Kp,T", [(1, 3, 20), (3, 4, 30), (5, 10, 50), (6, 40, 60), (14, 43, 100), (8, 70, 90)])
def test_phi_cholesky_matches_kronecker_oracle(handle, M, Kp, T):
    X, Y, U, d_sqrt, V, PHI0, PHI, z = chol_inputs(M, Kp, T, 100 + M)
    ref = orc.sample_PHI(PHI, PHI0, Y, X, U, d_sqrt, V, z)
    out = handle.sample_PHI_cholesky(PHI, PHI0, Y, X, U, d_sqrt, V, z=z)
    assert per_eq_relerr(out, ref) < TOL
    # the input PHI is not modified (the reference takes it by value)
    assert not np.shares_memory(out, PHI)


def test_phi_cholesky_config4_shape_spotcheck(handle):
    """M=50, Kp=101 (two lags + intercept) - big enough for several column blocks; the as-written
    oracle needs the 50x larger Kronecker design, so compare against the collapsed oracle
    (itself pinned to the Kronecker form in tests/test_oracle_regression.py)."""
    X, Y, U, d_sqrt, V, PHI0, PHI, z = chol_inputs(50, 101, 200, 7)
    ref = orc.sample_PHI_collapsed(PHI, PHI0, Y, X, U, d_sqrt, V, z)
    out = handle.sample_PHI_cholesky(PHI, PHI0, Y, X, U, d_sqrt, V, z=z)
    assert per_eq_relerr(out, ref) < TOL


@pytest.mark.parametrize("M,T", [(1, 10), (2, 20), (5, 50), (14, 100), (40, 120)])
def test_sample_U_matches_oracle(handle, M, T):
    rng = np.random.default_rng(M)
    E = rng.standard_normal((T, M)); d_sqrt = np.exp(rng.normal(0, 0.3, (T, M)))
    nU = M * (M - 1) // 2
    V = np.exp(rng.normal(0, 2, nU)); z = rng.standard_normal(nU)
    ref = orc.sample_U(np.eye(M), E, V, d_sqrt, z)
    out = handle.sample_U(E, V, d_sqrt, z=z)
    assert np.allclose(np.diag(out), 1.0) and np.all(np.tril(out, -1) == 0)
    if M > 1:
        iu = np.triu_indices(M, 1)
        assert np.max(np.abs(out[iu] - ref[iu])) / np.max(np.abs(ref[iu])) < TOL


def test_geweke_sample_phi_cholesky_gpu(handle):
    """tests/testthat/test-bvar.R:21-92 with bayesianVARs:::sample_PHI_cholesky replaced by the C ABI."""
    rng = np.random.default_rng(123)
    p1, p2 = geweke_phi_cholesky(lambda *a: handle.sample_PHI_cholesky(*a[:-1], z=a[-1]), 3000, rng)
    assert p1 > 0.01 and p2 > 0.01
