"""GPU parity of irf_cpp (src/irf.cpp:468-524) through the C ABI against the oracle
(pure recursion in FP64: tolerance 1e-12 relative to the largest response), plus
the structural checks of tests/testthat/test-irf.R:92 (output dimensions)."""
import numpy as np
import pytest

from oracle.irf import irf_cpp as oracle_irf

pytestmark = pytest.mark.gpu


def draws(M, p, R, nf, seed, intercept=True):
    rng = np.random.default_rng(seed)
    Kp = M * p + (1 if intercept else 0)
    coef = 0.3 * rng.standard_normal((Kp, M, R)) / np.sqrt(M)
    fl = rng.standard_normal((M, nf, R)) if nf else np.zeros((0, 0, 0))
    U = rng.standard_normal((M * (M - 1) // 2, R)) * 0.3
    lv = rng.normal(-1, 0.5, (M + nf, R))
    return coef, fl, U, lv


@pytest.mark.parametrize("M,p,R,nf,ahead", [(5, 5, 12, 2, 8), (3, 1, 7, 1, 4), (10, 2, 30, 3, 12), (4, 2, 5, 0, 6), (20, 2, 9, 0, 5)])
def test_irf_matches_oracle(handle, M, p, R, nf, ahead):
    coef, fl, U, lv = draws(M, p, R, nf, 10 + M)
    if nf:
        shocks = np.eye(nf)
        ref = oracle_irf(coef, fl, None, lv, shocks, ahead)
        out = handle.irf_cpp(coef, fl, None, lv, shocks, ahead)
    else:
        shocks = np.eye(M)
        ref = oracle_irf(coef, None, U, lv[:M], shocks, ahead)
        out = handle.irf_cpp(coef, None, U, lv[:M], shocks, ahead)
    assert out.shape == (M, shocks.shape[1], ahead + 1, R)          # test-irf.R:92
    assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_irf_rotation_custom_shocks_and_no_intercept(handle):
    M, p, R, nf = 6, 3, 11, 3
    coef, fl, U, lv = draws(M, p, R, nf, 3, intercept=False)
    rng = np.random.default_rng(4)
    rot = np.stack([np.linalg.qr(rng.standard_normal((nf, nf)))[0] for _ in range(R)], axis=2)
    shocks = rng.standard_normal((nf, 2))
    ref = oracle_irf(coef, fl, None, lv, shocks, 7, rot)
    out = handle.irf_cpp(coef, fl, None, lv, shocks, 7, rot)
    assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    # cholesky model with many shocks (more than one shock chunk per draw) and a rotation
    rotM = np.stack([np.linalg.qr(rng.standard_normal((M, M)))[0] for _ in range(R)], axis=2)
    sh = rng.standard_normal((M, 11))
    ref = oracle_irf(coef, None, U, lv[:M], sh, 5, rotM)
    out = handle.irf_cpp(coef, None, U, lv[:M], sh, 5, rotM)
    assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_irf_linearity_at_scale(handle):
    """Size-independent property at a larger shape (M=50, p=4, 64 draws): responses are linear in the shocks."""
    M, p, R, nf = 50, 4, 64, 6
    coef, fl, U, lv = draws(M, p, R, nf, 9)
    a = handle.irf_cpp(coef, fl, None, lv, np.eye(nf), 12)
    w = np.random.default_rng(1).standard_normal((nf, 1))
    b = handle.irf_cpp(coef, fl, None, lv, w, 12)
    comb = np.einsum("mshr,s->mhr", a, w[:, 0])
    assert np.max(np.abs(comb - b[:, 0])) <= 1e-11 * np.max(np.abs(b))


@pytest.mark.parametrize("M,p,R,nf,ahead", [
    (100, 4, 40, 1, 12),    # config 3 slice (321 KB): cluster of 2 CTAs
    (131, 4, 23, 2, 5),     # 4 CTAs, columns not divisible by the cluster size
    (200, 4, 50, 1, 12),    # config 5 slice (1.28 MB): cluster of 8 CTAs, more draws than resident clusters
    (200, 4, 9, 6, 3),      # 6 shocks: chunks of 4 so that the slice still fits next to the predictor rows
    (50, 4, 700, 6, 4),     # one CTA per draw, persistent CTAs stride over many draws
    (260, 4, 3, 1, 2),      # slice too large for 8 CTAs: many-CTA streaming kernel
])
def test_irf_cluster_forms_match_oracle(handle, M, p, R, nf, ahead):
    """The cluster form of K6 (PHI slice resident in the shared memory of 1/2/4/8 CTAs, responses exchanged through
    distributed shared memory) against the oracle recursion, every cluster size and the fallback."""
    coef, fl, U, lv = draws(M, p, R, nf, 100 + M)
    shocks = np.eye(nf)
    ref = oracle_irf(coef, fl, None, lv, shocks, ahead)
    out = handle.irf_cpp(coef, fl, None, lv, shocks, ahead)
    assert out.shape == (M, nf, ahead + 1, R)
    assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_irf_cluster_identity_shocks_no_factors_no_U(handle):
    """dim == M shocks without factors or U (plain reduced-form responses), 13 shocks -> two chunks."""
    M, p, R = 13, 2, 21
    coef, fl, U, lv = draws(M, p, R, 0, 77)
    shocks = np.eye(M)
    ref = oracle_irf(coef, None, None, None, shocks, 6)
    out = handle.irf_cpp(coef, None, None, None, shocks, 6)
    assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("force", ["0", "1"])
def test_irf_forced_forms_agree_with_oracle(handle, monkeypatch, force):
    """Both forms of K6 on the same small shapes (BVB_IRF_CLUSTER forces one): 1-CTA clusters, several shock chunks,
    a rotation, no intercept."""
    monkeypatch.setenv("BVB_IRF_CLUSTER", force)
    for (M, p, R, nf, ahead, icpt) in [(50, 4, 300, 6, 4, True), (7, 3, 40, 3, 9, False), (24, 2, 17, 11, 3, True)]:
        coef, fl, U, lv = draws(M, p, R, nf, 200 + M, intercept=icpt)
        rng = np.random.default_rng(M)
        rot = np.stack([np.linalg.qr(rng.standard_normal((nf, nf)))[0] for _ in range(R)], axis=2)
        shocks = rng.standard_normal((nf, nf + 2))
        ref = oracle_irf(coef, fl, None, lv, shocks, ahead, rot)
        out = handle.irf_cpp(coef, fl, None, lv, shocks, ahead, rot)
        assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
