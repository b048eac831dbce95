"""Host-side mirror of bvar()'s prologue (R/bvar_wrapper.R:200-399): the "empirical" prior settings."""
import numpy as np

from bayesianvars_b200 import bvar as B


def r_MP_V_prior_prep(s, K, M, icpt):
    """Literal transcription of R/aux_functions.R:81-108 (1-based indices)."""
    V = np.zeros((K, M))
    for i in range(1, M + 1):
        for j in range(1, K + 1):
            if j == K and icpt:
                V[j - 1, i - 1] = s[i - 1]
            elif (j - i) % M == 0:
                V[j - 1, i - 1] = 1 / (np.ceil(j / M) ** 2)
            else:
                r = np.ceil(j / M)
                if j % M == 0:
                    V[j - 1, i - 1] = s[i - 1] / (r * r * s[M - 1])
                elif j % M == 1:
                    V[j - 1, i - 1] = s[i - 1] / (r * r * s[0])
                else:
                    V[j - 1, i - 1] = s[i - 1] / (r * r * s[j % M - 1])
    return V.reshape(-1, order="F")


def test_minnesota_prep_matches_r_index_logic():
    s = np.random.default_rng(0).uniform(0.5, 2, 5)
    for icpt in (True, False):
        K = 5 * 3 + int(icpt)
        np.testing.assert_allclose(B.MP_V_prior_prep(s, K, 5, icpt), r_MP_V_prior_prep(s, K, 5, icpt))


def test_mp_sigma_sq_is_ar_residual_variance():
    rng = np.random.default_rng(1)
    y = np.zeros((300, 2))
    for t in range(1, 300):
        y[t] = 0.6 * y[t - 1] + rng.standard_normal(2) * np.array([1.0, 2.0])
    s = B.MP_sigma_sq(y, 6)
    assert abs(s[0] - 1.0) < 0.25 and abs(s[1] - 4.0) < 1.0


def test_prepare_fills_hmp_and_semiautomatic_ssvs_idempotently():
    rng = np.random.default_rng(2)
    data = np.cumsum(rng.standard_normal((120, 4)), axis=0) * 0.1 + rng.standard_normal((120, 4))
    ps = B.specify_prior_sigma(4, type="cholesky")
    pp = B.specify_prior_phi(4, 2, "HMP")
    B.prepare(data, 2, 10.0, pp, ps)
    v = pp["prior_phi_cpp"]["V_i_prep"]
    assert v.shape == (9 * 4,) and (v > 0).all() and v[0] == 1.0          # own first lag of equation 1
    pp = B.specify_prior_phi(4, 2, "SSVS")
    B.prepare(data, 2, 10.0, pp, ps)
    t0 = pp["prior_phi_cpp"]["SSVS_tau0"].copy()
    B.prepare(data, 2, 10.0, pp, ps)
    np.testing.assert_array_equal(t0, pp["prior_phi_cpp"]["SSVS_tau0"])   # not rescaled twice
    ratio = pp["prior_phi_cpp"]["SSVS_tau1"] / t0
    np.testing.assert_allclose(ratio, 100.0 / 0.01)
    assert t0.shape == (2 * 16,) and np.ptp(t0) > 0                        # scaled by the flat-posterior standard errors
