"""Pins the oracle restatement of out_of_sample (src/utilities_cpp.cpp:216-393) against an
independent brute-force companion-form computation (homoscedastic case: the h-step predictive
law of a VAR(p) is N(J F^h x_T, sum_s Psi_s Sigma Psi_s'))."""
import numpy as np

from oracle import predict as op


def companion(PHI, M, p, intercept):
    Kp = PHI.shape[0]
    F = np.zeros((Kp, Kp))
    F[:, :M] = PHI
    for l in range(p - 1):
        F[l * M:(l + 1) * M, (l + 1) * M:(l + 2) * M] = np.eye(M)
    if intercept:
        F[Kp - 1, Kp - 1] = 1.0
    return F


def test_predictive_mean_and_covariance_match_companion_form():
    rng = np.random.default_rng(1)
    M, p, R, r = 3, 2, 2, 1
    Kp = M * p + 1
    PHI = 0.3 * rng.standard_normal((Kp, M, R))
    facload = rng.standard_normal((M, r, R))
    logvar_T = rng.normal(-1, 0.3, (M + r, R))
    x = np.concatenate([rng.standard_normal(M * p), [1.0]])
    ahead = np.array([1, 3, 4])
    Y_obs = rng.standard_normal((3, M))
    none = np.zeros((0, R))
    res = op.out_of_sample(1, x, PHI, None, facload, logvar_T, ahead, none, none, none, np.zeros(0, int), True, True, Y_obs,
                           True, np.array([0, 2]), True, np.zeros((R, 4, 1, 0)), np.zeros((R, 3, 1, M)))
    for rr in range(R):
        F = companion(PHI[:, :, rr], M, p, True)
        Sigma = op.build_sigma(True, facload[:, :, rr], logvar_T[:, rr], r, M, None)
        Q = np.zeros((Kp, Kp)); Q[:M, :M] = Sigma
        cov = np.zeros((Kp, Kp)); mean = x.copy()
        for h in range(1, 5):
            mean = mean @ F
            cov = F.T @ cov @ F + Q
            if h in ahead:
                c = list(ahead).index(h)
                np.testing.assert_allclose(res["predictions"][c, :, rr], mean[:M], rtol=1e-10, atol=1e-12)   # z_pred = 0
                ll = op.dmvnrm_log(Y_obs[c], mean[:M], cov[:M, :M])
                np.testing.assert_allclose(res["LPL_draws"][c, rr], ll, rtol=1e-9)
                v = [0, 2]
                np.testing.assert_allclose(res["LPL_sub_draws"][c, rr], op.dmvnrm_log(Y_obs[c, v], mean[v], cov[np.ix_(v, v)]), rtol=1e-9)


def test_cholesky_sigma_and_sv_forecast_moments():
    rng = np.random.default_rng(2)
    M = 4
    u = 0.3 * rng.standard_normal(M * (M - 1) // 2)
    lv = rng.normal(0, 0.5, M)
    S = op.build_sigma(False, None, lv, 0, M, u)
    U = np.eye(M); iu = np.triu_indices(M, 1); o = np.lexsort((iu[0], iu[1])); U[iu[0][o], iu[1][o]] = u
    Ui = np.linalg.inv(U)
    np.testing.assert_allclose(S, Ui.T @ np.diag(np.exp(lv)) @ Ui, rtol=1e-12)
    # SV forecast: with z = 0 the h-step log-variance is mu + phi^h (h_T - mu)
    R, Kp = 1, M + 1
    PHI = np.zeros((Kp, M, R)); x = np.ones(Kp)
    mu, phi, sig = np.full((M, R), -1.0), np.full((M, R), 0.9), np.full((M, R), 0.2)
    res = op.out_of_sample(1, x, PHI, u[:, None], None, lv[:, None], np.array([2]), mu, phi, sig, np.arange(M), False, True,
                           np.zeros((1, M)), False, np.array([0]), False, np.zeros((R, 2, 1, M)), None)
    h2 = -1.0 + 0.81 * (lv + 1.0)
    S2 = op.build_sigma(False, None, h2, 0, M, u)
    # PHI = 0: predictive covariance at h = 2 is Sigma(h2) only
    np.testing.assert_allclose(res["LPL_draws"][0, 0], op.dmvnrm_log(np.zeros(M), np.zeros(M), S2), rtol=1e-10)
