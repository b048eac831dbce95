"""Small statistical helpers shared by the distributional tests."""
import numpy as np
from scipy import stats


def batch_means_var(y, nbatch=40):
    """Variance of the mean of an autocorrelated series via batch means
    (stands in for coda::spectrum0.ar used by tests/testthat/test-bvar.R:23-35)."""
    y = np.asarray(y, dtype=float)
    m = y.size // nbatch
    b = y[: m * nbatch].reshape(nbatch, m).mean(axis=1)
    return b.var(ddof=1) / nbatch


def quantile_disagreement(a, b, levels=(0.1, 0.5, 0.9), nbatch=30):
    """Posterior QUANTILES of two autocorrelated chains within Monte-Carlo error (north star: "posterior means and
    quantiles must agree within Monte-Carlo standard error").  For each level the pooled quantile c is taken and the
    two chains' frequencies of {draw <= c} are compared as means of indicator series (batch-means standard errors).
    Returns the largest excess of |freq_a - freq_b| over 5 standard errors (<= slack means agreement)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    worst = -np.inf
    for q in levels:
        c = np.quantile(np.concatenate([a, b]), q)
        ia, ib = (a <= c).astype(float), (b <= c).astype(float)
        se = np.sqrt(batch_means_var(ia, nbatch) + batch_means_var(ib, nbatch))
        worst = max(worst, abs(ia.mean() - ib.mean()) - 5 * se)
    return worst


def geweke_pvalue(x_iid, y_mcmc):
    """Two-sided p-value of mean(x) == mean(y), x iid, y autocorrelated
    (tests/testthat/test-bvar.R:23-35)."""
    se2 = np.var(x_iid, ddof=1) / len(x_iid) + batch_means_var(y_mcmc)
    zstat = (np.mean(x_iid) - np.mean(y_mcmc)) / np.sqrt(se2)
    return 2 * stats.norm.sf(abs(zstat))


def geweke_phi_cholesky(sampler, draws, rng, n=50, M=5, K=10):
    """Successive-conditional simulation of tests/testthat/test-bvar.R:37-81.
    `sampler(PHI, PHI_prior, Y, X, U, d_sqrt_mat, V_prior, z)` -> new PHI."""
    U_inv = np.eye(M)
    iu = np.triu_indices(M, 1)
    # R fills upper.tri column-major
    order = np.lexsort((iu[0], iu[1]))
    vals = 0.1 + 0.1 * np.arange((M * M - M) // 2)
    U_inv[(iu[0][order], iu[1][order])] = vals
    U = np.linalg.solve(U_inv, np.eye(M))
    d_sqrt = np.sqrt(np.full(M, 20.0))
    d_sqrt_mat = np.tile(d_sqrt, (n, 1))
    SIGMA_chol = np.diag(d_sqrt) @ U_inv
    X = rng.standard_normal((n, K))
    V_prior = np.ones((K, M))
    PHI_prior = np.zeros((K, M))
    PHI = rng.standard_normal((K, M))
    q = np.empty(draws)
    for r in range(draws):
        Y = X @ PHI + rng.standard_normal((n, M)) @ SIGMA_chol
        PHI = sampler(PHI, PHI_prior, Y, X, U, d_sqrt_mat, V_prior, rng.standard_normal((K, M)))
        q[r] = np.sum(PHI ** 2)
    q_iid = np.sum(rng.standard_normal((draws, K * M)) ** 2, axis=1)
    p1 = geweke_pvalue(q_iid, q)
    p2 = stats.kstest(q[::5], "chi2", args=(K * M,)).pvalue
    return p1, p2
