"""The statistical yardsticks of the posterior-parity tests, checked on series with a known law (CPU)."""
import numpy as np

from tests.stat_helpers import batch_means_var, quantile_disagreement


def ar1(rng, n, phi, mu=0.0, scale=1.0):
    e = rng.standard_normal(n)
    x = np.empty(n)
    x[0] = e[0]
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    return mu + scale * x


def test_quantile_disagreement_accepts_equal_laws_and_rejects_shifted_or_rescaled_ones():
    rng = np.random.default_rng(0)
    same = [quantile_disagreement(ar1(rng, 3000, 0.5), ar1(rng, 1500, 0.5)) for _ in range(60)]
    assert max(same) < 0.02                      # the slack used by the GPU chain tests
    shifted = [quantile_disagreement(ar1(rng, 3000, 0.5), ar1(rng, 1500, 0.5, mu=1.0)) for _ in range(10)]
    assert min(shifted) > 0.02
    rescaled = [quantile_disagreement(ar1(rng, 3000, 0.5), ar1(rng, 1500, 0.5, scale=2.5)) for _ in range(10)]
    assert min(rescaled) > 0.02


def test_batch_means_variance_tracks_autocorrelation():
    rng = np.random.default_rng(1)
    n, phi = 20000, 0.8
    v = np.mean([batch_means_var(ar1(rng, n, phi), 40) for _ in range(20)])
    asym = 1.0 / (1.0 - phi) ** 2 / n            # long-run variance of the mean of an AR(1) with unit innovations
    assert 0.6 * asym < v < 1.5 * asym
