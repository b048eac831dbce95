"""Edge shapes of bvar() + predict() + irf() through the C ABI: univariate, no lags, no factors, T < Kp, no intercept,
homoskedastic, restricted / column-wise / normal loadings priors, the maximum number of factors."""
import numpy as np
import pytest

from bayesianvars_b200 import bvar as B
from bayesianvars_b200.predict import predict

pytestmark = pytest.mark.gpu

CASES = [
    dict(M=1, lags=1, T=60, sigma=dict(type="factor", factor_factors=0), phi="HS"),
    dict(M=1, lags=2, T=60, sigma=dict(type="cholesky"), phi="DL"),
    dict(M=2, lags=1, T=12, sigma=dict(type="factor", factor_factors=1), phi="R2D2"),
    dict(M=3, lags=0, T=40, sigma=dict(type="cholesky"), phi="HS"),
    dict(M=4, lags=2, T=50, sigma=dict(type="factor", factor_factors=2, factor_heteroskedastic=False), phi="NG"),
    dict(M=4, lags=2, T=50, sigma=dict(type="cholesky", cholesky_heteroscedastic=False), phi="SSVS"),
    dict(M=5, lags=1, T=70, sigma=dict(type="factor", factor_factors=3, factor_restrict="upper"), phi="HMP"),
    dict(M=5, lags=3, T=9, sigma=dict(type="factor", factor_factors=1), phi="normal"),      # T < Kp
    dict(M=6, lags=2, T=30, sigma=dict(type="factor", factor_factors=1), phi="DL", intercept=False),
    dict(M=3, lags=1, T=40, sigma=dict(type="factor", factor_factors=1, factor_priorfacloadtype="normal"), phi="HS"),
    dict(M=3, lags=1, T=40, sigma=dict(type="factor", factor_factors=2, factor_priorfacloadtype="colwiseng",
                                       factor_interweaving=1), phi="HS"),
    dict(M=17, lags=1, T=40, sigma=dict(type="factor", factor_factors=16), phi="HS"),       # FSV_RMAX factors
]


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"M{c['M']}p{c['lags']}T{c['T']}-{c['sigma']['type']}-{c['phi']}")
def test_edge_shape_runs_end_to_end(handle, c):
    rng = np.random.default_rng(0)
    M, lags, T = c["M"], c["lags"], c["T"]
    data = rng.standard_normal((T + lags, M)) * 0.5
    pp = B.specify_prior_phi(M, lags, c["phi"])
    ps = B.specify_prior_sigma(M, **c["sigma"])
    res = B.bvar(data, lags=lags, draws=30, burnin=10, prior_phi=pp, prior_sigma=ps, handle=handle,
                 prior_intercept=(10.0 if c.get("intercept", True) else False), rng=np.random.default_rng(1), sv_keep="all")
    Kp = M * lags + (1 if c.get("intercept", True) else 0)
    assert res["PHI"].shape == (Kp, M, 30)
    for k in ("PHI", "logvar", "sv_para", "V_prior", "facload", "fac", "U"):
        assert np.isfinite(res[k]).all(), k
    assert res["logvar"].shape[0] == T                      # sv_keep = "all"
    if lags > 0:
        out = predict(res, ahead=[1, 2], stable=False, LPL=True, Y_obs=rng.standard_normal((2, M)), handle=handle)
        assert out["predictions"].shape == (2, M, 30)
        assert np.isfinite(out["predictions"]).all() and np.isfinite(out["LPL"]).all()
        nf = res["facload"].shape[1]
        irf = handle.irf_cpp(res["PHI"], res["facload"] if nf else None, res["U"] if res["U"].size else None,
                             res["logvar"][-1], np.eye(nf if nf else M)[:, :1], 3)
        assert irf.shape == (M, 1, 4, 30) and np.isfinite(irf).all()
