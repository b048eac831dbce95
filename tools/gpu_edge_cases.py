"""Edge-case smoke of bvar() through the C ABI (GPU box)."""
import sys, traceback
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
from bayesianvars_b200.predict import predict
h = Handle(0, seed=5)
rng = np.random.default_rng(0)
cases = [
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
    dict(M=3, lags=1, T=40, sigma=dict(type="factor", factor_factors=2, factor_priorfacloadtype="colwiseng", factor_interweaving=1), phi="HS"),
    dict(M=17, lags=1, T=40, sigma=dict(type="factor", factor_factors=16), phi="HS"),
]
bad = 0
for c in cases:
    M, lags, T = c["M"], c["lags"], c["T"]
    data = rng.standard_normal((T + lags, M)) * 0.5
    try:
        pp = B.specify_prior_phi(M, lags, c["phi"])
        ps = B.specify_prior_sigma(M, **c["sigma"])
        res = B.bvar(data, lags=lags, draws=30, burnin=10, prior_phi=pp, prior_sigma=ps, handle=h,
                     prior_intercept=(10.0 if c.get("intercept", True) else False), rng=np.random.default_rng(1), sv_keep="all")
        ok = all(np.isfinite(res[k]).all() for k in ("PHI", "logvar", "sv_para", "V_prior", "facload", "fac", "U"))
        msg = ""
        if lags > 0:
            out = predict(res, ahead=[1, 2], stable=False, LPL=True, Y_obs=rng.standard_normal((2, M)), handle=h)
            ok = ok and np.isfinite(out["predictions"]).all() and np.isfinite(out["LPL"]).all()
            nf = res["facload"].shape[1]
            irf = h.irf_cpp(res["PHI"], res["facload"] if nf else None, res["U"] if res["U"].size else None,
                            res["logvar"][-1], np.eye(nf if nf else M)[:, :1], 3)
            ok = ok and np.isfinite(irf).all()
        print(("ok  " if ok else "BAD ") + str({k: v for k, v in c.items()}), res["PHI"].shape, msg, flush=True)
        bad += not ok
    except Exception as e:
        bad += 1
        print("EXC " + str(c), repr(e)[:300], flush=True)
print("failures:", bad)
