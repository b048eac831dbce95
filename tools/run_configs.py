"""Smoke + timing of the BASELINE.json config shapes (synthetic data) on one GPU.
usage: python tools/run_configs.py [which ...]   which in {1,2,3,4}"""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
from bayesianvars_b200.predict import predict
from bench import synth_data

which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
h = Handle(0, seed=7)


def run(name, M, p, T, pp, ps, sweeps, huge=False):
    data = synth_data(M, p, T)
    prep = B.prepare(data, p, 10.0, pp, ps, rng=np.random.default_rng(1))
    t0 = time.perf_counter()
    smp = B.Sampler(h, prep, pp, ps, draws=sweeps, burnin=0, expert_huge=huge)
    smp.run()
    dt = time.perf_counter() - t0
    tm = smp.timing()
    ok = bool(np.isfinite(smp.out["PHI"]).all())
    print(json.dumps(dict(config=name, finite=ok, e2e_sweeps_per_s=round(sweeps / dt, 1),
                          dev_ms_per_sweep=round(tm["total_ms"] / max(tm["sweeps"], 1), 4),
                          phases_ms={k: round(float(v), 4) for k, v in tm["phase_ms"].items()})), flush=True)
    res = dict(smp.out)
    res["fac"] = np.transpose(res["fac"], (1, 0, 2))
    res.update(Y=prep["Y"], X=prep["X"], lags=p, intercept=True, sigma_type=ps["prior_sigma_cpp"]["type"],
               heteroscedastic=np.asarray(ps["prior_sigma_cpp"]["sv_heteroscedastic"], bool),
               datamat=np.hstack([prep["Y"], prep["X"]]), Yraw=data)
    return res


if 1 in which:
    M, p, T = 50, 4, 300
    run("configs[1] M=50 p=4 T=300 factor(6) HS", M, p, T, B.specify_prior_phi(M, p, "HS"),
        B.specify_prior_sigma(M, type="factor", factor_factors=6), 200)
if 2 in which:
    M, p, T = 100, 4, 500
    run("configs[2] M=100 p=4 T=500 factor(1) DL", M, p, T, B.specify_prior_phi(M, p, "DL", DL_a="1/K"),
        B.specify_prior_sigma(M, type="factor", factor_factors=1), 200)
if 3 in which:
    M, p, T = 100, 4, 500
    run("configs[3] M=100 p=4 T=500 cholesky SSVS", M, p, T, B.specify_prior_phi(M, p, "SSVS"),
        B.specify_prior_sigma(M, type="cholesky", cholesky_U_prior="SSVS"), 20)
    run("configs[3] M=100 p=4 T=500 cholesky HMP", M, p, T, B.specify_prior_phi(M, p, "HMP"),
        B.specify_prior_sigma(M, type="cholesky", cholesky_U_prior="HMP"), 20)
if 4 in which:
    M, p, T = 200, 4, 500
    res = run("configs[4] M=200 p=4 T=500 factor(1) R2D2 huge", M, p, T, B.specify_prior_phi(M, p, "R2D2"),
              B.specify_prior_sigma(M, type="factor", factor_factors=1), 100, huge=True)
    R = res["PHI"].shape[2]
    t0 = time.perf_counter()
    out = h.irf_cpp(res["PHI"], res["facload"], None, res["logvar"][-1], np.eye(1), 12)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(irf_draws=R, ahead=12, seconds=round(dt, 4), dev_ms=h.last_timing()["rest_ms"], finite=bool(np.isfinite(out).all()))), flush=True)
    t0 = time.perf_counter()
    pr = predict(res, ahead=list(range(1, 13)), stable=False, handle=h)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(predict_draws=R, ahead="1:12", seconds=round(dt, 4), dev_ms=h.last_timing()["rest_ms"],
                          finite=bool(np.isfinite(pr["predictions"]).all()))), flush=True)
