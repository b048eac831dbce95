"""BASELINE.json configs run end to end on the GPU and checked against the oracle chain.

configs[0]: the README demonstration (/root/reference/README.md:57-70) - usmacro_growth, 10 variables x 100,
lags 2, 2000 + 1000 sweeps, default priors (HS on PHI, factor-SV with one factor), sv_keep = "all", then
predict(ahead = 1:4, LPL = TRUE, Y_obs = test_data).
configs[2]: the north-star shape (M = 100, p = 4, T = 500, factor-SV r = 1, DL) against the committed 600-sweep
oracle chain (tools/make_golden_config3.py).
Tolerance: posterior means within 5 Monte-Carlo standard errors (batch means of both chains) plus a small absolute
slack written at each assertion."""
from pathlib import Path

import numpy as np
import pytest

from bayesianvars_b200 import bvar as B
from bayesianvars_b200.predict import predict
from oracle import gibbs as og
from oracle import predict as opred
from tests.stat_helpers import batch_means_var, quantile_disagreement

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
README_VARS = ["GDPC1", "PCECC96", "GPDIC1", "AWHMAN", "GDPCTPI", "CES2000000008x", "FEDFUNDS", "GS10", "EXUSUKx", "S&P 500"]


def test_config0_readme_demo(handle):
    g = np.load(GOLD / "usmacro_growth.npz")
    cols = [list(g["colnames"]).index(v) for v in README_VARS]
    train = 100.0 * g["values"][:237, cols]             # usmacro_growth[1:237, ...]
    test = 100.0 * g["values"][237:241, cols]           # usmacro_growth[238:241, ...]
    M, p = 10, 2
    pp, ps = B.specify_prior_phi(M, p), B.specify_prior_sigma(M)     # package defaults: HS, factor-SV with one factor
    mod = B.bvar(train, lags=p, draws=2000, burnin=1000, prior_phi=pp, prior_sigma=ps, sv_keep="all", handle=handle,
                 rng=np.random.default_rng(537))
    assert mod["PHI"].shape == (21, 10, 2000) and mod["logvar"].shape == (235, 11, 2000) and mod["fac"].shape == (235, 1, 2000)
    assert np.isfinite(mod["PHI"]).all() and np.isfinite(mod["logvar"]).all()
    # the oracle chain on the same prepared inputs
    prep = B.prepare(train, p, 10.0, pp, ps, rng=np.random.default_rng(537))
    orc = og.run(prep, pp, ps, draws=1500, burnin=700, rng=np.random.default_rng(11), sv_keep="all")
    bad = 0
    for i in range(21):
        for j in range(M):
            a, b = mod["PHI"][i, j], orc["PHI"][i, j]
            se = np.sqrt(batch_means_var(a, 25) + batch_means_var(b, 25))
            scale = max(1.0, np.abs(b).mean())           # intercepts of series in percent x 100 are O(1..10)
            assert abs(a.mean() - b.mean()) < 5 * se + 0.03 * scale, (i, j, a.mean(), b.mean(), se)
            bad += quantile_disagreement(a, b) >= 0.03
    assert bad <= 4, bad                                 # 210 coefficients x 3 quantiles at 5 MCSE: a handful may exceed
    # time-varying volatilities: posterior mean path of every series' log total variance  log(e^{h_it} + lambda_i^2 e^{h_ft})
    # (the split between a tiny idiosyncratic variance and the factor part is weakly identified and mixes slowly:
    # GDPC1's idiosyncratic log-variance sits near -7 +- 1.5 in both chains, its total variance agrees)
    def tot(o):
        return np.log(np.exp(o["logvar"][:, :M]) + o["facload"][None, :, 0, :] ** 2 * np.exp(o["logvar"][:, M:M + 1])).mean(axis=2)
    lg, lo = tot(mod), tot(orc)
    assert np.abs(lg - lo).mean() < 0.1 and np.abs(lg - lo).max() < 0.6, (np.abs(lg - lo).mean(), np.abs(lg - lo).max())
    # idiosyncratic paths of the series that are not absorbed by the factor
    ig, io = mod["logvar"][:, :M].mean(axis=2), orc["logvar"][:, :M].mean(axis=2)
    free = io.mean(axis=0) > -4.0
    assert free.sum() >= 6 and np.abs(ig - io)[:, free].mean() < 0.15 and np.abs(ig - io)[:, free].max() < 1.0
    # predict(mod, ahead = 1:4, LPL = TRUE, Y_obs = test_data)
    pred = predict(mod, ahead=[1, 2, 3, 4], LPL=True, Y_obs=test, handle=handle)
    assert pred["predictions"].shape[:2] == (4, 10) and pred["LPL"].shape == (4,) and np.isfinite(pred["LPL"]).all()
    # the same posterior draws through the oracle's out_of_sample: LPL agrees within Monte-Carlo error of the SV forecasts
    st = pred["predictions"].shape[2]
    omod = dict(mod)
    x = np.concatenate([mod["datamat"][-1, :p * M], [1.0]])
    from bayesianvars_b200.predict import stable_bvar
    ms = stable_bvar(mod, quiet=True)
    R = ms["PHI"].shape[2]
    assert st == R
    svi = np.arange(M + 1)
    rng = np.random.default_rng(5)
    want = opred.out_of_sample(1, x, ms["PHI"], None, ms["facload"], ms["logvar"][-1], np.arange(1, 5), ms["sv_para"][0],
                               ms["sv_para"][1], ms["sv_para"][2], svi, True, True, test, False, np.zeros(1, int), True,
                               rng.standard_normal((R, 4, 1, M + 1)), rng.standard_normal((R, 4, 1, M)))
    mx = want["LPL_draws"].max(axis=1) - 700.0
    lpl_o = np.log(np.exp(want["LPL_draws"] - mx[:, None]).mean(axis=1)) + mx
    assert np.abs(pred["LPL"] - lpl_o).max() < 0.35, (pred["LPL"], lpl_o)
    # predictive medians of GDP growth agree
    med_g, med_o = np.median(pred["predictions"], axis=2), np.median(want["predictions"], axis=2)
    sd_o = want["predictions"].std(axis=2)
    assert (np.abs(med_g - med_o) < 0.15 * sd_o + 0.05).all()
    del omod


def test_config3_chain_matches_oracle(handle):
    """M = 100, p = 4, T = 500, DL, r = 1 (the bench workload, same data and start values) for 1500 sweeps against
    the committed oracle chain."""
    from bench import make_problem
    gold = np.load(GOLD / "config3_chain.npz")
    data, pp, ps, prep = make_problem()
    smp = B.Sampler(handle, prep, pp, ps, draws=1200, burnin=300)
    smp.run()
    out = smp.out
    assert np.isfinite(out["PHI"]).all()
    rows, cols, eqs = gold["rows"], gold["cols"], gold["eqs"]
    M = prep["M"]
    sel = np.r_[eqs, M]
    g_phi, o_phi = out["PHI"][rows, cols, :], gold["PHI"]
    bad = 0
    for i in range(rows.size):
        a, b = g_phi[i], o_phi[i]
        se = np.sqrt(batch_means_var(a, 20) + batch_means_var(b, 20))
        assert abs(a.mean() - b.mean()) < 5 * se + 0.01, (rows[i], cols[i], a.mean(), b.mean(), se)
        # 10 / 50 / 90 % quantiles (the DL posterior is a spike with heavy tails: moments of 400 draws are too noisy)
        bad += quantile_disagreement(a, b) >= 0.03
    assert bad <= 4, bad
    for k in range(3):
        for s in range(sel.size):
            a, b = out["sv_para"][k, sel[s]], gold["sv_para"][k, s]
            se = np.sqrt(batch_means_var(a, 20) + batch_means_var(b, 20))
            assert abs(a.mean() - b.mean()) < 5 * se + 0.05, ("sv_para", k, sel[s], a.mean(), b.mean(), se)
    a, b = np.log(out["V_prior"][:-1]).mean(axis=(0, 1)), gold["logV_mean"]
    se = np.sqrt(batch_means_var(a, 20) + batch_means_var(b, 20))
    assert abs(a.mean() - b.mean()) < 5 * se + 0.1, ("logV", a.mean(), b.mean(), se)
    # log-variances at T of the idiosyncratic series (the factor's level is not identified separately from the loadings)
    a, b = out["logvar"][-1][eqs], gold["logvar_last"][:-1]
    assert np.abs(a.mean(axis=1) - b.mean(axis=1)).max() < 0.25
    # the factor's share sum_i lambda_i^2 e^{h_fT}: the synthetic data has no common factor, so this quantity is not
    # stationary within the test's horizon - it drifts towards zero for thousands of sweeps in BOTH chains (golden: 2.5 at
    # sweep 200, 0.15 at sweep 600; CUDA chain 0.05 over sweeps 300..600, different seeds give 0.02..0.5).  What is
    # compared is what the posterior says: next to the idiosyncratic variances (sum over the series ~ 100) it is negligible.
    a = ((out["facload"][:, 0, :] ** 2).sum(axis=0) * np.exp(out["logvar"][-1][M]))[:300]
    b = (gold["facload_sq"] * np.exp(gold["logvar_last"][-1]))[100:]
    idio = np.exp(out["logvar"][-1][:M]).sum(axis=0).mean()
    assert np.median(a) < 0.02 * idio and np.median(b) < 0.02 * idio, (np.median(a), np.median(b), idio)
