"""Generate the committed golden fixtures under tests/golden/.

The reference itself cannot be run here (no R / Rcpp / Armadillo, SURVEY.md
§8c), so the fixtures hold the LAPACK-faithful oracle's outputs on seeded
inputs.  They freeze the oracle (so a later edit of oracle/ cannot silently move
the target) and give the GPU parity tests value-level vectors that travel to the
GPU box.  Run from the repo root:  python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import regression as orc  # noqa: E402
from tools.synth import phi_draw_inputs  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def phi_factor_cases():
    cases = {}
    for name, (M, p, T, r, prior, seed) in {
        "tiny_r0": (3, 1, 40, 0, "flat", 11),
        "demo_like": (10, 2, 235, 1, "dl", 12),
        "fx_like": (14, 3, 100, 2, "dl", 13),
        "two_blocks": (8, 5, 90, 3, "dl", 14),
    }.items():
        d = phi_draw_inputs(M, p, T, r, seed=seed, prior=prior)
        out = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                    d["facload"], d["fac"], False, d["z"])
        cases[name] = dict(M=M, p=p, T=T, r=r, prior=prior, seed=seed, PHI=out)
    return cases


def entry_point_cases():
    """Self-contained (inputs + oracle outputs) vectors for the other entry points of the C ABI."""
    from oracle import irf as oirf, predict as opred
    rng = np.random.default_rng(2024)
    out = {}
    # sample_PHI_cholesky (as-written Kronecker restatement) and sample_U
    M, Kp, T = 4, 9, 40
    X = rng.standard_normal((T, Kp)); X[:, -1] = 1.0
    Y = rng.standard_normal((T, M))
    U = np.triu(0.3 * rng.standard_normal((M, M)), 1) + np.eye(M)
    d_sqrt = np.exp(rng.normal(0, 0.5, (T, M)))
    V = np.exp(rng.normal(-2, 3, (Kp, M)))
    PHI0 = 0.1 * rng.standard_normal((Kp, M)); PHI = rng.standard_normal((Kp, M)); z = rng.standard_normal((Kp, M))
    out["phi_cholesky"] = dict(X=X, Y=Y, U=U, d_sqrt=d_sqrt, V=V, PHI0=PHI0, PHI=PHI, z=z,
                               out=orc.sample_PHI(PHI, PHI0, Y, X, U, d_sqrt, V, z))
    nU = M * (M - 1) // 2
    E = rng.standard_normal((T, M)); VU = np.exp(rng.normal(0, 2, nU)); zu = rng.standard_normal(nU)
    out["sample_u"] = dict(E=E, V=VU, d_sqrt=d_sqrt, z=zu, out=orc.sample_U(np.eye(M), E, VU, d_sqrt, zu))
    # expert_huge branch of sample_PHI_factor (Kp = 16 > T = 12)
    d = phi_draw_inputs(5, 3, 12, 1, seed=77)
    delta = rng.standard_normal((12, 5))
    out["phi_huge"] = dict(**{k: d[k] for k in ("Y", "X", "logvar", "V", "PHI0", "facload", "fac", "z")}, delta=delta,
                           out=orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                                     d["facload"], d["fac"], True, d["z"], delta))
    # irf_cpp: factor model, 2 lags + intercept, 3 draws
    M, p, R, nf, ahead = 4, 2, 3, 2, 5
    coef = 0.2 * rng.standard_normal((M * p + 1, M, R)); fl = rng.standard_normal((M, nf, R)); lv = rng.normal(-1, 0.3, (M + nf, R))
    shocks = np.eye(nf)
    out["irf_factor"] = dict(coefficients=coef, factor_loadings=fl, logvar_t=lv, shocks=shocks, ahead=ahead,
                             out=oirf.irf_cpp(coef, fl, None, lv, shocks, ahead))
    Uv = 0.3 * rng.standard_normal((M * (M - 1) // 2, R)); lvc = rng.normal(-1, 0.3, (M, R))
    out["irf_cholesky"] = dict(coefficients=coef, U_vecs=Uv, logvar_t=lvc, shocks=np.eye(M), ahead=ahead,
                               out=oirf.irf_cpp(coef, None, Uv, lvc, np.eye(M), ahead))
    # out_of_sample: factor model with SV on every series, LPL + sub-vector LPL + predictive draws
    each, ah = 2, np.array([1, 3])
    x = np.concatenate([rng.standard_normal(M * p), [1.0]])
    svi = np.arange(M + nf)
    mu = rng.normal(-1, 0.3, (M + nf, R)); ph = rng.uniform(0.7, 0.98, (M + nf, R)); sg = rng.uniform(0.05, 0.4, (M + nf, R))
    yobs = rng.standard_normal((2, M)); voi = np.array([0, 2])
    zl = rng.standard_normal((R, 3, each, M + nf)); zp = rng.standard_normal((R, 2, each, M))
    res = opred.out_of_sample(each, x, coef, None, fl, lv, ah, mu, ph, sg, svi, True, True, yobs, True, voi, True, zl, zp)
    out["predict_factor"] = dict(each=each, X_T_plus_1=x, PHI=coef, facload=fl, logvar_T=lv, ahead=ah, sv_mu=mu, sv_phi=ph,
                                 sv_sigma=sg, sv_indicator=svi, Y_obs=yobs, VoI=voi, z_logvar=zl, z_pred=zp,
                                 **{"out_" + k: v for k, v in res.items()})
    return out


if __name__ == "__main__":
    e = entry_point_cases()
    flat = {}
    for k, v in e.items():
        for kk, vv in v.items():
            flat[f"{k}/{kk}"] = np.asarray(vv)
    np.savez_compressed(OUT / "entry_points.npz", **flat)
    print("wrote", OUT / "entry_points.npz", sorted(e))
    c = phi_factor_cases()
    flat = {}
    for k, v in c.items():
        for kk, vv in v.items():
            flat[f"{k}/{kk}"] = np.asarray(vv)
    np.savez_compressed(OUT / "phi_factor.npz", **flat)
    print("wrote", OUT / "phi_factor.npz", sorted(c))
