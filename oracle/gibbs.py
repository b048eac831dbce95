"""Oracle: the Gibbs loop of bvar_cpp() (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/bvar_cpp.cpp:55-163 (initialisation), :498-611
(PHI draw, tolerance clamp, hyper-parameters, residuals), :617-647 (Sigma_t via
oracle.fsv) and :751-836 (storage) on the CPU with NumPy / SciPy-LAPACK.  It is
the checker for the distributional parity tests and the timed arm of bench.py's
`cpu_baseline` / `--impl reference` (the reference itself needs R, Rcpp,
RcppArmadillo, stochvol, factorstochvol and GIGrvg, none of which exist here).

Inputs are the dicts produced by the product's host-side configuration mirror
(`bayesianvars_b200.bvar.prepare / specify_prior_*`), i.e. the same lists the
reference's bvar_cpp() receives.
"""
from __future__ import annotations

import numpy as np

from . import fsv as ofsv
from . import hyper as ohyp
from . import regression as oreg


def init_state(prep, prior_phi, prior_sigma):
    M, K, Kp, T = prep["M"], prep["K"], prep["Kp"], prep["T"]
    pc, sc = prior_phi["prior_phi_cpp"], prior_sigma["prior_sigma_cpp"]
    n = K * M
    sv = prep["startvals"]
    st = dict(PHI=np.array(sv["PHI"], dtype=float, order="F", copy=True))
    prior = pc["prior"]
    i_vec = prep["i_vec"]
    i_ocl = np.flatnonzero(i_vec != 0)
    ivs = i_vec[i_ocl]
    st["i_ocl"], st["i_i"] = i_ocl, np.flatnonzero(i_vec == 0)
    G = int(pc["n_groups"])
    st.update(**{"lambda": np.full(n, 1.0 / max(n, 1)), "psi": np.ones(n), "xi": np.full(G, 0.5),      # :83-85
                 "a": np.array(pc["a"], float).copy() if np.size(pc["a"]) == G else np.zeros(G),
                 "b": np.array(pc["b"], float).copy() if np.size(pc["b"]) == G else np.zeros(G),
                 "c": np.array(pc["c"], float).copy() if np.size(pc["c"]) == G else np.zeros(G),
                 "theta": np.full(n, 1.0 / max(n, 1)), "zeta": np.full(G, 10.0), "nu": np.ones(n), "varpi": np.ones(G),
                 "gammas": np.zeros(n), "lambda_1": 0.04, "lambda_2": 0.0016})
    V_i = np.ones(n)
    if prior == "normal":
        V_i = np.array(pc["V_i"], float).copy()
    elif prior == "DL":
        V_i = st["psi"] * st["lambda"] ** 2                                                             # :105-107
    elif prior == "GT":
        V_i = st["psi"] * st["lambda"] / 2 if pc["GT_priorkernel"] == "exponential" else st["lambda"].copy()
    elif prior == "HS":
        V_i = st["theta"] * st["zeta"][0]
    elif prior == "SSVS":
        st["p_i"] = np.array(pc["SSVS_p"], float).copy()
        V_i = np.asarray(pc["SSVS_tau1"]) ** 2 if pc["SSVS_hyper"] else np.asarray(pc["SSVS_tau0"]) ** 2
    elif prior == "HMP":
        st["i_ol"], st["i_cl"] = np.flatnonzero(ivs > 0), np.flatnonzero(ivs < 0)
        Vp = np.asarray(pc["V_i_prep"], float)[i_ocl]
        st["V_prep_small"] = Vp
        V_i[st["i_ol"]] = st["lambda_1"] * Vp[st["i_ol"]]
        V_i[st["i_cl"]] = st["lambda_2"] * Vp[st["i_cl"]]
    st["V_i"] = V_i.astype(float).copy()
    if prior in ("DL", "GT", "HS", "SSVS"):
        st["groups_idx"] = [np.flatnonzero(ivs == g) for g in pc["groups"]]
    V_long = np.zeros(Kp * M)
    V_long[st["i_i"]] = prep["priorIntercept"]
    V_long[i_ocl] = st["V_i"]
    st["V_i_long"] = V_long
    if sc["type"] == "factor":
        R = sc["factor_factors"]
        restr = np.asarray(sc["factor_restrinv"]).reshape(M, R, order="F").astype(bool) if R else np.zeros((M, 0), bool)
        st["facload"] = np.where(restr, np.array(sv["facload"], float).reshape(M, R), 0.0) if R else np.zeros((M, 0))
        st["tau2"] = np.where(restr, np.array(sv["tau2"], float).reshape(M, R), 0.0) if R else np.zeros((M, 0))
        st["fac"] = np.array(sv["fac"], float).reshape(R, T).copy() if R else np.zeros((0, T))
        st["logvar"] = np.array(sv["sv_logvar"], float, copy=True)
        st["logvar0"] = np.array(sv["sv_logvar0"], float, copy=True)
        st["sv_para"] = np.array(sv["sv_para"], float, copy=True)
        st["lambda2"] = np.zeros(max(M, R, 1))
        cfg = dict(sc)
        cfg["factor_restrinv"] = restr
        cfg["factor_priorhomoskedastic"] = np.asarray(sc["factor_priorhomoskedastic"]).reshape(M, 2, order="F")
        cfg["_svprior"] = ofsv.make_sv_prior(sc, M, R)
        st["_fsvcfg"] = cfg
    return st


def sweep(st, prep, prior_phi, prior_sigma, rep, burnin, rng, huge=False):
    """One iteration of the loop at src/bvar_cpp.cpp:498-845 (factor structure)."""
    M, K, Kp, T = prep["M"], prep["K"], prep["Kp"], prep["T"]
    Y, X, PHI0 = prep["Y"], prep["X"], prep["PHI0"]
    pc = prior_phi["prior_phi_cpp"]
    prior = pc["prior"]
    PHI_tol = prior_phi["general_settings"]["PHI_tol"]
    V_prior = st["V_i_long"].reshape(Kp, M, order="F")
    if K > 0:                                                                                            # :507
        z = rng.standard_normal((Kp, M))
        delta = rng.standard_normal((T, M)) if huge else None
        st["PHI"] = oreg.sample_PHI_factor(st["PHI"], PHI0, Y, X, st["logvar"][:, :M], V_prior, st["facload"],
                                           st["fac"], huge, z, delta)
        if not np.isfinite(st["PHI"]).all():
            raise FloatingPointError(f"non-finite PHI in rep {rep + 1}.")
    PHI_diff = st["PHI"] - PHI0
    if prior in ("DL", "GT") and K > 0:                                                                  # :534-559
        d = PHI_diff[:K]
        zero = d == 0
        sign = np.where(rng.uniform(size=d.shape) < 0.5, 1.0, -1.0)
        d[zero] = (sign * PHI_tol)[zero]
        small_pos = (d > 0) & (d < PHI_tol)
        small_neg = (d < 0) & (d > -PHI_tol)
        d[small_pos] = PHI_tol
        d[small_neg] = -PHI_tol
        st["PHI"][:K] = PHI0[:K] + d
    coefs = PHI_diff.reshape(-1, order="F")[st["i_ocl"]]
    if prior in ("DL", "GT", "SSVS", "HS"):                                                              # :563-599
        for g in range(len(pc["groups"])):
            if prior == "DL":
                ohyp.sample_V_i_DL(st, coefs, g, pc, rng)
            elif prior == "GT":
                ohyp.sample_V_i_GT(st, coefs, g, pc, rng)
            elif prior == "SSVS":
                if rep > 0.1 * burnin or pc["SSVS_hyper"]:
                    ohyp.sample_V_i_SSVS_beta(st, coefs, g, pc, rng)
            else:
                ohyp.sample_V_i_HS(st, coefs, g, rng)
    elif prior == "HMP":
        if rep > 0.1 * burnin:                                                                           # :602
            ohyp.sample_V_i_HMP(st, coefs, st["V_prep_small"], st["i_ol"], st["i_cl"], pc, rng)
    st["V_i_long"][st["i_ocl"]] = st["V_i"]                                                              # :609
    resid = Y - X @ st["PHI"] if Kp > 0 else Y.copy()                                                    # :611
    ofsv.update_fsv(st, np.ascontiguousarray(resid.T), st["_fsvcfg"], rng)                               # :617


def run(prep, prior_phi, prior_sigma, draws, burnin, thin=1, sv_keep="last", rng=None, huge=False, state=None):
    """Returns the reference's list elements (same layouts as src/bvar_cpp.cpp:409-482)."""
    rng = np.random.default_rng(0) if rng is None else rng
    M, K, Kp, T = prep["M"], prep["K"], prep["Kp"], prep["T"]
    pc = prior_phi["prior_phi_cpp"]
    st = init_state(prep, prior_phi, prior_sigma) if state is None else state
    R = st["facload"].shape[1]
    S = M + R
    nsave = draws // thin
    tv = T if sv_keep == "all" else 1
    n, G = K * M, int(pc["n_groups"])
    prior = pc["prior"]
    rows = {"DL": 2 * G + 2 * n, "GT": 2 * G + 2 * n, "HS": 2 * G + 2 * n, "SSVS": 2 * n, "HMP": 2}.get(prior, 0)
    out = dict(PHI=np.zeros((Kp, M, nsave)), V_prior=np.zeros((Kp, M, nsave)), logvar=np.zeros((tv, S, nsave)),
               sv_para=np.zeros((3, S, nsave)), phi_hyperparameter=np.zeros((rows, nsave)),
               facload=np.zeros((M, R, nsave)), fac=np.zeros((R, tv, nsave)))
    for rep in range(draws + burnin):
        sweep(st, prep, prior_phi, prior_sigma, rep, burnin, rng, huge)
        if rep >= burnin and (rep + 1 - burnin) % thin == 0:                                             # :751
            k = (rep + 1 - burnin) // thin - 1
            out["PHI"][:, :, k] = st["PHI"]
            out["V_prior"][:, :, k] = st["V_i_long"].reshape(Kp, M, order="F")
            out["logvar"][:, :, k] = st["logvar"][T - tv:]
            out["sv_para"][:, :, k] = st["sv_para"]
            if prior in ("DL", "GT"):
                out["phi_hyperparameter"][:, k] = np.concatenate([st["a"], st["xi"], st["lambda"], st["psi"]])
            elif prior == "HS":
                out["phi_hyperparameter"][:, k] = np.concatenate([st["zeta"], st["varpi"], st["theta"], st["nu"]])
            elif prior == "SSVS":
                out["phi_hyperparameter"][:, k] = np.concatenate([st["gammas"], st["p_i"]])
            elif prior == "HMP":
                out["phi_hyperparameter"][:, k] = [st["lambda_1"], st["lambda_2"]]
            out["facload"][:, :, k] = st["facload"]
            out["fac"][:, :, k] = st["fac"][:, T - tv:]
    out["state"] = st
    return out
