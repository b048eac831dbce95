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
    if sc["type"] == "cholesky":
        st["U"] = np.array(sv["U"], float, order="F", copy=True)
        st["logvar"] = np.array(sv["sv_logvar"], float, copy=True)
        st["logvar0"] = np.array(sv["sv_logvar0"], float, copy=True)
        st["sv_para"] = np.array(sv["sv_para"], float, copy=True)
        st["d_sqrt"] = np.exp(st["logvar"] / 2.0)                                                          # :404
        st["facload"], st["fac"] = np.zeros((M, 0)), np.zeros((0, T))
        nU = M * (M - 1) // 2
        up = sc["cholesky_U_prior"]
        u = dict(**{"lambda": np.full(nU, 1.0 / max(nU, 1)), "psi": np.ones(nU), "xi": np.array([0.5]),
                    "a": np.array([float(sc["cholesky_a"])]), "b": np.array([float(sc["cholesky_b"])]),
                    "c": np.array([float(sc["cholesky_c"])]), "theta": np.full(nU, 1.0 / max(nU, 1)),
                    "zeta": np.array([10.0]), "nu": np.ones(nU), "varpi": np.array([1.0]), "gammas": np.zeros(nU),
                    "groups_idx": [np.arange(nU)], "lambda_3": 0.001})
        if up == "normal":
            u["V_i"] = np.array(sc["cholesky_V_i"], float)[:nU].copy()
        elif up == "DL":
            u["V_i"] = u["psi"] * u["lambda"]                                                             # :214
        elif up == "GT":
            u["V_i"] = u["psi"] * u["lambda"] / 2 if sc["cholesky_GT_priorkernel"] == "exponential" else u["lambda"].copy()
        elif up == "HS":
            u["V_i"] = u["theta"] * 10.0
        elif up == "SSVS":
            u["p_i"] = np.array(sc["cholesky_SSVS_p"], float)[:nU].copy()
            u["V_i"] = np.asarray(sc["cholesky_SSVS_tau0"], float)[:nU] ** 2
        elif up == "HMP":
            u["V_i"] = np.full(nU, 0.001)
        st["ustate"] = u
        st["ucfg"] = dict(GL_tol=sc["cholesky_GL_tol"], DL_plus=sc["cholesky_DL_plus"], DL_hyper=sc["cholesky_DL_hyper"],
                          GT_hyper=sc["cholesky_GT_hyper"], GT_vs=sc["cholesky_GT_vs"],
                          GT_priorkernel=sc["cholesky_GT_priorkernel"], a_vec=sc["cholesky_a_vec"],
                          a_weight=sc["cholesky_a_weight"], norm_consts=sc["cholesky_norm_consts"],
                          c_vec=sc["cholesky_c_vec"], c_rel_a=sc["cholesky_c_rel_a"],
                          SSVS_tau0=np.asarray(sc["cholesky_SSVS_tau0"], float), SSVS_tau1=np.asarray(sc["cholesky_SSVS_tau1"], float),
                          SSVS_hyper=sc["cholesky_SSVS_hyper"], SSVS_s_a=sc["cholesky_SSVS_s_a"], SSVS_s_b=sc["cholesky_SSVS_s_b"])
        phi = np.asarray(sc["sv_priorphi"], float)
        ps2 = np.asarray(sc["sv_priorsigma2"], float).reshape(M, 2, order="F")
        from . import sv as osv
        st["_svprior"] = osv.make_prior(M, mu_mean=sc["sv_priormu"][0], mu_sd=sc["sv_priormu"][1], phi_a=phi[0], phi_b=phi[1],
                                        s2_rate=ps2[:, 1], h0_var=np.asarray(sc["sv_priorh0"], float), mu_const=False)
    return st


def _sweep_sigma_cholesky(st, resid, prior_sigma, rng):
    """src/bvar_cpp.cpp:655-747."""
    from . import sv as osv
    from .gig import rgig
    sc = prior_sigma["prior_sigma_cpp"]
    T, M = resid.shape
    nU = M * (M - 1) // 2
    L_tol = prior_sigma["general_settings"]["cholesky_U_tol"]
    up = sc["cholesky_U_prior"]
    u_state, cfg = st["ustate"], st["ucfg"]
    iu = np.triu_indices(M, 1)
    order = np.lexsort((iu[0], iu[1]))                     # trimatu_ind: column-major upper triangle
    ri, ci = iu[0][order], iu[1][order]
    if nU > 0:
        st["U"] = oreg.sample_U(st["U"], resid, u_state["V_i"], st["d_sqrt"], rng.standard_normal(nU))    # :659
        u = st["U"][ri, ci]
        if up in ("DL", "GT"):                                                                             # :667-685
            zero = u == 0
            u[zero] = np.where(rng.uniform(size=zero.sum()) < 0.5, L_tol, -L_tol)
            u[(u > 0) & (u < L_tol)] = L_tol
            u[(u < 0) & (u > -L_tol)] = -L_tol
            st["U"][ri, ci] = u
        if up == "DL":
            ohyp.sample_V_i_DL(u_state, u, 0, cfg, rng)
        elif up == "GT":
            ohyp.sample_V_i_GT(u_state, u, 0, cfg, rng)
        elif up == "SSVS":
            ohyp.sample_V_i_SSVS_beta(u_state, u, 0, cfg, rng)
        elif up == "HS":
            ohyp.sample_V_i_HS(u_state, u, 0, rng)
        elif up == "HMP":                                                                                  # :370-377
            s1, r1 = sc["cholesky_lambda_3"]
            u_state["lambda_3"] = float(rgig(s1 - nU // 2, (u ** 2).sum(), 2.0 * r1, rng))
            u_state["V_i"][:] = u_state["lambda_3"]
        st["u"] = u
    str_resid = resid @ st["U"]                                                                            # :725
    het = np.asarray(sc["sv_heteroscedastic"]).astype(bool)
    off = np.asarray(sc["cholesky_sv_offset"], float)
    with np.errstate(divide="ignore"):
        ystar = np.log(str_resid ** 2 + off[None, :])
    h, h0, mu, phi, sigma, _ = osv.update_fast_sv(ystar, st["logvar"], st["logvar0"], st["sv_para"][0], st["sv_para"][1],
                                                  st["sv_para"][2], st["_svprior"], rng)
    ph = np.asarray(sc["cholesky_priorhomoscedastic"], float).reshape(M, 2, order="F")
    for j in range(M):
        if het[j]:
            st["logvar"][:, j], st["logvar0"][j] = h[:, j], h0[j]
            st["sv_para"][:, j] = [mu[j], phi[j], sigma[j]]
        else:                                                                                              # :728-732
            d_i = (ph[j, 1] + 0.5 * (str_resid[:, j] ** 2).sum()) / rng.gamma(ph[j, 0] + 0.5 * T)
            st["logvar"][:, j] = np.log(d_i)
    st["d_sqrt"] = np.exp(st["logvar"] / 2.0)


def sweep(st, prep, prior_phi, prior_sigma, rep, burnin, rng, huge=False):
    """One iteration of the loop at src/bvar_cpp.cpp:498-845 (factor structure)."""
    M, K, Kp, T = prep["M"], prep["K"], prep["Kp"], prep["T"]
    Y, X, PHI0 = prep["Y"], prep["X"], prep["PHI0"]
    pc = prior_phi["prior_phi_cpp"]
    prior = pc["prior"]
    PHI_tol = prior_phi["general_settings"]["PHI_tol"]
    V_prior = st["V_i_long"].reshape(Kp, M, order="F")
    chol = prior_sigma["prior_sigma_cpp"]["type"] == "cholesky"
    if K > 0 and chol:                                                                                   # :510
        z = rng.standard_normal((Kp, M))
        st["PHI"] = oreg.sample_PHI_collapsed(st["PHI"], PHI0, Y, X, st["U"], st["d_sqrt"], V_prior, z)
    elif K > 0:                                                                                          # :507
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
    if chol:
        _sweep_sigma_cholesky(st, resid, prior_sigma, rng)
    else:
        ofsv.update_fsv(st, np.ascontiguousarray(resid.T), st["_fsvcfg"], rng)                           # :617


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
    nU = M * (M - 1) // 2 if "U" in st else 0
    out = dict(U=np.zeros((nU, nsave)), PHI=np.zeros((Kp, M, nsave)), V_prior=np.zeros((Kp, M, nsave)), logvar=np.zeros((tv, S, nsave)),
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
            if nU:
                out["U"][:, k] = st["u"]
            out["facload"][:, :, k] = st["facload"]
            out["fac"][:, :, k] = st["fac"][:, T - tv:]
    out["state"] = st
    return out
