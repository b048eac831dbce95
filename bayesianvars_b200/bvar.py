"""Host-side mirror of the reference's `bvar()` front end (R/bvar_wrapper.R:178-504)
and its prior constructors (`specify_prior_phi`, R/utility_functions.R:1984-2320;
`specify_prior_sigma`, :1011-1583), driving the sampler through the C ABI
(`bvb_gibbs_init` / `bvb_gibbs_run`) exactly as the Rcpp shim does: the two
"*_cpp" dicts below carry the same keys as the R lists `bvar_cpp()` receives and
are marshalled 1:1 into `bvb_prior_phi` / `bvb_prior_sigma`.

Only configuration, data preparation (embed, PHI0, i_vec, start values) and
result naming live here - no sampling arithmetic.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib
from ._lib import fmat, fptr
from .api import Handle

PRIOR_CODES = {"normal": 0, "DL": 1, "GT": 2, "HS": 3, "SSVS": 4, "HMP": 5, "nothing": 6}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


# --------------------------------------------------------------------------- ABI structs
class PriorPhi(C.Structure):
    _fields_ = [("prior", C.c_int), ("V_i", _dp), ("n_groups", C.c_int), ("groups", _ip),
                ("a", _dp), ("b", _dp), ("c", _dp), ("GL_tol", C.c_double),
                ("a_vec", _dp), ("a_weight", _dp), ("norm_consts", _dp), ("c_vec", _dp), ("grid_len", C.c_int),
                ("c_rel_a", C.c_int), ("DL_hyper", C.c_int), ("DL_plus", C.c_int), ("GT_hyper", C.c_int),
                ("GT_vs", C.c_double), ("GT_priorkernel", C.c_int),
                ("SSVS_tau0", _dp), ("SSVS_tau1", _dp), ("SSVS_s_a", C.c_double), ("SSVS_s_b", C.c_double),
                ("SSVS_hyper", C.c_int), ("SSVS_p", _dp), ("V_i_prep", _dp),
                ("lambda_1", C.c_double * 2), ("lambda_2", C.c_double * 2)]


class PriorSigma(C.Structure):
    _fields_ = [("type", C.c_int), ("cholesky_U_prior", C.c_int), ("cholesky_V_i", _dp),
                ("cholesky_a", C.c_double), ("cholesky_b", C.c_double), ("cholesky_c", C.c_double),
                ("cholesky_GL_tol", C.c_double),
                ("cholesky_a_vec", _dp), ("cholesky_a_weight", _dp), ("cholesky_norm_consts", _dp),
                ("cholesky_c_vec", _dp), ("cholesky_grid_len", C.c_int),
                ("cholesky_c_rel_a", C.c_int), ("cholesky_DL_hyper", C.c_int), ("cholesky_DL_plus", C.c_int),
                ("cholesky_GT_hyper", C.c_int), ("cholesky_GT_vs", C.c_double), ("cholesky_GT_priorkernel", C.c_int),
                ("cholesky_SSVS_tau0", _dp), ("cholesky_SSVS_tau1", _dp),
                ("cholesky_SSVS_s_a", C.c_double), ("cholesky_SSVS_s_b", C.c_double),
                ("cholesky_SSVS_hyper", C.c_int), ("cholesky_SSVS_p", _dp),
                ("cholesky_lambda_3", C.c_double * 2),
                ("cholesky_priorhomoscedastic", _dp), ("cholesky_sv_offset", _dp),
                ("factor_factors", C.c_int), ("factor_shrink_a", _dp), ("factor_shrink_c", _dp),
                ("factor_shrink_d", _dp), ("factor_ngprior", C.c_int), ("factor_columnwise", C.c_int),
                ("factor_facloadtol", C.c_double), ("factor_restrinv", _ip),
                ("factor_priorhomoskedastic", _dp), ("factor_interweaving", C.c_int),
                ("sv_heteroscedastic", _ip), ("sv_priormu", C.c_double * 2), ("sv_priorphi", C.c_double * 4),
                ("sv_priorsigma2", _dp), ("sv_priorh0", _dp)]


class StartVals(C.Structure):
    _fields_ = [(k, _dp) for k in ("PHI", "U", "sv_para", "sv_logvar", "sv_logvar0", "facload", "fac", "tau2")]


class Draws(C.Structure):
    _fields_ = [(k, _dp) for k in ("PHI", "U", "logvar", "sv_para", "phi_hyperparameter", "u_hyperparameter",
                                   "V_prior", "facload", "fac")]


def _register_gibbs_signatures():
    vp = C.c_void_p
    _lib.SIGNATURES.update({
        "bvb_gibbs_init": (C.c_int, [vp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, _dp, _dp, C.POINTER(PriorPhi), C.POINTER(PriorSigma),
                                     C.POINTER(StartVals), _ip, C.c_double, C.c_double, C.c_int, C.POINTER(Draws)]),
        "bvb_gibbs_run": (C.c_int, [vp, C.c_int, _ip]),
        "bvb_gibbs_timing": (C.c_int, [vp, _dp]),
        "bvb_gibbs_profile_sweep": (C.c_int, [vp, _dp, _ip]),
        "bvb_gibbs_kernel_times": (C.c_int, [vp, _dp, C.c_int]),
        "bvb_tolerance_clamp": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_uint32]),
        "bvb_gibbs_state": (C.c_int, [vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
        "bvb_comm_unique_id": (C.c_int, [C.c_char_p]),
        "bvb_comm_init": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p]),
        "bvb_create_multi": (C.c_int, [C.c_int, _ip, C.c_uint64, C.POINTER(vp)]),
        "bvb_gibbs_init_multi": (C.c_int, [C.c_int, C.POINTER(vp), _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, _dp, _dp, C.POINTER(PriorPhi), C.POINTER(PriorSigma),
                                           C.POINTER(StartVals), _ip, C.c_double, C.c_double, C.c_int, C.POINTER(Draws)]),
        "bvb_gibbs_run_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.c_int, _ip]),
        "bvb_destroy_multi": (None, [C.c_int, C.POINTER(vp)]),
    })


_register_gibbs_signatures()


# --------------------------------------------------------------------------- priors
def specify_prior_phi(M, lags=1, prior="HS", priormean=0.0, PHI_tol=1e-18, DL_a="1/K", DL_tol=0.0,
                      R2D2_a=0.1, R2D2_b=0.5, R2D2_tol=0.0, NG_a=0.1, NG_b=1.0, NG_c=1.0, NG_tol=0.0,
                      SSVS_c0=0.01, SSVS_c1=100.0, SSVS_semiautomatic=True, SSVS_p=0.5,
                      HMP_lambda1=(0.01, 0.01), HMP_lambda2=(0.01, 0.01), normal_sds=10.0,
                      global_grouping="global", DL_plus=False, DL_b=0.5, DL_c=None):
    """R/utility_functions.R:1984-2320 -> {"prior_phi_cpp": {...27 keys...}, "general_settings": {...}}."""
    M, lags = int(M), int(lags)
    K, n = lags * M, lags * M * M
    cpp = dict(prior=prior if lags > 0 else "nothing", V_i=np.zeros(1), n_groups=1, groups=np.ones(1, np.int32),
               a=np.zeros(1), b=np.zeros(1), c=np.zeros(1), GL_tol=0.0, a_vec=np.zeros(0), a_weight=np.zeros(0),
               norm_consts=np.zeros(0), c_vec=np.zeros(0), c_rel_a=False, DL_hyper=False, DL_plus=False,
               GT_hyper=False, GT_vs=1.0, GT_priorkernel="normal", SSVS_tau0=np.zeros(1), SSVS_tau1=np.zeros(1),
               SSVS_s_a=1.0, SSVS_s_b=1.0, SSVS_hyper=False, SSVS_p=np.zeros(1), V_i_prep=np.zeros(1),
               lambda_1=np.asarray(HMP_lambda1, float), lambda_2=np.asarray(HMP_lambda2, float))
    if lags == 0:
        return dict(prior_phi_cpp=cpp, general_settings=dict(M=M, lags=0, PHI0=np.zeros((0, M)), PHI_tol=0.0,
                                                             i_mat=np.zeros((0, M), np.int32), SSVS_semiautomatic=False))
    if prior not in ("DL", "HMP", "SSVS", "normal", "R2D2", "HS", "NG"):
        raise ValueError("Argument 'prior' must be one of 'DL', 'HS', 'NG', 'R2D2', 'SSVS', 'HMP' or 'normal'.")
    if prior in ("DL", "SSVS", "R2D2", "HS", "NG"):
        if isinstance(global_grouping, np.ndarray):
            i_mat = global_grouping.astype(np.int32)
            if i_mat.shape != (K, M):
                raise ValueError("Argument 'global_grouping' is not of dimension 'c(lags*M,M)'!")
        elif global_grouping == "global":
            i_mat = np.ones((K, M), np.int32)
        elif global_grouping == "fol":
            i_mat = np.ones((K, M), np.int32)
            i_mat[np.arange(M), np.arange(M)] = 2
        elif global_grouping == "olcl-lagwise":
            i_mat = np.ones((K, M), np.int32)
            i_mat[np.arange(M), np.arange(M)] = 2
            for j in range(1, lags):
                i_mat[j * M:(j + 1) * M, :] = i_mat[(j - 1) * M:j * M, :] + 2
        elif global_grouping == "equation-wise":
            i_mat = np.tile(np.arange(1, M + 1, dtype=np.int32), (K, 1))
        elif global_grouping == "covariate-wise":
            i_mat = np.tile(np.arange(1, K + 1, dtype=np.int32)[:, None], (1, M))
        else:
            raise ValueError("Argument 'global_grouping' must be one of 'global', 'equation-wise', "
                             "'covariate-wise', 'olcl-lagwise' or 'fol'.")
        flat = i_mat.reshape(-1, order="F")
        _, first = np.unique(flat, return_index=True)
        cpp["groups"] = flat[np.sort(first)].astype(np.int32)   # unique() keeps first-appearance order
        cpp["n_groups"] = int(cpp["groups"].size)
    else:  # HMP / normal: +lag on own lags, -lag on cross lags  (:2078-2085)
        i1 = -np.ones((M, M), np.int32)
        i1[np.arange(M), np.arange(M)] = 1
        i_mat = np.concatenate([i1 * j for j in range(1, lags + 1)], axis=0).astype(np.int32)
    G = cpp["n_groups"]
    pm = np.atleast_1d(np.asarray(priormean, float))
    if pm.ndim == 2:
        PHI0 = pm.copy()
    else:
        if pm.size == 1:
            pm = np.repeat(pm, M)
        PHI0 = np.zeros((K, M))
        PHI0[np.arange(M), np.arange(M)] = pm
    gs = dict(M=M, lags=lags, PHI0=PHI0, PHI_tol=PHI_tol, global_grouping=global_grouping, i_mat=i_mat,
              SSVS_semiautomatic=False)

    def grid(a):
        a = np.asarray(a, float)
        from scipy.special import gammaln
        return a[:, 0].copy(), a[:, 1].copy(), gammaln(a[:, 0])

    if prior == "DL":
        cpp["GL_tol"] = DL_tol
        if isinstance(DL_a, str):
            if DL_a not in ("1/K", "1/n"):
                raise ValueError("DL_a must be numeric, '1/K', '1/n' or a 2-column grid")
            cpp["a"] = np.full(G, 1.0 / K if DL_a == "1/K" else 1.0 / n)
        elif np.ndim(DL_a) == 2:
            cpp["DL_hyper"] = True
            cpp["a_vec"], cpp["a_weight"], cpp["norm_consts"] = grid(DL_a)
            cpp["a"] = np.full(G, cpp["a_vec"][0])
        else:
            cpp["a"] = np.resize(np.asarray(DL_a, float), G)
        if DL_plus:
            cpp["DL_plus"] = True
            cpp["b"] = np.full(G, float(DL_b))
            if DL_c is None:
                cpp["c"] = 0.5 * cpp["a"]
                if cpp["DL_hyper"]:
                    cpp["c_vec"] = 0.5 * cpp["a_vec"]
                    cpp["c_rel_a"] = True
            else:
                cpp["c"] = np.full(G, float(DL_c))
        else:
            cpp["b"] = np.zeros(G)
            cpp["c"] = np.zeros(G)
    elif prior in ("R2D2", "NG"):
        cpp["prior"] = "GT"
        a_in = R2D2_a if prior == "R2D2" else NG_a
        cpp["b"] = np.full(G, float(R2D2_b if prior == "R2D2" else NG_b))
        cpp["GT_vs"] = 0.5 if prior == "R2D2" else 1.0
        cpp["GT_priorkernel"] = "exponential" if prior == "R2D2" else "normal"
        cpp["GL_tol"] = R2D2_tol if prior == "R2D2" else NG_tol
        if np.ndim(a_in) == 2:
            cpp["GT_hyper"] = True
            cpp["a_vec"], cpp["a_weight"], cpp["norm_consts"] = grid(a_in)
            cpp["a"] = np.full(G, cpp["a_vec"][0])   # the reference initialises by sample(); any grid point is valid
        else:
            cpp["a"] = np.resize(np.asarray(a_in, float), G)
        if prior == "R2D2":   # c = "0.5*a"
            cpp["c_rel_a"] = True
            cpp["c"] = 0.5 * cpp["a"]
            if cpp["GT_hyper"]:
                cpp["c_vec"] = 0.5 * cpp["a_vec"]
        else:
            cpp["c"] = np.full(G, float(NG_c))
            if cpp["GT_hyper"]:
                cpp["c_vec"] = np.zeros_like(cpp["a_vec"])
    elif prior == "SSVS":
        if np.size(SSVS_p) == 2:
            cpp["SSVS_s_a"], cpp["SSVS_s_b"] = (float(x) for x in SSVS_p)
            cpp["SSVS_p"] = np.full(n, 0.5)
            cpp["SSVS_hyper"] = True
        else:
            cpp["SSVS_p"] = np.full(n, float(SSVS_p))
        cpp["SSVS_tau0"] = np.resize(np.asarray(SSVS_c0, float), n)
        cpp["SSVS_tau1"] = np.resize(np.asarray(SSVS_c1, float), n)
        gs["SSVS_semiautomatic"] = bool(SSVS_semiautomatic)
        gs["SSVS_c0"], gs["SSVS_c1"] = cpp["SSVS_tau0"].copy(), cpp["SSVS_tau1"].copy()   # unscaled (prepare() rescales)
    elif prior == "normal":
        cpp["V_i"] = np.resize(np.asarray(normal_sds, float) ** 2, n)
    return dict(prior_phi_cpp=cpp, general_settings=gs)


def specify_prior_sigma(M, type="factor", factor_factors=1, factor_restrict="none",
                        factor_priorfacloadtype="rowwiseng", factor_priorfacload=0.1, factor_facloadtol=1e-14,
                        factor_priorng=(1.0, 1.0), factor_priormu=(0.0, 10.0), factor_priorphiidi=(10.0, 3.0),
                        factor_priorphifac=(10.0, 3.0), factor_priorsigmaidi=1.0, factor_priorsigmafac=1.0,
                        factor_priorh0idi="stationary", factor_priorh0fac="stationary",
                        factor_heteroskedastic=True, factor_priorhomoskedastic=None, factor_interweaving=4,
                        cholesky_U_prior="HS", cholesky_U_tol=1e-18, cholesky_heteroscedastic=True,
                        cholesky_priormu=(0.0, 100.0), cholesky_priorphi=(20.0, 1.5),
                        cholesky_priorsigma2=(0.5, 0.5), cholesky_priorh0="stationary",
                        cholesky_priorhomoscedastic=None, cholesky_DL_a="1/n", cholesky_DL_tol=0.0,
                        cholesky_R2D2_a=0.4, cholesky_R2D2_b=0.5, cholesky_R2D2_tol=0.0,
                        cholesky_NG_a=0.5, cholesky_NG_b=0.5, cholesky_NG_c=0.5, cholesky_NG_tol=0.0,
                        cholesky_SSVS_c0=0.001, cholesky_SSVS_c1=1.0, cholesky_SSVS_p=0.5,
                        cholesky_HMP_lambda3=(0.01, 0.01), cholesky_normal_sds=10.0, expert_sv_offset=0.0):
    """R/utility_functions.R:1011-1583 -> {"prior_sigma_cpp": {...39 keys...}, "general_settings": {...}}."""
    M = int(M)
    cpp = dict(type=type)
    gs = dict(M=M, cholesky_U_tol=cholesky_U_tol)
    if type == "factor":
        r = int(factor_factors)
        restr = np.zeros((M, r), bool)
        if factor_restrict == "upper":
            restr[np.triu_indices(M, 1, r)] = True
        cpp["factor_factors"] = r
        cpp["factor_restrinv"] = (~restr).astype(np.int32)
        cpp["factor_ngprior"] = factor_priorfacloadtype != "normal"
        cpp["factor_columnwise"] = factor_priorfacloadtype == "colwiseng"
        units = r if cpp["factor_columnwise"] else M
        if factor_priorfacloadtype == "normal":
            gs["factor_starttau2"] = np.full((M, r), float(factor_priorfacload) ** 2)
            sa = sc = sd = np.zeros(max(units, 1))
        else:
            gs["factor_starttau2"] = np.ones((M, r))
            sa = np.full(max(units, 1), float(factor_priorfacload))
            sc = np.full(max(units, 1), float(factor_priorng[0]))
            sd = np.full(max(units, 1), float(factor_priorng[1]))
        cpp["factor_shrinkagepriors"] = dict(a=sa, c=sc, d=sd)
        cpp["factor_facloadtol"] = float(factor_facloadtol)
        cpp["factor_interweaving"] = int(factor_interweaving)
        het = np.resize(np.asarray(factor_heteroskedastic, bool), M + r) if np.size(factor_heteroskedastic) != 2 else \
            np.concatenate([np.repeat(bool(factor_heteroskedastic[0]), M), np.repeat(bool(factor_heteroskedastic[1]), r)])
        if not het[M:].all() and cpp["factor_interweaving"] in (2, 4):
            cpp["factor_interweaving"] = 3
        cpp["sv_heteroscedastic"] = het.astype(np.int32)
        cpp["sv_priormu"] = np.asarray(factor_priormu, float)
        cpp["sv_priorphi"] = np.concatenate([np.asarray(factor_priorphiidi, float), np.asarray(factor_priorphifac, float)])
        psig = np.concatenate([np.resize(np.asarray(factor_priorsigmaidi, float), M),
                               np.resize(np.asarray(factor_priorsigmafac, float), r)])
        cpp["sv_priorsigma2"] = np.asfortranarray(np.column_stack([np.full(M + r, 0.5), 0.5 / psig]))

        def h0(v, k):
            return np.full(k, -1.0) if isinstance(v, str) else np.resize(np.asarray(v, float) ** 2, k)
        cpp["sv_priorh0"] = np.concatenate([h0(factor_priorh0idi, M), h0(factor_priorh0fac, r)])
        if factor_priorhomoskedastic is None:
            factor_priorhomoskedastic = np.tile([1.1, 0.055], (M, 1))
        cpp["factor_priorhomoskedastic"] = np.asfortranarray(np.asarray(factor_priorhomoskedastic, float))
    elif type == "cholesky":
        n_U = (M * M - M) // 2
        cpp["factor_factors"] = 0
        het = bool(cholesky_heteroscedastic)
        cpp["sv_heteroscedastic"] = np.full(M, int(het), np.int32)
        cpp["sv_priormu"] = np.asarray(cholesky_priormu, float)
        cpp["sv_priorphi"] = np.asarray(cholesky_priorphi, float)
        ps2 = np.asarray(cholesky_priorsigma2, float)
        cpp["sv_priorsigma2"] = np.asfortranarray(np.tile(ps2, (M, 1)) if ps2.ndim == 1 else ps2)
        cpp["sv_priorh0"] = np.full(M, -1.0) if isinstance(cholesky_priorh0, str) else np.resize(np.asarray(cholesky_priorh0, float) ** 2, M)
        cpp["cholesky_sv_offset"] = np.resize(np.asarray(expert_sv_offset, float), M)
        ph = np.array([0.01, 0.01]) if cholesky_priorhomoscedastic is None else np.asarray(cholesky_priorhomoscedastic, float)
        cpp["cholesky_priorhomoscedastic"] = np.asfortranarray(np.tile(ph, (M, 1)) if ph.ndim == 1 else ph)
        up = cholesky_U_prior
        cpp.update(cholesky_U_prior=up, cholesky_V_i=np.zeros(max(n_U, 1)), cholesky_a=0.0, cholesky_b=0.0, cholesky_c=0.0,
                   cholesky_GL_tol=0.0, cholesky_a_vec=np.zeros(0), cholesky_a_weight=np.zeros(0),
                   cholesky_norm_consts=np.zeros(0), cholesky_c_vec=np.zeros(0), cholesky_c_rel_a=False,
                   cholesky_DL_hyper=False, cholesky_DL_plus=False, cholesky_GT_hyper=False, cholesky_GT_vs=1.0,
                   cholesky_GT_priorkernel="normal", cholesky_SSVS_tau0=np.zeros(max(n_U, 1)),
                   cholesky_SSVS_tau1=np.zeros(max(n_U, 1)), cholesky_SSVS_s_a=1.0, cholesky_SSVS_s_b=1.0,
                   cholesky_SSVS_hyper=False, cholesky_SSVS_p=np.zeros(max(n_U, 1)),
                   cholesky_lambda_3=np.asarray(cholesky_HMP_lambda3, float))
        if up == "DL":
            cpp["cholesky_GL_tol"] = cholesky_DL_tol
            cpp["cholesky_a"] = 1.0 / max(n_U, 1) if cholesky_DL_a == "1/n" else float(cholesky_DL_a)
        elif up == "R2D2":
            cpp.update(cholesky_U_prior="GT", cholesky_a=float(cholesky_R2D2_a), cholesky_b=float(cholesky_R2D2_b),
                       cholesky_c=0.5 * float(cholesky_R2D2_a), cholesky_c_rel_a=True, cholesky_GT_vs=0.5,
                       cholesky_GT_priorkernel="exponential", cholesky_GL_tol=cholesky_R2D2_tol)
        elif up == "NG":
            cpp.update(cholesky_U_prior="GT", cholesky_a=float(cholesky_NG_a), cholesky_b=float(cholesky_NG_b),
                       cholesky_c=float(cholesky_NG_c), cholesky_GT_vs=1.0, cholesky_GT_priorkernel="normal",
                       cholesky_GL_tol=cholesky_NG_tol)
        elif up == "SSVS":
            if np.size(cholesky_SSVS_p) == 2:
                cpp["cholesky_SSVS_s_a"], cpp["cholesky_SSVS_s_b"] = (float(x) for x in cholesky_SSVS_p)
                cpp["cholesky_SSVS_hyper"] = True
                cholesky_SSVS_p = 0.5
            cpp["cholesky_SSVS_tau0"] = np.full(max(n_U, 1), float(cholesky_SSVS_c0))
            cpp["cholesky_SSVS_tau1"] = np.full(max(n_U, 1), float(cholesky_SSVS_c1))
            cpp["cholesky_SSVS_p"] = np.full(max(n_U, 1), float(cholesky_SSVS_p))
        elif up == "normal":
            cpp["cholesky_V_i"] = np.resize(np.asarray(cholesky_normal_sds, float) ** 2, max(n_U, 1))
        elif up not in ("HS", "HMP"):
            raise ValueError("Argument 'cholesky_U_prior' must be one of 'DL', 'R2D2', 'NG', 'HS', 'SSVS', 'HMP' or 'normal'.")
    else:
        raise ValueError("type must be 'factor' or 'cholesky'")
    return dict(prior_sigma_cpp=cpp, general_settings=gs)


# --------------------------------------------------------------------------- sharding
def shard_equations(M, rank, world):
    """Contiguous block of equations owned by `rank` (mirrors bvb_gibbs_init: blocks of
    ceil(M / world); trailing ranks may own nothing).  Returns (first, count)."""
    nloc_max = -(-M // world)
    j0 = min(M, rank * nloc_max)
    return j0, max(0, min(M, j0 + nloc_max) - j0)


def shard_draws(R, rank, world):
    """Contiguous range of stored draws a rank processes in irf / predict (disjoint output slices)."""
    per = -(-R // world)
    r0 = min(R, rank * per)
    return r0, max(0, min(R, r0 + per) - r0)


# --------------------------------------------------------------------------- bvar()
def MP_sigma_sq(Y, p):
    """Residual variances of univariate AR(p) fits (R/aux_functions.R:63-79)."""
    Y = np.asarray(Y, float)
    T, M = Y.shape
    out = np.empty(M)
    for i in range(M):
        y = Y[:, i]
        Xi = np.column_stack([y[p - l:T - l] for l in range(1, p + 1)])
        yi = y[p:]
        phi = np.linalg.lstsq(Xi, yi, rcond=None)[0]
        out[i] = ((yi - Xi @ phi) ** 2).sum() / (T - p)
    return out


def MP_V_prior_prep(sigma_sq, K, M, intercept):
    """Minnesota prior variances up to the lambdas (R/aux_functions.R:81-108); K counts the intercept row."""
    V = np.zeros((K, M))
    for i in range(M):
        for j in range(K):            # j + 1 is R's 1-based regressor index
            if j == K - 1 and intercept:
                V[j, i] = sigma_sq[i]
            else:
                r = j // M + 1
                V[j, i] = 1.0 / r ** 2 if (j - i) % M == 0 else sigma_sq[i] / (r ** 2 * sigma_sq[j % M])
    return V.reshape(-1, order="F")


def _embed(Y_tmp, lags, intercept):
    """R's embed(Y_tmp, lags+1)[, -(1:M)] (+ ones column): R/bvar_wrapper.R:228-235."""
    Traw, M = Y_tmp.shape
    cols = [Y_tmp[lags - l:Traw - l, :] for l in range(1, lags + 1)]
    X = np.concatenate(cols, axis=1) if lags > 0 else np.zeros((Traw, 0))
    if intercept:
        X = np.concatenate([X, np.ones((Traw - lags, 1))], axis=1)
    return np.asfortranarray(X), np.asfortranarray(Y_tmp[lags:, :])


def prepare(data, lags, prior_intercept, prior_phi, prior_sigma, startvals=None, rng=None):
    """Everything bvar() computes before calling bvar_cpp (R/bvar_wrapper.R:200-399)."""
    rng = np.random.default_rng(0) if rng is None else rng
    Y_tmp = np.asarray(data, float)
    Traw, M = Y_tmp.shape
    lags = int(lags)
    K = lags * M
    intercept = 0 if prior_intercept is False or prior_intercept is None else 1
    X, Y = _embed(Y_tmp, lags, intercept)
    Tobs = Traw - lags
    Kp = K + intercept
    gs = prior_phi["general_settings"]
    if gs["M"] != M or gs["lags"] != lags:
        raise ValueError("prior_phi does not match data / lags")
    i_mat = np.asarray(gs["i_mat"], np.int32).reshape(K, M)
    if intercept:
        i_mat = np.concatenate([i_mat, np.zeros((1, M), np.int32)], axis=0)
        pi = np.atleast_1d(np.asarray(prior_intercept, float)) ** 2          # variances (:279)
        priorIntercept = np.repeat(pi, M) if pi.size == 1 else pi
    else:
        priorIntercept = np.zeros(0)
    i_vec = np.ascontiguousarray(i_mat.reshape(-1, order="F"), dtype=np.int32)
    PHI0 = np.zeros((Kp, M), order="F")
    if lags > 0:
        PHI0[:K, :] = gs["PHI0"]
    sv = dict(startvals or {})
    pcpp = prior_sigma["prior_sigma_cpp"]
    if "PHI" not in sv:
        # posterior mean under a flat conjugate prior (:297-320)
        if Kp > 0:
            P0 = np.eye(Kp) / 1e3
            V_post = np.linalg.inv(P0 + X.T @ X)
            sv["PHI"] = V_post @ (P0 @ PHI0 + X.T @ Y)
        else:
            sv["PHI"] = np.zeros((0, M))
    # "empirical" prior settings (R/bvar_wrapper.R:321-344); written into the prior object like bvar() does with its copy
    ppc = prior_phi["prior_phi_cpp"]
    if ppc["prior"] == "SSVS" and gs.get("SSVS_semiautomatic", False) and K > 0:
        P0 = np.eye(Kp) / 1e3
        V_post_flat = np.linalg.inv(P0 + X.T @ X)
        PHI_flat = V_post_flat @ (P0 @ PHI0 + X.T @ Y)
        E = Y - X @ PHI_flat
        S_post = np.eye(M) + E.T @ E + (PHI_flat - PHI0).T @ P0 @ (PHI_flat - PHI0)
        Sigma_flat = S_post / (M + 2 + Tobs - M - 1)
        sigma_phi = np.sqrt(np.outer(np.diag(Sigma_flat), np.diag(V_post_flat)).reshape(-1))   # diag(Sigma %x% V), equation-major
        sigma_phi = sigma_phi[i_vec != 0]
        if sigma_phi.size != lags * M * M:
            raise ValueError(f"{sigma_phi.size} length(sigma_phi)!= lags*M^2")
        ppc["SSVS_tau0"] = gs["SSVS_c0"] * sigma_phi
        ppc["SSVS_tau1"] = gs["SSVS_c1"] * sigma_phi
    elif ppc["prior"] == "HMP" and np.size(ppc.get("V_i_prep", ())) != Kp * M:
        ppc["V_i_prep"] = MP_V_prior_prep(MP_sigma_sq(Y_tmp, 6), Kp, M, intercept > 0)
    if pcpp["type"] == "cholesky":
        if "U" not in sv:
            # R/bvar_wrapper.R:310-319: Sigma_flat from the flat conjugate posterior, U = L_flat
            P0 = np.eye(Kp) / 1e3 if Kp else np.zeros((0, 0))
            PHI_flat = np.linalg.inv(P0 + X.T @ X) @ (P0 @ PHI0 + X.T @ Y) if Kp else np.zeros((0, M))
            E = Y - X @ PHI_flat if Kp else Y
            S_post = np.eye(M) + E.T @ E + ((PHI_flat - PHI0).T @ P0 @ (PHI_flat - PHI0) if Kp else 0.0)
            Sigma_flat = S_post / (M + 2 + Tobs - M - 1)
            Uc = np.linalg.cholesky(Sigma_flat).T            # R's chol(): upper
            L_inv = Uc / np.sqrt(np.diag(Uc) ** 2)[:, None]   # U / sqrt(D) divides row-wise (column recycling)
            sv["U"] = np.linalg.solve(L_inv, np.eye(M))       # backsolve(L_inv, diag(M))
        sv.setdefault("sv_para", np.asfortranarray(np.vstack([np.full(M, -10.0), np.full(M, 0.9), np.full(M, 0.2)])))
        sv.setdefault("sv_logvar0", np.full(M, -10.0))
        sv.setdefault("sv_logvar", np.full((Tobs, M), -10.0, order="F"))
    if pcpp["type"] == "factor":
        r = pcpp["factor_factors"]
        S = M + r
        # random start values as R/bvar_wrapper.R:365-385 (own generator; R's stream is not reproducible)
        sv.setdefault("facload", rng.normal(0, 0.5, (M, r)) ** 2)
        sv.setdefault("fac", rng.normal(0, 0.1, (r, Tobs)))
        mu = np.concatenate([-3.0 + rng.standard_normal(M), np.zeros(r)])
        phi = 0.8 + np.minimum(rng.normal(0, 0.06, S), 0.095)
        sigma = 0.1 + rng.gamma(1.0, 1.0 / 10.0, S)
        sv.setdefault("sv_para", np.asfortranarray(np.vstack([mu, phi, sigma])))
        lv = mu[0] + rng.standard_normal((Tobs, S))
        het = np.asarray(pcpp["sv_heteroscedastic"], bool)
        lv[:, M:][:, ~het[M:]] = 0.0
        sv.setdefault("sv_logvar", np.asfortranarray(lv))
        sv.setdefault("sv_logvar0", mu[0] + rng.standard_normal(S))
        sv.setdefault("tau2", np.asarray(prior_sigma["general_settings"]["factor_starttau2"], float))
    return dict(Y=Y, X=X, M=M, T=Tobs, K=K, Kp=Kp, intercept=intercept, priorIntercept=priorIntercept, PHI0=PHI0,
                i_vec=i_vec, i_mat=i_mat, startvals=sv)


class _Keep:
    """Keeps the numpy buffers behind the ctypes pointers alive."""

    def __init__(self):
        self.refs = []

    def d(self, a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64) if np.ndim(a) <= 1 else np.asfortranarray(a, dtype=np.float64)
        self.refs.append(a)
        return a.ctypes.data_as(_dp)

    def i(self, a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.int32) if np.ndim(a) <= 1 else np.asfortranarray(a, dtype=np.int32)
        self.refs.append(a)
        return a.ctypes.data_as(_ip)


def marshal_prior_phi(cpp, keep):
    p = PriorPhi()
    p.prior = PRIOR_CODES[cpp["prior"]]
    p.V_i = keep.d(cpp["V_i"])
    p.n_groups = int(cpp["n_groups"])
    p.groups = keep.i(cpp["groups"])
    p.a, p.b, p.c = keep.d(cpp["a"]), keep.d(cpp["b"]), keep.d(cpp["c"])
    p.GL_tol = float(cpp["GL_tol"])
    p.grid_len = int(np.size(cpp["a_vec"]))
    if p.grid_len:
        p.a_vec, p.a_weight = keep.d(cpp["a_vec"]), keep.d(cpp["a_weight"])
        p.norm_consts = keep.d(cpp["norm_consts"])
        p.c_vec = keep.d(cpp["c_vec"] if np.size(cpp["c_vec"]) == p.grid_len else np.zeros(p.grid_len))
    p.c_rel_a, p.DL_hyper, p.DL_plus, p.GT_hyper = (int(bool(cpp[k])) for k in ("c_rel_a", "DL_hyper", "DL_plus", "GT_hyper"))
    p.GT_vs = float(cpp["GT_vs"])
    p.GT_priorkernel = 1 if cpp["GT_priorkernel"] == "exponential" else 0
    p.SSVS_tau0, p.SSVS_tau1, p.SSVS_p = keep.d(cpp["SSVS_tau0"]), keep.d(cpp["SSVS_tau1"]), keep.d(cpp["SSVS_p"])
    p.SSVS_s_a, p.SSVS_s_b = float(cpp["SSVS_s_a"]), float(cpp["SSVS_s_b"])
    p.SSVS_hyper = int(bool(cpp["SSVS_hyper"]))
    p.V_i_prep = keep.d(cpp["V_i_prep"])
    p.lambda_1 = (C.c_double * 2)(*np.asarray(cpp["lambda_1"], float))
    p.lambda_2 = (C.c_double * 2)(*np.asarray(cpp["lambda_2"], float))
    return p


def marshal_prior_sigma(cpp, keep):
    s = PriorSigma()
    if cpp["type"] == "factor":
        s.type = 0
        s.factor_factors = int(cpp["factor_factors"])
        sp = cpp["factor_shrinkagepriors"]
        s.factor_shrink_a, s.factor_shrink_c, s.factor_shrink_d = keep.d(sp["a"]), keep.d(sp["c"]), keep.d(sp["d"])
        s.factor_ngprior, s.factor_columnwise = int(bool(cpp["factor_ngprior"])), int(bool(cpp["factor_columnwise"]))
        s.factor_facloadtol = float(cpp["factor_facloadtol"])
        s.factor_restrinv = keep.i(cpp["factor_restrinv"])
        s.factor_priorhomoskedastic = keep.d(cpp["factor_priorhomoskedastic"])
        s.factor_interweaving = int(cpp["factor_interweaving"])
        s.sv_heteroscedastic = keep.i(cpp["sv_heteroscedastic"])
        s.sv_priormu = (C.c_double * 2)(*cpp["sv_priormu"])
        s.sv_priorphi = (C.c_double * 4)(*cpp["sv_priorphi"])
        s.sv_priorsigma2 = keep.d(cpp["sv_priorsigma2"])
        s.sv_priorh0 = keep.d(cpp["sv_priorh0"])
    else:
        s.type = 1
        s.factor_factors = 0
        s.cholesky_U_prior = PRIOR_CODES[cpp["cholesky_U_prior"]]
        s.cholesky_V_i = keep.d(cpp["cholesky_V_i"])
        for k in ("a", "b", "c", "GL_tol", "GT_vs", "SSVS_s_a", "SSVS_s_b"):
            setattr(s, "cholesky_" + k, float(cpp["cholesky_" + k]))
        s.cholesky_grid_len = int(np.size(cpp["cholesky_a_vec"]))
        if s.cholesky_grid_len:
            for k in ("a_vec", "a_weight", "norm_consts"):
                setattr(s, "cholesky_" + k, keep.d(cpp["cholesky_" + k]))
            cv = cpp["cholesky_c_vec"]
            s.cholesky_c_vec = keep.d(cv if np.size(cv) == s.cholesky_grid_len else np.zeros(s.cholesky_grid_len))
        for k in ("c_rel_a", "DL_hyper", "DL_plus", "GT_hyper", "SSVS_hyper"):
            setattr(s, "cholesky_" + k, int(bool(cpp["cholesky_" + k])))
        s.cholesky_GT_priorkernel = 1 if cpp["cholesky_GT_priorkernel"] == "exponential" else 0
        s.cholesky_SSVS_tau0, s.cholesky_SSVS_tau1 = keep.d(cpp["cholesky_SSVS_tau0"]), keep.d(cpp["cholesky_SSVS_tau1"])
        s.cholesky_SSVS_p = keep.d(cpp["cholesky_SSVS_p"])
        s.cholesky_lambda_3 = (C.c_double * 2)(*cpp["cholesky_lambda_3"])
        s.cholesky_priorhomoscedastic = keep.d(cpp["cholesky_priorhomoscedastic"])
        s.cholesky_sv_offset = keep.d(cpp["cholesky_sv_offset"])
        s.sv_heteroscedastic = keep.i(cpp["sv_heteroscedastic"])
        s.sv_priormu = (C.c_double * 2)(*cpp["sv_priormu"])
        pphi = np.asarray(cpp["sv_priorphi"], float)
        s.sv_priorphi = (C.c_double * 4)(*np.resize(pphi, 4))
        s.sv_priorsigma2 = keep.d(cpp["sv_priorsigma2"])
        s.sv_priorh0 = keep.d(cpp["sv_priorh0"])
    return s


class MultiHandle:
    """n GPUs driven from THIS process (bvb_create_multi): the equation-sharded chain without a launcher.  Pass it to
    Sampler / bvar() in place of a Handle; rank 0's arrays receive the draws."""

    def __init__(self, devices, seed=0):
        from ._lib import load
        self._lib = load()
        self.n = len(devices)
        self.seed = int(seed)
        dv = (C.c_int * self.n)(*[int(d) for d in devices])
        self._hs = (C.c_void_p * self.n)()
        rc = self._lib.bvb_create_multi(self.n, dv, C.c_uint64(self.seed), self._hs)
        if rc:
            from ._lib import BvbError
            raise BvbError(rc, "bvb_create_multi failed (needs %d sm_100 devices and libnccl.so.2)" % self.n)
        self._h = self._hs[0]

    def _check(self, rc):
        if rc:
            from ._lib import BvbError
            last = None
            for i in range(self.n):     # the failing rank carries the message
                step, sweep, eq = C.c_int(), C.c_int(), C.c_int()
                buf = C.create_string_buffer(512)
                self._lib.bvb_last_error(self._hs[i], C.byref(step), C.byref(sweep), C.byref(eq), buf, 512)
                if buf.value:
                    last = BvbError(rc, "rank %d: %s" % (i, buf.value.decode()), step.value, sweep.value, eq.value)
                    break
            raise last or BvbError(rc, "multi-GPU call failed")

    def close(self):
        if getattr(self, "_hs", None) is not None and self.n:
            self._lib.bvb_destroy_multi(self.n, self._hs)
            self.n = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sampler:
    """A chain set up on the device (bvb_gibbs_init); `run(n)` advances it."""

    def __init__(self, handle: Handle, prep, prior_phi, prior_sigma, draws, burnin, thin=1, sv_keep="last",
                 expert_huge=False, out=None):
        self.h = handle
        self.lib = handle._lib
        self.prep = prep
        self.keep = _Keep()
        M, T, K, Kp = prep["M"], prep["T"], prep["K"], prep["Kp"]
        pcpp, scpp = prior_phi["prior_phi_cpp"], prior_sigma["prior_sigma_cpp"]
        r = scpp.get("factor_factors", 0)
        S = M + r
        chol = scpp["type"] == "cholesky"
        n_U = (M * M - M) // 2 if chol else 0
        up = scpp.get("cholesky_U_prior", "normal")
        uh_rows = {"DL": 2 + 2 * n_U, "GT": 2 + 2 * n_U, "HS": 2 + 2 * n_U, "SSVS": 2 * n_U, "HMP": 1}.get(up, 0) if chol and n_U else 0
        self.nsave = draws // thin
        self.draws, self.burnin, self.thin = draws, burnin, thin
        n = K * M
        G = int(pcpp["n_groups"])
        hyper_rows = {"DL": 2 * G + 2 * n, "GT": 2 * G + 2 * n, "HS": 2 * G + 2 * n, "SSVS": 2 * n, "HMP": 2}.get(pcpp["prior"], 0)
        tv = T if sv_keep == "all" else 1
        ns = self.nsave
        self.out = dict(
            PHI=np.zeros((Kp, M, ns), order="F"), V_prior=np.zeros((Kp, M, ns), order="F"),
            logvar=np.zeros((tv, S, ns), order="F"), sv_para=np.zeros((3, S, ns), order="F"),
            phi_hyperparameter=np.zeros((hyper_rows, ns), order="F"),
            facload=np.zeros((M, r, ns), order="F"), fac=np.zeros((r, tv, ns), order="F"),
            U=np.zeros((n_U, ns if chol else 0), order="F"), u_hyperparameter=np.zeros((uh_rows, ns if uh_rows else 0), order="F"))
        if out is not None:   # caller-owned (already allocated / touched) output arrays of exactly these shapes
            for k, a in self.out.items():
                b = out[k]
                if b.shape != a.shape or b.dtype != np.float64 or not (b.flags.f_contiguous or b.size == 0):
                    raise ValueError(f"out[{k!r}]: expected a float64 Fortran-ordered array of shape {a.shape}")
            self.out = {k: out[k] for k in self.out}
        d = Draws()
        for k in ("PHI", "V_prior", "logvar", "sv_para", "phi_hyperparameter", "facload", "fac", "U", "u_hyperparameter"):
            if self.out[k].size:
                setattr(d, k, self.out[k].ctypes.data_as(_dp))
        self._draws = d
        sv = prep["startvals"]
        st = StartVals()
        for k in ("PHI", "U", "sv_para", "sv_logvar", "sv_logvar0", "facload", "fac", "tau2"):
            if k in sv and np.size(sv[k]):
                setattr(st, k, self.keep.d(sv[k]))
        pp = marshal_prior_phi(pcpp, self.keep)
        ps = marshal_prior_sigma(scpp, self.keep)
        multi = isinstance(handle, MultiHandle)
        init = (lambda *a: self.lib.bvb_gibbs_init_multi(handle.n, handle._hs, *a)) if multi else (lambda *a: self.lib.bvb_gibbs_init(self.h._h, *a))
        rc = init(
            self.keep.d(prep["Y"]), self.keep.d(prep["X"]) if Kp else None, M, T, K, draws, burnin, thin,
            1 if sv_keep == "all" else 0, prep["intercept"], self.keep.d(prep["priorIntercept"]) if prep["intercept"] else None,
            self.keep.d(prep["PHI0"]) if Kp else None, C.byref(pp), C.byref(ps), C.byref(st), self.keep.i(prep["i_vec"]) if K else None,
            float(prior_phi["general_settings"]["PHI_tol"]), float(prior_sigma["general_settings"]["cholesky_U_tol"]),
            int(bool(expert_huge)), C.byref(d))
        self.h._check(rc)
        self.done = 0

    def run(self, n_sweeps=None):
        tot = self.draws + self.burnin
        n = tot - self.done if n_sweeps is None else n_sweeps
        done = C.c_int(0)
        if isinstance(self.h, MultiHandle):
            rc = self.lib.bvb_gibbs_run_multi(self.h.n, self.h._hs, int(n), C.byref(done))
        else:
            rc = self.lib.bvb_gibbs_run(self.h._h, int(n), C.byref(done))
        self.h._check(rc)
        self.done = done.value
        return self.done

    def profile_sweep(self):
        """One further sweep in its serial, event-instrumented form; returns ms per phase."""
        out = np.zeros(8)
        done = C.c_int(0)
        self.h._check(self.lib.bvb_gibbs_profile_sweep(self.h._h, fptr(out), C.byref(done)))
        self.done = done.value
        return dict(zip(["prep", "k1", "k2", "comm", "clamp_hyper", "resid", "sv", "store"], out))

    def kernel_times(self, reset=False):
        """Live %globaltimer spans of K1 (Z tiles) and K2 accumulated by the sweeps since the last reset."""
        out = np.zeros(3)
        self.h._check(self.lib.bvb_gibbs_kernel_times(self.h._h, fptr(out), 1 if reset else 0))
        n = int(out[2])
        return dict(sweeps=n, k1_ms=out[0] / n if n else 0.0, k2_ms=out[1] / n if n else 0.0)

    def timing(self):
        out = np.zeros(10)
        n = self.lib.bvb_gibbs_timing(self.h._h, fptr(out))
        names = ["prep", "k1", "k2", "comm", "clamp_hyper", "resid", "sv", "store"]
        return dict(sweeps=n, total_ms=out[0], launches_per_sweep=int(out[1]), phase_ms=dict(zip(names, out[2:10])))

    def state(self):
        p = self.prep
        M, T, Kp = p["M"], p["T"], p["Kp"]
        S = self.out["sv_para"].shape[1]
        r = S - M
        st = dict(PHI=np.zeros((Kp, M), order="F"), V_prior=np.zeros((Kp, M), order="F"),
                  logvar=np.zeros((T, S), order="F"), sv_para=np.zeros((3, S), order="F"),
                  facload=np.zeros((M, r), order="F"), fac=np.zeros((r, T), order="F"), U=np.zeros((M, M), order="F"))
        rc = self.lib.bvb_gibbs_state(self.h._h, *(fptr(st[k]) if st[k].size else None
                                                   for k in ("PHI", "V_prior", "logvar", "sv_para", "facload", "fac")),
                                      fptr(st["U"]) if self.out["U"].size else None)
        self.h._check(rc)
        return st


def bvar(data, lags=1, draws=1000, burnin=1000, thin=1, prior_intercept=10.0, prior_phi=None, prior_sigma=None,
         sv_keep="last", quiet=True, startvals=None, expert_huge=False, handle=None, seed=1, rng=None):
    """Markov Chain Monte Carlo sampling for a Bayesian VAR (R/bvar_wrapper.R:178).

    Returns a dict with the reference's list elements: PHI, U, logvar, sv_para,
    phi_hyperparameter, u_hyperparameter, V_prior, facload, fac (fac permuted to
    (T|1, r, nsave) like bvar_wrapper.R:484), plus Y, X, lags, bench."""
    data = np.asarray(data, float)
    M = data.shape[1]
    prior_phi = prior_phi or specify_prior_phi(M, lags)
    prior_sigma = prior_sigma or specify_prior_sigma(M)
    prep = prepare(data, lags, prior_intercept, prior_phi, prior_sigma, startvals, rng)
    own = handle is None
    h = handle or Handle(0, seed=seed)
    try:
        t0 = time.perf_counter()
        smp = Sampler(h, prep, prior_phi, prior_sigma, draws, burnin, thin, sv_keep, expert_huge)
        tot = draws + burnin
        while smp.done < tot:
            smp.run(256)   # the shim's interrupt / progress cadence (bvar_cpp.cpp:501-503)
            if not quiet:
                print(f"\r###  {smp.done} / {tot} ### ({100.0 * smp.done / tot:3.0f}%) ###", end="")
        bench = time.perf_counter() - t0
        res = dict(smp.out)
        res["fac"] = np.transpose(res["fac"], (1, 0, 2))
        scpp = prior_sigma["prior_sigma_cpp"]
        res.update(Y=prep["Y"], X=prep["X"], lags=lags, bench=bench, timing=smp.timing(),
                   prior_phi=prior_phi, prior_sigma=prior_sigma, Yraw=data, intercept=bool(prep["intercept"]),
                   sigma_type=scpp["type"], heteroscedastic=np.asarray(scpp["sv_heteroscedastic"], bool),
                   datamat=np.hstack([prep["Y"], prep["X"]]))   # bvar_wrapper.R:489-493
        return res
    finally:
        if own:
            h.close()
