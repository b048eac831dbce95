"""Host-side mirror of `predict.bayesianVARs_bvar` (R/utility_functions.R:2406-2530) and
`stable_bvar` (R/bvar_wrapper.R:582-640) for the dict returned by `bvar()`.  The per-draw
work (`out_of_sample`, src/utilities_cpp.cpp:216-393) runs on the device through
`bvb_predict`; this module only gathers its arguments and post-processes the draws the way
the R method does."""
from __future__ import annotations

import numpy as np

from .api import Handle

_PER_DRAW = ("PHI", "U", "logvar", "sv_para", "phi_hyperparameter", "u_hyperparameter", "V_prior", "facload", "fac")


def get_companion(PHI, intercept=None):
    """Companion matrix of one coefficient draw (lags*M [+1] x M)."""
    Kp, M = PHI.shape
    p = Kp // M
    K = p * M
    F = np.zeros((K, K))
    F[:, :M] = PHI[:K]
    if p > 1:
        F[:K - M, M:] = np.eye(K - M)
    return F.T


def stable_bvar(mod, quiet=False):
    """Keep the draws whose companion matrix has spectral radius < 1 (bvar_wrapper.R:590-627)."""
    PHI = mod["PHI"]
    draws = PHI.shape[2]
    keep = np.array([np.max(np.abs(np.linalg.eigvals(get_companion(PHI[:, :, i])))) < 1 for i in range(draws)], bool)
    if not quiet:
        print(f"\nOriginal 'bayesianVARs_bvar' object consists of {draws} posterior draws.\n")
        print(f"\nDetected {int((~keep).sum())} unstable draws.\n")
    out = dict(mod)
    for k in _PER_DRAW:
        a = mod.get(k)
        if a is not None and a.ndim and a.shape[-1] == draws:
            out[k] = np.asfortranarray(a[..., keep])
    out["stable"] = True
    return out


def _logmeanexp_rows(a, nan_ok=False):
    mx = (np.nanmax if nan_ok else np.max)(a, axis=1) - 700.0
    mean = (np.nanmean if nan_ok else np.mean)(np.exp(a - mx[:, None]), axis=1)
    return np.log(mean) + mx


def predict(mod, ahead=1, each=1, stable=True, simulate_predictive=True, LPL=False, Y_obs=None, LPL_VoI=None,
            handle=None, seed=1, variables=None):
    """predict.bayesianVARs_bvar: predictive draws and log predictive likelihoods."""
    if not simulate_predictive and not LPL:
        raise ValueError("At least one of 'simulate_predictive' or 'LPL' must be set 'TRUE'!")
    ahead = np.sort(np.atleast_1d(np.asarray(ahead, int)))
    factor = mod["sigma_type"] == "factor"
    if stable and not mod.get("stable", False):
        mod = stable_bvar(mod, quiet=True)
    sv_indicator = np.flatnonzero(np.asarray(mod["heteroscedastic"], bool))
    M = mod["Y"].shape[1]
    sv_para = mod["sv_para"]
    sv_mu, sv_phi, sv_sigma = (np.asfortranarray(sv_para[k, sv_indicator, :]) for k in range(3))
    sv_h_T = mod["logvar"][-1, :, :]
    x = np.asarray(mod["datamat"][-1, :mod["lags"] * M], float)
    if mod["intercept"]:
        x = np.concatenate([x, [1.0]])
    LPL_subset, VoI = False, np.zeros(1, int)
    if LPL:
        if LPL_VoI is not None:
            LPL_subset = True
            if all(isinstance(v, str) for v in LPL_VoI):
                if variables is None:
                    raise ValueError("Cannot find variables of interest specified in 'LPL_VoI' in the data! \n")
                VoI = np.array([i for i, v in enumerate(variables) if v in set(LPL_VoI)], int)
                if VoI.size != len(LPL_VoI):
                    raise ValueError("Cannot find variables of interest specified in 'LPL_VoI' in the data! \n")
            else:
                VoI = np.asarray(LPL_VoI, int) - 1          # R indices are 1-based (:2484)
        Y_obs = np.asarray(Y_obs, float).reshape(ahead.size, M)
    own = handle is None
    h = handle or Handle(0, seed=seed)
    try:
        ret = h.out_of_sample(each, x, mod["PHI"], mod.get("U"), mod.get("facload"), sv_h_T, ahead, sv_mu, sv_phi, sv_sigma,
                              sv_indicator, factor, LPL, Y_obs, LPL_subset, VoI, simulate_predictive)
    finally:
        if own:
            h.close()
    if LPL:
        ret["LPL"] = _logmeanexp_rows(ret["LPL_draws"])
        ret["LPL_univariate"] = np.log(ret["PL_univariate_draws"].mean(axis=2))
        if LPL_subset:
            sub = ret["LPL_sub_draws"]
            ret["LPL_VoI"] = _logmeanexp_rows(sub, nan_ok=bool(np.isnan(sub).any()))
            ret["VoI"] = VoI
    else:
        for k in ("LPL_draws", "PL_univariate_draws", "LPL_sub_draws"):
            ret.pop(k, None)
    if not simulate_predictive:
        ret.pop("predictions", None)
    ret["Y_obs"] = Y_obs
    ret["Yraw"] = mod.get("Yraw")
    ret["ahead"] = ahead
    return ret
