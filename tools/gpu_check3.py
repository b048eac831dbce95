"""First full-sweep run on the GPU: small model sanity + config-3 timing."""
import sys, json, time
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B

def sim_var(M, p, T, seed=1):
    rng = np.random.default_rng(seed)
    A = 0.5 * np.eye(M) + 0.05 * rng.standard_normal((M, M)) * (rng.uniform(size=(M, M)) < 0.1)
    Y = np.zeros((T + p + 50, M))
    lam = 0.5 * rng.standard_normal(M)
    for t in range(1, Y.shape[0]):
        f = rng.standard_normal() * np.exp(0.3 * np.sin(t / 30.0))
        Y[t] = Y[t - 1] @ A.T + lam * f + 0.5 * rng.standard_normal(M)
    return Y[50:], A

if __name__ == '__main__':
  h = Handle(0, seed=11)
  for (M, p, T, prior, draws, burn) in [(4, 1, 200, "HS", 2000, 1000), (6, 2, 150, "DL", 1000, 500), (6, 2, 150, "R2D2", 1000, 500),
                                        (6, 2, 150, "NG", 1000, 500), (6, 2, 150, "SSVS", 1000, 500), (6, 2, 150, "HMP", 500, 500),
                                        (6, 2, 150, "normal", 500, 500), (100, 4, 500, "DL", 200, 56)]:
      Y, A = sim_var(M, p, T)
      pp = B.specify_prior_phi(M, p, prior)
      if prior == "HMP":
          pp["prior_phi_cpp"]["V_i_prep"] = np.ones((p * M + 1) * M)
      if prior == "SSVS":
          pp["general_settings"]["SSVS_semiautomatic"] = False
      ps = B.specify_prior_sigma(M, factor_factors=1)
      t0 = time.time()
      try:
          res = B.bvar(Y, lags=p, draws=draws, burnin=burn, prior_phi=pp, prior_sigma=ps, handle=h, sv_keep="last")
      except Exception as e:
          print(f"M={M} {prior}: FAILED {e!r}", flush=True)
          continue
      dt = time.time() - t0
      phi_mean = res["PHI"].mean(axis=2)
      err = np.abs(phi_mean[:M, :] - A.T).max()
      tm = res["timing"]
      print(f"M={M} p={p} T={T} {prior}: {draws+burn} sweeps in {dt:.2f}s ({(draws+burn)/dt:.0f}/s) finite={np.isfinite(res['PHI']).all()} "
            f"max|E[PHI_1]-A'|={err:.3f} sv_phi_mean={res['sv_para'][1].mean():.3f} sigma_mean={res['sv_para'][2].mean():.3f} "
            f"launches={tm['launches_per_sweep']} phases_us={ {k: round(v*1e3,1) for k,v in tm['phase_ms'].items()} }", flush=True)
