import sys
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle, BvbError
from bayesianvars_b200 import bvar as B
from tools.gpu_check3 import sim_var

h = Handle(0, seed=11)
for prior, chunk in [("DL", 256), ("DL", 64), ("DL", 7), ("NG", 256)]:
    M, p, T = 6, 2, 150
    Y, A = sim_var(M, p, T)
    pp = B.specify_prior_phi(M, p, prior)
    ps = B.specify_prior_sigma(M, factor_factors=1)
    prep = B.prepare(Y, p, 10.0, pp, ps)
    smp = B.Sampler(h, prep, pp, ps, draws=1000, burnin=500)
    try:
        while smp.done < 1500:
            smp.run(chunk)
        print(prior, chunk, "ok", np.isfinite(smp.out["PHI"]).all(), smp.out["PHI"][:, :, -1].ravel()[:3])
    except BvbError as e:
        print(prior, chunk, "failed:", e, "done", smp.done)
        st = smp.state()
        V = st["V_prior"]
        print(" V_prior: min", np.nanmin(V), "max", np.nanmax(V), "nan", np.isnan(V).sum(), "inf", np.isinf(V).sum(), "zero", (V == 0).sum())
        print(" PHI finite", np.isfinite(st["PHI"]).all(), "absmax", np.nanmax(np.abs(st["PHI"])), "logvar range", np.nanmin(st["logvar"]), np.nanmax(st["logvar"]), np.isfinite(st["logvar"]).all())
        print(" sv_para", st["sv_para"][:, :7].round(3).tolist(), "facload", st["facload"].ravel().round(3).tolist())
