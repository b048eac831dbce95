import sys
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
from bench import synth_data
h = Handle(0, seed=7)
for (M, p, T, phi, up) in [(10, 2, 200, "HMP", "HMP"), (30, 4, 500, "HMP", "HMP"), (100, 4, 500, "HMP", "HS"), (100, 4, 500, "HS", "HMP"), (100, 4, 500, "HMP", "HMP"), (100, 1, 500, "HMP", "HMP")]:
    data = synth_data(M, p, T)
    pp = B.specify_prior_phi(M, p, phi)
    ps = B.specify_prior_sigma(M, type="cholesky", cholesky_U_prior=up)
    prep = B.prepare(data, p, 10.0, pp, ps, rng=np.random.default_rng(1))
    try:
        smp = B.Sampler(h, prep, pp, ps, draws=10, burnin=0)
        smp.run()
        print(M, p, T, phi, up, "ok", float(np.abs(smp.out["PHI"]).max()), smp.out["phi_hyperparameter"][:, -1][:4], flush=True)
    except Exception as e:
        print(M, p, T, phi, up, "FAIL", str(e)[:150], flush=True)
    V = prep.get("V_i", None)
print({k: (np.min(v), np.max(v)) for k, v in prep["startvals"].items() if np.size(v)})
