"""A few sweeps at the configs[1] shape (M=50, p=4, T=300, 6 factors, HS) - used under ncu."""
import sys
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
from bench import synth_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
M, p, T = 50, 4, 300
data = synth_data(M, p, T)
pp = B.specify_prior_phi(M, p, "HS")
ps = B.specify_prior_sigma(M, type="factor", factor_factors=6)
prep = B.prepare(data, p, 10.0, pp, ps, rng=np.random.default_rng(1))
h = Handle(0, seed=3)
smp = B.Sampler(h, prep, pp, ps, draws=n, burnin=0)
smp.run()
print("ok", smp.done, smp.timing())
