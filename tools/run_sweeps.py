"""Run a few full Gibbs sweeps at the bench workload (used under ncu)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
data, pp, ps, prep = bench.make_problem()
h = Handle(0, seed=2024)
smp = B.Sampler(h, prep, pp, ps, draws=n, burnin=0)
smp.run()
print("ok", smp.done, float(np.abs(smp.out["PHI"]).max()), smp.timing())
