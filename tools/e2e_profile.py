"""Where does a fresh K-sweep chain through the host API spend its wall time?  (cProfile of Sampler() + run())
usage: python tools/e2e_profile.py [K]"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
import bench
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
data, pp, ps, prep = bench.make_problem()
h = Handle(0, seed=1)
def chain(out=None):
    smp = B.Sampler(h, prep, pp, ps, draws=K, burnin=0, thin=1, sv_keep="last", out=out)
    while smp.done < K:
        smp.run(256)
    return smp
s0 = chain()
out = s0.out
for _ in range(2):
    chain(out)
t0 = time.perf_counter(); chain(out); t1 = time.perf_counter()
print("one chain: %.3f ms" % (1e3 * (t1 - t0)))
pr = cProfile.Profile(); pr.enable(); chain(out); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
