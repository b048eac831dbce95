"""Where does the end-to-end time of a fresh chain go?  (GPU box only.)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h = Handle(0, seed=1)
data, pp, ps, prep = bench.make_problem()
for rep in range(2):
    t0 = time.perf_counter()
    smp = B.Sampler(h, prep, pp, ps, draws=K, burnin=0, thin=1, sv_keep="last")
    t1 = time.perf_counter()
    marks = []
    while smp.done < K:
        ta = time.perf_counter()
        smp.run(256)
        marks.append(time.perf_counter() - ta)
    t2 = time.perf_counter()
    print(f"rep {rep}: ctor {1e3*(t1-t0):.1f} ms, runs {[round(1e3*m,1) for m in marks]} ms, total {1e3*(t2-t0):.1f} ms -> {K/(t2-t0):.1f} sweeps/s; dev {smp.timing()}")
    del smp
