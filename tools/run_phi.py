"""Run the factor PHI draw (K1+K2) a few times at one shape; used under ncu."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from tools.synth import phi_draw_inputs

M, p, T, r = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (100, 4, 500, 1)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
d = phi_draw_inputs(M, p, T, r)
h = Handle(0, seed=1)
for i in range(reps):
    out = h.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], z=d["z"])
    print(json.dumps(h.last_timing()), float(np.abs(out).max()))
