"""K1/K2 variant timing at config 3 + correctness spot check (env BVB_K1_MW)."""
import json, sys, os
import numpy as np
sys.path.insert(0, ".")
from bayesianvars_b200 import Handle
from oracle import regression as orc
from tools.synth import phi_draw_inputs
h = Handle(0, seed=7)
for (M, p, T, r) in [(14, 3, 100, 2), (50, 4, 300, 6), (100, 4, 500, 1)]:
    d = phi_draw_inputs(M, p, T, r)
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], False, d["z"])
    best = None
    for i in range(5):
        out = h.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], z=d["z"])
        t = h.last_timing()
        if best is None or t["k1_ms"] + t["k2_ms"] < best["k1_ms"] + best["k2_ms"]: best = t
    err = (np.max(np.abs(out - ref), axis=0) / np.max(np.abs(ref), axis=0)).max()
    print(f"MW={os.environ.get('BVB_K1_MW','4')} M={M} Kp={d['X'].shape[1]} T={T}: err={err:.2e} k1={best['k1_ms']*1e3:.1f}us k2={best['k2_ms']*1e3:.1f}us", flush=True)
