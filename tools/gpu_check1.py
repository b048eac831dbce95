"""First GPU trip: FP64 peaks, K1+K2 parity vs the oracle, first timings."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from bayesianvars_b200 import Handle
from oracle import regression as orc
from tools.synth import phi_draw_inputs

h = Handle(0, seed=7)
print("peak", json.dumps(h.measure_fp64_peak()), flush=True)

def relerr(a, b):
    return float(np.max(np.abs(a - b), axis=0).max() / max(np.abs(b).max(), 1e-300)), \
           float((np.max(np.abs(a - b), axis=0) / np.maximum(np.max(np.abs(b), axis=0), 1e-300)).max())

for (M, p, T, r, prior) in [(3, 1, 40, 0, "flat"), (10, 2, 235, 1, "dl"), (14, 3, 100, 2, "dl"),
                            (50, 4, 300, 6, "dl"), (100, 4, 500, 1, "dl")]:
    d = phi_draw_inputs(M, p, T, r, prior=prior)
    t0 = time.time()
    ref = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                d["facload"], d["fac"], False, d["z"])
    t_cpu = time.time() - t0
    try:
        out = h.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], z=d["z"])
        out = h.sample_PHI_factor(d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"], d["facload"], d["fac"], z=d["z"])
        e_all, e_eq = relerr(out, ref)
        print(f"M={M} p={p} T={T} r={r} Kp={d['X'].shape[1]}: relerr(global)={e_all:.3e} relerr(per-eq max)={e_eq:.3e} "
              f"cpu_oracle={t_cpu*1e3:.1f}ms timing={json.dumps(h.last_timing())}", flush=True)
    except Exception as ex:
        print(f"M={M} p={p} T={T}: FAILED {ex!r}", flush=True)
