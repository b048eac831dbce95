"""K1 / K2 timings of one conditional PHI draw (bvb_phi_draw_factor) at the headline system size (Kp = 401) for
several equation counts (100 = one GPU, 50 / 25 / 13 = the shards of 2 / 4 / 8 GPUs), with the parity error against
the oracle.  BVB_K2_TILED=0 selects the one-CTA-per-equation kernels.   usage: python tools/bench_k2.py [M ...]"""
import json
import os
import sys
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from oracle import regression as orc
from tools.synth import phi_draw_inputs

Ms = [int(a) for a in sys.argv[1:] if a.isdigit()] or [100, 50, 25, 13]
p, T = 4, 500
h = Handle(0, seed=3)
peak = h.measure_fp64_peak()["dmma_tflops"]
for M in Ms:
    # M equations of the 100-variable system: the regressors keep Kp = 401
    d = phi_draw_inputs(100, p, T, 1, seed=7)
    sel = slice(0, M)
    Y, lv, V, P0, z = (np.asfortranarray(d[k][:, sel]) for k in ("Y", "logvar", "V", "PHI0", "z"))
    fl = np.asfortranarray(d["facload"][sel, :])
    best = dict(k1_ms=1e9, k2_ms=1e9)
    for _ in range(6):
        out = h.sample_PHI_factor(P0, Y, d["X"], lv, V, fl, d["fac"], z=z)
        t = h.last_timing()
        best = {k: min(best[k], t[k]) for k in best}
    ref = orc.sample_PHI_factor(P0, P0, Y, d["X"], lv, V, fl, d["fac"], False, z)
    err = float((np.max(np.abs(out - ref), axis=0) / np.max(np.abs(ref), axis=0)).max())
    Kp = d["X"].shape[1]
    gf2 = M * (Kp ** 3 / 3 + 2 * Kp ** 2) / 1e9
    gf1 = M * (T * Kp * (Kp + 1) + 2 * T * Kp) / 1e9
    print(json.dumps(dict(M=M, Kp=Kp, tiled=os.environ.get("BVB_K2_TILED", "1"), k1_us=round(1e3 * best["k1_ms"], 1),
                          k2_us=round(1e3 * best["k2_ms"], 1), k2_tflops=round(gf2 / best["k2_ms"], 2),
                          k2_frac=round(gf2 / best["k2_ms"] / peak, 3), k1_frac=round(gf1 / best["k1_ms"] / peak, 3),
                          rel_err=err)), flush=True)
