"""Kernel-only timings of the per-draw kernels (K6 IRF, K7 predict, GIG) at BASELINE config shapes, with the
achieved HBM GB/s the north star asks for (algorithmic bytes / device time between CUDA events around the kernel).
usage: python tools/bench_kernels.py [irf] [gig] [predict] [--draws R]
"""
import sys, json, time
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle

args = [a for a in sys.argv[1:] if a in ("irf", "gig", "predict")]
which = args or ["irf", "gig", "predict"]
R = int(sys.argv[sys.argv.index("--draws") + 1]) if "--draws" in sys.argv else 1000
try:
    peaks = json.load(open("MEASURED_PEAKS.json"))
    hbm_peak = float(peaks.get("hbm_gbs", 6553.6))
except Exception:
    hbm_peak = 6553.6
h = Handle(0, seed=11)
rng = np.random.default_rng(5)


def irf_case(name, M, p, r, R, ahead, reps=3):
    Kp = M * p + 1
    PHI = np.asfortranarray(rng.standard_normal((Kp, M, R)) * 0.02)
    fl = np.asfortranarray(rng.standard_normal((M, r, R)))
    lv = np.asfortranarray(rng.standard_normal((M + r, R)) * 0.1)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = h.irf_cpp(PHI, fl, None, lv, np.eye(r), ahead)
        wall = time.perf_counter() - t0
        best = min(best, h.last_timing()["rest_ms"])
    gb = PHI.nbytes / 1e9
    print(json.dumps(dict(kernel="K6 irf", case=name, draws=R, ahead=ahead, shocks=r, dev_ms=round(best, 3),
                          us_per_draw=round(1e3 * best / R, 3), algorithmic_GB=round(gb, 3),
                          achieved_GBps=round(gb / (best * 1e-3), 1), hbm_peak_GBps=hbm_peak,
                          frac=round(gb / (best * 1e-3) / hbm_peak, 4), wall_s_last=round(wall, 3),
                          finite=bool(np.isfinite(out).all()))), flush=True)


if "irf" in which:
    irf_case("configs[4] M=200 p=4 r=1", 200, 4, 1, R, 12)
    irf_case("configs[2] M=100 p=4 r=1", 100, 4, 1, 4 * R, 12)
    irf_case("configs[1] M=50 p=4 r=6", 50, 4, 6, 8 * R, 12)

if "gig" in which:
    n = 4_000_000
    lam = np.full(n, -0.5); chi = rng.gamma(2.0, 1.0, n) * 1e-3 + 1e-8; psi = np.full(n, 2.0)
    for _ in range(3):
        out = h.my_gig(1, lam, chi, psi)
        ms = h.last_timing()["rest_ms"]
    gb = 4 * 8 * n / 1e9   # three parameters in, one variate out
    print(json.dumps(dict(kernel="bvb_gig", n=n, dev_ms=round(ms, 3), variates_per_us=round(n / ms / 1e3, 1),
                          achieved_GBps=round(gb / (ms * 1e-3), 1), hbm_peak_GBps=hbm_peak,
                          frac=round(gb / (ms * 1e-3) / hbm_peak, 4))), flush=True)
