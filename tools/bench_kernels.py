"""Kernel-only timings of the per-draw kernels (K6 IRF, K7 predict, GIG) at BASELINE config shapes, with the
achieved HBM GB/s the north star asks for (algorithmic bytes / device time between CUDA events around the kernel).
usage: python tools/bench_kernels.py [irf] [gig] [predict] [--draws R]
"""
import sys, json, time
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle

args = [a for a in sys.argv[1:] if a in ("irf", "gig", "predict", "vcov")]
which = args or ["irf", "gig", "predict", "vcov"]
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

if "predict" in which:
    # K7 at the config-5 shape with every SM busy (2 resident chains per SM): FP64 tensor roofline
    M, p, nf, ahead = 200, 4, 1, np.arange(1, 13)
    Rp = 296 if "--draws" not in sys.argv else R
    Kp = M * p + 1
    PHI = np.asfortranarray(rng.standard_normal((Kp, M, Rp)) * 0.3 / np.sqrt(M * p))
    fl = rng.standard_normal((M, nf, Rp))
    lv = rng.normal(-1.0, 0.3, (M + nf, Rp))
    nsv = M + nf
    mu = rng.normal(-1.0, 0.3, (nsv, Rp)); ph = rng.uniform(0.8, 0.98, (nsv, Rp)); sg = rng.uniform(0.05, 0.3, (nsv, Rp))
    x = np.concatenate([rng.standard_normal(M * p), [1.0]])
    yobs = rng.standard_normal((ahead.size, M))
    t0 = time.perf_counter()
    out = h.out_of_sample(1, x, PHI, None, fl, lv, ahead, mu, ph, sg, np.arange(nsv), True, True, yobs, False, None, True,
                          None, None)
    wall = time.perf_counter() - t0
    ms = h.last_timing()["rest_ms"]
    peak = h.measure_fp64_peak()
    dmma = float(peak["dmma_tflops"] if isinstance(peak, dict) else peak[0])
    # algorithmic flops per chain: horizons i >= 1: PHI' S (2 M mn^2) + (PHI' S) PHI (2 M^2 mn), mn = min(i, p) M; Cholesky M^3/3
    fl_chain = 0.0
    for i in range(int(ahead.max())):
        mn = min(i, p) * M
        fl_chain += 2.0 * M * mn * mn + 2.0 * M * M * mn + M ** 3 / 3.0
    tf = fl_chain * Rp / (ms * 1e-3) / 1e12
    print(json.dumps(dict(kernel="K7 predict", case="configs[4] M=200 p=4 r=1 ahead=1:12, LPL + predictive draws", draws=Rp,
                          dev_ms=round(ms, 2), ms_per_draw=round(ms / Rp, 4), algorithmic_GFLOP_per_draw=round(fl_chain / 1e9, 3),
                          achieved_TFLOPs=round(tf, 2), dmma_peak_TFLOPs=round(dmma, 1), frac=round(tf / dmma, 4),
                          wall_s=round(wall, 3), finite=bool(np.isfinite(out["predictions"]).all()))), flush=True)

if "vcov" in which:
    # K9 vcov at the config-3 shape: output-write bound (T * M * M doubles per draw)
    T, M, nf, Rv = 500, 100, 1, 40
    fl = rng.standard_normal((M, nf, Rv))
    lv = rng.normal(-1.0, 0.3, (T, M + nf, Rv))
    best = 1e30
    for _ in range(2):
        out = h.vcov_cpp(True, fl, lv, None, M, nf)
        best = min(best, h.last_timing()["rest_ms"])
    gb = out.nbytes / 1e9
    print(json.dumps(dict(kernel="K9 vcov", case="configs[2] T=500 M=100 r=1 (factor)", draws=Rv, dev_ms=round(best, 3),
                          algorithmic_GB_written=round(gb, 3), achieved_GBps=round(gb / (best * 1e-3), 1),
                          hbm_peak_GBps=hbm_peak, frac=round(gb / (best * 1e-3) / hbm_peak, 4),
                          finite=bool(np.isfinite(out).all()))), flush=True)
