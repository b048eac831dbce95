#!/usr/bin/env python
"""bench.py - Gibbs draws/sec of the bayesianVARs hot path on B200.

One "step" = one full Gibbs sweep of bvar_cpp()'s loop body
(src/bvar_cpp.cpp:498-845): PHI draw (K1 pair-GEMM + K2 Cholesky/solve/draw),
tolerance clamp, hyper-parameter update, residuals, factor-SV update, storage.
Workload at N=1: BASELINE.json configs[2] = synthetic M=100, p=4, T=500,
factor SV (r=1), Dirichlet-Laplace prior - the shape `metric` is quoted on.

    python bench.py --gpus N --steps K --warmup W          # this repo (CUDA)
    python bench.py --impl reference --steps K --warmup W  # CPU restatement of the reference

value  = sweeps/s with the chain resident in HBM: W warm-up sweeps, then REPS back-to-back regions of exactly K
         sweeps each, every region bracketed by barrier + synchronize and timed with CUDA events on the sampler
         stream (max over ranks); `value` is K / median region time, all repetitions are in `repetitions_ms`.
         Nothing but graph-replayed sweeps runs inside a region (no instrumented sweep, no subprocess).
e2e    = sweeps/s through the public host API (bvar()-equivalent: host buffers in, caller-owned draw arrays out;
         H2D of the data + per-sweep D2H of the stored draw inside the timed region), median of 3 fresh chains.
roofline = the longer of K1 / K2, from their live %globaltimer spans accumulated over the timed regions.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(M=100, p=4, T=500, r=1, prior="DL")
METRIC = "Gibbs draws/sec (M=100,p=4,T=500 factor-SV, DL)"
REPS = 7


def workload_config(n_gpus):
    """The `config` object - identical keys and values for both arms."""
    w = WORKLOAD
    return dict(workload=f"M={w['M']} p={w['p']} T={w['T']} factor-SV r={w['r']} prior={w['prior']} (BASELINE configs[2])",
                n_gpus=n_gpus, every_sweep_stored=True,
                l2="state per sweep (Z 322 MB + packed systems 65 MB) exceeds the 126 MB L2; no flush needed")


def synth_data(M, p, T, seed=123):
    """Reference benchmark protocol (vignettes/scalability-vignette.qmd:88-125): iid N(0,1) data."""
    return np.random.default_rng(seed).standard_normal((T + p, M))


def algorithmic_gflop(M, Kp, T, r):
    """BASELINE.md section 2 / SURVEY.md section 8(d): per-sweep algorithmic work (GFLOP)."""
    syrk = M * T * Kp * (Kp + 1)
    rhs = 2 * M * T * Kp
    chol = M * Kp ** 3 / 3
    solves = 2 * M * Kp ** 2
    resid = 2 * T * Kp * M
    yhat = 2 * T * r * M
    return dict(k1=(syrk + rhs) / 1e9, k2=(chol + solves) / 1e9, total=(syrk + rhs + chol + solves + resid + yhat) / 1e9)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons through NVML inside this process (no fork): samples continuously, keeps only the
    samples taken while `active` is set (the timed regions)."""

    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.active = threading.Event()
        self._halt = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        nv = self.nvml
        while not self._halt.is_set():
            if self.active.is_set():
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                    self.samples.append(mhz)
                    for nm, bit in self.REASONS.items():
                        if bits & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
                time.sleep(0.0005)
            else:
                time.sleep(0.0002)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples), how="NVML in-process, timed regions only")


def trace(msg):
    """Stage markers on stderr (BENCH_TRACE=1): where a multi-rank run is when it is killed."""
    if os.environ.get("BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def make_problem():
    from bayesianvars_b200 import bvar as B
    w = WORKLOAD
    data = synth_data(w["M"], w["p"], w["T"])
    pp = B.specify_prior_phi(w["M"], w["p"], w["prior"], DL_a="1/K")
    ps = B.specify_prior_sigma(w["M"], type="factor", factor_factors=w["r"])
    prep = B.prepare(data, w["p"], 10.0, pp, ps, rng=np.random.default_rng(123))
    return data, pp, ps, prep


def cpu_reference_rate(prep, pp, ps, sweeps, warm=1):
    """The oracle port of the reference sweep (oracle.gibbs), timed on the host cores."""
    from oracle import gibbs as og
    rng = np.random.default_rng(5)
    st = og.init_state(prep, pp, ps)
    for rep in range(warm):
        og.sweep(st, prep, pp, ps, rep, 0, rng)
    t0 = time.perf_counter()
    for rep in range(sweeps):
        og.sweep(st, prep, pp, ps, warm + rep, 0, rng)
    return sweeps / (time.perf_counter() - t0)


def vignette_anchor():
    """The reference's own published timing (vignettes/huge_all_combinations.RData row 24: M = 50, n = 100, p = 4, factor
    r = 1, SV, default HS prior: 1.69 s user time for bvar(draws = 10, burnin = 0) on one core of an i7-10610U),
    re-run with the oracle port on ONE thread of this host: the same protocol (iid N(0,1) data, seed 123)."""
    from threadpoolctl import threadpool_limits
    from bayesianvars_b200 import bvar as B
    from oracle import gibbs as og
    M, p, n = 50, 4, 100
    data = np.random.default_rng(123).standard_normal((n + p, M))
    pp = B.specify_prior_phi(M, p, "HS")
    ps = B.specify_prior_sigma(M, type="factor", factor_factors=1)
    prep = B.prepare(data, p, 10.0, pp, ps, rng=np.random.default_rng(123))
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        og.run(prep, pp, ps, draws=10, burnin=0, rng=np.random.default_rng(1))
        sec = time.perf_counter() - t0
    return dict(shape="M=50 n=100 p=4 factor r=1 SV, HS prior, 10 sweeps", port_seconds_1thread=sec,
                published_seconds=1.69, published_on="i7-10610U, one core, R (vignettes/huge_all_combinations.RData row 24)",
                ratio_port_over_published=sec / 1.69)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baselines(prep, pp, ps, n_all=12, n_one=4):
    """All-core and single-thread figures of the oracle port, plus the vignette anchor."""
    from threadpoolctl import threadpool_limits
    cores = blas_threads()
    rate = cpu_reference_rate(prep, pp, ps, n_all, 1)
    with threadpool_limits(limits=1):
        rate1 = cpu_reference_rate(prep, pp, ps, n_one, 1)
    cpu = dict(value=rate, unit="sweeps/s", cores=cores, kind="port",
               sample=f"{n_all} full sweeps of oracle.gibbs on the same inputs (NumPy/SciPy-LAPACK restatement of bvar_cpp, "
                      f"OpenBLAS {cores} threads); the R reference cannot run here")
    cpu1 = dict(value=rate1, unit="sweeps/s", cores=1, kind="port",
                sample=f"{n_one} full sweeps of oracle.gibbs with the BLAS pool limited to one thread (the vignette's 'one core' setting)")
    return cpu, cpu1


def draw_sharded_postprocessing(h, rank, world, barrier, max_over_ranks):
    """irf() and predict() over stored posterior draws, the draws sharded over the ranks
    (contiguous ranges, disjoint output slices, no collective).  Shape: BASELINE configs[4] (M = 200, p = 4, Kp = 801,
    one factor), irf(ahead = 12) over 4096 draws and predict(ahead = 1:12) over 592 draws in total whatever N is
    (strong scaling).  Host arrays in and out (the C-ABI copies them), wall clock between barriers, max over ranks."""
    from bayesianvars_b200 import bvar as B
    M, p, r = 200, 4, 1
    Kp = M * p + 1
    out = {}
    rng = np.random.default_rng(99)
    base = 32
    PHIb = np.asfortranarray(rng.standard_normal((Kp, M, base)) * 0.3 / np.sqrt(M * p))
    flb = np.asfortranarray(rng.standard_normal((M, r, base)))
    lvb = np.asfortranarray(rng.normal(-1.0, 0.3, (M + r, base)))
    for name, R_total in (("irf", 4096), ("predict", 592)):
        r0, cnt = B.shard_draws(R_total, rank, world)
        reps = -(-max(cnt, 1) // base)
        PHI = np.asfortranarray(np.tile(PHIb, (1, 1, reps))[:, :, :max(cnt, 1)])
        fl = np.asfortranarray(np.tile(flb, (1, 1, reps))[:, :, :max(cnt, 1)])
        lv = np.asfortranarray(np.tile(lvb, (1, reps))[:, :max(cnt, 1)])
        if name == "irf":
            run = lambda: h.irf_cpp(PHI, fl, None, lv, np.eye(r), 12)
        else:
            nsv = M + r
            mu = np.full((nsv, PHI.shape[2]), -1.0); ph = np.full_like(mu, 0.9); sg = np.full_like(mu, 0.2)
            x = np.concatenate([rng.standard_normal(M * p), [1.0]])
            yobs = rng.standard_normal((12, M))
            run = lambda: h.out_of_sample(1, x, PHI, None, fl, lv, np.arange(1, 13), mu, ph, sg, np.arange(nsv), True, True,
                                          yobs, False, None, True, None, None)
        if cnt > 0:
            run()                      # warm-up (workspaces, kernel attributes)
        barrier()
        t0 = time.perf_counter()
        res = run() if cnt > 0 else None
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0)
        dev_ms = max_over_ranks(h.last_timing()["rest_ms"] if cnt > 0 else 0.0)
        ok = True if res is None else bool(np.isfinite(res if name == "irf" else res["predictions"]).all())
        out[name] = dict(draws_total=R_total, draws_this_rank=cnt, wall_s=sec, draws_per_s=R_total / sec,
                         kernel_ms_max_over_ranks=dev_ms, finite=ok)
        del PHI, fl, lv, res
    out["shape"] = "BASELINE configs[4]: M=200 p=4 Kp=801 r=1, irf ahead=12 / predict ahead=1:12, LPL + predictive draws"
    out["scaling"] = "strong (total draws fixed), draws sharded in contiguous ranges, no collective"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, pp, ps, prep = make_problem()
    cores = blas_threads()
    # each "step" = one full CPU sweep (~0.4 s at this shape on 16 threads): K and W are honoured as given
    steps, warm = args.steps, args.warmup
    rate = cpu_reference_rate(prep, pp, ps, steps, warm)
    line = dict(metric=METRIC, value=rate, unit="sweeps/s", n_gpus=args.gpus,
                steps=steps, warmup=warm, ms_per_step=1e3 / rate, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=workload_config(args.gpus),
                cpu_baseline=dict(value=rate, unit="sweeps/s", cores=cores, kind="port",
                                  sample=f"{steps} full sweeps of oracle.gibbs (NumPy/SciPy-LAPACK restatement of bvar_cpp, "
                                         f"OpenBLAS {cores} threads); the R reference cannot run here"),
                e2e=dict(value=rate, unit="sweeps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from bayesianvars_b200 import Handle
    from bayesianvars_b200 import bvar as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    data, pp, ps, prep = make_problem()
    w = WORKLOAD
    Kp = prep["Kp"]
    W, K = args.warmup, args.steps

    # ---------------- N > 1: a short fixed-seed chain on ONE GPU first (rank 0), to compare the sharded chain with
    parity_ref = None
    PAR_SWEEPS = 12
    trace("problem ready")
    if world > 1 and rank == 0:
        with Handle(local_rank, seed=4242) as h1:
            s1 = B.Sampler(h1, prep, pp, ps, draws=PAR_SWEEPS, burnin=0, thin=1, sv_keep="last")
            s1.run()
            parity_ref = {k: s1.out[k].copy() for k in ("PHI", "logvar", "sv_para", "V_prior")}
            del s1

    trace("one-GPU parity chain done")
    h = Handle(local_rank, seed=2024)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (B.C.c_char * 128)()
            assert h._lib.bvb_comm_unique_id(buf) == 0
            uid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        h._check(h._lib.bvb_comm_init(h._h, rank, world, bytes(uid.cpu().tolist())))
    trace("communicator ready")
    peak = h.measure_fp64_peak() if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- N > 1 parity: the same fixed-seed chain, equations sharded, against the one-GPU chain
    parity = None
    if world > 1:
        seed_keep = h.seed
        h.reseed(4242)
        sp = B.Sampler(h, prep, pp, ps, draws=PAR_SWEEPS, burnin=0, thin=1, sv_keep="last")
        trace("sharded parity chain initialised")
        sp.run()
        trace("sharded parity chain done")
        if rank == 0:
            parity = {}
            for k, ref in parity_ref.items():
                got = sp.out[k]
                parity[k] = float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300))
        del sp
        h.reseed(seed_keep)
        barrier()

    # ---------------- value: device-resident chain, REPS regions of K timed sweeps after W warm-up sweeps
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    smp = B.Sampler(h, prep, pp, ps, draws=W + 1 + REPS * K, burnin=0, thin=1, sv_keep="last")
    trace("main chain initialised")
    smp.run(W)                        # warm-up (includes the eager first sweep + graph capture)
    trace("warm-up done")
    phases = smp.profile_sweep()      # per-phase events of one serial sweep: OUTSIDE the timed regions
    trace("profile sweep done")
    smp.kernel_times(reset=True)
    reps_ms = []
    for _rep in range(REPS):
        barrier()
        clocks.active.set()
        smp.run(K)
        barrier()
        clocks.active.clear()
        tm = smp.timing()
        assert tm["sweeps"] == K
        reps_ms.append(max_over_ranks(tm["total_ms"]))
    kt = smp.kernel_times()
    trace("timed regions done")
    clk = clocks.stop() if rank == 0 else None
    dev_ms = float(np.median(reps_ms))
    value = K / (dev_ms * 1e-3)
    stored_bytes = sum(smp.out[k][..., 0].nbytes for k in ("PHI", "V_prior", "logvar", "sv_para", "phi_hyperparameter", "facload", "fac"))
    out_template = {k: v.shape for k, v in smp.out.items()}
    launches_per_sweep = tm["launches_per_sweep"]
    del smp

    # ---------------- e2e: bvar()-equivalent through the host API, fresh chain of K sweeps
    # the caller-owned output arrays exist (and are resident) before the call, like R's would be; first-touch page
    # faults of a fresh VM are not part of the path being measured
    out_arrays = {k: np.zeros((*shp[:-1], K if shp[-1] else 0), order="F") for k, shp in out_template.items()}
    for a in out_arrays.values():
        a.fill(0.0)
    e2e_s = []
    for _rep in range(3):
        barrier()
        t0 = time.perf_counter()
        smp2 = B.Sampler(h, prep, pp, ps, draws=K, burnin=0, thin=1, sv_keep="last", out=out_arrays)
        while smp2.done < K:
            smp2.run(256)
        barrier()
        e2e_s.append(max_over_ranks(time.perf_counter() - t0))
        if rank == 0:   # (the stored draws land in rank 0's arrays only)
            assert np.isfinite(smp2.out["PHI"]).all() and np.abs(smp2.out["PHI"][..., 0]).max() > 0
        del smp2
    e2e = K / float(np.median(e2e_s))
    h2d = (prep["Y"].nbytes + prep["X"].nbytes + 3 * prep["PHI0"].nbytes + prep["startvals"]["sv_logvar"].nbytes) / K

    trace("e2e done")
    sharded = None
    if not args.no_draw_sharded:
        sharded = draw_sharded_postprocessing(h, rank, world, barrier, max_over_ranks)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    gf = algorithmic_gflop(w["M"], Kp, w["T"], w["r"])
    nloc = -(-w["M"] // world)
    frac_local = nloc / w["M"]
    kern = {}
    for name, key in (("k1_pairgemm", "k1"), ("k2_cholsolve", "k2")):
        ms = kt[key + "_ms"]
        ach = gf[key] * frac_local / (ms * 1e-3) / 1e3 if ms > 0 else None   # TFLOP/s
        kern[name] = dict(ms=ms, achieved=ach, share_of_sweep=ms / (1e3 / value) if value else None,
                          frac_of_self_measured_dmma_peak=(ach / peak["dmma_tflops"]) if ach else None)
    dom = max(kern, key=lambda k: kern[k]["ms"])
    traffic, traffic_src = None, None
    prof = ROOT / "profiles" / "ncu_summary.json"
    if prof.exists():
        try:
            summ = json.loads(prof.read_text())
            ent = summ.get(dom) or summ.get(dom + "_v2") or {}
            traffic = ent.get("dram_bytes_per_launch")
            traffic_src = f"profiles/ncu_summary.json (ncu --set full capture of {dom}, committed; not re-measured in this run)"
        except Exception:
            traffic = None
    roofline = dict(bound="tensor", kernel=dom, achieved=kern[dom]["achieved"], peak=peak["dmma_tflops"], unit="TFLOP/s",
                    frac=(kern[dom]["achieved"] / peak["dmma_tflops"]) if kern[dom]["achieved"] else None, traffic=traffic,
                    traffic_source=traffic_src,
                    peak_source="self-measured: FP64 DMMA m8n8k4 register-resident micro-kernel run in this process "
                                "(MEASURED_PEAKS.json has no FP64 entry)",
                    timing_source=f"live %globaltimer spans (first CTA start to last CTA end) of every launch inside the "
                                  f"{REPS} timed regions, mean of {kt['sweeps']} launches",
                    algorithmic_gflop_per_launch=gf["k1" if dom == "k1_pairgemm" else "k2"] * frac_local,
                    kernels=kern, dfma_peak_tflops=peak["dfma_tflops"],
                    phases_ms_serial_sweep=phases)
    cpu = cpu1 = anchor = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, cpu1 = cpu_baselines(prep, pp, ps)
        anchor = vignette_anchor()
    cfg = workload_config(world)
    line = dict(metric=METRIC, value=value, unit="sweeps/s", n_gpus=world,
                steps=K, warmup=W, ms_per_step=dev_ms / K, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=cfg,
                parallelism=("single GPU" if world == 1 else f"equations sharded over {world} GPUs"),
                repetitions=REPS, repetitions_ms=reps_ms, clocks=clk,
                e2e=dict(value=e2e, unit="sweeps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=stored_bytes,
                         repetitions_s=e2e_s,
                         note="median of 3 fresh chains, each: bvb_gibbs_init (H2D of data, priors, start values) + K sweeps, "
                              "every sweep's draw copied into caller-owned (pre-allocated) host arrays"),
                gpu_launches=launches_per_sweep * K, roofline=roofline, cpu_baseline=cpu, cpu_baseline_1thread=cpu1,
                vignette_anchor=anchor, draw_sharded=sharded)
    if parity is not None:
        line["parity_vs_n1"] = dict(max_rel_err=parity, sweeps=PAR_SWEEPS,
                                    note="fixed-seed chain with the equations sharded over the ranks against the same chain "
                                         "on one GPU, computed in this run (max |diff| / max |ref| over all stored draws)")
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-draw-sharded", action="store_true", help="skip the irf / predict draw-sharding section")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if args.steps == 256 and args.warmup == 32:    # no flags given: a CPU run that still ends within minutes
            args.steps, args.warmup = 20, 3
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
