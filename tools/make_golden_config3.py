"""Golden fixture for the chain-level check at BASELINE configs[2] (M = 100, p = 4, T = 500, factor-SV r = 1, DL).

The oracle chain (oracle.gibbs, the NumPy / SciPy-LAPACK restatement of bvar_cpp) needs ~0.5 s per sweep at this
shape, too slow to repeat inside the GPU test-suite, so a 600-sweep chain is run ONCE here and the draws of a subset
of the state are committed: tests/golden/config3_chain.npz.  tests/test_gpu_gibbs.py::test_config3_chain_matches_oracle
compares the CUDA chain on the same inputs with it (posterior means within 5 Monte-Carlo standard errors).

    python tools/make_golden_config3.py          # ~6 minutes on 8 cores
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import make_problem  # noqa: E402  (the bench workload: data seed 123, start values seed 123)
from oracle import gibbs as og  # noqa: E402

DRAWS, BURN = 400, 200


def subset(Kp, M):
    """Coefficients kept: the own first lags of 12 equations, a block of cross lags, every intercept of those."""
    eqs = np.arange(0, M, 9)[:12]
    rows, cols = [], []
    for j in eqs:
        rows += [j, (j + 1) % M, M + j, 2 * M + j, 3 * M + (j + 7) % M, Kp - 1]
        cols += [j] * 6
    return np.asarray(rows), np.asarray(cols), eqs


def main():
    data, pp, ps, prep = make_problem()
    t0 = time.time()
    o = og.run(prep, pp, ps, draws=DRAWS, burnin=BURN, rng=np.random.default_rng(77))
    rows, cols, eqs = subset(prep["Kp"], prep["M"])
    out = dict(rows=rows, cols=cols, eqs=eqs, draws=DRAWS, burnin=BURN,
               PHI=o["PHI"][rows, cols, :], sv_para=o["sv_para"][:, np.r_[eqs, prep["M"]], :],
               logvar_last=o["logvar"][-1][np.r_[eqs, prep["M"]], :],
               logV_mean=np.log(o["V_prior"][:-1]).mean(axis=(0, 1)),
               facload_sq=(o["facload"][:, 0, :] ** 2).sum(axis=0))
    np.savez_compressed(ROOT / "tests" / "golden" / "config3_chain.npz", **out)
    print(f"oracle chain: {DRAWS + BURN} sweeps in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
