"""Generate the committed golden fixtures under tests/golden/.

The reference itself cannot be run here (no R / Rcpp / Armadillo, SURVEY.md
§8c), so the fixtures hold the LAPACK-faithful oracle's outputs on seeded
inputs.  They freeze the oracle (so a later edit of oracle/ cannot silently move
the target) and give the GPU parity tests value-level vectors that travel to the
GPU box.  Run from the repo root:  python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import regression as orc  # noqa: E402
from tools.synth import phi_draw_inputs  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def phi_factor_cases():
    cases = {}
    for name, (M, p, T, r, prior, seed) in {
        "tiny_r0": (3, 1, 40, 0, "flat", 11),
        "demo_like": (10, 2, 235, 1, "dl", 12),
        "fx_like": (14, 3, 100, 2, "dl", 13),
        "two_blocks": (8, 5, 90, 3, "dl", 14),
    }.items():
        d = phi_draw_inputs(M, p, T, r, seed=seed, prior=prior)
        out = orc.sample_PHI_factor(d["PHI0"], d["PHI0"], d["Y"], d["X"], d["logvar"], d["V"],
                                    d["facload"], d["fac"], False, d["z"])
        cases[name] = dict(M=M, p=p, T=T, r=r, prior=prior, seed=seed, PHI=out)
    return cases


if __name__ == "__main__":
    c = phi_factor_cases()
    flat = {}
    for k, v in c.items():
        for kk, vv in v.items():
            flat[f"{k}/{kk}"] = np.asarray(vv)
    np.savez_compressed(OUT / "phi_factor.npz", **flat)
    print("wrote", OUT / "phi_factor.npz", sorted(c))
