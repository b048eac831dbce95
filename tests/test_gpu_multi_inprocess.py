"""bvb_create_multi / bvb_gibbs_init_multi / bvb_gibbs_run_multi: several GPUs driven from ONE process (what an R session
needs - no launcher).  The equation-sharded chain must reproduce the one-GPU chain with the same seed to the accuracy the
process-per-GPU form does (bench.py's parity_vs_n1: <= 1e-9 relative over every stored array; the tiled K2 of the shards
sums in a different order, so not bit for bit).  Skipped on a box with fewer than two GPUs."""
import numpy as np
import pytest
import torch

from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_in_one_process_match_one_gpu():
    rng = np.random.default_rng(5)
    M, p, T = 24, 2, 120
    Y = rng.standard_normal((T + p, M)).cumsum(axis=0) * 0.05 + rng.standard_normal((T + p, M))
    pp = B.specify_prior_phi(M, p, "DL")
    ps = B.specify_prior_sigma(M, factor_factors=1)
    prep = B.prepare(Y, p, 10.0, pp, ps, rng=np.random.default_rng(3))
    h1 = Handle(0, seed=77)
    s1 = B.Sampler(h1, prep, pp, ps, draws=30, burnin=10)
    s1.run()
    hm = B.MultiHandle([0, 1], seed=77)
    sm = B.Sampler(hm, prep, pp, ps, draws=30, burnin=10)
    while sm.done < 40:
        sm.run(16)                      # chunked, like the shim's interruptible loop
    for k in ("PHI", "logvar", "sv_para", "V_prior", "facload"):
        a, b = s1.out[k], sm.out[k]
        assert np.isfinite(b).all()
        assert np.abs(a - b).max() <= 1e-7 * np.abs(a).max(), k
    hm.close()
    h1.close()
