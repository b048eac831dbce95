"""tests/golden/entry_points.npz (generator: tools/make_golden.py) holds self-contained input/output vectors for every
entry point of the C ABI besides the factor PHI draw.  CPU: the oracle still reproduces them (a later edit of oracle/
cannot silently move the target).  GPU: the CUDA path matches them through the C ABI, with nothing but the fixture and
the library on the GPU box."""
import numpy as np
import pytest

from oracle import irf as oirf
from oracle import predict as opred
from oracle import regression as orc

RTOL = 1e-9


@pytest.fixture(scope="module")
def g(request):
    z = np.load(request.config.rootpath / "tests" / "golden" / "entry_points.npz")
    out = {}
    for k in z.files:
        case, name = k.split("/")
        out.setdefault(case, {})[name] = z[k]
    return out


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


# ------------------------------------------------------------------ CPU: oracle vs fixtures
def test_oracle_reproduces_fixtures(g):
    c = g["phi_cholesky"]
    assert relerr(orc.sample_PHI(c["PHI"], c["PHI0"], c["Y"], c["X"], c["U"], c["d_sqrt"], c["V"], c["z"]), c["out"]) < 1e-12
    c = g["sample_u"]
    assert relerr(orc.sample_U(np.eye(4), c["E"], c["V"], c["d_sqrt"], c["z"]), c["out"]) < 1e-12
    c = g["phi_huge"]
    out = orc.sample_PHI_factor(c["PHI0"], c["PHI0"], c["Y"], c["X"], c["logvar"], c["V"], c["facload"], c["fac"], True,
                                c["z"], c["delta"])
    assert relerr(out, c["out"]) < 1e-10
    c = g["irf_factor"]
    assert relerr(oirf.irf_cpp(c["coefficients"], c["factor_loadings"], None, c["logvar_t"], c["shocks"], int(c["ahead"])), c["out"]) < 1e-12
    c = g["irf_cholesky"]
    assert relerr(oirf.irf_cpp(c["coefficients"], None, c["U_vecs"], c["logvar_t"], c["shocks"], int(c["ahead"])), c["out"]) < 1e-12
    c = g["predict_factor"]
    res = opred.out_of_sample(int(c["each"]), c["X_T_plus_1"], c["PHI"], None, c["facload"], c["logvar_T"], c["ahead"], c["sv_mu"],
                              c["sv_phi"], c["sv_sigma"], c["sv_indicator"], True, True, c["Y_obs"], True, c["VoI"], True,
                              c["z_logvar"], c["z_pred"])
    for k in ("LPL_draws", "PL_univariate_draws", "LPL_sub_draws", "predictions"):
        assert relerr(res[k], c["out_" + k]) < 1e-12, k


# ------------------------------------------------------------------ GPU: CUDA path vs fixtures
@pytest.mark.gpu
def test_cuda_matches_fixtures(handle, g):
    c = g["phi_cholesky"]
    out = handle.sample_PHI_cholesky(c["PHI"], c["PHI0"], c["Y"], c["X"], c["U"], c["d_sqrt"], c["V"], z=c["z"])
    assert relerr(out, c["out"]) < RTOL
    c = g["sample_u"]
    out = handle.sample_U(c["E"], c["V"], c["d_sqrt"], z=c["z"])
    iu = np.triu_indices(4, 1)
    assert relerr(out[iu], c["out"][iu]) < RTOL
    c = g["phi_huge"]
    out = handle.sample_PHI_factor(c["PHI0"], c["Y"], c["X"], c["logvar"], c["V"], c["facload"], c["fac"], huge=True,
                                   z=c["z"], delta=c["delta"])
    assert relerr(out, c["out"]) < RTOL
    c = g["irf_factor"]
    out = handle.irf_cpp(c["coefficients"], c["factor_loadings"], None, c["logvar_t"], c["shocks"], int(c["ahead"]))
    assert relerr(out, c["out"]) < 1e-12
    c = g["irf_cholesky"]
    out = handle.irf_cpp(c["coefficients"], None, c["U_vecs"], c["logvar_t"], c["shocks"], int(c["ahead"]))
    assert relerr(out, c["out"]) < 1e-12
    c = g["predict_factor"]
    res = handle.out_of_sample(int(c["each"]), c["X_T_plus_1"], c["PHI"], None, c["facload"], c["logvar_T"], c["ahead"],
                               c["sv_mu"], c["sv_phi"], c["sv_sigma"], c["sv_indicator"], True, True, c["Y_obs"], True,
                               c["VoI"], True, z_logvar=c["z_logvar"], z_pred=c["z_pred"])
    for k in ("LPL_draws", "PL_univariate_draws", "LPL_sub_draws", "predictions"):
        assert relerr(res[k], c["out_" + k]) < RTOL, k
