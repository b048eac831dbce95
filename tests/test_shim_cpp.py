"""The Rcpp shim (r_shim/bvar_cpp_shim.cpp) compiled against the Rcpp stand-in of tests/r_stub and RUN: a C++ caller of
the C ABI with exactly the struct layouts, argument order and ownership rules the shim uses inside R.

The driver (tests/r_stub/shim_driver.cpp) receives the argument lists of bvar_cpp() - built here the way R's
specify_prior_* constructors build them, placeholders for the unused keys included
(R/utility_functions.R:374-420, 1585-1617) - calls bvar_cpp(), irf_cpp(), out_of_sample(), my_gig(), vcov_cpp(),
predh() and sample_PHI_cholesky(), and writes the results back.  The chain must equal, bit for bit, the chain the
ctypes host runs with the same Philox key."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bayesianvars_b200 import bvar as B

ROOT = Path(__file__).resolve().parent.parent
STUB = ROOT / "tests" / "r_stub"
DRIVER = STUB / "shim_driver"


def build_driver():
    srcs = [STUB / "shim_driver.cpp", STUB / "Rcpp.h", ROOT / "r_shim" / "bvar_cpp_shim.cpp", ROOT / "include" / "bvb.h"]
    if DRIVER.exists() and all(DRIVER.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return DRIVER
    from bayesianvars_b200.build import build_lib
    build_lib()
    cmd = ["g++", "-std=c++17", "-O1", f"-I{STUB}", f"-I{ROOT / 'include'}", str(STUB / "shim_driver.cpp"), "-o", str(DRIVER),
           f"-L{ROOT / 'bayesianvars_b200'}", "-lbvb", "-Wl,-rpath," + str(ROOT / "bayesianvars_b200")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return DRIVER


# ---- the driver's file format -------------------------------------------------------------------------------
def _w(o):
    if o is None:
        return b"\x00"
    if isinstance(o, dict):
        return b"\x05" + struct.pack("<i", len(o)) + b"".join(_ws(k) + _w(v) for k, v in o.items())
    if isinstance(o, str):
        return b"\x04" + _ws(o)
    if isinstance(o, (bool, np.bool_)):
        return b"\x03" + struct.pack("<ii", 1, 1) + struct.pack("<i", int(o))
    a = np.asarray(o)
    if a.dtype == bool:
        kind, a = 3, a.astype("<i4")
    elif np.issubdtype(a.dtype, np.integer):
        kind, a = 2, a.astype("<i4")
    else:
        kind, a = 1, a.astype("<f8")
    a = np.atleast_1d(a)
    return bytes([kind]) + struct.pack("<i", a.ndim) + struct.pack(f"<{a.ndim}i", *a.shape) + np.asfortranarray(a).tobytes(order="F")


def _ws(s):
    b = s.encode()
    return struct.pack("<i", len(b)) + b


def _r(buf, o=0):
    k = buf[o]; o += 1
    if k == 0:
        return None, o
    if k == 4:
        n, = struct.unpack_from("<i", buf, o); return buf[o + 4:o + 4 + n].decode(), o + 4 + n
    if k == 5:
        n, = struct.unpack_from("<i", buf, o); o += 4
        d = {}
        for _ in range(n):
            ln, = struct.unpack_from("<i", buf, o); name = buf[o + 4:o + 4 + ln].decode(); o += 4 + ln
            d[name], o = _r(buf, o)
        return d, o
    nd, = struct.unpack_from("<i", buf, o); o += 4
    dims = struct.unpack_from(f"<{nd}i", buf, o); o += 4 * nd
    n = int(np.prod(dims))
    dt = "<f8" if k == 1 else "<i4"
    a = np.frombuffer(buf, dtype=dt, count=n, offset=o).reshape(dims, order="F")
    return a, o + n * (8 if k == 1 else 4)


# ---- R's list constructors (placeholders for the keys a structure does not use) ---------------------------------
def r_prior_sigma_list(cpp):
    out = dict(type="", cholesky_U_prior="", cholesky_V_i=np.zeros(1), cholesky_a=0.0, cholesky_b=0.0, cholesky_c=0.0,
               cholesky_GL_tol=0.0, cholesky_a_vec=np.zeros(1), cholesky_a_weight=np.zeros(1), cholesky_norm_consts=np.zeros(1),
               cholesky_c_vec=np.zeros(1), cholesky_c_rel_a=False, cholesky_DL_hyper=False, cholesky_DL_plus=False,
               cholesky_GT_hyper=False, cholesky_GT_vs=0.0, cholesky_GT_priorkernel="", cholesky_SSVS_tau0=np.zeros(1),
               cholesky_SSVS_tau1=np.zeros(1), cholesky_SSVS_s_a=0.0, cholesky_SSVS_s_b=0.0, cholesky_SSVS_hyper=False,
               cholesky_SSVS_p=np.zeros(1), cholesky_lambda_3=np.zeros(1), cholesky_priorhomoscedastic=np.zeros((1, 1)),
               cholesky_sv_offset=np.zeros(1), factor_factors=0,
               factor_shrinkagepriors=dict(a=np.zeros(1), c=np.zeros(1), d=np.zeros(1)), factor_ngprior=False,
               factor_columnwise=False, factor_facloadtol=0.0, factor_restrinv=np.zeros((1, 1), np.int32),
               factor_priorhomoskedastic=np.zeros((1, 1)), factor_interweaving=0, sv_heteroscedastic=np.zeros(1),
               sv_priormu=np.zeros(1), sv_priorphi=np.zeros(1), sv_priorsigma2=np.zeros((1, 1)), sv_priorh0=np.zeros(1))
    for k, v in cpp.items():
        if k == "sv_heteroscedastic":
            v = np.asarray(v, bool)
        out[k] = v
    return out


def r_startvals_list(prep, chol):
    sv = prep["startvals"]
    M, T = prep["M"], prep["T"]
    z = np.zeros((0, 0))
    return dict(PHI=sv["PHI"], U=sv.get("U", z) if chol else z, sv_para=sv["sv_para"], sv_logvar=sv["sv_logvar"],
                sv_logvar0=sv["sv_logvar0"],
                factor_startval=dict(facload=sv.get("facload", z), fac=sv.get("fac", z), tau2=sv.get("tau2", z)))


def run_driver(tmp_path, prep, pp, ps, draws, burnin, thin, tvp_keep, huge=False):
    chol = ps["prior_sigma_cpp"]["type"] == "cholesky"
    args = dict(Y=prep["Y"], X=prep["X"], M=prep["M"], T=prep["T"], K=prep["K"], draws=draws, burnin=burnin, thin=thin,
                tvp_keep=tvp_keep, intercept=prep["intercept"], priorIntercept=prep["priorIntercept"], PHI0=prep["PHI0"],
                priorPHI_in=dict(pp["prior_phi_cpp"]), priorSigma_in=r_prior_sigma_list(ps["prior_sigma_cpp"]),
                Rstartvals_in=r_startvals_list(prep, chol), i_mat=prep["i_mat"], i_vec=prep["i_vec"],
                PHI_tol=float(pp["general_settings"]["PHI_tol"]), L_tol=float(ps["general_settings"]["cholesky_U_tol"]),
                huge=bool(huge))
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    fin.write_bytes(_w(args))
    p = subprocess.run([str(build_driver()), str(fin), str(fout)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    out, _ = _r(fout.read_bytes())
    return out


def test_driver_compiles_and_file_format_round_trips():
    """CPU: the shim + stand-in compile with g++ and link against libbvb.so; the test's own file format round-trips."""
    assert build_driver().exists()
    obj = dict(a=np.arange(6.0).reshape(2, 3), b=dict(c="x", d=np.array([1, 2], np.int32)), e=True, f=None)
    back, _ = _r(_w(obj))
    assert np.array_equal(back["a"], obj["a"]) and back["b"]["c"] == "x" and np.array_equal(back["b"]["d"], [1, 2])
    assert back["e"][0] == 1 and back["f"] is None


@pytest.mark.gpu
@pytest.mark.parametrize("sigma", ["factor", "cholesky"])
def test_shim_runs_the_same_chain_as_the_ctypes_host(tmp_path, sigma):
    from tests.test_gpu_gibbs import sim_var
    from bayesianvars_b200 import Handle
    M, p, T = 6, 2, 90
    Y, _ = sim_var(M, p, T)
    pp = B.specify_prior_phi(M, p, "DL")
    ps = B.specify_prior_sigma(M, factor_factors=2) if sigma == "factor" else B.specify_prior_sigma(M, type="cholesky", cholesky_U_prior="HS")
    prep = B.prepare(Y, p, 10.0, pp, ps, rng=np.random.default_rng(3))
    draws, burnin, thin = 40, 15, 2
    out = run_driver(tmp_path, prep, pp, ps, draws, burnin, thin, "all")
    res = out["bvar"]
    nsave, Kp, S = draws // thin, prep["Kp"], M + (2 if sigma == "factor" else 0)
    assert res["PHI"].shape == (Kp, M, nsave) and res["logvar"].shape == (prep["T"], S, nsave) and res["sv_para"].shape == (3, S, nsave)
    seed = (int(res_seed := out["seed_hi_lo"][0]) << 32) | int(out["seed_hi_lo"][1])
    del res_seed
    with Handle(0, seed=seed) as h:
        smp = B.Sampler(h, prep, pp, ps, draws=draws, burnin=burnin, thin=thin, sv_keep="all")
        smp.run()
        keys = ["PHI", "V_prior", "logvar", "sv_para", "phi_hyperparameter"] + (["facload", "fac"] if sigma == "factor" else ["U", "u_hyperparameter"])
        for k in keys:
            assert res[k].shape == smp.out[k].shape, (k, res[k].shape, smp.out[k].shape)
            assert np.array_equal(res[k], smp.out[k]), k
    if sigma == "factor":
        assert res["U"].size == 0 and res["u_hyperparameter"].size == 0
    else:
        nU = M * (M - 1) // 2
        assert res["U"].shape == (nU, nsave) and res["u_hyperparameter"].shape == (2 + 2 * nU, nsave)
        assert res["facload"].size == 0 and np.isfinite(out["phi_chol"]).all() and out["phi_chol"].shape == (Kp, M)
    # the other entry points, fed from the chain inside the driver
    from oracle.irf import irf_cpp as oracle_irf
    lvT = res["logvar"][-1]
    n_sh = 2
    shocks = np.eye(2 if sigma == "factor" else M)[:, :n_sh]
    irf = np.stack([out["irf"][str(r)] for r in range(nsave)], axis=-1)
    want = oracle_irf(res["PHI"], res["facload"] if sigma == "factor" else None, None if sigma == "factor" else res["U"], lvT, shocks, 5)
    assert irf.shape == want.shape and np.max(np.abs(irf - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    pr = out["pred"]
    assert pr["predictions"].shape == (2, M, 2 * nsave) and pr["LPL_draws"].shape == (2, 2 * nsave) and pr["LPL_sub_draws"].shape == (2, 2 * nsave)
    assert np.isfinite(pr["predictions"]).all() and np.isfinite(pr["LPL_draws"]).all() and (pr["PL_univariate_draws"] > 0).all()
    assert out["gig"].shape == (1000, 3) and (out["gig"] > 0).all()
    assert out["vcov"].shape == (prep["T"] * M, M, nsave) and out["predh"].shape == (2, S, nsave)
    from oracle.fitted import vcov_cpp as oracle_vcov
    wv = oracle_vcov(sigma == "factor", res["facload"], res["logvar"], res["U"], M, 2 if sigma == "factor" else 0)
    assert np.max(np.abs(out["vcov"] - wv)) <= 1e-12 * np.max(np.abs(wv))
