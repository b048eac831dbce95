"""GPU parity: bvb_predict (K7) against the oracle restatement of out_of_sample
(src/utilities_cpp.cpp:216-393) with injected normals.  Tolerance 1e-9 relative
(north_star: injected-normal conditional draws), through the C ABI."""
import numpy as np
import pytest

from oracle import predict as op

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def make_case(seed, M, p, intercept, R, each, ahead, factor, nf=2, sv="all", scale=0.25):
    rng = np.random.default_rng(seed)
    Kp = M * p + (1 if intercept else 0)
    PHI = scale / np.sqrt(M) * rng.standard_normal((Kp, M, R))
    x = rng.standard_normal(Kp)
    if intercept:
        x[-1] = 1.0
    ahead = np.asarray(ahead, int)
    na, mx = ahead.size, int(ahead.max())
    if factor:
        facload = rng.standard_normal((M, nf, R))
        U = None
        nlv = M + nf
    else:
        facload = None
        U = 0.3 * rng.standard_normal((M * (M - 1) // 2, R))
        nlv = M
    logvar_T = rng.normal(-1.0, 0.4, (nlv, R))
    if sv == "all":
        ind = np.arange(nlv)
    elif sv == "none":
        ind = np.zeros(0, int)
    else:
        ind = np.arange(0, nlv, 2)
    nsv = ind.size
    mu = rng.normal(-1.0, 0.3, (nsv, R))
    phi = rng.uniform(0.7, 0.98, (nsv, R))
    sig = rng.uniform(0.05, 0.4, (nsv, R))
    Y_obs = rng.standard_normal((na, M))
    VoI = np.sort(rng.choice(M, size=max(1, M // 2), replace=False))
    zl = rng.standard_normal((R, mx, each, nsv))
    zp = rng.standard_normal((R, na, each, M))
    return dict(each=each, X_T_plus_1=x, PHI=PHI, U=U, facload=facload, logvar_T=logvar_T, ahead=ahead, sv_mu=mu, sv_phi=phi,
                sv_sigma=sig, sv_indicator=ind, factor=factor, LPL=True, Y_obs=Y_obs, LPL_subset=True, VoI=VoI,
                simulate_predictive=True, z_logvar=zl, z_pred=zp)


def check(handle, c):
    ref = op.out_of_sample(**c)
    got = handle.out_of_sample(**c)
    for k in ("LPL_draws", "LPL_sub_draws", "predictions"):
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=1e-10, err_msg=k)
    np.testing.assert_allclose(got["PL_univariate_draws"], ref["PL_univariate_draws"], rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("cfg", [
    dict(seed=1, M=5, p=2, intercept=True, R=4, each=3, ahead=[1, 3, 4], factor=True),
    dict(seed=2, M=4, p=1, intercept=False, R=3, each=2, ahead=[2], factor=False, sv="some"),
    dict(seed=3, M=4, p=3, intercept=True, R=2, each=2, ahead=[1, 2, 3, 4, 5, 6], factor=False),
    dict(seed=4, M=3, p=4, intercept=True, R=5, each=1, ahead=[2, 7], factor=True, nf=1, sv="none"),
    dict(seed=5, M=6, p=2, intercept=False, R=3, each=2, ahead=[1, 2], factor=True, nf=0, sv="some"),
    dict(seed=6, M=1, p=2, intercept=True, R=3, each=2, ahead=[1, 3], factor=False),
])
def test_out_of_sample_matches_oracle(handle, cfg):
    check(handle, make_case(**cfg))


def test_out_of_sample_tiles_cross_boundaries(handle):
    # Kp = 141 > two 64-wide GEMM tiles, M = 70 > one tile
    check(handle, make_case(seed=7, M=70, p=2, intercept=True, R=2, each=2, ahead=[1, 2, 4], factor=True, nf=3))
    check(handle, make_case(seed=8, M=33, p=3, intercept=True, R=2, each=1, ahead=[3, 5], factor=False))


def test_out_of_sample_config3_shape(handle):
    # M = 100, p = 4 (Kp = 401): 7 Cholesky panels of 16 columns, ping-pong companion covariance over 6 horizons (> p)
    check(handle, make_case(seed=11, M=100, p=4, intercept=True, R=2, each=1, ahead=[1, 4, 6], factor=True, nf=2))


def test_many_chains_stride_over_ctas(handle):
    # more chains than resident CTAs (2 x 148): R*each = 700
    check(handle, make_case(seed=9, M=3, p=2, intercept=True, R=350, each=2, ahead=[1, 2], factor=True, nf=1))


def test_lpl_only_and_predict_only(handle):
    c = make_case(seed=10, M=5, p=2, intercept=True, R=3, each=2, ahead=[1, 2], factor=True)
    ref = op.out_of_sample(**c)
    a = handle.out_of_sample(**{**c, "simulate_predictive": False, "z_pred": None})
    np.testing.assert_allclose(a["LPL_draws"], ref["LPL_draws"], rtol=RTOL)
    assert a["predictions"].size == 0
    b = handle.out_of_sample(**{**c, "LPL": False, "LPL_subset": False})
    np.testing.assert_allclose(b["predictions"], ref["predictions"], rtol=RTOL, atol=1e-10)
    assert b["LPL_draws"].size == 0


def test_device_normals_have_the_predictive_law(handle):
    # homoscedastic, one stored draw, many replicates: predictions ~ N(mean_h, cov_h); moments from the oracle with z = 0
    c = make_case(seed=11, M=4, p=2, intercept=True, R=1, each=20000, ahead=[1, 3], factor=True, sv="none")
    c0 = {**c, "each": 1, "z_logvar": np.zeros((1, 3, 1, 0)), "z_pred": np.zeros((1, 2, 1, 4))}
    mean = op.out_of_sample(**c0)["predictions"][:, :, 0]
    got = handle.out_of_sample(**{**c, "z_logvar": None, "z_pred": None})
    pred = got["predictions"]
    assert np.isfinite(pred).all()
    for k in range(2):
        x = pred[k].T                      # each x M
        se = x.std(axis=0) / np.sqrt(x.shape[0])
        assert np.all(np.abs(x.mean(axis=0) - mean[k]) < 5 * se)
    # two calls with the same handle seed reproduce
    again = handle.out_of_sample(**{**c, "z_logvar": None, "z_pred": None})
    np.testing.assert_array_equal(again["predictions"], pred)
    # replicates differ
    assert np.unique(pred[0, 0]).size > 19000


def test_sv_device_normals_spread_logvariances(handle):
    c = make_case(seed=12, M=3, p=1, intercept=True, R=2, each=4000, ahead=[4], factor=False)
    got = handle.out_of_sample(**{**c, "z_logvar": None, "z_pred": None})
    ll = got["LPL_draws"]
    assert np.isfinite(ll).all() and ll.std() > 0


def test_argument_errors(handle):
    from bayesianvars_b200._lib import BvbError
    c = make_case(seed=13, M=3, p=1, intercept=True, R=2, each=1, ahead=[1, 2], factor=True)
    with pytest.raises(BvbError) as e:
        handle.out_of_sample(**{**c, "ahead": np.array([2, 1]), "z_logvar": None, "z_pred": None})
    assert e.value.code == 2
    # not positive definite: negative infinite log-variance is fine, NaN loading is not
    bad = {**c, "facload": np.full_like(c["facload"], np.nan)}
    with pytest.raises(BvbError) as e:
        handle.out_of_sample(**bad)
    assert e.value.code == 3 and "chol()" in str(e.value)


@pytest.mark.parametrize("sigma", ["factor", "cholesky"])
def test_predict_method_on_a_fitted_model(handle, sigma):
    """predict.bayesianVARs_bvar (R/utility_functions.R:2406-2530) on a homoscedastic fit: the log predictive
    likelihood draws are deterministic functions of the stored draws -> compare with the oracle."""
    from bayesianvars_b200 import bvar as bv
    from bayesianvars_b200.predict import predict, stable_bvar
    rng = np.random.default_rng(20)
    M, lags, T = 3, 2, 80
    data = np.cumsum(0.1 * rng.standard_normal((T + 4, M)), axis=0) * 0.2 + rng.standard_normal((T + 4, M))
    train, test = data[:-4], data[-4:]
    pp = bv.specify_prior_phi(M, lags, prior="HS")
    if sigma == "factor":
        ps = bv.specify_prior_sigma(M, type="factor", factor_factors=1, factor_heteroskedastic=False)
    else:
        ps = bv.specify_prior_sigma(M, type="cholesky", cholesky_heteroscedastic=False)
    mod = bv.bvar(train, lags=lags, draws=60, burnin=40, prior_phi=pp, prior_sigma=ps, handle=handle, rng=np.random.default_rng(3))
    out = predict(mod, ahead=[1, 2, 3, 4], LPL=True, Y_obs=test, LPL_VoI=[1, 3], handle=handle)
    st = stable_bvar(mod, quiet=True)
    R = st["PHI"].shape[2]
    assert out["predictions"].shape == (4, M, R) and np.isfinite(out["predictions"]).all()
    x = np.concatenate([train[-1], train[-2], [1.0]])
    none = np.zeros((0, R))
    ref = op.out_of_sample(1, x, st["PHI"], st["U"] if sigma == "cholesky" else None,
                           st["facload"] if sigma == "factor" else None, st["logvar"][-1], np.array([1, 2, 3, 4]), none, none, none,
                           np.zeros(0, int), sigma == "factor", True, test, True, np.array([0, 2]), False,
                           np.zeros((R, 4, 1, 0)), None)
    np.testing.assert_allclose(out["LPL_draws"], ref["LPL_draws"], rtol=RTOL)
    np.testing.assert_allclose(out["LPL_sub_draws"], ref["LPL_sub_draws"], rtol=RTOL)
    mx = ref["LPL_draws"].max(axis=1) - 700
    np.testing.assert_allclose(out["LPL"], np.log(np.exp(ref["LPL_draws"] - mx[:, None]).mean(axis=1)) + mx, rtol=1e-9)
    assert out["LPL_univariate"].shape == (4, M) and out["LPL_VoI"].shape == (4,)
