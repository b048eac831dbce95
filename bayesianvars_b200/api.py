"""Host-side mirror of the reference's Rcpp interface for the hot path.

Function names, argument meaning and error behaviour follow the reference's
exported C++ (`src/sample_coefficients.h:9-55`, `src/RcppExports.cpp`) so the
parity tests read like the reference's own tests.  Every call goes through the
C ABI of libbvb.so with HOST buffers - exactly what the Rcpp shim does.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BvbError, fmat, fptr


class Handle:
    """One per process / GPU (bvb_create / bvb_destroy)."""

    def __init__(self, device: int = 0, seed: int = 1):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.bvb_create(C.byref(h), int(device), C.c_uint64(seed))
        if rc != _lib.BVB_OK:
            raise BvbError(rc, "bvb_create failed: no usable sm_100a CUDA device (there is no CPU fallback)")
        self._h = h
        self.device = device
        self.seed = int(seed)

    def reseed(self, seed: int):
        """New Philox key for the calls that follow (bvb_set_seed)."""
        self._check(self._lib.bvb_set_seed(self._h, C.c_uint64(int(seed))))
        self.seed = int(seed)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bvb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- helpers -----------------------------------------------------------
    def _check(self, rc):
        if rc == _lib.BVB_OK:
            return
        step, sweep, eq = C.c_int(), C.c_int(), C.c_int()
        buf = C.create_string_buffer(512)
        self._lib.bvb_last_error(self._h, C.byref(step), C.byref(sweep), C.byref(eq), buf, 512)
        raise BvbError(rc, buf.value.decode(), step.value, sweep.value, eq.value)

    def last_timing(self):
        out = np.zeros(4)
        n = self._lib.bvb_last_timing(self._h, fptr(out))
        return {"prep_ms": out[0], "k1_ms": out[1], "k2_ms": out[2], "rest_ms": out[3], "launches": n}

    def measure_fp64_peak(self):
        out = np.zeros(3)
        self._check(self._lib.bvb_measure_fp64_peak(self._h, fptr(out)))
        return {"dmma_tflops": out[0], "dfma_tflops": out[1], "sm_mhz": out[2]}

    # -- a3: sample_PHI_factor (sample_coefficients.cpp:84-114) -------------
    def sample_PHI_factor(self, PHI_prior, Y, X, logvaridi, V_prior, facload=None, fac=None,
                          huge=False, z=None, delta=None):
        Y, X, logvaridi = fmat(Y), fmat(X), fmat(logvaridi)
        V_prior, PHI_prior = fmat(V_prior), fmat(PHI_prior)
        T, Kp = X.shape
        M = Y.shape[1]
        r = 0 if fac is None or np.size(fac) == 0 else np.shape(fac)[0]
        facload = fmat(facload) if r else None
        fac = fmat(fac) if r else None
        z = fmat(z)
        delta = fmat(delta)
        out = np.empty((Kp, M), order="F")
        rc = self._lib.bvb_phi_draw_factor(self._h, T, Kp, M, r, fptr(X), fptr(Y), fptr(logvaridi),
                                           fptr(V_prior), fptr(PHI_prior), fptr(facload), fptr(fac),
                                           fptr(z), fptr(delta), int(bool(huge)), fptr(out))
        self._check(rc)
        return out

    # -- a4x: sample_PHI_cholesky (sample_coefficients.cpp:149-157) ---------
    def sample_PHI_cholesky(self, PHI, PHI_prior, Y, X, U, d_sqrt, V_prior, z=None):
        PHI, PHI_prior, Y, X = fmat(PHI), fmat(PHI_prior), fmat(Y), fmat(X)
        U, d_sqrt, V_prior, z = fmat(U), fmat(d_sqrt), fmat(V_prior), fmat(z)
        T, Kp = X.shape
        M = Y.shape[1]
        out = np.empty((Kp, M), order="F")
        rc = self._lib.bvb_phi_draw_cholesky(self._h, T, Kp, M, fptr(PHI), fptr(PHI_prior), fptr(Y),
                                             fptr(X), fptr(U), fptr(d_sqrt), fptr(V_prior), fptr(z), fptr(out))
        self._check(rc)
        return out

    # -- a5: sample_U (sample_coefficients.cpp:160-188) ----------------------
    def sample_U(self, resid, V_i_U, d_sqrt, z=None):
        resid, d_sqrt = fmat(resid), fmat(d_sqrt)
        V_i_U = np.ascontiguousarray(V_i_U, dtype=np.float64)
        z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        T, M = resid.shape
        out = np.empty((M, M), order="F")
        rc = self._lib.bvb_sample_u(self._h, T, M, fptr(resid), fptr(V_i_U), fptr(d_sqrt), fptr(z), fptr(out))
        self._check(rc)
        return out

    # -- a11: my_gig (gig_vec.cpp:20-43) --------------------------------------
    def my_gig(self, n, lam, chi, psi, stream=0):
        lam = np.atleast_1d(np.asarray(lam, dtype=np.float64))
        chi = np.atleast_1d(np.asarray(chi, dtype=np.float64))
        psi = np.atleast_1d(np.asarray(psi, dtype=np.float64))
        m = max(lam.size, chi.size, psi.size)
        out = np.empty((n, m), order="F")
        rc = self._lib.bvb_gig(self._h, int(n), fptr(lam), lam.size, fptr(chi), chi.size, fptr(psi), psi.size,
                               C.c_uint64(stream), fptr(out))
        self._check(rc)
        return out

    # -- a12: tolerance clamp (src/bvar_cpp.cpp:534-559) ---------------------------
    def tolerance_clamp(self, PHI, PHI0, K, tol, sweep=0):
        """|PHI - PHI0| >= tol on the first K rows; exact zeros get a random sign.  Returns the clamped copy."""
        from . import bvar as _b  # registers the signature
        out = np.asfortranarray(PHI, dtype=np.float64).copy(order="F")
        P0 = np.asfortranarray(PHI0, dtype=np.float64)
        Kp, M = out.shape
        self._check(self._lib.bvb_tolerance_clamp(self._h, int(K), Kp, M, fptr(out), fptr(P0), float(tol), int(sweep)))
        return out

    # -- a14: irf_cpp (src/irf.cpp:468-524) ---------------------------------------
    def irf_cpp(self, coefficients, factor_loadings, U_vecs, logvar_t, shocks, ahead, rotation=None):
        """Returns the (variables, shocks, 1+ahead, draws) array R/irf.R:275-276 builds."""
        coefficients = np.asfortranarray(coefficients, dtype=np.float64)
        Kp, M, R = coefficients.shape
        nf = 0 if factor_loadings is None or np.size(factor_loadings) == 0 else np.shape(factor_loadings)[1]
        has_U = 0 if U_vecs is None or np.size(U_vecs) == 0 or nf > 0 else 1
        fl = np.asfortranarray(factor_loadings, dtype=np.float64) if nf else None
        uv = np.asfortranarray(U_vecs, dtype=np.float64) if has_U else None
        lv = np.asfortranarray(np.asarray(logvar_t, dtype=np.float64).reshape(-1, R)) if logvar_t is not None and np.size(logvar_t) else None
        shocks = np.asfortranarray(shocks, dtype=np.float64)
        dim, n_shocks = shocks.shape
        rot = np.asfortranarray(rotation, dtype=np.float64) if rotation is not None else None
        out = np.empty((M, n_shocks, ahead + 1, R), order="F")
        rc = self._lib.bvb_irf(self._h, Kp, M, R, nf, has_U, 0 if lv is None else lv.shape[0], dim, n_shocks, int(ahead),
                               fptr(coefficients.reshape(-1, order="F")), fptr(fl.reshape(-1, order="F")) if nf else None,
                               fptr(uv) if has_U else None, fptr(lv) if lv is not None else None, fptr(shocks),
                               fptr(rot.reshape(-1, order="F")) if rot is not None else None, fptr(out.reshape(-1, order="F")))
        self._check(rc)
        return out

    # -- section 8(f) rank 2: predh / vcov_cpp (src/utilities_cpp.cpp:181-212, 436-472) --
    def predh(self, logvar_T, ahead, each, sv_mu, sv_phi, sv_sigma, z=None):
        """Same arguments as the Rcpp export (R/RcppExports.R:54); returns (len(ahead), M, R*each).
        z (optional, parity tests): the injected standard normals, shape (len(ahead), each, M, R)."""
        lv = np.asfortranarray(logvar_T, dtype=np.float64)
        M, R = lv.shape
        ahead = np.ascontiguousarray(ahead, dtype=np.int32)
        out = np.empty((ahead.size, M, R * int(each)), order="F")
        F = lambda a: fptr(np.asfortranarray(a, dtype=np.float64).reshape(-1, order="F"))
        zz = None
        if z is not None:   # (na, each, M, R) -> C-ordered [na][each][R][M]
            zz = np.ascontiguousarray(np.transpose(np.asarray(z, dtype=np.float64), (0, 1, 3, 2)))
        rc = self._lib.bvb_predh(self._h, M, R, int(each), ahead.ctypes.data_as(C.POINTER(C.c_int)), int(ahead.size), F(lv),
                                 F(sv_mu), F(sv_phi), F(sv_sigma), fptr(zz.reshape(-1)) if zz is not None else None,
                                 fptr(out.reshape(-1, order="F")))
        self._check(rc)
        return out

    def vcov_cpp(self, factor, facload, logvar, U, M, factors):
        """Same arguments as the Rcpp export (R/RcppExports.R:66); returns (T*M, M, nsave)."""
        lv = np.asfortranarray(logvar, dtype=np.float64)
        T, S, R = lv.shape
        factor = bool(factor)
        nf = int(factors) if factor else 0
        fl = np.asfortranarray(facload, dtype=np.float64) if factor and nf > 0 else None
        uv = np.asfortranarray(U, dtype=np.float64) if not factor and M > 1 else None
        out = np.empty((T * M, M, R), order="F")
        rc = self._lib.bvb_vcov(self._h, int(factor), T, int(M), nf, R, fptr(fl.reshape(-1, order="F")) if fl is not None else None,
                                fptr(lv.reshape(-1, order="F")), fptr(uv.reshape(-1, order="F")) if uv is not None else None,
                                fptr(out.reshape(-1, order="F")))
        self._check(rc)
        return out

    def insample(self, X, PHI, U, facload, logvar, prediction, factor, z=None):
        """Same arguments as the Rcpp export (R/RcppExports.R:62); returns y_pred (T, M, nsave).
        z (optional, parity tests): injected normals, shape (T, nsave, M)."""
        X = np.asfortranarray(X, dtype=np.float64)
        PHI = np.asfortranarray(PHI, dtype=np.float64)
        T, Kp = X.shape
        _, M, R = PHI.shape
        factor, prediction = bool(factor), bool(prediction)
        nf = (0 if facload is None or np.size(facload) == 0 else np.shape(facload)[1]) if factor else 0
        F = lambda a: fptr(np.asfortranarray(a, dtype=np.float64).reshape(-1, order="F"))
        fl = F(facload) if (prediction and nf) else None
        uv = F(U) if (prediction and not factor and M > 1) else None
        lv = F(logvar) if prediction else None
        zz = fptr(np.ascontiguousarray(z, dtype=np.float64).reshape(-1)) if (z is not None and prediction) else None
        out = np.empty((T, M, R), order="F")
        rc = self._lib.bvb_insample(self._h, T, Kp, M, R, nf, int(factor), int(prediction), F(X), F(PHI), uv, fl, lv, zz,
                                    fptr(out.reshape(-1, order="F")))
        self._check(rc)
        return out

    def sv_latent_draw(self, ystar, mixind, mu, phi, sigma, h0_var=-1.0, n_rep=1):
        """Test hook: n_rep draws of h_{0:T} | ystar, r, (mu, phi, sigma) from the series kernel's parallel state draw;
        returns (n_rep, T+1)."""
        import ctypes as C
        ystar = np.ascontiguousarray(ystar, dtype=np.float64)
        mixind = np.ascontiguousarray(mixind, dtype=np.uint8)
        T = ystar.size
        out = np.empty((int(n_rep), T + 1))
        rc = self._lib.bvb_sv_latent_draw(self._h, T, int(n_rep), fptr(ystar), mixind.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                          float(mu), float(phi), float(sigma), float(h0_var), fptr(out.reshape(-1)))
        self._check(rc)
        return out

    # -- a15: out_of_sample (src/utilities_cpp.cpp:216-393) ----------------------------
    def out_of_sample(self, each, X_T_plus_1, PHI, U, facload, logvar_T, ahead, sv_mu, sv_phi, sv_sigma,
                      sv_indicator, factor, LPL, Y_obs, LPL_subset, VoI, simulate_predictive,
                      z_logvar=None, z_pred=None):
        """Same argument order and 0-based index vectors as the Rcpp export
        (R/utility_functions.R:2469-2485); returns the four arrays of its List."""
        PHI = np.asfortranarray(PHI, dtype=np.float64)
        Kp, M, R = PHI.shape
        factor = bool(factor)
        nf = (0 if facload is None or np.size(facload) == 0 else np.shape(facload)[1]) if factor else 0
        fl = np.asfortranarray(facload, dtype=np.float64) if nf else None
        uv = None if factor or M < 2 else np.asfortranarray(np.asarray(U, dtype=np.float64).reshape(-1, R))
        lv = np.asfortranarray(np.asarray(logvar_T, dtype=np.float64).reshape(-1, R))
        ahead = np.ascontiguousarray(ahead, dtype=np.int32)
        svi = np.ascontiguousarray(sv_indicator if sv_indicator is not None else [], dtype=np.int32)
        nsv = svi.size
        mu = np.asfortranarray(np.asarray(sv_mu, dtype=np.float64).reshape(-1, R)) if nsv else None
        ph = np.asfortranarray(np.asarray(sv_phi, dtype=np.float64).reshape(-1, R)) if nsv else None
        sg = np.asfortranarray(np.asarray(sv_sigma, dtype=np.float64).reshape(-1, R)) if nsv else None
        x = np.ascontiguousarray(X_T_plus_1, dtype=np.float64).ravel()
        na, nsave = ahead.size, R * int(each)
        LPL, LPL_subset, simulate_predictive = bool(LPL), bool(LPL_subset), bool(simulate_predictive)
        yo = np.asfortranarray(Y_obs, dtype=np.float64) if LPL else None
        voi = np.ascontiguousarray(VoI if VoI is not None else [], dtype=np.int32)
        zl = None if z_logvar is None or not nsv else np.ascontiguousarray(z_logvar, dtype=np.float64)
        zp = None if z_pred is None else np.ascontiguousarray(z_pred, dtype=np.float64)
        if zl is not None:
            assert zl.shape == (R, int(ahead.max()), each, nsv)
        if zp is not None:
            assert zp.shape == (R, na, each, M)
        LPLd = np.zeros((na, nsave), order="F") if LPL else np.zeros((0, 0))
        PLu = np.zeros((na, M, nsave), order="F") if LPL else np.zeros((0, 0, 0))
        LPLs = np.zeros((na, nsave), order="F") if LPL and LPL_subset else np.zeros((0, 0))
        pred = np.zeros((na, M, nsave), order="F") if simulate_predictive else np.zeros((0, 0, 0))
        dp, ip = _lib._dp, _lib._ip

        def P(a):
            return None if a is None or a.size == 0 else a.ctypes.data_as(dp)

        def I(a):
            return None if a is None or a.size == 0 else a.ctypes.data_as(ip)

        rc = self._lib.bvb_predict(self._h, int(each), Kp, M, R, nf, lv.shape[0], int(factor), P(x), P(PHI), P(uv), P(fl),
                                   P(lv), I(ahead), na, P(mu), P(ph), P(sg), I(svi), nsv, int(LPL), P(yo),
                                   int(LPL_subset), I(voi), voi.size, int(simulate_predictive), P(zl), P(zp),
                                   P(LPLd), P(PLu), P(LPLs), P(pred))
        self._check(rc)
        return dict(LPL_draws=LPLd, PL_univariate_draws=PLu, LPL_sub_draws=LPLs, predictions=pred)
