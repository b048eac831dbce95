/*
 * bvb.h - C ABI of libbvb.so: the B200 (sm_100a) implementation of the
 * bayesianVARs Gibbs coefficient-draw hot path.
 *
 * Drop-in boundary: the reference's Rcpp entry points
 *   _bayesianVARs_bvar_cpp            (src/RcppExports.cpp:16-44  -> src/bvar_cpp.cpp:10-30)
 *   _bayesianVARs_sample_PHI_cholesky (src/RcppExports.cpp:224-236 -> src/sample_coefficients.cpp:149-157)
 *   _bayesianVARs_my_gig              (src/RcppExports.cpp:46-58  -> src/gig_vec.cpp:20-43)
 *   _bayesianVARs_irf_cpp             (src/RcppExports.cpp:148-163 -> src/irf.cpp:468-524)
 *   _bayesianVARs_out_of_sample       (src/RcppExports.cpp:165-191 -> src/utilities_cpp.cpp:216-393)
 * keep their R-visible signatures; an Rcpp shim (r_shim/)
 * unpacks the SEXP arguments and forwards to the functions declared here.
 *
 * Conventions
 *  - every matrix is IEEE double, column-major, dense, leading dimension = rows
 *    (exactly what REAL(x) of an R matrix / arma::mat::memptr() points to);
 *  - all pointers are HOST pointers; inputs are read-only, outputs are
 *    caller-owned and never freed or reallocated by the library;
 *  - no exceptions cross the ABI: functions return BVB_OK or an error code and
 *    bvb_last_error() gives step / sweep / equation and a message so the shim
 *    can re-raise with the reference's wording (bvar_cpp.cpp:512,522,530;
 *    sample_coefficients.cpp:106,140,178);
 *  - there is no CPU fallback: without a usable CUDA device bvb_create fails.
 */
#ifndef BVB_H
#define BVB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVB_OK 0
#define BVB_ERR_CUDA 1        /* CUDA runtime / driver error, no device */
#define BVB_ERR_ARG 2         /* invalid argument */
#define BVB_ERR_NUMERIC 3     /* Cholesky pivot <= 0 or non-finite draw */
#define BVB_ERR_UNSUPPORTED 4 /* shape outside what the kernels were built for */
#define BVB_ERR_COMM 5        /* NCCL / peer-access failure */

typedef struct bvb_handle bvb_handle;

/* Library / device lifetime.  `device` is the CUDA ordinal this handle (one per
 * process / rank) drives.  `seed` keys the Philox4x32-10 counter-based RNG; the
 * Rcpp shim draws it from R's RNG so set.seed() keeps runs reproducible. */
int bvb_create(bvb_handle** out, int device, uint64_t seed);
void bvb_destroy(bvb_handle* h);
const char* bvb_version(void);

/* Last error of this handle.  step: 1 = PHI draw, 2 = hyper-parameters,
 * 3 = Sigma_t update, 4 = irf, 5 = predict; sweep and equation are 1-based as
 * in the reference's messages (0 = not applicable).  Returns the error code. */
int bvb_last_error(const bvb_handle* h, int* step, int* sweep, int* equation,
                   char* msg, int msg_len);

/* ---- a2/a3: sample_PHI_factor (src/sample_coefficients.cpp:84-114) ----------
 * One conditional draw of all M columns of PHI under the factor error
 * structure.  X is T x Kp (intercept last column), Y T x M, logvar_idi T x M
 * (idiosyncratic log-variances), V_prior / PHI0 Kp x M, facload M x r, fac r x T
 * (r may be 0, then both may be NULL).
 * z: Kp x M standard normals consumed exactly as the reference consumes
 *    R::norm_rand (column j by equation j); NULL -> drawn on device (Philox).
 * delta: T x M, only read when huge != 0 (the T normals of
 *    sample_coefficients.cpp:47); NULL -> drawn on device.
 * PHI_out: Kp x M. */
int bvb_phi_draw_factor(bvb_handle* h, int T, int Kp, int M, int r,
                        const double* X, const double* Y, const double* logvar_idi,
                        const double* V_prior, const double* PHI0,
                        const double* facload, const double* fac,
                        const double* z, const double* delta, int huge,
                        double* PHI_out);

/* ---- a4/a4x: sample_PHI_cholesky (src/sample_coefficients.cpp:118-157) -------
 * Corrected triangular algorithm: equations are drawn in order, equation j sees
 * the already-updated columns < j.  PHI_in is the current value (not modified),
 * U is M x M upper unitriangular, d_sqrt T x M.  z as above (NULL -> Philox). */
int bvb_phi_draw_cholesky(bvb_handle* h, int T, int Kp, int M,
                          const double* PHI_in, const double* PHI0,
                          const double* Y, const double* X, const double* U,
                          const double* d_sqrt, const double* V_prior,
                          const double* z, double* PHI_out);

/* ---- a5: sample_U (src/sample_coefficients.cpp:160-188) ----------------------
 * resid T x M, V_i_U length M(M-1)/2 (trimatu_ind order), z length M(M-1)/2 in
 * the reference's consumption order (NULL -> Philox).  U_out M x M. */
int bvb_sample_u(bvb_handle* h, int T, int M, const double* resid,
                 const double* V_i_U, const double* d_sqrt, const double* z,
                 double* U_out);

/* ---- a11: my_gig (src/gig_vec.cpp:20-43) -------------------------------------
 * out is n x m column-major with m = max(len_lambda, len_chi, len_psi) and the
 * three parameter vectors recycled to length m (rep_len); density
 * x^(lambda-1) exp(-(chi/x + psi x)/2).  `stream` offsets the Philox counter so
 * successive calls give fresh variates. */
int bvb_gig(bvb_handle* h, int n, const double* lambda, int len_lambda,
            const double* chi, int len_chi, const double* psi, int len_psi,
            uint64_t stream, double* out);

/* ---- diagnostics -------------------------------------------------------------
 * Measured FP64 peaks on this device (own micro-kernels, register-resident):
 * out[0] = DMMA m8n8k4 TFLOP/s, out[1] = DFMA TFLOP/s, out[2] = SM clock MHz
 * seen by the kernel (cycles / elapsed). */
int bvb_measure_fp64_peak(bvb_handle* h, double* out3);

/* Device time in ms of the kernels launched by the last API call on this
 * handle: out[0] = prep (weights/pairs), out[1] = K1 pair-GEMM, out[2] = K2
 * Cholesky/solve/draw, out[3] = rest.  Returns number of kernel launches. */
int bvb_last_timing(const bvb_handle* h, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* BVB_H */
