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
/* Re-key the Philox generator of an existing handle (takes effect for the calls that
 * follow; call before bvb_gibbs_init for a new chain).  Lets a host keep ONE handle -
 * one CUDA context, one set of cached workspaces - per process, as the Rcpp shim does,
 * and still draw a fresh key from R's RNG at every entry point. */
int bvb_set_seed(bvb_handle* h, uint64_t seed);

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

/* ---- a1/a6-a10/a12/a13: bvar_cpp() - the Gibbs sampler (src/bvar_cpp.cpp:10-857) ----
 * The three structs mirror, key by key, the R lists bvar_cpp() receives
 * (prior_phi_cpp: R/utility_functions.R:1585-1617, prior_sigma_cpp: :374-420,
 * startvals: R/bvar_wrapper.R:352-385); the Rcpp shim fills them from the lists.
 * Vectors are (pointer, implied length); n = K*M coefficients without intercepts,
 * ordered as vec(PHI[0:K, ]) (column-major), n_U = M(M-1)/2 in trimatu_ind order. */
#define BVB_PRIOR_NORMAL 0
#define BVB_PRIOR_DL 1
#define BVB_PRIOR_GT 2   /* R2D2: kernel exponential, vs 1/2;  NG: kernel normal, vs 1 */
#define BVB_PRIOR_HS 3
#define BVB_PRIOR_SSVS 4
#define BVB_PRIOR_HMP 5
#define BVB_PRIOR_NOTHING 6 /* lags = 0 (R/utility_functions.R:2001) */

typedef struct bvb_prior_phi {
  int prior;
  const double* V_i;        /* [n]  prior "normal": fixed prior variances */
  int n_groups;
  const int* groups;        /* [n_groups] group ids as they appear in i_vec */
  const double* a;          /* [n_groups] */
  const double* b;          /* [n_groups] */
  const double* c;          /* [n_groups] */
  double GL_tol;
  const double* a_vec;      /* [grid_len] */
  const double* a_weight;   /* [grid_len] */
  const double* norm_consts;/* [grid_len] */
  const double* c_vec;      /* [grid_len] */
  int grid_len;
  int c_rel_a, DL_hyper, DL_plus, GT_hyper;
  double GT_vs;
  int GT_priorkernel;       /* 0 = "normal", 1 = "exponential" */
  const double* SSVS_tau0;  /* [n] */
  const double* SSVS_tau1;  /* [n] */
  double SSVS_s_a, SSVS_s_b;
  int SSVS_hyper;
  const double* SSVS_p;     /* [n] */
  const double* V_i_prep;   /* [n + M*intercept] (HMP) */
  double lambda_1[2];       /* HMP shape, rate */
  double lambda_2[2];
} bvb_prior_phi;

#define BVB_SIGMA_FACTOR 0
#define BVB_SIGMA_CHOLESKY 1

typedef struct bvb_prior_sigma {
  int type;                      /* BVB_SIGMA_FACTOR | BVB_SIGMA_CHOLESKY */
  /* cholesky_* keys: same meaning as bvb_prior_phi for the n_U free elements of U */
  int cholesky_U_prior;
  const double* cholesky_V_i;
  double cholesky_a, cholesky_b, cholesky_c, cholesky_GL_tol;
  const double* cholesky_a_vec; const double* cholesky_a_weight;
  const double* cholesky_norm_consts; const double* cholesky_c_vec;
  int cholesky_grid_len;
  int cholesky_c_rel_a, cholesky_DL_hyper, cholesky_DL_plus, cholesky_GT_hyper;
  double cholesky_GT_vs;
  int cholesky_GT_priorkernel;
  const double* cholesky_SSVS_tau0; const double* cholesky_SSVS_tau1;
  double cholesky_SSVS_s_a, cholesky_SSVS_s_b;
  int cholesky_SSVS_hyper;
  const double* cholesky_SSVS_p;
  double cholesky_lambda_3[2];
  const double* cholesky_priorhomoscedastic; /* M x 2 */
  const double* cholesky_sv_offset;          /* [M] */
  /* factor_* keys */
  int factor_factors;
  const double* factor_shrink_a;  /* [M] rowwise | [r] columnwise */
  const double* factor_shrink_c;
  const double* factor_shrink_d;
  int factor_ngprior, factor_columnwise;
  double factor_facloadtol;
  const int* factor_restrinv;     /* M x r, 1 = unrestricted */
  const double* factor_priorhomoskedastic; /* M x 2 */
  int factor_interweaving;
  /* sv_* keys */
  const int* sv_heteroscedastic;  /* [M (+ r)] */
  double sv_priormu[2];           /* mean, sd */
  double sv_priorphi[4];          /* idi a,b (, fac a,b) */
  const double* sv_priorsigma2;   /* (M + r) x 2: shape, rate */
  const double* sv_priorh0;       /* [M (+ r)]  <= 0: stationary */
} bvb_prior_sigma;

typedef struct bvb_startvals {
  const double* PHI;         /* Kp x M */
  const double* U;           /* M x M (cholesky) */
  const double* sv_para;     /* 3 x (M + r) */
  const double* sv_logvar;   /* T x (M + r) */
  const double* sv_logvar0;  /* [M + r] */
  const double* facload;     /* M x r */
  const double* fac;         /* r x T */
  const double* tau2;        /* M x r */
} bvb_startvals;

/* Caller-owned draw storage, laid out exactly like the arrays bvar_cpp() returns
 * (src/bvar_cpp.cpp:409-482, :751-836); nsave = draws / thin.  NULL = not wanted. */
typedef struct bvb_draws {
  double* PHI;                /* Kp x M x nsave */
  double* U;                  /* n_U x nsave */
  double* logvar;             /* (T | 1) x (M + r) x nsave */
  double* sv_para;            /* 3 x (M + r) x nsave */
  double* phi_hyperparameter; /* size x nsave (rows as bvar_cpp.cpp:758-790) */
  double* u_hyperparameter;
  double* V_prior;            /* Kp x M x nsave */
  double* facload;            /* M x r x nsave */
  double* fac;                /* r x (T | 1) x nsave */
} bvb_draws;

/* Set up a chain on the handle's device.  Argument meaning = bvar_cpp()'s
 * (src/bvar_cpp.cpp:10-30); tvp_keep_all: 1 = "all", 0 = "last".  i_vec is the
 * vectorised indicator matrix ((K + intercept) * M ints; 0 = intercept).  Everything
 * is copied to the device; the caller's inputs are never modified.
 * Multi-GPU (factor structure): call bvb_comm_init first; equations are then
 * sharded over the ranks, every rank passes identical arguments, and rank 0's
 * bvb_draws receives the stored draws. */
int bvb_gibbs_init(bvb_handle* h, const double* Y, const double* X, int M, int T, int K,
                   int draws, int burnin, int thin, int tvp_keep_all, int intercept,
                   const double* priorIntercept, const double* PHI0,
                   const bvb_prior_phi* prior_phi, const bvb_prior_sigma* prior_sigma,
                   const bvb_startvals* startvals, const int* i_vec,
                   double PHI_tol, double L_tol, int huge, const bvb_draws* out);

/* Run up to max_sweeps further sweeps (the shim calls this with 256 so it can
 * poll R interrupts and print progress like bvar_cpp.cpp:501-503,838-843).
 * *done receives the number of sweeps completed so far (== draws + burnin at the
 * end).  Stored draws are written into the bvb_draws given to bvb_gibbs_init. */
int bvb_gibbs_run(bvb_handle* h, int max_sweeps, int* done);

/* Device time of the graph-replayed sweeps of the last bvb_gibbs_run call: out[0]
 * total ms (CUDA events on the sampler stream), out[1] = kernels launched per sweep,
 * out[2..9] = ms of {prep, K1, K2, comm, clamp + hyper, resid, Sigma_t, store} as
 * measured by the last bvb_gibbs_profile_sweep (0 before one ran).  Returns the
 * number of sweeps out[0] covers (the eager first sweep of a chain is excluded). */
int bvb_gibbs_timing(const bvb_handle* h, double* out10);

/* One further sweep of the chain (counted and stored like any other) launched in its
 * serial form with CUDA events between the phases; out8 = ms of the eight phases
 * above.  bvb_gibbs_run itself never instruments a sweep, so callers decide where
 * this lies relative to their own timed regions. */
int bvb_gibbs_profile_sweep(bvb_handle* h, double* out8, int* done);

/* Live kernel timers: K1 (pair-GEMM Z tiles) and K2 (Cholesky / solve / draw) stamp
 * the device %globaltimer of their first CTA start and last CTA end in every sweep
 * (graph replay included); out3 = {K1 ms, K2 ms, sweeps} accumulated since the last
 * reset (reset != 0 clears the accumulators after reading). */
int bvb_gibbs_kernel_times(bvb_handle* h, double* out3, int reset);

/* ---- a12: the tolerance clamp of src/bvar_cpp.cpp:534-559 on host arrays ---------
 * PHI_io / PHI0 are Kp x M; on the K lag rows |PHI - PHI0| is pushed up to tol and an
 * exact zero becomes +tol or -tol with probability 1/2 (R::rbinom(1, 0.5) there, the
 * Philox stream of (handle seed, element, sweep) here).  The Gibbs sweep runs the same
 * kernel on device state; this entry point exists so that the exact-zero branch, which
 * a chain never reaches, can be tested. */
int bvb_tolerance_clamp(bvb_handle* h, int K, int Kp, int M, double* PHI_io,
                        const double* PHI0, double tol, uint32_t sweep);

/* Copy the current sampler state back (for tests): any pointer may be NULL. */
int bvb_gibbs_state(bvb_handle* h, double* PHI, double* V_prior, double* logvar,
                    double* sv_para, double* facload, double* fac, double* U);

/* Multi-GPU plumbing: one process per GPU.  unique_id is the 128-byte
 * ncclUniqueId produced by bvb_comm_unique_id on rank 0 and broadcast by the
 * launcher (torch.distributed / MPI / a file). */
/* (The per-sweep exchange of the freshly drawn PHI columns is a push over NVLink peer memory fused into the solve kernel,
 * with CUDA IPC handles exchanged once through the communicator; the communicator's allgather is the fallback when the
 * peers' buffers cannot be opened - BVB_NCCL_EXCHANGE=1 forces it.) */
int bvb_comm_unique_id(char* unique_id_128);
int bvb_comm_init(bvb_handle* h, int rank, int world, const char* unique_id_128);

/* One process, several GPUs (what an R session is: no launcher, no MPI).  bvb_create_multi makes n handles on the
 * given devices with ranks 0..n-1 and an in-process NCCL communicator; bvb_gibbs_init_multi sets the same chain up on
 * every rank (identical arguments, rank 0's bvb_draws receives the stored draws); bvb_gibbs_run_multi advances all
 * ranks together (one host thread per rank for the duration of the call) - the multi-GPU form of the bvar_cpp() loop
 * (src/bvar_cpp.cpp:498-845) behind the same chunked, interruptible call as bvb_gibbs_run.  n = 1 is the plain path. */
int bvb_create_multi(int n, const int* devices, uint64_t seed, bvb_handle** handles);
int bvb_gibbs_init_multi(int n, bvb_handle** handles, const double* Y, const double* X, int M, int T, int K,
                         int draws, int burnin, int thin, int tvp_keep_all, int intercept,
                         const double* priorIntercept, const double* PHI0,
                         const bvb_prior_phi* prior_phi, const bvb_prior_sigma* prior_sigma,
                         const bvb_startvals* startvals, const int* i_vec,
                         double PHI_tol, double L_tol, int huge, const bvb_draws* out);
int bvb_gibbs_run_multi(int n, bvb_handle** handles, int max_sweeps, int* done);
void bvb_destroy_multi(int n, bvb_handle** handles);

/* ---- a14: irf_cpp (src/irf.cpp:468-524) -------------------------------------------
 * Impulse responses for R stored posterior draws.  coefficients Kp x M x R;
 * factor_loadings M x n_factors x R (n_factors = 0: NULL); U_vecs n_U x R
 * (has_U = 0: NULL); logvar_t ld_logvar x R (log-variances at the chosen time
 * point); shocks dim x n_shocks with dim = n_factors (factor), M (otherwise);
 * rotation dim x dim x R or NULL.  out: M x n_shocks x (ahead+1) x R - the array
 * R/irf.R:275-276 builds from the returned field<cube>.  Draws are sharded over
 * the ranks of bvb_comm_init by the caller (disjoint slices, no collective). */
int bvb_irf(bvb_handle* h, int Kp, int M, int R, int n_factors, int has_U, int ld_logvar,
            int dim, int n_shocks, int ahead, const double* coefficients,
            const double* factor_loadings, const double* U_vecs, const double* logvar_t,
            const double* shocks, const double* rotation, double* out);

/* ---- a15: out_of_sample (src/utilities_cpp.cpp:216-393) ---------------------------
 * Predictive simulation and log predictive likelihoods over R stored draws, `each`
 * replicates per draw (chain (r, j) writes slice r + j*R, as :352/:364/:376).
 * PHI Kp x M x R; U n_U x R (cholesky) or facload M x n_factors x R (factor);
 * logvar_T ld_logvar x R (log-variances at T, idiosyncratic then factor);
 * ahead[n_ahead] strictly increasing horizons (1-based); sv_mu/phi/sigma n_sv x R
 * and sv_indicator[n_sv] (0-based rows of logvar_T that follow an SV process);
 * Y_obs n_ahead x M (LPL only); VoI[n_voi] 0-based (LPL_subset only).
 * z_logvar / z_pred: NULL = counter-based device normals; otherwise the standard
 * normals the reference takes from R::norm_rand, C-ordered [R][max_ahead][each][n_sv]
 * (:325-327) and [R][n_ahead][each][M] (:371-374) - used by the parity tests.
 * Outputs (column-major, any unused one may be NULL): LPL_draws n_ahead x nsave,
 * PL_univariate_draws n_ahead x M x nsave, LPL_sub_draws n_ahead x nsave,
 * predictions n_ahead x M x nsave, nsave = R*each. */
int bvb_predict(bvb_handle* h, int each, int Kp, int M, int R, int n_factors, int ld_logvar,
                int factor, const double* X_T_plus_1, const double* PHI, const double* U,
                const double* facload, const double* logvar_T, const int* ahead, int n_ahead,
                const double* sv_mu, const double* sv_phi, const double* sv_sigma,
                const int* sv_indicator, int n_sv, int LPL, const double* Y_obs,
                int LPL_subset, const int* VoI, int n_voi, int simulate_predictive,
                const double* z_logvar, const double* z_pred, double* LPL_draws,
                double* PL_univariate_draws, double* LPL_sub_draws, double* predictions);

/* ---- section 8(f) rank 2: predh (src/utilities_cpp.cpp:181-212) ----------------------
 * Log-variance forecasts h_{T+ahead} | h_T of the AR(1) SV processes.  logvar_T, sv_mu, sv_phi, sv_sigma M x R
 * (column-major, one column per stored draw); ahead[n_ahead] strictly increasing (1-based).  z: NULL = counter-based
 * device normals, otherwise the normals of `norm_rand.imbue(R::norm_rand)` (:202-203), C-ordered
 * [n_ahead][each][R][M].  out: n_ahead x M x (R*each) column-major; replicate j fills slices j*R .. (j+1)*R-1 (:205). */
int bvb_predh(bvb_handle* h, int M, int R, int each, const int* ahead, int n_ahead, const double* logvar_T,
              const double* sv_mu, const double* sv_phi, const double* sv_sigma, const double* z, double* out);

/* ---- section 8(f) rank 2: vcov_cpp (src/utilities_cpp.cpp:436-472) -------------------
 * Posterior draws of Sigma_t.  logvar T x (M + n_factors) x R (factor) or T x M x R (cholesky); facload
 * M x n_factors x R; U n_U x R (upper-triangular elements in trimatu_ind order).  out: (T*M) x M x R column-major,
 * slice i = the cube (T, M, M) flattened as the reference does (:448-449): element (t, a, b) at row t + T*a, column b. */
int bvb_vcov(bvb_handle* h, int factor, int T, int M, int n_factors, int R, const double* facload,
             const double* logvar, const double* U, double* out);

/* ---- section 8(f) rank 2: insample (src/utilities_cpp.cpp:396-433, predict_y :69-93) -----
 * In-sample fit x_t' PHI_r for every t and stored draw r, plus - prediction != 0 - the error term z' chol_upper(Sigma_tr).
 * X T x Kp; PHI Kp x M x R; U n_U x R (cholesky); facload M x n_factors x R (factor); logvar T x (M + n_factors | M) x R.
 * z: NULL = counter-based device normals, otherwise the normals the reference takes from R::norm_rand, C-ordered
 * [T][R][M] (its loop order: t outer, draw inner, :417-418).  out: T x M x R column-major (the returned cube). */
int bvb_insample(bvb_handle* h, int T, int Kp, int M, int R, int n_factors, int factor, int prediction,
                 const double* X, const double* PHI, const double* U, const double* facload,
                 const double* logvar, const double* z, double* out);

/* ---- test hook: the latent-state draw of the SV update ---------------------------------
 * n_rep independent draws of h_{0:T} | ystar_{1:T}, mixture indicators r_{1:T} (0..9), mu, phi, sigma from the parallel
 * ("perturb, then solve") form of the all-without-a-loop sampler the series kernel runs; the reference takes this step
 * from stochvol's update_fast_sv (called at src/bvar_cpp.cpp:761-777).  h0_var <= 0: stationary prior on h_0.
 * out: n_rep x (T+1), replica-major.  The tests compare sample mean / covariance with Omega^-1 c and Omega^-1. */
int bvb_sv_latent_draw(bvb_handle* h, int T, int n_rep, const double* ystar, const unsigned char* mixind, double mu,
                       double phi, double sigma, double h0_var, double* out);

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
