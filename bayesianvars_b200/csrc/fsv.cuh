// Device-side state of the Sigma_t block (factor SV or M univariate SVs).
#pragma once
#include "bvb_internal.cuh"
#include "gig.cuh"
#include "sv_math.cuh"

namespace bvb {

constexpr int FSV_RMAX = 16;     // factors handled by the per-thread (register / local) small systems
constexpr int FSV_RMAX_BIG = 64; // beyond FSV_RMAX the r x r systems live in shared memory, one CTA per row / time point
                                 // (the reference's own benchmark runs r = 50, vignettes/scalability-vignette.qmd:95)

struct FsvDev {
  int M = 0, T = 0, r = 0;
  // inputs
  const double* resid = nullptr;    // T x M   (VAR residuals; orthogonalised residuals for cholesky)
  // state
  double* logvar = nullptr;         // T x (M+r)
  double* logvar0 = nullptr;        // [M+r]
  double* sv_para = nullptr;        // 3 x (M+r): mu, phi, sigma
  double* facload = nullptr;        // M x r
  double* fac = nullptr;            // r x T
  double* tau2 = nullptr;           // M x r
  double* lambda2 = nullptr;        // [M] rowwise | [r] columnwise
  // scratch (series-major)
  double* ystar = nullptr;          // [S][T]
  double* eps = nullptr;            // [S][T+1]
  unsigned char* mixind = nullptr;  // [S][T]
  double* scratch = nullptr;        // unused by the CTA-per-series kernel
  double* eidi = nullptr;           // T x M idiosyncratic residuals (series-major)
  double* wexp = nullptr;           // T x M exp(-h) of the idiosyncratic series (refreshed by the series kernel)
  double* Wk1 = nullptr;            // [M][Tp] weight operand of the pair-GEMM (written by the series kernel when set)
  int Tp = 0;
  int* status = nullptr;
  int prof = 0;                     // BVB_FSV_PROF: clock64 phase profile of the series kernel (series 0)
  // configuration
  const int* hetero = nullptr;      // [M+r]
  const int* restr = nullptr;       // M x r, 1 = unrestricted
  const double* offset = nullptr;   // [M] sv offset
  const double* priorsigma2 = nullptr;  // (M+r) x 2
  const double* priorh0 = nullptr;      // [M+r]
  const double* priorhomo = nullptr;    // M x 2
  const double* shrink_a = nullptr; const double* shrink_c = nullptr; const double* shrink_d = nullptr;
  double priormu[2] = {0, 10};
  double priorphi[4] = {10, 3, 10, 3};
  double facloadtol = 1e-14;
  int ngprior = 1, columnwise = 0, interweaving = 4;
};

cudaError_t launch_fsv_update(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s, cudaEvent_t ng_done = nullptr);
cudaError_t launch_fsv_series(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s);
cudaError_t launch_fsv_rest(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s, cudaEvent_t ng_done = nullptr);
bool fsv_has_ng(const FsvDev& d);
void fsv_print_profile();
cudaError_t launch_sv_latent_test(int T, int n_rep, const double* ystar, const unsigned char* mixind, double mu, double phi, double sigma,
                                  double h0_var, uint64_t seed, double* out, cudaStream_t s);
cudaError_t launch_fsv_ng(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s);

}  // namespace bvb
