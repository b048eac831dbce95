// Cholesky error structure (Sigma_t = U^-T D_t U^-1): device context + launchers of
// the triangular-algorithm PHI draw (sample_coefficients.cpp:118-146) and of
// sample_U (sample_coefficients.cpp:160-188).
#pragma once
#include "bvb_internal.cuh"

namespace bvb {

struct CholPhiCtx {
  int T = 0, Tp = 0, Kp = 0, M = 0;
  // inputs
  const double* X = nullptr;      // T x Kp
  const double* Y = nullptr;      // T x M
  const double* PHI0 = nullptr;   // Kp x M
  const double* V = nullptr;      // Kp x M prior variances
  const double* U = nullptr;      // M x M upper unitriangular
  const double* dsq = nullptr;    // T x M   d_sqrt
  const double* Zn = nullptr;     // Kp x M  standard normals
  double* PHI = nullptr;          // Kp x M  in: current value, out: new draw
  // pair-GEMM / factorisation state
  const double* A = nullptr;      // pair-product matrix of X
  const int2* lut = nullptr;
  const int* colbase = nullptr;
  int nZtiles = 0, ntiles = 0;
  long ldo = 0;
  double* W = nullptr;            // [Nalloc][Tp]
  const double* WYzero = nullptr; // [Nalloc][Tp] all zero
  double* omega = nullptr;        // [M][ldo]
  // scratch
  double* d2inv = nullptr;        // T x M
  double* resid = nullptr;        // T x M
  double* XPHI = nullptr;         // T x M
  double* Q = nullptr;            // T x M  (Y - X PHI) U
  double* rvec = nullptr;         // [T]
  double* bvec = nullptr;         // [Kp]
  int* status = nullptr;          // [M]
};

// One sequential sweep over the M equations (all launches on `s`).
cudaError_t launch_chol_phi(const CholPhiCtx& c, cudaStream_t s);

struct CholUCtx {
  int T = 0, Tp = 0, M = 0, nu = 0;   // nu = M - 1 regressors (padded systems)
  const double* resid = nullptr;  // T x M
  const double* dsq = nullptr;    // T x M
  const double* V_i_U = nullptr;  // [M(M-1)/2]
  const double* zflat = nullptr;  // [M(M-1)/2] standard normals in the reference's consumption order
  double* U = nullptr;            // M x M out
  // pair-GEMM state for X = resid[:, 0:nu]
  double* A = nullptr;
  const int2* lut = nullptr;
  const int2* pairs = nullptr;
  const int* colbase = nullptr;
  int nZtiles = 0, ntiles = 0, nArows = 0;
  long ldo = 0;
  double *W = nullptr, *WY = nullptr, *omega = nullptr, *Vu = nullptr, *Zu = nullptr, *phi = nullptr;
  int* status = nullptr;          // [M-1]
};
cudaError_t launch_sample_u(const CholUCtx& c, cudaStream_t s);

// resid = Y - X PHI  (shared with the Gibbs loop)
cudaError_t launch_resid(double* resid, const double* Y, const double* X, const double* PHI, int T, int Kp, int M,
                         cudaStream_t s);

}  // namespace bvb
