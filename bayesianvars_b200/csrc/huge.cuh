// expert_huge branch: device context + launchers (see huge.cu).
#pragma once
#include "bvb_internal.cuh"

namespace bvb {

struct HugeCtx {
  int T = 0, Kp = 0, Kp_pad = 0, M = 0, r = 0;
  // inputs
  const double* X = nullptr;        // T x Kp
  const double* Y = nullptr;        // T x M
  const double* logvar = nullptr;   // T x M idiosyncratic log-variances
  const double* V = nullptr;        // Kp x M
  const double* PHI0 = nullptr;     // Kp x M
  const double* facload = nullptr;  // M x r
  const double* fac = nullptr;      // r x T
  const double* z1 = nullptr;       // Kp x M normals (sample_coefficients.cpp:42)
  const double* delta = nullptr;    // T x M normals (:47)
  double* PHI = nullptr;            // Kp x M out
  // pair-GEMM over observation pairs
  const double* A = nullptr;        // [nArows][Kp_pad]
  const int2* lut_h = nullptr;
  const int2* pairs_h = nullptr;
  const int* colbase_h = nullptr;
  int nZtiles = 0, ntiles = 0;
  long ldo = 0;
  double* Wv = nullptr;             // [Nalloc][Kp_pad] prior variances (B operand)
  const double* WYzero = nullptr;   // [Nalloc][Kp_pad] zeros
  double* omega = nullptr;          // [M][ldo] packed T x T systems
  // scratch
  double* u = nullptr;              // Kp x M
  double* nrm = nullptr;            // T x M
  double* wsol = nullptr;           // T x M
  const double* zero_tm = nullptr;  // T x M zeros
  int* status = nullptr;
};

cudaError_t launch_huge_setup(double* XT, double* A, const double* X, const int2* pairs, int T, int Kp, int Kp_pad,
                              int nArows, cudaStream_t s);
cudaError_t launch_phi_huge(const HugeCtx& c, cudaStream_t s);

}  // namespace bvb
