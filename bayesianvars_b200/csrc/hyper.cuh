// Device-side description of one hierarchical shrinkage prior block (the PHI
// coefficients, or the free elements of U under the cholesky structure) and the
// launchers of the K3 kernels that update it.
//
// Mirrors sample_V_i_DL / _GT / _HS / _SSVS_beta / _HMP / _U_HMP
// (src/sample_coefficients.cpp:190-377) and the dispatch in
// src/bvar_cpp.cpp:563-609 (PHI) / :688-721 (U).
#pragma once
#include "bvb_internal.cuh"
#include "gig.cuh"

namespace bvb {

constexpr int HYP_NRED = 2;       // reductions per group per pass
constexpr int HYP_THREADS = 256;

// per-group scalars (device array gpar[G][HYP_GP])
enum { GP_A = 0, GP_B = 1, GP_C = 2, GP_XI = 3, GP_ZETA = 4, GP_VARPI = 5, GP_AUX = 6, GP_N = 7, HYP_GP = 8 };

struct HyperDev {
  int prior = BVB_PRIOR_NORMAL;
  int n = 0, G = 0, nblocks = 0;
  int* grp = nullptr;          // [n] group index (0..G-1) or -1
  double* coef = nullptr;      // [n] coefficient minus prior mean (clamped for DL/GT)
  double* V_i = nullptr;       // [n]
  double* loc1 = nullptr;      // [n] lambda (DL/GT) | theta (HS) | gammas (SSVS)
  double* loc2 = nullptr;      // [n] psi    (DL/GT) | nu    (HS) | p_i    (SSVS)
  double* gpar = nullptr;      // [G][HYP_GP]
  // any number of groups (global_grouping = "equation-wise" / "covariate-wise" give M / K of them): the members of
  // group g are gmem[gstart[g] .. gstart[g+1]) in ascending coefficient order, one CTA per group reduces them in a
  // fixed order
  const int* gstart = nullptr; // [G + 1]
  const int* gmem = nullptr;   // [number of grouped coefficients]
  double* red = nullptr;       // [HYP_NRED][n] per-coefficient terms of the group sums
  double* gsum = nullptr;      // [G][HYP_NRED]   reduced sums of the last pass
  double* gpart = nullptr;     // [G][nsplit] partial sums of a split group reduction
  int* gcnt = nullptr;         // [G] CTAs of a split reduction that have arrived (self-resetting)
  int nsplit = 1;              // CTAs per group in the reductions (few large groups: up to 32)
  // grids for the hyper-hyper update of `a` (DL/GT)
  double *a_vec = nullptr, *a_weight = nullptr, *norm_consts = nullptr, *c_vec = nullptr;
  int grid_len = 0;
  double tol = 0.0, vs = 1.0;
  int hyper = 0, dl_plus = 0, c_rel_a = 0, kernel_exp = 0;
  // SSVS
  double *tau0 = nullptr, *tau1 = nullptr;
  double s_a = 1.0, s_b = 1.0;
  // HMP: groups are 0 = own-lag, 1 = cross-lag; gpar[g][GP_A..GP_B] = shape, rate
  double* V_prep = nullptr;    // [n]
  uint32_t stream_base = 0;    // Philox stream ids stream_base .. stream_base+3
  int always_active = 0;       // U block: SSVS / HMP updates run from the first sweep (bvar_cpp.cpp:708-716)
  int* status = nullptr;       // set to 1 when a GIG rejection loop is exhausted (the variate is NaN, never a rejected candidate)
};

// One full hyper-parameter update (up to four kernels).  The sweep index is read
// from device memory (the launches are replayed from a CUDA graph); SSVS and HMP
// only start after 10% of the burn-in (bvar_cpp.cpp:585,602).
cudaError_t launch_hyper_update(const HyperDev& hd, uint64_t seed, const uint32_t* sweep, int burnin, cudaStream_t s);

}  // namespace bvb
