// Gibbs sampler orchestration on the device: the loop body of bvar_cpp()
// (src/bvar_cpp.cpp:498-845) as one CUDA graph per sweep, with every piece of
// sampler state resident in HBM and stored draws staged through a device ring
// and pinned host memory into the caller's (R-owned) arrays.
//
// One sweep, factor structure (bvar_cpp.cpp:519-647):
//   prep (W = exp(-h), WY)  -> K1 pair-GEMM -> normals -> K2 Cholesky/solve/draw
//   [-> NCCL allgather of the PHI columns when equations are sharded]
//   -> tolerance clamp (:534-559) -> K3 hyper-parameters (:563-609)
//   -> residuals Y - X PHI (:611) -> K8 factor-SV update (:617-647) -> store (:751-836)
// The sweep counter lives in device memory so the captured graph is replayed
// unchanged; all Philox streams are keyed by (purpose, element, sweep).
#include "gibbs.cuh"
#include <chrono>
#include <thread>
#include "chol.cuh"

#include <nccl.h>
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <new>

using namespace bvb;

// NCCL is bound at run time (dlopen), not at link time: a process that also
// imports PyTorch already has PyTorch's bundled libnccl.so.2 loaded, and a second,
// older system copy pulled in through DT_NEEDED would shadow its symbols.
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api;
  if (api.lib) return api;
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!api.lib) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
  api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
  return api;
}
}  // namespace

namespace bvb {

// Philox stream ids: one named constant per purpose (hyper-parameter blocks use base .. base + 3)
constexpr uint32_t STREAM_PHI_Z = 0x10, STREAM_CLAMP = 0x11, STREAM_U_Z = 0x13, STREAM_HUGE_DELTA = 0x14, STREAM_CLAMP_U = 0x15,
                   STREAM_PHI_HYPER = 0x20, STREAM_U_HYPER = 0x30;

// ------------------------------------------------------------------ kernels
__global__ void gibbs_normals_kernel(double* out, int n, int M, int ld, int j0, uint64_t seed, const uint32_t* sweep) {
  // z[i, j] for the local equations j0..j0+M-1; stream element = global (i, j) so
  // the draw does not depend on how equations are sharded
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * M) return;
  const int j = (int)(idx / n), i = (int)(idx % n);
  RngStream rng(seed, STREAM_PHI_Z, (uint32_t)((size_t)(j0 + j) * n + i), *sweep);
  out[(size_t)j * ld + i] = rng.normal();
}

// Start of a factor-structure sweep, one launch: blocks [0, M) build the GEMM's B operands W = exp(-h), WY = W * yhat
// (prep.cu: prep_factor_kernel) for every equation, the remaining blocks draw the normals of the local equations.
__global__ void __launch_bounds__(256) gibbs_prep_normals_kernel(double* __restrict__ W, double* __restrict__ WY,
                                                                 const double* __restrict__ Y, const double* __restrict__ logvar,
                                                                 const double* __restrict__ facload, const double* __restrict__ fac,
                                                                 int T, int Tp, int M, int r, double* zout, int n, int Mloc,
                                                                 int j0, uint64_t seed, const uint32_t* sweep, int write_w) {
  if ((int)blockIdx.x < M) {
    const int j = blockIdx.x;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      double yh = Y[(size_t)j * T + t];
      for (int k = 0; k < r; ++k) yh -= facload[(size_t)k * M + j] * fac[(size_t)t * r + k];
      const double nrm = exp(-0.5 * logvar[(size_t)j * T + t]);   // squared like the reference's X_norm'X_norm
      const double w = nrm * nrm;
      if (write_w) W[(size_t)j * Tp + t] = w;   // (0: the SV series kernel wrote it and the Z tiles may be reading it)
      WY[(size_t)j * Tp + t] = w * yh;
    }
    return;
  }
  const size_t idx = (size_t)(blockIdx.x - M) * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * Mloc) return;
  const int j = (int)(idx / n), i = (int)(idx % n);
  RngStream rng(seed, STREAM_PHI_Z, (uint32_t)((size_t)(j0 + j) * n + i), *sweep);
  zout[(size_t)j * n + i] = rng.normal();
}

// PHI_diff = PHI - PHI0 on the K lag rows, pushed away from zero for DL / GT
// (bvar_cpp.cpp:534-559); also the non-finite check of :529-531.
//
// Thread 0 also closes the live kernel timers of this sweep: K1 (Z tiles) and K2 stamped the device-wide
// %globaltimer of their first CTA start / last CTA end into kt[0..3] (atomicMin / atomicMax); both are complete here
// and the next Z tiles cannot start before the SV series kernel, so the spans are added to kt[4] / kt[5] and re-armed.
__global__ void gibbs_clamp_kernel(double* PHI, const double* PHI0, double* coef, int K, int Kp, int M,
                                   int do_clamp, double tol, uint64_t seed, const uint32_t* sweep, int* status,
                                   unsigned long long* kt) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && kt) {
    if (kt[1] > kt[0]) kt[4] += kt[1] - kt[0];
    if (kt[3] > kt[2]) kt[5] += kt[3] - kt[2];
    if (kt[3] > kt[2]) kt[6] += 1;
    kt[0] = ~0ull; kt[1] = 0; kt[2] = ~0ull; kt[3] = 0;
  }
  if (idx >= (size_t)Kp * M) return;
  const int j = (int)(idx / Kp), k = (int)(idx % Kp);
  double v = PHI[idx];
  if (!isfinite(v)) atomicCAS(&status[0], 0, 100);
  if (k >= K) return;
  double dlt = v - PHI0[idx];
  if (do_clamp) {
    if (dlt == 0.0) {
      RngStream rng(seed, STREAM_CLAMP, (uint32_t)idx, *sweep);
      dlt = (rng.uniform() < 0.5) ? tol : -tol;   // R::rbinom(1, 0.5) == 0 -> +tol
      PHI[idx] = PHI0[idx] + dlt;
    } else if (dlt < tol && dlt > 0.0) {
      dlt = tol;
      PHI[idx] = PHI0[idx] + tol;
    } else if (dlt > -tol && dlt < 0.0) {
      dlt = -tol;
      PHI[idx] = PHI0[idx] - tol;
    }
  }
  coef[(size_t)j * K + k] = dlt;
}

__global__ void gibbs_scatter_V_kernel(double* Vprior, const double* V_i, int K, int Kp, int M) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)K * M) return;
  const int j = (int)(idx / K), k = (int)(idx % K);
  Vprior[(size_t)j * Kp + k] = V_i[idx];
}

// resid = Y - X PHI  (T x M, bvar_cpp.cpp:611).  One CTA per 8 x 8 output tile, its four warps split the Kp
// contraction (fixed-order reduction through shared memory): DMMA m8n8k4 with both operands read straight from L2
// (X 1.6 MB and PHI 0.3 MB stay resident); A[m][k] = X[t0+m, k], B[k][n] = PHI[k, j0+n].  The kernel is pure L2
// latency (every trip of the k loop is one round trip), so the contraction is cut four ways: 24 -> 9 us.
constexpr int RESID_SPLIT = 4;
__global__ void __launch_bounds__(32 * RESID_SPLIT) gibbs_resid_kernel(double* __restrict__ resid, const double* __restrict__ Y,
                                                                       const double* __restrict__ X, const double* __restrict__ PHI,
                                                                       int T, int Kp, int M) {
  __shared__ double part[RESID_SPLIT][64];
  const int lane = threadIdx.x & 31, wk = threadIdx.x >> 5;
  const int ntj = (M + 7) / 8;
  const int tile = blockIdx.x;
  const int t0 = (tile / ntj) * 8, j0 = (tile % ntj) * 8;
  const int fr = lane >> 2, fk = lane & 3;
  const int tr = min(t0 + fr, T - 1), jc = min(j0 + fr, M - 1);
  const double* xa = X + tr;
  const double* pb = PHI + (size_t)jc * Kp;
  // this warp's slice of k: multiples of 4
  const int steps = (Kp + 3) / 4, per = (steps + RESID_SPLIT - 1) / RESID_SPLIT;
  const int kbeg = min(Kp, 4 * per * wk), kend = min(Kp, 4 * per * (wk + 1));
  double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
  int k = kbeg;
  for (; k + 32 <= kend; k += 32) {   // 8 k4-steps per trip: 16 independent L2 loads in flight per lane
    double av[8], bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { av[q] = xa[(size_t)(k + 4 * q + fk) * T]; bv[q] = pb[k + 4 * q + fk]; }
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      dmma884(c0, c1, av[q], bv[q]);
      dmma884(d0, d1, av[q + 1], bv[q + 1]);
    }
  }
  {   // remainder of the slice (< 8 k4-steps), loads issued together
    double av[8], bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = k + 4 * q + fk;
      av[q] = kk < kend ? xa[(size_t)kk * T] : 0.0;
      bv[q] = kk < kend ? pb[kk] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      dmma884(c0, c1, av[q], bv[q]);
      dmma884(d0, d1, av[q + 1], bv[q + 1]);
    }
  }
  c0 += d0; c1 += d1;
  part[wk][fr * 8 + 2 * fk] = c0;
  part[wk][fr * 8 + 2 * fk + 1] = c1;
  __syncthreads();
  if (wk == 0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = fr * 8 + 2 * fk + e;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < RESID_SPLIT; ++w) v += part[w][idx];
      const int t = t0 + fr, j = j0 + 2 * fk + e;
      if (t < T && j < M) resid[(size_t)j * T + t] = Y[(size_t)j * T + t] - v;
    }
  }
}

// Right-hand sides b_j = X'(w_j . yhat_j) + PHI0_j / v_j of the local equations, scattered into the packed systems
// (what K1 does with its X tiles).  A separate small contraction so that K1's Z tiles - 99.5 % of its work, which only
// depend on W - can be computed before yhat of the next sweep exists.  Same layout as the residual kernel: one CTA per
// 8 x 8 output tile, four warps split the contraction over t (both operands are t-contiguous and L2-resident).
constexpr int RHS_SPLIT = 4;
__global__ void __launch_bounds__(32 * RHS_SPLIT) gibbs_k1_rhs_kernel(const double* __restrict__ AX, const int2* __restrict__ lutX,
                                                                     const double* __restrict__ WY, const double* __restrict__ V,
                                                                     const double* __restrict__ PHI0, double* __restrict__ omega,
                                                                     int nrows, int Tp, int Mloc, int n, long ldo) {
  __shared__ double part[RHS_SPLIT][64];
  const int lane = threadIdx.x & 31, wk = threadIdx.x >> 5;
  const int ntj = (Mloc + 7) / 8;
  const int tile = blockIdx.x;
  const int r0 = (tile / ntj) * 8, j0 = (tile % ntj) * 8;
  const int fr = lane >> 2, fk = lane & 3;
  const double* xa = AX + (size_t)min(r0 + fr, nrows - 1) * Tp;
  const double* pb = WY + (size_t)min(j0 + fr, Mloc - 1) * Tp;
  const int steps = (Tp + 3) / 4, per = (steps + RHS_SPLIT - 1) / RHS_SPLIT;
  const int kbeg = min(Tp, 4 * per * wk), kend = min(Tp, 4 * per * (wk + 1));
  double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
  int k = kbeg;
  for (; k + 32 <= kend; k += 32) {
    double av[8], bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { av[q] = xa[k + 4 * q + fk]; bv[q] = pb[k + 4 * q + fk]; }
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      dmma884(c0, c1, av[q], bv[q]);
      dmma884(d0, d1, av[q + 1], bv[q + 1]);
    }
  }
  {
    double av[8], bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = k + 4 * q + fk;
      av[q] = kk < kend ? xa[kk] : 0.0;
      bv[q] = kk < kend ? pb[kk] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      dmma884(c0, c1, av[q], bv[q]);
      dmma884(d0, d1, av[q + 1], bv[q + 1]);
    }
  }
  c0 += d0; c1 += d1;
  part[wk][fr * 8 + 2 * fk] = c0;
  part[wk][fr * 8 + 2 * fk + 1] = c1;
  __syncthreads();
  if (wk == 0) {
    const int row = r0 + fr;
    const int2 l = row < nrows ? lutX[row] : make_int2(-1, -1);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = fr * 8 + 2 * fk + e;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < RHS_SPLIT; ++w) v += part[w][idx];
      const int j = j0 + 2 * fk + e;
      if (l.x >= 0 && j < Mloc) {
        if (l.y >= 0 && PHI0) v += PHI0[(size_t)j * n + l.y] / V[(size_t)j * n + l.y];
        omega[(size_t)j * ldo + l.x] = v;
      }
    }
  }
}

cudaError_t launch_resid(double* resid, const double* Y, const double* X, const double* PHI, int T, int Kp, int M,
                         cudaStream_t s) {
  const int tiles = ((T + 7) / 8) * ((M + 7) / 8);
  gibbs_resid_kernel<<<tiles, 32 * RESID_SPLIT, 0, s>>>(resid, Y, X, PHI, T, Kp, M);
  return cudaGetLastError();
}

// Copy the current state into slot `slot` of the device ring if this sweep is stored.
__global__ void gibbs_store_kernel(const GibbsStore st, uint32_t* sweep, int* nstored) {
  // nstored[0] = highest slot written + 1; nstored[1] = blocks of this launch that are done (the last one advances
  // the sweep counter - every block has read it by then - so there is no separate one-thread "advance" launch)
  const int rep = (int)*sweep;
  const bool stored = !(rep < st.burnin || ((rep + 1 - st.burnin) % st.thin) != 0);
  const int slot = stored ? (((rep + 1 - st.burnin) / st.thin) - 1) - *st.first_slot : -1;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (slot >= 0 && slot < st.ring_slots) {
    double* dst = st.ring + (size_t)slot * st.slot_doubles;
    for (int s = 0; s < st.nseg; ++s) {
      const GibbsSeg& sg = st.seg[s];
      const unsigned n = (unsigned)sg.rows * (unsigned)sg.cols, rows = (unsigned)sg.rows;
      for (unsigned i = (unsigned)tid; i < n; i += (unsigned)nth) {
        const unsigned c = i / rows, rr = i - c * rows;
        dst[sg.dst_off + i] = sg.src[(size_t)c * sg.src_ld + sg.src_row0 + rr];
      }
    }
    if (tid == 0) atomicMax(nstored, slot + 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(nstored + 1, 1) == (int)gridDim.x - 1) {
      nstored[1] = 0;
      *sweep = (uint32_t)rep + 1u;
    }
  }
}

// Cholesky structure: push |U_ij| (i < j) away from zero for DL / GT priors
// (bvar_cpp.cpp:667-685) and gather u = U(trimatu_ind) (:688) - column-major upper
// triangle, element (i, j) at j(j-1)/2 + i.
__global__ void chol_clamp_u_kernel(double* U, double* uvec, int M, int do_clamp, double tol, uint64_t seed,
                                    const uint32_t* sweep) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * M) return;
  const int j = (int)(idx / M), i = (int)(idx % M);
  if (i >= j) return;
  double v = U[idx];
  if (do_clamp) {
    if (v == 0.0) {
      RngStream rng(seed, STREAM_CLAMP_U, (uint32_t)idx, *sweep);
      v = (rng.uniform() < 0.5) ? tol : -tol;
    } else if (v < tol && v > 0.0) v = tol;
    else if (v > -tol && v < 0.0) v = -tol;
    U[idx] = v;
  }
  uvec[(size_t)j * (j - 1) / 2 + i] = v;
}

// str_resid = resid U (bvar_cpp.cpp:725)
__global__ void chol_strres_kernel(double* out, const double* resid, const double* U, int T, int M) {
  const int i = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double q = 0.0;
    for (int m = 0; m <= i; ++m) q = fma(resid[(size_t)m * T + t], U[(size_t)i * M + m], q);
    out[(size_t)i * T + t] = q;
  }
}

__global__ void chol_dsq_kernel(double* dsq, const double* logvar, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dsq[i] = exp(0.5 * logvar[i]);
}

__global__ void gibbs_flat_normals_kernel(double* out, size_t n, uint64_t seed, uint32_t stream, const uint32_t* sweep) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RngStream rng(seed, stream, (uint32_t)i, *sweep);
  out[i] = rng.normal();
}

}  // namespace bvb

// ------------------------------------------------------------------ host side
namespace {

#define GB_TRY(expr) BVB_CUDA_TRY(expr)

int fail_arg(bvb_handle* h, const char* msg) {
  bvb_set_error(h, BVB_ERR_ARG, "bvb_gibbs_init: %s", msg);
  return BVB_ERR_ARG;
}

template <typename T>
int upload(bvb_handle* h, GBuf<T>& b, const T* src, size_t n) {
  if (n == 0) return BVB_OK;
  GB_TRY(b.ensure(n));
  if (src) {
    // through the chain's pinned staging arena when it fits: a copy from pageable memory is staged by the runtime
    // synchronously, 10-20 us apiece for the ~45 small arrays of a chain set-up plus 0.3 ms for X; from pinned memory the
    // calls return at once and the transfers overlap the rest of the set-up.  (The arena is rewound at the next
    // bvb_gibbs_init, after a stream synchronisation.)
    Gibbs* G = static_cast<Gibbs*>(h->gibbs);
    const size_t bytes = n * sizeof(T), al = (bytes + 255) & ~(size_t)255;
    if (G && G->stage && G->stage_off + al <= G->stage_cap) {
      memcpy(G->stage + G->stage_off, src, bytes);
      GB_TRY(cudaMemcpyAsync(b.p, G->stage + G->stage_off, bytes, cudaMemcpyHostToDevice, h->stream));
      G->stage_off += al;
    } else {
      GB_TRY(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    }
  }
  return BVB_OK;
}

// Fill a HyperDev (+ its buffers) from the ABI description of one prior block.
int setup_hyper(bvb_handle* h, GibbsHyper& H, int prior, int n, const std::vector<int>& grp, int G,
                const std::vector<double>& gpar, const double* V_i_fixed, double tol, const double* a_vec,
                const double* a_weight, const double* norm_consts, const double* c_vec, int grid_len, int hyper,
                int dl_plus, int c_rel_a, int kernel_exp, double vs, const double* tau0, const double* tau1,
                double s_a, double s_b, const double* p_init, const std::vector<double>& V_prep,
                const std::vector<double>& loc1_init, const std::vector<double>& loc2_init,
                const std::vector<double>& V_init, uint32_t stream_base) {
  HyperDev& d = H.d;
  d.prior = prior; d.n = n; d.G = G; d.nblocks = (n + HYP_THREADS - 1) / HYP_THREADS;
  d.tol = tol; d.vs = vs; d.hyper = hyper; d.dl_plus = dl_plus; d.c_rel_a = c_rel_a; d.kernel_exp = kernel_exp;
  d.s_a = s_a; d.s_b = s_b; d.grid_len = grid_len; d.stream_base = stream_base;
  if (n == 0) return BVB_OK;
  int rc;
  if ((rc = upload(h, H.grp, grp.data(), n))) return rc;
  {
    // member lists of the groups (CSR, ascending coefficient index): one CTA per group reduces them in a fixed order
    const int Gs = std::max(G, 1);
    std::vector<int> gstart((size_t)Gs + 1, 0), gmem;
    for (int i = 0; i < n; ++i) if (grp[i] >= 0 && grp[i] < Gs) gstart[grp[i] + 1] += 1;
    for (int g = 0; g < Gs; ++g) gstart[g + 1] += gstart[g];
    gmem.assign((size_t)std::max(gstart[Gs], 1), 0);
    std::vector<int> fill(gstart.begin(), gstart.end() - 1);
    for (int i = 0; i < n; ++i) if (grp[i] >= 0 && grp[i] < Gs) gmem[fill[grp[i]]++] = i;
    if ((rc = upload(h, H.gstart, gstart.data(), gstart.size()))) return rc;
    if ((rc = upload(h, H.gmem, gmem.data(), gmem.size()))) return rc;
  }
  if ((rc = upload(h, H.gpar, gpar.data(), (size_t)std::max(G, 1) * HYP_GP))) return rc;
  GB_TRY(H.coef.ensure(n));
  if ((rc = upload(h, H.V_i, V_init.data(), n))) return rc;
  if ((rc = upload(h, H.loc1, loc1_init.data(), n))) return rc;
  if ((rc = upload(h, H.loc2, loc2_init.data(), n))) return rc;
  GB_TRY(H.red.ensure((size_t)n * HYP_NRED));
  GB_TRY(H.gsum.ensure((size_t)std::max(G, 1) * HYP_NRED));
  // few large groups (one global group of 40 000 coefficients at the headline shape): several CTAs per group
  const int nsplit = std::max(1, std::min(32, (int)((size_t)n / ((size_t)std::max(G, 1) * 2048))));
  GB_TRY(H.gpart.ensure((size_t)std::max(G, 1) * nsplit));
  GB_TRY(H.gcnt.ensure((size_t)std::max(G, 1), true));
  if (grid_len > 0) {
    if ((rc = upload(h, H.a_vec, a_vec, grid_len))) return rc;
    if ((rc = upload(h, H.a_weight, a_weight, grid_len))) return rc;
    if ((rc = upload(h, H.norm_consts, norm_consts, grid_len))) return rc;
    if ((rc = upload(h, H.c_vec, c_vec, grid_len))) return rc;
  }
  if (prior == BVB_PRIOR_SSVS) {
    if ((rc = upload(h, H.tau0, tau0, n))) return rc;
    if ((rc = upload(h, H.tau1, tau1, n))) return rc;
  }
  if (prior == BVB_PRIOR_HMP) if ((rc = upload(h, H.V_prep, V_prep.data(), n))) return rc;
  (void)V_i_fixed; (void)p_init;
  d.grp = H.grp.p; d.coef = H.coef.p; d.V_i = H.V_i.p; d.loc1 = H.loc1.p; d.loc2 = H.loc2.p; d.gpar = H.gpar.p;
  d.gstart = H.gstart.p; d.gmem = H.gmem.p; d.red = H.red.p; d.gsum = H.gsum.p; d.gpart = H.gpart.p; d.gcnt = H.gcnt.p; d.nsplit = nsplit; d.a_vec = H.a_vec.p; d.a_weight = H.a_weight.p;
  d.norm_consts = H.norm_consts.p; d.c_vec = H.c_vec.p; d.tau0 = H.tau0.p; d.tau1 = H.tau1.p; d.V_prep = H.V_prep.p;
  return BVB_OK;
}

void add_seg(GibbsStore& st, const double* src, int rows, int cols, int src_ld, int src_row0, size_t& off) {
  GibbsSeg& s = st.seg[st.nseg++];
  s.src = src; s.rows = rows; s.cols = cols; s.src_ld = src_ld; s.src_row0 = src_row0; s.dst_off = off;
  off += (size_t)rows * cols;
}

}  // namespace

// Z tiles of the pair-GEMM (Omega_j = Z W_j + diag(1/v_j)) of the local equations from the current W and V_prior.
static cudaError_t enqueue_k1z(Gibbs& G, cudaStream_t s) {
  const int Kp = G.Kp;
  K1Params k1{};
  k1.A = G.A.p; k1.W = G.W.p + (size_t)G.j0 * G.geom.Tp; k1.WY = G.WY.p + (size_t)G.j0 * G.geom.Tp; k1.lut = G.lut.p;
  k1.V = G.Vprior.p + (size_t)G.j0 * Kp; k1.PHI0 = G.PHI0.p + (size_t)G.j0 * Kp; k1.omega = G.omega.p;
  k1.Tp = G.geom.Tp; k1.nZtiles = G.geom.nZtiles; k1.ntiles = G.geom.nZtiles;     // no X tiles: gibbs_k1_rhs_kernel
  k1.M = G.nloc; k1.n = Kp; k1.ldo = G.geom.ldo;
  k1.tstamp = G.ktime.p;
  return launch_k1(k1, s);
}

// Normals of the local equations, WY = w * yhat (and W unless the Z tiles already own it) for the next PHI draw.
static cudaError_t enqueue_prep(bvb_handle* h, Gibbs& G, cudaStream_t s, int write_w) {
  const size_t nz = (size_t)G.Kp * G.nloc;
  gibbs_prep_normals_kernel<<<(unsigned)(G.M + (nz + 255) / 256), 256, 0, s>>>(G.W.p, G.WY.p, G.Y.p, G.logvar.p, G.facload.p,
                                                                              G.fac.p, G.T, G.geom.Tp, G.M, G.r, G.Zn.p, G.Kp,
                                                                              G.nloc, G.j0, G.seed, G.sweep.p, write_w);
  return cudaGetLastError();
}

// Right-hand sides of the local systems from WY (K1's X tiles as a separate small contraction).
static cudaError_t enqueue_k1rhs(Gibbs& G, cudaStream_t s) {
  const int Kp = G.Kp, nrows = G.geom.nXtiles * K1_BM;
  const int tiles = ((nrows + 7) / 8) * ((G.nloc + 7) / 8);
  gibbs_k1_rhs_kernel<<<tiles, 32 * RHS_SPLIT, 0, s>>>(G.A.p + (size_t)G.geom.nZtiles * K1_BM * G.geom.Tp,
                                                      G.lut.p + (size_t)G.geom.nZtiles * K1_BM,
                                                      G.WY.p + (size_t)G.j0 * G.geom.Tp, G.Vprior.p + (size_t)G.j0 * Kp,
                                                      G.PHI0.p + (size_t)G.j0 * Kp, G.omega.p, nrows, G.geom.Tp, G.nloc, Kp,
                                                      G.geom.ldo);
  return cudaGetLastError();
}

// ---- peer exchange of the freshly drawn PHI columns -------------------------------------------------------------------
// Every rank owns a contiguous block of columns.  push: CTA b copies this rank's block straight into peer p's exchange
// buffer (remote stores over NVLink), every thread fences at system scope, then one thread stamps p's flag for this rank
// with (epoch << 32 | sweep + 1).  pull: CTA b spins on the flag of source p (acquire at system scope, bounded), then
// copies p's block from the local exchange buffer into PHI.  The buffer is double-buffered by sweep parity: a peer can
// be at most one sweep ahead (its next K2 needs THIS rank's push), so it never overwrites a block still being read; and
// by the parity of the chain count, so the first sweep of a recycled chain cannot touch the last sweep of the previous one.
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr int XCH_SLICES = 8;    // CTAs per peer: one CTA pushing 160 KB of 8-byte remote stores took 30 us
__global__ void __launch_bounds__(512) gibbs_push_kernel(const double* __restrict__ mine, size_t blk, double* const* __restrict__ peer_xbuf,
                                                         unsigned long long* const* __restrict__ peer_xflag, int* __restrict__ cnt, int rank,
                                                         int world, unsigned long long epoch_base, const uint32_t* __restrict__ sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int pi = (int)blockIdx.x / XCH_SLICES, sl = (int)blockIdx.x % XCH_SLICES;
  const int p = pi < rank ? pi : pi + 1;
  double* dst = peer_xbuf[p] + (((epoch_base >> 32) & 1ull) * 2 + (sweep & 1u)) * (size_t)world * blk + (size_t)rank * blk;
  const size_t per = (blk + XCH_SLICES - 1) / XCH_SLICES, i0 = (size_t)sl * per, i1 = i0 + per < blk ? i0 + per : blk;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = mine[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(cnt + pi, 1);
    __threadfence();
    if (old == XCH_SLICES - 1) {                      // last slice for this peer: everything is out, stamp its flag
      cnt[pi] = 0;
      __threadfence_system();
      st_release_sys_u64(peer_xflag[p] + rank, epoch_base + sweep + 1ull);
    }
  }
}
__global__ void __launch_bounds__(512) gibbs_pull_kernel(double* __restrict__ PHI, const double* __restrict__ xbuf,
                                                         const unsigned long long* __restrict__ xflag, size_t blk, int rank, int world,
                                                         unsigned long long epoch_base, const uint32_t* __restrict__ sweep_p, int* status) {
  const uint32_t sweep = *sweep_p;
  const int pi = (int)blockIdx.x / XCH_SLICES, sl = (int)blockIdx.x % XCH_SLICES;
  const int p = pi < rank ? pi : pi + 1;
  if (threadIdx.x == 0) {
    const unsigned long long want = epoch_base + sweep + 1ull;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(xflag + p) < want) {
      if (clock64() - t0 > (1ll << 36)) { atomicExch(status + 3, 1 + p); break; }   // ~35 s: a peer is gone; reported, not hung
    }
  }
  __syncthreads();
  const double* src = xbuf + (((epoch_base >> 32) & 1ull) * 2 + (sweep & 1u)) * (size_t)world * blk + (size_t)p * blk;
  double* dst = PHI + (size_t)p * blk;
  const size_t per = (blk + XCH_SLICES - 1) / XCH_SLICES, i0 = (size_t)sl * per, i1 = i0 + per < blk ? i0 + per : blk;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = __ldcg(src + i);   // (written by a peer: never through L1)
}

// Does the sweep graph of this chain run in its rotated form (the NEXT sweep's K1 Z tiles under this sweep's tail)?
static bool gibbs_rotates(const Gibbs& G) {
  return getenv("BVB_NO_ROTATE") == nullptr && G.K > 0 && G.sigma_type != BVB_SIGMA_CHOLESKY && !G.huge && G.side != nullptr &&
         G.ev_series != nullptr && G.ev_k1z != nullptr;
}

// One sweep's worth of launches on h->stream (captured into a graph by the caller).
static int enqueue_sweep(bvb_handle* h, Gibbs& G) {
  cudaStream_t s = h->stream;
  const int M = G.M, T = G.T, K = G.K, Kp = G.Kp;
  const uint32_t* sw = G.sweep.p;
  cudaEvent_t* ev = G.ev;
  const bool timing = G.timing_pass;
  auto mark = [&](int i) { if (timing) cudaEventRecord(ev[i], s); };
  mark(0);
  if (K > 0 && G.sigma_type == BVB_SIGMA_CHOLESKY) {
    // ---- 1) PHI | .  (cholesky structure: corrected triangular algorithm, bvar_cpp.cpp:510)
    const size_t nz = (size_t)Kp * M;
    gibbs_normals_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, s>>>(G.Zn.p, Kp, M, Kp, 0, G.seed, sw);
    mark(1);
    GB_TRY(launch_chol_phi(G.cphi, s));
    mark(2); mark(3); mark(4);
    const int do_clamp = (G.hphi.d.prior == BVB_PRIOR_DL || G.hphi.d.prior == BVB_PRIOR_GT) ? 1 : 0;
    gibbs_clamp_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, s>>>(G.PHI.p, G.PHI0.p, G.hphi.d.coef, K, Kp, M, do_clamp,
                                                                   G.PHI_tol, G.seed, sw, G.status.p, nullptr);
  } else if (K > 0 && G.huge) {
    // ---- 1) PHI | .  (factor structure, expert_huge: T x T systems; every rank draws all equations)
    const size_t nz = (size_t)Kp * M, nd = (size_t)T * M;
    gibbs_normals_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, s>>>(G.Zn.p, Kp, M, Kp, 0, G.seed, sw);
    gibbs_flat_normals_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, s>>>(G.hdelta.p, nd, G.seed, STREAM_HUGE_DELTA, sw);
    mark(1);
    GB_TRY(launch_phi_huge(G.chuge, s));
    mark(2); mark(3); mark(4);
    const int do_clamp = (G.hphi.d.prior == BVB_PRIOR_DL || G.hphi.d.prior == BVB_PRIOR_GT) ? 1 : 0;
    gibbs_clamp_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, s>>>(G.PHI.p, G.PHI0.p, G.hphi.d.coef, K, Kp, M, do_clamp,
                                                                   G.PHI_tol, G.seed, sw, G.status.p, nullptr);
  } else if (K > 0) {  // bvar_cpp.cpp:507: PHI is only drawn when there are lagged regressors
    // ---- 1) PHI | .  (factor structure: equations independent)
    // K1 is split.  The Z tiles (X'diag(w)X, 99.5 % of its work) only need W = exp(-h), which is final once the SV series
    // kernel of the PREVIOUS sweep has run, so that sweep computed them on the side stream under its own tail, and its
    // main stream finished with this sweep's normals, WY and right-hand sides (enqueue_assemble below).  Only the first
    // sweep of a chain (and the timing pass, so that phases "prep" and "k1" show their cost) assemble here.
    if (!G.k1z_valid || timing) {
      GB_TRY(enqueue_prep(h, G, s, 1));
      mark(1);
      GB_TRY(enqueue_k1z(G, s));
      GB_TRY(enqueue_k1rhs(G, s));
    } else {
      mark(1);
    }
    mark(2);
    K2Params k2{};
    k2.omega = G.omega.p; k2.colbase = G.colbase.p; k2.z = G.Zn.p; k2.phi = G.PHI.p + (size_t)G.j0 * Kp;
    k2.status = G.eqstatus.p + G.j0; k2.n = Kp; k2.M = G.nloc; k2.ldo = G.geom.ldo; k2.ldz = Kp; k2.ldphi = Kp;
    k2.tstamp = G.ktime.p + 2;
    bool xch_fused = false;
    if (G.world > 1 && G.p2p && G.nloc > 0) {
      k2.xch.peer_xbuf = G.peer_xbuf.p; k2.xch.peer_xflag = G.peer_xflag.p; k2.xch.cnt = G.xcnt.p + G.world;   // (own counter word)
      k2.xch.rank = G.rank; k2.xch.world = G.world; k2.xch.blk = (size_t)G.nloc_max * Kp;
      k2.xch.epoch_base = G.epoch_base; k2.xch.sweep_p = sw;
      k2.xch_fused = &xch_fused;
    }
    GB_TRY(launch_k2(k2, s));
    mark(3);
    if (G.world > 1) {
      // ranks own contiguous blocks of nloc_max columns (PHI is padded to world*nloc_max
      // columns), so the gather is in place: sendbuff == recvbuff + rank*count
      const size_t blk = (size_t)G.nloc_max * Kp;
      if (G.p2p) {
        const unsigned xg = (unsigned)((G.world - 1) * XCH_SLICES);
        // (the tiled K2's back-solve pushes the columns itself; the stand-alone push serves the other K2 forms)
        if (!xch_fused) gibbs_push_kernel<<<xg, 512, 0, s>>>(G.PHI.p + (size_t)G.rank * blk, blk, G.peer_xbuf.p, G.peer_xflag.p, G.xcnt.p, G.rank, G.world,
                                             G.epoch_base, sw);
        gibbs_pull_kernel<<<xg, 512, 0, s>>>(G.PHI.p, G.xbuf.p, G.xflag.p, blk, G.rank, G.world, G.epoch_base, sw, G.status.p);
      } else {
        ncclResult_t nr = nccl_api().AllGather(G.PHI.p + (size_t)G.rank * blk, G.PHI.p, blk, ncclDouble, (ncclComm_t)G.comm, s);
        if (nr != ncclSuccess) { bvb_set_error(h, BVB_ERR_COMM, "ncclAllGather: %s", nccl_api().GetErrorString(nr)); return BVB_ERR_COMM; }
      }
    }
    mark(4);
    const size_t nphi = (size_t)Kp * M;
    const int do_clamp = (G.hphi.d.prior == BVB_PRIOR_DL || G.hphi.d.prior == BVB_PRIOR_GT) ? 1 : 0;
    gibbs_clamp_kernel<<<(unsigned)((nphi + 255) / 256), 256, 0, s>>>(G.PHI.p, G.PHI0.p, G.hphi.d.coef, K, Kp, M, do_clamp,
                                                                     G.PHI_tol, G.seed, sw, G.status.p, G.ktime.p);
  } else {
    mark(1); mark(2); mark(3); mark(4);
  }
  // ---- 2) hyper-parameters of the prior on PHI.  They only read PHI and write V_prior / their own state, which
  // nothing below touches, so outside the timing pass they run on a side stream next to resid + Sigma_t (a fork /
  // join in the captured graph); the join is before the store.
#ifdef BVB_DEBUG_SWITCHES
  static const bool dbg_skip_hyper = getenv("BVB_DEBUG_SKIP_HYPER") != nullptr;   // timing experiments only (wrong draws)
#else
  constexpr bool dbg_skip_hyper = false;   // (switches that change the draws are compiled out of release builds)
#endif
  const bool has_hyper = !dbg_skip_hyper && G.hphi.d.n > 0 && G.hphi.d.prior != BVB_PRIOR_NORMAL && G.hphi.d.prior != BVB_PRIOR_NOTHING;
  static const bool dbg_fork_timing = getenv("BVB_DEBUG_FORK_TIMING") != nullptr;
  const bool forked = has_hyper && (!timing || dbg_fork_timing) && G.side != nullptr;
  const bool rotate = gibbs_rotates(G);                          // (read per enqueue: tests compare both forms)
  const bool no_ng_side = getenv("BVB_NG_INLINE") != nullptr;
  const bool ng_side = forked && !no_ng_side && G.sigma_type != BVB_SIGMA_CHOLESKY && fsv_has_ng(G.fsv) && G.ev_ng != nullptr;
  {
    cudaStream_t hs = forked ? G.side : s;
    if (forked) {
      GB_TRY(cudaEventRecord(G.ev_fork, s));
      GB_TRY(cudaStreamWaitEvent(hs, G.ev_fork, 0));
      // the NG hyper-parameters of the factor loadings only read last sweep's loadings: first on the side stream, the
      // Sigma_t branch waits for them in front of its loadings kernel (19 us off the critical path)
      if (ng_side) {
        GB_TRY(launch_fsv_ng(G.fsv, G.seed, sw, hs));
        GB_TRY(cudaEventRecord(G.ev_ng, hs));
      }
    }
    if (has_hyper) {
      GB_TRY(launch_hyper_update(G.hphi.d, G.seed, sw, G.burnin, hs));
      const size_t nk = (size_t)K * M;
      gibbs_scatter_V_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, hs>>>(G.Vprior.p, G.hphi.d.V_i, K, Kp, M);
    }
    if (forked) GB_TRY(cudaEventRecord(G.ev_join, hs));
    if (timing) cudaEventRecord(ev[5], hs);
  }
  // ---- resid = Y - X PHI
  if (Kp > 0) {
    const int tiles = ((T + 7) / 8) * ((M + 7) / 8);
    gibbs_resid_kernel<<<tiles, 32 * RESID_SPLIT, 0, s>>>(G.resid.p, G.Y.p, G.X.p, G.PHI.p, T, Kp, M);
  }
  mark(6);
  // ---- 3) Sigma_t
  if (G.sigma_type == BVB_SIGMA_CHOLESKY) {
    const int M2 = M * M;
    if (G.nU > 0) {
      gibbs_flat_normals_kernel<<<(G.nU + 255) / 256, 256, 0, s>>>(G.zflat.p, (size_t)G.nU, G.seed, STREAM_U_Z, sw);
      GB_TRY(launch_sample_u(G.cu, s));                                                       // bvar_cpp.cpp:659
      const int clampU = (G.hU.d.prior == BVB_PRIOR_DL || G.hU.d.prior == BVB_PRIOR_GT) ? 1 : 0;
      chol_clamp_u_kernel<<<(M2 + 255) / 256, 256, 0, s>>>(G.U.p, G.hU.d.coef, M, clampU, G.L_tol, G.seed, sw);   // :667-688
      GB_TRY(launch_hyper_update(G.hU.d, G.seed, sw, G.burnin, s));                          // :689-721
    }
    chol_strres_kernel<<<M, 256, 0, s>>>(G.strres.p, G.resid.p, G.U.p, T, M);                 // :725
    GB_TRY(launch_fsv_update(G.fsv, G.seed, sw, s));                                         // :727-746
    const size_t tm = (size_t)T * M;
    chol_dsq_kernel<<<(unsigned)((tm + 255) / 256), 256, 0, s>>>(G.dsq.p, G.logvar.p, tm);
  } else {
    {
      // rotated graph: the series kernel writes next sweep's W itself (every idiosyncratic series passes through it)
      FsvDev fs = G.fsv;
      if (rotate) { fs.Wk1 = G.W.p; fs.Tp = G.geom.Tp; }
      GB_TRY(launch_fsv_series(fs, G.seed, sw, s));
    }
    if (rotate && !timing) {
      // the idiosyncratic log-variances are final: the next sweep's Z tiles run on the (low-priority) side stream,
      // behind the hyper-parameters (V_prior), under loadings -> interweaving -> factors -> store
      GB_TRY(cudaEventRecord(G.ev_series, s));
      GB_TRY(cudaStreamWaitEvent(G.side, G.ev_series, 0));
      GB_TRY(enqueue_k1z(G, G.side));
      GB_TRY(cudaEventRecord(G.ev_k1z, G.side));
    }
    GB_TRY(launch_fsv_rest(G.fsv, G.seed, sw, s, ng_side ? G.ev_ng : nullptr));
  }
  mark(7);
  if (forked) GB_TRY(cudaStreamWaitEvent(s, G.ev_join, 0));
  // ---- store + advance
  // (two CTAs per SM: with 64 CTAs every thread copied ten elements one L2 round trip at a time, 14 us for 1.3 MB)
  gibbs_store_kernel<<<2 * (h->num_sms > 0 ? h->num_sms : 148), 256, 0, s>>>(G.store, G.sweep.p, G.nstored.p);   // also advances the sweep counter
  mark(8);
  if (rotate) {
    // the store advanced the sweep counter: normals, WY and right-hand sides of the NEXT sweep (disjoint from the entries
    // of `omega` the Z tiles are writing), then the join with the Z tiles
    if (timing) GB_TRY(enqueue_k1z(G, s));   // serial form of the same tail (the timing pass measures phase by phase)
    GB_TRY(enqueue_prep(h, G, s, 0));
    GB_TRY(enqueue_k1rhs(G, s));
    if (!timing) GB_TRY(cudaStreamWaitEvent(s, G.ev_k1z, 0));
  }
  G.k1z_valid = rotate;
  GB_TRY(cudaGetLastError());
  return BVB_OK;
}

static void gibbs_free(bvb_handle* h) {
  if (!h->gibbs) return;
  Gibbs* G = static_cast<Gibbs*>(h->gibbs);
  if (G->fsv.prof) fsv_print_profile();
  if (G->graph_exec) cudaGraphExecDestroy(G->graph_exec);
  if (G->stage) cudaFreeHost(G->stage);
  for (void* q : G->ipc_opened) cudaIpcCloseMemHandle(q);
  for (auto& e : G->ev) if (e) cudaEventDestroy(e);
  if (G->ev_fork) cudaEventDestroy(G->ev_fork);
  if (G->ev_join) cudaEventDestroy(G->ev_join);
  if (G->ev_ng) cudaEventDestroy(G->ev_ng);
  if (G->ev_series) cudaEventDestroy(G->ev_series);
  if (G->ev_k1z) cudaEventDestroy(G->ev_k1z);
  if (G->side) cudaStreamDestroy(G->side);
  if (G->copy) cudaStreamDestroy(G->copy);
  if (G->first_slot_host) cudaFreeHost(G->first_slot_host);
  for (int b = 0; b < 2; ++b) if (G->ev_sub[b]) cudaEventDestroy(G->ev_sub[b]);
  if (G->ev_chunk[0]) cudaEventDestroy(G->ev_chunk[0]);
  if (G->ev_chunk[1]) cudaEventDestroy(G->ev_chunk[1]);
  for (int b = 0; b < 2; ++b) if (G->ev_pin[b]) cudaEventDestroy(G->ev_pin[b]);
  delete G;
  h->gibbs = nullptr;
}

void bvb_gibbs_release(bvb_handle* h) { gibbs_free(h); }

extern "C" {

// Multi-process form: the peers' exchange buffers are opened through CUDA IPC; the 128 bytes of handles per rank travel
// through one NCCL allgather at set-up (NCCL stays the bootstrap, it is no longer in the sweep).
static int p2p_bootstrap_ipc(bvb_handle* h, Gibbs& G) {
  for (void* q : G.ipc_opened) cudaIpcCloseMemHandle(q);
  G.ipc_opened.clear();
  struct Pack { cudaIpcMemHandle_t x, f; };
  static_assert(sizeof(Pack) == 128, "two IPC handles");
  cudaStream_t s = h->stream;
  GBuf<double> send, recv;
  GB_TRY(send.ensure(16)); GB_TRY(recv.ensure((size_t)16 * G.world));
  Pack mine;
  GB_TRY(cudaIpcGetMemHandle(&mine.x, G.xbuf.p));
  GB_TRY(cudaIpcGetMemHandle(&mine.f, G.xflag.p));
  GB_TRY(cudaMemcpyAsync(send.p, &mine, sizeof(Pack), cudaMemcpyHostToDevice, s));
  ncclResult_t nr = nccl_api().AllGather(send.p, recv.p, 16, ncclDouble, (ncclComm_t)G.comm, s);
  if (nr != ncclSuccess) { bvb_set_error(h, BVB_ERR_COMM, "ncclAllGather (set-up): %s", nccl_api().GetErrorString(nr)); return BVB_ERR_COMM; }
  GB_TRY(cudaStreamSynchronize(s));
  std::vector<Pack> all((size_t)G.world);
  GB_TRY(cudaMemcpy(all.data(), recv.p, sizeof(Pack) * G.world, cudaMemcpyDeviceToHost));
  std::vector<double*> px((size_t)G.world);
  std::vector<unsigned long long*> pf((size_t)G.world);
  for (int p = 0; p < G.world; ++p) {
    if (p == G.rank) { px[p] = G.xbuf.p; pf[p] = G.xflag.p; continue; }
    void *a = nullptr, *b = nullptr;
    GB_TRY(cudaIpcOpenMemHandle(&a, all[p].x, cudaIpcMemLazyEnablePeerAccess));
    G.ipc_opened.push_back(a);
    GB_TRY(cudaIpcOpenMemHandle(&b, all[p].f, cudaIpcMemLazyEnablePeerAccess));
    G.ipc_opened.push_back(b);
    px[p] = static_cast<double*>(a); pf[p] = static_cast<unsigned long long*>(b);
  }
  GB_TRY(cudaMemcpy(G.peer_xbuf.p, px.data(), sizeof(double*) * G.world, cudaMemcpyHostToDevice));
  GB_TRY(cudaMemcpy(G.peer_xflag.p, pf.data(), sizeof(unsigned long long*) * G.world, cudaMemcpyHostToDevice));
  return BVB_OK;
}

int bvb_gibbs_init(bvb_handle* h, const double* Y, const double* X, int M, int T, int K, int draws, int burnin,
                   int thin, int tvp_keep_all, int intercept, const double* priorIntercept, const double* PHI0,
                   const bvb_prior_phi* pp, const bvb_prior_sigma* ps, const bvb_startvals* sv, const int* i_vec,
                   double PHI_tol, double L_tol, int huge, const bvb_draws* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 0; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (!Y || M <= 0 || T <= 0 || K < 0 || draws < 0 || burnin < 0 || thin < 1 || !pp || !ps || !sv || !out)
    return fail_arg(h, "null or out-of-range argument");
  const int Kp = K + (intercept ? 1 : 0);
  if (Kp > 0 && (!X || !PHI0 || !sv->PHI)) return fail_arg(h, "X / PHI0 / startvals.PHI missing");
  if (K > 0 && !i_vec) return fail_arg(h, "i_vec missing");
  const bool chol = ps->type == BVB_SIGMA_CHOLESKY;
  if (chol && h->world > 1) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "cholesky structure does not shard (Gauss-Seidel over equations): run one chain per GPU");
    return BVB_ERR_UNSUPPORTED;
  }
  const int r = chol ? 0 : ps->factor_factors;
  if (r < 0 || r > FSV_RMAX_BIG) { bvb_set_error(h, BVB_ERR_UNSUPPORTED, "factor_factors=%d outside [0,%d]", r, FSV_RMAX_BIG); return BVB_ERR_UNSUPPORTED; }
  const bool use_huge = huge && !chol && K > 0;   // 'expert_huge=TRUE' is ignored for the cholesky structure (R/bvar_wrapper.R:389-391)
  if (use_huge && !k2_supported(T)) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "expert_huge: T=%d exceeds the in-SM panel of the T x T solve", T);
    return BVB_ERR_UNSUPPORTED;
  }
  if (Kp > 0 && !use_huge && !k2_supported(Kp)) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "Kp=%d exceeds the in-SM panel of the standard path; use expert_huge", Kp);
    return BVB_ERR_UNSUPPORTED;
  }
  GB_TRY(cudaSetDevice(h->device));
  GB_TRY(cudaDeviceSynchronize());
  const bool init_prof = getenv("BVB_INIT_PROF") != nullptr;
  auto init_t0 = std::chrono::steady_clock::now();
  auto init_mark = [&](const char* what) {
    if (!init_prof) return;
    cudaStreamSynchronize(h->stream);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[bvb init] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - init_t0).count());
    init_t0 = t;
  };
  // A handle that runs chain after chain of one shape (the usual case: bvar() called repeatedly, the bench's end-to-end
  // repetitions) keeps its device buffers, streams and events: every buffer below is (re)filled by this function and
  // GBuf::ensure only reallocates when a buffer must grow.  Freeing and reallocating ~0.5 GB cost 6 ms per chain.
  Gibbs* Gp = static_cast<Gibbs*>(h->gibbs);
  const int world_now = h->world > 0 ? h->world : 1;
  const bool reuse = Gp && Gp->M == M && Gp->T == T && Gp->K == K && Gp->Kp == Kp && Gp->r == r && Gp->sigma_type == ps->type &&
                     Gp->huge == (use_huge ? 1 : 0) && Gp->world == world_now && Gp->rank == h->rank &&
                     Gp->tvp_all == (tvp_keep_all ? 1 : 0) && !getenv("BVB_NO_REUSE");
  if (reuse) {
    GB_TRY(cudaStreamSynchronize(h->stream));
    // the executable sweep graph is kept: kernel arguments (seed, priors) change, the topology does not, so the first
    // run of the new chain re-captures one sweep and patches the executable in place (cudaGraphExecUpdate)
    Gp->graph_stale = Gp->graph_exec != nullptr;
    Gp->ready = false; Gp->k1z_valid = false; Gp->timed = false; Gp->timing_pass = false; Gp->rep = 0;
    for (auto& v : Gp->phase_ms) v = 0;
  } else {
    gibbs_free(h);
    Gp = new (std::nothrow) Gibbs();
    if (!Gp) return BVB_ERR_CUDA;
    h->gibbs = Gp;
  }
  init_mark(reuse ? "previous chain recycled" : "free previous chain");
  gbuf_stream() = h->stream;
  Gibbs& G = *Gp;
  if (!G.stage) {
    const size_t cap = (size_t)16 << 20;
    if (cudaMallocHost(reinterpret_cast<void**>(&G.stage), cap) == cudaSuccess) G.stage_cap = cap;
    else { G.stage = nullptr; G.stage_cap = 0; cudaGetLastError(); }
  }
  G.stage_off = 0;   // (the stream was synchronised above when the chain is recycled; a new Gibbs has a new arena)
  G.M = M; G.T = T; G.K = K; G.Kp = Kp; G.intercept = intercept ? 1 : 0; G.r = r; G.S = M + r;
  G.draws = draws; G.burnin = burnin; G.thin = thin; G.tvp_all = tvp_keep_all ? 1 : 0;
  G.nsave = draws / thin; G.PHI_tol = PHI_tol; G.L_tol = L_tol; G.out = *out;
  G.sigma_type = ps->type;
  G.huge = use_huge ? 1 : 0;
  G.rank = h->rank; G.world = h->world > 0 ? h->world : 1; G.comm = h->comm;
  G.seed = h->seed;   // the chain keeps the key it was set up with (bvb_set_seed only affects later chains)
  G.nloc_max = (M + G.world - 1) / G.world;
  G.j0 = std::min(M, G.rank * G.nloc_max);
  GB_TRY(G.first_slot.ensure(1, true));
  G.nloc = std::max(0, std::min(M, G.j0 + G.nloc_max) - G.j0);
  const int S = G.S;
  int rc;
  cudaStream_t s = h->stream;

  // ---- data and PHI block
  if ((rc = upload(h, G.Y, Y, (size_t)T * M))) return rc;
  GB_TRY(G.resid.ensure((size_t)T * M));
  GB_TRY(G.status.ensure(8, true));
  GB_TRY(G.sweep.ensure(1, true));
  GB_TRY(G.nstored.ensure(2, true));
  GB_TRY(G.ktime.ensure(8, true));
  {
    const unsigned long long init[8] = {~0ull, 0, ~0ull, 0, 0, 0, 0, 0};
    GB_TRY(cudaMemcpyAsync(G.ktime.p, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  }
  GB_TRY(G.eqstatus.ensure(std::max(M, 1), true));
  if (Kp > 0) {
    // (a recycled chain on the same regressors - the usual case: bvar() again on the same data - keeps X and the
    // pair-product matrix Z on the device)
    const bool sameX = reuse && G.Xhost.size() == (size_t)T * Kp && memcmp(G.Xhost.data(), X, sizeof(double) * (size_t)T * Kp) == 0;
    if (!sameX) { if ((rc = upload(h, G.X, X, (size_t)T * Kp))) return rc; }
    if ((rc = upload(h, G.PHI0, PHI0, (size_t)Kp * M))) return rc;
    GB_TRY(G.PHI.ensure((size_t)Kp * std::max(M, G.nloc_max * G.world)));
    if ((rc = upload(h, G.PHI, sv->PHI, (size_t)Kp * M))) return rc;
    // peer exchange instead of the per-sweep NCCL allgather (BVB_NCCL_EXCHANGE=1 keeps the collective)
    G.p2p = false; G.p2p_wanted = false;
    if (G.world > 1 && ps->type != BVB_SIGMA_CHOLESKY && !use_huge && !getenv("BVB_NCCL_EXCHANGE")) {
      const size_t blk = (size_t)G.nloc_max * Kp;
      const double* old_x = G.xbuf.p;
      const unsigned long long* old_f = G.xflag.p;
      GB_TRY(G.xbuf.ensure((size_t)4 * G.world * blk));
      GB_TRY(G.xflag.ensure((size_t)G.world));
      GB_TRY(G.peer_xbuf.ensure((size_t)G.world)); GB_TRY(G.peer_xflag.ensure((size_t)G.world));
      GB_TRY(G.xcnt.ensure((size_t)G.world + 1, true));
      h->p2p_epoch += 1;                       // (every rank sets up the same chains in the same order)
      G.epoch_base = h->p2p_epoch << 32;
      G.p2p_wanted = true;
      if (!h->inproc) {
        if (old_x != G.xbuf.p || old_f != G.xflag.p || G.ipc_opened.empty()) {
          // all ranks or none: a rank that cannot open its peers' buffers (IPC disabled in a container, no peer access)
          // must not leave the others waiting on flags - the outcome is agreed on through one more tiny allgather, and
          // the per-sweep NCCL allgather remains the fallback
          const int brc = p2p_bootstrap_ipc(h, G);
          GBuf<double> okv, okall;
          GB_TRY(okv.ensure(1)); GB_TRY(okall.ensure((size_t)G.world));
          const double mine_ok = (brc == BVB_OK) ? 1.0 : 0.0;
          GB_TRY(cudaMemcpyAsync(okv.p, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, h->stream));
          ncclResult_t nr = nccl_api().AllGather(okv.p, okall.p, 1, ncclDouble, (ncclComm_t)G.comm, h->stream);
          if (nr != ncclSuccess) { bvb_set_error(h, BVB_ERR_COMM, "ncclAllGather (set-up): %s", nccl_api().GetErrorString(nr)); return BVB_ERR_COMM; }
          GB_TRY(cudaStreamSynchronize(h->stream));
          std::vector<double> oks((size_t)G.world, 0.0);
          GB_TRY(cudaMemcpy(oks.data(), okall.p, sizeof(double) * G.world, cudaMemcpyDeviceToHost));
          bool all_ok = true;
          for (double v : oks) all_ok = all_ok && v > 0.5;
          if (!all_ok) {
            for (void* q : G.ipc_opened) cudaIpcCloseMemHandle(q);
            G.ipc_opened.clear();
            cudaGetLastError();
            h->err_code = 0; h->err_msg[0] = 0;
          }
          G.p2p = all_ok;
        } else {
          G.p2p = true;
        }
      }                                        // (in-process ranks are wired by bvb_gibbs_init_multi)
    }
    // the pair geometry (80 k pairs at the headline shape: 0.4 ms of host loops and 1.3 MB of tables) only depends on
    // (T, Kp): a recycled chain keeps it, on the host and on the device
    if (!reuse || G.geom.T != T || G.geom.n != Kp) {
      make_pair_geom(G.geom, T, Kp);
      if ((rc = upload(h, G.lut, G.geom.lut.data(), G.geom.nArows))) return rc;
      if ((rc = upload(h, G.pairs, G.geom.pairs.data(), G.geom.nArows))) return rc;
      if ((rc = upload(h, G.colbase, G.geom.colbase.data(), G.geom.colbase.size()))) return rc;
    }
    if (!use_huge) GB_TRY(G.A.ensure((size_t)G.geom.nArows * G.geom.Tp));
    const int Nalloc = ((M + 7) / 8) * 8 + K1_NMAX;
    GB_TRY(G.W.ensure((size_t)Nalloc * G.geom.Tp, true));
    GB_TRY(G.WY.ensure((size_t)Nalloc * G.geom.Tp, true));
    if (!use_huge) GB_TRY(G.omega.ensure((size_t)std::max(G.nloc, 1) * G.geom.ldo));
    GB_TRY(G.Zn.ensure((size_t)Kp * std::max(use_huge ? M : G.nloc, 1)));
    // the pair-product matrix Z is built ONCE per chain (X never changes) - and kept when the next chain has the same X
    if (!sameX) G.Xhost.assign(X, X + (size_t)T * Kp);
    if (!use_huge && !sameX) GB_TRY(launch_build_A(G.A.p, G.X.p, G.pairs.p, T, G.geom.Tp, Kp, G.geom.nArows, 0, s));
    if (use_huge) {
      make_pair_geom(G.geomH, Kp, T);   // pairs of observations, contraction over the coefficients
      const int Kp_pad = G.geomH.Tp;
      if ((rc = upload(h, G.luth, G.geomH.lut.data(), G.geomH.nArows))) return rc;
      if ((rc = upload(h, G.pairsh, G.geomH.pairs.data(), G.geomH.nArows))) return rc;
      if ((rc = upload(h, G.colbaseh, G.geomH.colbase.data(), G.geomH.colbase.size()))) return rc;
      const size_t tm2 = (size_t)T * M;
      GB_TRY(G.XT.ensure((size_t)T * Kp)); GB_TRY(G.Ah.ensure((size_t)G.geomH.nArows * Kp_pad));
      GB_TRY(G.Wv.ensure((size_t)Nalloc * Kp_pad, true)); GB_TRY(G.WYzh.ensure((size_t)Nalloc * Kp_pad, true));
      GB_TRY(G.omegah.ensure((size_t)M * G.geomH.ldo)); GB_TRY(G.hu.ensure((size_t)Kp * M)); GB_TRY(G.hnrm.ensure(tm2));
      GB_TRY(G.hwsol.ensure(tm2)); GB_TRY(G.hzero.ensure(tm2, true)); GB_TRY(G.hdelta.ensure(tm2));
      GB_TRY(launch_huge_setup(G.XT.p, G.Ah.p, G.X.p, G.pairsh.p, T, Kp, Kp_pad, G.geomH.nArows, s));
    }
  }

  init_mark("data + pair products");
  // ---- prior variances V_prior (Kp x M): V_i on the lag rows, priorIntercept on the last row
  const int n = K * M;
  std::vector<double> Vprior((size_t)std::max(Kp * M, 1), 1.0), V_init((size_t)std::max(n, 1), 1.0);
  std::vector<double> loc1((size_t)std::max(n, 1), 0.0), loc2((size_t)std::max(n, 1), 1.0), V_prep_small;
  std::vector<int> grp((size_t)std::max(n, 1), -1);
  std::vector<double> gpar;
  int Gn = 0;
  const int prior = (K == 0) ? BVB_PRIOR_NOTHING : pp->prior;
  if (K > 0) {
    // i_vec_small = i_vec without intercept entries, same order as vec(PHI[0:K,])
    std::vector<int> ivs((size_t)n);
    for (int j = 0; j < M; ++j) for (int k = 0; k < K; ++k) ivs[(size_t)j * K + k] = i_vec[(size_t)j * Kp + k];
    if (prior == BVB_PRIOR_HMP) {
      Gn = 2;  // 0 = own-lag (i_vec_small > 0), 1 = cross-lag (< 0)   bvar_cpp.cpp:46-49
      gpar.assign((size_t)Gn * HYP_GP, 0.0);
      double cnt[2] = {0, 0};
      V_prep_small.resize(n);
      if (!pp->V_i_prep) return fail_arg(h, "HMP needs V_i_prep");
      for (int j = 0; j < M; ++j) for (int k = 0; k < K; ++k) {
        const size_t i = (size_t)j * K + k;
        V_prep_small[i] = pp->V_i_prep[(size_t)j * Kp + k];   // V_i_prep(i_ocl)
        grp[i] = ivs[i] > 0 ? 0 : (ivs[i] < 0 ? 1 : -1);
        if (grp[i] >= 0) cnt[grp[i]] += 1;
      }
      const double lam_init[2] = {0.04, 0.0016};              // bvar_cpp.cpp:144-145
      for (int g = 0; g < 2; ++g) {
        gpar[g * HYP_GP + GP_A] = (g == 0 ? pp->lambda_1 : pp->lambda_2)[0];
        gpar[g * HYP_GP + GP_B] = (g == 0 ? pp->lambda_1 : pp->lambda_2)[1];
        gpar[g * HYP_GP + GP_AUX] = lam_init[g];
        gpar[g * HYP_GP + GP_N] = cnt[g];
      }
      for (int i = 0; i < n; ++i) V_init[i] = (grp[i] >= 0 ? lam_init[grp[i]] : 1.0) * V_prep_small[i];
    } else if (prior != BVB_PRIOR_NORMAL) {
      Gn = pp->n_groups;
      if (Gn <= 0 || !pp->groups) return fail_arg(h, "n_groups / groups missing");
      gpar.assign((size_t)Gn * HYP_GP, 0.0);
      for (int g = 0; g < Gn; ++g) {
        double cnt = 0;
        for (int i = 0; i < n; ++i) if (ivs[i] == pp->groups[g]) { grp[i] = g; cnt += 1; }
        gpar[g * HYP_GP + GP_A] = pp->a ? pp->a[g] : 0.0;
        gpar[g * HYP_GP + GP_B] = pp->b ? pp->b[g] : 0.0;
        gpar[g * HYP_GP + GP_C] = pp->c ? pp->c[g] : 0.0;
        gpar[g * HYP_GP + GP_XI] = 0.5;       // bvar_cpp.cpp:85
        gpar[g * HYP_GP + GP_ZETA] = 10.0;    // :125
        gpar[g * HYP_GP + GP_VARPI] = 1.0;    // :127
        gpar[g * HYP_GP + GP_N] = cnt;
      }
    }
    switch (prior) {
      case BVB_PRIOR_NORMAL:
        if (!pp->V_i) return fail_arg(h, "prior normal needs V_i");
        for (int i = 0; i < n; ++i) V_init[i] = pp->V_i[i];
        break;
      case BVB_PRIOR_DL:  // lambda = 1/n, psi = 1, V = psi lambda^2  (:83-84,:105-107)
        for (int i = 0; i < n; ++i) { loc1[i] = 1.0 / n; loc2[i] = 1.0; V_init[i] = loc1[i] * loc1[i]; }
        break;
      case BVB_PRIOR_GT:  // :113-119
        for (int i = 0; i < n; ++i) { loc1[i] = 1.0 / n; loc2[i] = 1.0; V_init[i] = pp->GT_priorkernel ? 0.5 * loc1[i] : loc1[i]; }
        break;
      case BVB_PRIOR_HS:  // theta = 1/n, nu = 1, zeta = 10  (:123-129)
        for (int i = 0; i < n; ++i) { loc1[i] = 1.0 / n; loc2[i] = 1.0; V_init[i] = loc1[i] * 10.0; }
        break;
      case BVB_PRIOR_SSVS:  // :131-145
        if (!pp->SSVS_tau0 || !pp->SSVS_tau1 || !pp->SSVS_p) return fail_arg(h, "SSVS needs tau0/tau1/p");
        for (int i = 0; i < n; ++i) {
          loc1[i] = 0.0; loc2[i] = pp->SSVS_p[i];
          V_init[i] = pp->SSVS_hyper ? pp->SSVS_tau1[i] * pp->SSVS_tau1[i] : pp->SSVS_tau0[i] * pp->SSVS_tau0[i];
        }
        break;
      default: break;
    }
    for (int j = 0; j < M; ++j) for (int k = 0; k < K; ++k) Vprior[(size_t)j * Kp + k] = V_init[(size_t)j * K + k];
  }
  if (intercept) {
    if (!priorIntercept) return fail_arg(h, "priorIntercept missing");
    for (int j = 0; j < M; ++j) Vprior[(size_t)j * Kp + K] = priorIntercept[j];
  }
  if (Kp > 0) if ((rc = upload(h, G.Vprior, Vprior.data(), (size_t)Kp * M))) return rc;
  if (gpar.size() < (size_t)std::max(Gn, 1) * HYP_GP) gpar.resize((size_t)std::max(Gn, 1) * HYP_GP, 0.0);
  const int is_dl = prior == BVB_PRIOR_DL;
  rc = setup_hyper(h, G.hphi, prior, n, grp, Gn, gpar, pp->V_i, pp->GL_tol, pp->a_vec, pp->a_weight, pp->norm_consts,
                   pp->c_vec, (prior == BVB_PRIOR_DL || prior == BVB_PRIOR_GT) ? pp->grid_len : 0,
                   is_dl ? pp->DL_hyper : (prior == BVB_PRIOR_GT ? pp->GT_hyper : (prior == BVB_PRIOR_SSVS ? pp->SSVS_hyper : 0)),
                   pp->DL_plus, pp->c_rel_a, pp->GT_priorkernel, pp->GT_vs, pp->SSVS_tau0, pp->SSVS_tau1,
                   pp->SSVS_s_a, pp->SSVS_s_b, pp->SSVS_p, V_prep_small, loc1, loc2, V_init, STREAM_PHI_HYPER);
  if (rc) return rc;
  G.hphi.d.status = G.status.p + 2;

  init_mark("prior block");
  // ---- Sigma_t block (factor SV)
  FsvDev& F = G.fsv;
  F.M = M; F.T = T; F.r = r;
  if (!sv->sv_logvar || !sv->sv_para || !sv->sv_logvar0) return fail_arg(h, "sv start values missing");
  if ((rc = upload(h, G.logvar, sv->sv_logvar, (size_t)T * S))) return rc;
  if ((rc = upload(h, G.logvar0, sv->sv_logvar0, S))) return rc;
  if ((rc = upload(h, G.sv_para, sv->sv_para, (size_t)3 * S))) return rc;
  std::vector<int> restr((size_t)std::max(M * r, 1), 1);
  std::vector<double> facload((size_t)std::max(M * r, 1), 0.0), tau2((size_t)std::max(M * r, 1), 1.0);
  if (r > 0) {
    if (!sv->facload || !sv->fac || !sv->tau2) return fail_arg(h, "factor start values missing");
    for (size_t i = 0; i < (size_t)M * r; ++i) {
      restr[i] = ps->factor_restrinv ? (ps->factor_restrinv[i] != 0) : 1;
      facload[i] = restr[i] ? sv->facload[i] : 0.0;    // bvar_cpp.cpp:305-312
      tau2[i] = restr[i] ? sv->tau2[i] : 0.0;
    }
    if ((rc = upload(h, G.fac, sv->fac, (size_t)r * T))) return rc;
  }
  if ((rc = upload(h, G.facload, facload.data(), (size_t)std::max(M * r, 1)))) return rc;
  if ((rc = upload(h, G.tau2, tau2.data(), (size_t)std::max(M * r, 1)))) return rc;
  if ((rc = upload(h, G.restr, restr.data(), (size_t)std::max(M * r, 1)))) return rc;
  GB_TRY(G.lambda2.ensure(std::max(M, std::max(r, 1)), true));
  std::vector<int> hetero(S, 1);
  std::vector<double> offset(M, 0.0), ps2((size_t)2 * S, 0.5), ph0(S, -1.0), phomo((size_t)2 * M, 1.0);
  for (int i = 0; i < S; ++i) {
    if (ps->sv_heteroscedastic) hetero[i] = ps->sv_heteroscedastic[i] != 0;
    if (ps->sv_priorsigma2) { ps2[i] = ps->sv_priorsigma2[i]; ps2[S + i] = ps->sv_priorsigma2[S + i]; }
    if (ps->sv_priorh0) ph0[i] = ps->sv_priorh0[i];
  }
  if (!chol && ps->factor_priorhomoskedastic) for (int i = 0; i < 2 * M; ++i) phomo[i] = ps->factor_priorhomoskedastic[i];
  if (chol && ps->cholesky_priorhomoscedastic) for (int i = 0; i < 2 * M; ++i) phomo[i] = ps->cholesky_priorhomoscedastic[i];
  if (chol && ps->cholesky_sv_offset) for (int i = 0; i < M; ++i) offset[i] = ps->cholesky_sv_offset[i];
  if ((rc = upload(h, G.hetero, hetero.data(), S))) return rc;
  if ((rc = upload(h, G.offset, offset.data(), M))) return rc;
  if ((rc = upload(h, G.priorsigma2, ps2.data(), (size_t)2 * S))) return rc;
  if ((rc = upload(h, G.priorh0, ph0.data(), S))) return rc;
  if ((rc = upload(h, G.priorhomo, phomo.data(), (size_t)2 * M))) return rc;
  const int units = ps->factor_columnwise ? r : M;
  if (r > 0 && ps->factor_ngprior) {
    if (!ps->factor_shrink_a || !ps->factor_shrink_c || !ps->factor_shrink_d) return fail_arg(h, "factor_shrinkagepriors missing");
    if ((rc = upload(h, G.shrink_a, ps->factor_shrink_a, units))) return rc;
    if ((rc = upload(h, G.shrink_c, ps->factor_shrink_c, units))) return rc;
    if ((rc = upload(h, G.shrink_d, ps->factor_shrink_d, units))) return rc;
  }
  GB_TRY(G.ystar.ensure((size_t)T * S));
  GB_TRY(G.eps.ensure((size_t)(T + 1) * S));
  GB_TRY(G.mixind.ensure((size_t)T * S));
  GB_TRY(G.svscratch.ensure((size_t)3 * (T + 1) * S));
  GB_TRY(G.eidi.ensure((size_t)T * M));
  GB_TRY(G.wexp.ensure((size_t)T * M));
  F.resid = G.resid.p; F.logvar = G.logvar.p; F.logvar0 = G.logvar0.p; F.sv_para = G.sv_para.p;
  F.facload = G.facload.p; F.fac = G.fac.p; F.tau2 = G.tau2.p; F.lambda2 = G.lambda2.p;
  F.ystar = G.ystar.p; F.eps = G.eps.p; F.mixind = G.mixind.p; F.scratch = G.svscratch.p; F.eidi = G.eidi.p; F.wexp = G.wexp.p;
  F.status = G.status.p + 1;
  F.hetero = G.hetero.p; F.restr = G.restr.p; F.offset = G.offset.p; F.priorsigma2 = G.priorsigma2.p;
  F.priorh0 = G.priorh0.p; F.priorhomo = G.priorhomo.p;
  F.shrink_a = G.shrink_a.p; F.shrink_c = G.shrink_c.p; F.shrink_d = G.shrink_d.p;
  F.priormu[0] = ps->sv_priormu[0]; F.priormu[1] = ps->sv_priormu[1];
  for (int i = 0; i < 4; ++i) F.priorphi[i] = ps->sv_priorphi[chol ? i % 2 : i];
  F.facloadtol = ps->factor_facloadtol; F.ngprior = ps->factor_ngprior; F.columnwise = ps->factor_columnwise;
  F.interweaving = ps->factor_interweaving > 4 ? 4 : ps->factor_interweaving;
  F.prof = getenv("BVB_FSV_PROF") ? 1 : 0;
  if (Kp == 0) GB_TRY(cudaMemcpyAsync(G.resid.p, G.Y.p, sizeof(double) * (size_t)T * M, cudaMemcpyDeviceToDevice, s));
  if (use_huge) {
    HugeCtx& c = G.chuge;
    c.T = T; c.Kp = Kp; c.Kp_pad = G.geomH.Tp; c.M = M; c.r = r; c.X = G.X.p; c.Y = G.Y.p; c.logvar = G.logvar.p;
    c.V = G.Vprior.p; c.PHI0 = G.PHI0.p; c.facload = G.facload.p; c.fac = G.fac.p; c.z1 = G.Zn.p; c.delta = G.hdelta.p;
    c.PHI = G.PHI.p; c.A = G.Ah.p; c.lut_h = G.luth.p; c.pairs_h = G.pairsh.p; c.colbase_h = G.colbaseh.p;
    c.nZtiles = G.geomH.nZtiles; c.ntiles = G.geomH.nZtiles + G.geomH.nXtiles; c.ldo = G.geomH.ldo; c.Wv = G.Wv.p;
    c.WYzero = G.WYzh.p; c.omega = G.omegah.p; c.u = G.hu.p; c.nrm = G.hnrm.p; c.wsol = G.hwsol.p; c.zero_tm = G.hzero.p;
    c.status = G.eqstatus.p;
  }

  // ---- cholesky structure: U, its prior block, and the contexts of the two sequential draws
  int priorU = BVB_PRIOR_NORMAL;
  const int nU = M * (M - 1) / 2;
  G.nU = chol ? nU : 0;
  if (chol) {
    if (!sv->U) return fail_arg(h, "startvals.U missing");
    const size_t tm = (size_t)T * M;
    if ((rc = upload(h, G.U, sv->U, (size_t)M * M))) return rc;
    GB_TRY(G.dsq.ensure(tm)); GB_TRY(G.d2inv.ensure(tm)); GB_TRY(G.XPHI.ensure(tm)); GB_TRY(G.Q.ensure(tm));
    GB_TRY(G.strres.ensure(tm)); GB_TRY(G.rvec.ensure(T)); GB_TRY(G.bvec.ensure(std::max(Kp, 1)));
    GB_TRY(G.zflat.ensure(std::max(nU, 1))); GB_TRY(G.ustatus.ensure(M, true));
    chol_dsq_kernel<<<(unsigned)((tm + 255) / 256), 256, 0, s>>>(G.dsq.p, G.logvar.p, tm);   // bvar_cpp.cpp:404
    F.resid = G.strres.p;   // the SV kernels see the orthogonalised residuals
    if (Kp > 0) {
      const int Nalloc = ((M + 7) / 8) * 8 + K1_NMAX;
      GB_TRY(G.WYzero.ensure((size_t)Nalloc * G.geom.Tp, true));
      CholPhiCtx& c = G.cphi;
      c.T = T; c.Tp = G.geom.Tp; c.Kp = Kp; c.M = M; c.X = G.X.p; c.Y = G.Y.p; c.PHI0 = G.PHI0.p; c.V = G.Vprior.p;
      c.U = G.U.p; c.dsq = G.dsq.p; c.Zn = G.Zn.p; c.PHI = G.PHI.p; c.A = G.A.p; c.lut = G.lut.p; c.colbase = G.colbase.p;
      c.nZtiles = G.geom.nZtiles; c.ntiles = G.geom.nZtiles + G.geom.nXtiles; c.ldo = G.geom.ldo; c.W = G.W.p;
      c.WYzero = G.WYzero.p; c.omega = G.omega.p; c.d2inv = G.d2inv.p; c.resid = G.resid.p; c.XPHI = G.XPHI.p; c.Q = G.Q.p;
      c.rvec = G.rvec.p; c.bvec = G.bvec.p; c.status = G.eqstatus.p;
    }
    // prior on the free elements of U (bvar_cpp.cpp:166-245)
    priorU = ps->cholesky_U_prior;
    std::vector<double> l1((size_t)std::max(nU, 1), 0.0), l2((size_t)std::max(nU, 1), 1.0), Vi((size_t)std::max(nU, 1), 1.0), Vprep;
    std::vector<int> grpU((size_t)std::max(nU, 1), 0);
    std::vector<double> gp(HYP_GP, 0.0);
    gp[GP_A] = ps->cholesky_a; gp[GP_B] = ps->cholesky_b; gp[GP_C] = ps->cholesky_c; gp[GP_XI] = 0.5;
    gp[GP_ZETA] = 10.0; gp[GP_VARPI] = 1.0; gp[GP_N] = nU;
    switch (priorU) {
      case BVB_PRIOR_NORMAL:
        if (nU > 0 && !ps->cholesky_V_i) return fail_arg(h, "cholesky_V_i missing");
        for (int i = 0; i < nU; ++i) Vi[i] = ps->cholesky_V_i[i];
        break;
      case BVB_PRIOR_DL:   // V_i_U = psi_U % lambda_U (:214)
        for (int i = 0; i < nU; ++i) { l1[i] = 1.0 / nU; l2[i] = 1.0; Vi[i] = l1[i]; }
        break;
      case BVB_PRIOR_GT:
        for (int i = 0; i < nU; ++i) { l1[i] = 1.0 / nU; l2[i] = 1.0; Vi[i] = ps->cholesky_GT_priorkernel ? 0.5 * l1[i] : l1[i]; }
        break;
      case BVB_PRIOR_HS:
        for (int i = 0; i < nU; ++i) { l1[i] = 1.0 / nU; l2[i] = 1.0; Vi[i] = l1[i] * 10.0; }
        break;
      case BVB_PRIOR_SSVS:
        if (nU > 0 && (!ps->cholesky_SSVS_tau0 || !ps->cholesky_SSVS_tau1 || !ps->cholesky_SSVS_p)) return fail_arg(h, "cholesky SSVS keys missing");
        for (int i = 0; i < nU; ++i) { l1[i] = 0.0; l2[i] = ps->cholesky_SSVS_p[i]; Vi[i] = ps->cholesky_SSVS_tau0[i] * ps->cholesky_SSVS_tau0[i]; }
        break;
      case BVB_PRIOR_HMP:  // lambda_3 = 0.001, V_i_U = lambda_3 (:239-245); sample_V_i_U_HMP (:370-377)
        Vprep.assign((size_t)std::max(nU, 1), 1.0);
        gp[GP_A] = ps->cholesky_lambda_3[0]; gp[GP_B] = ps->cholesky_lambda_3[1]; gp[GP_AUX] = 0.001;
        for (int i = 0; i < nU; ++i) Vi[i] = 0.001;
        break;
      default: return fail_arg(h, "unknown cholesky_U_prior");
    }
    const int hyperU = priorU == BVB_PRIOR_DL ? ps->cholesky_DL_hyper : (priorU == BVB_PRIOR_GT ? ps->cholesky_GT_hyper
                       : (priorU == BVB_PRIOR_SSVS ? ps->cholesky_SSVS_hyper : 0));
    rc = setup_hyper(h, G.hU, priorU, nU, grpU, 1, gp, ps->cholesky_V_i, ps->cholesky_GL_tol, ps->cholesky_a_vec,
                     ps->cholesky_a_weight, ps->cholesky_norm_consts, ps->cholesky_c_vec,
                     (priorU == BVB_PRIOR_DL || priorU == BVB_PRIOR_GT) ? ps->cholesky_grid_len : 0, hyperU,
                     ps->cholesky_DL_plus, ps->cholesky_c_rel_a, ps->cholesky_GT_priorkernel, ps->cholesky_GT_vs,
                     ps->cholesky_SSVS_tau0, ps->cholesky_SSVS_tau1, ps->cholesky_SSVS_s_a, ps->cholesky_SSVS_s_b,
                     ps->cholesky_SSVS_p, Vprep, l1, l2, Vi, STREAM_U_HYPER);
    if (rc) return rc;
    G.hU.d.always_active = 1;
    G.hU.d.status = G.status.p + 2;
    if (nU > 0) {
      const int nu = M - 1;
      if (!k2_supported(nu)) { bvb_set_error(h, BVB_ERR_UNSUPPORTED, "M=%d exceeds the in-SM panel of sample_U", M); return BVB_ERR_UNSUPPORTED; }
      make_pair_geom(G.geomU, T, nu);
      if ((rc = upload(h, G.lutu, G.geomU.lut.data(), G.geomU.nArows))) return rc;
      if ((rc = upload(h, G.pairsu, G.geomU.pairs.data(), G.geomU.nArows))) return rc;
      if ((rc = upload(h, G.colbaseu, G.geomU.colbase.data(), G.geomU.colbase.size()))) return rc;
      const int Nal = ((nu + 7) / 8) * 8 + K1_NMAX;
      GB_TRY(G.Au.ensure((size_t)G.geomU.nArows * G.geomU.Tp));
      GB_TRY(G.Wu.ensure((size_t)Nal * G.geomU.Tp, true)); GB_TRY(G.WYu.ensure((size_t)Nal * G.geomU.Tp, true));
      GB_TRY(G.omegau.ensure((size_t)nu * G.geomU.ldo)); GB_TRY(G.Vu.ensure((size_t)nu * nu));
      GB_TRY(G.Zu.ensure((size_t)nu * nu)); GB_TRY(G.phiu.ensure((size_t)nu * nu));
      CholUCtx& c = G.cu;
      c.T = T; c.Tp = G.geomU.Tp; c.M = M; c.nu = nu; c.resid = G.resid.p; c.dsq = G.dsq.p; c.V_i_U = G.hU.d.V_i;
      c.zflat = G.zflat.p; c.U = G.U.p; c.A = G.Au.p; c.lut = G.lutu.p; c.pairs = G.pairsu.p; c.colbase = G.colbaseu.p;
      c.nZtiles = G.geomU.nZtiles; c.ntiles = G.geomU.nZtiles + G.geomU.nXtiles; c.nArows = G.geomU.nArows; c.ldo = G.geomU.ldo;
      c.W = G.Wu.p; c.WY = G.WYu.p; c.omega = G.omegau.p; c.Vu = G.Vu.p; c.Zu = G.Zu.p; c.phi = G.phiu.p; c.status = G.ustatus.p;
    }
  }

  init_mark("Sigma block");
  // ---- stored-draw ring: one slot = everything bvar_cpp.cpp:751-836 copies per stored sweep
  GibbsStore& st = G.store;
  st.burnin = burnin; st.thin = thin; st.nseg = 0;
  size_t off = 0;
  G.off_PHI = off;      if (Kp > 0) add_seg(st, G.PHI.p, Kp, M, Kp, 0, off);
  G.off_Vprior = off;   if (Kp > 0) add_seg(st, G.Vprior.p, Kp, M, Kp, 0, off);
  const int lv_rows = G.tvp_all ? T : 1;
  G.off_logvar = off;   add_seg(st, G.logvar.p, lv_rows, S, T, G.tvp_all ? 0 : T - 1, off);
  G.off_svpara = off;   add_seg(st, G.sv_para.p, 3, S, 3, 0, off);
  G.off_hyper = off; G.hyper_rows = 0;
  if (prior == BVB_PRIOR_DL || prior == BVB_PRIOR_GT || prior == BVB_PRIOR_HS) {
    // rows: a|zeta [G], xi|varpi [G], lambda|theta [n], psi|nu [n]   (bvar_cpp.cpp:758-779)
    const int c0 = (prior == BVB_PRIOR_HS) ? GP_ZETA : GP_A, c1 = (prior == BVB_PRIOR_HS) ? GP_VARPI : GP_XI;
    add_seg(st, G.hphi.d.gpar, 1, Gn, HYP_GP, c0, off);
    add_seg(st, G.hphi.d.gpar, 1, Gn, HYP_GP, c1, off);
    add_seg(st, G.hphi.d.loc1, n, 1, n, 0, off);
    add_seg(st, G.hphi.d.loc2, n, 1, n, 0, off);
    G.hyper_rows = 2 * Gn + 2 * n;
  } else if (prior == BVB_PRIOR_SSVS) {
    add_seg(st, G.hphi.d.loc1, n, 1, n, 0, off);
    add_seg(st, G.hphi.d.loc2, n, 1, n, 0, off);
    G.hyper_rows = 2 * n;
  } else if (prior == BVB_PRIOR_HMP) {
    add_seg(st, G.hphi.d.gpar, 1, 2, HYP_GP, GP_AUX, off);
    G.hyper_rows = 2;
  }
  G.off_U = off; G.off_uhyper = off; G.uhyper_rows = 0;
  if (chol && nU > 0) {
    add_seg(st, G.hU.d.coef, nU, 1, nU, 0, off);   // u = U(trimatu_ind), refreshed by the clamp kernel every sweep
    G.off_uhyper = off;
    if (priorU == BVB_PRIOR_DL || priorU == BVB_PRIOR_GT || priorU == BVB_PRIOR_HS) {
      const int c0 = (priorU == BVB_PRIOR_HS) ? GP_ZETA : GP_A, c1 = (priorU == BVB_PRIOR_HS) ? GP_VARPI : GP_XI;
      add_seg(st, G.hU.d.gpar, 1, 1, HYP_GP, c0, off);
      add_seg(st, G.hU.d.gpar, 1, 1, HYP_GP, c1, off);
      add_seg(st, G.hU.d.loc1, nU, 1, nU, 0, off);
      add_seg(st, G.hU.d.loc2, nU, 1, nU, 0, off);
      G.uhyper_rows = 2 + 2 * nU;
    } else if (priorU == BVB_PRIOR_SSVS) {
      add_seg(st, G.hU.d.loc1, nU, 1, nU, 0, off);
      add_seg(st, G.hU.d.loc2, nU, 1, nU, 0, off);
      G.uhyper_rows = 2 * nU;
    } else if (priorU == BVB_PRIOR_HMP) {
      add_seg(st, G.hU.d.gpar, 1, 1, HYP_GP, GP_AUX, off);
      G.uhyper_rows = 1;
    }
  }
  G.off_facload = off;  if (r > 0) add_seg(st, G.facload.p, M, r, M, 0, off);
  const int fac_cols = G.tvp_all ? T : 1;
  G.off_fac = off;      if (r > 0) add_seg(st, G.fac.p + (G.tvp_all ? 0 : (size_t)(T - 1) * r), r, fac_cols, r, 0, off);
  st.slot_doubles = off;
  // ring sized for one sub-chunk of bvb_gibbs_run (<= 32 sweeps) and bounded to 1 GiB
  size_t slots = 32;
  while (slots > 1 && slots * off * sizeof(double) > ((size_t)1 << 30)) slots /= 2;
  st.ring_slots = (int)slots;
  GB_TRY(G.ring.ensure(slots * off));
  st.ring = G.ring.p;
  st.first_slot = G.first_slot.p;
  for (int b = 0; b < 2; ++b) {
    const size_t need = slots * off * sizeof(double);
    if (h->pin_bytes[b] < need) {
      if (h->pin[b]) cudaFreeHost(h->pin[b]);
      h->pin[b] = nullptr; h->pin_bytes[b] = 0;
      GB_TRY(cudaMallocHost(&h->pin[b], need));
      h->pin_bytes[b] = need;
    }
    G.pinned[b] = static_cast<double*>(h->pin[b]);
  }
  init_mark("ring + pinned staging");
  for (auto& e : G.ev) if (!e) GB_TRY(cudaEventCreate(&e));
  if (getenv("BVB_NO_FORK") && G.side) { cudaStreamDestroy(G.side); G.side = nullptr; }   // (a recycled chain honours the switch)
  if (!getenv("BVB_NO_FORK") && !G.side) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    GB_TRY(cudaStreamCreateWithPriority(&G.side, cudaStreamNonBlocking, lo));
  }
  if (G.side && !G.ev_fork) {
    GB_TRY(cudaEventCreateWithFlags(&G.ev_fork, cudaEventDisableTiming));
    GB_TRY(cudaEventCreateWithFlags(&G.ev_join, cudaEventDisableTiming));
    GB_TRY(cudaEventCreateWithFlags(&G.ev_ng, cudaEventDisableTiming));
    GB_TRY(cudaEventCreateWithFlags(&G.ev_series, cudaEventDisableTiming));
    GB_TRY(cudaEventCreateWithFlags(&G.ev_k1z, cudaEventDisableTiming));
  }
  if (!G.copy) {
    GB_TRY(cudaEventCreate(&G.ev_chunk[0]));
    GB_TRY(cudaEventCreate(&G.ev_chunk[1]));
    GB_TRY(cudaStreamCreateWithFlags(&G.copy, cudaStreamNonBlocking));
    GB_TRY(cudaMallocHost(&G.first_slot_host, sizeof(int) * Gibbs::FS_SLOTS));
    for (int b = 0; b < 2; ++b) GB_TRY(cudaEventCreateWithFlags(&G.ev_sub[b], cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) GB_TRY(cudaEventCreateWithFlags(&G.ev_pin[b], cudaEventDisableTiming));
  }
  GB_TRY(cudaStreamSynchronize(s));
  G.rep = 0;
  G.ready = true;
  return BVB_OK;
}

// Copy slots [first, first+count) of the pinned chunk into the caller's arrays.
static void scatter_to_caller(Gibbs& G, const double* chunk, int first, int count) {
  const size_t sd = G.store.slot_doubles;
  const int M = G.M, Kp = G.Kp, S = G.S, T = G.T, r = G.r;
  const int lv_rows = G.tvp_all ? T : 1, fac_cols = G.tvp_all ? T : 1;
  for (int q = 0; q < count; ++q) {
    const double* src = chunk + (size_t)q * sd;
    const size_t d = (size_t)(first + q);
    if (G.out.PHI && Kp > 0) memcpy(G.out.PHI + d * Kp * M, src + G.off_PHI, sizeof(double) * Kp * M);
    if (G.out.V_prior && Kp > 0) memcpy(G.out.V_prior + d * Kp * M, src + G.off_Vprior, sizeof(double) * Kp * M);
    if (G.out.logvar) memcpy(G.out.logvar + d * lv_rows * S, src + G.off_logvar, sizeof(double) * lv_rows * S);
    if (G.out.sv_para) memcpy(G.out.sv_para + d * 3 * S, src + G.off_svpara, sizeof(double) * 3 * S);
    if (G.out.phi_hyperparameter && G.hyper_rows > 0)
      memcpy(G.out.phi_hyperparameter + d * G.hyper_rows, src + G.off_hyper, sizeof(double) * G.hyper_rows);
    if (G.out.U && G.nU > 0) memcpy(G.out.U + d * G.nU, src + G.off_U, sizeof(double) * G.nU);
    if (G.out.u_hyperparameter && G.uhyper_rows > 0)
      memcpy(G.out.u_hyperparameter + d * G.uhyper_rows, src + G.off_uhyper, sizeof(double) * G.uhyper_rows);
    if (G.out.facload && r > 0) memcpy(G.out.facload + d * M * r, src + G.off_facload, sizeof(double) * M * r);
    if (G.out.fac && r > 0) memcpy(G.out.fac + d * r * fac_cols, src + G.off_fac, sizeof(double) * r * fac_cols);
  }
}

// Number of stored draws after `rep` sweeps.
static inline int stored_count(const Gibbs& G, int rep) { return (rep <= G.burnin) ? 0 : (rep - G.burnin) / G.thin; }

// One sweep launched eagerly (first sweep of a chain: lazily sized workspaces, kernel attributes; or the per-phase
// timing pass).  It is an ordinary sweep of the chain: the sweep counter advances and its draw is stored.
static int run_eager_sweep(bvb_handle* h, Gibbs& G, bool timing) {
  cudaStream_t s = h->stream;
  const int stored_before = stored_count(G, G.rep);
  G.first_slot_host[0] = stored_before;
  GB_TRY(cudaMemcpyAsync(G.first_slot.p, &G.first_slot_host[0], sizeof(int), cudaMemcpyHostToDevice, s));
  G.timing_pass = timing;
  int rc = enqueue_sweep(h, G);
  G.timing_pass = false;
  if (rc) return rc;
  G.rep += 1;
  if (stored_count(G, G.rep) > stored_before) {   // this sweep was stored in ring slot 0
    GB_TRY(cudaMemcpyAsync(G.pinned[0], G.ring.p, sizeof(double) * G.store.slot_doubles, cudaMemcpyDeviceToHost, s));
    GB_TRY(cudaStreamSynchronize(s));
    if (G.rank == 0) scatter_to_caller(G, G.pinned[0], stored_before, 1);
  } else {
    GB_TRY(cudaStreamSynchronize(s));
  }
  if (timing) {
    for (int i = 0; i < 8; ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, G.ev[i], G.ev[i + 1]) != cudaSuccess) { ms = 0; cudaGetLastError(); }
      G.phase_ms[i] = ms;
    }
    G.timed = true;
  }
  return BVB_OK;
}

// Device status words -> error code + the reference's wording.
static int check_status(bvb_handle* h, Gibbs& G) {
  int hst[8];
  GB_TRY(cudaMemcpy(hst, G.status.p, sizeof(hst), cudaMemcpyDeviceToHost));
  std::vector<int> eqs(G.M);
  GB_TRY(cudaMemcpy(eqs.data(), G.eqstatus.p, sizeof(int) * G.M, cudaMemcpyDeviceToHost));
  for (int j = 0; j < G.M; ++j) {
    if (eqs[j] != 0) {
      h->err_step = 1; h->err_eq = j + 1; h->err_sweep = G.rep;
      bvb_set_error(h, BVB_ERR_NUMERIC, "%s failed in run <= %i: univariate_regression_update() failed in %ith eqation: chol(): decomposition failed",
                    G.sigma_type == BVB_SIGMA_CHOLESKY ? "sample_PHI()" : "sample_PHI_factor()", G.rep, j + 1);   // bvar_cpp.cpp:512, :522
      return BVB_ERR_NUMERIC;
    }
  }
  if (G.sigma_type == BVB_SIGMA_CHOLESKY && G.nU > 0) {
    std::vector<int> us(G.M);
    GB_TRY(cudaMemcpy(us.data(), G.ustatus.p, sizeof(int) * G.M, cudaMemcpyDeviceToHost));
    for (int e = 0; e + 1 < G.M; ++e)
      if (us[e] != 0) {
        h->err_step = 3; h->err_eq = e + 2; h->err_sweep = G.rep;
        bvb_set_error(h, BVB_ERR_NUMERIC, "sample_U() failed in run <= %i: univariate_regression_update() failed in equation %i: chol(): decomposition failed", G.rep, e + 2);
        return BVB_ERR_NUMERIC;
      }
  }
  if (hst[0] != 0) {
    h->err_step = 1; h->err_sweep = G.rep;
    bvb_set_error(h, BVB_ERR_NUMERIC, "non-finite PHI in rep <= %i.", G.rep);
    return BVB_ERR_NUMERIC;
  }
  if (hst[1] != 0) {
    h->err_step = 3; h->err_sweep = G.rep;
    bvb_set_error(h, BVB_ERR_NUMERIC, "factorstochvol::update_fsv() failed in run <= %i: small system not positive definite (%s)", G.rep,
                  hst[1] == 1 ? "loadings" : "factors");
    return BVB_ERR_NUMERIC;
  }
  if (hst[3] != 0) {
    h->err_step = 1; h->err_sweep = G.rep;
    bvb_set_error(h, BVB_ERR_COMM, "peer exchange: no PHI block from rank %i within ~35 s (run <= %i)", hst[3] - 1, G.rep);
    return BVB_ERR_COMM;
  }
  if (hst[2] != 0) {
    h->err_step = 2; h->err_sweep = G.rep;
    bvb_set_error(h, BVB_ERR_NUMERIC, "GIG sampler: rejection loop exhausted in run <= %i (parameters outside the sampler's range)", G.rep);
    return BVB_ERR_NUMERIC;
  }
  return BVB_OK;
}

int bvb_gibbs_run(bvb_handle* h, int max_sweeps, int* done) {
  if (!h || !h->gibbs) return BVB_ERR_ARG;
  Gibbs& G = *static_cast<Gibbs*>(h->gibbs);
  if (!G.ready) return BVB_ERR_ARG;
  GB_TRY(cudaSetDevice(h->device));
  gbuf_stream() = h->stream;   // (thread-local: bvb_gibbs_run_multi drives every rank from its own thread)
  cudaStream_t s = h->stream;
  const int tot = G.draws + G.burnin;
  int todo = std::min(max_sweeps, tot - G.rep);
  if (todo < 0) todo = 0;
  h->err_step = 0;
  int ran = 0;
  // First sweep ever: eager (allocates lazily-sized workspaces, sets kernel
  // attributes), then one sweep's launches are captured into a graph.
  if ((!G.graph_exec || G.graph_stale) && todo > 0) {
    const bool run_prof = getenv("BVB_INIT_PROF") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
      if (!run_prof) return;
      auto t = std::chrono::steady_clock::now();
      fprintf(stderr, "[bvb run ] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t0).count());
      t0 = t;
    };
    int rc;
    const bool recycled = G.graph_exec != nullptr && G.graph_stale;
    if (!recycled) {
      rc = run_eager_sweep(h, G, false);
      if (rc) return rc;
      mark("eager first sweep");
      todo -= 1;
      ran += 1;
    } else if (gibbs_rotates(G) && !G.k1z_valid) {
      // A recycled chain (same shape as the one before: workspaces sized, kernel attributes set) needs no eager sweep:
      // what the first sweep does ahead of the steady-state graph - W, normals, the K1 Z tiles and right-hand sides -
      // is launched here, and every sweep including the first replays the patched graph (0.5 ms per chain).
      GB_TRY(enqueue_prep(h, G, s, 1));
      GB_TRY(enqueue_k1z(G, s));
      GB_TRY(enqueue_k1rhs(G, s));
      G.k1z_valid = true;
      mark("first-sweep prologue");
    }
    cudaGraph_t graph;
    GB_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_sweep(h, G);
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc) return rc;
    GB_TRY(ce);
    mark("capture");
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    G.launches_per_sweep = (int)nn;
    if (const char* dot = getenv("BVB_GRAPH_DOT")) cudaGraphDebugDotPrint(graph, dot, 0);
    bool patched = false;
    if (G.graph_exec && G.graph_stale) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(G.graph_exec, graph, &info) == cudaSuccess) patched = true;
      else { cudaGetLastError(); cudaGraphExecDestroy(G.graph_exec); G.graph_exec = nullptr; }
    }
    if (!patched) GB_TRY(cudaGraphInstantiate(&G.graph_exec, graph, 0));
    G.graph_stale = false;
    cudaGraphDestroy(graph);
    mark(patched ? "graph exec update" : "graph instantiate");
  }
  // Sub-chunks of <= 8 sweeps that store <= half a ring.  The two ring halves alternate: the D2H copy of a sub-chunk's
  // stored draws runs on the copy stream (behind an event, off the sampler stream) while the next sub-chunk fills the
  // other half, and the host scatters the PREVIOUS sub-chunk into the caller's arrays meanwhile.
  const int half = std::max(1, G.store.ring_slots / 2);
  const bool two_halves = G.store.ring_slots >= 2;
  GB_TRY(cudaEventRecord(G.ev_chunk[0], s));
  struct Pending { int buf, first, count; } pend = {-1, 0, 0};
  auto flush = [&](Pending& pd) -> int {
    if (pd.buf < 0) return BVB_OK;
    GB_TRY(cudaEventSynchronize(G.ev_pin[pd.buf]));
    if (G.rank == 0) scatter_to_caller(G, G.pinned[pd.buf], pd.first, pd.count);
    pd.buf = -1;
    return BVB_OK;
  };
  int flip = 0, timed_sweeps = 0, nchunk = 0;
  bool closed = false;
  bool used[2] = {false, false};
  while (todo > 0) {
    const int stored_before = stored_count(G, G.rep);
    int nsw = std::min(todo, 8);   // (short sub-chunks: the copy-out and the host-side scatter of a sub-chunk hide behind the next one, only the last is exposed)
    while (nsw > 1 && stored_count(G, G.rep + nsw) - stored_before > half) --nsw;
    // ring half `flip` is free once the copy that last read it is done
    if (used[flip]) GB_TRY(cudaStreamWaitEvent(s, G.ev_pin[flip], 0));
    if (!two_halves && used[flip ^ 1]) GB_TRY(cudaStreamWaitEvent(s, G.ev_pin[flip ^ 1], 0));
    // stored draw that maps to ring slot 0 (pinned staging, one word per sub-chunk of this call: a pageable source
    // would make the copy - and with it the enqueue of the next sub-chunk - wait for the stream)
    if (nchunk > 0 && nchunk % (Gibbs::FS_SLOTS - 1) == 0) GB_TRY(cudaStreamSynchronize(s));   // staging words are reused
    int* fs = &G.first_slot_host[1 + (nchunk++ % (Gibbs::FS_SLOTS - 1))];
    *fs = stored_before - (two_halves ? flip * half : 0);
    GB_TRY(cudaMemcpyAsync(G.first_slot.p, fs, sizeof(int), cudaMemcpyHostToDevice, s));
#ifdef BVB_DEBUG_SWITCHES
    static const bool dbg_eager = getenv("BVB_DEBUG_EAGER") != nullptr;
#else
    constexpr bool dbg_eager = false;
#endif
    for (int i = 0; i < nsw; ++i) {
      if (dbg_eager) { int rc = enqueue_sweep(h, G); if (rc) return rc; }
      else GB_TRY(cudaGraphLaunch(G.graph_exec, s));
    }
    G.rep += nsw;
    ran += nsw;
    todo -= nsw;
    timed_sweeps += nsw;
    // the closing event goes in right behind the last sweep: recorded after the host-side scatter below it would carry
    // the host's copy time (seen as 0.86 instead of 0.76 ms per sweep for calls of 16 + 4 sweeps)
    if (todo == 0) { GB_TRY(cudaEventRecord(G.ev_chunk[1], s)); closed = true; }
    const int cnt = stored_count(G, G.rep) - stored_before;
    Pending cur = {-1, 0, 0};
    if (cnt > 0) {
      GB_TRY(cudaEventRecord(G.ev_sub[flip], s));
      GB_TRY(cudaStreamWaitEvent(G.copy, G.ev_sub[flip], 0));
      GB_TRY(cudaMemcpyAsync(G.pinned[flip], G.ring.p + (size_t)(two_halves ? flip * half : 0) * G.store.slot_doubles,
                             sizeof(double) * (size_t)cnt * G.store.slot_doubles, cudaMemcpyDeviceToHost, G.copy));
      GB_TRY(cudaEventRecord(G.ev_pin[flip], G.copy));
      used[flip] = true;
      cur = {flip, stored_before, cnt};
      flip ^= 1;
    }
    int rc = flush(pend);   // overlaps with the sweeps just queued
    if (rc) return rc;
    pend = cur;
  }
  if (!closed) GB_TRY(cudaEventRecord(G.ev_chunk[1], s));
  {
    int rc = flush(pend);
    if (rc) return rc;
  }
  GB_TRY(cudaStreamSynchronize(s));
  GB_TRY(cudaStreamSynchronize(G.copy));
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, G.ev_chunk[0], G.ev_chunk[1]);
    G.last_chunk_ms = ms;                 // device time of the graph-replayed sweeps of this call
    G.last_chunk_sweeps = timed_sweeps;   // (the eager first sweep of a chain is not part of it)
  }
  if (done) *done = G.rep;
  (void)ran;
  return check_status(h, G);
}

// The tolerance clamp of bvar_cpp.cpp:534-559 on host arrays (the kernel the sweep runs): |PHI - PHI0| on the K lag
// rows is pushed up to `tol`, exact zeros get +tol or -tol with probability 1/2 each (R::rbinom(1, 0.5) there, the
// Philox stream (STREAM_CLAMP, element, sweep) here).  Test hook for the exact-zero branch, which a chain never reaches.
int bvb_tolerance_clamp(bvb_handle* h, int K, int Kp, int M, double* PHI_io, const double* PHI0, double tol, uint32_t sweep) {
  if (!h || !PHI_io || !PHI0 || K < 0 || Kp < K || M <= 0) return BVB_ERR_ARG;
  GB_TRY(cudaSetDevice(h->device));
  const size_t nphi = (size_t)Kp * M;
  double *dphi = nullptr, *dphi0 = nullptr, *dcoef = nullptr;
  uint32_t* dsw = nullptr;
  int* dst = nullptr;
  auto cleanup = [&]() { cudaFree(dphi); cudaFree(dphi0); cudaFree(dcoef); cudaFree(dsw); cudaFree(dst); };
  cudaError_t e = cudaMalloc(&dphi, nphi * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dphi0, nphi * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dcoef, std::max<size_t>((size_t)K * M, 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dsw, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&dst, sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(dphi, PHI_io, nphi * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dphi0, PHI0, nphi * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dsw, &sweep, sizeof(uint32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dst, 0, sizeof(int));
  if (e == cudaSuccess) {
    gibbs_clamp_kernel<<<(unsigned)((nphi + 255) / 256), 256, 0, h->stream>>>(dphi, dphi0, dcoef, K, Kp, M, 1, tol, h->seed, dsw, dst, nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaMemcpy(PHI_io, dphi, nphi * sizeof(double), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) { bvb_set_error(h, BVB_ERR_CUDA, "bvb_tolerance_clamp: %s", cudaGetErrorString(e)); return BVB_ERR_CUDA; }
  return BVB_OK;
}

// One further sweep of the chain, launched eagerly in its serial form with CUDA events between the phases (the Z tiles
// are recomputed at its start so that phase "k1" shows their cost).  out8 = ms of {prep, K1, K2, comm, clamp + hyper,
// resid, Sigma_t, store}.  Outside every timed region: bvb_gibbs_run itself never instruments a sweep.
int bvb_gibbs_profile_sweep(bvb_handle* h, double* out8, int* done) {
  if (!h || !h->gibbs) return BVB_ERR_ARG;
  Gibbs& G = *static_cast<Gibbs*>(h->gibbs);
  if (!G.ready || G.rep >= G.draws + G.burnin) return BVB_ERR_ARG;
  GB_TRY(cudaSetDevice(h->device));
  int rc = run_eager_sweep(h, G, true);
  if (rc) return rc;
  if (out8) for (int i = 0; i < 8; ++i) out8[i] = G.phase_ms[i];
  if (done) *done = G.rep;
  return check_status(h, G);
}

// Live kernel timers (device %globaltimer spans accumulated by the sweeps themselves, graph replay included):
// out3 = {K1 Z-tile ms, K2 ms, sweeps counted} since the last reset.
int bvb_gibbs_kernel_times(bvb_handle* h, double* out3, int reset) {
  if (!h || !h->gibbs) return BVB_ERR_ARG;
  Gibbs& G = *static_cast<Gibbs*>(h->gibbs);
  GB_TRY(cudaSetDevice(h->device));
  GB_TRY(cudaStreamSynchronize(h->stream));
  unsigned long long kt[8];
  GB_TRY(cudaMemcpy(kt, G.ktime.p, sizeof(kt), cudaMemcpyDeviceToHost));
  if (out3) { out3[0] = (double)kt[4] * 1e-6; out3[1] = (double)kt[5] * 1e-6; out3[2] = (double)kt[6]; }
  if (reset) {
    const unsigned long long init[8] = {~0ull, 0, ~0ull, 0, 0, 0, 0, 0};
    GB_TRY(cudaMemcpy(G.ktime.p, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  return BVB_OK;
}

int bvb_gibbs_timing(const bvb_handle* h, double* out10) {
  if (!h || !h->gibbs || !out10) return 0;
  const Gibbs& G = *static_cast<const Gibbs*>(h->gibbs);
  out10[0] = G.last_chunk_ms;
  out10[1] = G.launches_per_sweep;
  for (int i = 0; i < 8; ++i) out10[2 + i] = G.phase_ms[i];
  return G.last_chunk_sweeps;
}

int bvb_gibbs_state(bvb_handle* h, double* PHI, double* V_prior, double* logvar, double* sv_para, double* facload,
                    double* fac, double* U) {
  if (!h || !h->gibbs) return BVB_ERR_ARG;
  Gibbs& G = *static_cast<Gibbs*>(h->gibbs);
  GB_TRY(cudaSetDevice(h->device));
  GB_TRY(cudaStreamSynchronize(h->stream));
  const size_t d = sizeof(double);
  if (PHI && G.Kp) GB_TRY(cudaMemcpy(PHI, G.PHI.p, d * G.Kp * G.M, cudaMemcpyDeviceToHost));
  if (V_prior && G.Kp) GB_TRY(cudaMemcpy(V_prior, G.Vprior.p, d * G.Kp * G.M, cudaMemcpyDeviceToHost));
  if (logvar) GB_TRY(cudaMemcpy(logvar, G.logvar.p, d * G.T * G.S, cudaMemcpyDeviceToHost));
  if (sv_para) GB_TRY(cudaMemcpy(sv_para, G.sv_para.p, d * 3 * G.S, cudaMemcpyDeviceToHost));
  if (facload && G.r) GB_TRY(cudaMemcpy(facload, G.facload.p, d * G.M * G.r, cudaMemcpyDeviceToHost));
  if (fac && G.r) GB_TRY(cudaMemcpy(fac, G.fac.p, d * G.r * G.T, cudaMemcpyDeviceToHost));
  if (U && G.U.p) GB_TRY(cudaMemcpy(U, G.U.p, d * G.M * G.M, cudaMemcpyDeviceToHost));
  return BVB_OK;
}

}  // extern "C"
void bvb_comm_release(bvb_handle* h) {
  if (h && h->comm) { nccl_api().CommDestroy((ncclComm_t)h->comm); h->comm = nullptr; }
}
extern "C" {
int bvb_comm_unique_id(char* unique_id_128) {
  if (!unique_id_128) return BVB_ERR_ARG;
  ncclUniqueId id;
  if (!nccl_api().ok || nccl_api().GetUniqueId(&id) != ncclSuccess) return BVB_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(unique_id_128, &id, 128);
  return BVB_OK;
}

int bvb_comm_init(bvb_handle* h, int rank, int world, const char* unique_id_128) {
  if (!h || world < 1 || rank < 0 || rank >= world) return BVB_ERR_ARG;
  h->rank = rank; h->world = world;
  if (world == 1) return BVB_OK;
  if (!unique_id_128) return BVB_ERR_ARG;
  GB_TRY(cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(&id, unique_id_128, 128);
  if (!nccl_api().ok) { bvb_set_error(h, BVB_ERR_COMM, "libnccl.so.2 could not be loaded"); return BVB_ERR_COMM; }
  ncclComm_t comm;
  ncclResult_t nr = nccl_api().CommInitRank(&comm, world, id, rank);
  if (nr != ncclSuccess) { bvb_set_error(h, BVB_ERR_COMM, "ncclCommInitRank: %s", nccl_api().GetErrorString(nr)); return BVB_ERR_COMM; }
  h->comm = comm;
  return BVB_OK;
}

// ---- one process, several GPUs (the R package is a single process: no launcher, no MPI) ----------------------------
// n handles on n devices with ranks 0..n-1 and an in-process NCCL communicator; the equation-sharded sweep is the same
// code as in the process-per-GPU form, each rank driven by its own host thread inside bvb_gibbs_run_multi.
int bvb_create_multi(int n, const int* devices, uint64_t seed, bvb_handle** handles) {
  if (n < 1 || !devices || !handles) return BVB_ERR_ARG;
  for (int i = 0; i < n; ++i) handles[i] = nullptr;
  char id[128];
  if (n > 1) { const int rc = bvb_comm_unique_id(id); if (rc) return rc; }
  for (int i = 0; i < n; ++i) {
    const int rc = bvb_create(&handles[i], devices[i], seed);
    if (rc) { for (int j = 0; j < i; ++j) { bvb_destroy(handles[j]); handles[j] = nullptr; } return rc; }
  }
  std::vector<int> rcs((size_t)n, BVB_OK);
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i) handles[i]->inproc = (n > 1) ? 1 : 0;
  for (int i = 0; i < n; ++i) th.emplace_back([&, i]() { rcs[i] = bvb_comm_init(handles[i], i, n, n > 1 ? id : nullptr); });   // (ncclCommInitRank blocks until every rank has joined)
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rcs[i]) { const int rc = rcs[i]; for (int j = 0; j < n; ++j) { bvb_destroy(handles[j]); handles[j] = nullptr; } return rc; }
  return BVB_OK;
}

int bvb_gibbs_init_multi(int n, bvb_handle** handles, const double* Y, const double* X, int M, int T, int K, int draws, int burnin,
                         int thin, int tvp_keep_all, int intercept, const double* priorIntercept, const double* PHI0,
                         const bvb_prior_phi* prior_phi, const bvb_prior_sigma* prior_sigma, const bvb_startvals* startvals,
                         const int* i_vec, double PHI_tol, double L_tol, int huge, const bvb_draws* out) {
  if (n < 1 || !handles) return BVB_ERR_ARG;
  for (int i = 0; i < n; ++i) {   // (no collective in the set-up: one rank after the other, rank 0 owns the output arrays)
    if (!handles[i]) return BVB_ERR_ARG;
    const int rc = bvb_gibbs_init(handles[i], Y, X, M, T, K, draws, burnin, thin, tvp_keep_all, intercept, priorIntercept, PHI0, prior_phi,
                                  prior_sigma, startvals, i_vec, PHI_tol, L_tol, huge, out);
    if (rc) return rc;
  }
  // peer exchange: the ranks live in this process, so the tables hold plain peer pointers (peer access enabled once)
  bool wire = n > 1;
  for (int i = 0; i < n && wire; ++i) {
    const Gibbs* Gi = static_cast<const Gibbs*>(handles[i]->gibbs);
    wire = Gi && Gi->p2p_wanted && Gi->world == n;
  }
  if (wire) {
    std::vector<double*> px((size_t)n);
    std::vector<unsigned long long*> pf((size_t)n);
    for (int j = 0; j < n; ++j) { Gibbs* Gj = static_cast<Gibbs*>(handles[j]->gibbs); px[j] = Gj->xbuf.p; pf[j] = Gj->xflag.p; }
    for (int i = 0; i < n && wire; ++i) {
      if (cudaSetDevice(handles[i]->device) != cudaSuccess) { wire = false; break; }
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const cudaError_t pe = cudaDeviceEnablePeerAccess(handles[j]->device, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (pe != cudaSuccess) { cudaGetLastError(); wire = false; break; }
      }
      Gibbs* Gi = static_cast<Gibbs*>(handles[i]->gibbs);
      if (wire && (cudaMemcpy(Gi->peer_xbuf.p, px.data(), sizeof(double*) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
                   cudaMemcpy(Gi->peer_xflag.p, pf.data(), sizeof(unsigned long long*) * n, cudaMemcpyHostToDevice) != cudaSuccess)) wire = false;
    }
    for (int i = 0; i < n; ++i) static_cast<Gibbs*>(handles[i]->gibbs)->p2p = wire;   // (all or none: otherwise the NCCL allgather)
  }
  return BVB_OK;
}

int bvb_gibbs_run_multi(int n, bvb_handle** handles, int max_sweeps, int* done) {
  if (n < 1 || !handles) return BVB_ERR_ARG;
  if (n == 1) return bvb_gibbs_run(handles[0], max_sweeps, done);
  std::vector<int> rcs((size_t)n, BVB_OK), dn((size_t)n, 0);
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i) th.emplace_back([&, i]() { rcs[i] = bvb_gibbs_run(handles[i], max_sweeps, &dn[i]); });
  for (auto& t : th) t.join();
  if (done) *done = dn[0];
  for (int i = 0; i < n; ++i) if (rcs[i]) return rcs[i];   // (the failing rank's message: bvb_last_error(handles[i]))
  return BVB_OK;
}

void bvb_destroy_multi(int n, bvb_handle** handles) {
  if (!handles) return;
  for (int i = 0; i < n; ++i) { if (handles[i]) bvb_destroy(handles[i]); handles[i] = nullptr; }
}

}  // extern "C"
