// K6 - impulse-response recursion over stored posterior draws.
//
// Replaces irf_cpp / shift_and_insert (src/irf.cpp:457-524): per draw the impact
// response (Lambda diag(sigma) shocks for the factor model, U^-T diag(sqrt d) shocks
// for the cholesky model) followed by `ahead` products predictors x PHI with a lag
// shift.  Draws are independent, and so are the shocks of one draw (each shock is
// one row of the predictor matrix), so the grid is (draw, chunk of <= 8 shocks):
// the 8 x Kp predictor rows live in shared memory, one warp owns one output
// variable at a time and streams the matching PHI column (coalesced) once per
// horizon; consecutive CTAs work on the same draw, so its PHI slice is served
// from L2 after the first touch.  HBM-bound for factor models (few shocks).
// Quirks Q1/Q2 of SURVEY.md are reproduced on purpose.
#include "bvb_internal.cuh"

namespace bvb {

constexpr int IRF_SC_MAX = 8;   // shocks per CTA: 1, 2, 4 or 8 (template parameter: a factor model has r shocks)
constexpr int IRF_THREADS = 256;

struct IrfParams {
  const double* coef;     // Kp x M x R
  const double* facload;  // M x nf x R
  const double* Uvecs;    // nU x R
  const double* logvar;   // ld x R
  const double* shocks;   // dim x n_shocks
  const double* rotation; // dim x dim x R
  double* out;            // M x n_shocks x (ahead+1) x R
  long long* prof;        // BVB_IRF_PROF=1: cycles of CTA 0 per phase (load issue, impact, load wait, horizons, flush)
  int Kp, M, R, nf, hasU, ld, dim, n_shocks, ahead;
};

template <int IRF_SC>
__global__ void __launch_bounds__(IRF_THREADS) irf_kernel(const IrfParams p) {
  extern __shared__ double sm[];
  const int M = p.M, Kp = p.Kp;
  double* pred = sm;                       // [2][SC][Kp]
  double* rs = pred + 2 * IRF_SC * Kp;     // [dim][SC] rotated, scaled shocks
  double* ir0 = rs + p.dim * IRF_SC;       // [M][SC]
  const int r = blockIdx.x / ((p.n_shocks + IRF_SC - 1) / IRF_SC);
  const int s0 = (blockIdx.x % ((p.n_shocks + IRF_SC - 1) / IRF_SC)) * IRF_SC;
  const int ns = min(IRF_SC, p.n_shocks - s0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = IRF_THREADS / 32;
  const double* PHI = p.coef + (size_t)r * Kp * M;
  // rotated_shocks = rotation[r] * shocks | shocks, rows scaled by sqrt of the variances (irf.cpp:493-503)
  for (int i = tid; i < p.dim * IRF_SC; i += IRF_THREADS) {
    const int row = i / IRF_SC, s = i % IRF_SC;
    double v = 0.0;
    if (s < ns) {
      if (p.rotation) {
        const double* rot = p.rotation + (size_t)r * p.dim * p.dim;
        for (int k = 0; k < p.dim; ++k) v = fma(rot[(size_t)k * p.dim + row], p.shocks[(size_t)(s0 + s) * p.dim + k], v);
      } else {
        v = p.shocks[(size_t)(s0 + s) * p.dim + row];
      }
      if (p.nf > 0 || p.hasU) v *= exp(0.5 * p.logvar[(size_t)r * p.ld + row]);   // Q1: head(n_factors) of logvar_t
    }
    rs[i] = v;
  }
  __syncthreads();
  // impact responses
  if (p.nf > 0) {
    const double* L = p.facload + (size_t)r * M * p.nf;
    for (int i = tid; i < M * IRF_SC; i += IRF_THREADS) {
      const int m = i / IRF_SC, s = i % IRF_SC;
      double v = 0.0;
      for (int k = 0; k < p.nf; ++k) v = fma(L[(size_t)k * M + m], rs[k * IRF_SC + s], v);
      ir0[i] = v;
    }
  } else if (p.hasU) {
    // solve(trimatl(U'), rs): forward substitution, one thread per shock (unit diagonal)
    const double* uv = p.Uvecs + (size_t)r * ((size_t)M * (M - 1) / 2);
    if (tid < IRF_SC) {
      const int s = tid;
      for (int i = 0; i < M; ++i) {
        double v = rs[i * IRF_SC + s];
        // (U')[i, k] = U[k, i], k < i, stored at i(i-1)/2 + k
        for (int k = 0; k < i; ++k) v = fma(-uv[(size_t)i * (i - 1) / 2 + k], ir0[k * IRF_SC + s], v);
        ir0[i * IRF_SC + s] = v;
      }
    }
  } else {
    for (int i = tid; i < M * IRF_SC; i += IRF_THREADS) ir0[i] = rs[i];
  }
  __syncthreads();
  const size_t ostride_t = (size_t)M * p.n_shocks;
  double* out = p.out + (size_t)r * ostride_t * (p.ahead + 1);
  for (int i = tid; i < M * ns; i += IRF_THREADS) {
    const int m = i % M, s = i / M;
    out[(size_t)(s0 + s) * M + m] = ir0[m * IRF_SC + s];
  }
  // current_predictors = [ir0', 0 ...]
  for (int i = tid; i < IRF_SC * Kp; i += IRF_THREADS) {
    const int s = i / Kp, k = i % Kp;
    pred[i] = (k < M) ? ir0[k * IRF_SC + s] : 0.0;
  }
  __syncthreads();
  int cur = 0;
  for (int t = 1; t <= p.ahead; ++t) {
    const double* pc = pred + cur * IRF_SC * Kp;
    double* pn = pred + (cur ^ 1) * IRF_SC * Kp;
    // new[s][m] = sum_k pred[s][k] PHI[k, m]: one warp per variable m, lanes over k
    for (int m = warp; m < M; m += nw) {
      const double* col = PHI + (size_t)m * Kp;
      double acc[IRF_SC];
#pragma unroll
      for (int s = 0; s < IRF_SC; ++s) acc[s] = 0.0;
      // four independent loads in flight per lane: the PHI stream (L2 after the first horizon) is latency-bound
      int k = lane;
      for (; k + 96 < Kp; k += 128) {
        const double c0 = col[k], c1 = col[k + 32], c2 = col[k + 64], c3 = col[k + 96];
#pragma unroll
        for (int s = 0; s < IRF_SC; ++s) {
          const double* ps = pc + s * Kp + k;
          acc[s] = fma(ps[0], c0, acc[s]);
          acc[s] = fma(ps[32], c1, acc[s]);
          acc[s] = fma(ps[64], c2, acc[s]);
          acc[s] = fma(ps[96], c3, acc[s]);
        }
      }
      for (; k < Kp; k += 32) {
        const double c = col[k];
#pragma unroll
        for (int s = 0; s < IRF_SC; ++s) acc[s] = fma(pc[s * Kp + k], c, acc[s]);
      }
#pragma unroll
      for (int s = 0; s < IRF_SC; ++s) acc[s] = warp_sum(acc[s]);
      if (lane == 0) {
#pragma unroll
        for (int s = 0; s < IRF_SC; ++s) {
          pn[s * Kp + m] = acc[s];
          if (s < ns) out[(size_t)t * ostride_t + (size_t)(s0 + s) * M + m] = acc[s];
        }
      }
    }
    // shift_and_insert (irf.cpp:457-465): columns M..Kp-2 take the values M to their left (Q2)
    for (int i = tid; i < IRF_SC * Kp; i += IRF_THREADS) {
      const int s = i / Kp, k = i % Kp;
      if (k >= M) pn[i] = (k <= Kp - 2) ? pc[s * Kp + k - M] : pc[i];
    }
    __syncthreads();
    cur ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// K6 cluster form: the PHI slice of a draw stays ON CHIP for all horizons.
//
// ncu on the one-CTA-per-draw kernel above (profiles/ncu_summary_r1e_irf.json): with hundreds of
// draws in flight the PHI slices (1.28 MB each at M = 200, p = 4) thrash the 126 MB L2, every
// horizon re-reads its slice from DRAM (12x the algorithmic bytes, L2 hit rate 9 %).  Here a
// thread-block cluster of C CTAs (1, 2, 4 or 8) owns one draw at a time: CTA c keeps the columns
// [c nc, (c+1) nc) of PHI in its shared memory (read from HBM exactly once), computes those
// entries of the next response, and pushes them into the predictor rows of EVERY CTA of the
// cluster through distributed shared memory; one cluster barrier per horizon.  Clusters are
// persistent and stride over the draws; the next draw's columns are prefetched into L2 while the
// horizons of the current one run.  Same per-column summation order as the kernel above.
// ---------------------------------------------------------------------------------------------
// Block size is chosen at launch: one warp per resident column when they fit (25 warps for the 25 columns a CTA of an
// 8-CTA cluster holds at M = 200), so a horizon is ONE pass over the shared-memory slice - its floor is the 128 B/clk
// shared-memory read of the slice (160 KB: 0.65 us).  MAXT bounds the registers: 1024 threads for <= 2 shocks.

constexpr int IRFC_OSTAGE = 1536;   // doubles of output staging per CTA (at least one horizon is always staged)

__device__ __forceinline__ void irf_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64(double* local_ptr, unsigned rank, double v) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;\n" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gptr), "r"(bytes) : "memory");
}

template <int SC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) irf_cluster_kernel(const IrfParams p, int C, int nc) {
  const int IRFC_THREADS = blockDim.x;
  extern __shared__ __align__(16) double sm[];
  const int M = p.M, Kp = p.Kp;
  double* phis = sm;                               // [nc * Kp + 2]  this CTA's PHI columns (+1 slot of alignment slack)
  double* pred = phis + (((size_t)nc * Kp + 3) & ~(size_t)1);   // [2][SC][Kp]
  double* rs = pred + 2 * SC * Kp;                 // [dim][SC]
  double* ir0 = rs + p.dim * SC;                   // [M][SC]
  double* ostage = ir0 + M * SC;                   // [hcap][SC][nc] responses waiting to be written out
  constexpr int CPW = SC <= 2 ? 2 : 1;             // columns per warp and pass
  constexpr int NACC = SC <= 2 ? 4 : (SC == 4 ? 2 : 1);
  const int hcap = max(1, IRFC_OSTAGE / (SC * nc));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = IRFC_THREADS / 32;
  const int crank = blockIdx.x % C, cid = blockIdx.x / C, nclusters = gridDim.x / C;
  const int m0 = crank * nc;
  const int mycols = max(0, min(nc, M - m0));
  const int L = mycols * Kp;                       // doubles in this CTA's contiguous column block
  const int nchunks = (p.n_shocks + SC - 1) / SC;
  const size_t ostride_t = (size_t)M * p.n_shocks;

  const bool prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
  long long tprev = prof ? clock64() : 0;
  long long tacc[6] = {0, 0, 0, 0, 0, 0};          // kept in registers: a global += per tick would stall the warp on its load
#define IRF_TICK(slot) do { if (prof) { const long long tn = clock64(); tacc[slot] += tn - tprev; tprev = tn; } } while (0)
  for (int r = cid; r < p.R; r += nclusters) {
    const double* PHI = p.coef + (size_t)r * Kp * M;
    const double* src = PHI + (size_t)m0 * Kp;
    // the column block is contiguous in global memory: copy it flat, shifted by one slot when it is only 8-byte aligned
    const int a = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
    double* colbase = phis + a;
    if (L > 0) {
      if (a && tid == 0) cp_async8(colbase, src);
      const int i0 = a;                            // first 16-byte aligned element
      for (int i = i0 + 2 * tid; i + 1 < L; i += 2 * IRFC_THREADS) cp_async16(colbase + i, src + i);
      if (((L - i0) & 1) && tid == 32) cp_async8(colbase + L - 1, src + L - 1);
    }
    cp_async_commit();
    IRF_TICK(0);
    double* out = p.out + (size_t)r * ostride_t * (p.ahead + 1);

    for (int ch = 0; ch < nchunks; ++ch) {
      const int s0 = ch * SC;
      const int ns = min(SC, p.n_shocks - s0);
      // rotated_shocks = rotation[r] * shocks | shocks, rows scaled by sqrt of the variances (irf.cpp:493-503)
      for (int i = tid; i < p.dim * SC; i += IRFC_THREADS) {
        const int row = i / SC, s = i % SC;
        double v = 0.0;
        if (s < ns) {
          if (p.rotation) {
            const double* rot = p.rotation + (size_t)r * p.dim * p.dim;
            for (int k = 0; k < p.dim; ++k) v = fma(rot[(size_t)k * p.dim + row], p.shocks[(size_t)(s0 + s) * p.dim + k], v);
          } else {
            v = p.shocks[(size_t)(s0 + s) * p.dim + row];
          }
          if (p.nf > 0) v *= exp(0.5 * p.logvar[(size_t)r * p.ld + row]);   // Q1: head(n_factors) of logvar_t
        }
        rs[i] = v;
      }
      __syncthreads();
      if (p.nf > 0) {
        const double* Lm = p.facload + (size_t)r * M * p.nf;
        for (int i = tid; i < M * SC; i += IRFC_THREADS) {
          const int m = i / SC, s = i % SC;
          double v = 0.0;
          for (int k = 0; k < p.nf; ++k) v = fma(Lm[(size_t)k * M + m], rs[k * SC + s], v);
          ir0[i] = v;
        }
      } else {
        for (int i = tid; i < M * SC; i += IRFC_THREADS) ir0[i] = rs[i];
      }
      __syncthreads();
      if (crank == 0)
        for (int i = tid; i < M * ns; i += IRFC_THREADS) {
          const int m = i % M, s = i / M;
          out[(size_t)(s0 + s) * M + m] = ir0[m * SC + s];
        }
      for (int i = tid; i < SC * Kp; i += IRFC_THREADS) {
        const int s = i / Kp, k = i % Kp;
        pred[i] = (k < M) ? ir0[k * SC + s] : 0.0;
      }
      IRF_TICK(1);
      if (ch == 0) {
        cp_async_wait<0>();
        // next draw of this cluster -> L2 while the horizons run (16-byte granules; the edges are fetched by the copy)
        const int rn = r + nclusters;
        if (rn < p.R && warp == 0 && L > 8) {
          const char* nb = reinterpret_cast<const char*>(p.coef + (size_t)rn * Kp * M + (size_t)m0 * Kp);
          const char* lo = reinterpret_cast<const char*>((reinterpret_cast<uintptr_t>(nb) + 15) & ~(uintptr_t)15);
          const size_t tot = ((size_t)L * 8 - (size_t)(lo - nb)) & ~(size_t)15;
          const size_t per = ((tot / 32) + 15) & ~(size_t)15;
          const size_t b = (size_t)lane * per;
          if (b < tot) prefetch_l2_bulk(lo + b, (unsigned)min(per, tot - b));
        }
      }
      __syncthreads();
      IRF_TICK(2);
      int cur = 0;
      for (int t = 1; t <= p.ahead; ++t) {
        const double* pc = pred + cur * SC * Kp;
        double* pn = pred + (cur ^ 1) * SC * Kp;
        // new[s][m] = sum_k pred[s][k] PHI[k, m] for this CTA's columns: one warp per CPW columns (the predictor row is
        // read once for both), lanes over k.  The responses are staged in shared memory and written to global memory
        // every `hcap` horizons: a global store in flight makes the release of the cluster barrier wait for its ack.
        for (int j = warp; j < mycols; j += CPW * nw) {
          // NACC independent partial sums per (column, shock): the k loop is a chain of dependent DFMAs otherwise
          double acc4[NACC][CPW][SC];
#pragma unroll
          for (int u = 0; u < NACC; ++u)
#pragma unroll
            for (int c = 0; c < CPW; ++c)
#pragma unroll
              for (int s = 0; s < SC; ++s) acc4[u][c][s] = 0.0;
          const double* col = colbase + (size_t)j * Kp;
          const bool two = (CPW > 1) && (j + nw < mycols);
          const double* colb = two ? col + (size_t)nw * Kp : col;
          int k = lane;
          for (; k + 32 * (NACC - 1) < Kp; k += 32 * NACC) {
#pragma unroll
            for (int u = 0; u < NACC; ++u) {
              double cv[CPW];
              cv[0] = col[k + 32 * u];
              if (CPW > 1) cv[CPW - 1] = colb[k + 32 * u];
#pragma unroll
              for (int s = 0; s < SC; ++s) {
                const double pv = pc[s * Kp + k + 32 * u];
#pragma unroll
                for (int c = 0; c < CPW; ++c) acc4[u][c][s] = fma(pv, cv[c], acc4[u][c][s]);
              }
            }
          }
          for (; k < Kp; k += 32) {
            double cv[CPW];
            cv[0] = col[k];
            if (CPW > 1) cv[CPW - 1] = colb[k];
#pragma unroll
            for (int s = 0; s < SC; ++s) {
              const double pv = pc[s * Kp + k];
#pragma unroll
              for (int c = 0; c < CPW; ++c) acc4[0][c][s] = fma(pv, cv[c], acc4[0][c][s]);
            }
          }
          double acc[CPW][SC];
#pragma unroll
          for (int c = 0; c < CPW; ++c)
#pragma unroll
            for (int s = 0; s < SC; ++s) {
              double v = acc4[0][c][s];
#pragma unroll
              for (int u = 1; u < NACC; ++u) v += acc4[u][c][s];
              acc[c][s] = v;
            }
#pragma unroll
          for (int c = 0; c < CPW; ++c)
#pragma unroll
            for (int s = 0; s < SC; ++s) acc[c][s] = warp_sum(acc[c][s]);
#pragma unroll
          for (int c = 0; c < CPW; ++c) {
            if (c > 0 && !two) break;
            const int jj = j + c * nw;
            if (lane < C) {
#pragma unroll
              for (int s = 0; s < SC; ++s) st_cluster_f64(pn + s * Kp + m0 + jj, (unsigned)lane, acc[c][s]);
            }
            if (lane == 0) {
#pragma unroll
              for (int s = 0; s < SC; ++s) ostage[((size_t)((t - 1) % hcap) * SC + s) * nc + jj] = acc[c][s];
            }
          }
        }
        // shift_and_insert (irf.cpp:457-465): columns M..Kp-2 take the values M to their left (Q2); local rows only
        for (int i = tid; i < SC * Kp; i += IRFC_THREADS) {
          const int s = i / Kp, k = i % Kp;
          if (k >= M) pn[i] = (k <= Kp - 2) ? pc[s * Kp + k - M] : pc[i];
        }
        IRF_TICK(3);
        irf_cluster_sync();                          // every CTA's pushes have landed; pc is dead everywhere
        IRF_TICK(4);
        cur ^= 1;
        if (t % hcap == 0 || t == p.ahead) {         // flush the staged responses of horizons (t0, t]
          const int t0 = ((t - 1) / hcap) * hcap;
          const int cnt = (t - t0) * SC * mycols;
          for (int i = tid; i < cnt; i += IRFC_THREADS) {
            const int jj = i % mycols, sq = (i / mycols) % SC, tl = i / (mycols * SC);
            if (sq < ns) out[(size_t)(t0 + 1 + tl) * ostride_t + (size_t)(s0 + sq) * M + m0 + jj] = ostage[((size_t)tl * SC + sq) * nc + jj];
          }
          __syncthreads();
          IRF_TICK(5);
        }
      }
      __syncthreads();
    }
    if (nchunks == 0) cp_async_wait<0>();
    __syncthreads();
  }
  if (prof)
    for (int i = 0; i < 6; ++i) p.prof[i] = tacc[i];
#undef IRF_TICK
}

template <int SC>
static cudaError_t launch_irf_sc(const IrfParams& p, cudaStream_t s) {
  const size_t smem = ((size_t)2 * SC * p.Kp + (size_t)p.dim * SC + (size_t)p.M * SC) * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(irf_kernel<SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int chunks = (p.n_shocks + SC - 1) / SC;
  irf_kernel<SC><<<p.R * chunks, IRF_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

template <int SC>
static cudaError_t launch_irf_cluster_sc(const IrfParams& p, int C, int nc, size_t smem, cudaStream_t s) {
  constexpr int MAXT = 512;
  auto kern = irf_cluster_kernel<SC, MAXT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int maxw = MAXT / 32;
  const int items = SC <= 2 ? (nc + 1) / 2 : nc;       // (pairs of) columns, see CPW in the kernel
  const int rounds = (items + maxw - 1) / maxw;
  int warps = (items + rounds - 1) / rounds;
  if (warps < 8) warps = 8;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3((unsigned)(32 * warps));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)C);
  int maxc = 0;
  e = cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg);
  if (e != cudaSuccess) return e;
  if (maxc < 1) return cudaErrorInvalidConfiguration;
  const int ncl = p.R < maxc ? p.R : maxc;
  cfg.gridDim = dim3((unsigned)(ncl * C));
  return cudaLaunchKernelEx(&cfg, kern, p, C, nc);
}

static size_t irf_cluster_smem(const IrfParams& p, int SC, int nc) {
  const size_t stage = (size_t)SC * nc > (size_t)IRFC_OSTAGE ? (size_t)SC * nc : (size_t)IRFC_OSTAGE;
  return ((((size_t)nc * p.Kp + 3) & ~(size_t)1) + (size_t)2 * SC * p.Kp + (size_t)p.dim * SC + (size_t)p.M * SC + stage) * sizeof(double);
}

template <int SC>
static size_t irf_stream_footprint(const IrfParams& p) {
  // bytes of PHI the one-CTA-per-draw kernel keeps in flight: resident CTAs x slice (consecutive CTAs share a draw)
  const size_t smem = ((size_t)2 * SC * p.Kp + (size_t)p.dim * SC + (size_t)p.M * SC) * sizeof(double);
  if (smem > (size_t)227 * 1024) return ~(size_t)0;
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(irf_kernel<SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, irf_kernel<SC>, IRF_THREADS, smem);
  const int chunks = (p.n_shocks + SC - 1) / SC;
  size_t draws = ((size_t)sms * (size_t)(occ > 0 ? occ : 1) + chunks - 1) / chunks;
  if (draws > (size_t)p.R) draws = (size_t)p.R;
  return draws * (size_t)p.Kp * p.M * sizeof(double);
}

cudaError_t launch_irf(const IrfParams& p, cudaStream_t s) {
  int sc0 = 1;
  while (sc0 < IRF_SC_MAX && sc0 < p.n_shocks) sc0 <<= 1;
  // The streaming kernel is fine while the PHI slices of the draws in flight stay L2-resident (small models); beyond
  // that every horizon re-reads its slice from DRAM and the cluster form (slice resident in the shared memory of 1..8
  // CTAs) takes over whenever it fits.  The impact response of the cholesky model is a serial substitution per shock,
  // which the many-CTA kernel hides better.
  const char* fenv = getenv("BVB_IRF_CLUSTER");   // 0 / 1 force the streaming / cluster form (tests)
  const int forced = fenv ? atoi(fenv) : -1;
  const size_t fp = sc0 == 1 ? irf_stream_footprint<1>(p) : sc0 == 2 ? irf_stream_footprint<2>(p)
                  : sc0 == 4 ? irf_stream_footprint<4>(p) : irf_stream_footprint<8>(p);
  const bool want_cluster = forced >= 0 ? forced != 0 : fp > ((size_t)64 << 20);
  if (!p.hasU && want_cluster) {
    for (int SC = sc0; SC >= 1; SC >>= 1)
      for (int C = 1; C <= 8; C <<= 1) {
        const int nc = (p.M + C - 1) / C;
        const size_t smem = irf_cluster_smem(p, SC, nc);
        if (smem > (size_t)227 * 1024) continue;
        switch (SC) {
          case 1: return launch_irf_cluster_sc<1>(p, C, nc, smem, s);
          case 2: return launch_irf_cluster_sc<2>(p, C, nc, smem, s);
          case 4: return launch_irf_cluster_sc<4>(p, C, nc, smem, s);
          default: return launch_irf_cluster_sc<8>(p, C, nc, smem, s);
        }
      }
  }
  if (p.n_shocks <= 1) return launch_irf_sc<1>(p, s);
  if (p.n_shocks <= 2) return launch_irf_sc<2>(p, s);
  if (p.n_shocks <= 4) return launch_irf_sc<4>(p, s);
  return launch_irf_sc<IRF_SC_MAX>(p, s);
}

}  // namespace bvb

using namespace bvb;

extern "C" int bvb_irf(bvb_handle* h, int Kp, int M, int R, int n_factors, int has_U, int ld_logvar, int dim, int n_shocks,
                       int ahead, const double* coefficients, const double* factor_loadings, const double* U_vecs,
                       const double* logvar_t, const double* shocks, const double* rotation, double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 4; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (Kp <= 0 || M <= 0 || R < 0 || n_factors < 0 || dim <= 0 || n_shocks <= 0 || ahead < 0 || !coefficients || !shocks || !out ||
      (n_factors > 0 && !factor_loadings) || (has_U && !U_vecs) || ((n_factors > 0 || has_U) && !logvar_t) || Kp < M ||
      (n_factors > 0 ? dim != n_factors : dim != M)) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_irf: invalid argument");
    return BVB_ERR_ARG;
  }
  if (R == 0) return BVB_OK;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t d = sizeof(double);
  const size_t nU = (size_t)M * (M - 1) / 2;
  double *dc = nullptr, *dl = nullptr, *du = nullptr, *dv = nullptr, *ds = nullptr, *dr = nullptr, *dout = nullptr;
  auto fin = [&](int rc) {
    for (double* q : {dc, dl, du, dv, ds, dr, dout}) if (q) cudaFree(q);
    return rc;
  };
  const size_t out_n = (size_t)M * n_shocks * (ahead + 1) * R;
#define IRF_TRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { bvb_set_error(h, BVB_ERR_CUDA, "bvb_irf: %s", cudaGetErrorString(_e)); return fin(BVB_ERR_CUDA); } } while (0)
  IRF_TRY(cudaMalloc(&dc, d * Kp * M * R));
  IRF_TRY(cudaMalloc(&ds, d * dim * n_shocks));
  IRF_TRY(cudaMalloc(&dout, d * out_n));
  IRF_TRY(cudaMemcpyAsync(dc, coefficients, d * Kp * M * R, cudaMemcpyHostToDevice, s));
  IRF_TRY(cudaMemcpyAsync(ds, shocks, d * dim * n_shocks, cudaMemcpyHostToDevice, s));
  if (n_factors > 0) {
    IRF_TRY(cudaMalloc(&dl, d * M * n_factors * R));
    IRF_TRY(cudaMemcpyAsync(dl, factor_loadings, d * M * n_factors * R, cudaMemcpyHostToDevice, s));
  }
  if (has_U && nU > 0) {
    IRF_TRY(cudaMalloc(&du, d * nU * R));
    IRF_TRY(cudaMemcpyAsync(du, U_vecs, d * nU * R, cudaMemcpyHostToDevice, s));
  }
  if (logvar_t && ld_logvar > 0) {
    IRF_TRY(cudaMalloc(&dv, d * ld_logvar * R));
    IRF_TRY(cudaMemcpyAsync(dv, logvar_t, d * ld_logvar * R, cudaMemcpyHostToDevice, s));
  }
  if (rotation) {
    IRF_TRY(cudaMalloc(&dr, d * dim * dim * R));
    IRF_TRY(cudaMemcpyAsync(dr, rotation, d * dim * dim * R, cudaMemcpyHostToDevice, s));
  }
  IrfParams p{};
  p.coef = dc; p.facload = dl; p.Uvecs = du; p.logvar = dv; p.shocks = ds; p.rotation = dr; p.out = dout;
  p.Kp = Kp; p.M = M; p.R = R; p.nf = n_factors; p.hasU = (has_U && n_factors == 0) ? 1 : 0; p.ld = ld_logvar; p.dim = dim;
  p.n_shocks = n_shocks; p.ahead = ahead;
  long long* dprof = nullptr;
  if (getenv("BVB_IRF_PROF")) {
    if (cudaMalloc(&dprof, 8 * sizeof(long long)) == cudaSuccess) cudaMemsetAsync(dprof, 0, 8 * sizeof(long long), s);
    else dprof = nullptr;
  }
  p.prof = dprof;
  IRF_TRY(cudaEventRecord(h->ev[0], s));
  IRF_TRY(launch_irf(p, s));
  IRF_TRY(cudaEventRecord(h->ev[1], s));
  if (dprof) {
    long long hp[8];
    cudaMemcpyAsync(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(dprof);
    fprintf(stderr, "irf cluster ticks (CTA 0, all its draws): load-issue %lld impact %lld load-wait %lld compute %lld barrier %lld flush %lld\n",
            hp[0], hp[1], hp[2], hp[3], hp[4], hp[5]);
  }
  IRF_TRY(cudaMemcpyAsync(out, dout, d * out_n, cudaMemcpyDeviceToHost, s));
  IRF_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = 1;
#undef IRF_TRY
  return fin(BVB_OK);
}
