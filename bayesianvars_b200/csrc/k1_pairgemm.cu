// K1 - batched weighted SYRK as ONE tall-skinny FP64 tensor-core GEMM.
//
// Replaces, for all equations at once, the reference's per-equation
//   X_norm = X.each_col() % normalizer.col(j);          sample_coefficients.cpp:99
//   Omega  = X_norm.t()*X_norm; Omega.diag() += 1/v;    sample_coefficients.cpp:55-56
//   b      = X_norm.t()*y_norm + prior_mean/v;          sample_coefficients.cpp:61-64
//
// B200 design: every equation shares the regressor matrix X and differs only
// in the diagonal weight w_j[t] = exp(-h_tj), so
//   Omega_j[a,b] = sum_t (X[t,a] X[t,b]) * w_j[t]      ==   (Z W)[pair(a,b), j]
// with Z[pair, t] = X[t,a] X[t,b] built ONCE per chain (X never changes inside a
// Gibbs run).  The M symmetric rank-T updates become one dense GEMM with no
// wasted upper-triangle flops, no per-equation rescaled copy of X, and the
// weights as the (tiny, L2-resident) B operand.  The right-hand sides
//   b_j = X' (w_j . yhat_j)  are extra row tiles of the same GEMM (A = X', B = WY).
// The epilogue adds the prior terms (1/v on the diagonal, phi0/v on the rhs) and
// scatters into the packed augmented layout K2 factorises in place.
//
// Tiling: CTA = MW x 2 warps, 16*MW A-rows x (8*NT <= 104) equations, K chunks of
// 20 (3-stage cp.async ring, 160-byte rows -> conflict-free DMMA fragment loads:
// row stride 20 doubles == 4 mod 16).  Each warp owns 16 rows x half the columns
// as 2 x ceil(NT/2) m8n8k4 DMMA accumulators, which keeps it under 128 registers
// so 4 warps per scheduler are resident (the DMMA pipe needs that many to stay
// fed: with 2 warps per scheduler ncu showed 58% pipe activity).
#include "bvb_internal.cuh"
#include <cuda.h>      // CUtensorMap types only: the encoder is fetched with cudaGetDriverEntryPoint, nothing links to libcuda
#include <stdlib.h>

namespace bvb {

// DMMA work of one pipeline stage for a warp that owns NTC n8 tiles.  NTC is a
// template parameter (not a runtime predicate): ptxas if-converts `if (nt < ntc)`
// into predicated DMMAs, and a predicated-off DMMA still occupies the tensor pipe.
template <int NTC, int NTW>
__device__ __forceinline__ void k1_stage_mma(double (&acc)[2][NTW][2], const double* a, const double* b) {
#pragma unroll
  for (int ks = 0; ks < K1_BK / 4; ++ks) {
    const double a0 = a[4 * ks], a1 = a[8 * K1_BK + 4 * ks];
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      const double bv = b[nt * 8 * K1_BK + 4 * ks];
      dmma884(acc[0][nt][0], acc[0][nt][1], a0, bv);
      dmma884(acc[1][nt][0], acc[1][nt][1], a1, bv);
    }
  }
}

// NT = n8 tiles per CTA (split between two n-warps), MW = m-warps (16 rows each).
template <int NT, int MW>
__global__ void __launch_bounds__(MW * 64, (MW <= 3 ? 3 : 2)) k1_pairgemm(const K1Params p) {
  constexpr int BM = 16 * MW;
  constexpr int NTHREADS = MW * 64;
  constexpr int NB = 8 * NT;                       // equations handled by this CTA
  constexpr int NTW = (NT + 1) / 2;                // n8 tiles per warp (second n-warp may own one less)
  constexpr int A_STAGE = BM * K1_BK;              // doubles
  constexpr int B_STAGE = NB * K1_BK;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;                               // [STAGES][BM][20]
  double* sB = smem + K1_STAGES * A_STAGE;         // [STAGES][NB][20]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) bvb_stamp_begin(p.tstamp);
  const int wm = warp >> 1, wn = warp & 1;
  const int nt0 = wn * NTW;                        // first n-tile of this warp
  const int ntc = wn == 0 ? NTW : NT - NTW;        // number of n-tiles of this warp
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * NB;
  const bool ztile = tile < p.nZtiles;
  const double* __restrict__ Ag = p.A + (size_t)tile * BM * p.Tp;
  const double* __restrict__ Bg = (ztile ? p.W : p.WY) + (size_t)n0 * p.Tp;
  const int nk = p.Tp / K1_BK;

  auto load_stage = [&](int stage, int kc) {
    const double* a = Ag + kc * K1_BK;
    const double* b = Bg + kc * K1_BK;
    double* da = sA + stage * A_STAGE;
    double* db = sB + stage * B_STAGE;
    constexpr int PIECES = K1_BK / 2;              // 16-byte pieces per row
    for (int i = tid; i < BM * PIECES; i += NTHREADS) {
      int r = i / PIECES, c = i - r * PIECES;
      cp_async16(da + r * K1_BK + 2 * c, a + (size_t)r * p.Tp + 2 * c);
    }
    for (int i = tid; i < NB * PIECES; i += NTHREADS) {
      int r = i / PIECES, c = i - r * PIECES;
      cp_async16(db + r * K1_BK + 2 * c, b + (size_t)r * p.Tp + 2 * c);
    }
  };

  double acc[2][NTW][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;

#pragma unroll
  for (int s = 0; s < K1_STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  const int fr = lane >> 2, fk = lane & 3;         // fragment row / k index
  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<K1_STAGES - 2>();
    __syncthreads();
    {
      int nxt = kc + K1_STAGES - 1;
      if (nxt < nk) load_stage(nxt % K1_STAGES, nxt);
      cp_async_commit();
    }
    const double* a = sA + (kc % K1_STAGES) * A_STAGE + (wm * 16 + fr) * K1_BK + fk;
    const double* b = sB + (kc % K1_STAGES) * B_STAGE + (nt0 * 8 + fr) * K1_BK + fk;
    if (wn == 0) k1_stage_mma<NTW, NTW>(acc, a, b);
    else if (NT - NTW > 0) k1_stage_mma<(NT - NTW > 0 ? NT - NTW : 1), NTW>(acc, a, b);
  }
  cp_async_wait<0>();

  // Epilogue: add prior terms, scatter into the packed augmented systems.
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int row = tile * BM + wm * 16 + m * 8 + fr;
    const int2 l = p.lut[row];
    if (l.x < 0) continue;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt >= ntc) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = n0 + (nt0 + nt) * 8 + 2 * fk + e;
        if (j >= p.M) continue;
        double v = acc[m][nt][e];
        if (p.rowscale) {
          if (!ztile) continue;   // the right-hand side of the T x T system is written by huge_rhs_kernel
          const int2 ts = p.pairs[row];
          v = v * p.rowscale[(size_t)j * p.n + ts.x] * p.rowscale[(size_t)j * p.n + ts.y] + (ts.x == ts.y ? 1.0 : 0.0);
        } else if (l.y >= 0) {
          const double pv = p.V[(size_t)j * p.n + l.y];
          if (ztile) v += 1.0 / pv;
          else if (p.PHI0) v += p.PHI0[(size_t)j * p.n + l.y] / pv;
        }
        p.omega[(size_t)j * p.ldo + l.x] = v;
      }
    }
  }
  if (p.tstamp) {   // (uniform branch) last CTA end: after every warp's stores were issued
    __syncthreads();
    if (tid == 0) bvb_stamp_end(p.tstamp);
  }
}

// ---------------------------------------------------------------------------------------------
// TMA variant.  Same tiling and arithmetic; the operand tiles arrive as two 2-D tensor-map copies per
// stage (A: BM rows x 20 columns, W / WY: 8 NT rows x 20 columns, dense 160-byte rows = the layout the
// fragment loads already use) issued by one thread, completion on an mbarrier per stage.  Why: a
// cp.async.cg stream fills shared memory one 32-byte sector at a time and that contends with the LDS
// operand loads of the DMMA warps (tools/microbench/stream_vs_dmma.cu: at 42 B/clk/SM the DMMA rate
// halves, an aligned TMA stream costs it nothing); K1 streams ~14 B/clk/SM.  It also removes ~20
// address/issue instructions per thread and stage.
// ---------------------------------------------------------------------------------------------
struct K1Maps { CUtensorMap a, w, wy; };

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

template <int NT, int MW, int ST>
__global__ void __launch_bounds__(MW * 64, (MW <= 3 ? 3 : (MW <= 4 ? 2 : 1))) k1_pairgemm_tma(const K1Params p, const __grid_constant__ K1Maps maps) {
  constexpr int BM = 16 * MW;
  constexpr int NB = 8 * NT;
  constexpr int NTW = (NT + 1) / 2;
  constexpr int A_STAGE = BM * K1_BK;              // doubles (BM x 160 B: a multiple of 128 B)
  constexpr int B_STAGE = NB * K1_BK;
  constexpr int STAGE = A_STAGE + ((B_STAGE + 15) / 16) * 16;   // keep every stage 128-byte aligned
  extern __shared__ __align__(128) double smem_t[];
  __shared__ __align__(8) uint64_t full[ST];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) bvb_stamp_begin(p.tstamp);
  const int wm = warp >> 1, wn = warp & 1;
  const int nt0 = wn * NTW;
  const int ntc = wn == 0 ? NTW : NT - NTW;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * NB;
  const bool ztile = tile < p.nZtiles;
  const int nk = p.Tp / K1_BK;
  if (tid == 0) {
    for (int i = 0; i < ST; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto load_stage = [&](int stage, int kc) {   // thread 0 only
    double* da = smem_t + (size_t)stage * STAGE;
    mbar_arrive_expect_tx(&full[stage], (uint32_t)((A_STAGE + B_STAGE) * sizeof(double)));
    tma_load_2d(da, &maps.a, kc * K1_BK, tile * BM, &full[stage]);
    tma_load_2d(da + A_STAGE, ztile ? &maps.w : &maps.wy, kc * K1_BK, n0, &full[stage]);
  };

  double acc[2][NTW][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ST - 1; ++s)
      if (s < nk) load_stage(s, s);
  }
  const int fr = lane >> 2, fk = lane & 3;
  for (int kc = 0; kc < nk; ++kc) {
    const int st = kc % ST;
    __syncthreads();                               // every warp is done with the stage refilled below
    if (tid == 0) {
      const int nxt = kc + ST - 1;
      if (nxt < nk) { fence_proxy_async(); load_stage(nxt % ST, nxt); }
    }
    mbar_wait(&full[st], (uint32_t)((kc / ST) & 1));
    const double* a = smem_t + (size_t)st * STAGE + (wm * 16 + fr) * K1_BK + fk;
    const double* b = smem_t + (size_t)st * STAGE + A_STAGE + (nt0 * 8 + fr) * K1_BK + fk;
    if (wn == 0) k1_stage_mma<NTW, NTW>(acc, a, b);
    else if (NT - NTW > 0) k1_stage_mma<(NT - NTW > 0 ? NT - NTW : 1), NTW>(acc, a, b);
  }

  // Epilogue: add prior terms, scatter into the packed augmented systems.
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int row = tile * BM + wm * 16 + m * 8 + fr;
    const int2 l = p.lut[row];
    if (l.x < 0) continue;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt >= ntc) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = n0 + (nt0 + nt) * 8 + 2 * fk + e;
        if (j >= p.M) continue;
        double v = acc[m][nt][e];
        if (p.rowscale) {
          if (!ztile) continue;   // the right-hand side of the T x T system is written by huge_rhs_kernel
          const int2 ts = p.pairs[row];
          v = v * p.rowscale[(size_t)j * p.n + ts.x] * p.rowscale[(size_t)j * p.n + ts.y] + (ts.x == ts.y ? 1.0 : 0.0);
        } else if (l.y >= 0) {
          const double pv = p.V[(size_t)j * p.n + l.y];
          if (ztile) v += 1.0 / pv;
          else if (p.PHI0) v += p.PHI0[(size_t)j * p.n + l.y] / pv;
        }
        p.omega[(size_t)j * p.ldo + l.x] = v;
      }
    }
  }
  if (p.tstamp) {   // (uniform branch) last CTA end: after every warp's stores were issued
    __syncthreads();
    if (tid == 0) bvb_stamp_end(p.tstamp);
  }
}

// ---------------------------------------------------------------------------------------------
// Stream-K form of the TMA kernel.  1272 tiles over 296 resident CTAs are 4.3 waves: ncu shows the SM
// sub-partitions idle 11% of the kernel (the tail).  Here the work is the flat list of (tile, K-chunk)
// units, cut into one contiguous range per resident CTA, so every CTA gets the same number of chunks
// and the TMA pipeline runs straight through tile boundaries.  A tile cut by a range boundary is
// finished by the CTA that owns its FIRST chunks (it reaches the tile at the end of its range): the CTA
// owning the last chunks reaches them at the very start of its range, parks its accumulators in a
// workspace slot and raises a flag; the finisher adds them (fixed order: deterministic) and runs the
// epilogue.  The finisher only ever waits for a CTA with a higher index that started at the same time
// (all CTAs are co-resident, checked with the occupancy API), never for work that lies in the future.
// MEASURED: parity-green but 381 us against 317 us for the plain tiled TMA kernel at the headline shape - the
// resident CTAs of an SM now reach their tile ends (epilogue, pipeline drain) in lock step instead of staggered,
// and the kernel needs 126 registers.  Kept behind BVB_K1_STREAMK=1 as a recorded experiment.
// ---------------------------------------------------------------------------------------------
template <int NT, int MW, int ST>
__global__ void __launch_bounds__(MW * 64, 2) k1_pairgemm_sk(const K1Params p, const __grid_constant__ K1Maps maps,
                                                             double* __restrict__ ws, int* __restrict__ flags, int ntiles_rows) {
  constexpr int BM = 16 * MW;
  constexpr int NTHREADS = MW * 64;
  constexpr int NB = 8 * NT;
  constexpr int NTW = (NT + 1) / 2;
  constexpr int A_STAGE = BM * K1_BK;
  constexpr int B_STAGE = NB * K1_BK;
  constexpr int STAGE = A_STAGE + ((B_STAGE + 15) / 16) * 16;
  constexpr int NACC = 2 * NTW * 2;                  // accumulator doubles per thread
  extern __shared__ __align__(128) double smem_k[];
  __shared__ __align__(8) uint64_t full[ST];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int nt0 = wn * NTW;
  const int ntc = wn == 0 ? NTW : NT - NTW;
  const int nk = p.Tp / K1_BK;
  const int fr = lane >> 2, fk = lane & 3;
  // my range of units; unit u = (tile u / nk, chunk u % nk); tile = rowtile + ntiles_rows * ytile
  const long long U = (long long)ntiles_rows * gridDim.y * nk;   // (gridDim.y = 1: one equation tile)
  const long long u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < ST; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto load_unit = [&](int stage, long long u) {     // thread 0 only
    const int t = (int)(u / nk), kc = (int)(u - (long long)t * nk);
    const int rt = t % ntiles_rows, yt = t / ntiles_rows;
    double* da = smem_k + (size_t)stage * STAGE;
    mbar_arrive_expect_tx(&full[stage], (uint32_t)((A_STAGE + B_STAGE) * sizeof(double)));
    tma_load_2d(da, &maps.a, kc * K1_BK, rt * BM, &full[stage]);
    tma_load_2d(da + A_STAGE, rt < p.nZtiles ? &maps.w : &maps.wy, kc * K1_BK, yt * NB, &full[stage]);
  };

  double acc[2][NTW][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ST - 1; ++s)
      if (u0 + s < u1) load_unit(s, u0 + s);
  }
  int it = 0;                                        // units consumed by this CTA (stage / parity bookkeeping)
  for (long long u = u0; u < u1; ++u, ++it) {
    const int st = it % ST;
    __syncthreads();                                 // every warp is done with the stage refilled below
    if (tid == 0) {
      const long long nxt = u + ST - 1;
      if (nxt < u1) { fence_proxy_async(); load_unit((it + ST - 1) % ST, nxt); }
    }
    mbar_wait(&full[st], (uint32_t)((it / ST) & 1));
    const double* a = smem_k + (size_t)st * STAGE + (wm * 16 + fr) * K1_BK + fk;
    const double* b = smem_k + (size_t)st * STAGE + A_STAGE + (nt0 * 8 + fr) * K1_BK + fk;
    if (wn == 0) k1_stage_mma<NTW, NTW>(acc, a, b);
    else if (NT - NTW > 0) k1_stage_mma<(NT - NTW > 0 ? NT - NTW : 1), NTW>(acc, a, b);

    const int t = (int)(u / nk), kc = (int)(u - (long long)t * nk);
    const bool tile_done = (kc == nk - 1), range_done = (u + 1 == u1);
    if (!tile_done && !range_done) continue;
    if (tile_done && u - kc < u0) {
      // This tile was started by the previous CTA, which reaches it at the END of its range; I reach it at the START
      // of mine.  So I park my partial sums (the tile's last chunks) right away and the previous CTA finishes the
      // tile when it gets there - nobody waits for work that lies in the future.
      double* slot = ws + (size_t)blockIdx.x * (NACC * NTHREADS);
      int i = 0;
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e, ++i) { slot[(size_t)i * NTHREADS + tid] = acc[m][nt][e]; acc[m][nt][e] = 0.0; }
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(&flags[blockIdx.x], 1);
      continue;
    }
    if (!tile_done) {
      // my range ends inside this tile: the next CTA has (long ago) parked the rest of it; fixed order: mine + theirs
      if (tid == 0) {
        while (atomicAdd(&flags[blockIdx.x + 1], 0) == 0) __nanosleep(64);
      }
      __syncthreads();
      __threadfence();
      const double* slot = ws + (size_t)(blockIdx.x + 1) * (NACC * NTHREADS);
      int i = 0;
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e, ++i) acc[m][nt][e] += __ldcg(slot + (size_t)i * NTHREADS + tid);
      __syncthreads();
      if (tid == 0) atomicExch(&flags[blockIdx.x + 1], 0);   // consumed: the slot is free for the next launch
    }
    // epilogue of tile t: add prior terms, scatter into the packed augmented systems
    {
      const int rt = t % ntiles_rows, yt = t / ntiles_rows;
      const bool ztile = rt < p.nZtiles;
      const int n0 = yt * NB;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int row = rt * BM + wm * 16 + m * 8 + fr;
        const int2 l = p.lut[row];
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = n0 + (nt0 + nt) * 8 + 2 * fk + e;
            double v = acc[m][nt][e];
            acc[m][nt][e] = 0.0;
            if (l.x < 0 || nt >= ntc || j >= p.M) continue;
            if (p.rowscale) {
              if (!ztile) continue;   // the right-hand side of the T x T system is written by huge_rhs_kernel
              const int2 ts = p.pairs[row];
              v = v * p.rowscale[(size_t)j * p.n + ts.x] * p.rowscale[(size_t)j * p.n + ts.y] + (ts.x == ts.y ? 1.0 : 0.0);
            } else if (l.y >= 0) {
              const double pv = p.V[(size_t)j * p.n + l.y];
              if (ztile) v += 1.0 / pv;
              else if (p.PHI0) v += p.PHI0[(size_t)j * p.n + l.y] / pv;
            }
            p.omega[(size_t)j * p.ldo + l.x] = v;
          }
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// rows x Tp doubles, row-major (K-major), box = box_rows x K1_BK
static bool make_map(CUtensorMap* m, const double* base, int rows, int Tp, int box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)Tp, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)Tp * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)K1_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NT, int MW, int ST>
static cudaError_t launch_k1_tma_cfg(const K1Params& p, cudaStream_t s, bool* used) {
  constexpr int BM = 16 * MW, NB = 8 * NT;
  constexpr int STAGE = BM * K1_BK + ((NB * K1_BK + 15) / 16) * 16;
  *used = false;
  const int rows = p.ntiles * K1_BM;
  dim3 grid(rows / BM, (p.M + NB - 1) / NB);
  K1Maps maps;
  const int wrows = (int)grid.y * NB;
  if (!make_map(&maps.a, p.A, rows, p.Tp, BM) || !make_map(&maps.w, p.W, wrows, p.Tp, NB) ||
      !make_map(&maps.wy, p.WY, wrows, p.Tp, NB))
    return cudaSuccess;                              // no encoder / unsupported shape: the caller falls back to cp.async
  const size_t smem = (size_t)ST * STAGE * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(k1_pairgemm_tma<NT, MW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  K1Params q = p;
  q.nZtiles = p.nZtiles * K1_BM / BM;
  k1_pairgemm_tma<NT, MW, ST><<<grid, MW * 64, smem, s>>>(q, maps);
  *used = true;
  return cudaGetLastError();
}

template <int NT, int MW>
static cudaError_t launch_k1_cfg(const K1Params& p, cudaStream_t s) {
  const size_t smem = (size_t)K1_STAGES * (16 * MW + 8 * NT) * K1_BK * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(k1_pairgemm<NT, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int rows = p.ntiles * K1_BM;               // geometry is laid out in 64-row tiles
  dim3 grid(rows / (16 * MW), (p.M + 8 * NT - 1) / (8 * NT));
  K1Params q = p;
  q.nZtiles = p.nZtiles * K1_BM / (16 * MW);
  k1_pairgemm<NT, MW><<<grid, MW * 64, smem, s>>>(q);
  return cudaGetLastError();
}

static double* g_sk_ws = nullptr;
static int* g_sk_flags = nullptr;
static size_t g_sk_ws_bytes = 0;
static int g_sk_flags_n = 0;

template <int NT, int MW, int ST>
static cudaError_t launch_k1_sk_cfg(const K1Params& p, cudaStream_t s, bool* used) {
  constexpr int BM = 16 * MW, NB = 8 * NT;
  constexpr int STAGE = BM * K1_BK + ((NB * K1_BK + 15) / 16) * 16;
  constexpr int NACC = 2 * ((NT + 1) / 2) * 2;
  *used = false;
  const int rows = p.ntiles * K1_BM;
  const int ntr = rows / BM, ny = (p.M + NB - 1) / NB;
  const int nk = p.Tp / K1_BK;
  const size_t smem = (size_t)ST * STAGE * sizeof(double);
  static int num_sms = 0;
  if (num_sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
  cudaError_t e = cudaFuncSetAttribute(k1_pairgemm_sk<NT, MW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_pairgemm_sk<NT, MW, ST>, MW * 64, smem);
  if (e != cudaSuccess) return e;
  if (per_sm > 2) per_sm = 2;
  const int resident = per_sm * num_sms;
  const long long tiles = (long long)ntr * ny;
  // every CTA needs a range of at least one whole tile (then a tile is cut at most once)
  int G = (int)(tiles < resident ? tiles : resident);
  if (G < 1 || per_sm < 1 || (tiles * nk) / G < nk) return cudaSuccess;
  K1Maps maps;
  if (!make_map(&maps.a, p.A, rows, p.Tp, BM) || !make_map(&maps.w, p.W, ny * NB, p.Tp, NB) ||
      !make_map(&maps.wy, p.WY, ny * NB, p.Tp, NB))
    return cudaSuccess;
  const size_t need = (size_t)G * NACC * (MW * 64) * sizeof(double);
  if (need > g_sk_ws_bytes) {
    if (g_sk_ws) cudaFree(g_sk_ws);
    e = cudaMalloc(&g_sk_ws, need);
    if (e != cudaSuccess) { g_sk_ws = nullptr; g_sk_ws_bytes = 0; return e; }
    g_sk_ws_bytes = need;
  }
  if (G > g_sk_flags_n) {
    if (g_sk_flags) cudaFree(g_sk_flags);
    e = cudaMalloc(&g_sk_flags, sizeof(int) * 2048);
    if (e != cudaSuccess) { g_sk_flags = nullptr; g_sk_flags_n = 0; return e; }
    e = cudaMemsetAsync(g_sk_flags, 0, sizeof(int) * 2048, s);
    if (e != cudaSuccess) return e;
    g_sk_flags_n = 2048;
  }
  K1Params q = p;
  q.nZtiles = p.nZtiles * K1_BM / BM;
  if (ny != 1) return cudaSuccess;                   // one equation tile (M <= 104) only; wider systems fall back
  k1_pairgemm_sk<NT, MW, ST><<<dim3(G, 1), MW * 64, smem, s>>>(q, maps, g_sk_ws, g_sk_flags, ntr);
  *used = true;
  return cudaGetLastError();
}

template <int NT>
static cudaError_t launch_k1_nt(const K1Params& p, cudaStream_t s) {
  static int mw = 0;
  if (mw == 0) {
    const char* e = getenv("BVB_K1_MW");
    mw = e ? atoi(e) : 2;   // measured at the headline shape: 32-row tiles 328 us, 48-row 377 us, 64-row 335 us
    if (mw != 2 && mw != 3 && mw != 4) mw = 2;
  }
  static int tma = -1, st = 0, mwt = 0;
  if (tma < 0) {
    const char* e = getenv("BVB_K1_TMA"); tma = e ? atoi(e) : 1;
    e = getenv("BVB_K1_ST"); st = e ? atoi(e) : 3;
    e = getenv("BVB_K1_MW"); mwt = e ? atoi(e) : 4;      // TMA kernel: 64-row tiles by default
  }
  static int sk = -1;
  if (sk < 0) { const char* e = getenv("BVB_K1_STREAMK"); sk = e ? atoi(e) : 0; }   // off: measured 381 us vs 317 us
  if (tma && sk && mwt == 4) {
    bool used = false;
    cudaError_t e = (st == 4) ? launch_k1_sk_cfg<NT, 4, 4>(p, s, &used) : launch_k1_sk_cfg<NT, 4, 3>(p, s, &used);
    if (e != cudaSuccess || used) return e;
  }
  if (tma) {
    bool used = false;
    cudaError_t e;
    if (mwt == 2) e = st == 4 ? launch_k1_tma_cfg<NT, 2, 4>(p, s, &used) : launch_k1_tma_cfg<NT, 2, 3>(p, s, &used);
    else if (mwt == 8) e = st == 4 ? launch_k1_tma_cfg<NT, 8, 4>(p, s, &used) : launch_k1_tma_cfg<NT, 8, 3>(p, s, &used);
    else e = st == 4 ? launch_k1_tma_cfg<NT, 4, 4>(p, s, &used) : launch_k1_tma_cfg<NT, 4, 3>(p, s, &used);
    if (e != cudaSuccess || used) return e;
  }
  return mw == 2 ? launch_k1_cfg<NT, 2>(p, s) : mw == 3 ? launch_k1_cfg<NT, 3>(p, s) : launch_k1_cfg<NT, 4>(p, s);
}

cudaError_t launch_k1(const K1Params& p, cudaStream_t s) {
  const int m = p.M;
  // Smallest column block that covers M in the fewest CTA columns.
  int blocks = (m + K1_NMAX - 1) / K1_NMAX;
  {
    // BVB_K1_NSPLIT=c: at least c CTA columns (narrower equation blocks, more and shorter CTAs: a smaller last wave at
    // the price of re-reading A once per column)
    static int nsplit = -1;
    if (nsplit < 0) { const char* e = getenv("BVB_K1_NSPLIT"); nsplit = e ? atoi(e) : 0; }
    if (nsplit > blocks && nsplit <= (m + 7) / 8) blocks = nsplit;
  }
  const int per = (m + blocks - 1) / blocks;       // equations per CTA column
  const int nt = (per + 7) / 8;
  if (nt <= 1) return launch_k1_nt<1>(p, s);
  if (nt <= 2) return launch_k1_nt<2>(p, s);
  if (nt <= 4) return launch_k1_nt<4>(p, s);
  if (nt <= 7) return launch_k1_nt<7>(p, s);
  if (nt <= 10) return launch_k1_nt<10>(p, s);
  return launch_k1_nt<13>(p, s);
}

}  // namespace bvb
