// K2 - batched Cholesky factorisation + forward solve + N(0,1)-perturbed mean
// draw, one CTA per equation.
//
// Replaces the reference's per-equation (sample_coefficients.cpp:58-70)
//   R = chol(Omega,"upper"); Rinv = inv(trimatu(R));
//   mean = Rinv*(Rinv.t()*b);  draw = mean + Rinv*z
// by the algebraically identical   draw = L^-T (L^-1 b + z),  L = R'.
//
// Layout: the packed augmented system written by K1 (lower triangle of Omega by
// columns, each column followed by its right-hand-side entry as row n), so the
// forward solve L^-1 b falls out of the factorisation as the last row of L.
//
// Algorithm (left-looking, block size 32, in place, the factor stays L2-resident):
//   for each 32-column block k:
//     acc    = L[rows, 0:32k] * L[block rows, 0:32k]^T       (DMMA m8n8k4, 16 warps,
//                                                             3-stage cp.async ring over L)
//     panel  = A[rows >= 32k, block k] - acc                 (smem, aliases the ring)
//     L_kk   = chol(panel diag 32x32), Linv = L_kk^-1        (one warp, registers,
//                                                             rsqrt pivots, smem broadcast)
//     panel[sub rows] = panel[sub rows] * Linv^T             (DMMA, triangular skip)
//     write the panel back; keep Linv for the back-solve
//   back-solve  L^T phi = (L^-1 b) + z  block by block from the bottom, using the
//   saved 32x32 inverses so no 32-step dependent chain is left in it.
#include "bvb_internal.cuh"

namespace bvb {

constexpr int K2_NB = 32;
constexpr int K2_BKC = 16;   // columns of L per pipeline stage (one per warp in the loader)
constexpr int K2_STAGES = 3;
constexpr int K2_LS = 36;    // row stride of the Linv^T tile (== 4 mod 16)
constexpr int K2_NWARPS = 16;
constexpr int K2_MAXQ = 4;   // m8 tiles per warp: 16 * 4 * 8 = 512 panel rows max

__host__ __device__ inline int k2_row_stride(int n) {
  int rmax = (n + 1 > 33 ? n + 1 : 33) + 8;
  return ((rmax + 15) / 16) * 16 + 4;
}
__host__ __device__ inline int k2_nblocks(int n) { return (n + K2_NB - 1) / K2_NB; }

// smem: [stage ring | panel (aliased)] [linvT] [ysol] [sblk] [colbuf 2x32] [rsbuf 32] [colbase ints]
size_t k2_smem_bytes(int n) {
  size_t rs = k2_row_stride(n);
  size_t ring = (size_t)K2_STAGES * K2_BKC * rs, panel = (size_t)K2_NB * rs;
  size_t d = (ring > panel ? ring : panel) + (size_t)K2_NB * K2_LS + (size_t)(k2_nblocks(n) * K2_NB + 32) +
             32 + 64 + 32 + (size_t)(n + 3) / 2 + 8;
  return d * sizeof(double);
}


// DMMA work of one pipeline stage for a warp that owns NQ m8 tiles (mt = warp +
// NWARPS*q).  NQ is a template parameter on purpose: a runtime `if (mt < mtiles)`
// is if-converted by ptxas into PREDICATED DMMAs, and a predicated-off DMMA
// still occupies the tensor pipe for its full 16 cycles (measured: the loop ran
// at MAXQ tiles' cost for every warp, 4x the useful work).
template <int NQ>
__device__ __forceinline__ void k2_stage_mma(double (&acc)[K2_MAXQ][4][2], const double* sb, int RS, int warp, int fr, int fk) {
#pragma unroll
  for (int ks = 0; ks < K2_BKC / 4; ++ks) {
    const double* col = sb + (4 * ks + fk) * RS + fr;
    double b[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) b[nt] = col[8 * nt];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double a = col[8 * (warp + K2_NWARPS * q)];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884(acc[q][nt][0], acc[q][nt][1], a, b[nt]);
    }
  }
}
template <>
__device__ __forceinline__ void k2_stage_mma<0>(double (&)[K2_MAXQ][4][2], const double*, int, int, int, int) {}

__global__ void __launch_bounds__(K2_NWARPS * 32, 1) k2_cholsolve(const K2Params p, double* __restrict__ linv_ws) {
  constexpr int NWARPS = K2_NWARPS, MAXQ = K2_MAXQ, NTHREADS = K2_NWARPS * 32;
  extern __shared__ __align__(16) double smem[];
  const int n = p.n, n1 = n + 1;
  const int RS = k2_row_stride(n);
  const int nbk = k2_nblocks(n);
  const size_t ring_d = (size_t)K2_STAGES * K2_BKC * RS, panel_d = (size_t)K2_NB * RS;
  double* stage = smem;                                   // [STAGES][BKC][RS]
  double* panel = smem;                                   // [32][RS]   panel[c][li]  (aliases the ring)
  double* linvT = smem + (ring_d > panel_d ? ring_d : panel_d);  // [32][36] linvT[c'][c] = Linv[c][c']
  double* ysol = linvT + K2_NB * K2_LS;                   // [nbk*32 + 32]
  double* sblk = ysol + nbk * K2_NB + 32;                 // [32]
  double* colbuf = sblk + 32;                             // [2][32]
  double* rsbuf = colbuf + 64;                            // [32]
  int* colb = reinterpret_cast<int*>(rsbuf + 32);         // [n]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int eq = blockIdx.x;
  double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  double* __restrict__ lws = linv_ws + (size_t)eq * nbk * (K2_NB * K2_NB);
  int fail = 0;  // meaningful on warp 0 only
  // optional phase profile (cycles, CTA 0 / thread 0): gemm, fill, factor, invert, trsm, writeback, backsolve
  long long tprev = 0;
  const bool prof = (p.prof != nullptr) && eq == 0 && tid == 0;
  if (prof) tprev = clock64();
#define K2_TICK(slot) do { if (prof) { const long long tn = clock64(); p.prof[slot] += tn - tprev; tprev = tn; } } while (0)

  for (int i = tid; i < n; i += NTHREADS) colb[i] = p.colbase[i];
  __syncthreads();

  for (int k = 0; k < nbk; ++k) {
    const int c0 = k * K2_NB;
    const int w = min(K2_NB, n - c0);
    const int kb_end = c0 + w;
    const int nsub = n1 - kb_end;
    const int R = K2_NB + nsub;
    const bool full = (w == K2_NB);
    const int mtiles = (R + 7) >> 3;
    // local row -> global row (or -1 for the virtual identity rows of a partial block)
    auto grow = [&](int li) -> int { return li < K2_NB ? (li < w ? c0 + li : -1) : kb_end + li - K2_NB; };

    // ---- B: left-looking update with the already-factorised columns -----------
    // chunk = 16 columns of L, one column per warp in the loader (coalesced 16-byte cp.async)
    const int nchunks = c0 / K2_BKC;
    auto load_chunk = [&](int st, int ch) {
      double* sdst = stage + (size_t)st * (K2_BKC * RS) + warp * RS;
      const double* src = om + colb[ch * K2_BKC + warp];
      if (full) {
        const int npair = (R + 1) >> 1;
        for (int pr = lane; pr < npair; pr += 32) cp_async16(sdst + 2 * pr, src + c0 + 2 * pr);
      } else {
        for (int li = lane; li < R; li += 32) {
          const int g = grow(li);
          if (g >= 0) cp_async8(sdst + li, src + g);
          else sdst[li] = 0.0;
        }
      }
    };

    const int nq = (mtiles > warp) ? (mtiles - warp + NWARPS - 1) / NWARPS : 0;
    double acc[MAXQ][4][2];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;

#pragma unroll
    for (int s = 0; s < K2_STAGES - 1; ++s) {
      if (s < nchunks) load_chunk(s, s);
      cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
      cp_async_wait<K2_STAGES - 2>();
      __syncthreads();
      {
        const int nxt = ch + K2_STAGES - 1;
        if (nxt < nchunks && p.debug_mode != 2) load_chunk(nxt % K2_STAGES, nxt);
        cp_async_commit();
      }
      const double* sb = stage + (size_t)(ch % K2_STAGES) * (K2_BKC * RS);
      if (p.debug_mode == 1) continue;
      switch (nq) {   // warp-uniform: number of m8 tiles this warp owns in this block column
        case 1: k2_stage_mma<1>(acc, sb, RS, warp, fr, fk); break;
        case 2: k2_stage_mma<2>(acc, sb, RS, warp, fr, fk); break;
        case 3: k2_stage_mma<3>(acc, sb, RS, warp, fr, fk); break;
        case 4: k2_stage_mma<4>(acc, sb, RS, warp, fr, fk); break;
        default: break;
      }
    }
    cp_async_wait<0>();
    __syncthreads();   // ring is dead from here on; its memory becomes the panel
    K2_TICK(0);

    // ---- A: panel = A_orig - acc, straight from the accumulator fragments -------
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int mt = warp + NWARPS * q;
      const int li = 8 * mt + fr;
      if (mt < mtiles && li < R) {
        const int g = grow(li);
        double orig[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * fk + e;
            double v;
            if (c < w) v = (g >= c0 + c) ? om[colb[c0 + c] + g] : 0.0;
            else v = (li == c) ? 1.0 : 0.0;
            orig[nt][e] = v;
          }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * fk + e;
            double v = orig[nt][e];
            if (c < w && g >= c0 + c) v -= acc[q][nt][e];
            panel[c * RS + li] = v;
          }
      }
    }
    __syncthreads();
    K2_TICK(1);

    // ---- C: factor the 32x32 diagonal block and invert it (warp 0) -------------
    // Lane i holds row i.  Right-looking, software-pipelined: as soon as column c
    // is scaled, column c+1 is brought up to date and the NEXT pivot's
    // shuffle + rsqrt chain is started before the remaining (independent) rank-1
    // updates are issued, so the 32-pivot dependent chain is ~(shfl + rsqrt + mul + fma).
    if (warp == 0) {
      double a[K2_NB];
#pragma unroll
      for (int c = 0; c < K2_NB; ++c) a[c] = (c <= lane) ? panel[c * RS + lane] : 0.0;
      double x = __shfl_sync(0xffffffffu, a[0], 0);
      if (!(x > 0.0) || !isfinite(x)) { if (fail == 0) fail = c0 + 1; x = 1.0; }
      double rs = rsqrt(x);
#pragma unroll
      for (int c = 0; c < K2_NB; ++c) {
        const double l = (lane >= c) ? a[c] * rs : 0.0;   // lane c: x * rsqrt(x) = sqrt(x)
        a[c] = l;
        double* cb = colbuf + (c & 1) * 32;
        cb[lane] = l;
        if (lane == c) rsbuf[c] = rs;
        __syncwarp();
        if (c + 1 < K2_NB) {
          a[c + 1] = fma(-l, cb[c + 1], a[c + 1]);
          x = __shfl_sync(0xffffffffu, a[c + 1], c + 1);
          if (!(x > 0.0) || !isfinite(x)) { if (fail == 0) fail = c0 + c + 2; x = 1.0; }
          rs = rsqrt(x);
        }
#pragma unroll
        for (int c2 = c + 2; c2 < K2_NB; ++c2) a[c2] = fma(-l, cb[c2], a[c2]);
      }
#pragma unroll
      for (int c = 0; c < K2_NB; ++c) panel[c * RS + lane] = a[c];
      __syncwarp();
      K2_TICK(2);
      // lane j solves L x = e_j -> column j of Linv.  Right-looking substitution: once
      // x[k] is known all later partial sums are updated (independent FMAs), so the
      // dependent chain is one fma + one mul per row.
      double sacc[K2_NB];
#pragma unroll
      for (int i = 0; i < K2_NB; ++i) sacc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int kk = 0; kk < K2_NB; ++kk) {
        const double xk = (kk >= lane) ? sacc[kk] * rsbuf[kk] : 0.0;
        sacc[kk] = xk;
#pragma unroll
        for (int i = kk + 1; i < K2_NB; ++i) sacc[i] = fma(-panel[kk * RS + i], xk, sacc[i]);
      }
#pragma unroll
      for (int i = 0; i < K2_NB; ++i) {
        linvT[lane * K2_LS + i] = sacc[i];                                        // linvT[c'=j][c=i] = Linv[i][j]
        lws[(size_t)k * (K2_NB * K2_NB) + i * K2_NB + lane] = sacc[i];            // ws[k][c'=i][c=j] = Linv[i][j]
      }
    }
    __syncthreads();
    K2_TICK(3);

    // ---- D: sub-diagonal rows  X = A * Linv^T  (DMMA, skipping the zero triangle)
    for (int mt = 4 + warp; mt < mtiles; mt += NWARPS) {
      double t[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) t[nt][0] = t[nt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < K2_NB / 4; ++ks) {
        const double av = panel[(4 * ks + fk) * RS + 8 * mt + fr];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (4 * ks <= 8 * nt + 7) {
            const double bv = linvT[(4 * ks + fk) * K2_LS + 8 * nt + fr];
            dmma884(t[nt][0], t[nt][1], av, bv);
          }
        }
      }
      __syncwarp();
      const int li = 8 * mt + fr;
      if (li < R) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) panel[(8 * nt + 2 * fk + e) * RS + li] = t[nt][e];
      }
    }
    __syncthreads();
    K2_TICK(4);

    // ---- E: write the factorised panel back; keep the forward-solved rhs --------
    for (int c = warp; c < w; c += NWARPS) {
      double* dst = om + colb[c0 + c];
      const double* src = panel + c * RS;
      for (int li = lane; li < R; li += 32) {
        const int g = grow(li);
        if (g >= c0 + c) dst[g] = src[li];
      }
    }
    if (tid < K2_NB) ysol[c0 + tid] = (tid < w) ? panel[tid * RS + (R - 1)] : 0.0;
    __syncthreads();
    K2_TICK(5);
  }

  if (p.mode == 1) {   // factor only: the caller solves later (triangular algorithm)
    if (tid == 0 && fail) atomicCAS(&p.status[eq], 0, fail);
    return;
  }
  // ---- back-solve  L^T phi = y + z ------------------------------------------------
  for (int i = tid; i < n; i += NTHREADS) ysol[i] += p.z[(size_t)eq * p.ldz + i];
  __syncthreads();
  for (int k = nbk - 1; k >= 0; --k) {
    const int c0 = k * K2_NB;
    const int w = min(K2_NB, n - c0);
    const int kb_end = c0 + w;
    for (int c = warp; c < K2_NB; c += NWARPS) {
      double s = 0.0;
      if (c < w) {
        const double* colp = om + colb[c0 + c];
        for (int i = kb_end + lane; i < n; i += 32) s = fma(colp[i], ysol[i], s);
        s = warp_sum(s);
      }
      if (lane == 0) sblk[c] = (c < w) ? ysol[c0 + c] - s : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      const double* wsk = lws + (size_t)k * (K2_NB * K2_NB);
      double xv0 = 0.0, xv1 = 0.0;
#pragma unroll
      for (int cp = 0; cp < K2_NB; cp += 2) {
        xv0 = fma(wsk[cp * K2_NB + lane], sblk[cp], xv0);
        xv1 = fma(wsk[(cp + 1) * K2_NB + lane], sblk[cp + 1], xv1);
      }
      if (lane < w) ysol[c0 + lane] = xv0 + xv1;
    }
    __syncthreads();
  }
  bool bad = false;
  for (int i = tid; i < n; i += NTHREADS) {
    const double v = ysol[i];
    p.phi[(size_t)eq * p.ldphi + i] = v;
    bad |= !isfinite(v);
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  K2_TICK(6);
  // sticky: the first failure of an equation is kept until the host clears it
  if (tid == 0 && (fail || anybad)) atomicCAS(&p.status[eq], 0, fail ? fail : -1);
}


// Solve-only companion (triangular algorithm, sample_coefficients.cpp:118-146): the
// systems were factorised in batch (mode 1); equation `eq`'s right-hand side only
// becomes known once the previous equations have been drawn.  One CTA:
//   forward   y_blk = Linv_kk (b_blk - L[blk, 0:c0] y_{0:c0})
//   backward  phi   = L^-T (y + z)            (same block scheme as above)
__global__ void __launch_bounds__(K2_NWARPS * 32, 1) k2_solve_kernel(const K2Params p, const double* __restrict__ linv_ws,
                                                                     const double* __restrict__ bvec, int eq0) {
  constexpr int NWARPS = K2_NWARPS, NTHREADS = K2_NWARPS * 32;
  extern __shared__ __align__(16) double smem[];
  const int n = p.n;
  const int nbk = k2_nblocks(n);
  double* ysol = smem;                       // [nbk*32 + 32]
  double* part = ysol + nbk * K2_NB + 32;    // [NWARPS][32]
  double* sblk = part + NWARPS * K2_NB;      // [32]
  int* colb = reinterpret_cast<int*>(sblk + 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int eq = eq0 + blockIdx.x;
  const double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  const double* __restrict__ lws = linv_ws + (size_t)eq * nbk * (K2_NB * K2_NB);
  const double* __restrict__ b = bvec + (size_t)blockIdx.x * n;
  for (int i = tid; i < n; i += NTHREADS) colb[i] = p.colbase[i];
  for (int i = tid; i < nbk * K2_NB + 32; i += NTHREADS) ysol[i] = 0.0;
  __syncthreads();
  // ---- forward
  for (int k = 0; k < nbk; ++k) {
    const int c0 = k * K2_NB;
    const int w = min(K2_NB, n - c0);
    // lanes = the 32 rows of the block, warps split the c0 previous columns
    double s = 0.0;
    if (lane < w) {
      for (int c = warp; c < c0; c += NWARPS) s = fma(om[colb[c] + c0 + lane], ysol[c], s);
    }
    part[warp * K2_NB + lane] = s;
    __syncthreads();
    if (warp == 0) {
      double t = 0.0;
#pragma unroll
      for (int ww = 0; ww < NWARPS; ++ww) t += part[ww * K2_NB + lane];
      sblk[lane] = (lane < w) ? b[c0 + lane] - t : 0.0;
      __syncwarp();
      // y[c] = sum_{c' <= c} Linv[c][c'] s[c'],  ws[c][c'] = Linv[c][c'] -> lane = c'?  use lane = c:
      const double* wsk = lws + (size_t)k * (K2_NB * K2_NB);
      double yv = 0.0;
#pragma unroll 8
      for (int cp = 0; cp < K2_NB; ++cp) yv = fma(wsk[lane * K2_NB + cp], sblk[cp], yv);
      if (lane < w) ysol[c0 + lane] = yv;
    }
    __syncthreads();
  }
  // ---- backward
  for (int i = tid; i < n; i += NTHREADS) ysol[i] += p.z[(size_t)blockIdx.x * p.ldz + i];
  __syncthreads();
  for (int k = nbk - 1; k >= 0; --k) {
    const int c0 = k * K2_NB;
    const int w = min(K2_NB, n - c0);
    const int kb_end = c0 + w;
    for (int c = warp; c < K2_NB; c += NWARPS) {
      double s = 0.0;
      if (c < w) {
        const double* colp = om + colb[c0 + c];
        for (int i = kb_end + lane; i < n; i += 32) s = fma(colp[i], ysol[i], s);
        s = warp_sum(s);
      }
      if (lane == 0) sblk[c] = (c < w) ? ysol[c0 + c] - s : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      const double* wsk = lws + (size_t)k * (K2_NB * K2_NB);
      double xv0 = 0.0, xv1 = 0.0;
#pragma unroll
      for (int cp = 0; cp < K2_NB; cp += 2) {
        xv0 = fma(wsk[cp * K2_NB + lane], sblk[cp], xv0);
        xv1 = fma(wsk[(cp + 1) * K2_NB + lane], sblk[cp + 1], xv1);
      }
      if (lane < w) ysol[c0 + lane] = xv0 + xv1;
    }
    __syncthreads();
  }
  bool bad = false;
  for (int i = tid; i < n; i += NTHREADS) {
    const double v = ysol[i];
    p.phi[(size_t)blockIdx.x * p.ldphi + i] = v;
    bad |= !isfinite(v);
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (tid == 0 && anybad) atomicCAS(&p.status[eq], 0, -1);
}

static double* g_linv_ws = nullptr;
static size_t g_linv_ws_bytes = 0;

cudaError_t launch_k2(const K2Params& p, cudaStream_t s) {
  const int nbk = k2_nblocks(p.n);
  const size_t need = (size_t)p.M * nbk * K2_NB * K2_NB * sizeof(double);
  if (need > g_linv_ws_bytes) {
    if (g_linv_ws) cudaFree(g_linv_ws);
    cudaError_t e = cudaMalloc(&g_linv_ws, need);
    if (e != cudaSuccess) { g_linv_ws = nullptr; g_linv_ws_bytes = 0; return e; }
    g_linv_ws_bytes = need;
  }
  const size_t smem = k2_smem_bytes(p.n);
  if (smem > 227 * 1024 || p.n + 1 > K2_NWARPS * K2_MAXQ * 8) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k2_cholsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k2_cholsolve<<<p.M, K2_NWARPS * 32, smem, s>>>(p, g_linv_ws);
  return cudaGetLastError();
}

// Solve `count` already-factorised equations eq0..eq0+count-1 (p.omega / status are
// the batch arrays; p.z, p.phi and bvec are indexed from the first solved equation).
cudaError_t launch_k2_solve(const K2Params& p, const double* bvec, int eq0, int count, cudaStream_t s) {
  if (!g_linv_ws) return cudaErrorInvalidValue;
  const int nbk = k2_nblocks(p.n);
  const size_t smem = ((size_t)(nbk * K2_NB + 32) + K2_NWARPS * K2_NB + 32) * sizeof(double) + (size_t)(p.n + 4) * sizeof(int);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k2_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k2_solve_kernel<<<count, K2_NWARPS * 32, smem, s>>>(p, g_linv_ws, bvec, eq0);
  return cudaGetLastError();
}

}  // namespace bvb
