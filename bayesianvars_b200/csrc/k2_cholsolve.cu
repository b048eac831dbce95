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

// v1 processes a block column in ROW TILES of the 32 block rows + at most `rt` sub-diagonal rows, so the
// shared-memory panel no longer has to hold a whole column of the system: rt = n + 1 (one tile, the whole
// column) whenever that fits, else K2_RT.  Row stride of the staged columns for a given rt:
constexpr int K2_RT = 320;   // sub-diagonal rows per tile of a system too tall for one tile (multiple of 16)
__host__ __device__ inline int k2_tile_rows(int n, int rt) { return (n + 1 < K2_NB + rt) ? n + 1 : K2_NB + rt; }
__host__ __device__ inline int k2_row_stride_rt(int n, int rt) { return k2_row_stride(k2_tile_rows(n, rt) - 1); }

// smem: [stage ring | panel (aliased)] [linvT] [ysol] [sblk] [colbuf 2x32] [rsbuf 32] [colbase ints]
static size_t k2_smem_bytes_rt(int n, int rt) {
  size_t rs = k2_row_stride_rt(n, rt);
  size_t ring = (size_t)K2_STAGES * K2_BKC * rs, panel = (size_t)K2_NB * rs;
  size_t d = (ring > panel ? ring : panel) + (size_t)K2_NB * K2_LS + (size_t)(k2_nblocks(n) * K2_NB + 32) +
             32 + 64 + 32 + (size_t)(n + 3) / 2 + 8;
  return d * sizeof(double);
}
size_t k2_smem_bytes(int n) { return k2_smem_bytes_rt(n, n + 1); }
// largest sub-row tile for a system of size n (0: not even the tiled form fits - n beyond ~9000)
int k2_pick_rt(int n) {
  if (k2_smem_bytes_rt(n, n + 1) <= (size_t)227 * 1024 && n + 1 <= K2_NWARPS * K2_MAXQ * 8) return n + 1;
  if (k2_smem_bytes_rt(n, K2_RT) <= (size_t)227 * 1024) return K2_RT;
  return 0;
}
bool k2_supported(int n) { return n > 0 && k2_pick_rt(n) > 0; }


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
      const double a = col[8 * (warp + K2_NWARPS * q)];   // `warp` = first tile of this warp (mt0 + warp id)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884(acc[q][nt][0], acc[q][nt][1], a, b[nt]);
    }
  }
}
template <>
__device__ __forceinline__ void k2_stage_mma<0>(double (&)[K2_MAXQ][4][2], const double*, int, int, int, int) {}

// Cluster form (CL = true): a thread-block cluster of `csize` CTAs shares one equation.  Every block column is cut into
// csize row tiles (the owner of block column k - rank k mod csize - takes the 32 block rows + the first tile and
// factorises the diagonal block); the inverse of the diagonal block travels through global memory (lws), two cluster
// barriers per block column order "Linv ready" and "column written back".  All reads of data another CTA wrote bypass
// L1 (cp.async.cg, ld.global.cg).  Used when equations x csize <= SMs (multi-GPU shards, small M): the per-equation
// latency chain is what bounds K2, and this is the only way more SMs shorten it.
__device__ __forceinline__ void k2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
template <bool CL>
__global__ void __launch_bounds__(K2_NWARPS * 32, 1) k2_cholsolve(const K2Params p, double* __restrict__ linv_ws, int rt, int csize) {
  constexpr int NWARPS = K2_NWARPS, MAXQ = K2_MAXQ, NTHREADS = K2_NWARPS * 32;
  extern __shared__ __align__(16) double smem[];
  const int n = p.n, n1 = n + 1;
  const int RS = k2_row_stride_rt(n, rt);   // (cluster form: rt = rows of the tallest tile, computed by the launcher)
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
  if (tid == 0) bvb_stamp_begin(p.tstamp);
  const int eq = CL ? blockIdx.x / csize : blockIdx.x;
  const int crank = CL ? blockIdx.x % csize : 0;          // rank in the cluster (cluster dims = (csize, 1, 1))
  double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  double* __restrict__ lws = linv_ws + (size_t)eq * nbk * (K2_NB * K2_NB);
  int fail = 0;  // meaningful on warp 0 only
  // optional phase profile (cycles, CTA 0 / thread 0): gemm, fill, factor, invert, trsm, writeback, backsolve
  long long tprev = 0;
  const bool prof = (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
  if (prof) tprev = clock64();
#define K2_TICK(slot) do { if (prof) { const long long tn = clock64(); p.prof[slot] += tn - tprev; tprev = tn; } } while (0)

  for (int i = tid; i < n; i += NTHREADS) colb[i] = p.colbase[i];
  __syncthreads();

  for (int k = 0; k < nbk; ++k) {
    const int c0 = k * K2_NB;
    const int w = min(K2_NB, n - c0);
    const int kb_end = c0 + w;
    const int nsub_all = n1 - kb_end;
    const bool full = (w == K2_NB);
   // row tiles: the 32 block rows + sub-diagonal rows [sub0, sub0 + nsub); tile 0 also factorises the diagonal block
   // cluster form: one tile per CTA and block column, ownership of the diagonal block rotates with k
   const int rtk = CL ? (((nsub_all + csize - 1) / csize + 7) & ~7) : rt;
   const int my_tile = CL ? (crank - k % csize + csize) % csize : 0;
   const bool has_tile = !CL || my_tile * rtk < nsub_all;
   if (CL && !has_tile) k2_cluster_sync();              // "Linv ready" of a block column this CTA has no rows in
   for (int sub0 = my_tile * rtk; sub0 < nsub_all; sub0 += (CL ? n1 : rtk)) {
    const bool first = (sub0 == 0);
    const int nsub = min(rtk, nsub_all - sub0);
    const int R = K2_NB + nsub;
    const int mtiles = (R + 7) >> 3;
    const int mt0 = first ? 0 : 4;                       // later tiles only produce their sub-diagonal rows
    const int li0 = first ? 0 : K2_NB;
    // local row -> global row (or -1 for the virtual identity rows of a partial block)
    auto grow = [&](int li) -> int { return li < K2_NB ? (li < w ? c0 + li : -1) : kb_end + sub0 + li - K2_NB; };

    // ---- B: left-looking update with the already-factorised columns -----------
    // chunk = 16 columns of L, one column per warp in the loader (coalesced 16-byte cp.async)
    const int nchunks = c0 / K2_BKC;
    auto load_chunk = [&](int st, int ch) {
      double* sdst = stage + (size_t)st * (K2_BKC * RS) + warp * RS;
      const double* src = om + colb[ch * K2_BKC + warp];
      if (full && first) {
        const int npair = (R + 1) >> 1;
        for (int pr = lane; pr < npair; pr += 32) cp_async16(sdst + 2 * pr, src + c0 + 2 * pr);
      } else if (full) {
        if (lane < K2_NB / 2) cp_async16(sdst + 2 * lane, src + c0 + 2 * lane);
        const int npair = (nsub + 1) >> 1;
        for (int pr = lane; pr < npair; pr += 32) cp_async16(sdst + K2_NB + 2 * pr, src + kb_end + sub0 + 2 * pr);
      } else {
        for (int li = lane; li < R; li += 32) {
          const int g = grow(li);
          if (g < 0) sdst[li] = 0.0;
          else if (CL) sdst[li] = __ldcg(src + g);
          else cp_async8(sdst + li, src + g);
        }
      }
    };

    const int wt = mt0 + warp;                           // first tile of this warp
    const int nq = (mtiles > wt) ? (mtiles - wt + NWARPS - 1) / NWARPS : 0;
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
        case 1: k2_stage_mma<1>(acc, sb, RS, wt, fr, fk); break;
        case 2: k2_stage_mma<2>(acc, sb, RS, wt, fr, fk); break;
        case 3: k2_stage_mma<3>(acc, sb, RS, wt, fr, fk); break;
        case 4: k2_stage_mma<4>(acc, sb, RS, wt, fr, fk); break;
        default: break;
      }
    }
    cp_async_wait<0>();
    __syncthreads();   // ring is dead from here on; its memory becomes the panel
    if (prof && k < 16) p.prof[8 + k] += clock64() - tprev;
    K2_TICK(0);

    // ---- A: panel = A_orig - acc, straight from the accumulator fragments -------
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int mt = wt + NWARPS * q;
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
            if (c < w) v = (g >= c0 + c) ? (CL ? __ldcg(om + colb[c0 + c] + g) : om[colb[c0 + c] + g]) : 0.0;
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
    if (warp == 0 && first) {
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
    if (CL) {
      k2_cluster_sync();                                 // "Linv ready": the owner's stores to lws are released
      if (!first) {
        const double* wsk = lws + (size_t)k * (K2_NB * K2_NB);
        for (int t = tid; t < K2_NB * K2_NB; t += NTHREADS) linvT[(t & 31) * K2_LS + (t >> 5)] = __ldcg(wsk + t);
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
      for (int li = li0 + lane; li < R; li += 32) {
        const int g = grow(li);
        if (g >= c0 + c) dst[g] = src[li];
      }
    }
    if (!CL && tid < K2_NB && sub0 + nsub == nsub_all) ysol[c0 + tid] = (tid < w) ? panel[tid * RS + (R - 1)] : 0.0;   // rhs row: last tile
    __syncthreads();
    K2_TICK(5);
   }
   if (CL) k2_cluster_sync();                            // "column written back": the next block column may read it
  }
  if (CL) {
    if (crank != 0) {                                    // the solve is rank 0's; report this CTA's pivots and leave
      if (tid == 0 && fail) atomicCAS(&p.status[eq], 0, fail);
      return;
    }
    for (int c = tid; c < nbk * K2_NB; c += NTHREADS) ysol[c] = (c < n) ? __ldcg(om + colb[c] + n) : 0.0;   // the rhs row of L
    __syncthreads();
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
        for (int i = kb_end + lane; i < n; i += 32) s = fma(CL ? __ldcg(colp + i) : colp[i], ysol[i], s);
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
        xv0 = fma(CL ? __ldcg(wsk + cp * K2_NB + lane) : wsk[cp * K2_NB + lane], sblk[cp], xv0);
        xv1 = fma(CL ? __ldcg(wsk + (cp + 1) * K2_NB + lane) : wsk[(cp + 1) * K2_NB + lane], sblk[cp + 1], xv1);
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
  if (tid == 0) bvb_stamp_end(p.tstamp);
  // sticky: the first failure of an equation is kept until the host clears it
  if (tid == 0 && (fail || anybad)) atomicCAS(&p.status[eq], 0, fail ? fail : -1);
}


// Solve-only companion (triangular algorithm, sample_coefficients.cpp:118-146): the systems were factorised
// in batch (mode 1); equation `eq`'s right-hand side only becomes known once the previous equations have been
// drawn, so this kernel sits on the sequential critical path M times per sweep.  One CTA per solve, bounded by
// streaming L (2 x 650 KB at n = 401) from L2: both passes walk L by column panels P_k = L[kb_end.., block k]
// (contiguous runs), prefetched with cp.async two chunks ahead of the dependency chain, which then only touches
// shared memory:
//   forward  (right-looking)  y_k = Linv_kk b_k ;  b[rows below] -= P_k y_k
//   backward                  x_k = Linv_kk^T ( (y+z)_k - P_k^T x[rows below] )
constexpr int K2S_RC = 176;                    // rows of a staged panel chunk (multiple of 16)
constexpr int K2S_RCS = K2S_RC + 4;            // its row stride (== 4 mod 16)
constexpr int K2S_STAGES = 4;                  // three items in flight ahead of the one being consumed
constexpr int K2S_LS = 33;                     // row stride of a staged Linv block

__global__ void __launch_bounds__(K2_NWARPS * 32, 1) k2_solve_kernel(const K2Params p, const double* __restrict__ linv_ws,
                                                                     const double* __restrict__ bvec, int eq0) {
  constexpr int NWARPS = K2_NWARPS, NTHREADS = K2_NWARPS * 32;
  constexpr int STAGE_D = K2_NB * K2S_RCS;                      // >= 32*32: a stage also holds one Linv block
  extern __shared__ __align__(16) double smem[];
  const int n = p.n;
  const int nbk = k2_nblocks(n);
  double* ring = smem;                                          // [STAGES][32][RCS] panel chunks (column-major) / Linv blocks
  double* xs = ring + (size_t)K2S_STAGES * STAGE_D;             // [nbk*32 + 32] running rhs, then the solution
  double* ybuf = xs + nbk * K2_NB + 32;                         // [32] y_k of the current block
  double* sblk = ybuf + 32;                                     // [32] column sums of the backward pass
  int* colb = reinterpret_cast<int*>(sblk + 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int eq = eq0 + blockIdx.x;
  const double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  const double* __restrict__ lws = linv_ws + (size_t)eq * nbk * (K2_NB * K2_NB);
  const double* __restrict__ b = bvec + (size_t)blockIdx.x * n;
  for (int i = tid; i < n; i += NTHREADS) colb[i] = p.colbase[i];
  for (int i = tid; i < nbk * K2_NB + 32; i += NTHREADS) xs[i] = (i < n) ? b[i] : 0.0;
  __syncthreads();

  // The item stream, in the order the solve consumes it:
  //   forward  k = 0..nbk-1 :  Linv_k, then the chunks of P_k (rows kb_end.. in steps of RC)
  //   backward k = nbk-1..0 :  the chunks of P_k, then Linv_k
  // Fetch cursor: (pass, k, r0) with r0 < 0 meaning "the Linv item of block k"; pass == 2 is the end.
  auto kb_end_of = [&](int k) { return min(n, (k + 1) * K2_NB); };
  int fpass = 0, fk = 0, fr0 = -1;
  auto advance = [&]() {
    if (fpass == 0) {
      if (fr0 < 0) fr0 = kb_end_of(fk); else fr0 += K2S_RC;
      if (fr0 >= n) {                                    // block exhausted: next block's Linv, or turn around
        if (fk + 1 < nbk) { ++fk; fr0 = -1; }
        else { fpass = 1; fr0 = kb_end_of(fk); if (fr0 >= n) fr0 = -1; }
      }
    } else if (fpass == 1) {
      if (fr0 >= 0) { fr0 += K2S_RC; if (fr0 >= n) fr0 = -1; }
      else {
        --fk;
        if (fk < 0) fpass = 2;
        else { fr0 = kb_end_of(fk); if (fr0 >= n) fr0 = -1; }
      }
    }
  };
  int fetched = 0, consumed = 0;
  auto issue_next = [&]() {
    if (fpass < 2) {
      double* dst = ring + (size_t)(fetched % K2S_STAGES) * STAGE_D;
      if (fr0 < 0) {
        // one Linv block, staged with row stride 33 (row c of Linv is read by lane c: stride 32 would be a 32-way
        // bank conflict); 8-byte copies because the padded rows are not 16-byte aligned
        const double* src = lws + (size_t)fk * (K2_NB * K2_NB);
        for (int t = tid; t < K2_NB * K2_NB; t += NTHREADS) cp_async8(dst + (t >> 5) * K2S_LS + (t & 31), src + t);
      } else {
        const int c0 = fk * K2_NB, w = min(K2_NB, n - c0);
        const int nr = min(K2S_RC, n - fr0);
        const int col = tid >> 4, t16 = tid & 15;        // 32 columns x 16 threads
        if (col < w) {
          const double* src = om + colb[c0 + col] + fr0;
          double* d = dst + (size_t)col * K2S_RCS;
          const int npair = (nr + 1) >> 1;
          for (int pr = t16; pr < npair; pr += 16) cp_async16(d + 2 * pr, src + 2 * pr);
        }
      }
      ++fetched;
      advance();
    }
    cp_async_commit();
  };
  // wait for the next item of the stream, refill the stage freed by the previous one, return the item's stage
  auto next_item = [&]() -> const double* {
    cp_async_wait<K2S_STAGES - 2>();
    __syncthreads();
    issue_next();
    const double* st = ring + (size_t)(consumed % K2S_STAGES) * STAGE_D;
    ++consumed;
    return st;
  };
  for (int sgi = 0; sgi < K2S_STAGES - 1; ++sgi) issue_next();

  // ---- forward (right-looking):  y_k = Linv_kk b_k ;  b[rows below] -= P_k y_k
  for (int k = 0; k < nbk; ++k) {
    const int c0 = k * K2_NB, w = min(K2_NB, n - c0), kb_end = c0 + w;
    {
      const double* wsk = next_item();                   // wsk[c*33 + c'] = Linv[c][c']
      if (warp == 0) {
        double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
        for (int cp = 0; cp < K2_NB; cp += 4) {
          y0 = fma(wsk[lane * K2S_LS + cp], (cp < w) ? xs[c0 + cp] : 0.0, y0);
          y1 = fma(wsk[lane * K2S_LS + cp + 1], (cp + 1 < w) ? xs[c0 + cp + 1] : 0.0, y1);
          y2 = fma(wsk[lane * K2S_LS + cp + 2], (cp + 2 < w) ? xs[c0 + cp + 2] : 0.0, y2);
          y3 = fma(wsk[lane * K2S_LS + cp + 3], (cp + 3 < w) ? xs[c0 + cp + 3] : 0.0, y3);
        }
        double yv = (y0 + y1) + (y2 + y3);
        yv = (lane < w) ? yv : 0.0;
        ybuf[lane] = yv;
        if (lane < w) xs[c0 + lane] = yv;
      }
    }
    for (int r0 = kb_end; r0 < n; r0 += K2S_RC) {
      const double* P = next_item();                     // (its barrier also publishes ybuf)
      const int nr = min(K2S_RC, n - r0);
      if (tid < nr) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
        for (int c = 0; c < K2_NB; c += 2) {
          s0 = fma(P[(size_t)c * K2S_RCS + tid], (c < w) ? ybuf[c] : 0.0, s0);
          s1 = fma(P[(size_t)(c + 1) * K2S_RCS + tid], (c + 1 < w) ? ybuf[c + 1] : 0.0, s1);
        }
        xs[r0 + tid] -= s0 + s1;
      }
    }
  }
  // ---- backward:  L^T phi = y + z ;  x_k = Linv_kk^T ( (y+z)_k - P_k^T x[rows below] )
  __syncthreads();
  for (int i = tid; i < n; i += NTHREADS) xs[i] += p.z[(size_t)blockIdx.x * p.ldz + i];
  for (int k = nbk - 1; k >= 0; --k) {
    const int c0 = k * K2_NB, w = min(K2_NB, n - c0), kb_end = c0 + w;
    __syncthreads();                                     // previous block's x_k (and the z add) visible; sblk free
    if (tid < K2_NB) sblk[tid] = 0.0;
    for (int r0 = kb_end; r0 < n; r0 += K2S_RC) {
      const double* P = next_item();
      const int nr = min(K2S_RC, n - r0);
      for (int c = warp; c < w; c += NWARPS) {           // one warp per column: no race on sblk[c]
        const double* colp = P + (size_t)c * K2S_RCS;
        double sc = 0.0;
        for (int i = lane; i < nr; i += 32) sc = fma(colp[i], xs[r0 + i], sc);
        sc = warp_sum(sc);
        if (lane == 0) sblk[c] += sc;
      }
    }
    const double* wsk = next_item();                     // (its barrier also publishes sblk)
    if (warp == 0) {
      // x[c'] = sum_{c >= c'} Linv[c][c'] (y[c] - s[c])
      double xv0 = 0.0, xv1 = 0.0;
#pragma unroll
      for (int cp = 0; cp < K2_NB; cp += 2) {
        xv0 = fma(wsk[cp * K2S_LS + lane], (cp < w) ? xs[c0 + cp] - sblk[cp] : 0.0, xv0);
        xv1 = fma(wsk[(cp + 1) * K2S_LS + lane], (cp + 1 < w) ? xs[c0 + cp + 1] - sblk[cp + 1] : 0.0, xv1);
      }
      __syncwarp();
      if (lane < w) xs[c0 + lane] = xv0 + xv1;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  bool bad = false;
  for (int i = tid; i < n; i += NTHREADS) {
    const double v = xs[i];
    p.phi[(size_t)blockIdx.x * p.ldphi + i] = v;
    bad |= !isfinite(v);
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (tid == 0 && anybad) atomicCAS(&p.status[eq], 0, -1);
}

// ---------------------------------------------------------------------------------------------
// K2 v2 - the same factorisation, warp-specialised with one block column of LOOK-AHEAD.
//
// In v1 the 32x32 diagonal factor + inverse (one warp, a 32-pivot dependent chain) runs while
// the other 15 warps wait, and every warp alternates between issuing cp.async and DMMA in
// lock step, so neither the L2 stream (latency-bound at ~24 B/clk/SM) nor the tensor pipe is
// kept busy.  Here the 16 warps have three roles:
//   warp 0      panel warp : factorises + inverts the diagonal block k
//   warps 1..15 GEMM warps : part A = update of block k+1 with the columns [0, 32k) that do not
//                            depend on block k (overlaps the panel warp; streamed from L2 through
//                            a cp.async ring whose loads are issued in the shadow of the DMMAs,
//                            a quarter of the next chunk after every k-step), part B = the 32
//                            columns of block k read straight from the shared-memory panel,
//                            TRSM, write-back, fill.
// (Measured dead ends: one dedicated loader warp sustains only ~4 B/clk of cp.async - the
// per-warp limit on copies in flight - and 1-D TMA bulk copies of these 1-3 KB columns were
// ~15% slower than 16-byte cp.async spread over all warps.)
// The panel and the ring no longer alias; the ring gets a per-block row stride and as many
// stages as fit (2..4).  Arithmetic and summation order are those of v1.
// ---------------------------------------------------------------------------------------------
constexpr int K2V_G = K2_NWARPS - 1;          // GEMM warps
constexpr int K2V_GT = K2V_G * 32;            // GEMM threads
constexpr int K2V_MAXS = 4;                   // ring stages

__host__ __device__ inline int k2v_rs_local(int R) { return ((R + 15) / 16) * 16 + 4; }
__host__ __device__ inline size_t k2v_misc_doubles(int n) {
  return (size_t)K2_NB * K2_LS + (size_t)(k2_nblocks(n) * K2_NB + 32) + 32 + 64 + 32 + (size_t)(n + 3) / 2 + 8;
}
// smem: [panel 32 x RS] [ring >= 32 x RS] [linvT] [ysol] [sblk] [colbuf] [rsbuf] [colbase ints]
size_t k2v_smem_bytes(int n, size_t* ring_doubles) {
  const size_t rs = k2_row_stride(n);
  const size_t misc = k2v_misc_doubles(n);
  const size_t cap = (size_t)227 * 1024 / sizeof(double);
  size_t ring = (size_t)K2V_MAXS * K2_BKC * rs;
  if ((size_t)K2_NB * rs + ring + misc > cap) ring = cap - misc - (size_t)K2_NB * rs;
  if (ring_doubles) *ring_doubles = ring;
  return ((size_t)K2_NB * rs + ring + misc) * sizeof(double);
}
bool k2v_fits(int n) {
  const size_t rs = k2_row_stride(n);
  return ((size_t)2 * K2_NB * rs + k2v_misc_doubles(n)) * sizeof(double) <= (size_t)227 * 1024 && n + 1 <= K2V_G * K2_MAXQ * 8;
}

__device__ __forceinline__ void k2v_bar_gemm() { asm volatile("bar.sync 1, %0;" ::"n"(K2V_GT) : "memory"); }
__device__ __forceinline__ void k2v_bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }
// stage MMA for GEMM warp g owning the m8 tiles g + K2V_G q.  NQ is a template parameter on purpose (see v1).
template <int NQ>
__device__ __forceinline__ void k2v_stage_mma(double (&acc)[K2_MAXQ][4][2], const double* sb, int RS, int g, int fr, int fk) {
#pragma unroll
  for (int ks = 0; ks < K2_BKC / 4; ++ks) {
    const double* col = sb + (4 * ks + fk) * RS + fr;
    double b[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) b[nt] = col[8 * nt];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double a = col[8 * (g + K2V_G * q)];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884(acc[q][nt][0], acc[q][nt][1], a, b[nt]);
    }
  }
}
// (every call site inlines all four bodies, and the kernel is instruction-cache bound between phases: keep the
// number of call sites at two - the part A chunk loop and the part B loop)
__device__ __forceinline__ void k2v_stage(int nq, double (&acc)[K2_MAXQ][4][2], const double* sb, int RS, int g, int fr, int fk) {
  switch (nq) {
    case 1: k2v_stage_mma<1>(acc, sb, RS, g, fr, fk); break;
    case 2: k2v_stage_mma<2>(acc, sb, RS, g, fr, fk); break;
    case 3: k2v_stage_mma<3>(acc, sb, RS, g, fr, fk); break;
    case 4: k2v_stage_mma<4>(acc, sb, RS, g, fr, fk); break;
    default: break;
  }
}

// Part A without the shared-memory ring (p.direct, BVB_K2_DIRECT=0 switches it off): every A element of the update is
// used by exactly one warp (the owner of its m8 tile), so staging it through shared memory buys no reuse - each GEMM
// warp loads its own fragments straight from L2 (8 rows x 4 columns per load: four 64-byte segments), two k4-steps in
// flight, no barrier in the loop.  The B fragments (the 32 rows of block k+1) are the same for all 15 warps and come
// through L1.  Same k order, same accumulators: bit-identical to the ring form.  Measured 333.8 -> 328.6 us: the
// stream through shared memory was NOT what held part A at 8.5 B/clk/SM.  (Also measured: dealing the n8 units of the
// last, mostly empty round of m8 tiles one at a time over the warps - 47 tiles are 3 full rounds + 2 tiles - is SLOWER,
// 408 us: the extra units have no operand reuse and their second pass over k runs at two loads per DMMA.)
template <int NQ>
__device__ __forceinline__ void k2v_direct_mma(double (&acc)[K2_MAXQ][4][2], const double* __restrict__ om,
                                               const int* colb, int c0n, int c0, int g, int fr, int fk) {
  const double* rowbase = om + c0n + fr;
#pragma unroll 1
  for (int c = 0; c < c0; c += 8) {
    double av[2][NQ], bv[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const double* base = rowbase + colb[c + 4 * ks + fk];
#pragma unroll
      for (int q = 0; q < NQ; ++q) av[ks][q] = __ldcg(base + 8 * (g + K2V_G * q));
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bv[ks][nt] = __ldca(base + 8 * nt);
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[q][nt][0], acc[q][nt][1], av[ks][q], bv[ks][nt]);
  }
}
__device__ __forceinline__ void k2v_direct(int nq, double (&acc)[K2_MAXQ][4][2], const double* om, const int* colb, int c0n,
                                           int c0, int g, int fr, int fk) {
  switch (nq) {
    case 1: k2v_direct_mma<1>(acc, om, colb, c0n, c0, g, fr, fk); break;
    case 2: k2v_direct_mma<2>(acc, om, colb, c0n, c0, g, fr, fk); break;
    case 3: k2v_direct_mma<3>(acc, om, colb, c0n, c0, g, fr, fk); break;
    case 4: k2v_direct_mma<4>(acc, om, colb, c0n, c0, g, fr, fk); break;
    default: break;
  }
}

__global__ void __launch_bounds__(K2_NWARPS * 32, 1) k2_cholsolve_v2(const K2Params p, double* __restrict__ linv_ws, int ring_doubles) {
  constexpr int NWARPS = K2_NWARPS, MAXQ = K2_MAXQ, NTHREADS = K2_NWARPS * 32;
  extern __shared__ __align__(16) double smem[];
  const int n = p.n, n1 = n + 1;
  const int RS = k2_row_stride(n);
  const int nbk = k2_nblocks(n);
  double* panel = smem;                                   // [32][RS]
  double* ring = smem + (size_t)K2_NB * RS;               // [ring_doubles]
  double* linvT = ring + ring_doubles;                    // [32][36]
  double* ysol = linvT + K2_NB * K2_LS;                   // [nbk*32 + 32]
  double* sblk = ysol + nbk * K2_NB + 32;                 // [32]
  double* colbuf = sblk + 32;                             // [2][32]
  double* rsbuf = colbuf + 64;                            // [32]
  int* colb = reinterpret_cast<int*>(rsbuf + 32);         // [n]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) bvb_stamp_begin(p.tstamp);
  const int eq = blockIdx.x;
  double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  double* __restrict__ lws = linv_ws + (size_t)eq * nbk * (K2_NB * K2_NB);
  int fail = 0;  // meaningful on warp 0 only
  // phase profile (cycles seen by GEMM warp 1 lane 0 of CTA 0): 0 part A, 1 wait for the panel warp, 2 trsm,
  // 3 writeback + part B, 4 fill, 6 back-solve
  long long tprev = 0;
  const bool prof = (p.prof != nullptr) && eq == 0 && tid == 32;
  if (prof) tprev = clock64();
#define K2V_TICK(slot) do { if (prof) { const long long tn = clock64(); p.prof[slot] += tn - tprev; tprev = tn; } } while (0)

  for (int i = tid; i < n; i += NTHREADS) colb[i] = p.colbase[i];
  __syncthreads();

  auto blk_w = [&](int kk) { return min(K2_NB, n - kk * K2_NB); };
  auto blk_R = [&](int kk) { return K2_NB + n1 - (kk * K2_NB + blk_w(kk)); };

  // All roles execute the same sequence of CTA-wide barriers: two for block 0, then per block column
  //   #1 Linv(k) ready + part A(k+1) done   #2 TRSM(k) done   #3 panel(k) dead, prefetch landed   #4 panel(k+1) filled
  if (warp == 0) {
    // =================================== panel warp ===================================
    k2v_bar_all();   // block 0 prefetch landed
    k2v_bar_all();   // panel(0) filled
    const bool profp = (p.prof != nullptr) && eq == 0 && tid == 0;
    long long tp = profp ? clock64() : 0;
#define K2V_PTICK(slot) do { if (profp) { const long long tn = clock64(); p.prof[slot] += tn - tp; tp = tn; } } while (0)
    for (int k = 0; k < nbk; ++k) {
      const int c0 = k * K2_NB;
      const int w = min(K2_NB, n - c0);
      const int R = K2_NB + n1 - (c0 + w);
#ifdef K2V_COMPACT_PANEL   // 60 KB less code but 1.6x slower under contention with the GEMM warps (measured): off
      // ---- C: factor the 32x32 diagonal block and invert it.  Lane i holds row i in registers as a window that
      // slides with the column (a[u] = D[i][c+u]), so the column loop is NOT unrolled: the fully unrolled form is
      // ~75 KB of straight-line code that evicts the GEMM warps' loop from the instruction cache once per block.
      // Same arithmetic and order as v1; window slots of columns >= 32 hold don't-care values.
      double a[K2_NB];
#pragma unroll
      for (int u = 0; u < K2_NB; ++u) a[u] = (u <= lane) ? panel[u * RS + lane] : 0.0;
      double x = __shfl_sync(0xffffffffu, a[0], 0);
      if (!(x > 0.0) || !isfinite(x)) { if (fail == 0) fail = c0 + 1; x = 1.0; }
      double rs = rsqrt(x);
#pragma unroll 1
      for (int c = 0; c < K2_NB; ++c) {
        const double l = (lane >= c) ? a[0] * rs : 0.0;   // lane c: x * rsqrt(x) = sqrt(x)
        panel[c * RS + lane] = l;                          // final column c (zeros above the diagonal)
        if (lane == c) rsbuf[c] = rs;
        __syncwarp();
        const double* colc = panel + c * RS + c;           // colc[u] = L[c+u][c]
        a[0] = fma(-l, colc[1], a[1]);                     // next column first: its pivot heads the dependent chain
        if (c + 1 < K2_NB) {
          x = __shfl_sync(0xffffffffu, a[0], (c + 1) & 31);
          if (!(x > 0.0) || !isfinite(x)) { if (fail == 0) fail = c0 + c + 2; x = 1.0; }
          rs = rsqrt(x);
        }
#pragma unroll
        for (int u = 2; u < K2_NB; ++u) a[u - 1] = fma(-l, colc[u], a[u]);
        a[K2_NB - 1] = 0.0;
      }
      __syncwarp();
      // lane j solves L x = e_j -> column j of Linv (right-looking substitution, same sliding window)
      double sacc[K2_NB];
#pragma unroll
      for (int u = 0; u < K2_NB; ++u) sacc[u] = (u == lane) ? 1.0 : 0.0;
#pragma unroll 1
      for (int kk = 0; kk < K2_NB; ++kk) {
        const double xk = (kk >= lane) ? sacc[0] * rsbuf[kk] : 0.0;
        linvT[lane * K2_LS + kk] = xk;                                            // linvT[c'=j][c=i] = Linv[i][j]
        lws[(size_t)k * (K2_NB * K2_NB) + kk * K2_NB + lane] = xk;                // ws[k][c'=i][c=j] = Linv[i][j]
        const double* colk = panel + kk * RS + kk;
#pragma unroll
        for (int u = 1; u < K2_NB; ++u) sacc[u - 1] = fma(-colk[u], xk, sacc[u]);
        sacc[K2_NB - 1] = 0.0;
      }
#else
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
        linvT[lane * K2_LS + i] = sacc[i];
        lws[(size_t)k * (K2_NB * K2_NB) + i * K2_NB + lane] = sacc[i];
      }
#endif
      K2V_PTICK(24);
      k2v_bar_all();   // #1
      K2V_PTICK(25);
      k2v_bar_all();   // #2
      K2V_PTICK(26);
      ysol[c0 + lane] = (lane < w) ? panel[lane * RS + (R - 1)] : 0.0;            // forward-solved rhs (row n of L)
      k2v_bar_all();   // #3
      K2V_PTICK(27);
      k2v_bar_all();   // #4
      K2V_PTICK(28);
    }
  } else {
    // =================================== GEMM warps ===================================
    const int fr = lane >> 2, fk = lane & 3;
    // (measured: handing the tiles of the last, partial round to the GEMM warps that share a sub-partition with the panel
    // warp - warps 4, 8, 12 as g = 0, 1, 2 - costs 5 us: the panel warp's pivot chain is slowed more than the tiles gain)
    const int g = warp - 1;                                  // 0..14, owns m8 tiles g + 15 q
    const int tg = tid - 32;                                 // 0..479
    // original entries of a FULL block column kk (rows c0.., 32 columns) -> ring as [c][RSl]
    auto prefetch_block = [&](int kk) {
      const int c0 = kk * K2_NB, R = blk_R(kk), RSl = k2v_rs_local(R);
      const int col = tg / 15, r0 = tg - 15 * col;
      const double* src = om + colb[c0 + col] + c0;
      double* dst = ring + (size_t)col * RSl;
      const int npair = (R + 1) >> 1;
      for (int pr = r0; pr < npair; pr += 15) cp_async16(dst + 2 * pr, src + 2 * pr);
    };
    double acc[MAXQ][4][2];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;
    // panel(kk) = A_orig - acc   (`pre`: A_orig comes from the prefetched copy in the ring)
    auto fill_panel = [&](int kk, bool pre) {
      const int c0 = kk * K2_NB, w = blk_w(kk), kb_end = c0 + w, R = blk_R(kk), RSl = k2v_rs_local(R);
      const int mtiles = (R + 7) >> 3;
#pragma unroll
      for (int q = 0; q < MAXQ; ++q) {
        const int mt = g + K2V_G * q;
        const int li = 8 * mt + fr;
        if (mt < mtiles && li < R) {
          const int gr = li < K2_NB ? (li < w ? c0 + li : -1) : kb_end + li - K2_NB;
          double orig[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 8 * nt + 2 * fk + e;
              double v;
              if (c < w) v = (gr >= c0 + c) ? (pre ? ring[(size_t)c * RSl + li] : om[colb[c0 + c] + gr]) : 0.0;
              else v = (li == c) ? 1.0 : 0.0;
              orig[nt][e] = v;
            }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 8 * nt + 2 * fk + e;
              double v = orig[nt][e];
              if (c < w && gr >= c0 + c) v -= acc[q][nt][e];
              panel[c * RS + li] = v;
            }
        }
      }
    };

    // k = -1 is a virtual iteration: nothing to factor yet, it only prefetches and fills block 0 (acc = 0), so
    // fill_panel / prefetch_block have a single call site each.
    for (int k = -1; k < nbk; ++k) {
      const bool real = (k >= 0);
      const int c0 = k * K2_NB;
      const int w = min(K2_NB, n - c0);
      const int kb_end = c0 + w;
      const int R = K2_NB + n1 - kb_end;
      const int mtiles = real ? (R + 7) >> 3 : 0;
      auto grow = [&](int li) -> int { return li < K2_NB ? (li < w ? c0 + li : -1) : kb_end + li - K2_NB; };
      // next block column
      const bool have_next = (k + 1 < nbk);
      const int c0n = c0 + K2_NB;
      const int wn = have_next ? min(K2_NB, n - c0n) : 0;
      const int Rn = have_next ? K2_NB + n1 - (c0n + wn) : 0;
      const bool fulln = (wn == K2_NB);
      const int mtn = (Rn + 7) >> 3;
      const int nqn = (have_next && mtn > g) ? (mtn - g + K2V_G - 1) / K2V_G : 0;
      const int RSl = k2v_rs_local(Rn);
      const int stage_d = K2_BKC * RSl;

      if (have_next && real) {
        // ---- part A of block k+1: columns [0, 32k), while warp 0 factorises block k
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;
        // A ring stage holds `cm` sub-chunks of 16 columns: late block columns have few rows, so 16 columns are too
        // little work per barrier (measured ~1.6k cycles per 16-column chunk regardless of its size); cm = 1, 2, 4 or 8.
        const int sub_d = K2_BKC * RSl;                      // one 16-column sub-chunk
        int cm = 1;
        if (ring_doubles >= 2 * 8 * sub_d) cm = 8; else if (ring_doubles >= 2 * 4 * sub_d) cm = 4; else if (ring_doubles >= 2 * 2 * sub_d) cm = 2;
        const int stage_cd = cm * sub_d;
        const int nsubs = c0 / K2_BKC;                       // 16-column sub-chunks to go through
        const int nchunks = (nsubs + cm - 1) / cm;
        int S = ring_doubles / stage_cd;
        S = S > 4 ? 4 : S;                                   // >= 2 by construction of the ring
        const int lcol = tg / 30, lr0 = tg - 30 * lcol;      // loader: 16 columns x 30 threads
        const int npair = (Rn + 1) >> 1;
        auto load_chunk = [&](int st, int ch) {
          for (int u = 0; u < cm; ++u) {
            const int sc = ch * cm + u;
            if (sc >= nsubs) break;
            double* sdst = ring + (size_t)st * stage_cd + (size_t)u * sub_d + lcol * RSl;
            const double* src = om + colb[sc * K2_BKC + lcol];
            if (fulln) {
              for (int pr = lr0; pr < npair; pr += 30) cp_async16(sdst + 2 * pr, src + c0n + 2 * pr);
            } else {
              for (int li = lr0; li < Rn; li += 30) {
                const int gg = li < K2_NB ? (li < wn ? c0n + li : -1) : c0n + wn + li - K2_NB;
                if (gg >= 0) cp_async8(sdst + li, src + gg);
                else sdst[li] = 0.0;
              }
            }
          }
        };
        if (p.direct && fulln) {
          k2v_direct(nqn, acc, om, colb, c0n, c0, g, fr, fk);
        } else {
        for (int s = 0; s < S - 1; ++s) {
          if (s < nchunks) load_chunk(s, s);
          cp_async_commit();
        }
        for (int ch = 0; ch < nchunks; ++ch) {
          if (S == 2) cp_async_wait<0>(); else if (S == 3) cp_async_wait<1>(); else cp_async_wait<2>();
          k2v_bar_gemm();                                    // chunk ch landed everywhere; chunk ch-1's stage is free
          {
            const int nxt = ch + S - 1;
            if (nxt < nchunks) load_chunk(nxt % S, nxt);      // issued before the MMAs: the copies are latency-bound
            cp_async_commit();
          }
          const int usub = min(cm, nsubs - ch * cm);
          for (int u = 0; u < usub; ++u)
            k2v_stage(nqn, acc, ring + (size_t)(ch % S) * stage_cd + (size_t)u * sub_d, RSl, g, fr, fk);
        }
        cp_async_wait<0>();
        }
        k2v_bar_gemm();                                      // ring is free
      }
      if (prof && real && k + 1 < 16) p.prof[8 + k + 1] += clock64() - tprev;
      K2V_TICK(0);
      if (have_next && fulln) prefetch_block(k + 1);
      cp_async_commit();
      if (real) k2v_bar_all();                               // #1: Linv ready, part A done
      K2V_TICK(1);

      // ---- D: sub-diagonal rows  X = A * Linv^T  (DMMA, skipping the zero triangle)
      for (int mt = 4 + g; mt < mtiles; mt += K2V_G) {
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
      if (real) k2v_bar_all();                               // #2: panel(k) = L[:, block k]
      K2V_TICK(2);

      // ---- E: write the factorised panel back; part B of block k+1 straight from the panel
      for (int c = g; real && c < w; c += K2V_G) {
        double* dst = om + colb[c0 + c];
        const double* src = panel + c * RS;
        for (int li = lane; li < R; li += 32) {
          const int gg = grow(li);
          if (gg >= c0 + c) dst[gg] = src[li];
        }
      }
      cp_async_wait<0>();                                    // the prefetched A_orig(k+1) has landed
      if (have_next && real && !fulln) {
        // partial (last) block: its rows are not contiguous in the panel.  Gather the two chunks of block k through
        // the row map into the ring (nothing was prefetched for a partial block, the ring is free)
        const int lcol = tg / 30, lr0 = tg - 30 * lcol;      // 16 columns x 30 threads
        for (int hh = 0; hh < 2; ++hh) {
          double* sdst = ring + (size_t)hh * stage_d + lcol * RSl;
          const double* src = panel + (size_t)(hh * K2_BKC + lcol) * RS;
          for (int li = lr0; li < Rn; li += 30) {
            // block k+1 local row -> global row -> panel(k) local row (panel(k) is a full block: row = 32 + g - c0n)
            const int gg = li < K2_NB ? (li < wn ? c0n + li : -1) : c0n + wn + li - K2_NB;
            sdst[li] = (gg >= 0) ? src[K2_NB + gg - c0n] : 0.0;
          }
        }
        k2v_bar_gemm();
      }
      if (have_next && real) {
        // part B: the 32 columns of block k.  Full block: the rows of block k+1 are the panel rows 32.. and its own
        // 32 rows are the B operand; partial block: the gathered copy in the ring.
        const double* pb = fulln ? panel + K2_NB : ring;
        const int pbRS = fulln ? RS : RSl;
        for (int hh = 0; hh < 2; ++hh) k2v_stage(nqn, acc, pb + (size_t)hh * K2_BKC * pbRS, pbRS, g, fr, fk);
      }
      k2v_bar_all();                                         // #3: panel(k) dead, prefetch visible
      K2V_TICK(3);
      if (have_next) fill_panel(k + 1, fulln);
      k2v_bar_all();                                         // #4: panel(k+1) ready
      K2V_TICK(4);
    }
  }

  if (p.mode == 1) {   // factor only: the caller solves later (triangular algorithm)
    if (tid == 0 && fail) atomicCAS(&p.status[eq], 0, fail);
    return;
  }
  // ---- back-solve  L^T phi = y + z ------------------------------------------------
  for (int i = tid; i < n; i += NTHREADS) ysol[i] += p.z[(size_t)eq * p.ldz + i];
  __syncthreads();
  // Every load of a block column is in flight before the first use (<= 12 values per lane and column at n <= 407;
  // the rolled form took one L2 round trip per four rows), and warp 0 fetches Linv_k before the dot products instead
  // of after the barrier behind them: 5 k -> 2 k cycles per block column.
  for (int k = nbk - 1; k >= 0; --k) {
    const int c0 = k * K2_NB;
    const int w = min(K2_NB, n - c0);
    const int rb = c0 + K2_NB;
    double lk[K2_NB];
    if (warp == 0) {
      const double* wsk = lws + (size_t)k * (K2_NB * K2_NB);
#pragma unroll
      for (int c = 0; c < K2_NB; ++c) lk[c] = wsk[c * K2_NB + lane];
    }
    constexpr int MAXR = 12;                                 // rows below a block: <= 407 - 32 = 375 < 12 * 32
    double v0[MAXR], v1[MAXR];
    const int ca = 2 * warp, cb2 = 2 * warp + 1;             // NWARPS = 16: two columns of the block per warp
    const double* pa = om + colb[min(c0 + ca, n - 1)];
    const double* pb = om + colb[min(c0 + cb2, n - 1)];
#pragma unroll
    for (int j = 0; j < MAXR; ++j) {
      const int i = min(rb + lane + 32 * j, n - 1);          // (clamped: always a valid address, masked below)
      v0[j] = pa[i];
      v1[j] = pb[i];
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < MAXR; ++j) {
      const int i = rb + lane + 32 * j;
      const double yi = (i < n) ? ysol[min(i, n - 1)] : 0.0;
      s0 = fma(v0[j], yi, s0);
      s1 = fma(v1[j], yi, s1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
      sblk[ca] = (ca < w) ? ysol[c0 + ca] - s0 : 0.0;
      sblk[cb2] = (cb2 < w) ? ysol[c0 + cb2] - s1 : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double xv0 = 0.0, xv1 = 0.0, xv2 = 0.0, xv3 = 0.0;
#pragma unroll
      for (int cp = 0; cp < K2_NB; cp += 4) {
        xv0 = fma(lk[cp], sblk[cp], xv0);
        xv1 = fma(lk[cp + 1], sblk[cp + 1], xv1);
        xv2 = fma(lk[cp + 2], sblk[cp + 2], xv2);
        xv3 = fma(lk[cp + 3], sblk[cp + 3], xv3);
      }
      if (lane < w) ysol[c0 + lane] = (xv0 + xv1) + (xv2 + xv3);
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
  K2V_TICK(6);
  if (tid == 0) bvb_stamp_end(p.tstamp);
  if (tid == 0 && (fail || anybad)) atomicCAS(&p.status[eq], 0, fail ? fail : -1);
}

// The [M][nbk][32 x 32] inverses of the diagonal blocks, shared by every K2 form and by the solve-only kernel; one
// workspace per device (a process may drive several GPUs: bvb_create_multi).
static double* g_linv_ws[16] = {nullptr};
static size_t g_linv_ws_bytes[16] = {0};

double* k2_linv_workspace(size_t need) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (need > g_linv_ws_bytes[dev]) {
    if (g_linv_ws[dev]) cudaFree(g_linv_ws[dev]);
    if (cudaMalloc(&g_linv_ws[dev], need) != cudaSuccess) { g_linv_ws[dev] = nullptr; g_linv_ws_bytes[dev] = 0; return nullptr; }
    g_linv_ws_bytes[dev] = need;
  }
  return g_linv_ws[dev];
}

cudaError_t launch_k2_tiled(const K2Params& p, cudaStream_t s);   // k2_tiled.cu

cudaError_t launch_k2(const K2Params& p, cudaStream_t s) {
  const int nbk = k2_nblocks(p.n);
  const size_t need = (size_t)p.M * nbk * K2_NB * K2_NB * sizeof(double);
  double* const linv_ws = k2_linv_workspace(need);
  if (!linv_ws) return cudaErrorMemoryAllocation;
  {
    // The tile-queue form (k2_tiled.cu) when the equations alone cannot fill the chip: measured at n = 401 on 148 SMs,
    // tiled / one-CTA-per-equation: 13 equations 201 / 303 us, 25: 213 / 307, 50: 230 / 334, 100: 459 / 319 (the look-ahead
    // kernel below keeps the headline shape).  BVB_K2_TILED=0 / 1 forces one or the other; the phase-profile and
    // timing-experiment switches of the kernels below select them too.
    static int tiled = -2, num_sms_t = 0;
    if (tiled == -2) {
      const char* ev = getenv("BVB_K2_TILED");
      tiled = ev ? atoi(ev) : -1;
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms_t, cudaDevAttrMultiProcessorCount, dev);
    }
    static const bool tprof = getenv("BVB_K2T_PROF") != nullptr;   // phase profile of the tiled form's chain step
    const bool want = tiled == 1 || (tiled == -1 && 2 * p.M <= num_sms_t && p.n >= 96);
    if (want && (p.prof == nullptr || tprof) && p.debug_mode == 0) {
      const cudaError_t te = launch_k2_tiled(p, s);
      if (te != cudaErrorNotSupported) return te;     // (more equations than chain CTAs fit: the kernels below)
    }
  }
  // Thread-block clusters when the equations leave SMs idle: csize = 8 or 4 with M * csize <= SMs.
  {
    static int num_sms = 0, forced = -1;
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
      const char* ev = getenv("BVB_K2_CLUSTER");
      forced = ev ? atoi(ev) : -1;
    }
    int cs = 1;
    // (measured at n = 401: 13 equations 338 -> 287 us with 8 CTAs, 25 equations 344 -> 305 us with 4; two CTAs per
    // equation only break even with the look-ahead kernel, so they are not used)
    for (int c = 8; c >= 4; c >>= 1)
      if (p.M * c <= num_sms && p.n + 1 - K2_NB >= 64 * c) { cs = c; break; }
    if (forced == 0 || forced == 1) cs = 1;
    else if (forced == 2 || forced == 4 || forced == 8) cs = (p.n + 1 - K2_NB >= 8 * forced) ? forced : 1;
    if (cs > 1) {
      const int rtc = ((((p.n + 1 - K2_NB) + cs - 1) / cs) + 7) & ~7;      // tallest tile (block column 0)
      if (k2_smem_bytes_rt(p.n, rtc) <= (size_t)227 * 1024 && K2_NB + rtc <= K2_NWARPS * K2_MAXQ * 8) {
        const size_t smemc = k2_smem_bytes_rt(p.n, rtc);
        cudaError_t ec = cudaFuncSetAttribute(k2_cholsolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc);
        if (ec != cudaSuccess) return ec;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.M * cs));
        cfg.blockDim = dim3(K2_NWARPS * 32);
        cfg.dynamicSmemBytes = smemc;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k2_cholsolve<true>, p, linv_ws, rtc, cs);
      }
    }
  }
  static const bool force_v1 = getenv("BVB_K2_V1") != nullptr;
  if (!force_v1 && k2v_fits(p.n)) {
    size_t ring = 0;
    const size_t smem2 = k2v_smem_bytes(p.n, &ring);
    cudaError_t e2 = cudaFuncSetAttribute(k2_cholsolve_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e2 != cudaSuccess) return e2;
    K2Params q = p;
    { const char* ed = getenv("BVB_K2_DIRECT"); if (ed) q.direct = atoi(ed); }
    k2_cholsolve_v2<<<q.M, K2_NWARPS * 32, smem2, s>>>(q, linv_ws, (int)ring);
    return cudaGetLastError();
  }
  const int rt = k2_pick_rt(p.n);
  if (rt == 0) return cudaErrorInvalidValue;
  const size_t smem = k2_smem_bytes_rt(p.n, rt);
  cudaError_t e = cudaFuncSetAttribute(k2_cholsolve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k2_cholsolve<false><<<p.M, K2_NWARPS * 32, smem, s>>>(p, linv_ws, rt, 1);
  return cudaGetLastError();
}

// Solve `count` already-factorised equations eq0..eq0+count-1 (p.omega / status are
// the batch arrays; p.z, p.phi and bvec are indexed from the first solved equation).
cudaError_t launch_k2_solve(const K2Params& p, const double* bvec, int eq0, int count, cudaStream_t s) {
  double* const linv_ws = k2_linv_workspace(0);      // (sized by the factorisation that came before)
  if (!linv_ws) return cudaErrorInvalidValue;
  const int nbk = k2_nblocks(p.n);
  const size_t smem = ((size_t)K2S_STAGES * K2_NB * K2S_RCS + (size_t)(nbk * K2_NB + 32) + 64) * sizeof(double) +
                      (size_t)(p.n + 4) * sizeof(int);
  if (smem > (size_t)227 * 1024) return cudaErrorInvalidValue;
  {
    cudaError_t e = cudaFuncSetAttribute(k2_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k2_solve_kernel<<<count, K2_NWARPS * 32, smem, s>>>(p, linv_ws, bvec, eq0);
  return cudaGetLastError();
}

}  // namespace bvb
