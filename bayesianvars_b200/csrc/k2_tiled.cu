// K2 (tiled form) - batched Cholesky + forward solve as a task queue over 32 x 32 tiles, then the back-solve / draw.
//
// Same contract and same storage as k2_cholsolve.cu (packed augmented systems written by K1, factorised in place,
// 32 x 32 inverses of the diagonal blocks in `linv_ws`; sample_coefficients.cpp:58-70), different schedule.  The
// one-CTA-per-equation kernels there leave a third of the chip idle at the headline shape (100 equations, 148 SMs) and
// walk every equation's 13 block columns in lock step inside one CTA (321 us however few equations a GPU owns).  Here
// the left-looking block algorithm is cut into tile tasks that any CTA can run:
//
//   DIAG(e, k)        diagonal tile of block column k of equation e:  A_kk + D_k - L(k,k-1) L(k,k-1)' -> chol, inverse
//   SUB(e, k, I0:I1)  row blocks I0..I1-1 (<= 128 rows) of block column k:  A - L(rows, <k) L(k, <k)'  (FP64 DMMA,
//                     operands straight from L2, columns consumed as they become final), times Linv_k', stored; and
//                     for every row block I >= k+2 of the tile the contribution - L(I,k) L(I,k)' is folded into the
//                     running diagonal accumulator D_I, so DIAG(I) never walks its own history.
//
// Tasks sit in ONE list in a topological order (stage k: DIAG(.,k), the far tiles of column k-1, the first tile of
// column k); persistent CTAs take the next index with an atomic and wait for what a task needs through per-(equation,
// row block) progress counters (release / acquire at gpu scope).  A waiting CTA only ever waits for tasks with a
// smaller index, which some resident CTA has already taken, so the scheme cannot deadlock whatever the number of
// resident CTAs, and every tile is computed by exactly one task from fixed inputs in a fixed order: results do not
// depend on which CTA ran what, on the grid size, or on how many equations share the launch (bitwise identical for
// 13 or 100 equations - the equation-sharded multi-GPU chain is the single-GPU chain).
#include "bvb_internal.cuh"

#include <stdlib.h>
#include <algorithm>
#include <utility>
#include <vector>

namespace bvb {

double* k2_linv_workspace(size_t bytes);   // k2_cholsolve.cu: the shared [M][nbk][32 x 32] buffer of diagonal-block inverses

constexpr int KT_NB = 32;
constexpr int KT_WARPS = 4, KT_THREADS = KT_WARPS * 32, KT_MAXQ = 4;
constexpr int KT_MAXROWS = KT_WARPS * KT_MAXQ * 8;   // 128 rows per SUB task
constexpr int KT_RS = KT_MAXROWS + 4;                // row stride of the panel (== 4 mod 16: conflict-free fragment loads)
constexpr int KT_LS = 36;                            // row stride of the 32 x 32 helper tiles
#ifndef KT_DEPTH
#define KT_DEPTH 2                                   // k4-steps of operands in flight per warp in the history product
#endif
#ifndef KT_CTAS
#define KT_CTAS 2                                    // resident CTAs per SM the register budget is set for
#endif

struct KtTask { int eq, k, i0, i1; };                // i1 == i0: DIAG(eq, k); else SUB over row blocks [i0, i1)

struct KtArgs {
  K2Params p;
  const KtTask* tasks;
  int ntasks;
  int* head;        // queue head (zeroed before the launch)
  int* progress;    // [M][nrb]: leading block columns of row block I that are final (zeroed before the launch)
  int* dprogress;   // [M][nrb]: 1 once the diagonal accumulator D_I is out
  int* pprogress;   // [M][nrb]: 1 once the look-ahead product P_I is out
  double* P;        // [M][nbk][32*32] column-major: sum_{b <= I-3} L(I,b) L(I-1,b)'  (the old history of the chain's tile (I, I-1))
  double* D;        // [M][nbk][32*32] column-major: - sum_{b <= I-2} L(I,b) L(I,b)'  (lower triangle), written by one task
  double* lws;      // [M][nbk][32*32] row-major inverses of the diagonal blocks
  int nbk, nrb;     // block columns; row blocks = nbk + 1 (the last one is the right-hand-side row n)
  int ncrit;        // CTAs [0, ncrit) walk the diagonal chains; `tasks` holds the far tiles only
};

// Counters are POLLED with relaxed loads and the acquire fence comes once, after the value has been seen (kt_acquire):
// a gpu-scope acquire (like __threadfence) invalidates the SM's L1, and a CTA spinning on ld.acquire would keep wiping
// the L1 lines its SM neighbours read their shared B operand through.  Publishing is bar.sync + ONE st.release.gpu by
// thread 0 (the barrier orders every thread's stores before it - the CUTLASS semaphore pattern), no per-thread fence.
__device__ __forceinline__ int kt_ld_acquire(const int* p) {   // relaxed poll; pair with kt_acquire()
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void kt_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void kt_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void kt_st_relaxed(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int kt_rowstart(int I, int nbk, int n) { return I < nbk ? KT_NB * I : n; }
__device__ __forceinline__ int kt_rowend(int I, int nbk, int n) { return I < nbk ? min(KT_NB * I + KT_NB, n) : n + 1; }

// acc[q][nt] += A(rows of m-tile q, columns [cbeg, cend)) * B(rows brow[nt], same columns)'.  Operands come straight from
// L2 / L1 into registers, one k4-step ahead of the DMMAs: A entries (this warp's own rows, read once) with ld.cg, B
// entries (the 32 rows of block k, read by every warp of every CTA working on this block column) through L1 - safe
// because tiles are only read once final and a 128-byte line never spans two tiles (prep.cu: 16-double column bases).
// PA (look-ahead product): the task that owns row block k+2 of block column k also forms, from the A fragments of its
// first 32 rows, P = sum_{b < k} L(k+2, b) L(k+1, b)' - the old part of the history of the tile the diagonal chain
// will need one step later (B rows brow2 = row block k+1) - so the chain only adds the newest block column itself.
// and D = sum_{b < k} L(k+2, b) L(k+2, b)' (B rows brow3 = its own first 32 rows), the history of the diagonal block two
// steps ahead: with this task's own block column added at the end, D_{k+2} is complete - written by ONE task.
template <int NQ, int DEPTH, bool PA>
__device__ __forceinline__ void kt_acc(double (&acc)[KT_MAXQ][4][2], double (&acc2)[4][2], double (&acc3)[4][2], const double* __restrict__ om,
                                       const int* colb, int cbeg, int cend, const int (&arow)[KT_MAXQ], const int (&brow)[4],
                                       const int (&brow2)[4], const int (&brow3)[4], int fk) {
  if (cbeg >= cend) return;
  double ar[DEPTH][NQ], br[DEPTH][4], br2[DEPTH][PA ? 8 : 1];
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) {
    if (cbeg + 4 * d < cend) {
      const double* col = om + colb[cbeg + 4 * d + fk];
#pragma unroll
      for (int q = 0; q < NQ; ++q) ar[d][q] = __ldcg(col + arow[q]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) br[d][nt] = col[brow[nt]];
      if (PA) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { br2[d][nt] = col[brow2[nt]]; br2[d][4 + nt] = col[brow3[nt]]; }
      }
    }
  }
  for (int c = cbeg; c < cend; c += 4 * DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (c + 4 * d < cend) {
        double ac[NQ], bc[4], bc2[PA ? 8 : 1];
#pragma unroll
        for (int q = 0; q < NQ; ++q) ac[q] = ar[d][q];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bc[nt] = br[d][nt];
        if (PA) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) bc2[nt] = br2[d][nt];
        }
        if (c + 4 * (d + DEPTH) < cend) {
          const double* col = om + colb[c + 4 * (d + DEPTH) + fk];
#pragma unroll
          for (int q = 0; q < NQ; ++q) ar[d][q] = __ldcg(col + arow[q]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) br[d][nt] = col[brow[nt]];
          if (PA) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { br2[d][nt] = col[brow2[nt]]; br2[d][4 + nt] = col[brow3[nt]]; }
          }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc[q][nt][0], acc[q][nt][1], ac[q], bc[nt]);
        if (PA) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            dmma884(acc2[nt][0], acc2[nt][1], ac[0], bc2[nt]);
            dmma884(acc3[nt][0], acc3[nt][1], ac[0], bc2[4 + nt]);
          }
        }
      }
    }
  }
}
template <int NQ>
__device__ __forceinline__ void kt_acc_any(bool pa, double (&acc)[KT_MAXQ][4][2], double (&acc2)[4][2], double (&acc3)[4][2],
                                           const double* __restrict__ om, const int* colb, int cbeg, int cend, const int (&arow)[KT_MAXQ],
                                           const int (&brow)[4], const int (&brow2)[4], const int (&brow3)[4], int fk) {
  if (pa) kt_acc<NQ, 1, true>(acc, acc2, acc3, om, colb, cbeg, cend, arow, brow, brow2, brow3, fk);   // (three products: one step ahead keeps it in registers)
  else kt_acc<NQ, KT_DEPTH, false>(acc, acc2, acc3, om, colb, cbeg, cend, arow, brow, brow2, brow3, fk);
}

// The same product for ONE warp working alone on a 32-row tile (the chain step): all four m-tiles, k4-steps
// [sbeg, send) of the history, FOUR steps of operands in flight in named registers (an L2 round trip is ~1000 cycles
// under load, four steps of DMMA issue).
#define KT_OP_LOAD(A, B, step)                                                   \
  {                                                                              \
    const double* col_ = om + colb[4 * (step) + fk];                             \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) A[q] = __ldcg(col_ + arow[q]); \
    _Pragma("unroll") for (int nt = 0; nt < 4; ++nt) B[nt] = col_[brow[nt]];     \
  }
#define KT_OP_STEP(A, B, step_next, have_next)                                                         \
  {                                                                                                    \
    double ta_[4], tb_[4];                                                                             \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) ta_[q] = A[q];                                       \
    _Pragma("unroll") for (int nt = 0; nt < 4; ++nt) tb_[nt] = B[nt];                                  \
    if (have_next) KT_OP_LOAD(A, B, step_next)                                                         \
    _Pragma("unroll") for (int q = 0; q < 4; ++q)                                                      \
      _Pragma("unroll") for (int nt = 0; nt < 4; ++nt) dmma884(acc[q][nt][0], acc[q][nt][1], ta_[q], tb_[nt]); \
  }
__device__ __forceinline__ void kt_acc_chain(double (&acc)[KT_MAXQ][4][2], const double* __restrict__ om, const int* colb, int sbeg,
                                             int send, const int (&arow)[KT_MAXQ], const int (&brow)[4], int fk) {
  double a0[4], b0[4], a1[4], b1[4], a2[4], b2[4], a3[4], b3[4];
  if (sbeg < send) KT_OP_LOAD(a0, b0, sbeg)
  if (sbeg + 1 < send) KT_OP_LOAD(a1, b1, sbeg + 1)
  if (sbeg + 2 < send) KT_OP_LOAD(a2, b2, sbeg + 2)
  if (sbeg + 3 < send) KT_OP_LOAD(a3, b3, sbeg + 3)
  for (int s0 = sbeg; s0 < send; s0 += 4) {
    KT_OP_STEP(a0, b0, s0 + 4, s0 + 4 < send)
    if (s0 + 1 < send) KT_OP_STEP(a1, b1, s0 + 5, s0 + 5 < send)
    if (s0 + 2 < send) KT_OP_STEP(a2, b2, s0 + 6, s0 + 6 < send)
    if (s0 + 3 < send) KT_OP_STEP(a3, b3, s0 + 7, s0 + 7 < send)
  }
}
#undef KT_OP_LOAD
#undef KT_OP_STEP

// G(8 rows of this warp x 8 NTC columns) = X(rows rb + 8 warp ..) X(rows rb ..)' over the 32 columns of a tile held in
// shared memory as t[c][row] (row stride RSX); only the n-tiles at or below the diagonal (NTC = warp + 1) are formed.
template <int NTC>
__device__ __forceinline__ void kt_syrk32(double (&g)[4][2], const double* t, int RSX, int rb, int warp, int fr, int fk) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) g[nt][0] = g[nt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < KT_NB / 4; ++ks) {
    const double* col = t + (4 * ks + fk) * RSX + rb;
    const double av = col[8 * warp + fr];
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) dmma884(g[nt][0], g[nt][1], av, col[8 * nt + fr]);
  }
}
__device__ __forceinline__ void kt_syrk32_dispatch(double (&g)[4][2], const double* t, int RSX, int rb, int warp, int fr, int fk) {
  switch (warp) {
    case 0: kt_syrk32<1>(g, t, RSX, rb, warp, fr, fk); break;
    case 1: kt_syrk32<2>(g, t, RSX, rb, warp, fr, fk); break;
    case 2: kt_syrk32<3>(g, t, RSX, rb, warp, fr, fk); break;
    default: kt_syrk32<4>(g, t, RSX, rb, warp, fr, fk); break;
  }
}

struct KtCtx {
  double *panel, *aux, *linvT, *dn0, *dn1, *rsbuf, *colbuf;
  int* colb;
  int* s_avail;
  int tid, lane, warp, fr, fk;
};

// ================================ SUB(eq, k, [i0, i1)) ================================
__device__ __forceinline__ void kt_sub(const KtArgs& a, const KtCtx& x, int eq, int k, int i0, int i1) {
  const K2Params& p = a.p;
  const int n = p.n, nbk = a.nbk, nrb = a.nrb;
  const int tid = x.tid, lane = x.lane, warp = x.warp, fr = x.fr, fk = x.fk;
  double* panel = x.panel; double* linvT = x.linvT;
  const int* colb = x.colb;
  double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  int* prog = a.progress + (size_t)eq * nrb;
  const double* lwsk = a.lws + ((size_t)eq * nbk + k) * (KT_NB * KT_NB);
  const int c0 = k * KT_NB;
  const int w = min(KT_NB, n - c0);
  const int r_lo = kt_rowstart(i0, nbk, n), r_hi = kt_rowend(i1 - 1, nbk, n);
  const int R = r_hi - r_lo;
  const int mtiles = (R + 7) >> 3;
  const int nq = mtiles > warp ? (mtiles - warp + KT_WARPS - 1) / KT_WARPS : 0;   // m-tiles warp, warp+4, ..
  int arow[KT_MAXQ], brow[4];
#pragma unroll
  for (int q = 0; q < KT_MAXQ; ++q) arow[q] = min(r_lo + 8 * (warp + KT_WARPS * q) + fr, n);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) brow[nt] = min(c0 + 8 * nt + fr, n);
  // the accumulators start at -A (the tile's original entries, in flight while the first dependencies are polled), so the
  // history product lands on top of them and the tile is  C = -(acc)  without a second pass over global memory
  double acc[KT_MAXQ][4][2];
#pragma unroll
  for (int q = 0; q < KT_MAXQ; ++q) {
    const int li = 8 * (warp + KT_WARPS * q) + fr;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * nt + 2 * fk + e;
        acc[q][nt][e] = (li < R && c < w) ? -__ldcg(om + colb[c0 + c] + r_lo + li) : 0.0;
      }
  }
  // the nearest far tile also forms the look-ahead product for the chain (see kt_acc)
  const bool pa = (i0 == k + 2);   // (k = 0: no history yet, but D_2 still gets this block column's term)
  int brow2[4], brow3[4];
  double acc2[4][2], acc3[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    brow2[nt] = min(c0 + KT_NB + 8 * nt + fr, n);
    brow3[nt] = min(r_lo + 8 * nt + fr, n);
    acc2[nt][0] = acc2[nt][1] = acc3[nt][0] = acc3[nt][1] = 0.0;
  }

  const bool profb = p.prof != nullptr && eq == 0 && pa && tid == 0;   // the look-ahead tasks of equation 0
  long long tb = profb ? clock64() : 0;
#define KB_TICK(slot) do { if (profb) { const long long tn = clock64(); p.prof[slot] += tn - tb; tb = tn; } } while (0)
  if (profb) p.prof[39] += 1;
  // ---- history: block columns b < k, consumed as they become final for row block k (the B operand) and for my rows
  int bdone = 0;
  while (bdone < k) {
    if (warp == 0) {
      // one request per poll: lane I reads the counter of row block I (the equation's counters share a cache line) and the
      // warp takes the minimum over the ones this task depends on; waiting CTAs back off (hundreds of them polled the
      // same few lines every ~50 cycles: the L2 slices holding them were saturated and every hop of the chain paid for it)
      int av;
      unsigned ns = 32;
      for (;;) {
        int v = 0x7fffffff;
        for (int I = lane; I < nrb; I += 32)
          if (I == k || (pa && I == k + 1) || (I >= i0 && I < i1)) v = min(v, kt_ld_acquire(prog + I));
        av = min(__reduce_min_sync(0xffffffffu, v), k);
        if (av > bdone) break;
        __nanosleep(ns);
        ns = min(ns * 2, 1024u);
      }
      if (lane == 0) { kt_acquire(); *x.s_avail = av; }
    }
    __syncthreads();
    const int avail = *x.s_avail;
    __syncthreads();
    KB_TICK(32);
    const int cb = bdone * KT_NB, ce = avail * KT_NB;
    switch (nq) {
      case 1: kt_acc_any<1>(pa, acc, acc2, acc3, om, colb, cb, ce, arow, brow, brow2, brow3, fk); break;
      case 2: kt_acc_any<2>(pa, acc, acc2, acc3, om, colb, cb, ce, arow, brow, brow2, brow3, fk); break;
      case 3: kt_acc_any<3>(pa, acc, acc2, acc3, om, colb, cb, ce, arow, brow, brow2, brow3, fk); break;
      case 4: kt_acc_any<4>(pa, acc, acc2, acc3, om, colb, cb, ce, arow, brow, brow2, brow3, fk); break;
      default: break;
    }
    bdone = avail;
    KB_TICK(33);
  }
  if (pa) {
    // P(k+2, k+1) over block columns < k: out to the chain of this equation (rows 8 warp .. of the first row block)
    double* Pw = a.P + ((size_t)eq * nbk + (k + 2)) * (KT_NB * KT_NB);
    if (nq > 0) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) __stcg(Pw + (8 * nt + 2 * fk + e) * KT_NB + 8 * warp + fr, acc2[nt][e]);
    }
      __syncthreads();
    if (tid == 0) kt_st_release(a.pprogress + (size_t)eq * nrb + k + 2, 1);
  }

  // ---- C = A - history = -acc into the panel (rows beyond the tile and columns beyond the block are zero)
#pragma unroll
  for (int q = 0; q < KT_MAXQ; ++q) {
    const int mt = warp + KT_WARPS * q;
    if (mt < mtiles) {
      const int li = 8 * mt + fr;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 8 * nt + 2 * fk + e;
          panel[c * KT_RS + li] = (li < R && c < w) ? -acc[q][nt][e] : 0.0;
        }
    }
  }
  KB_TICK(34);
  // ---- the inverse of the diagonal block (DIAG(eq, k))
  if (tid == 0) { unsigned ns = 20; while (kt_ld_acquire(prog + k) < k + 1) { __nanosleep(ns); ns = min(ns * 2, 320u); } kt_acquire(); }
  __syncthreads();
  KB_TICK(35);
  for (int idx = tid; idx < KT_NB * KT_NB; idx += KT_THREADS) linvT[(idx >> 5) * KT_LS + (idx & 31)] = __ldcg(lwsk + idx);   // row-major
  __syncthreads();
  // ---- X = C Linv'  (DMMA, skipping the zero triangle); every warp works on its own rows only
  for (int mt = warp; mt < mtiles; mt += KT_WARPS) {
    double xt[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) xt[nt][0] = xt[nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KT_NB / 4; ++ks) {
      const double av = panel[(4 * ks + fk) * KT_RS + 8 * mt + fr];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (4 * ks <= 8 * nt + 7) dmma884(xt[nt][0], xt[nt][1], av, linvT[(8 * nt + fr) * KT_LS + 4 * ks + fk]);
      }
    }
    __syncwarp();
    const int li = 8 * mt + fr;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * nt + 2 * fk + e;
        panel[c * KT_RS + li] = (li < R && c < w) ? xt[nt][e] : 0.0;
      }
  }
  __syncthreads();
  // ---- store the tile
  for (int c = warp; c < w; c += KT_WARPS) {
    double* dst = om + colb[c0 + c] + r_lo;
    const double* src = panel + c * KT_RS;
    for (int li = lane; li < R; li += 32) dst[li] = src[li];
  }
  __syncthreads();
  if (tid == 0) {
    kt_acquire();                                    // one gpu-scope fence for the whole tile, then plain flags
    for (int I = i0; I < i1; ++I) kt_st_relaxed(prog + I, k + 1);
  }
  KB_TICK(36);
  // ---- the diagonal accumulator of my first row block (look-ahead task only): D_{k+2} = -(sum_{b < k} + this block column)
  if (pa && i0 < nbk) {
    double g[4][2];
    kt_syrk32_dispatch(g, panel, KT_RS, 0, warp, fr, fk);
    double* DI = a.D + ((size_t)eq * nbk + i0) * (KT_NB * KT_NB);
    const int r = 8 * warp + fr;
    if (nq > 0) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 8 * nt + 2 * fk + e;
          if (nt <= warp && c <= r) __stcg(DI + c * KT_NB + r, -(acc3[nt][e] + g[nt][e]));
        }
    }
      __syncthreads();
    if (tid == 0) kt_st_release(a.dprogress + (size_t)eq * nrb + i0, 1);
  }
  KB_TICK(38);
#undef KB_TICK
}

// ================================ one step of an equation's diagonal chain ================================
// DIAG(eq, k) and the tile below it, SUB(eq, k, row block k+1), fused.  While warp 0 factorises and inverts the
// diagonal block (a 32-step dependent chain), warps 1-3 fetch the tile's original entries and form its history product
// (each a third of the block columns, all 32 rows); then warp 0 publishes the block and its inverse while warps 1-3 do
// the triangular solve with Linv straight from shared memory.  The solved tile stays in shared memory (`aux`): it is the
// newest block column's term of the NEXT diagonal block, so a CTA that walks a single equation (`carry`) never reads it
// back.  Same arithmetic as the separate tasks except that the history of the tile is summed in three parts (fixed order).
__device__ __forceinline__ void kt_chain_step(const KtArgs& a, const KtCtx& x, int eq, int k, bool carry) {
  const K2Params& p = a.p;
  const int n = p.n, nbk = a.nbk, nrb = a.nrb;
  const int tid = x.tid, lane = x.lane, warp = x.warp, fr = x.fr, fk = x.fk;
  double* panel = x.panel; double* aux = x.aux; double* linvT = x.linvT; double* rsbuf = x.rsbuf; double* colbuf = x.colbuf;
  const int* colb = x.colb;
  double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  int* prog = a.progress + (size_t)eq * nrb;
  double* lwsk = a.lws + ((size_t)eq * nbk + k) * (KT_NB * KT_NB);
  const int c0 = k * KT_NB;
  const int w = min(KT_NB, n - c0);
  const int r_lo = kt_rowstart(k + 1, nbk, n), r_hi = kt_rowend(k + 1, nbk, n);   // the tile below: row block k+1
  const int R = r_hi - r_lo;
  const bool prof0 = p.prof != nullptr && blockIdx.x == 0 && tid == 0, prof1 = p.prof != nullptr && blockIdx.x == 0 && tid == 32;
  long long tp = (prof0 || prof1) ? clock64() : 0;
#define KT_TICK(cond, slot) do { if (cond) { const long long tn = clock64(); p.prof[slot] += tn - tp; tp = tn; } } while (0)
  if (k > 0) {
    if (tid == 0) {
      if (!carry) while (kt_ld_acquire(prog + k) < k) __nanosleep(20);   // (carry: this CTA released it itself one step ago)
      if (k >= 2) while (kt_ld_acquire(a.dprogress + (size_t)eq * nrb + k) < 1) __nanosleep(20);
      kt_acquire();
    }
    __syncthreads();
  }
  KT_TICK(prof0, 0);
  {
    // diagonal block A_kk + D_k (and L(k,k-1) unless carried over): warp w takes columns w, w+4, ..; every load of the
    // eight columns is issued before the first use
    const double* Dk = a.D + ((size_t)eq * nbk + k) * (KT_NB * KT_NB);
    double va[8], vd[8], vx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = warp + KT_WARPS * i;
      const bool in = c < w && lane >= c && lane < w;
      va[i] = in ? __ldcg(om + colb[min(c0 + c, n - 1)] + c0 + lane) : 0.0;
      vd[i] = (in && k >= 2) ? __ldcg(Dk + c * KT_NB + lane) : 0.0;
      vx[i] = (k > 0 && !carry && lane < w) ? __ldcg(om + colb[c0 - KT_NB + c] + c0 + lane) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = warp + KT_WARPS * i;
      panel[c * KT_RS + lane] = (c < w) ? va[i] + vd[i] : ((lane == c) ? 1.0 : 0.0);   // identity padding of a partial block
      if (k > 0 && !carry) aux[c * KT_LS + lane] = vx[i];
    }
  }
  __syncthreads();
  KT_TICK(prof0, 1);
  if (k > 0) {
    // the newest block column's term: - L(k,k-1) L(k,k-1)'  (older ones arrived through D)
    double g[4][2];
    kt_syrk32_dispatch(g, aux, KT_LS, 0, warp, fr, fk);
    const int r = 8 * warp + fr;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * nt + 2 * fk + e;
        if (nt <= warp && c <= r) panel[c * KT_RS + r] -= g[nt][e];
      }
    __syncthreads();
  }
  KT_TICK(prof0, 2);
  if (prof1) tp = clock64();
  if (warp == 0) {
    int fail = 0;
    double av[KT_NB];     // row `lane` of the factor
#pragma unroll
    for (int c = 0; c < KT_NB; ++c) av[c] = (c <= lane) ? panel[c * KT_RS + lane] : 0.0;
    double xx = __shfl_sync(0xffffffffu, av[0], 0);
    if (!(xx > 0.0) || !isfinite(xx)) { if (fail == 0) fail = c0 + 1; xx = 1.0; }
    double rs = rsqrt(xx);
#pragma unroll
    for (int c = 0; c < KT_NB; ++c) {
      const double l = (lane >= c) ? av[c] * rs : 0.0;
      av[c] = l;
      double* cb = colbuf + (c & 1) * 32;
      cb[lane] = l;
      if (lane == c) rsbuf[c] = rs;
      __syncwarp();
      if (c + 1 < KT_NB) {
        av[c + 1] = fma(-l, cb[c + 1], av[c + 1]);
        xx = __shfl_sync(0xffffffffu, av[c + 1], c + 1);
        if (!(xx > 0.0) || !isfinite(xx)) { if (fail == 0) fail = c0 + c + 2; xx = 1.0; }
        rs = rsqrt(xx);
      }
#pragma unroll
      for (int c2 = c + 2; c2 < KT_NB; ++c2) av[c2] = fma(-l, cb[c2], av[c2]);
    }
#pragma unroll
    for (int c = 0; c < KT_NB; ++c) panel[c * KT_RS + lane] = av[c];
    __syncwarp();
    KT_TICK(prof0, 3);
    double sacc[KT_NB];
#pragma unroll
    for (int i = 0; i < KT_NB; ++i) sacc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int kk = 0; kk < KT_NB; ++kk) {
      const double xk = (kk >= lane) ? sacc[kk] * rsbuf[kk] : 0.0;
      sacc[kk] = xk;
#pragma unroll
      for (int i = kk + 1; i < KT_NB; ++i) sacc[i] = fma(-panel[kk * KT_RS + i], xk, sacc[i]);
    }
#pragma unroll
    for (int i = 0; i < KT_NB; ++i) linvT[i * KT_LS + lane] = sacc[i];   // row-major Linv[i][j], lane j (conflict-free both ways)
    if (lane == 0 && fail) atomicCAS(&p.status[eq], 0, fail);
    KT_TICK(prof0, 4);
  } else {
    // ---- warps 1-3: original entries of the tile below -> aux is still needed by nobody else (the syrk above is done),
    // but it becomes the solved tile at the end, so the originals go to panel rows 32.. of a scratch column block:
    // panel[c][96 + li] is free (parts use rows 32 w .. 32 w + 31 for w = 1..3 -> rows 32..127; rows 0..31 the factor).
    // They are kept in registers instead: 1024 entries over 96 threads, issued now, consumed after the history.
    // (`aux` is free between the syrk above and the solved tile written at the end: asynchronous 8-byte copies, no registers)
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      const int idx = (tid - 32) + 96 * i;
      const int c = idx >> 5, li = idx & 31;
      if (idx < KT_NB * KT_NB && li < R && c < w) cp_async8(aux + c * KT_LS + li, om + colb[c0 + c] + r_lo + li);
    }
    cp_async_commit();
    // history of the tile below: sum_{b < k} L(k+1, b) L(k, b)'
    double pold[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) pold[i] = 0.0;
    double acc[KT_MAXQ][4][2];
#pragma unroll
    for (int q = 0; q < KT_MAXQ; ++q)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;
    if (k > 0) {
      int arow[KT_MAXQ], brow[4];
#pragma unroll
      for (int q = 0; q < KT_MAXQ; ++q) arow[q] = min(r_lo + 8 * q + fr, n);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) brow[nt] = min(c0 + 8 * nt + fr, n);
      // the old part of the history (block columns b <= k-2) arrives as P from the bulk task that owned these rows one
      // block column ago; only the newest block column (b = k-1, a bulk tile) is formed here, a third per warp
      if (tid == 32) {
        if (k >= 2) while (kt_ld_acquire(a.pprogress + (size_t)eq * nrb + k + 1) < 1) __nanosleep(20);
        while (kt_ld_acquire(prog + k + 1) < k) __nanosleep(20);
        kt_acquire();
      }
      asm volatile("bar.sync 1, 96;" ::: "memory");
      KT_TICK(prof1, 13);
      if (k >= 2) {
        const double* Pw = a.P + ((size_t)eq * nbk + (k + 1)) * (KT_NB * KT_NB);
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const int idx = (tid - 32) + 96 * i;
          pold[i] = (idx < KT_NB * KT_NB) ? __ldcg(Pw + idx) : 0.0;     // idx = c * 32 + li
        }
      }
      {
        const int s0 = 8 * (k - 1);
        const int sb = s0 + (warp == 1 ? 0 : (warp == 2 ? 3 : 6)), se = s0 + (warp == 1 ? 3 : (warp == 2 ? 6 : 8));
        kt_acc_chain(acc, om, colb, sb, se, arow, brow, fk);
      }
    }
    KT_TICK(prof1, 14);
    double* part = panel + 32 * warp;   // rows [32 warp, 32 warp + 32) of every panel column
#pragma unroll
    for (int q = 0; q < KT_MAXQ; ++q)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) part[(8 * nt + 2 * fk + e) * KT_RS + 8 * q + fr] = acc[q][nt][e];
    cp_async_wait<0>();
    asm volatile("bar.sync 1, 96;" ::: "memory");
    // C = A - (part 1 + part 2 + part 3) -> panel rows 32..63 (each thread sums its own entries)
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      const int idx = (tid - 32) + 96 * i;
      if (idx < KT_NB * KT_NB) {
        const int c = idx >> 5, li = idx & 31;
        double* pc = panel + c * KT_RS + li;
        pc[32] = (li < R && c < w) ? aux[c * KT_LS + li] - (((pold[i] + pc[32]) + pc[64]) + pc[96]) : 0.0;
      }
    }
  }
  __syncthreads();     // Linv (warp 0) and C (warps 1-3) are in shared memory
  KT_TICK(prof0, 6);
  // the diagonal block and its inverse go out (all threads, from shared memory); they are published with the tile below
  // after ONE fence at the end of the step
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = warp + KT_WARPS * i;
    if (c < w && lane >= c && lane < w) om[colb[c0 + c] + c0 + lane] = panel[c * KT_RS + lane];
    lwsk[c * KT_NB + lane] = linvT[c * KT_LS + lane];
  }
  {
    // X = C Linv' : warp w takes rows 8 w .. 8 w + 7; the solved tile goes to global memory and to `aux`
    double xt[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) xt[nt][0] = xt[nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KT_NB / 4; ++ks) {
      const double avv = panel[(4 * ks + fk) * KT_RS + 32 + 8 * warp + fr];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (4 * ks <= 8 * nt + 7) dmma884(xt[nt][0], xt[nt][1], avv, linvT[(8 * nt + fr) * KT_LS + 4 * ks + fk]);
      }
    }
    __syncthreads();   // every warp has read the originals in `aux` (sum) and its C rows before `aux` is overwritten
    const int li = 8 * warp + fr;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * nt + 2 * fk + e;
        const bool in = li < R && c < w;
        if (in) om[colb[c0 + c] + r_lo + li] = xt[nt][e];
        aux[c * KT_LS + li] = in ? xt[nt][e] : 0.0;      // next step's L(k+1, k) (rows beyond the block: zero)
      }
  }
  __syncthreads();
  KT_TICK(prof0, 8);
  if (tid == 0) { kt_st_release(prog + k, k + 1); kt_st_release(prog + k + 1, k + 1); }
  KT_TICK(prof0, 9);
#undef KT_TICK
}

// ================================ an equation's diagonal chain, warp-specialised ================================
// One chain CTA walks ONE equation (the launch gives every equation its own chain CTA).  Per block column k the only
// strictly sequential work is the 32 pivots of the diagonal block - everything else is arranged around them:
//   warp 0  factorises the diagonal block, one column at a time, and announces every finished column on a shared-memory
//           mbarrier; the next pivot never leaves the registers of its lane (no shared-memory round trip);
//   warp 2  inverts the block by substitution, one column behind warp 0 (the triangular solves below, the far tiles and
//           the back-solve all use Linv);
//   warps 1, 3  form C = A(k+1,k) - P - L(k+1,k-1) L(k,k-1)' as soon as the far tile L(k+1,k-1) of the previous block
//           column has arrived (operands straight from L2 into DMMA fragments), then X = C Linv' (DMMA), 16 rows each;
//   warp 3  also polls the counters, publishes the step (one gpu-scope fence, two flags) and stages the NEXT diagonal
//           block (original entries + accumulator D) while warps 0-2 wait for nothing but that.
// The solved tile L(k+1,k) stays in shared memory (`aux`): it is the newest term of the next diagonal block.
constexpr int KT_BAR_DN = 1, KT_BAR_S = 2, KT_BAR_C = 3, KT_BAR_L = 4, KT_BAR_PUB = 5;
__device__ __forceinline__ void kt_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void kt_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void kt_chain_run(const KtArgs& a, const KtCtx& x, int eq, uint64_t* colbar /* [32], count 1 */, long long* sprof /* [24] shared */) {
  const K2Params& p = a.p;
  const int n = p.n, nbk = a.nbk, nrb = a.nrb;
  const int tid = x.tid, lane = x.lane, warp = x.warp, fr = x.fr, fk = x.fk;
  double* Lc = x.panel;                       // [32][KT_RS], rows 0..31: columns of the factor (+ identity padding)
  double* cb = x.panel + 32;                  //              rows 32..63: C, the tile below before its solve
  double* aux = x.aux;                        // [32][KT_LS]  L(k, k-1), then L(k+1, k)
  double* lv = x.linvT;                       // [32][KT_LS]  Linv of the diagonal block, row-major
  double* dnb[2] = {x.dn0, x.dn1};            // [32][KT_LS]  diagonal block k (even / odd k): A_kk + D_k
  double* rsbuf = x.rsbuf;
  const int* colb = x.colb;
  double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  int* prog = a.progress + (size_t)eq * nrb;
  int* dprog = a.dprogress + (size_t)eq * nrb;
  int* pprog = a.pprogress + (size_t)eq * nrb;
  const bool prof0 = p.prof != nullptr && blockIdx.x == 0 && tid == 0, prof1 = p.prof != nullptr && blockIdx.x == 0 && tid == 32;
  const bool prof3 = p.prof != nullptr && blockIdx.x == 0 && tid == 96;
  long long tp = (prof0 || prof1 || prof3) ? clock64() : 0;
  // (phase cycles accumulate in shared memory - a global read-modify-write per tick cost an L2 round trip itself)
#define KT_TICK(cond, slot) do { if (cond) { const long long tn = clock64(); sprof[slot] += tn - tp; tp = tn; } } while (0)

  // diagonal block kk into dst: lower triangle of A + D, identity padding of a partial block (one warp)
  auto load_diag = [&](int kk, double* dst) {
    const int c0 = kk * KT_NB, w = min(KT_NB, n - c0);
    const double* Dk = a.D + ((size_t)eq * nbk + kk) * (KT_NB * KT_NB);
    double va[KT_NB], vd[KT_NB];
#pragma unroll
    for (int c = 0; c < KT_NB; ++c) {
      const bool in = c < w && lane >= c && lane < w;
      va[c] = in ? __ldcg(om + colb[min(c0 + c, n - 1)] + c0 + lane) : 0.0;
      vd[c] = (in && kk >= 2) ? __ldcg(Dk + c * KT_NB + lane) : 0.0;
    }
    KT_TICK(prof3, 22);
    if (prof3 && va[0] == 12345.678) sprof[23] += 1;   // (forces the first load to have landed)
    KT_TICK(prof3, 23);
#pragma unroll
    for (int c = 0; c < KT_NB; ++c) dst[c * KT_LS + lane] = (c < w) ? va[c] + vd[c] : ((lane == c) ? 1.0 : 0.0);
  };

  if (warp == 3) {
    load_diag(0, dnb[0]);
    kt_bar_arrive(KT_BAR_DN, KT_THREADS);
  }
  for (int k = 0; k < nbk; ++k) {
    const int c0 = k * KT_NB;
    const int w = min(KT_NB, n - c0);
    const int r_lo = kt_rowstart(k + 1, nbk, n), r_hi = kt_rowend(k + 1, nbk, n);   // the tile below: row block k+1
    const int R = r_hi - r_lo;
    double* dn = dnb[k & 1];
    const uint32_t par = (uint32_t)(k & 1);   // every column barrier completes one phase per block column
    if (warp < 3) {
      kt_bar_sync(KT_BAR_DN, KT_THREADS);     // diagonal block staged (warp 3), L(k,k-1) in aux (warps 1 and 3, previous step)
      KT_TICK(prof0, 0);
      if (k > 0) {
        // the newest block column's term: - L(k,k-1) L(k,k-1)'.  m-tiles: warp 0 -> 0 and 1, warp 1 -> 2, warp 2 -> 3
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int mt = (warp == 0) ? h : warp + 1;
          if (warp == 0 || h == 0) {
            double g[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) g[nt][0] = g[nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KT_NB / 4; ++ks) {
              const double* col = aux + (4 * ks + fk) * KT_LS;
              const double av = col[8 * mt + fr];
#pragma unroll
              for (int nt = 0; nt < 4; ++nt)
                if (nt <= mt) dmma884(g[nt][0], g[nt][1], av, col[8 * nt + fr]);
            }
            const int r = 8 * mt + fr;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = 8 * nt + 2 * fk + e;
                if (nt <= mt && c <= r) dn[c * KT_LS + r] -= g[nt][e];
              }
          }
        }
      }
      kt_bar_sync(KT_BAR_S, 96);
      KT_TICK(prof0, 2);
    }
    if (warp == 0) {
      // ---- factorisation of the diagonal block: lane = row
      int fail = 0;
      double av[KT_NB];
#pragma unroll
      for (int c = 0; c < KT_NB; ++c) av[c] = (c <= lane) ? dn[c * KT_LS + lane] : 0.0;
      double xx = __shfl_sync(0xffffffffu, av[0], 0);
      if (!(xx > 0.0) || !isfinite(xx)) { fail = c0 + 1; xx = 1.0; }
      double rs = rsqrt(xx);
#pragma unroll
      for (int c = 0; c < KT_NB; ++c) {
        const double l = (lane >= c) ? av[c] * rs : 0.0;
        av[c] = l;
        Lc[c * KT_RS + lane] = l;
        if (lane == c) rsbuf[c] = rs;
        if (c + 1 < KT_NB) {
          const double dpiv = fma(-l, l, av[c + 1]);          // lane c+1: its own pivot, straight from registers
          xx = __shfl_sync(0xffffffffu, dpiv, c + 1);
          if (!(xx > 0.0) || !isfinite(xx)) { if (fail == 0) fail = c0 + c + 2; xx = 1.0; }
          rs = rsqrt(xx);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&colbar[c]);    // (release.cta: the column and rsbuf[c] are visible to whoever sees the phase flip)
#pragma unroll
        for (int c2 = c + 1; c2 < KT_NB; ++c2) av[c2] = fma(-l, Lc[c * KT_RS + c2], av[c2]);
      }
      KT_TICK(prof0, 3);
#pragma unroll
      for (int c = 0; c < KT_NB; ++c)
        if (c < w && lane >= c && lane < w) om[colb[c0 + c] + c0 + lane] = av[c];
      if (lane == 0 && fail) atomicCAS(&p.status[eq], 0, fail);
      kt_bar_arrive(KT_BAR_PUB, KT_THREADS);
      KT_TICK(prof0, 8);
    } else if (warp == 2) {
      // ---- inverse of the diagonal block, lane = column of Linv, one column behind the factorisation
      double* lwsk = a.lws + ((size_t)eq * nbk + k) * (KT_NB * KT_NB);
      double sacc[KT_NB];
#pragma unroll
      for (int i = 0; i < KT_NB; ++i) sacc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int kk = 0; kk < KT_NB; ++kk) {
        mbar_wait(&colbar[kk], par);
        const double xk = (kk >= lane) ? sacc[kk] * rsbuf[kk] : 0.0;
        sacc[kk] = xk;
#pragma unroll
        for (int i = kk + 1; i < KT_NB; ++i) sacc[i] = fma(-Lc[kk * KT_RS + i], xk, sacc[i]);
      }
#pragma unroll
      for (int i = 0; i < KT_NB; ++i) lv[i * KT_LS + lane] = sacc[i];       // row-major Linv[i][j], lane j
      kt_bar_arrive(KT_BAR_L, 96);
#pragma unroll
      for (int i = 0; i < KT_NB; ++i) lwsk[i * KT_NB + lane] = sacc[i];
      kt_bar_arrive(KT_BAR_PUB, KT_THREADS);
    } else {
      // ---- warps 1 and 3: the tile below.  Warp 3 polls for the far tile of the previous block column first.
      if (warp == 3) {
        if (k > 0 && lane == 0) {
          if (k >= 2) while (kt_ld_acquire(pprog + k + 1) < 1) { }
          while (kt_ld_acquire(prog + k + 1) < k) { }
          kt_acquire();
        }
        __syncwarp();
        KT_TICK(prof3, 16);
      } else {
        KT_TICK(prof1, 12);
      }
      kt_bar_sync(KT_BAR_C, 64);
      KT_TICK(prof1, 13);
      const int mt0 = (warp == 1) ? 0 : 2;
      // C = A - P - L(k+1,k-1) L(k,k-1)' for m-tiles mt0, mt0+1: every operand straight from L2 into its fragment
      const double* Pw = a.P + ((size_t)eq * nbk + (k + 1)) * (KT_NB * KT_NB);
      double ao[2][KT_NB / 4], co[2][4][2], po[2][4][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // every load is unconditional on a clamped (always valid) address and masked afterwards: predicated loads made
        // the compiler split them into branchy groups with the DMMAs in between - several L2 round trips instead of one
        const int li = 8 * (mt0 + h) + fr;
        const int lic = r_lo + min(li, R - 1);
        const int cprev = max(c0 - KT_NB, 0);
#pragma unroll
        for (int ks = 0; ks < KT_NB / 4; ++ks) ao[h][ks] = __ldcg(om + colb[cprev + 4 * ks + fk] + lic);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * fk + e;
            co[h][nt][e] = __ldcg(om + colb[min(c0 + c, n - 1)] + lic);
            po[h][nt][e] = __ldcg(Pw + c * KT_NB + li);
          }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool rin = 8 * (mt0 + h) + fr < R;
#pragma unroll
        for (int ks = 0; ks < KT_NB / 4; ++ks) ao[h][ks] = (k >= 1 && rin) ? ao[h][ks] : 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) po[h][nt][e] = (k >= 2) ? po[h][nt][e] : 0.0;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double g[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) g[nt][0] = g[nt][1] = 0.0;
        if (k > 0) {
#pragma unroll
          for (int ks = 0; ks < KT_NB / 4; ++ks) {
            const double* bcol = aux + (4 * ks + fk) * KT_LS;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(g[nt][0], g[nt][1], ao[h][ks], bcol[8 * nt + fr]);
          }
        }
        const int li = 8 * (mt0 + h) + fr;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * fk + e;
            cb[c * KT_RS + li] = (li < R && c < w) ? co[h][nt][e] - (po[h][nt][e] + g[nt][e]) : 0.0;
          }
      }
      KT_TICK(prof1, 14);
      kt_bar_sync(KT_BAR_L, 96);              // Linv in shared memory (warp 2), every read of the old aux done
      // X = C Linv'  (skipping the zero triangle)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mt = mt0 + h;
        double xt[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) xt[nt][0] = xt[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KT_NB / 4; ++ks) {
          const double avv = cb[(4 * ks + fk) * KT_RS + 8 * mt + fr];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            if (4 * ks <= 8 * nt + 7) dmma884(xt[nt][0], xt[nt][1], avv, lv[(8 * nt + fr) * KT_LS + 4 * ks + fk]);
        }
        const int li = 8 * mt + fr;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * fk + e;
            const bool in = li < R && c < w;
            if (in) om[colb[c0 + c] + r_lo + li] = xt[nt][e];
            aux[c * KT_LS + li] = in ? xt[nt][e] : 0.0;      // next step's L(k+1, k)
          }
      }
      KT_TICK(prof1, 15);
      if (warp == 1) {
        kt_bar_arrive(KT_BAR_PUB, KT_THREADS);
      } else {
        KT_TICK(prof3, 17);
        kt_bar_sync(KT_BAR_PUB, KT_THREADS);    // warps 0-2 have issued every store of this step
        if (lane == 0) {
          kt_acquire();                         // (fence.acq_rel.gpu: one fence, then two relaxed flags)
          kt_st_relaxed(prog + k, k + 1);
          kt_st_relaxed(prog + k + 1, k + 1);
        }
        __syncwarp();
        KT_TICK(prof3, 18);
        if (k + 1 < nbk) {
          if (k + 1 >= 2 && lane == 0) {
            while (kt_ld_acquire(dprog + k + 1) < 1) { }
            kt_acquire();
          }
          __syncwarp();
          KT_TICK(prof3, 19);
          if (prof3) {   // probe: one dependent L2 read (a word another SM wrote a while ago)
            const long long t0 = clock64();
            const int v = kt_ld_acquire(dprog + k + 1);
            if (v < 0) sprof[23] += 1;
            const long long t1 = clock64();
            sprof[20] += t1 - t0;
            tp = clock64();
          }
          load_diag(k + 1, dnb[(k + 1) & 1]);
          KT_TICK(prof3, 21);
          kt_bar_arrive(KT_BAR_DN, KT_THREADS);
        }
      }
    }
  }
  __syncthreads();
  if (p.prof != nullptr && blockIdx.x == 0 && tid < 24) p.prof[tid] += sprof[tid];
#undef KT_TICK
}

// Scheduling.  The diagonal chain of an equation - DIAG(e,0), its first sub-diagonal tile, DIAG(e,1), .. - is latency:
// 13 sequential steps however many SMs there are.  The first `ncrit` CTAs own these chains statically (equation e
// belongs to CTA e mod ncrit, which walks the block columns in order), so a chain step never queues behind bulk work;
// every other CTA - and the chain CTAs once their chains are done - pulls the far tiles (row blocks >= k+2 of block
// column k, in stage order) from one atomic counter.  Chain tasks wait for bulk tasks of the previous block column
// only and bulk tasks for chain tasks of their own or earlier block columns, both classes are consumed in order, so
// by induction over the block column nothing can wait forever as long as the grid is co-resident (it is sized by the
// occupancy query) and has at least one CTA of each class.
__global__ void __launch_bounds__(KT_THREADS, KT_CTAS) k2_tiled_kernel(const KtArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task, s_avail;
  __shared__ __align__(8) uint64_t s_colbar[KT_NB];
  __shared__ long long s_prof[24];
  KtCtx x;
  x.panel = smem;                                        // [32][KT_RS]  panel[c][local row]
  x.aux = x.panel + KT_NB * KT_RS;                       // [32][KT_LS]  L(k, k-1) / the originals of the tile below
  x.linvT = x.aux + KT_NB * KT_LS;                       // [32][KT_LS]  Linv of the diagonal block, row-major
  x.dn0 = x.linvT + KT_NB * KT_LS;                       // [2][32][KT_LS]  diagonal-block staging of the chain (even / odd block column)
  x.dn1 = x.dn0 + KT_NB * KT_LS;
  x.rsbuf = x.dn1 + KT_NB * KT_LS;                       // [32]
  x.colbuf = x.rsbuf + 32;                               // [2][32]
  x.colb = reinterpret_cast<int*>(x.colbuf + 64);        // [n]
  x.s_avail = &s_avail;
  x.tid = threadIdx.x; x.lane = x.tid & 31; x.warp = x.tid >> 5; x.fr = x.lane >> 2; x.fk = x.lane & 3;
  const K2Params& p = a.p;
  const int tid = x.tid;
  if (tid == 0) bvb_stamp_begin(p.tstamp);
  for (int i = tid; i < p.n; i += KT_THREADS) x.colb[i] = p.colbase[i];
  __syncthreads();

  if ((int)blockIdx.x < a.ncrit) {
    if (p.M <= a.ncrit) {                // one equation per chain CTA: the warp-specialised walk
      if (tid < KT_NB) mbar_init(&s_colbar[tid], 1);
      if (tid < 24) s_prof[tid] = 0;
      mbar_fence_init();
      __syncthreads();
      kt_chain_run(a, x, blockIdx.x, s_colbar, s_prof);
    } else {
      for (int k = 0; k < a.nbk; ++k)
        for (int e = blockIdx.x; e < p.M; e += a.ncrit) kt_chain_step(a, x, e, k, false);
    }
  }
  for (;;) {
    if (tid == 0) s_task = atomicAdd(a.head, 1);
    __syncthreads();
    const int ti = s_task;
    __syncthreads();
    if (ti >= a.ntasks) break;
    const KtTask t = a.tasks[ti];
    kt_sub(a, x, t.eq, t.k, t.i0, t.i1);
  }
  if (tid == 0) bvb_stamp_end(p.tstamp);
}

// Back-solve  L' phi = (L^-1 b) + z  and the draw, one CTA per equation: block rows from the bottom, the saved inverses
// of the diagonal blocks instead of 32-step substitution chains (the tail of k2_cholsolve_v2 as its own kernel).
constexpr int KTB_WARPS = 8;
__global__ void __launch_bounds__(KTB_WARPS * 32) k2_tiled_backsolve(const K2Params p, const double* __restrict__ lws, int nbk) {
  extern __shared__ __align__(16) double smem[];
  double* ysol = smem;                         // [nbk * 32 + 32]
  double* sblk = ysol + nbk * KT_NB + 32;      // [32]
  int* colb = reinterpret_cast<int*>(sblk + 32);
  const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, eq = blockIdx.x;
  const double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  const double* __restrict__ lwe = lws + (size_t)eq * nbk * (KT_NB * KT_NB);
  for (int i = tid; i < n; i += KTB_WARPS * 32) colb[i] = p.colbase[i];
  __syncthreads();
  for (int i = tid; i < nbk * KT_NB + 32; i += KTB_WARPS * 32)
    ysol[i] = (i < n) ? om[colb[i] + n] + p.z[(size_t)eq * p.ldz + i] : 0.0;   // row n of L = the forward-solved rhs
  __syncthreads();
  for (int k = nbk - 1; k >= 0; --k) {
    const int c0 = k * KT_NB, w = min(KT_NB, n - c0), kb_end = c0 + w;
    for (int c = warp; c < KT_NB; c += KTB_WARPS) {
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
      const double* wsk = lwe + (size_t)k * (KT_NB * KT_NB);
      double xv0 = 0.0, xv1 = 0.0;
#pragma unroll
      for (int cp = 0; cp < KT_NB; cp += 2) {
        xv0 = fma(wsk[cp * KT_NB + lane], sblk[cp], xv0);
        xv1 = fma(wsk[(cp + 1) * KT_NB + lane], sblk[cp + 1], xv1);
      }
      if (lane < w) ysol[c0 + lane] = xv0 + xv1;
    }
    __syncthreads();
  }
  bool bad = false;
  for (int i = tid; i < n; i += KTB_WARPS * 32) {
    const double v = ysol[i];
    p.phi[(size_t)eq * p.ldphi + i] = v;
    bad |= !isfinite(v);
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (tid == 0) {
    if (anybad) atomicCAS(&p.status[eq], 0, -1);
    bvb_stamp_end(p.tstamp);
  }
}

// Fast form for n <= 416 (the headline Kp = 401): right-looking over the block rows.  Thread c keeps the 32 entries
// L(32 k .. 32 k + 31, c) of the CURRENT block row in registers and already holds the next block row's (loaded one
// step ahead, so no L2 round trip sits on the 13-step chain); all inverses of the diagonal blocks and the solution
// vector live in shared memory.  Per step: x_k = Linv_k' y_k (warp 0), then y_c -= L(k, c)' x_k for every column c < 32 k.
constexpr int KTF_THREADS = 384;
__device__ __forceinline__ void ktf_load_rows(double (&buf)[KT_NB], const double* __restrict__ om, const int* colb, int k, int n, int tid) {
  if (tid < KT_NB * k) {
    const double* src = om + colb[tid] + KT_NB * k;
    const int wk = min(KT_NB, n - KT_NB * k);
#pragma unroll
    for (int r = 0; r < KT_NB; ++r) buf[r] = (r < wk) ? __ldcg(src + r) : 0.0;
  }
}
__device__ __forceinline__ void ktf_step(const double (&cur)[KT_NB], int k, int n, const double* linv, double* y, double* xs,
                                         const K2Params& p, int eq, int tid, int lane, int warp, bool& bad) {
  const int c0 = KT_NB * k, w = min(KT_NB, n - c0);
  if (warp == 0) {
    const double* lk = linv + (size_t)k * (KT_NB * KT_NB);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int cp = 0; cp < KT_NB; cp += 4) {
      s0 = fma(lk[cp * KT_NB + lane], y[c0 + cp], s0);
      s1 = fma(lk[(cp + 1) * KT_NB + lane], y[c0 + cp + 1], s1);
      s2 = fma(lk[(cp + 2) * KT_NB + lane], y[c0 + cp + 2], s2);
      s3 = fma(lk[(cp + 3) * KT_NB + lane], y[c0 + cp + 3], s3);
    }
    const double xv = (lane < w) ? (s0 + s1) + (s2 + s3) : 0.0;
    xs[lane] = xv;
    if (lane < w) { p.phi[(size_t)eq * p.ldphi + c0 + lane] = xv; bad |= !isfinite(xv); }
  }
  __syncthreads();
  if (tid < c0) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int r = 0; r < KT_NB; r += 4) {
      s0 = fma(cur[r], xs[r], s0);
      s1 = fma(cur[r + 1], xs[r + 1], s1);
      s2 = fma(cur[r + 2], xs[r + 2], s2);
      s3 = fma(cur[r + 3], xs[r + 3], s3);
    }
    y[tid] -= (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
}
__global__ void __launch_bounds__(KTF_THREADS, 1) k2_tiled_backsolve_fast(const K2Params p, const double* __restrict__ lws, int nbk) {
  extern __shared__ __align__(16) double smem[];
  double* linv = smem;                                   // [nbk][32*32] row-major
  double* y = linv + (size_t)nbk * KT_NB * KT_NB;        // [nbk*32 + 32]
  double* xs = y + nbk * KT_NB + 32;                     // [32]
  int* colb = reinterpret_cast<int*>(xs + 32);
  const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, eq = blockIdx.x;
  const double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  const double* __restrict__ lwe = lws + (size_t)eq * nbk * (KT_NB * KT_NB);
  for (int i = tid; i < n; i += KTF_THREADS) colb[i] = p.colbase[i];
  __syncthreads();
  double bufA[KT_NB], bufB[KT_NB];
  ktf_load_rows(bufA, om, colb, nbk - 1, n, tid);
  for (int i = tid; i < nbk * KT_NB * KT_NB; i += KTF_THREADS) linv[i] = __ldcg(lwe + i);
  for (int i = tid; i < nbk * KT_NB + 32; i += KTF_THREADS)
    y[i] = (i < n) ? __ldcg(om + colb[i] + n) + p.z[(size_t)eq * p.ldz + i] : 0.0;   // row n of L = the forward-solved rhs
  __syncthreads();
  bool bad = false;
  for (int k = nbk - 1; k >= 0; k -= 2) {
    if (k > 0) ktf_load_rows(bufB, om, colb, k - 1, n, tid);
    ktf_step(bufA, k, n, linv, y, xs, p, eq, tid, lane, warp, bad);
    if (k == 0) break;
    if (k > 1) ktf_load_rows(bufA, om, colb, k - 2, n, tid);
    ktf_step(bufB, k - 1, n, linv, y, xs, p, eq, tid, lane, warp, bad);
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (tid == 0) {
    if (anybad) atomicCAS(&p.status[eq], 0, -1);
    bvb_stamp_end(p.tstamp);
  }
}

// ------------------------------------------------------------------ host side
namespace {
// Per-shape launch state (task list, queue head + progress counters, diagonal accumulators).  Entries are never
// freed or reallocated once built: a captured sweep graph keeps their device pointers, and one sweep may alternate
// between shapes (cholesky structure: the PHI systems and the U systems).  Launches of one shape on one device must not
// overlap in time (they share the counters): the library issues them on the handle's sampler stream only.
struct KtCache {
  int n = -1, M = -1;
  KtTask* tasks = nullptr;
  int ntasks = 0;
  int* ctrl = nullptr;       // [1 + 3 M nrb]: queue head, tile / D / P progress counters
  double* D = nullptr;
  double* P = nullptr;
};
constexpr int KT_CACHE = 64;
KtCache g_kt[16][KT_CACHE];
int g_kt_used[16] = {0};

void kt_build_tasks(std::vector<KtTask>& out, int n, int M) {
  const int nbk = (n + KT_NB - 1) / KT_NB;
  auto rowstart = [&](int I) { return I < nbk ? KT_NB * I : n; };
  auto rowend = [&](int I) { return I < nbk ? std::min(KT_NB * I + KT_NB, n) : n + 1; };
  // far tiles of block column k: row blocks k+2 .. nbk in chunks of at most KT_MAXROWS rows
  auto far_chunks = [&](int k, std::vector<std::pair<int, int>>& ch) {
    ch.clear();
    int I = k + 2;
    while (I <= nbk) {
      int J = I, rows = 0;
      while (J <= nbk && rows + (rowend(J) - rowstart(J)) <= KT_MAXROWS) { rows += rowend(J) - rowstart(J); ++J; }
      ch.emplace_back(I, J);
      I = J;
    }
  };
  std::vector<std::pair<int, int>> ch;
  out.clear();
  // the bulk queue: far tiles in stage order, the chunk nearest to the diagonal first (the chain needs it soonest)
  for (int k = 0; k + 1 < nbk; ++k) {
    far_chunks(k, ch);
    for (auto& c : ch)
      for (int e = 0; e < M; ++e) out.push_back({e, k, c.first, c.second});
  }
}
}  // namespace

size_t k2_tiled_smem_bytes(int n) {
  return ((size_t)KT_NB * KT_RS + 4 * (size_t)KT_NB * KT_LS + 32 + 64) * sizeof(double) + (size_t)(n + 4) * sizeof(int);
}

// mode 0: factor + forward solve + back-solve / draw; mode 1: factor (+ forward solve) only.
cudaError_t launch_k2_tiled(const K2Params& p, cudaStream_t s) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 16) return cudaErrorInvalidDevice;
  const int nbk = (p.n + KT_NB - 1) / KT_NB, nrb = nbk + 1;
  const size_t ctrl_ints = 1 + 3 * (size_t)p.M * nrb;
  const size_t D_bytes = (size_t)p.M * nbk * KT_NB * KT_NB * sizeof(double);
  KtCache* Cp = nullptr;
  for (int i = 0; i < g_kt_used[dev]; ++i)
    if (g_kt[dev][i].n == p.n && g_kt[dev][i].M == p.M) { Cp = &g_kt[dev][i]; break; }
  if (!Cp) {
    if (g_kt_used[dev] >= KT_CACHE) return cudaErrorMemoryAllocation;   // (64 distinct system shapes in one process)
    KtCache C;
    std::vector<KtTask> tasks;
    kt_build_tasks(tasks, p.n, p.M);
    e = cudaMalloc(&C.tasks, std::max<size_t>(tasks.size(), 1) * sizeof(KtTask));
    if (e != cudaSuccess) return e;
    // (synchronous copy: the host vector dies here; a shape is only ever built once)
    e = cudaMemcpy(C.tasks, tasks.data(), tasks.size() * sizeof(KtTask), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&C.ctrl, ctrl_ints * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&C.D, D_bytes);
    // (P is indexed by row block 2..nbk: the right-hand-side row block nbk of the LAST equation lies one tile past M*nbk)
    if (e == cudaSuccess) e = cudaMalloc(&C.P, D_bytes + KT_NB * KT_NB * sizeof(double));
    if (e != cudaSuccess) { cudaFree(C.tasks); cudaFree(C.ctrl); cudaFree(C.D); cudaFree(C.P); return e; }
    C.ntasks = (int)tasks.size();
    C.n = p.n; C.M = p.M;
    g_kt[dev][g_kt_used[dev]] = C;
    Cp = &g_kt[dev][g_kt_used[dev]++];
  }
  KtCache& C = *Cp;
  double* lws = k2_linv_workspace(D_bytes);
  if (!lws) return cudaErrorMemoryAllocation;
  e = cudaMemsetAsync(C.ctrl, 0, ctrl_ints * sizeof(int), s);
  if (e != cudaSuccess) return e;

  const size_t smem = k2_tiled_smem_bytes(p.n);
  e = cudaFuncSetAttribute(k2_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  static int num_sms[16] = {0}, per_sm[16] = {0};
  static size_t per_sm_smem[16] = {0};
  if (num_sms[dev] == 0) cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
  if (per_sm[dev] == 0 || per_sm_smem[dev] != smem) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[dev], k2_tiled_kernel, KT_THREADS, smem);
    if (e != cudaSuccess) return e;
    per_sm_smem[dev] = smem;
    if (per_sm[dev] < 1) return cudaErrorLaunchOutOfResources;
  }
  static int cap = -1;
  if (cap < 0) { const char* ev = getenv("BVB_K2T_CTAS_PER_SM"); cap = ev ? atoi(ev) : 0; }
  int occ = per_sm[dev];
  if (cap > 0 && cap < occ) occ = cap;
  const int resident = occ * num_sms[dev];
  // chain CTAs: one per equation while that leaves two thirds of the grid to the bulk queue
  const int ncrit = (p.M <= resident / 2) ? p.M : std::max(1, std::min(p.M, resident / 3));
  const int grid = std::max(ncrit + 1, std::min(ncrit + C.ntasks, resident));
  KtArgs a{};
  a.ncrit = ncrit;

  a.p = p; a.tasks = C.tasks; a.ntasks = C.ntasks; a.head = C.ctrl; a.progress = C.ctrl + 1; a.dprogress = C.ctrl + 1 + (size_t)p.M * nrb; a.pprogress = C.ctrl + 1 + 2 * (size_t)p.M * nrb; a.D = C.D; a.P = C.P; a.lws = lws;
  a.nbk = nbk; a.nrb = nrb;
  if (p.mode != 1) a.p.tstamp = p.tstamp;   // (the span is closed by the back-solve kernel below: begin here, end there)
  k2_tiled_kernel<<<grid, KT_THREADS, smem, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess || p.mode == 1) return e;
  static int fastbs = -1;
  if (fastbs < 0) { const char* ev = getenv("BVB_K2T_FASTBS"); fastbs = ev ? atoi(ev) : 1; }
  if (fastbs && KT_NB * (nbk - 1) <= KTF_THREADS) {
    const size_t fsmem = ((size_t)nbk * KT_NB * KT_NB + (size_t)nbk * KT_NB + 64) * sizeof(double) + (size_t)(p.n + 4) * sizeof(int);
    if (fsmem <= (size_t)227 * 1024) {
      e = cudaFuncSetAttribute(k2_tiled_backsolve_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      if (e != cudaSuccess) return e;
      k2_tiled_backsolve_fast<<<p.M, KTF_THREADS, fsmem, s>>>(p, lws, nbk);
      return cudaGetLastError();
    }
  }
  const size_t bsmem = ((size_t)nbk * KT_NB + 64) * sizeof(double) + (size_t)(p.n + 4) * sizeof(int);
  if (bsmem > 48 * 1024) {
    e = cudaFuncSetAttribute(k2_tiled_backsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
    if (e != cudaSuccess) return e;
  }
  k2_tiled_backsolve<<<p.M, KTB_WARPS * 32, bsmem, s>>>(p, lws, nbk);
  return cudaGetLastError();
}

}  // namespace bvb
