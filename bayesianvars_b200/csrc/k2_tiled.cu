// K2 (tiled form) - batched Cholesky + forward solve as a task queue over 32 x 32 tiles, then the back-solve / draw.
//
// Same contract and same storage as k2_cholsolve.cu (packed augmented systems written by K1, factorised in place,
// 32 x 32 inverses of the diagonal blocks in `linv_ws`; sample_coefficients.cpp:58-70), different schedule.  The
// one-CTA-per-equation kernels there leave a third of the chip idle at the headline shape (100 equations, 148 SMs) and
// walk every equation's 13 block columns in lock step inside one CTA (321 us however few equations a GPU owns).  Here
// the left-looking block algorithm is cut into tile tasks that any CTA can run:
//
//   chain(e)          one CTA per equation walks the diagonal: block column k = the diagonal tile (A_kk + D_k -
//                     L(k,k-1) L(k,k-1)' -> Cholesky -> inverse) and the tile below it, by four warp roles (kt_chain_role)
//   SUB(e, k, I0:I1)  row blocks I0..I1-1 (<= 128 rows) of block column k, I0 >= k+2:  A - L(rows, <k) L(k, <k)'  (FP64
//                     DMMA, operands straight from L2, columns consumed as they become final), times Linv_k', stored; the
//                     task nearest the diagonal (I0 == k+2) also forms the look-ahead products P (old history of the
//                     chain's next tile) and D (the diagonal accumulator two block columns ahead).
//
// SUB tasks sit in ONE list in a topological order (stage k: the far tiles of column k, nearest chunk first);
// persistent CTAs take the next index with an atomic and wait for what a task needs through per-(equation, row block)
// progress counters (relaxed polls with back-off, one acquire fence; publication = barrier + one fence + relaxed
// flags).  A waiting CTA only ever waits for tasks with a smaller index, which some resident CTA has already taken, or
// for a chain CTA, all of which are resident, so the scheme cannot deadlock; every tile is computed by exactly one
// task from fixed inputs in a fixed order: results do not depend on which CTA ran what or on the grid size (the same
// launch is bitwise reproducible; against the one-CTA-per-equation kernels the summation order differs, ~1e-15).
// What bounds it, and the measurements behind the rolled / software-pipelined chain code: DESIGN.md, section K2.
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
  double *panel, *aux, *linvT, *dn0, *dn1, *dd0, *dd1, *lr, *rsbuf;
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
  // ---- X = C Linv'  (DMMA, skipping the zero triangle); every warp works on its own rows only.  Two passes: the first
  // row block (m-tiles 0..3, one per warp) is solved, stored straight from the fragments and PUBLISHED before the rest
  // of the tile is touched - it is the tile the equation's diagonal chain is waiting for (7.9 k cycles from the arrival
  // of Linv to the flag when the whole 128-row tile went out at once).
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int cbl = colb[min(c0 + lane, n - 1)];
    for (int mt = warp + (pass ? KT_WARPS : 0); mt < (pass ? mtiles : min(mtiles, KT_WARPS)); mt += KT_WARPS) {
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
          const bool in = li < R && c < w;
          const int cbc = __shfl_sync(0xffffffffu, cbl, c);
          if (in) om[cbc + r_lo + li] = xt[nt][e];
          panel[c * KT_RS + li] = in ? xt[nt][e] : 0.0;       // (kept for the diagonal accumulator below)
        }
    }
    __syncthreads();
    if (tid == 0) {
      kt_acquire();                                    // one gpu-scope fence per pass, then plain flags
      // (a task whose rows fit the first pass - the partial last row block with the right-hand-side row - is complete now)
      const int Ib = (pass == 0) ? i0 : i0 + 1, Ie = (pass == 0 && mtiles > KT_WARPS) ? i0 + 1 : i1;
      for (int I = Ib; I < Ie; ++I) kt_st_relaxed(prog + I, k + 1);
    }
    if (pass == 0 && mtiles <= KT_WARPS) break;
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

__device__ __forceinline__ void kt_ctx_init(KtCtx& x, double* smem) {
  x.panel = smem;                                        // [32][KT_RS]  panel[c][local row]
  x.aux = x.panel + KT_NB * KT_RS;                       // [32][KT_LS]  L(k, k-1) / the originals of the tile below
  x.linvT = x.aux + KT_NB * KT_LS;                       // [32][KT_LS]  Linv of the diagonal block, row-major
  x.dn0 = x.linvT + KT_NB * KT_LS;                       // [2][32][KT_LS]  diagonal-block staging of the chain (even / odd block column)
  x.dn1 = x.dn0 + KT_NB * KT_LS;
  x.dd0 = x.dn1 + KT_NB * KT_LS;
  x.dd1 = x.dd0 + KT_NB * KT_LS;
  x.lr = x.dd1 + KT_NB * KT_LS;                          // [32][KT_LS]  the chain's factor columns, relative layout
  x.rsbuf = x.lr + KT_NB * KT_LS;                        // [32] 1/l(c,c), [32] l(c,c)
  x.colb = reinterpret_cast<int*>(x.rsbuf + 32 + 64);    // [n]
  x.s_avail = nullptr;
  x.tid = threadIdx.x; x.lane = x.tid & 31; x.warp = x.tid >> 5; x.fr = x.lane >> 2; x.fk = x.lane & 3;
}

// rsqrt for the pivot chain: hardware seed (MUFU.RSQ64H) + one third-order step, 4 dependent FP64 operations; relative
// error <= 2 ulp on the pivots of an SPD block (tools/microbench/chol32.cu: 202 instead of 281 cycles per column).
__device__ __forceinline__ double kt_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = y * y;
  const double e = fma(-x, t, 1.0);
  const double pz = fma(0.375, e, 0.5);
  const double u = y * e;
  return fma(pz, u, y);
}

// (One instance per warp role, NOT inlined: as one function the four roles shared a register allocation at the 255 cap
// and ptxas serialised every group of loads - consecutive loads into one register - so that staging 64 values took
// 4.7 k cycles.  Shared-memory pointers are rebuilt from the extern array inside so they keep their address space.)
template <int ROLE>
__device__ __noinline__ void kt_chain_role(const KtArgs& a, int eq, uint64_t* colbar /* [32], count 1 */, long long* sprof /* [24] shared */) {
  extern __shared__ __align__(16) double smem[];
  KtCtx x;
  kt_ctx_init(x, smem);
  constexpr int warp = ROLE;
  const K2Params& p = a.p;
  const int n = p.n, nbk = a.nbk, nrb = a.nrb;
  const int tid = x.tid, lane = x.lane, fr = x.fr, fk = x.fk;
  double* Lr = x.lr;                          // [32][KT_LS]  column c of the factor below its diagonal: Lr[c][i] = l(c+1+i, c)
  double* cb = x.panel + 32;                  //              rows 32..63: originals of the tile below, then C
  double* pb = x.panel + 64;                  //              rows 64..95: look-ahead product P
  double* t2 = x.panel + 96;                  //              rows 96..127: L(k+1, k-1)
  double* aux = x.aux;                        // [32][KT_LS]  L(k, k-1), then L(k+1, k)
  double* lv = x.linvT;                       // [32][KT_LS]  Linv of the diagonal block, row-major
  double* dnb[2] = {x.dn0, x.dn1};            // [32][KT_LS]  diagonal block k (even / odd k): original entries, then - newest term
  double* ddb[2] = {x.dd0, x.dd1};            // [32][KT_LS]  its accumulator D_k (k >= 2)
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

  // diagonal block kk: original entries -> dA, accumulator D -> dD, 16-byte asynchronous copies of whole 32-row column
  // segments (no registers: register-staged loads were serialised by the allocator - consecutive loads into one
  // register - and cost 5 k cycles per block).  Entries outside the lower triangle / the partial block are masked by
  // the reader (warp 0).  One warp.
  auto stage_diag = [&](int kk, double* dA, double* dD) {
    const int c0 = kk * KT_NB;
    const double* Dk = a.D + ((size_t)eq * nbk + kk) * (KT_NB * KT_NB);
    const int ch = lane & 15, cl = lane >> 4;
    int cbv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cbv[i] = colb[min(c0 + cl + 2 * i, n - 1)];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = cl + 2 * i;
      cp_async16(dA + c * KT_LS + 2 * ch, om + cbv[i] + c0 + 2 * ch);
      if (kk >= 2) cp_async16(dD + c * KT_LS + 2 * ch, Dk + c * KT_NB + 2 * ch);
    }
    cp_async_commit();
    cp_async_wait<0>();
  };

  if (warp == 2) stage_diag(0, dnb[0], ddb[0]);
  if (warp == 3) kt_bar_arrive(KT_BAR_DN, KT_THREADS);
  for (int k = 0; k < nbk; ++k) {
    const int c0 = k * KT_NB;
    const int w = min(KT_NB, n - c0);
    const int r_lo = kt_rowstart(k + 1, nbk, n), r_hi = kt_rowend(k + 1, nbk, n);   // the tile below: row block k+1
    const int R = r_hi - r_lo;
    double* dn = x.dn0 + (k & 1) * (KT_NB * KT_LS);          // (dn0 / dn1 and dd0 / dd1 are adjacent)
    const double* dd = x.dd0 + (k & 1) * (KT_NB * KT_LS);
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
      // ---- factorisation of the diagonal block: lane = row.  ONE rolled loop over the columns: the register window slides
      // one entry per column (wv[i] = a(lane, c + i)), so every register index is static inside a loop body of ~100
      // instructions.  (The fully unrolled form is 2 700 instructions of straight-line code per warp role; with four
      // warp roles and the far-tile code streaming different instruction regions through one SM's instruction cache it
      // ran at HALF the speed it has alone - tools/microbench/icache_multi.cu.)  Column c of the factor goes to shared
      // memory RELATIVE to the diagonal (Lr[c][i] = l(c+1+i, c)): the loop reads 16 aligned LDS.128 whatever c is.
      int fail = 0;
      double wv[KT_NB];
#pragma unroll
      for (int c = 0; c < KT_NB; ++c) {
        // (unconditional loads, masks applied afterwards: a load inside a conditional expression became a branch per
        // element, and through a pointer picked from an array it was a GENERIC load - 4.7 k cycles for these 64 values)
        const double raw = dn[c * KT_LS + lane], dvr = dd[c * KT_LS + lane];
        const double dv = (k >= 2) ? dvr : 0.0;
        wv[c] = (c <= lane) ? ((c < w && lane < w) ? raw + dv : ((lane == c) ? 1.0 : 0.0)) : 0.0;   // identity padding
      }
      // Software-pipelined over the columns: body c applies column c (rank-1 update of the window) and, as soon as the
      // first updated entry exists, runs the pivot chain of column c+1 (shuffle -> rsqrt -> scale -> store) - the other
      // 29 DFMAs of the update issue in the shadow of that chain's latency (tools/microbench/chol32_rolled.cu: chain
      // 124 cycles, update 180 when they run one after the other).
      double xx = __shfl_sync(0xffffffffu, wv[0], 0);
      bool ok = (xx > 0.0) && (xx < 1.7e308);
      fail = ok ? 0 : c0 + 1;
      double rs = kt_rsqrt(ok ? xx : 1.0);
      double l = wv[0] * rs;                                   // column 0 (every lane >= 0)
      if (lane > 0) Lr[lane - 1] = l;
      if (lane == 0) { rsbuf[0] = rs; rsbuf[KT_NB] = l; }        // (rsbuf[32 + c] = l(c, c))
      const double l00 = l;
      __syncwarp();
      KT_TICK(prof0, 4);
#pragma unroll 1
      for (int c = 0; c < KT_NB - 1; ++c) {
        const double2* lcol = reinterpret_cast<const double2*>(Lr + c * KT_LS);
        double tl[KT_NB];
#pragma unroll
        for (int i = 0; i < KT_NB / 2; ++i) { const double2 v = lcol[i]; tl[2 * i] = v.x; tl[2 * i + 1] = v.y; }
        const double n0 = fma(-l, tl[0], wv[1]);               // a(lane, c+1), final
        const double n1 = fma(-l, tl[1], wv[2]);
        // column c+1: pivot = n0 of lane c+1
        xx = __shfl_sync(0xffffffffu, n0, c + 1);
        ok = (xx > 0.0) && (xx < 1.7e308);
        fail = (!ok && fail == 0) ? c0 + c + 2 : fail;
        rs = kt_rsqrt(ok ? xx : 1.0);
        const double ln = (lane > c) ? n0 * rs : 0.0;          // (lane >= c + 1)
        if (lane > c + 1) Lr[(c + 1) * KT_LS + (lane - c - 2)] = ln;
        if (lane == c + 1) { rsbuf[c + 1] = rs; rsbuf[KT_NB + c + 1] = ln; }
        // the rest of column c's update
#pragma unroll
        for (int i = 2; i < KT_NB - 1; ++i) wv[i] = fma(-l, tl[i], wv[i + 1]);
        wv[KT_NB - 1] = 0.0;
        wv[0] = n0; wv[1] = n1;
        l = ln;
        __syncwarp();
        // hand-off to the inverting warp: columns 0 .. c+1 and their rsbuf entries are in shared memory
        if (((c + 1) & 7) == 7 && lane == 0) mbar_arrive(&colbar[(c + 1) >> 3]);   // (release.cta)
      }
      KT_TICK(prof0, 3);
      // L(k,k) out to global memory (off the pivot chain: only the publication of the step waits for it).  Lane r holds
      // nothing of the finished columns any more, so they are read back: l(r, c) = Lr[c][r-c-1], l(c, c) = rsbuf[32 + c].
      {
        const int cbl = colb[min(c0 + lane, n - 1)];
        double vv[KT_NB];
#pragma unroll
        for (int c = 1; c < KT_NB; ++c) {
          const double dg = rsbuf[KT_NB + c], od = Lr[c * KT_LS + max(lane - c - 1, 0)];
          vv[c] = (lane == c) ? dg : od;
        }
        vv[0] = l00;
#pragma unroll
        for (int c = 0; c < KT_NB; ++c) {
          const int cbc = __shfl_sync(0xffffffffu, cbl, c);
          if (c < w && lane >= c && lane < w) om[cbc + c0 + lane] = vv[c];
        }
      }
      if (lane == 0 && fail) atomicCAS(&p.status[eq], 0, fail);
      kt_bar_arrive(KT_BAR_PUB, KT_THREADS);
      KT_TICK(prof0, 8);
    } else if (warp == 2) {
      // ---- inverse of the diagonal block, lane = column of Linv, one column behind the factorisation
      double* lwsk = a.lws + ((size_t)eq * nbk + k) * (KT_NB * KT_NB);
      // rolled like the factorisation: sv[j] = entry (kk + j, lane) of the running inverse; row kk is final at step kk
      double sv[KT_NB];
#pragma unroll
      for (int i = 0; i < KT_NB; ++i) sv[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll 1
      for (int kk = 0; kk < KT_NB; ++kk) {
        if ((kk & 7) == 0) mbar_wait(&colbar[kk >> 3], par);
        const double xk = (kk >= lane) ? sv[0] * rsbuf[kk] : 0.0;
        lv[kk * KT_LS + lane] = xk;                  // row-major Linv[kk][j], lane j
        lwsk[kk * KT_NB + lane] = xk;
        const double2* lcol = reinterpret_cast<const double2*>(Lr + kk * KT_LS);
        double tl[KT_NB];
#pragma unroll
        for (int i = 0; i < KT_NB / 2; ++i) { const double2 v = lcol[i]; tl[2 * i] = v.x; tl[2 * i + 1] = v.y; }
#pragma unroll
        for (int i = 0; i < KT_NB - 1; ++i) sv[i] = fma(-tl[i], xk, sv[i + 1]);
        sv[KT_NB - 1] = 0.0;
      }
      kt_bar_arrive(KT_BAR_L, 96);
      kt_bar_arrive(KT_BAR_PUB, KT_THREADS);
      // the next diagonal block (its accumulator D_{k+1} was published by the look-ahead task of block column k-1)
      if (k + 1 < nbk) {
        if (k + 1 >= 2 && lane == 0) {
          while (kt_ld_acquire(dprog + k + 1) < 1) { }
          kt_acquire();
        }
        __syncwarp();
        stage_diag(k + 1, dnb[(k + 1) & 1], ddb[(k + 1) & 1]);
      }
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
      const int cbk = colb[min(c0 + lane, n - 1)];   // column offsets of this block, handed round by shuffle (stores below)
      // operands of C = A - P - L(k+1,k-1) L(k,k-1)': the tile's original entries -> cb, P -> pb, L(k+1,k-1) -> t2, by
      // asynchronous copies of warps 1 and 3 (16-byte column segments for a full row block; the partial last row block
      // and the right-hand-side row go entry by entry)
      {
        const double* Pw = a.P + ((size_t)eq * nbk + (k + 1)) * (KT_NB * KT_NB);
        const int t64 = lane + (warp == 1 ? 0 : 32);
        const int cprev = max(c0 - KT_NB, 0);
        if (R == KT_NB) {
          const int ch = t64 & 15, cl = t64 >> 4;
          int cbc[8], cbq[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { cbc[i] = colb[min(c0 + cl + 4 * i, n - 1)]; cbq[i] = colb[cprev + cl + 4 * i]; }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = cl + 4 * i;
            cp_async16(cb + c * KT_RS + 2 * ch, om + cbc[i] + r_lo + 2 * ch);
            if (k >= 2) cp_async16(pb + c * KT_RS + 2 * ch, Pw + c * KT_NB + 2 * ch);
            if (k >= 1) cp_async16(t2 + c * KT_RS + 2 * ch, om + cbq[i] + r_lo + 2 * ch);
          }
        } else {
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int idx = t64 + 64 * i, c = idx >> 5, li = idx & 31;
            if (li < R && c < w) cp_async8(cb + c * KT_RS + li, om + colb[c0 + c] + r_lo + li);
            else cb[c * KT_RS + li] = 0.0;
            if (k >= 2 && li < R) cp_async8(pb + c * KT_RS + li, Pw + idx);
            else pb[c * KT_RS + li] = 0.0;
            if (k >= 1 && li < R) cp_async8(t2 + c * KT_RS + li, om + colb[cprev + c] + r_lo + li);
            else t2[c * KT_RS + li] = 0.0;
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
        kt_bar_sync(KT_BAR_C, 64);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mt = mt0 + h;
        double g[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) g[nt][0] = g[nt][1] = 0.0;
        if (k > 0) {
#pragma unroll
          for (int ks = 0; ks < KT_NB / 4; ++ks) {
            const double av = t2[(4 * ks + fk) * KT_RS + 8 * mt + fr];
            const double* bcol = aux + (4 * ks + fk) * KT_LS;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(g[nt][0], g[nt][1], av, bcol[8 * nt + fr]);
          }
        }
        const int li = 8 * mt + fr;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * fk + e;
            double* pc = cb + c * KT_RS + li;
            const double pv = (k >= 2) ? pb[c * KT_RS + li] : 0.0;
            *pc = (li < R && c < w) ? *pc - (pv + g[nt][e]) : 0.0;
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
            const int cbc = __shfl_sync(0xffffffffu, cbk, c);
            if (in) om[cbc + r_lo + li] = xt[nt][e];
            aux[c * KT_LS + li] = in ? xt[nt][e] : 0.0;      // next step's L(k+1, k)
          }
      }
      KT_TICK(prof1, 15);
      if (warp == 1) {
        kt_bar_arrive(KT_BAR_PUB, KT_THREADS);
      } else {
        if (k + 1 < nbk) kt_bar_arrive(KT_BAR_DN, KT_THREADS);   // aux is complete: the next block column may start
        KT_TICK(prof3, 17);
        kt_bar_sync(KT_BAR_PUB, KT_THREADS);    // warps 0-2 have issued every store of this step
        if (lane == 0) {
          kt_acquire();                         // (fence.acq_rel.gpu: one fence, then two relaxed flags)
          kt_st_relaxed(prog + k, k + 1);
          kt_st_relaxed(prog + k + 1, k + 1);
        }
        __syncwarp();
        KT_TICK(prof3, 18);
      }
    }
  }
  __syncthreads();
  if (p.prof != nullptr && blockIdx.x == 0 && tid < 24) p.prof[tid] += sprof[tid];
#undef KT_TICK
}

__device__ __forceinline__ void kt_chain_run(const KtArgs& a, int warp, int eq, uint64_t* colbar, long long* sprof) {
  switch (warp) {
    case 0: kt_chain_role<0>(a, eq, colbar, sprof); break;
    case 1: kt_chain_role<1>(a, eq, colbar, sprof); break;
    case 2: kt_chain_role<2>(a, eq, colbar, sprof); break;
    default: kt_chain_role<3>(a, eq, colbar, sprof); break;
  }
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
  kt_ctx_init(x, smem);
  x.s_avail = &s_avail;
  x.tid = threadIdx.x; x.lane = x.tid & 31; x.warp = x.tid >> 5; x.fr = x.lane >> 2; x.fk = x.lane & 3;
  const K2Params& p = a.p;
  const int tid = x.tid;
  if (tid == 0) bvb_stamp_begin(p.tstamp);
  for (int i = tid; i < p.n; i += KT_THREADS) x.colb[i] = p.colbase[i];
  __syncthreads();

  if ((int)blockIdx.x < a.ncrit) {
    // one equation per chain CTA (the launch guarantees ncrit == M)
    if (tid < KT_NB) mbar_init(&s_colbar[tid], 1);
    if (tid < 24) s_prof[tid] = 0;
    for (int i = tid; i < KT_NB * KT_LS; i += KT_THREADS) x.lr[i] = 0.0;   // (entries past a column's end are read, never used)
    mbar_fence_init();
    __syncthreads();
    kt_chain_run(a, x.warp, blockIdx.x, s_colbar, s_prof);
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

// Left-looking form for n <= 416 (default).  The right-looking kernel above gives every thread a COLUMN of the current
// block row: 32 consecutive rows of one column per thread, i.e. every warp-wide load touches 32 different 32-byte
// sectors for 8 useful bytes each - 51 us per launch whatever the number of equations (ncu, warm cache), a quarter of
// the tiled K2.  Here block column k is read the way it is stored: warp w owns columns 2w, 2w+1 of the block, its lanes
// stride over the rows below the block (coalesced 256-byte requests, <= 12 values per lane and column, all in flight
// at once), the two dot products with the already solved part of x are reduced by shuffle, and warp 0 applies
// Linv_k' (prefetched while the dot products run).  Two CTA barriers per block column.
constexpr int KTL_WARPS = 16, KTL_THREADS = KTL_WARPS * 32, KTL_MAXN = 416, KTL_MAXR = (KTL_MAXN - KT_NB) / 32;
__global__ void __launch_bounds__(KTL_THREADS, 1) k2_tiled_backsolve_ll(const K2Params p, const double* __restrict__ lws, int nbk) {
  __shared__ double y[KTL_MAXN + 32];
  __shared__ double sblk[KT_NB];
  __shared__ int colb[KTL_MAXN];
  const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, eq = blockIdx.x;
  const double* __restrict__ om = p.omega + (size_t)eq * p.ldo;
  const double* __restrict__ lwe = lws + (size_t)eq * nbk * (KT_NB * KT_NB);
  for (int i = tid; i < n; i += KTL_THREADS) colb[i] = p.colbase[i];
  __syncthreads();
  for (int i = tid; i < KTL_MAXN + 32; i += KTL_THREADS)
    y[i] = (i < n) ? __ldcg(om + colb[i] + n) + p.z[(size_t)eq * p.ldz + i] : 0.0;   // row n of L = the forward-solved rhs
  __syncthreads();
  bool bad = false;
  const int xworld = (p.xch.peer_xbuf != nullptr) ? p.xch.world : 1;
  const size_t xoff = xworld > 1 ? ((((p.xch.epoch_base >> 32) & 1ull) * 2 + (*p.xch.sweep_p & 1u)) * (size_t)p.xch.world + p.xch.rank) * p.xch.blk : 0;
  for (int k = nbk - 1; k >= 0; --k) {
    const int c0 = k * KT_NB, w = min(KT_NB, n - c0), rb = c0 + KT_NB;
    double lk[KT_NB];
    if (warp == 0) {
      const double* wsk = lwe + (size_t)k * (KT_NB * KT_NB);
#pragma unroll
      for (int c = 0; c < KT_NB; ++c) lk[c] = __ldcg(wsk + c * KT_NB + lane);
    }
    double v0[KTL_MAXR], v1[KTL_MAXR];
    const int ca = 2 * warp, cb2 = 2 * warp + 1;
    const double* pa = om + colb[min(c0 + ca, n - 1)];
    const double* pb = om + colb[min(c0 + cb2, n - 1)];
#pragma unroll
    for (int j = 0; j < KTL_MAXR; ++j) {
      const int i = min(rb + lane + 32 * j, n - 1);          // (clamped: always a valid address, masked below)
      v0[j] = __ldcg(pa + i);
      v1[j] = __ldcg(pb + i);
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < KTL_MAXR; ++j) {
      const int i = rb + lane + 32 * j;
      const double yi = (i < n) ? y[min(i, KTL_MAXN + 31)] : 0.0;
      s0 = fma(v0[j], yi, s0);
      s1 = fma(v1[j], yi, s1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
      sblk[ca] = (ca < w) ? y[c0 + ca] - s0 : 0.0;
      sblk[cb2] = (cb2 < w) ? y[c0 + cb2] - s1 : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
      for (int c = 0; c < KT_NB; c += 4) {
        x0 = fma(lk[c], sblk[c], x0);
        x1 = fma(lk[c + 1], sblk[c + 1], x1);
        x2 = fma(lk[c + 2], sblk[c + 2], x2);
        x3 = fma(lk[c + 3], sblk[c + 3], x3);
      }
      const double xv = (lane < w) ? (x0 + x1) + (x2 + x3) : 0.0;
      y[c0 + lane] = xv;
      if (lane < w) {
        p.phi[(size_t)eq * p.ldphi + c0 + lane] = xv;
        bad |= !isfinite(xv);
      }
    }
    __syncthreads();
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (xworld > 1) {
    // fused exchange: the solved column (complete in shared memory) straight into every peer's exchange buffer - one
    // coalesced remote store per thread and peer over NVLink, off the 13-step chain above
    for (int pi = 0; pi + 1 < xworld; ++pi) {
      const int pp = pi < p.xch.rank ? pi : pi + 1;
      double* dst = p.xch.peer_xbuf[pp] + xoff + (size_t)eq * p.ldphi;
      for (int i = tid; i < n; i += KTL_THREADS) dst[i] = y[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(p.xch.cnt, 1);
      __threadfence();
      if (old == (int)gridDim.x - 1) {               // every equation of this rank is out: stamp the peers' flags
        *p.xch.cnt = 0;
        __threadfence_system();
        const unsigned long long stamp = p.xch.epoch_base + *p.xch.sweep_p + 1ull;
        for (int pi = 0; pi + 1 < xworld; ++pi) {
          const int pp = pi < p.xch.rank ? pi : pi + 1;
          // (relaxed: the system-scope fence above already orders every pushed column before the flags - a release store
          // per peer put one MEMBAR.SYS in front of each of the world-1 flags, ~5 us apiece at 8 GPUs)
          asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.xch.peer_xflag[pp] + p.xch.rank), "l"(stamp) : "memory");
        }
      }
    }
  }
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
  return ((size_t)KT_NB * KT_RS + 7 * (size_t)KT_NB * KT_LS + 32 + 64) * sizeof(double) + (size_t)(n + 4) * sizeof(int);
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
  if (2 * p.M > resident) return cudaErrorNotSupported;   // every equation needs its own chain CTA next to at least as many far-tile CTAs
  const int ncrit = p.M;
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
  if (fastbs < 0) { const char* ev = getenv("BVB_K2T_FASTBS"); fastbs = ev ? atoi(ev) : 2; }   // 2: left-looking (default), 1: right-looking, 0: generic
  if (fastbs == 2 && p.n <= KTL_MAXN) {
    k2_tiled_backsolve_ll<<<p.M, KTL_THREADS, 0, s>>>(p, lws, nbk);
    if (p.xch_fused && p.xch.peer_xbuf != nullptr && p.xch.world > 1) *p.xch_fused = true;   // (this kernel pushed the columns itself)
    return cudaGetLastError();
  }
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
