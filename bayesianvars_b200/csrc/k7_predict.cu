// K7 - predictive simulation / log predictive likelihoods over stored draws.
//
// Replaces out_of_sample (src/utilities_cpp.cpp:216-393) with its helpers
// build_sigma (:45-67), PHI_power0 (:125-142), Sigma_pred_uncond (:144-178) and
// dmvnrm_arma_fast (:19-42).  Every (draw r, replicate j) pair is an independent
// chain over the horizons, so the grid is persistent CTAs striding over the
// R*each chains; a chain keeps its Kp x Kp companion covariance, the M x M
// innovation covariance and its Cholesky factor in a per-CTA global scratch slab
// (L2-resident) and runs the dense steps (PHI' S, S PHI, PHI-power) through a
// shared-memory tiled FP64 GEMM.  The reference's arithmetic is followed step by
// step (same recursion, same symmetrisation, chol -> inverse-free solves).
// Slice index of chain (r, j) is r + j*R (:352, :364, :376).
#include "bvb_internal.cuh"

namespace bvb {

constexpr int PR_THREADS = 256;

struct PredictParams {
  int each, Kp, M, R, nf, nU, ld_lv, na, max_ahead, nsv, factor, lpl, lpl_sub, nvoi, simulate, p;
  const double *x, *PHI, *U, *facload, *logvar_T, *sv_mu, *sv_phi, *sv_sigma, *Y_obs, *z_logvar, *z_pred;
  const int *ahead, *sv_ind, *voi;
  double *LPL, *PLu, *LPLs, *pred;
  long long* prof;          // BVB_PREDICT_PROF=1: cycles of CTA 0 per phase
  double* scratch;          // per CTA slab
  size_t slab;              // doubles per CTA
  uint64_t seed;
  int* status;
};

// C[m x n] (ldc) = op(A) * B (+ C if accumulate); A is k x m when transA (lda), else m x k.  All column-major in
// global memory.  64 x 64 output tile per pass, K chunks of 16 staged k-major in shared memory (row stride 68 == 4
// mod 16: conflict-free m8n8k4 fragment loads), 8 warps as 4 (m) x 2 (n), each 16 x 32 of the tile on FP64 DMMA.
constexpr int PR_LD = 68;
constexpr int PR_ST = 3;                              // cp.async ring stages (one K chunk of 16 each)
constexpr int PR_STAGE = 2 * 16 * PR_LD;              // doubles per stage: sA[16][PR_LD], sB[16][PR_LD]
// The operands live in global scratch (L2): a 3-stage cp.async ring keeps two K chunks in flight under the DMMAs of the
// current one, one CTA barrier per chunk (the first version loaded a chunk, synchronised, multiplied, synchronised: the
// tensor pipe idled for a full L2 round trip per 32 DMMAs).  Accumulation order over k is unchanged.
// transB: B is stored n x k (element (k, n) at B[k * ldb + n]); alpha scales the product (C = [C +] alpha * op(A) op(B)).
__device__ void cta_gemm(int m, int n, int k, const double* A, int lda, bool transA, const double* B, int ldb, double* C,
                         int ldc, bool accumulate, double* sbuf /*[PR_ST][PR_STAGE]*/, bool transB = false,
                         double alpha = 1.0, const int* bmap = nullptr /* B(k, n) lives at B[bmap[n] * ldb + bmap[k]] */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int nk = (k + 15) >> 4;
  for (int m0 = 0; m0 < m; m0 += 64) {
    for (int n0 = 0; n0 < n; n0 += 64) {
      double acc[2][4][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      auto issue = [&](int c) {
        double* sA = sbuf + (size_t)(c % PR_ST) * PR_STAGE;
        double* sB = sA + 16 * PR_LD;
        const int k0 = c << 4;
        for (int i = tid; i < 16 * 64; i += PR_THREADS) {
          // sA[kk][mm] = op(A)[m0+mm, k0+kk];  sB[kk][nn] = B[k0+kk, n0+nn].  The fastest index follows the operand's
          // contiguous dimension so the global reads coalesce.
          int kk, mm;
          if (transA) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
          const int gm = m0 + mm, gk = k0 + kk;
          double* da = sA + kk * PR_LD + mm;
          if (gm < m && gk < k) cp_async8(da, transA ? A + (size_t)gm * lda + gk : A + (size_t)gk * lda + gm);
          else *da = 0.0;
          int kb, nn;
          if (transB) { nn = i & 63; kb = i >> 6; } else { kb = i & 15; nn = i >> 4; }
          const int gn = n0 + nn, gkb = k0 + kb;
          double* db = sB + kb * PR_LD + nn;
          if (gn < n && gkb < k)
            cp_async8(db, bmap ? B + (size_t)bmap[gn] * ldb + bmap[gkb]
                               : (transB ? B + (size_t)gkb * ldb + gn : B + (size_t)gn * ldb + gkb));
          else *db = 0.0;
        }
      };
#pragma unroll
      for (int c = 0; c < PR_ST - 1; ++c) {
        if (c < nk) issue(c);
        cp_async_commit();
      }
      for (int c = 0; c < nk; ++c) {
        cp_async_wait<PR_ST - 2>();
        __syncthreads();                               // chunk c landed everywhere; the stage of chunk c-1 is free
        if (c + PR_ST - 1 < nk) issue(c + PR_ST - 1);
        cp_async_commit();
        const double* sA = sbuf + (size_t)(c % PR_ST) * PR_STAGE;
        const double* sB = sA + 16 * PR_LD;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double* ap = sA + (4 * ks + fk) * PR_LD + 16 * wm + fr;
          const double* bp = sB + (4 * ks + fk) * PR_LD + 32 * wn + fr;
          const double a0 = ap[0], a1 = ap[8];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double bv = bp[8 * b];
            dmma884(acc[0][b][0], acc[0][b][1], a0, bv);
            dmma884(acc[1][b][0], acc[1][b][1], a1, bv);
          }
        }
      }
      cp_async_wait<0>();
      __syncthreads();                                 // the ring is free for the next tile
      // accumulator fragment: rows 8a + fr, columns 8b + 2 fk + {0, 1}
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int gm = m0 + 16 * wm + 8 * a + fr, gn = n0 + 32 * wn + 8 * b + 2 * fk + e;
            if (gm < m && gn < n) {
              double* c = C + (size_t)gn * ldc + gm;
              *c = accumulate ? *c + alpha * acc[a][b][e] : alpha * acc[a][b][e];
            }
          }
    }
  }
  __syncthreads();
}

// In-place lower Cholesky of the n x n matrix L (ld), right-looking, whole CTA.  Returns false on a bad pivot.
__device__ bool cta_chol(int n, double* L, int ld, double* sh) {
  bool ok = true;
  for (int c = 0; c < n; ++c) {
    __syncthreads();
    const double d = L[(size_t)c * ld + c];
    if (!(d > 0.0)) ok = false;
    const double piv = sqrt(d > 0.0 ? d : 1.0);
    __syncthreads();
    for (int i = c + threadIdx.x; i < n; i += PR_THREADS) L[(size_t)c * ld + i] = (i == c) ? piv : L[(size_t)c * ld + i] / piv;
    __syncthreads();
    const int rem = n - c - 1;
    for (int idx = threadIdx.x; idx < rem * rem; idx += PR_THREADS) {
      const int a = c + 1 + idx % rem, b = c + 1 + idx / rem;
      if (a >= b) L[(size_t)b * ld + a] -= L[(size_t)c * ld + a] * L[(size_t)c * ld + b];
    }
  }
  __syncthreads();
  (void)sh;
  return ok;
}

// Blocked form: panels of 16 columns are factorised in shared memory (the GEMM ring, n <= 408 rows), the trailing
// matrix takes a rank-16 update on the DMMA GEMM.  (The column-at-a-time version above goes through L2 three barriers
// per column: 21 % of the kernel at M = 200.)  The strict upper triangle of L is scratch afterwards.
constexpr int PR_CNB = 16;
__device__ bool cta_chol_blocked(int n, double* L, int ld, double* sbuf) {
  const int tid = threadIdx.x;
  bool ok = true;
  for (int k0 = 0; k0 < n; k0 += PR_CNB) {
    const int nb = min(PR_CNB, n - k0), rows = n - k0;
    const int RS = rows | 1;
    double* P = sbuf;                                 // [nb][RS]
    for (int c = 0; c < nb; ++c)
      for (int r = tid; r < rows; r += PR_THREADS) P[c * RS + r] = L[(size_t)(k0 + c) * ld + k0 + r];
    __syncthreads();
    for (int c = 0; c < nb; ++c) {
      const double d = P[c * RS + c];
      if (!(d > 0.0)) ok = false;
      const double piv = sqrt(d > 0.0 ? d : 1.0);
      __syncthreads();
      for (int r = c + tid; r < rows; r += PR_THREADS) P[c * RS + r] = (r == c) ? piv : P[c * RS + r] / piv;
      __syncthreads();
      for (int r = c + 1 + tid; r < rows; r += PR_THREADS) {
        const double lr = P[c * RS + r];
        const int cend = min(nb - 1, r);
        for (int c2 = c + 1; c2 <= cend; ++c2) P[c2 * RS + r] -= lr * P[c * RS + c2];
      }
      __syncthreads();
    }
    for (int c = 0; c < nb; ++c)
      for (int r = c + tid; r < rows; r += PR_THREADS) L[(size_t)(k0 + c) * ld + k0 + r] = P[c * RS + r];
    __syncthreads();
    const int rem = rows - nb;
    if (rem > 0) {
      const double* Ap = L + (size_t)k0 * ld + k0 + nb;  // rem x nb panel below the diagonal block
      cta_gemm(rem, rem, nb, Ap, ld, false, Ap, ld, L + (size_t)(k0 + nb) * ld + k0 + nb, ld, true, sbuf, true, -1.0);
    }
  }
  return __syncthreads_and(ok ? 1 : 0) != 0;
}

// log N(x; mean, L L') for the n-vector d = x - mean (overwritten with L^-1 d); thread 0 returns the value.
__device__ double cta_mvn_logdens(int n, const double* L, int ld, double* d) {
  double out = 0.0;
  if (threadIdx.x == 0) {
    double ld_sum = 0.0, q = 0.0;
    for (int i = 0; i < n; ++i) {
      double v = d[i];
      for (int k = 0; k < i; ++k) v -= L[(size_t)k * ld + i] * d[k];
      v /= L[(size_t)i * ld + i];
      d[i] = v;
      q += v * v;
      ld_sum += log(L[(size_t)i * ld + i]);
    }
    out = -ld_sum - 0.5 * n * 1.8378770664093453 - 0.5 * q;   // log(2 pi)
  }
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(PR_THREADS) predict_kernel(const PredictParams p) {
  extern __shared__ __align__(16) double sbuf[];   // [PR_ST][PR_STAGE] GEMM operand ring, then int pmap[Kp]
  int* pmap = reinterpret_cast<int*>(sbuf + PR_ST * PR_STAGE);
  __shared__ double sres[2];
  double* sA = sbuf;
  const bool prof = p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = prof ? clock64() : 0;
#define PR_TICK(slot) do { if (prof) { const long long tn = clock64(); tacc[slot] += tn - tprev; tprev = tn; } } while (0)
  const int Kp = p.Kp, M = p.M, R = p.R, P = p.p, tid = threadIdx.x;
  double* base = p.scratch + (size_t)blockIdx.x * p.slab;
  double* SL = base;                               // Kp x Kp   companion covariance, lag blocks rotated by `off` (see below)
  double* SIG = SL + (size_t)Kp * Kp;              // M x M
  double* LCH = SIG + (size_t)M * M;               // M x M (also used for the VoI sub-matrix)
  double* UINV = LCH + (size_t)M * M;              // M x M
  double* xb0 = UINV + (size_t)M * M;              // [Kp] predictor row of the current horizon
  double* xb1 = xb0 + Kp;                          // [Kp]
  double* lvp = xb1 + Kp;                          // [M + nf]
  double* lvvar = lvp + (M + p.nf);                // [nsv]
  double* yp = lvvar + (p.nsv > 0 ? p.nsv : 1);    // [M]
  double* dv = yp + M;                             // [M]
  double* GB = dv + M;                             // M x pM    PHI' SL
  const int nsave = R * p.each;
  for (int chain = blockIdx.x; chain < nsave; chain += gridDim.x) {
    const int r = chain % R, j = chain / R;
    const double* PHI = p.PHI + (size_t)r * Kp * M;
    const double* lvT = p.logvar_T + (size_t)r * p.ld_lv;
    for (size_t i = tid; i < (size_t)Kp * Kp; i += PR_THREADS) SL[i] = 0.0;       // the intercept row / column stay 0
    int off = 0;                                                                  // logical lag block b lives in physical block (b + off) % P
    for (int i = tid; i < Kp; i += PR_THREADS) pmap[i] = i;
    double* xc = xb0;
    double* xn = xb1;
    for (int i = tid; i < Kp; i += PR_THREADS) xc[i] = p.x[i];
    for (int i = tid; i < p.nsv; i += PR_THREADS) lvvar[i] = 0.0;
    if (!p.factor) {
      // U_inv = inv(trimatu(Um)), one thread per column: back substitution on the unit upper triangle
      const double* u = p.U + (size_t)r * p.nU;
      for (int c = tid; c < M; c += PR_THREADS) {
        double* col = UINV + (size_t)c * M;
        for (int i = M - 1; i >= 0; --i) {
          double v = (i == c) ? 1.0 : 0.0;
          if (i < c) for (int k = i + 1; k <= c; ++k) v -= u[(size_t)k * (k - 1) / 2 + i] * col[k];
          col[i] = (i <= c) ? v : 0.0;
        }
      }
    }
    __syncthreads();
    int counter = 0;
    for (int i = 0; i < p.max_ahead; ++i) {
      // ---- log-variance forecast (:300-304, :325-329)
      for (int s = tid; s < M + p.nf; s += PR_THREADS) lvp[s] = lvT[s];
      __syncthreads();
      for (int s = tid; s < p.nsv; s += PR_THREADS) {
        const double mu = p.sv_mu[(size_t)r * p.nsv + s], ph = p.sv_phi[(size_t)r * p.nsv + s], sg = p.sv_sigma[(size_t)r * p.nsv + s];
        const int idx = p.sv_ind[s];
        const double mean = mu + pow(ph, (double)(i + 1)) * (lvT[idx] - mu);
        const double t = sg * pow(ph, (double)i);
        const double var = lvvar[s] + t * t;
        lvvar[s] = var;
        double z;
        if (p.z_logvar) z = p.z_logvar[(((size_t)r * p.max_ahead + i) * p.each + j) * p.nsv + s];
        else { RngStream g(p.seed, 0x700u, (uint32_t)(chain * p.max_ahead + i), (uint32_t)s); z = g.normal(); }
        lvp[idx] = mean + z * sqrt(var);
      }
      PR_TICK(0);
      // ---- predictive mean.  The reference carries PHI^(i+1) (PHI_power0 :125-142, a Kp x M x M product per horizon)
      // and forms X_T_plus_1 * PHI^(i+1) (:347); the same number is the iterated forecast yhat_i = x_i' PHI with the
      // predictor row shifted by one lag per horizon (x_{i+1} = [yhat_i, x_i[0:(p-1)M], intercept]) - a Kp x M
      // matrix-vector product instead of 18 % of the kernel; the two agree to 1e-16 relative (tests: 1e-9).
      {
        const int lane = tid & 31, warp = tid >> 5;
        for (int m2 = warp; m2 < M; m2 += PR_THREADS / 32) {
          const double* col = PHI + (size_t)m2 * Kp;
          double v = 0.0;
          for (int k = lane; k < Kp; k += 32) v = fma(xc[k], col[k], v);
          v = warp_sum(v);
          if (lane == 0) yp[m2] = v;
        }
        __syncthreads();
        for (int k = tid; k < Kp; k += PR_THREADS) xn[k] = (k < M) ? yp[k] : (k < P * M ? xc[k - M] : xc[k]);
        __syncthreads();
        double* t = xc; xc = xn; xn = t;
      }
      PR_TICK(1);
      // ---- Sigma (build_sigma :45-67, return_chol = false)
      for (int e = tid; e < M * M; e += PR_THREADS) {
        const int a = e % M, b = e / M;
        double v = 0.0;
        if (p.factor) {
          const double* L = p.facload + (size_t)r * M * p.nf;
          for (int k = 0; k < p.nf; ++k) v = fma(L[(size_t)k * M + a] * L[(size_t)k * M + b], exp(lvp[M + k]), v);
          if (a == b) v += exp(lvp[a]);
        } else {
          const int lo = a > b ? a : b;  // U_inv is upper triangular: rows m <= min(a, b)
          (void)lo;
          const int mm = a < b ? a : b;
          for (int m2 = 0; m2 <= mm; ++m2) v = fma(UINV[(size_t)a * M + m2] * UINV[(size_t)b * M + m2], exp(lvp[m2]), v);
        }
        SIG[e] = v;
      }
      __syncthreads();
      PR_TICK(2);
      // ---- companion covariance recursion (Sigma_pred_uncond :144-178), n_ahead = i
      if (P > 1) {
        if (i > 0) {
          // The reference builds tmp = [PHI' SL ; SL.rows(0,(p-1)M-1)], then SL(0:M,0:M) = tmp PHI, SL.cols(M,pM-1) =
          // tmp.cols(0,(p-1)M-1) and copies the upper triangle over the lower one: the covariance of the lags moves one
          // block down the diagonal, new[b+1, c+1] = old[b, c].  Here that move is free: the lag blocks are addressed
          // through a rotation (logical block b = physical block (b + off) % P, `pmap`), each horizon decrements `off`,
          // and only the new first block row / column are written - into the physical block the oldest lag vacates.
          // Same values as the reference (old is symmetric), none of its four Kp x Kp passes (34 % of the first version).
          const int mn = (i < P ? i : P) * M;
          const int sh = (P - 1) * M;
          cta_gemm(M, mn, mn, PHI, Kp, true, SL, Kp, GB, M, false, sbuf, false, 1.0, pmap);   // G = PHI[0:mn,:]' SL[0:mn,0:mn]
          PR_TICK(4);
          off = (off + P - 1) % P;
          __syncthreads();
          for (int k = tid; k < P * M; k += PR_THREADS) pmap[k] = ((k / M + off) % P) * M + k % M;
          __syncthreads();
          double* TL = SL + (size_t)(off * M) * Kp + off * M;                        // new top-left block
          cta_gemm(M, M, mn, GB, M, false, PHI, Kp, TL, Kp, false, sbuf);            // new[0:M,0:M] = G[0:M,0:mn] PHI[0:mn,:]
          PR_TICK(4);
          {
            const int lane = tid & 31, warp = tid >> 5, nwp = PR_THREADS / 32;
            for (int c2 = warp; c2 < sh; c2 += nwp) {                               // new[0:M, M+c] = G[:, c] (0 beyond mn)
              double* top = SL + (size_t)pmap[M + c2] * Kp + off * M;
              const double* g = GB + (size_t)c2 * M;
              for (int a2 = lane; a2 < M; a2 += 32) top[a2] = (c2 < mn) ? g[a2] : 0.0;
            }
            for (int a2 = warp; a2 < M; a2 += nwp) {                                // new[M+c, 0:M] = G[:, c]' (lower <- upper)
              double* dst = SL + (size_t)(off * M + a2) * Kp;
              for (int c2 = lane; c2 < sh; c2 += 32) dst[pmap[M + c2]] = (c2 < mn) ? GB[(size_t)c2 * M + a2] : 0.0;
            }
          }
          __syncthreads();
          for (int e = tid; e < M * M; e += PR_THREADS) {                           // strict lower of the top-left block <- upper
            const int row = e % M, col = e / M;
            if (row > col) TL[(size_t)col * Kp + row] = TL[(size_t)row * Kp + col];
          }
          __syncthreads();
        }
      } else if (P == 1) {
        if (i > 0) {
          cta_gemm(M, M, M, PHI, Kp, true, SL, Kp, GB, M, false, sbuf);
          cta_gemm(M, M, M, GB, M, false, PHI, Kp, SL, Kp, false, sbuf);
        }
      }
      const double* STL = SL + (size_t)(off * M) * Kp + off * M;   // top-left (current horizon) block
      for (int e = tid; e < M * M; e += PR_THREADS) SL[(size_t)(off * M + e / M) * Kp + off * M + (e % M)] += SIG[e];
      __syncthreads();
      PR_TICK(3);
      // ---- requested horizon? (:346-380)
      if (counter < p.na && i == p.ahead[counter] - 1) {
        for (int e = tid; e < M * M; e += PR_THREADS) LCH[e] = STL[(size_t)(e / M) * Kp + (e % M)];
        __syncthreads();
        PR_TICK(5);
        const bool ok = (M <= PR_ST * PR_STAGE / PR_CNB - 1) ? cta_chol_blocked(M, LCH, M, sbuf) : cta_chol(M, LCH, M, sA);
        PR_TICK(6);
        if (!ok && tid == 0) atomicCAS(p.status, 0, chain + 1);
        if (p.lpl) {
          for (int m2 = tid; m2 < M; m2 += PR_THREADS) {
            const double y = p.Y_obs[(size_t)m2 * p.na + counter];
            dv[m2] = y - yp[m2];
            const double sd = sqrt(STL[(size_t)m2 * Kp + m2]);
            const double zz = (y - yp[m2]) / sd;
            p.PLu[((size_t)chain * M + m2) * p.na + counter] = exp(-0.5 * zz * zz) / (sd * 2.5066282746310002);
          }
          __syncthreads();
          const double ll = cta_mvn_logdens(M, LCH, M, dv);
          if (tid == 0) p.LPL[(size_t)chain * p.na + counter] = ll;
        }
        if (p.simulate) {
          for (int m2 = tid; m2 < M; m2 += PR_THREADS) {
            // y_pred + rand_vec * chol(Sigma) (upper R = L'): component m2 = sum_{k <= m2} z_k L[m2, k]
            double v = yp[m2];
            for (int k = 0; k <= m2; ++k) {
              double z;
              if (p.z_pred) z = p.z_pred[(((size_t)r * p.na + counter) * p.each + j) * M + k];
              else { RngStream g(p.seed, 0x701u, (uint32_t)(chain * p.na + counter), (uint32_t)k); z = g.normal(); }
              v = fma(z, LCH[(size_t)k * M + m2], v);
            }
            p.pred[((size_t)chain * M + m2) * p.na + counter] = v;
          }
        }
        __syncthreads();
        if (p.lpl && p.lpl_sub) {
          const int nv = p.nvoi;
          for (int e = tid; e < nv * nv; e += PR_THREADS) LCH[e] = STL[(size_t)p.voi[e / nv] * Kp + p.voi[e % nv]];
          for (int e = tid; e < nv; e += PR_THREADS) dv[e] = p.Y_obs[(size_t)p.voi[e] * p.na + counter] - yp[p.voi[e]];
          __syncthreads();
          const bool ok2 = (nv <= PR_ST * PR_STAGE / PR_CNB - 1) ? cta_chol_blocked(nv, LCH, nv, sbuf) : cta_chol(nv, LCH, nv, sA);
          const double ll = cta_mvn_logdens(nv, LCH, nv, dv);
          if (tid == 0) p.LPLs[(size_t)chain * p.na + counter] = ok2 ? ll : nan("");
        }
        ++counter;   // every chain owns exactly one replicate j, so the horizon counter advances here
      }
      __syncthreads();
      PR_TICK(7);
    }
    (void)sres;
  }
  if (prof)
    for (int i = 0; i < 8; ++i) p.prof[i] = tacc[i];
#undef PR_TICK
}

cudaError_t launch_predict(PredictParams& p, int num_sms, cudaStream_t s, double** scratch_out) {
  const int nsave = p.R * p.each;
  int grid = 2 * num_sms;
  if (grid > nsave) grid = nsave;
  p.slab = (size_t)p.Kp * p.Kp + (size_t)3 * p.M * p.M + (size_t)2 * p.Kp + (size_t)(p.M + p.nf) +
           (size_t)(p.nsv > 0 ? p.nsv : 1) + (size_t)2 * p.M + (size_t)p.M * p.p * p.M + 8;
  cudaError_t e = cudaMalloc(scratch_out, p.slab * grid * sizeof(double));
  if (e != cudaSuccess) return e;
  p.scratch = *scratch_out;
  const size_t smem = (size_t)PR_ST * PR_STAGE * sizeof(double) + (size_t)(p.Kp + 2) * sizeof(int);
  e = cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  predict_kernel<<<grid, PR_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace bvb

using namespace bvb;

extern "C" int bvb_predict(bvb_handle* h, int each, int Kp, int M, int R, int n_factors, int ld_logvar, int factor,
                           const double* X_T_plus_1, const double* PHI, const double* U, const double* facload,
                           const double* logvar_T, const int* ahead, int n_ahead, const double* sv_mu,
                           const double* sv_phi, const double* sv_sigma, const int* sv_indicator, int n_sv, int LPL,
                           const double* Y_obs, int LPL_subset, const int* VoI, int n_voi, int simulate_predictive,
                           const double* z_logvar, const double* z_pred, double* LPL_draws, double* PL_univariate_draws,
                           double* LPL_sub_draws, double* predictions) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 5; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (each <= 0 || Kp <= 0 || M <= 0 || R < 0 || Kp < M || !X_T_plus_1 || !PHI || !logvar_T || !ahead || n_ahead <= 0 ||
      (factor ? (n_factors > 0 && !facload) : (M > 1 && !U)) || (n_sv > 0 && (!sv_mu || !sv_phi || !sv_sigma || !sv_indicator)) ||
      (LPL && (!Y_obs || !LPL_draws || !PL_univariate_draws)) || (LPL && LPL_subset && (!VoI || n_voi <= 0 || !LPL_sub_draws)) ||
      (simulate_predictive && !predictions) || ld_logvar < M + (factor ? n_factors : 0)) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_predict: invalid argument");
    return BVB_ERR_ARG;
  }
  if (R == 0) return BVB_OK;
  int max_ahead = 0;
  for (int i = 0; i < n_ahead; ++i) {
    if (ahead[i] <= 0 || (i > 0 && ahead[i] <= ahead[i - 1])) { bvb_set_error(h, BVB_ERR_ARG, "bvb_predict: ahead must be sorted and positive"); return BVB_ERR_ARG; }
    max_ahead = ahead[i];
  }
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t d = sizeof(double);
  const int nsave = R * each;
  const size_t nU = (size_t)M * (M - 1) / 2;
  std::vector<void*> allocs;
  auto fin = [&](int rc) { for (void* q : allocs) cudaFree(q); return rc; };
#define PR_TRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { bvb_set_error(h, BVB_ERR_CUDA, "bvb_predict: %s", cudaGetErrorString(_e)); return fin(BVB_ERR_CUDA); } } while (0)
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    *dst = nullptr;
    if (!src || bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    allocs.push_back(*dst);
    return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, s);
  };
  PredictParams p{};
  p.each = each; p.Kp = Kp; p.M = M; p.R = R; p.nf = factor ? n_factors : 0; p.nU = (int)nU; p.ld_lv = ld_logvar; p.na = n_ahead;
  p.max_ahead = max_ahead; p.nsv = n_sv; p.factor = factor ? 1 : 0; p.lpl = LPL ? 1 : 0; p.lpl_sub = (LPL && LPL_subset) ? 1 : 0;
  p.nvoi = n_voi; p.simulate = simulate_predictive ? 1 : 0; p.p = Kp / M; p.seed = h->seed;
  PR_TRY(up(X_T_plus_1, d * Kp, (void**)&p.x));
  PR_TRY(up(PHI, d * Kp * M * R, (void**)&p.PHI));
  if (!factor) PR_TRY(up(U, d * nU * R, (void**)&p.U));
  if (factor && n_factors > 0) PR_TRY(up(facload, d * M * n_factors * R, (void**)&p.facload));
  PR_TRY(up(logvar_T, d * ld_logvar * R, (void**)&p.logvar_T));
  PR_TRY(up(ahead, sizeof(int) * n_ahead, (void**)&p.ahead));
  if (n_sv > 0) {
    PR_TRY(up(sv_mu, d * n_sv * R, (void**)&p.sv_mu));
    PR_TRY(up(sv_phi, d * n_sv * R, (void**)&p.sv_phi));
    PR_TRY(up(sv_sigma, d * n_sv * R, (void**)&p.sv_sigma));
    PR_TRY(up(sv_indicator, sizeof(int) * n_sv, (void**)&p.sv_ind));
    PR_TRY(up(z_logvar, d * (size_t)R * max_ahead * each * n_sv, (void**)&p.z_logvar));
  }
  if (LPL) PR_TRY(up(Y_obs, d * n_ahead * M, (void**)&p.Y_obs));
  if (p.lpl_sub) PR_TRY(up(VoI, sizeof(int) * n_voi, (void**)&p.voi));
  if (simulate_predictive) PR_TRY(up(z_pred, d * (size_t)R * n_ahead * each * M, (void**)&p.z_pred));
  auto dalloc = [&](size_t n, double** q) -> cudaError_t {
    cudaError_t e = cudaMalloc(q, d * (n ? n : 1));
    if (e == cudaSuccess) allocs.push_back(*q);
    return e;
  };
  if (LPL) { PR_TRY(dalloc((size_t)n_ahead * nsave, &p.LPL)); PR_TRY(dalloc((size_t)n_ahead * M * nsave, &p.PLu)); }
  if (p.lpl_sub) PR_TRY(dalloc((size_t)n_ahead * nsave, &p.LPLs));
  if (simulate_predictive) PR_TRY(dalloc((size_t)n_ahead * M * nsave, &p.pred));
  PR_TRY(cudaMalloc(&p.status, sizeof(int)));
  allocs.push_back(p.status);
  PR_TRY(cudaMemsetAsync(p.status, 0, sizeof(int), s));
  double* scratch = nullptr;
  if (getenv("BVB_PREDICT_PROF")) {
    PR_TRY(cudaMalloc(&p.prof, 8 * sizeof(long long)));
    allocs.push_back(p.prof);
    PR_TRY(cudaMemsetAsync(p.prof, 0, 8 * sizeof(long long), s));
  }
  PR_TRY(cudaEventRecord(h->ev[0], s));
  cudaError_t le = launch_predict(p, h->num_sms, s, &scratch);
  if (scratch) allocs.push_back(scratch);
  PR_TRY(le);
  PR_TRY(cudaEventRecord(h->ev[1], s));
  if (LPL) {
    PR_TRY(cudaMemcpyAsync(LPL_draws, p.LPL, d * n_ahead * nsave, cudaMemcpyDeviceToHost, s));
    PR_TRY(cudaMemcpyAsync(PL_univariate_draws, p.PLu, d * n_ahead * M * nsave, cudaMemcpyDeviceToHost, s));
  }
  if (p.lpl_sub) PR_TRY(cudaMemcpyAsync(LPL_sub_draws, p.LPLs, d * n_ahead * nsave, cudaMemcpyDeviceToHost, s));
  if (simulate_predictive) PR_TRY(cudaMemcpyAsync(predictions, p.pred, d * n_ahead * M * nsave, cudaMemcpyDeviceToHost, s));
  int st = 0;
  PR_TRY(cudaMemcpyAsync(&st, p.status, sizeof(int), cudaMemcpyDeviceToHost, s));
  long long hp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (p.prof) PR_TRY(cudaMemcpyAsync(hp, p.prof, sizeof(hp), cudaMemcpyDeviceToHost, s));
  PR_TRY(cudaStreamSynchronize(s));
  if (p.prof)
    fprintf(stderr, "predict ticks (CTA 0): logvar %lld phi-power %lld sigma %lld copies+sym %lld gemm %lld mean+copy %lld chol %lld density+draw %lld\n",
            hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7]);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = 1;
#undef PR_TRY
  if (st != 0) {
    bvb_set_error(h, BVB_ERR_NUMERIC, "chol(): decomposition failed in the predictive covariance of draw %d", (st - 1) % R + 1);
    return fin(BVB_ERR_NUMERIC);
  }
  return fin(BVB_OK);
}
