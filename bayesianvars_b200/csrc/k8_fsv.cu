// K8 - Sigma_t update on the device: one sweep of the factor stochastic
// volatility model given the VAR residuals (factor structure), or of M
// independent univariate SV models given the orthogonalised residuals
// (cholesky structure).
//
// Replaces factorstochvol::update_fsv (src/bvar_cpp.cpp:617-647) and the loop over
// stochvol::update_fast_sv / the homoscedastic inverse-gamma draw
// (src/bvar_cpp.cpp:728-746).  Both packages are third-party and their source is
// not under /root/reference: this restates the published sampler
// (Kastner, Fruehwirth-Schnatter & Lopes 2017, JCGS 26, sections 2-3; Kastner 2019
// for the normal-gamma prior on the loadings) with the reference's settings:
// row-/column-wise NG prior or fixed tau2, restrictions on the loadings,
// shallow (1,3) or deep (2,4) interweaving through the diagonal / largest loading.
// Parity is unpinned (distributional only).
//
// Parallel shape: indicators + AWOL normals: one thread per (t, series);
// h / (mu,phi,sigma): one CTA per series (shared-memory recursion on one thread,
// sufficient statistics and element-wise parts on all 128);
// loadings: one warp per row; factors: one thread per t.
#include "fsv.cuh"

namespace bvb {

constexpr uint32_t STREAM_SV_ELEM = 0x100, STREAM_SV_SERIES = 0x101, STREAM_FSV_NG = 0x102,
                   STREAM_FSV_LOAD = 0x103, STREAM_FSV_IW = 0x104, STREAM_FSV_FAC = 0x105;

constexpr int SV_THREADS = 512;   // one row of the T+1 = 501 state equations per thread at the headline shape
constexpr int SV_PCR = SV_THREADS - 32;   // threads of the cyclic reduction (the last warp draws the parameter variates meanwhile)

// phase profile of the series kernel (series 0, thread 0; BVB_FSV_PROF=1 prints it when the chain is released)
__device__ long long g_fsv_prof[16];

template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* sh /* [N][SV_THREADS/32] */) {
  // fixed-order (deterministic) block reduction of N values; result broadcast to all threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    v[i] = warp_sum(v[i]);
    if (lane == 0) sh[i * (SV_THREADS / 32) + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SV_THREADS / 32; ++w) t += sh[i * (SV_THREADS / 32) + w];
    v[i] = t;
  }
  __syncthreads();
}

// One CTA per series: the element-wise parts and the sufficient statistics use
// all threads, only the (T+1)-step tridiagonal recursion and the scalar
// parameter draws run on thread 0, out of shared memory.
// "All without a loop" draw of h_{0:T} ~ N(Omega^-1 c, Omega^-1), in parallel, by the first SV_PCR threads of the CTA.
// A factorisation of the tridiagonal precision is a (T+1)-step dependent chain (501 steps x ~70 cycles even split in two
// from both ends: it was the bulk of the series kernel's time, and it does not shrink when series are sharded over GPUs).
// Instead: perturb, then solve.  Omega = B'B + D with B the bidiagonal AR(1) innovation map (row 0: sqrt(B0inv)/sigma;
// row t: (h_t - phi h_{t-1})/sigma) and D = diag(0, 1/v_1, .., 1/v_T) the mixture precisions, so
// eta = B' xi + sqrt(D) zeta with xi ~ N(0, I_{T+1}), zeta ~ N(0, I_T) has covariance Omega and  h = Omega^-1 (c + eta)
// has exactly the law above.  The solve is a parallel cyclic reduction: ceil(log2(T+1)) sweeps in which every row
// eliminates its two neighbours at distance 1, 2, 4, .. (Omega is strictly diagonally dominant: no pivoting).
// sm: 12 (T+1) doubles; rows 0..4 = od, c, lo, hh, ep (ep = xi filled by the caller); e2 = zeta (index 1..T).
// All SV_THREADS threads call this; the threads beyond SV_PCR run `other()` during the reduction.  Result in hh (row 3).
template <class F>
__device__ __forceinline__ void sv_state_draw_parallel(int T, int tid, double* sm, const double* ys, const unsigned char* rr,
                                                       const double* e2, double mu, double phi, double sigma, double h0_var,
                                                       F other, bool prof, long long& tp) {
  double* od = sm;
  double* c = od + (T + 1);
  double* lo = c + (T + 1);
  double* hh = lo + (T + 1);
  double* ep = hh + (T + 1);
  const double s2inv = 1.0 / (sigma * sigma);
  const double B0inv = (h0_var <= 0.0) ? (1.0 - phi * phi) : 1.0 / h0_var;
  for (int t = tid; t <= T; t += SV_THREADS)
    sv_system_row(t, T, t > 0 ? ys[t - 1] : 0.0, t > 0 ? rr[t - 1] : 0, mu, phi, s2inv, B0inv, od[t], c[t]);
  __syncthreads();
  const double off = -phi * s2inv;
  const double sinv = 1.0 / sigma;
  // PCR work arrays (double-buffered): sub-diagonal a, diagonal b (= od), super-diagonal g, rhs r
  double* pa[2] = {lo, sm + 5 * (size_t)(T + 1)};
  double* pb[2] = {od, sm + 6 * (size_t)(T + 1)};
  double* pg[2] = {hh, sm + 7 * (size_t)(T + 1)};
  double* pr_[2] = {c, sm + 8 * (size_t)(T + 1)};
  for (int t = tid; t <= T; t += SV_THREADS) {
    double eta;
    if (t == 0) eta = sqrt(B0inv) * sinv * ep[0] - phi * sinv * ep[1];
    else {
      double p_, m_, v2;
      sv_mix(rr[t - 1], p_, m_, v2);
      eta = sinv * ep[t] - (t < T ? phi * sinv * ep[t + 1] : 0.0) + e2[t] * rsqrt(v2);
    }
    c[t] += eta;
    lo[t] = (t > 0) ? off : 0.0;
    hh[t] = (t < T) ? off : 0.0;
  }
  __syncthreads();
  if (prof) { const long long tn = clock64(); g_fsv_prof[1] += tn - tp; tp = tn; }
  if (tid < SV_PCR) {
    // the diagonal travels with its reciprocal (q): one division per row and sweep
    double* pq[2] = {sm + 10 * (size_t)(T + 1), sm + 11 * (size_t)(T + 1)};
    for (int i = tid; i <= T; i += SV_PCR) pq[0][i] = 1.0 / od[i];
    asm volatile("bar.sync 1, %0;" ::"n"(SV_PCR) : "memory");
    int cur = 0;
    for (int st = 1; st <= T; st <<= 1) {
      const double *a0 = pa[cur], *b0 = pb[cur], *g0 = pg[cur], *r0 = pr_[cur], *q0 = pq[cur];
      double *a1 = pa[cur ^ 1], *b1 = pb[cur ^ 1], *g1 = pg[cur ^ 1], *r1 = pr_[cur ^ 1], *q1 = pq[cur ^ 1];
      for (int i = tid; i <= T; i += SV_PCR) {
        const int im = i - st, ip = i + st;
        const double ai = a0[i], gi = g0[i];
        double bn = b0[i], rn = r0[i], an = 0.0, gn = 0.0;
        if (im >= 0) { const double k1 = ai * q0[im]; bn -= g0[im] * k1; rn -= r0[im] * k1; an = -a0[im] * k1; }
        if (ip <= T) { const double k2 = gi * q0[ip]; bn -= a0[ip] * k2; rn -= r0[ip] * k2; gn = -g0[ip] * k2; }
        a1[i] = an; b1[i] = bn; g1[i] = gn; r1[i] = rn; q1[i] = 1.0 / bn;
      }
      cur ^= 1;
      asm volatile("bar.sync 1, %0;" ::"n"(SV_PCR) : "memory");
    }
    const double *qf = pq[cur], *rf = pr_[cur];
    double* hout = sm + 9 * (size_t)(T + 1);
    for (int i = tid; i <= T; i += SV_PCR) hout[i] = rf[i] * qf[i];
  } else {
    other(tid - SV_PCR);
  }
  __syncthreads();
  {
    const double* hout = sm + 9 * (size_t)(T + 1);
    for (int t = tid; t <= T; t += SV_THREADS) hh[t] = hout[t];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SV_THREADS) fsv_series_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[8 * (SV_THREADS / 32)];
  __shared__ double bc[4];
  __shared__ SvVariates rvs;
  const uint32_t sweep = *sweep_p;
  const int S = d.M + d.r, T = d.T, tid = threadIdx.x;
  const int s = blockIdx.x;
  const bool prof = d.prof && s == 0 && tid == 0;
  long long tp = prof ? clock64() : 0;
#define SV_TICK(slot) do { if (prof) { const long long tn = clock64(); g_fsv_prof[slot] += tn - tp; tp = tn; } } while (0)
  double* h = d.logvar + (size_t)s * T;
  if (!d.hetero[s]) {
    if (s < d.M) {
      // homoskedastic idiosyncratic variance: inverse gamma (factorstochvol; bvar_cpp.cpp:728-732 analogue)
      double ss[1] = {0.0};
      for (int t = tid; t < T; t += SV_THREADS) {
        double e = d.resid[(size_t)s * T + t];
        for (int k = 0; k < d.r; ++k) e -= d.facload[(size_t)k * d.M + s] * d.fac[(size_t)t * d.r + k];
        ss[0] += e * e;
      }
      block_sum<1>(ss, red);
      if (tid == 0) {
        RngStream rng(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
        bc[0] = log((d.priorhomo[d.M + s] + 0.5 * ss[0]) / rng.gamma(d.priorhomo[s] + 0.5 * T));
      }
      __syncthreads();
      for (int t = tid; t < T; t += SV_THREADS) {
        h[t] = bc[0];
        d.wexp[(size_t)s * T + t] = exp(-bc[0]);
        if (d.Wk1) { const double nrm = exp(-0.5 * bc[0]); d.Wk1[(size_t)s * d.Tp + t] = nrm * nrm; }
      }
    }  // homoskedastic factors keep h == 0
    return;
  }
  double* od = sm;                 // [T+1]
  double* c = od + (T + 1);        // [T+1]
  double* lo = c + (T + 1);        // [T+1]
  double* hh = lo + (T + 1);       // [T+1]  hh[0] = h0
  double* ep = hh + (T + 1);       // [T+1]
  double* ys = sm + 12 * (size_t)(T + 1);   // [T]   (sm + 5..11 (T+1): buffers of the cyclic reduction and its result)
  unsigned char* rr = reinterpret_cast<unsigned char*>(ys + T);  // [T]
  SvPrior pr;
  pr.mu_mean = d.priormu[0]; pr.mu_sd = d.priormu[1];
  pr.phi_a = d.priorphi[s < d.M ? 0 : 2]; pr.phi_b = d.priorphi[s < d.M ? 1 : 3];
  pr.s2_shape = d.priorsigma2[s]; pr.s2_rate = d.priorsigma2[S + s];
  pr.h0_var = d.priorh0[s];
  pr.mu_const = (s >= d.M) ? 1 : 0;
  double mu = pr.mu_const ? 0.0 : d.sv_para[3 * s + 0], phi = d.sv_para[3 * s + 1], sigma = d.sv_para[3 * s + 2];
  // y*_t = log(e_t^2 + offset), the mixture indicators given the current h, and the two noise vectors of the state draw
  // (this was a kernel of its own, one thread per (t, series), with its results going through global memory: 10 us plus
  // a launch boundary on the critical path of the sweep; the Philox counters are the same, so is the chain)
  double* e2s = sm + 5 * (size_t)(T + 1);   // [T+1] second noise vector; consumed before the cyclic reduction claims the buffer
  for (int t = tid; t <= T; t += SV_THREADS) {
    const size_t idx = (size_t)s * (T + 1) + t;
    RngStream rng(seed, STREAM_SV_ELEM, (uint32_t)idx, sweep);
    ep[t] = rng.normal();
    e2s[t] = rng.normal();   // the Box-Muller partner
    if (t < T) {
      double ysv;
      if (s < d.M) {
        double e = d.resid[(size_t)s * T + t];
        for (int k = 0; k < d.r; ++k) e -= d.facload[(size_t)k * d.M + s] * d.fac[(size_t)t * d.r + k];
        ysv = log(e * e + d.offset[s]);
      } else {
        const double f = d.fac[(size_t)t * d.r + (s - d.M)];
        ysv = log(f * f);
      }
      ys[t] = ysv;
      rr[t] = (unsigned char)sv_draw_indicator(ysv - h[t], rng.uniform());
    }
  }
  __syncthreads();
  SV_TICK(0);
  // (a) tridiagonal system and state draw (parallel "all without a loop")
  sv_state_draw_parallel(T, tid, sm, ys, rr, e2s, mu, phi, sigma, pr.h0_var, [&](int l3) {
    // last warp: the variates of the parameter updates (they do not depend on the state); disjoint counter ranges
    if (l3 < 5) {
      RngStream gq(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
      gq.ctr = 16u * (uint32_t)(l3 + 1);
      const double z = gq.normal();
      double* dst[5] = {&rvs.z1, &rvs.z2, &rvs.nz1, &rvs.nz0, &rvs.nzphi};
      *dst[l3] = z;
    } else if (l3 >= 8 && l3 < 11) {
      RngStream gq(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
      gq.ctr = 16u * (uint32_t)(l3);
      const double lu = log(gq.uniform());
      double* dst[3] = {&rvs.lu_s2, &rvs.lu_mh, &rvs.nlu};
      *dst[l3 - 8] = lu;
    } else if (l3 == 16) {
      RngStream gq(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
      gq.ctr = 4096u;
      rvs.gam = gq.gamma(sv_gamma_shape(T, pr));
    }
  }, prof, tp);
  SV_TICK(2);
  // (b) centered draw
  {
    double v[5] = {0, 0, 0, 0, 0};
    for (int t = 1 + tid; t <= T; t += SV_THREADS) {
      const double a = hh[t - 1], b = hh[t];
      v[0] += a; v[1] += b; v[2] += a * b; v[3] += a * a; v[4] += b * b;
    }
    block_sum<5>(v, red);
    const SvCSums cs = {v[0], v[1], v[2], v[3], v[4]};
    // residual sum of squares from the residuals themselves (see sv_draw_centered: the raw-sum form cancels)
    double rss[1] = {0.0};
    {
      double b0 = 0.0, b1 = phi, B00, B01, B11;
      if (!pr.mu_const) sv_centered_coef(T, cs, b0, b1, B00, B01, B11);
      for (int t = 1 + tid; t <= T; t += SV_THREADS) { const double e = hh[t] - b0 - b1 * hh[t - 1]; rss[0] += e * e; }
    }
    block_sum<1>(rss, red);
    SV_TICK(3);
    if (tid < 32) {   // warp 0, all lanes with identical arguments: the lanes share the terms of the MH ratio
      const SvVariates rv = rvs;
      sv_draw_centered(rv, T, cs, hh[0], mu, phi, sigma, pr, rss[0], tid);
      if (tid == 0) { bc[0] = mu; bc[1] = phi; bc[2] = sigma; }
    }
    __syncthreads();
    SV_TICK(4);
    mu = bc[0]; phi = bc[1]; sigma = bc[2];
  }
  // (b') non-centered draw (ASIS) and transformation back
  {
    const double mu_c = mu, sg_c = sigma;
    SvNSums ns = {0, 0, 0, 0, 0, 0, 0};
    for (int t = 1 + tid; t <= T; t += SV_THREADS)
      sv_nc_accumulate(ns, (hh[t] - mu_c) / sg_c, (hh[t - 1] - mu_c) / sg_c, ys[t - 1], rr[t - 1]);
    double v[7] = {ns.p00, ns.p01, ns.p11, ns.l0, ns.l1, ns.n_prev2, ns.n_cross};
    block_sum<7>(v, red);
    SV_TICK(5);
    if (tid < 32) {
      const SvVariates rv = rvs;
      SvNSums tot = {v[0], v[1], v[2], v[3], v[4], v[5], v[6]};
      double mu_n, sg_n;
      sv_draw_noncentered(rv, tot, (hh[0] - mu_c) / sg_c, mu_n, sg_n, phi, pr, tid);
      if (tid == 0) { bc[0] = mu_n; bc[1] = sg_n; bc[2] = phi; }
    }
    __syncthreads();
    SV_TICK(6);
    const double mu_n = bc[0], sg_n = bc[1];
    for (int t = 1 + tid; t <= T; t += SV_THREADS) {
      const double hv = mu_n + sg_n * ((hh[t] - mu_c) / sg_c);
      h[t - 1] = hv;
      if (s < d.M) {
        d.wexp[(size_t)s * T + t - 1] = exp(-hv);
        // the pair-GEMM's weight operand of the NEXT sweep, same expression as gibbs_prep_normals_kernel (bit-identical)
        if (d.Wk1) { const double nrm = exp(-0.5 * hv); d.Wk1[(size_t)s * d.Tp + t - 1] = nrm * nrm; }
      }
    }
    if (tid == 0) {
      d.logvar0[s] = mu_n + sg_n * ((hh[0] - mu_c) / sg_c);
      d.sv_para[3 * s + 0] = mu_n; d.sv_para[3 * s + 1] = bc[2]; d.sv_para[3 * s + 2] = fabs(sg_n);
    }
  }
  SV_TICK(7);
  if (prof) g_fsv_prof[15] += 1;
#undef SV_TICK
}

// Test hook (bvb_sv_latent_draw): n_rep independent draws of h_{0:T} | ystar, mixture indicators, (mu, phi, sigma) with
// the series kernel's parallel state draw; replica b uses Philox element index b.  out[b*(T+1) + t].
__global__ void __launch_bounds__(SV_THREADS) sv_latent_test_kernel(int T, const double* ystar, const unsigned char* mixind, double mu,
                                                                    double phi, double sigma, double h0_var, uint64_t seed, double* out) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, b = blockIdx.x;
  double* ep = sm + 4 * (size_t)(T + 1);
  double* ys = sm + 12 * (size_t)(T + 1);
  double* e2 = ys + T;                                                  // [T+1]
  unsigned char* rr = reinterpret_cast<unsigned char*>(e2 + T + 1);
  for (int t = tid; t < T; t += SV_THREADS) { ys[t] = ystar[t]; rr[t] = mixind[t]; }
  for (int t = tid; t <= T; t += SV_THREADS) {
    RngStream rng(seed, STREAM_SV_ELEM, (uint32_t)(b * (T + 1) + t), 0u);
    ep[t] = rng.normal();
    e2[t] = rng.normal();
  }
  __syncthreads();
  long long tp = 0;
  sv_state_draw_parallel(T, tid, sm, ys, rr, e2, mu, phi, sigma, h0_var, [](int) {}, false, tp);
  const double* hh = sm + 3 * (size_t)(T + 1);
  for (int t = tid; t <= T; t += SV_THREADS) out[(size_t)b * (T + 1) + t] = hh[t];
}
cudaError_t launch_sv_latent_test(int T, int n_rep, const double* ystar, const unsigned char* mixind, double mu, double phi, double sigma,
                                  double h0_var, uint64_t seed, double* out, cudaStream_t s) {
  const size_t smem = (size_t)(12 * (T + 1) + T + T + 1) * sizeof(double) + (size_t)T + 16;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) cudaFuncSetAttribute(sv_latent_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sv_latent_test_kernel<<<n_rep, SV_THREADS, smem, s>>>(T, ystar, mixind, mu, phi, sigma, h0_var, seed, out);
  return cudaGetLastError();
}

void fsv_print_profile() {
  long long hp[16];
  if (cudaMemcpyFromSymbol(hp, g_fsv_prof, sizeof(hp)) != cudaSuccess) return;
  if (hp[15] == 0) return;
  const double n = (double)hp[15];
  fprintf(stderr, "fsv_series_kernel (series 0, thread 0, cycles per launch over %lld launches): load %.0f system+eta %.0f cyclic-reduction %.0f "
                  "centered sums %.0f centered draw %.0f non-centered sums %.0f non-centered draw %.0f write-back %.0f\n",
          hp[15], hp[0] / n, hp[1] / n, hp[2] / n, hp[3] / n, hp[4] / n, hp[5] / n, hp[6] / n, hp[7] / n);
}

// Normal-gamma prior on the loadings: lambda2 then tau2 (Kastner 2019, eq. 6-7).
// One WARP per unit (row, or column for the column-wise prior): every lane carries the same stream, the GIG rejection
// loops run 32 attempts at a time (gig_draw_warp) - with one thread per unit the slowest of the M rejection loops
// set the pace: 27 us for 100 draws.
__global__ void __launch_bounds__(128) fsv_ng_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int units = d.columnwise ? d.r : d.M;
  if (u >= units) return;
  RngStream rng(seed, STREAM_FSV_NG, (uint32_t)u, sweep);
  const int len = d.columnwise ? d.M : d.r;
  double sum = 0.0;
  int cnt = 0;
  for (int q = 0; q < len; ++q) {
    const int i = d.columnwise ? q : u, j = d.columnwise ? u : q;
    if (d.restr[(size_t)j * d.M + i]) { sum += d.tau2[(size_t)j * d.M + i]; ++cnt; }
  }
  const double a = d.shrink_a[u];
  const double lam2 = rng.gamma(d.shrink_c[u] + a * cnt) / (d.shrink_d[u] + 0.5 * a * sum);
  if (lane == 0) d.lambda2[u] = lam2;
  for (int q = 0; q < len; ++q) {
    const int i = d.columnwise ? q : u, j = d.columnwise ? u : q;
    if (d.restr[(size_t)j * d.M + i]) {
      const double l = d.facload[(size_t)j * d.M + i];
      const double t2 = gig_draw_warp(rng, a - 0.5, l * l, a * lam2, lane);
      if (lane == 0) d.tau2[(size_t)j * d.M + i] = t2;
    }
  }
}

// In-place Cholesky solve + draw for a small dense SPD system (lower triangle in
// A, row-major n x n with stride FSV_RMAX): x = A^-1 b + L^-T z.
__device__ inline bool small_chol_draw(int n, double* A, double* b, RngStream& rng) {
  bool ok = true;
  for (int c = 0; c < n; ++c) {
    double dcc = A[c * FSV_RMAX + c];
    for (int k = 0; k < c; ++k) dcc -= A[c * FSV_RMAX + k] * A[c * FSV_RMAX + k];
    if (!(dcc > 0.0)) { ok = false; dcc = 1.0; }
    const double l = sqrt(dcc);
    A[c * FSV_RMAX + c] = l;
    for (int i = c + 1; i < n; ++i) {
      double v = A[i * FSV_RMAX + c];
      for (int k = 0; k < c; ++k) v -= A[i * FSV_RMAX + k] * A[c * FSV_RMAX + k];
      A[i * FSV_RMAX + c] = v / l;
    }
  }
  for (int i = 0; i < n; ++i) {  // forward: L y = b
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= A[i * FSV_RMAX + k] * b[k];
    b[i] = v / A[i * FSV_RMAX + i];
  }
  for (int i = 0; i < n; ++i) b[i] += rng.normal();
  for (int i = n - 1; i >= 0; --i) {  // backward: L' x = y + z
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= A[k * FSV_RMAX + i] * b[k];
    b[i] = v / A[i * FSV_RMAX + i];
  }
  return ok;
}

// Loadings: row i regresses its residuals on the unrestricted factors with
// weights exp(-h_it) and prior N(0, tau2_ij).  One warp per row: lanes split t.
__global__ void __launch_bounds__(128) fsv_loadings_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= d.M) return;
  int idx[FSV_RMAX];
  int ri = 0;
  for (int k = 0; k < d.r; ++k) if (d.restr[(size_t)k * d.M + i]) idx[ri++] = k;
  if (ri == 0) return;
  double A[FSV_RMAX * FSV_RMAX], b[FSV_RMAX];
  for (int p = 0; p < ri; ++p) { b[p] = 0.0; for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] = 0.0; }
  const double* y = d.resid + (size_t)i * d.T;
  const double* wx = d.wexp + (size_t)i * d.T;
  for (int t = lane; t < d.T; t += 32) {
    const double w = wx[t];
    const double* f = d.fac + (size_t)t * d.r;
    for (int p = 0; p < ri; ++p) {
      const double fw = f[idx[p]] * w;
      b[p] += fw * y[t];
      for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] += fw * f[idx[q]];
    }
  }
  for (int p = 0; p < ri; ++p) {
    b[p] = warp_sum(b[p]);
    for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] = warp_sum(A[p * FSV_RMAX + q]);
  }
  if (lane == 0) {
    for (int p = 0; p < ri; ++p) A[p * FSV_RMAX + p] += 1.0 / d.tau2[(size_t)idx[p] * d.M + i];
    RngStream rng(seed, STREAM_FSV_LOAD, (uint32_t)i, sweep);
    if (!small_chol_draw(ri, A, b, rng)) d.status[0] = 1;
    for (int p = 0; p < ri; ++p) {
      double l = b[p];
      if (fabs(l) < d.facloadtol) l = (l < 0.0) ? -d.facloadtol : d.facloadtol;  // factor_facloadtol
      d.facload[(size_t)idx[p] * d.M + i] = l;
    }
  }
}

// Interweaving (Kastner et al. 2017, section 3): one CTA (IW_THREADS) per factor; the column / time sums are block
// reductions in a fixed order, the scalar draw runs on thread 0.
constexpr int IW_THREADS = 256;
__device__ __forceinline__ double iw_block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();                                   // sh may still be read from the previous reduction
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < IW_THREADS / 32; ++w) t += sh[w];
  return t;
}
__global__ void __launch_bounds__(IW_THREADS) fsv_interweave_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  __shared__ double sh[IW_THREADS / 32];
  __shared__ double shb[IW_THREADS / 32];
  __shared__ int shi[IW_THREADS / 32];
  __shared__ double bc[3];
  __shared__ int bci[2];
  const uint32_t sweep = *sweep_p;
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (j >= d.r || d.interweaving == 0) return;
  const bool deep = (d.interweaving == 2 || d.interweaving == 4);
  const bool largest = (d.interweaving >= 3);
  // leading element: diagonal, or the largest absolute loading of the column
  int lead = j;
  if (largest) {
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = tid; i < d.M; i += IW_THREADS) {
      const double a = fabs(d.facload[(size_t)j * d.M + i]);
      if (d.restr[(size_t)j * d.M + i] && (a > best || (a == best && i < bi))) { best = a; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // max |loading|, ties -> smallest row (as a serial scan would pick)
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { shb[warp] = best; shi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < IW_THREADS / 32; ++w)
        if (shb[w] > best || (shb[w] == best && shi[w] < bi)) { best = shb[w]; bi = shi[w]; }
      bci[0] = bi;
    }
    __syncthreads();
    if (bci[0] != 0x7fffffff) lead = bci[0];
  }
  const double Lold = d.facload[(size_t)j * d.M + lead];
  if (!(fabs(Lold) > 0.0)) return;                   // uniform: every thread reads the same value
  // column statistics: psi_s = sum_i (Lambda_ij / Lambda_lj)^2 / tau2_ij, n_j = #unrestricted
  double psis = 0.0, nj = 0.0;
  for (int i = tid; i < d.M; i += IW_THREADS) {
    if (d.restr[(size_t)j * d.M + i]) {
      const double l = d.facload[(size_t)j * d.M + i] / Lold;
      psis += l * l / d.tau2[(size_t)j * d.M + i];
      nj += 1.0;
    }
  }
  psis = iw_block_sum(psis, sh);
  nj = iw_block_sum(nj, sh);
  double* hf = d.logvar + (size_t)(d.M + j) * d.T;
  double mu_old = 0.0;
  double stat = 0.0;                                 // shallow: chi; deep: sum_t (h_t + mu) - phi (h_{t-1} + mu)
  const bool hetero = d.hetero[d.M + j] != 0;
  double phi = 0.0, sigma = 0.0, h0o = 0.0;
  if (!deep) {
    for (int t = tid; t < d.T; t += IW_THREADS) { const double f = d.fac[(size_t)t * d.r + j]; stat += f * f * exp(-hf[t]); }
    stat = iw_block_sum(stat, sh) * Lold * Lold;
  } else if (hetero) {
    phi = d.sv_para[3 * (d.M + j) + 1]; sigma = d.sv_para[3 * (d.M + j) + 2];
    mu_old = log(Lold * Lold);
    h0o = d.logvar0[d.M + j] + mu_old;
    for (int t = tid; t < d.T; t += IW_THREADS) {
      const double prev = (t == 0) ? h0o : hf[t - 1] + mu_old;
      stat += (hf[t] + mu_old) - phi * prev;
    }
    stat = iw_block_sum(stat, sh);
  }
  if (tid == 0) {
    RngStream rng(seed, STREAM_FSV_IW, (uint32_t)j, sweep);
    double scale = 1.0, mu_prop = 0.0;   // new Lambda_.j = scale * old, new f_j = old / scale
    bool accept = false;
    if (!deep) {
      // shallow: Lambda_lj^2 | f*, Lambda* ~ GIG((n_j - T)/2, sum_t f*_t^2 e^{-h_t}, psi_s)
      const double x = gig_draw(rng, 0.5 * (nj - d.T), stat, psis);
      if (isfinite(x) && x > 0.0) { scale = sqrt(x) / fabs(Lold); accept = true; }
    } else if (hetero) {
      // deep: MH on mu* = log Lambda_lj^2 through the level of the factor log-variance
      const double den = d.T + BVB_SV_B011INV;
      const double gam_old = (1.0 - phi) * mu_old;
      const double gam_prop = stat / den + sigma / sqrt(den) * rng.normal();
      mu_prop = gam_prop / (1.0 - phi);
      double lr = 0.0;
      if (d.priorh0[d.M + j] <= 0.0) {
        const double sd0 = sigma / sqrt(1.0 - phi * phi);
        lr += sv_log_norm(h0o, mu_prop, sd0) - sv_log_norm(h0o, mu_old, sd0);
      } else {
        const double sd0 = sigma * sqrt(d.priorh0[d.M + j]);
        lr += sv_log_norm(h0o, mu_prop, sd0) - sv_log_norm(h0o, mu_old, sd0);
      }
      // prior of the rescaled column: x = e^mu has density x^{n_j/2 - 1} exp(-x psi_s / 2) dx
      lr += 0.5 * nj * (mu_prop - mu_old) - 0.5 * psis * (exp(mu_prop) - exp(mu_old));
      lr += sv_log_norm(gam_old, 0.0, sigma / sqrt(BVB_SV_B011INV)) - sv_log_norm(gam_prop, 0.0, sigma / sqrt(BVB_SV_B011INV));
      if (log(rng.uniform()) < lr) { accept = true; scale = exp(0.5 * mu_prop) / Lold; }
    }
    bc[0] = scale; bc[1] = mu_prop; bci[1] = accept ? 1 : 0;
  }
  __syncthreads();
  if (!bci[1]) return;
  const double scale = bc[0], mu_prop = bc[1];
  for (int i = tid; i < d.M; i += IW_THREADS) d.facload[(size_t)j * d.M + i] *= scale;
  for (int t = tid; t < d.T; t += IW_THREADS) d.fac[(size_t)t * d.r + j] /= scale;
  if (deep) {
    const double shift = mu_old - mu_prop;
    for (int t = tid; t < d.T; t += IW_THREADS) hf[t] += shift;
    if (tid == 0) d.logvar0[d.M + j] += shift;
  }
}

// Factors: T independent r-variate regressions with M observations each.
// Block = 32 time points x 8 slices of the M rows; slice sums are reduced through
// shared memory and slice 0 solves the r x r system.
constexpr int FAC_SLICES = 8;
__global__ void __launch_bounds__(32 * FAC_SLICES) fsv_factors_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  extern __shared__ double fsm[];   // [FAC_SLICES][nred][32]
  const uint32_t sweep = *sweep_p;
  const int tx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + tx;
  const int r = d.r;
  const int nsl = blockDim.x >> 5;     // row slices (FAC_SLICES unless the reduction buffer would not fit)
  const int nred = r * (r + 1) / 2 + r;
  double A[FSV_RMAX * FSV_RMAX], b[FSV_RMAX];
  for (int p = 0; p < r; ++p) { b[p] = 0.0; for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] = 0.0; }
  if (t < d.T) {
    for (int i = sl; i < d.M; i += nsl) {
      const double w = d.wexp[(size_t)i * d.T + t];
      const double y = d.resid[(size_t)i * d.T + t];
      for (int p = 0; p < r; ++p) {
        const double lw = d.facload[(size_t)p * d.M + i] * w;
        b[p] += lw * y;
        for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] += lw * d.facload[(size_t)q * d.M + i];
      }
    }
  }
  {
    int k = 0;
    for (int p = 0; p < r; ++p) {
      for (int q = 0; q <= p; ++q) fsm[((size_t)sl * nred + k++) * 32 + tx] = A[p * FSV_RMAX + q];
    }
    for (int p = 0; p < r; ++p) fsm[((size_t)sl * nred + k++) * 32 + tx] = b[p];
  }
  __syncthreads();
  if (sl != 0 || t >= d.T) return;
  {
    int k = 0;
    for (int p = 0; p < r; ++p)
      for (int q = 0; q <= p; ++q, ++k) {
        double v = 0.0;
        for (int z = 0; z < nsl; ++z) v += fsm[((size_t)z * nred + k) * 32 + tx];
        A[p * FSV_RMAX + q] = v;
      }
    for (int p = 0; p < r; ++p, ++k) {
      double v = 0.0;
      for (int z = 0; z < nsl; ++z) v += fsm[((size_t)z * nred + k) * 32 + tx];
      b[p] = v;
    }
  }
  for (int p = 0; p < r; ++p) A[p * FSV_RMAX + p] += exp(-d.logvar[(size_t)(d.M + p) * d.T + t]);
  RngStream rng(seed, STREAM_FSV_FAC, (uint32_t)t, sweep);
  if (!small_chol_draw(r, A, b, rng)) d.status[0] = 2;
  for (int p = 0; p < r; ++p) d.fac[(size_t)t * r + p] = b[p];
}


// ---------------------------------------------------------------------------------------------------------------
// r > FSV_RMAX: the r x r normal equations no longer fit a thread's registers.  One CTA per regression (a row of the
// loadings / a time point of the factors); the lower triangle is accumulated entry-parallel in shared memory (fixed
// summation order per entry), factorised column by column by the whole CTA, and thread 0 runs the two O(r^2)
// substitutions with the normals of the same Philox stream the small-r path uses.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BIG_THREADS = 128;
// A: n x n row-major (stride n) lower triangle in shared memory, b: rhs -> draw.  Returns false on a non-positive pivot.
__device__ inline bool block_chol_draw(int n, double* A, double* b, RngStream& rng, int* okflag) {
  const int tid = threadIdx.x;
  if (tid == 0) *okflag = 1;
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    if (tid == 0) {
      double dcc = A[c * n + c];
      if (!(dcc > 0.0)) { *okflag = 0; dcc = 1.0; }
      A[c * n + c] = sqrt(dcc);
    }
    __syncthreads();
    const double l = A[c * n + c];
    for (int i = c + 1 + tid; i < n; i += BIG_THREADS) A[i * n + c] /= l;
    __syncthreads();
    // trailing update of the lower triangle: entry (i, k), c < k <= i
    const int m = n - c - 1;
    for (int e = tid; e < m * (m + 1) / 2; e += BIG_THREADS) {
      int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      while (i * (i + 1) / 2 > e) --i;
      const int k = e - i * (i + 1) / 2;
      const int gi = c + 1 + i, gk = c + 1 + k;
      A[gi * n + gk] -= A[gi * n + c] * A[gk * n + c];
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int i = 0; i < n; ++i) {  // forward: L y = b
      double v = b[i];
      for (int k = 0; k < i; ++k) v -= A[i * n + k] * b[k];
      b[i] = v / A[i * n + i];
    }
    for (int i = 0; i < n; ++i) b[i] += rng.normal();
    for (int i = n - 1; i >= 0; --i) {  // backward: L' x = y + z
      double v = b[i];
      for (int k = i + 1; k < n; ++k) v -= A[k * n + i] * b[k];
      b[i] = v / A[i * n + i];
    }
  }
  __syncthreads();
  return *okflag != 0;
}

__global__ void __launch_bounds__(BIG_THREADS) fsv_loadings_big_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  extern __shared__ double bsm[];   // A[r*r] b[r] idx (ints)
  __shared__ int okflag, ri_s;
  const uint32_t sweep = *sweep_p;
  const int i = blockIdx.x, tid = threadIdx.x, r = d.r;
  double* A = bsm;
  double* b = A + r * r;
  int* idx = reinterpret_cast<int*>(b + r);
  if (tid == 0) {
    int ri = 0;
    for (int k = 0; k < r; ++k) if (d.restr[(size_t)k * d.M + i]) idx[ri++] = k;
    ri_s = ri;
  }
  __syncthreads();
  const int ri = ri_s;
  if (ri == 0) return;
  const double* y = d.resid + (size_t)i * d.T;
  const double* wx = d.wexp + (size_t)i * d.T;
  const int nent = ri * (ri + 1) / 2;
  for (int e = tid; e < nent + ri; e += BIG_THREADS) {
    double acc = 0.0;
    if (e < nent) {
      int p = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((p + 1) * (p + 2) / 2 <= e) ++p;
      while (p * (p + 1) / 2 > e) --p;
      const int q = e - p * (p + 1) / 2;
      const int kp = idx[p], kq = idx[q];
      for (int t = 0; t < d.T; ++t) acc += d.fac[(size_t)t * r + kp] * wx[t] * d.fac[(size_t)t * r + kq];
      if (p == q) acc += 1.0 / d.tau2[(size_t)kp * d.M + i];
      A[p * ri + q] = acc;
    } else {
      const int p = e - nent, kp = idx[p];
      for (int t = 0; t < d.T; ++t) acc += d.fac[(size_t)t * r + kp] * wx[t] * y[t];
      b[p] = acc;
    }
  }
  __syncthreads();
  RngStream rng(seed, STREAM_FSV_LOAD, (uint32_t)i, sweep);
  if (!block_chol_draw(ri, A, b, rng, &okflag) && tid == 0) d.status[0] = 1;
  for (int p = tid; p < ri; p += BIG_THREADS) {
    double l = b[p];
    if (fabs(l) < d.facloadtol) l = (l < 0.0) ? -d.facloadtol : d.facloadtol;
    d.facload[(size_t)idx[p] * d.M + i] = l;
  }
}

__global__ void __launch_bounds__(BIG_THREADS) fsv_factors_big_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  extern __shared__ double bsm[];   // A[r*r] b[r]
  __shared__ int okflag;
  const uint32_t sweep = *sweep_p;
  const int t = blockIdx.x, tid = threadIdx.x, r = d.r;
  double* A = bsm;
  double* b = A + r * r;
  const int nent = r * (r + 1) / 2;
  for (int e = tid; e < nent + r; e += BIG_THREADS) {
    double acc = 0.0;
    if (e < nent) {
      int p = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((p + 1) * (p + 2) / 2 <= e) ++p;
      while (p * (p + 1) / 2 > e) --p;
      const int q = e - p * (p + 1) / 2;
      for (int i = 0; i < d.M; ++i)
        acc += d.facload[(size_t)p * d.M + i] * d.wexp[(size_t)i * d.T + t] * d.facload[(size_t)q * d.M + i];
      if (p == q) acc += exp(-d.logvar[(size_t)(d.M + p) * d.T + t]);
      A[p * r + q] = acc;
    } else {
      const int p = e - nent;
      for (int i = 0; i < d.M; ++i)
        acc += d.facload[(size_t)p * d.M + i] * d.wexp[(size_t)i * d.T + t] * d.resid[(size_t)i * d.T + t];
      b[p] = acc;
    }
  }
  __syncthreads();
  RngStream rng(seed, STREAM_FSV_FAC, (uint32_t)t, sweep);
  if (!block_chol_draw(r, A, b, rng, &okflag) && tid == 0) d.status[0] = 2;
  for (int p = tid; p < r; p += BIG_THREADS) d.fac[(size_t)t * r + p] = b[p];
}

// The NG hyper-parameters of the loadings only read the loadings of the previous sweep: the Gibbs loop may run them on
// its side stream next to the SV series kernel (launch_fsv_ng there, `ng_done` here instead of the launch).
bool fsv_has_ng(const FsvDev& d) { return d.r > 0 && d.ngprior; }
cudaError_t launch_fsv_ng(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s) {
  const int units = d.columnwise ? d.r : d.M;
  fsv_ng_kernel<<<(units + 3) / 4, 128, 0, s>>>(d, seed, sweep);
  return cudaGetLastError();
}

// First half of the update: mixture indicators + AWOL normals, then h and (mu, phi, sigma) of every series.  The
// idiosyncratic log-variances are FINAL after this (the interweaving step only moves the factor log-variances), which is
// what lets the Gibbs loop start the next sweep's pair-GEMM here.
cudaError_t launch_fsv_series(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s) {
  const int S = d.M + d.r;
  const size_t smem = (size_t)(12 * (d.T + 1) + d.T) * sizeof(double) + (size_t)d.T + 16;
  if (smem > 48 * 1024) cudaFuncSetAttribute(fsv_series_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fsv_series_kernel<<<S, SV_THREADS, smem, s>>>(d, seed, sweep);
  return cudaGetLastError();
}

// Second half: NG hyper-parameters (or the event of their side-stream launch), loadings, interweaving, factors.
cudaError_t launch_fsv_rest(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s, cudaEvent_t ng_done) {
  if (d.r > 0) {
    if (d.ngprior) {
      if (ng_done) {
        cudaError_t e = cudaStreamWaitEvent(s, ng_done, 0);
        if (e != cudaSuccess) return e;
      } else {
        const int units = d.columnwise ? d.r : d.M;
        fsv_ng_kernel<<<(units + 3) / 4, 128, 0, s>>>(d, seed, sweep);
      }
    }
    const size_t bigsmem = ((size_t)d.r * d.r + d.r) * sizeof(double) + (size_t)d.r * sizeof(int) + 16;
    if (d.r > FSV_RMAX) {
      if (bigsmem > 48 * 1024) {
        cudaFuncSetAttribute(fsv_loadings_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bigsmem);
        cudaFuncSetAttribute(fsv_factors_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bigsmem);
      }
      fsv_loadings_big_kernel<<<d.M, BIG_THREADS, bigsmem, s>>>(d, seed, sweep);
    } else {
      fsv_loadings_kernel<<<(d.M + 3) / 4, 128, 0, s>>>(d, seed, sweep);
    }
    if (d.interweaving != 0) fsv_interweave_kernel<<<d.r, IW_THREADS, 0, s>>>(d, seed, sweep);
    if (d.r > FSV_RMAX) {
      fsv_factors_big_kernel<<<d.T, BIG_THREADS, bigsmem, s>>>(d, seed, sweep);
    } else {
      const size_t per_slice = (size_t)(d.r * (d.r + 1) / 2 + d.r) * 32 * sizeof(double);
      int nsl = FAC_SLICES;
      while (nsl > 1 && nsl * per_slice > (size_t)200 * 1024) --nsl;      // r > 13: fewer row slices
      const size_t fsmem = nsl * per_slice;
      if (fsmem > 48 * 1024) cudaFuncSetAttribute(fsv_factors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      fsv_factors_kernel<<<(d.T + 31) / 32, 32 * nsl, fsmem, s>>>(d, seed, sweep);
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_fsv_update(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s, cudaEvent_t ng_done) {
  cudaError_t e = launch_fsv_series(d, seed, sweep, s);
  if (e != cudaSuccess) return e;
  return launch_fsv_rest(d, seed, sweep, s, ng_done);
}

}  // namespace bvb
