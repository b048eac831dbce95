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

// ystar, mixture indicators (from the current h) and the AWOL normals.
// Series-major scratch: ystar / mixind [s*T + t], eps [s*(T+1) + t].
__global__ void fsv_elem_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int S = d.M + d.r;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)(d.T + 1) * S) return;
  const int s = (int)(idx / (d.T + 1)), t = (int)(idx % (d.T + 1));
  RngStream rng(seed, STREAM_SV_ELEM, (uint32_t)idx, sweep);
  d.eps[idx] = rng.normal();
  if (t >= d.T) return;
  double ys;
  if (s < d.M) {
    double e = d.resid[(size_t)s * d.T + t];
    for (int k = 0; k < d.r; ++k) e -= d.facload[(size_t)k * d.M + s] * d.fac[(size_t)t * d.r + k];
    d.eidi[(size_t)s * d.T + t] = e;
    ys = log(e * e + d.offset[s]);
  } else {
    const double f = d.fac[(size_t)t * d.r + (s - d.M)];
    ys = log(f * f);
  }
  d.ystar[(size_t)s * d.T + t] = ys;
  d.mixind[(size_t)s * d.T + t] = (unsigned char)sv_draw_indicator(ys - d.logvar[(size_t)s * d.T + t], rng.uniform());
}

constexpr int SV_THREADS = 128;

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
__global__ void __launch_bounds__(SV_THREADS) fsv_series_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[8 * (SV_THREADS / 32)];
  __shared__ double bc[4];
  __shared__ SvVariates rvs;
  const uint32_t sweep = *sweep_p;
  const int S = d.M + d.r, T = d.T, tid = threadIdx.x;
  const int s = blockIdx.x;
  double* h = d.logvar + (size_t)s * T;
  if (!d.hetero[s]) {
    if (s < d.M) {
      // homoskedastic idiosyncratic variance: inverse gamma (factorstochvol; bvar_cpp.cpp:728-732 analogue)
      double ss[1] = {0.0};
      for (int t = tid; t < T; t += SV_THREADS) { const double e = d.eidi[(size_t)s * T + t]; ss[0] += e * e; }
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
  double* ys = ep + (T + 1);       // [T]
  unsigned char* rr = reinterpret_cast<unsigned char*>(ys + T);  // [T]
  SvPrior pr;
  pr.mu_mean = d.priormu[0]; pr.mu_sd = d.priormu[1];
  pr.phi_a = d.priorphi[s < d.M ? 0 : 2]; pr.phi_b = d.priorphi[s < d.M ? 1 : 3];
  pr.s2_shape = d.priorsigma2[s]; pr.s2_rate = d.priorsigma2[S + s];
  pr.h0_var = d.priorh0[s];
  pr.mu_const = (s >= d.M) ? 1 : 0;
  double mu = pr.mu_const ? 0.0 : d.sv_para[3 * s + 0], phi = d.sv_para[3 * s + 1], sigma = d.sv_para[3 * s + 2];
  for (int t = tid; t < T; t += SV_THREADS) {
    ys[t] = d.ystar[(size_t)s * T + t];
    rr[t] = d.mixind[(size_t)s * T + t];
  }
  for (int t = tid; t <= T; t += SV_THREADS) ep[t] = d.eps[(size_t)s * (T + 1) + t];
  __syncthreads();
  // (a) tridiagonal system, AWOL draw
  {
    const double s2inv = 1.0 / (sigma * sigma);
    const double B0inv = (pr.h0_var <= 0.0) ? (1.0 - phi * phi) : 1.0 / pr.h0_var;
    for (int t = tid; t <= T; t += SV_THREADS)
      sv_system_row(t, T, t > 0 ? ys[t - 1] : 0.0, t > 0 ? rr[t - 1] : 0, mu, phi, s2inv, B0inv, od[t], c[t]);
    __syncthreads();
    // AWOL draw with a TWISTED factorisation of the tridiagonal precision (eliminate from both ends towards the
    // middle row m, substitute from m outwards): the same law as sv_awol's one-sided Cholesky - any factor F with
    // F F' = Omega maps N(0, I) noise to N(0, Omega^-1) - but two threads (in different warps) each walk half of the
    // dependent chain, which is what the kernel's time is made of (1002 steps x ~70 cycles on one thread).
    {
      const double off = -phi * s2inv;
      const int m = (T + 1) / 2;                       // rows 0..m-1 from the top, T..m+1 from the bottom
      if (tid == 0) {
        double lprev = 0.0, aprev = 0.0;
        double od_n = od[0], c_n = c[0];
        for (int t = 0; t < m; ++t) {
          const double od_t = od_n, c_t = c_n;
          od_n = od[t + 1]; c_n = c[t + 1];
          const double ri = rsqrt(od_t - lprev * lprev);
          const double at = (c_t - lprev * aprev) * ri;
          od[t] = ri; c[t] = at;
          lprev = off * ri;
          lo[t] = lprev;                               // couples row t with row t+1
          aprev = at;
        }
      } else if (tid >= 64 && tid < 69) {
        // Meanwhile warps 2 and 3 draw the variates of the parameter updates (they do not depend on the state): the
        // five normals in lock step on five lanes of warp 2, the three log-uniforms and the gamma on warp 3.
        // Disjoint counter ranges of the series' stream.
        RngStream gq(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
        gq.ctr = 16u * (uint32_t)(tid - 64 + 1);
        const double z = gq.normal();
        double* dst[5] = {&rvs.z1, &rvs.z2, &rvs.nz1, &rvs.nz0, &rvs.nzphi};
        *dst[tid - 64] = z;
      } else if (tid >= 96 && tid < 99) {
        RngStream gq(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
        gq.ctr = 16u * (uint32_t)(tid - 96 + 8);
        const double lu = log(gq.uniform());
        double* dst[3] = {&rvs.lu_s2, &rvs.lu_mh, &rvs.nlu};
        *dst[tid - 96] = lu;
      } else if (tid == 99) {
        RngStream gq(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
        gq.ctr = 4096u;
        rvs.gam = gq.gamma(sv_gamma_shape(T, pr));
      } else if (tid == 32) {
        double lprev = 0.0, aprev = 0.0;
        double od_n = od[T], c_n = c[T];
        for (int t = T; t > m; --t) {
          const double od_t = od_n, c_t = c_n;
          od_n = od[t - 1]; c_n = c[t - 1];
          const double ri = rsqrt(od_t - lprev * lprev);
          const double at = (c_t - lprev * aprev) * ri;
          od[t] = ri; c[t] = at;
          lprev = off * ri;
          lo[t] = lprev;                               // couples row t with row t-1
          aprev = at;
        }
      }
      __syncthreads();
      if (tid == 0) {                                  // middle row: both neighbours are eliminated
        const double lu = (m > 0) ? lo[m - 1] : 0.0, au = (m > 0) ? c[m - 1] : 0.0;
        const double ld = (m < T) ? lo[m + 1] : 0.0, ad = (m < T) ? c[m + 1] : 0.0;
        const double ri = rsqrt(od[m] - lu * lu - ld * ld);
        const double at = (c[m] - lu * au - ld * ad) * ri;
        od[m] = ri; c[m] = at;
        hh[m] = (at + ep[m]) * ri;
      }
      __syncthreads();
      if (tid == 0) {
        double hnext = hh[m];
        for (int t = m - 1; t >= 0; --t) {
          const double ht = (c[t] + ep[t] - lo[t] * hnext) * od[t];
          hh[t] = ht;
          hnext = ht;
        }
      } else if (tid == 32) {
        double hprev = hh[m];
        for (int t = m + 1; t <= T; ++t) {
          const double ht = (c[t] + ep[t] - lo[t] * hprev) * od[t];
          hh[t] = ht;
          hprev = ht;
        }
      }
    }
    __syncthreads();
  }
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
    if (tid < 32) {   // warp 0, all lanes with identical arguments: the lanes share the terms of the MH ratio
      const SvVariates rv = rvs;
      sv_draw_centered(rv, T, cs, hh[0], mu, phi, sigma, pr, rss[0], tid);
      if (tid == 0) { bc[0] = mu; bc[1] = phi; bc[2] = sigma; }
    }
    __syncthreads();
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
    if (tid < 32) {
      const SvVariates rv = rvs;
      SvNSums tot = {v[0], v[1], v[2], v[3], v[4], v[5], v[6]};
      double mu_n, sg_n;
      sv_draw_noncentered(rv, tot, (hh[0] - mu_c) / sg_c, mu_n, sg_n, phi, pr, tid);
      if (tid == 0) { bc[0] = mu_n; bc[1] = sg_n; bc[2] = phi; }
    }
    __syncthreads();
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
  const size_t nel = (size_t)(d.T + 1) * S;
  fsv_elem_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(d, seed, sweep);
  const size_t smem = (size_t)(6 * (d.T + 1)) * sizeof(double) + (size_t)d.T + 16;
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
