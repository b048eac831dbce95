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
// h / (mu,phi,sigma): one thread per series (time-major scratch, coalesced);
// loadings: one warp per row; factors: one thread per t.
#include "fsv.cuh"

namespace bvb {

constexpr uint32_t STREAM_SV_ELEM = 0x100, STREAM_SV_SERIES = 0x101, STREAM_FSV_NG = 0x102,
                   STREAM_FSV_LOAD = 0x103, STREAM_FSV_IW = 0x104, STREAM_FSV_FAC = 0x105;

// ystar, mixture indicators (from the current h) and the AWOL normals.
__global__ void fsv_elem_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int S = d.M + d.r;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)(d.T + 1) * S) return;
  const int s = (int)(idx % S), t = (int)(idx / S);  // time-major
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
  d.ystar[idx] = ys;
  d.mixind[idx] = (unsigned char)sv_draw_indicator(ys - d.logvar[(size_t)s * d.T + t], rng.uniform());
}

__global__ void fsv_series_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int S = d.M + d.r;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  RngStream rng(seed, STREAM_SV_SERIES, (uint32_t)s, sweep);
  double* h = d.logvar + (size_t)s * d.T;
  if (d.hetero[s]) {
    SvPrior pr;
    pr.mu_mean = d.priormu[0]; pr.mu_sd = d.priormu[1];
    pr.phi_a = d.priorphi[s < d.M ? 0 : 2]; pr.phi_b = d.priorphi[s < d.M ? 1 : 3];
    pr.s2_shape = d.priorsigma2[s]; pr.s2_rate = d.priorsigma2[S + s];
    pr.h0_var = d.priorh0[s];
    pr.mu_const = (s >= d.M) ? 1 : 0;
    double mu = d.sv_para[3 * s + 0], phi = d.sv_para[3 * s + 1], sigma = d.sv_para[3 * s + 2], h0 = d.logvar0[s];
    sv_update_series(rng, d.T, d.ystar + s, h, h0, mu, phi, sigma, d.mixind + s, d.eps + s, pr, d.scratch + s, S);
    d.sv_para[3 * s + 0] = mu; d.sv_para[3 * s + 1] = phi; d.sv_para[3 * s + 2] = sigma;
    d.logvar0[s] = h0;
  } else if (s < d.M) {
    // homoskedastic idiosyncratic variance: inverse gamma (factorstochvol; bvar_cpp.cpp:728-732 analogue)
    double ss = 0.0;
    for (int t = 0; t < d.T; ++t) { const double e = d.eidi[(size_t)s * d.T + t]; ss += e * e; }
    const double v = (d.priorhomo[d.M + s] + 0.5 * ss) / rng.gamma(d.priorhomo[s] + 0.5 * d.T);
    const double lv = log(v);
    for (int t = 0; t < d.T; ++t) h[t] = lv;
  }  // homoskedastic factors keep h == 0
}

// Normal-gamma prior on the loadings: lambda2 then tau2 (Kastner 2019, eq. 6-7).
__global__ void fsv_ng_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
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
  d.lambda2[u] = lam2;
  for (int q = 0; q < len; ++q) {
    const int i = d.columnwise ? q : u, j = d.columnwise ? u : q;
    if (d.restr[(size_t)j * d.M + i]) {
      const double l = d.facload[(size_t)j * d.M + i];
      d.tau2[(size_t)j * d.M + i] = gig_draw(rng, a - 0.5, l * l, a * lam2);
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
  const double* h = d.logvar + (size_t)i * d.T;
  for (int t = lane; t < d.T; t += 32) {
    const double w = exp(-h[t]);
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

// Interweaving (Kastner et al. 2017, section 3): one warp per factor.
__global__ void fsv_interweave_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int j = blockIdx.x, lane = threadIdx.x & 31;
  if (j >= d.r || d.interweaving == 0) return;
  const int S = d.M + d.r;
  const bool deep = (d.interweaving == 2 || d.interweaving == 4);
  const bool largest = (d.interweaving >= 3);
  // leading element: diagonal, or the largest absolute loading of the column
  int lead = j;
  if (largest) {
    double best = -1.0;
    for (int i = 0; i < d.M; ++i) {
      const double a = fabs(d.facload[(size_t)j * d.M + i]);
      if (d.restr[(size_t)j * d.M + i] && a > best) { best = a; lead = i; }
    }
  }
  const double Lold = d.facload[(size_t)j * d.M + lead];
  if (!(fabs(Lold) > 0.0)) return;
  // column statistics: psi_s = sum_i (Lambda_ij / Lambda_lj)^2 / tau2_ij, n_j = #unrestricted
  double psis = 0.0, nj = 0.0;
  for (int i = lane; i < d.M; i += 32) {
    if (d.restr[(size_t)j * d.M + i]) {
      const double l = d.facload[(size_t)j * d.M + i] / Lold;
      psis += l * l / d.tau2[(size_t)j * d.M + i];
      nj += 1.0;
    }
  }
  psis = warp_sum(psis);
  nj = warp_sum(nj);
  double* hf = d.logvar + (size_t)(d.M + j) * d.T;
  RngStream rng(seed, STREAM_FSV_IW, (uint32_t)j, sweep);
  double scale = 1.0;  // new Lambda_.j = scale * old, new f_j = old / scale
  bool accept = false;
  double mu_old = 0.0, mu_prop = 0.0;
  if (!deep) {
    // shallow: Lambda_lj^2 | f*, Lambda* ~ GIG((n_j - T)/2, sum_t f*_t^2 e^{-h_t}, psi_s)
    double chi = 0.0;
    for (int t = lane; t < d.T; t += 32) { const double f = d.fac[(size_t)t * d.r + j]; chi += f * f * exp(-hf[t]); }
    chi = warp_sum(chi) * Lold * Lold;
    const double x = gig_draw(rng, 0.5 * (nj - d.T), chi, psis);
    if (isfinite(x) && x > 0.0) { scale = sqrt(x) / fabs(Lold); accept = true; }
  } else if (d.hetero[d.M + j]) {
    // deep: MH on mu* = log Lambda_lj^2 through the level of the factor log-variance
    const double phi = d.sv_para[3 * (d.M + j) + 1], sigma = d.sv_para[3 * (d.M + j) + 2];
    mu_old = log(Lold * Lold);
    const double h0o = d.logvar0[d.M + j] + mu_old;
    double tmph = 0.0;
    for (int t = lane; t < d.T; t += 32) {
      const double prev = (t == 0) ? h0o : hf[t - 1] + mu_old;
      tmph += (hf[t] + mu_old) - phi * prev;
    }
    tmph = warp_sum(tmph);
    const double den = d.T + BVB_SV_B011INV;
    const double gam_old = (1.0 - phi) * mu_old;
    const double gam_prop = tmph / den + sigma / sqrt(den) * rng.normal();
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
  accept = __shfl_sync(0xffffffffu, accept ? 1 : 0, 0) != 0;
  scale = __shfl_sync(0xffffffffu, scale, 0);
  mu_prop = __shfl_sync(0xffffffffu, mu_prop, 0);
  if (!accept) return;
  for (int i = lane; i < d.M; i += 32) d.facload[(size_t)j * d.M + i] *= scale;
  for (int t = lane; t < d.T; t += 32) d.fac[(size_t)t * d.r + j] /= scale;
  if (deep) {
    const double shift = mu_old - mu_prop;
    for (int t = lane; t < d.T; t += 32) hf[t] += shift;
    if (lane == 0) d.logvar0[d.M + j] += shift;
  }
  (void)S;
}

// Factors: T independent r-variate regressions with M observations each.
__global__ void fsv_factors_kernel(const FsvDev d, uint64_t seed, const uint32_t* sweep_p) {
  const uint32_t sweep = *sweep_p;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.T) return;
  const int r = d.r;
  double A[FSV_RMAX * FSV_RMAX], b[FSV_RMAX];
  for (int p = 0; p < r; ++p) { b[p] = 0.0; for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] = 0.0; }
  for (int i = 0; i < d.M; ++i) {
    const double w = exp(-d.logvar[(size_t)i * d.T + t]);
    const double y = d.resid[(size_t)i * d.T + t];
    for (int p = 0; p < r; ++p) {
      const double lw = d.facload[(size_t)p * d.M + i] * w;
      b[p] += lw * y;
      for (int q = 0; q <= p; ++q) A[p * FSV_RMAX + q] += lw * d.facload[(size_t)q * d.M + i];
    }
  }
  for (int p = 0; p < r; ++p) A[p * FSV_RMAX + p] += exp(-d.logvar[(size_t)(d.M + p) * d.T + t]);
  RngStream rng(seed, STREAM_FSV_FAC, (uint32_t)t, sweep);
  if (!small_chol_draw(r, A, b, rng)) d.status[0] = 2;
  for (int p = 0; p < r; ++p) d.fac[(size_t)t * r + p] = b[p];
}

cudaError_t launch_fsv_update(const FsvDev& d, uint64_t seed, const uint32_t* sweep, cudaStream_t s) {
  const int S = d.M + d.r;
  const size_t nel = (size_t)(d.T + 1) * S;
  fsv_elem_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, s>>>(d, seed, sweep);
  fsv_series_kernel<<<(S + 31) / 32, 32, 0, s>>>(d, seed, sweep);
  if (d.r > 0) {
    if (d.ngprior) {
      const int units = d.columnwise ? d.r : d.M;
      fsv_ng_kernel<<<(units + 63) / 64, 64, 0, s>>>(d, seed, sweep);
    }
    fsv_loadings_kernel<<<(d.M + 3) / 4, 128, 0, s>>>(d, seed, sweep);
    if (d.interweaving != 0) fsv_interweave_kernel<<<d.r, 32, 0, s>>>(d, seed, sweep);
    fsv_factors_kernel<<<(d.T + 63) / 64, 64, 0, s>>>(d, seed, sweep);
  }
  return cudaGetLastError();
}

}  // namespace bvb
