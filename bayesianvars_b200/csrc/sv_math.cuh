// One sweep of the univariate stochastic-volatility sampler for one series.
//
// Replaces the reference's call into the third-party stochvol package,
//   stochvol::update_fast_sv(log(e^2 + offset), mu, phi, sigma, h0, h, mixind,
//                            prior_spec, expert)           src/bvar_cpp.cpp:742
// (and the same call made per series inside factorstochvol::update_fsv,
// src/bvar_cpp.cpp:617), with the expert settings of src/bvar_cpp.cpp:375-398:
// centered baseline + ASIS interweaving, Gamma(1/2, 1/(2 B_sigma)) prior on
// sigma^2 with independence proposal, normal proposal for phi with immediate
// rejection outside (-1,1), 2-block update when mu is free and a per-parameter
// update when mu is fixed at 0 (factor series).  stochvol's source is NOT under
// /root/reference; this restates the published algorithm:
//   Kastner & Fruehwirth-Schnatter (2014), CSDA 76 - ASIS for SV, sections 2.2-2.4
//   Omori, Chib, Shephard & Nakajima (2007) - 10-component mixture of log chi^2_1
//   McCausland, Miller & Pelletier (2011) - all-without-a-loop draw of h.
// Parity with stochvol is unpinned (distributional only).
//
// The pieces (tridiagonal system, AWOL recursion, parameter draws from
// sufficient statistics) are separate __host__ __device__ functions: the device
// kernel evaluates the sums and the element-wise parts with a whole CTA and runs
// only the recursion and the scalar draws on one thread; sv_update_series()
// composes the same pieces sequentially (host harness / reference semantics).
#pragma once
#include "common.cuh"

namespace bvb {

struct SvPrior {
  double mu_mean, mu_sd;       // mu ~ N(mu_mean, mu_sd^2)          (sv_priormu)
  double phi_a, phi_b;         // (phi+1)/2 ~ Beta(phi_a, phi_b)    (sv_priorphi)
  double s2_shape, s2_rate;    // sigma^2 ~ Gamma(shape, rate)      (sv_priorsigma2) shape is 1/2
  double h0_var;               // <= 0: stationary h0 ~ N(mu, sigma^2/(1-phi^2)); else h0 ~ N(mu, h0_var*sigma^2)
  int mu_const;                // 1: mu fixed at 0 (factor log-variances)
};

#define BVB_SV_B011INV 1e-8
#define BVB_SV_B022INV 1e-12

__host__ __device__ inline void sv_mix(int k, double& p, double& m, double& v2) {
  // Omori et al. (2007), Table 1
  const double P[10] = {0.00609, 0.04775, 0.13057, 0.20674, 0.22715, 0.18842, 0.12047, 0.05591, 0.01575, 0.00115};
  const double Mn[10] = {1.92677, 1.34744, 0.73504, 0.02266, -0.85173, -1.97278, -3.46788, -5.55246, -8.68384, -14.65000};
  const double V2[10] = {0.11265, 0.17788, 0.26768, 0.40611, 0.62699, 0.98583, 1.57469, 2.54498, 4.16591, 7.33342};
  p = P[k]; m = Mn[k]; v2 = V2[k];
}

// P(r_t = k | y*_t - h_t) by inverse-cdf sampling with one uniform.
__host__ __device__ inline int sv_draw_indicator(double resid /* y*_t - h_t */, double u) {
  double w[10], tot = 0.0, best = -1e300;
  double lw[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    double p, m, v2;
    sv_mix(k, p, m, v2);
    const double d = resid - m;
    lw[k] = log(p) - 0.5 * log(v2) - 0.5 * d * d / v2;
    best = fmax(best, lw[k]);
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) { w[k] = exp(lw[k] - best); tot += w[k]; }
  double target = u * tot, cum = 0.0;
  int r = 9;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    cum += w[k];
    if (target < cum) { r = k; break; }
  }
  return r;
}

// Sum of n independent log-density terms.  term(k) only SELECTS the data of term k (what, at which point, with which
// parameters and sign); the arithmetic - logs, divides - is the same code for every k.  lane < 0: one thread evaluates
// them in order (host, tests).  lane >= 0 (device): the 32 lanes of a warp call this together with IDENTICAL arguments,
// lane k evaluates term k and the warp adds them up.  (A lambda that computed its term inside a `switch (k)` made the
// warp run the ten branches one after the other: divergence serialised what was meant to be parallel.)
struct SvTerm {
  int kind;          // 0: normal log-kernel of x with mean m, sd s;  1: Beta((x+1)/2; a = m, b = s) log-kernel;  2: nothing
  double x, m, s, sign;
};
__host__ __device__ inline double sv_eval_term(const SvTerm& t) {
  if (t.kind == 2) return 0.0;
  // both kernels are "c1 log(u1) + c2 log(u2)"-shaped; evaluate with common code: normal -> -log(s) - z^2/2
  const bool beta = t.kind == 1;
  const double u1 = beta ? 0.5 * (1.0 + t.x) : t.s;
  const double u2 = beta ? 0.5 * (1.0 - t.x) : 1.0;
  const double c1 = beta ? (t.m - 1.0) : -1.0;
  const double c2 = beta ? (t.s - 1.0) : 0.0;
  const double z = beta ? 0.0 : (t.x - t.m) / t.s;
  return t.sign * (c1 * log(u1) + c2 * log(u2) - 0.5 * z * z);
}
template <class F>
__host__ __device__ inline double sv_sum_terms(int n, int lane, F term) {
#ifdef __CUDA_ARCH__
  if (lane >= 0) {
    SvTerm t = term(lane < n ? lane : 0);
    if (lane >= n) t.kind = 2;
    double v = sv_eval_term(t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
#endif
  double t = 0.0;
  for (int k = 0; k < n; ++k) t += sv_eval_term(term(k));
  return t;
}

// The random variates one series consumes in its parameter updates.  They do not depend on the state, so the
// device draws them on otherwise idle warps WHILE the AWOL chains run and hands them to the update functions; the
// host (tests) draws them in this order from one stream.
struct SvVariates {
  double gam;      // Gamma(cT, 1) of the sigma^2 proposal (centered)
  double lu_s2;    // log U of the sigma^2 independence-MH step
  double z1, z2;   // N(0,1): phi and gamma proposals (centered)
  double lu_mh;    // log U of the (gamma, phi) MH step
  double nz1, nz0; // N(0,1): sigma and mu draws (non-centered)
  double nzphi;    // N(0,1): phi proposal (non-centered)
  double nlu;      // log U of the non-centered phi MH step
};
__host__ __device__ inline double sv_gamma_shape(int T, const SvPrior& pr) { return pr.mu_const ? 0.5 * T : 0.5 * (T - 1.0); }
__host__ __device__ inline void sv_draw_variates(RngStream& g, int T, const SvPrior& pr, SvVariates& v) {
  v.gam = g.gamma(sv_gamma_shape(T, pr));
  v.lu_s2 = log(g.uniform());
  v.z1 = g.normal();
  v.z2 = g.normal();
  v.lu_mh = log(g.uniform());
  v.nz1 = g.normal();
  v.nz0 = g.normal();
  v.nzphi = g.normal();
  v.nlu = log(g.uniform());
}

__host__ __device__ inline double sv_log_beta_kernel(double phi, double a, double b) {
  // log density kernel of (phi+1)/2 ~ Beta(a, b)
  return (a - 1.0) * log(0.5 * (1.0 + phi)) + (b - 1.0) * log(0.5 * (1.0 - phi));
}
__host__ __device__ inline double sv_log_norm(double x, double mean, double sd) {
  const double z = (x - mean) / sd;
  return -log(sd) - 0.5 * z * z;
}

// Row t (0..T) of the tridiagonal posterior precision of h_{0:T} and of its
// linear term; the off-diagonal is -phi/sigma^2 everywhere.
__host__ __device__ inline void sv_system_row(int t, int T, double ystar_tm1, int r_tm1, double mu, double phi,
                                              double s2inv, double B0inv, double& od, double& c) {
  const double phi2 = phi * phi;
  if (t == 0) {
    od = (B0inv + phi2) * s2inv;
    c = mu * (B0inv - phi * (1.0 - phi)) * s2inv;
  } else {
    double p, m, v2;
    sv_mix(r_tm1, p, m, v2);
    const double prec = 1.0 / v2;
    od = prec + ((t < T) ? (1.0 + phi2) : 1.0) * s2inv;
    c = (ystar_tm1 - m) * prec + mu * ((t < T) ? (1.0 - phi) * (1.0 - phi) : (1.0 - phi)) * s2inv;
  }
}

// AWOL draw: Cholesky of the tridiagonal precision, forward solve, h = L^-T (L^-1 c + eps).
// In: od[0..T], c[0..T], eps[0..T].  Out: hh[0..T] (hh[0] = h0).  od / c / lo are
// overwritten (1/L_diag, forward-solved rhs, L_offdiag).
__host__ __device__ inline void sv_awol(int T, double off, double* od, double* c, double* lo, const double* eps,
                                        double* hh) {
  // The recursion is a (T+1)-step dependent chain; operands of step t+1 are
  // loaded before step t's chain is consumed so memory latency stays off it.
  double lprev = 0.0, aprev = 0.0;
  double od_n = od[0], c_n = c[0];
  for (int t = 0; t <= T; ++t) {
    const double od_t = od_n, c_t = c_n;
    if (t < T) { od_n = od[t + 1]; c_n = c[t + 1]; }
    const double d = od_t - lprev * lprev;    // L_diag^2
    const double ri = bvb_rsqrt(d);
    const double at = (c_t - lprev * aprev) * ri;
    od[t] = ri;
    c[t] = at;
    lprev = off * ri;                         // L_offdiag[t]
    lo[t] = lprev;
    aprev = at;
  }
  double hnext = 0.0;
  double num_n = c[T] + eps[T], lo_n = 0.0, ri_n = od[T];
  for (int t = T; t >= 0; --t) {
    const double num = num_n, lot = lo_n, rit = ri_n;
    if (t > 0) { num_n = c[t - 1] + eps[t - 1]; lo_n = lo[t - 1]; ri_n = od[t - 1]; }
    const double ht = (num - lot * hnext) * rit;
    hh[t] = ht;
    hnext = ht;
  }
}

// Sufficient statistics of the centered AR(1) regression h_t = gamma + phi h_{t-1} + sigma eta.
struct SvCSums { double s_prev, s_cur, s_cross, s_prev2, s_cur2; };

// Centered update of (mu, phi, sigma) given h_{0:T}.
// Posterior mean (b0, b1) of (gamma, phi) in the centered regression h_t = gamma + phi h_{t-1} + sigma eta_t under the
// N(0, sigma^2 diag(1/B011INV, 1/B022INV)) prior, and the posterior covariance factor B.
__host__ __device__ inline void sv_centered_coef(int T, const SvCSums& S, double& b0, double& b1, double& B00, double& B01,
                                                 double& B11) {
  const double x00 = T + BVB_SV_B011INV, x01 = S.s_prev, x11 = S.s_prev2 + BVB_SV_B022INV;
  const double det = x00 * x11 - x01 * x01;
  B00 = x11 / det; B01 = -x01 / det; B11 = x00 / det;
  b0 = B00 * S.s_cur + B01 * S.s_cross;
  b1 = B01 * S.s_cur + B11 * S.s_cross;
}

// `rss`: sum_t (h_t - b0 - b1 h_{t-1})^2 (mu_const: sum_t (h_t - phi h_{t-1})^2) accumulated from the residuals.  The
// scale of the sigma^2 draw is y'y - b'(X'X + B0^-1) b = rss + b' B0^-1 b; formed from the raw sums it cancels
// catastrophically once a series is nearly deterministic (sigma small) and can come out <= 0, which collapsed sigma to
// 1e-150 and the chain with it (seen after ~1900 sweeps of one test chain).  rss < 0 = not supplied: raw-sum form.
__host__ __device__ inline void sv_draw_centered(const SvVariates& rv, int T, const SvCSums& S, double h0, double& mu,
                                                 double& phi, double& sigma, const SvPrior& pr, double rss = -1.0,
                                                 int lane = -1) {
  if (!pr.mu_const) {
    // 2-block: sigma^2 marginally over (gamma, phi), then (gamma, phi) | sigma^2
    double b0, b1, B00, B01, B11;
    sv_centered_coef(T, S, b0, b1, B00, B01, B11);
    double CT = (rss >= 0.0) ? 0.5 * (rss + BVB_SV_B011INV * b0 * b0 + BVB_SV_B022INV * b1 * b1)
                             : 0.5 * (S.s_cur2 - (S.s_cur * b0 + S.s_cross * b1));
    if (!(CT > 0.0)) CT = 1e-300;
    const double s2_old = sigma * sigma;
    const double s2_prop = CT / rv.gam;
    if (rv.lu_s2 < (s2_old - s2_prop) * pr.s2_rate) sigma = sqrt(s2_prop);
    // (gamma, phi) ~ N(b, sigma^2 B)
    const double sg = sigma;
    const double phi_prop = b1 + sg * sqrt(B11) * rv.z1;
    const double cm = b0 + B01 / B11 * (phi_prop - b1);
    double cv = B00 - B01 * B01 / B11;
    if (cv < 0.0) cv = 0.0;
    const double gam_prop = cm + sg * sqrt(cv) * rv.z2;
    if (phi_prop > -1.0 && phi_prop < 1.0) {
      const double gam_old = mu * (1.0 - phi);
      const double mu_prop = gam_prop / (1.0 - phi_prop);
      const double mu_o = mu, phi_o = phi;
      const double sd0p = (pr.h0_var <= 0.0) ? sg / sqrt(1.0 - phi_prop * phi_prop) : sg * sqrt(pr.h0_var);
      const double sd0o = (pr.h0_var <= 0.0) ? sg / sqrt(1.0 - phi_o * phi_o) : sg * sqrt(pr.h0_var);
      const double sdg = sg / sqrt(BVB_SV_B011INV), sdp = sg / sqrt(BVB_SV_B022INV);
      const double lr = sv_sum_terms(10, lane, [&](int k) -> SvTerm {
        switch (k) {
          case 0: return SvTerm{0, h0, mu_prop, sd0p, 1.0};
          case 1: return SvTerm{0, h0, mu_o, sd0o, -1.0};
          case 2: return SvTerm{0, gam_prop, pr.mu_mean * (1.0 - phi_prop), pr.mu_sd * (1.0 - phi_prop), 1.0};
          case 3: return SvTerm{0, gam_old, pr.mu_mean * (1.0 - phi_o), pr.mu_sd * (1.0 - phi_o), -1.0};
          case 4: return SvTerm{1, phi_prop, pr.phi_a, pr.phi_b, 1.0};
          case 5: return SvTerm{1, phi_o, pr.phi_a, pr.phi_b, -1.0};
          case 6: return SvTerm{0, gam_old, 0.0, sdg, 1.0};
          case 7: return SvTerm{0, gam_prop, 0.0, sdg, -1.0};
          case 8: return SvTerm{0, phi_o, 0.0, sdp, 1.0};
          default: return SvTerm{0, phi_prop, 0.0, sdp, -1.0};
        }
      });
      if (rv.lu_mh < lr) { mu = mu_prop; phi = phi_prop; }
    }
  } else {
    // mu fixed at 0: sigma^2 | phi, then phi | sigma^2
    const double stat0 = (pr.h0_var <= 0.0) ? (1.0 - phi * phi) * h0 * h0 : h0 * h0 / pr.h0_var;
    double CT = (rss >= 0.0) ? 0.5 * (rss + stat0)
                             : 0.5 * (S.s_cur2 - 2.0 * phi * S.s_cross + phi * phi * S.s_prev2 + stat0);
    if (!(CT > 0.0)) CT = 1e-300;
    const double s2_old = sigma * sigma;
    const double s2_prop = CT / rv.gam;
    if (rv.lu_s2 < (s2_old - s2_prop) * pr.s2_rate) sigma = sqrt(s2_prop);
    const double sg = sigma;
    const double den = S.s_prev2 + BVB_SV_B022INV;
    const double phi_prop = S.s_cross / den + sg / sqrt(den) * rv.z1;
    if (phi_prop > -1.0 && phi_prop < 1.0) {
      const double phi_o = phi;
      const bool stat0 = pr.h0_var <= 0.0;
      const double sd0p = stat0 ? sg / sqrt(1.0 - phi_prop * phi_prop) : 1.0, sd0o = stat0 ? sg / sqrt(1.0 - phi_o * phi_o) : 1.0;
      const double sdp = sg / sqrt(BVB_SV_B022INV);
      const double lr = sv_sum_terms(6, lane, [&](int k) -> SvTerm {
        switch (k) {
          case 0: return SvTerm{stat0 ? 0 : 2, h0, 0.0, sd0p, 1.0};
          case 1: return SvTerm{stat0 ? 0 : 2, h0, 0.0, sd0o, -1.0};
          case 2: return SvTerm{1, phi_prop, pr.phi_a, pr.phi_b, 1.0};
          case 3: return SvTerm{1, phi_o, pr.phi_a, pr.phi_b, -1.0};
          case 4: return SvTerm{0, phi_o, 0.0, sdp, 1.0};
          default: return SvTerm{0, phi_prop, 0.0, sdp, -1.0};
        }
      });
      if (rv.lu_mh < lr) phi = phi_prop;
    }
  }
}

// Data sums of the non-centered regression y*_t - m_r = mu + sigma htilde_t + eps_t
// (weights 1/v_r^2) and of the AR(1) of htilde.
struct SvNSums { double p00, p01, p11, l0, l1, n_prev2, n_cross; };

__host__ __device__ inline void sv_nc_accumulate(SvNSums& S, double x_t, double x_tm1, double ystar_t, int r_t) {
  double p, m, v2;
  sv_mix(r_t, p, m, v2);
  const double wgt = 1.0 / v2, yv = ystar_t - m;
  S.p00 += wgt; S.p01 += wgt * x_t; S.l0 += wgt * yv;
  S.p11 += wgt * x_t * x_t; S.l1 += wgt * x_t * yv;
  S.n_prev2 += x_tm1 * x_tm1; S.n_cross += x_tm1 * x_t;
}

// Non-centered update: (mu, sigma) | htilde (Gaussian), then phi | htilde.
// Returns the signed sigma draw in sg_n (h = mu + sigma*htilde is invariant to its sign).
__host__ __device__ inline void sv_draw_noncentered(const SvVariates& rv, const SvNSums& S, double ht0, double& mu_n,
                                                    double& sg_n, double& phi, const SvPrior& pr, int lane = -1) {
  const double Bsig_inv = 2.0 * pr.s2_rate;   // sigma^2 ~ Gamma(1/2, rate) <=> sigma ~ N(0, 1/(2 rate))
  if (!pr.mu_const) {
    const double p00 = S.p00 + 1.0 / (pr.mu_sd * pr.mu_sd), p01 = S.p01, p11 = S.p11 + Bsig_inv;
    const double l0 = S.l0 + pr.mu_mean / (pr.mu_sd * pr.mu_sd), l1 = S.l1;
    const double det = p00 * p11 - p01 * p01;
    const double C00 = p11 / det, C01 = -p01 / det, C11 = p00 / det;
    const double m0 = C00 * l0 + C01 * l1, m1 = C01 * l0 + C11 * l1;
    sg_n = m1 + sqrt(C11) * rv.nz1;
    double cv = C00 - C01 * C01 / C11;
    if (cv < 0.0) cv = 0.0;
    mu_n = m0 + C01 / C11 * (sg_n - m1) + sqrt(cv) * rv.nz0;
  } else {
    const double p11 = S.p11 + Bsig_inv;
    mu_n = 0.0;
    sg_n = S.l1 / p11 + rv.nz1 / sqrt(p11);
  }
  const double den = S.n_prev2 + BVB_SV_B022INV;
  const double phi_prop = S.n_cross / den + rv.nzphi / sqrt(den);
  if (phi_prop > -1.0 && phi_prop < 1.0) {
    const double phi_o = phi;
    const bool stat0 = pr.h0_var <= 0.0;
    const double sd0p = stat0 ? 1.0 / sqrt(1.0 - phi_prop * phi_prop) : 1.0, sd0o = stat0 ? 1.0 / sqrt(1.0 - phi_o * phi_o) : 1.0;
    const double sdp = 1.0 / sqrt(BVB_SV_B022INV);
    const double lr = sv_sum_terms(6, lane, [&](int k) -> SvTerm {
      switch (k) {
        case 0: return SvTerm{stat0 ? 0 : 2, ht0, 0.0, sd0p, 1.0};
        case 1: return SvTerm{stat0 ? 0 : 2, ht0, 0.0, sd0o, -1.0};
        case 2: return SvTerm{1, phi_prop, pr.phi_a, pr.phi_b, 1.0};
        case 3: return SvTerm{1, phi_o, pr.phi_a, pr.phi_b, -1.0};
        case 4: return SvTerm{0, phi_o, 0.0, sdp, 1.0};
        default: return SvTerm{0, phi_prop, 0.0, sdp, -1.0};
      }
    });
    if (rv.nlu < lr) phi = phi_prop;
  }
}

// Steps (a) and (b) for one series, sequentially, given the freshly drawn mixture
// indicators (step (c), drawn element-wise with sv_draw_indicator BEFORE this
// call, from the current h, as stochvol orders them) and the T+1 standard normals
// of the AWOL draw.  ystar[t] = log(e_t^2 + offset), t = 0..T-1; h[0..T-1] are
// h_1..h_T; h0 the initial state; eps[0..T]; scratch: 4*(T+1) doubles.
__host__ __device__ inline void sv_update_series(RngStream& g, int T, const double* ystar, double* h, double& h0,
                                                 double& mu, double& phi, double& sigma, const unsigned char* r,
                                                 const double* eps, const SvPrior& pr, double* scratch) {
  if (pr.mu_const) mu = 0.0;
  double* od = scratch;
  double* c = scratch + (T + 1);
  double* lo = scratch + 2 * (T + 1);
  double* hh = scratch + 3 * (T + 1);
  const double s2inv = 1.0 / (sigma * sigma);
  const double B0inv = (pr.h0_var <= 0.0) ? (1.0 - phi * phi) : 1.0 / pr.h0_var;
  for (int t = 0; t <= T; ++t)
    sv_system_row(t, T, t > 0 ? ystar[t - 1] : 0.0, t > 0 ? r[t - 1] : 0, mu, phi, s2inv, B0inv, od[t], c[t]);
  sv_awol(T, -phi * s2inv, od, c, lo, eps, hh);
  SvCSums cs = {0, 0, 0, 0, 0};
  for (int t = 1; t <= T; ++t) {
    cs.s_prev += hh[t - 1]; cs.s_prev2 += hh[t - 1] * hh[t - 1];
    cs.s_cur += hh[t]; cs.s_cur2 += hh[t] * hh[t];
    cs.s_cross += hh[t - 1] * hh[t];
  }
  double rss = 0.0;
  if (!pr.mu_const) {
    double b0, b1, B00, B01, B11;
    sv_centered_coef(T, cs, b0, b1, B00, B01, B11);
    for (int t = 1; t <= T; ++t) { const double e = hh[t] - b0 - b1 * hh[t - 1]; rss += e * e; }
  } else {
    for (int t = 1; t <= T; ++t) { const double e = hh[t] - phi * hh[t - 1]; rss += e * e; }
  }
  SvVariates rv;
  sv_draw_variates(g, T, pr, rv);
  sv_draw_centered(rv, T, cs, hh[0], mu, phi, sigma, pr, rss);
  const double mu_c = mu, sg_c = sigma;
  SvNSums ns = {0, 0, 0, 0, 0, 0, 0};
  for (int t = 1; t <= T; ++t)
    sv_nc_accumulate(ns, (hh[t] - mu_c) / sg_c, (hh[t - 1] - mu_c) / sg_c, ystar[t - 1], r[t - 1]);
  double mu_n, sg_n;
  sv_draw_noncentered(rv, ns, (hh[0] - mu_c) / sg_c, mu_n, sg_n, phi, pr);
  for (int t = 1; t <= T; ++t) h[t - 1] = mu_n + sg_n * ((hh[t] - mu_c) / sg_c);
  h0 = mu_n + sg_n * ((hh[0] - mu_c) / sg_c);
  mu = mu_n;
  sigma = fabs(sg_n);
}

}  // namespace bvb
