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

__host__ __device__ inline double sv_log_beta_kernel(double phi, double a, double b) {
  // log density kernel of (phi+1)/2 ~ Beta(a, b)
  return (a - 1.0) * log(0.5 * (1.0 + phi)) + (b - 1.0) * log(0.5 * (1.0 - phi));
}
__host__ __device__ inline double sv_log_norm(double x, double mean, double sd) {
  const double z = (x - mean) / sd;
  return -log(sd) - 0.5 * z * z;
}

// Steps (a) and (b) for one series, given the freshly drawn mixture indicators
// (step (c), drawn element-wise with sv_draw_indicator BEFORE this call, from the
// current h, as stochvol orders them) and the T+1 standard normals of the AWOL
// draw.  ystar[t] = log(e_t^2 + offset), t = 0..T-1.  h[0..T-1] are h_1..h_T; h0 is
// the initial state.  eps[0..T]: normals.  scratch: 3*(T+1) doubles.
// `ld` is the element stride of ystar / r / eps / scratch (1 on the host; the
// number of series on the device, where those arrays are time-major so that the
// threads of a warp - one series each - touch consecutive addresses); h has
// stride 1 (a column of the T x (M+r) logvar matrix).
__host__ __device__ inline void sv_update_series(RngStream& g, int T, const double* ystar_, double* h,
                                                 double& h0, double& mu, double& phi, double& sigma,
                                                 const unsigned char* r_, const double* eps_,
                                                 const SvPrior& pr, double* scratch_, int ld) {
  if (pr.mu_const) mu = 0.0;
#define YSTAR(t) ystar_[(size_t)(t) * ld]
#define RIND(t) r_[(size_t)(t) * ld]
#define EPS(t) eps_[(size_t)(t) * ld]
#define RINV(t) scratch_[(size_t)(t) * ld]
#define LOFF(t) scratch_[((size_t)(T + 1) + (t)) * ld]
#define AFWD(t) scratch_[((size_t)2 * (T + 1) + (t)) * ld]

  // ---- (a) h_{0:T} | r, theta: tridiagonal precision, Cholesky, AWOL draw --------
  {
    const double s2inv = 1.0 / (sigma * sigma), phi2 = phi * phi;
    const double B0inv = (pr.h0_var <= 0.0) ? (1.0 - phi2) : 1.0 / pr.h0_var;
    const double off = -phi * s2inv;
    double lprev = 0.0, aprev = 0.0;
    for (int t = 0; t <= T; ++t) {
      double od, c;
      if (t == 0) {
        od = (B0inv + phi2) * s2inv;
        c = mu * (B0inv - phi * (1.0 - phi)) * s2inv;
      } else {
        double p, m, v2;
        sv_mix(RIND(t - 1), p, m, v2);
        const double prec = 1.0 / v2;
        od = prec + ((t < T) ? (1.0 + phi2) : 1.0) * s2inv;
        c = (YSTAR(t - 1) - m) * prec + mu * ((t < T) ? (1.0 - phi) * (1.0 - phi) : (1.0 - phi)) * s2inv;
      }
      const double d = (t == 0) ? od : od - lprev * lprev;   // L_diag^2
      const double ri = bvb_rsqrt(d);
      const double at = (c - lprev * aprev) * ri;
      RINV(t) = ri;
      AFWD(t) = at;
      lprev = off * ri;       // L_offdiag[t] = Omega_offdiag / L_diag[t]
      LOFF(t) = lprev;
      aprev = at;
    }
    double hnext = 0.0;
    for (int t = T; t >= 0; --t) {
      const double ht = (AFWD(t) + EPS(t) - ((t < T) ? LOFF(t) * hnext : 0.0)) * RINV(t);
      if (t == 0) h0 = ht; else h[t - 1] = ht;
      hnext = ht;
    }
  }

  // ---- (b) theta = (mu, phi, sigma): centered, then non-centered (ASIS) ----------
  // -- centered --
  {
    double s_prev = h0, s_cur = 0.0, s_cross = 0.0, s_prev2 = h0 * h0, s_cur2 = 0.0;
    s_cross = h0 * h[0];
    for (int t = 0; t < T; ++t) {
      s_cur += h[t];
      s_cur2 += h[t] * h[t];
      if (t < T - 1) { s_prev += h[t]; s_prev2 += h[t] * h[t]; s_cross += h[t] * h[t + 1]; }
    }
    // s_prev = sum h_{t-1}, s_cur = sum h_t, s_cross = sum h_{t-1} h_t, s_prev2 = sum h_{t-1}^2
    if (!pr.mu_const) {
      // 2-block: sigma^2 marginally over (gamma, phi), then (gamma, phi) | sigma^2
      const double x00 = T + BVB_SV_B011INV, x01 = s_prev, x11 = s_prev2 + BVB_SV_B022INV;
      const double det = x00 * x11 - x01 * x01;
      const double B00 = x11 / det, B01 = -x01 / det, B11 = x00 / det;
      const double b0 = B00 * s_cur + B01 * s_cross, b1 = B01 * s_cur + B11 * s_cross;
      const double cT = 0.5 * (T - 1.0);
      double CT = 0.5 * (s_cur2 - (s_cur * b0 + s_cross * b1));
      if (!(CT > 0.0)) CT = 1e-300;
      const double s2_old = sigma * sigma;
      const double s2_prop = CT / g.gamma(cT);
      if (log(g.uniform()) < (s2_old - s2_prop) * pr.s2_rate) sigma = sqrt(s2_prop);
      // (gamma, phi) ~ N(b, sigma^2 B)
      const double sg = sigma;
      const double phi_prop = b1 + sg * sqrt(B11) * g.normal();
      const double cm = b0 + B01 / B11 * (phi_prop - b1);
      double cv = B00 - B01 * B01 / B11;
      if (cv < 0.0) cv = 0.0;
      const double gam_prop = cm + sg * sqrt(cv) * g.normal();
      const double u = g.uniform();
      if (phi_prop > -1.0 && phi_prop < 1.0) {
        const double gam_old = mu * (1.0 - phi);
        const double mu_prop = gam_prop / (1.0 - phi_prop);
        double lr = 0.0;
        if (pr.h0_var <= 0.0) {
          lr += sv_log_norm(h0, mu_prop, sg / sqrt(1.0 - phi_prop * phi_prop)) -
                sv_log_norm(h0, mu, sg / sqrt(1.0 - phi * phi));
        } else {
          lr += sv_log_norm(h0, mu_prop, sg * sqrt(pr.h0_var)) - sv_log_norm(h0, mu, sg * sqrt(pr.h0_var));
        }
        lr += sv_log_norm(gam_prop, pr.mu_mean * (1.0 - phi_prop), pr.mu_sd * (1.0 - phi_prop)) -
              sv_log_norm(gam_old, pr.mu_mean * (1.0 - phi), pr.mu_sd * (1.0 - phi));
        lr += sv_log_beta_kernel(phi_prop, pr.phi_a, pr.phi_b) - sv_log_beta_kernel(phi, pr.phi_a, pr.phi_b);
        lr += sv_log_norm(gam_old, 0.0, sg / sqrt(BVB_SV_B011INV)) - sv_log_norm(gam_prop, 0.0, sg / sqrt(BVB_SV_B011INV));
        lr += sv_log_norm(phi, 0.0, sg / sqrt(BVB_SV_B022INV)) - sv_log_norm(phi_prop, 0.0, sg / sqrt(BVB_SV_B022INV));
        if (log(u) < lr) { mu = mu_prop; phi = phi_prop; }
      }
    } else {
      // mu fixed at 0: sigma^2 | phi, then phi | sigma^2
      const double stat0 = (pr.h0_var <= 0.0) ? (1.0 - phi * phi) * h0 * h0 : h0 * h0 / pr.h0_var;
      double CT = 0.5 * (s_cur2 - 2.0 * phi * s_cross + phi * phi * s_prev2 + stat0);
      if (!(CT > 0.0)) CT = 1e-300;
      const double cT = 0.5 * T;
      const double s2_old = sigma * sigma;
      const double s2_prop = CT / g.gamma(cT);
      if (log(g.uniform()) < (s2_old - s2_prop) * pr.s2_rate) sigma = sqrt(s2_prop);
      const double sg = sigma;
      const double den = s_prev2 + BVB_SV_B022INV;
      const double phi_prop = s_cross / den + sg / sqrt(den) * g.normal();
      const double u = g.uniform();
      if (phi_prop > -1.0 && phi_prop < 1.0) {
        double lr = 0.0;
        if (pr.h0_var <= 0.0)
          lr += sv_log_norm(h0, 0.0, sg / sqrt(1.0 - phi_prop * phi_prop)) - sv_log_norm(h0, 0.0, sg / sqrt(1.0 - phi * phi));
        lr += sv_log_beta_kernel(phi_prop, pr.phi_a, pr.phi_b) - sv_log_beta_kernel(phi, pr.phi_a, pr.phi_b);
        lr += sv_log_norm(phi, 0.0, sg / sqrt(BVB_SV_B022INV)) - sv_log_norm(phi_prop, 0.0, sg / sqrt(BVB_SV_B022INV));
        if (log(u) < lr) phi = phi_prop;
      }
    }
  }
  // -- non-centered: htilde = (h - mu)/sigma --
  {
    const double mu_c = mu, sg_c = sigma;
    // (mu, sigma) | htilde, r, y  (Gaussian regression, prior mu ~ N(b,B), sigma ~ N(0, B_sigma))
    const double Bsig_inv = 2.0 * pr.s2_rate;   // sigma^2 ~ Gamma(1/2, rate) <=> sigma ~ N(0, 1/(2 rate))
    double p00 = pr.mu_const ? 1.0 : 1.0 / (pr.mu_sd * pr.mu_sd), p01 = 0.0, p11 = Bsig_inv;
    double l0 = pr.mu_const ? 0.0 : pr.mu_mean / (pr.mu_sd * pr.mu_sd), l1 = 0.0;
    double n_prev2 = 0.0, n_cross = 0.0;
    const double ht0 = (h0 - mu_c) / sg_c;
    double hprev = ht0;
    for (int t = 0; t < T; ++t) {
      double p, m, v2;
      sv_mix(RIND(t), p, m, v2);
      const double wgt = 1.0 / v2, x = (h[t] - mu_c) / sg_c, yv = YSTAR(t) - m;
      if (!pr.mu_const) { p00 += wgt; p01 += wgt * x; l0 += wgt * yv; }
      p11 += wgt * x * x;
      l1 += wgt * x * yv;
      n_prev2 += hprev * hprev;
      n_cross += hprev * x;
      hprev = x;
    }
    double mu_n, sg_n;
    if (!pr.mu_const) {
      const double det = p00 * p11 - p01 * p01;
      const double C00 = p11 / det, C01 = -p01 / det, C11 = p00 / det;
      const double m0 = C00 * l0 + C01 * l1, m1 = C01 * l0 + C11 * l1;
      const double z1 = g.normal(), z0 = g.normal();
      sg_n = m1 + sqrt(C11) * z1;
      double cv = C00 - C01 * C01 / C11;
      if (cv < 0.0) cv = 0.0;
      mu_n = m0 + C01 / C11 * (sg_n - m1) + sqrt(cv) * z0;
    } else {
      mu_n = 0.0;
      sg_n = l1 / p11 + g.normal() / sqrt(p11);
    }
    // phi | htilde
    const double den = n_prev2 + BVB_SV_B022INV;
    const double phi_prop = n_cross / den + g.normal() / sqrt(den);
    const double u = g.uniform();
    if (phi_prop > -1.0 && phi_prop < 1.0) {
      double lr = 0.0;
      if (pr.h0_var <= 0.0)
        lr += sv_log_norm(ht0, 0.0, 1.0 / sqrt(1.0 - phi_prop * phi_prop)) - sv_log_norm(ht0, 0.0, 1.0 / sqrt(1.0 - phi * phi));
      lr += sv_log_beta_kernel(phi_prop, pr.phi_a, pr.phi_b) - sv_log_beta_kernel(phi, pr.phi_a, pr.phi_b);
      lr += sv_log_norm(phi, 0.0, 1.0 / sqrt(BVB_SV_B022INV)) - sv_log_norm(phi_prop, 0.0, 1.0 / sqrt(BVB_SV_B022INV));
      if (log(u) < lr) phi = phi_prop;
    }
    // back to the centered states: h = mu + sigma * htilde is invariant to the sign of sigma
    for (int t = 0; t < T; ++t) h[t] = mu_n + sg_n * ((h[t] - mu_c) / sg_c);
    h0 = mu_n + sg_n * ht0;
    mu = mu_n;
    sigma = fabs(sg_n);
  }
#undef YSTAR
#undef RIND
#undef EPS
#undef RINV
#undef LOFF
#undef AFWD
}

}  // namespace bvb
