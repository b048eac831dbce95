// Generalized inverse Gaussian variates, density  x^(lambda-1) exp(-(chi/x + psi x)/2).
//
// Replaces the reference's do_rgig() (src/sample_coefficients.cpp:16-22,
// src/gig_vec.cpp:30-36), which forwards to the third-party GIGrvg package
// (>= 0.7, source NOT in /root/reference).  This is a restatement of the
// published algorithm GIGrvg implements - Hoermann & Leydold (2014),
// "Generating generalized inverse Gaussian random variates", Stat. Comput. 24:
//   * chi ~ 0  -> gamma limit, psi ~ 0 -> inverse-gamma limit (tolerance 10*eps),
//   * lambda < 0 handled through X ~ GIG(lambda,chi,psi) <=> 1/X ~ GIG(-lambda,psi,chi),
//   * lambda > 2 or omega > 3            : ratio-of-uniforms with mode shift,
//   * lambda >= 1-2.25 omega^2 or omega > 0.2 : ratio-of-uniforms without shift,
//   * otherwise (0 <= lambda < 1, small omega) : three-part hat ("new approach"),
//   * |lambda| = 1/2 exactly: the (reciprocal) inverse Gaussian law, sampled in closed form (Michael et al. 1976).
// Value-level parity with GIGrvg is unpinned (different uniform stream); the tests
// are distributional (KS against scipy.stats.geninvgauss and the gamma limits).
#pragma once
#include "common.cuh"

namespace bvb {

#define BVB_GIG_ZTOL (2.220446049250313e-16 * 10.0)

__host__ __device__ inline double gig_mode(double lambda, double omega) {
  if (lambda >= 1.0) return (sqrt((lambda - 1.0) * (lambda - 1.0) + omega * omega) + (lambda - 1.0)) / omega;
  return omega / (sqrt((1.0 - lambda) * (1.0 - lambda) + omega * omega) + (1.0 - lambda));
}

// Ratio-of-uniforms without shift (lambda <= ~2, moderate omega).
__host__ __device__ inline double gig_rou_noshift(RngStream& g, double lambda, double omega) {
  const double t = 0.5 * (lambda - 1.0), s = 0.25 * omega;
  const double xm = gig_mode(lambda, omega);
  const double nc = t * log(xm) - s * (xm + 1.0 / xm);
  const double ym = ((lambda + 1.0) + sqrt((lambda + 1.0) * (lambda + 1.0) + omega * omega)) / omega;
  const double um = exp(0.5 * (lambda + 1.0) * log(ym) - s * (ym + 1.0 / ym) - nc);
  double X = xm;
  bool ok = false;
  for (int it = 0; it < 100000; ++it) {
    const double U = um * g.uniform();
    const double V = g.uniform();
    X = U / V;
    if (log(V) <= t * log(X) - s * (X + 1.0 / X) - nc) { ok = true; break; }
  }
  return ok ? X : nan("");   // exhausted rejection loop: never hand back an unaccepted candidate (the caller flags NaN)
}

// Ratio-of-uniforms with shift by the mode (lambda > 2 or omega > 3).
__host__ __device__ inline double gig_rou_shift(RngStream& g, double lambda, double omega) {
  const double t = 0.5 * (lambda - 1.0), s = 0.25 * omega;
  const double xm = gig_mode(lambda, omega);
  const double nc = t * log(xm) - s * (xm + 1.0 / xm);
  // extrema of (x - xm) sqrt(f(x)): roots of a cubic (Cardano, trigonometric form)
  const double a = -(2.0 * (lambda + 1.0) / omega + xm);
  const double b = (2.0 * (lambda - 1.0) * xm / omega - 1.0);
  const double c = xm;
  const double p = b - a * a / 3.0;
  const double q = (2.0 * a * a * a) / 27.0 - (a * b) / 3.0 + c;
  double arg = -q / (2.0 * sqrt(-(p * p * p) / 27.0));
  arg = fmin(1.0, fmax(-1.0, arg));
  const double fi = acos(arg);
  const double fak = 2.0 * sqrt(-p / 3.0);
  const double y1 = fak * cos(fi / 3.0) - a / 3.0;
  const double y2 = fak * cos(fi / 3.0 + 4.0 / 3.0 * 3.14159265358979323846) - a / 3.0;
  const double uplus = (y1 - xm) * exp(t * log(y1) - s * (y1 + 1.0 / y1) - nc);
  const double uminus = (y2 - xm) * exp(t * log(y2) - s * (y2 + 1.0 / y2) - nc);
  double X = xm;
  bool ok = false;
  for (int it = 0; it < 100000; ++it) {
    const double U = uminus + g.uniform() * (uplus - uminus);
    const double V = g.uniform();
    X = U / V + xm;
    if (X > 0.0 && log(V) <= t * log(X) - s * (X + 1.0 / X) - nc) { ok = true; break; }
    X = xm;
  }
  return ok ? X : nan("");
}

// Three-part hat for 0 <= lambda < 1 and small omega (Hoermann & Leydold, Alg. 3).  This is the branch the DL / NG /
// R2D2 local scales hit for coefficients shrunk to zero, i.e. almost all of them, so it is written in log space:
// x^y as exp(y log x) with the logs shared (no pow), the hat value carried as its logarithm so the acceptance test
// needs no exp, and log X produced by the inversion itself where possible.  One attempt consumes exactly one Philox
// block (two uniforms), so attempt number a of a draw is a pure function of (stream, counter base + a): the warp form
// below evaluates 32 attempts at once and takes the first accepted one - the same variate the serial loop returns.
struct GigHat {
  double lambda, omega, xm, x0, lm1, hw, lk0, A0, lk1, A1, lk2, A2, x0lam, k1inv, Atot, ea, c2;
};
__host__ __device__ inline GigHat gig_hat_setup(double lambda, double omega) {
  GigHat h;
  h.lambda = lambda; h.omega = omega;
  h.xm = gig_mode(lambda, omega);
  h.x0 = omega / (1.0 - lambda);
  h.lm1 = lambda - 1.0; h.hw = 0.5 * omega;
  h.lk0 = h.lm1 * log(h.xm) - h.hw * (h.xm + 1.0 / h.xm);      // log k0
  h.A0 = exp(h.lk0) * h.x0;
  const double lx0 = log(h.x0), l2w = log(2.0 / omega);
  const bool two = (h.x0 >= 2.0 / omega);                        // no middle part
  h.lk1 = 0.0; h.A1 = 0.0; h.x0lam = 0.0; h.k1inv = 0.0;
  if (two) {
    h.lk2 = h.lm1 * lx0;
    h.A2 = exp(h.lk2 - h.hw * h.x0) * 2.0 / omega;
  } else {
    h.lk1 = -omega;
    const double k1 = exp(h.lk1);
    h.k1inv = 1.0 / k1;
    if (lambda == 0.0) h.A1 = k1 * (2.0 * l2w - 0.6931471805599453);   // k1 log(2 / omega^2)
    else { h.x0lam = exp(lambda * lx0); h.A1 = k1 / lambda * (exp(lambda * l2w) - h.x0lam); }
    h.lk2 = h.lm1 * l2w;
    h.A2 = exp(h.lk2 - 1.0) * 2.0 / omega;
  }
  h.Atot = h.A0 + h.A1 + h.A2;
  const double aa = two ? h.x0 : 2.0 / omega;
  h.ea = exp(-h.hw * aa);
  h.c2 = omega / (2.0 * exp(h.lk2));
  return h;
}
// one attempt from the uniforms (u1, u2); true = accepted, X the candidate
__host__ __device__ inline bool gig_hat_attempt(const GigHat& h, double u1, double u2, double& X) {
  double V = h.Atot * u1;
  double lhx, lX;
  if (V <= h.A0) {
    X = h.x0 * V / h.A0;
    lX = log(X);
    lhx = h.lk0;
  } else {
    V -= h.A0;
    if (V <= h.A1) {
      if (h.lambda == 0.0) lX = log(h.omega) + h.k1inv * V;                   // X = omega exp(exp(omega) V)
      else lX = log(h.x0lam + h.lambda * h.k1inv * V) / h.lambda;             // X = (x0^lambda + lambda V / k1)^(1/lambda)
      X = exp(lX);
      lhx = h.lk1 + h.lm1 * lX;                                               // hat = k1 X^(lambda-1)
    } else {
      V -= h.A1;
      X = -2.0 / h.omega * log(h.ea - h.c2 * V);
      lX = log(X);
      lhx = h.lk2 - h.hw * X;                                                 // hat = k2 exp(-omega X / 2)
    }
  }
  return log(u2) + lhx <= h.lm1 * lX - h.hw * (X + 1.0 / X);
}
__host__ __device__ inline double gig_newapproach(RngStream& g, double lambda, double omega) {
  const GigHat h = gig_hat_setup(lambda, omega);
  g.have = 0;                                          // attempts start on a Philox block boundary
  double X = h.xm;
  bool ok = false;
  for (int it = 0; it < 100000; ++it) {
    const double u1 = g.uniform(), u2 = g.uniform();
    if (gig_hat_attempt(h, u1, u2, X)) { ok = true; break; }
  }
  return ok ? X : nan("");
}
#ifdef __CUDACC__
// Warp form: all 32 lanes hold the SAME stream state; returns the same variate in every lane and leaves the
// stream where the serial loop would have left it.
__device__ inline double gig_newapproach_warp(RngStream& g, double lambda, double omega, int lane) {
  const GigHat h = gig_hat_setup(lambda, omega);
  g.have = 0;
  const uint32_t c0 = g.ctr;
  double X = nan("");   // stays NaN if no attempt is accepted
  for (int round = 0; round < 3200; ++round) {
    RngStream a = g;
    a.ctr = c0 + (uint32_t)(round * 32 + lane);
    a.have = 0;
    const double u1 = a.uniform(), u2 = a.uniform();
    double Xt;
    const bool ok = gig_hat_attempt(h, u1, u2, Xt);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) {
      const int w = __ffs(m) - 1;
      X = __shfl_sync(0xffffffffu, Xt, w);
      g.ctr = c0 + (uint32_t)(round * 32 + w + 1);
      g.have = 0;
      break;
    }
  }
  return X;
}
#endif

// Inverse Gaussian IG(mean mu, shape sh) by the transformation of Michael, Schucany & Haas (1976), smaller root in
// its cancellation-free form.  GIG(-1/2, chi, psi) = IG(sqrt(chi/psi), chi) exactly, and GIG(1/2, chi, psi) is its
// reciprocal with the roles of chi and psi swapped - the DL psi draw and the R2D2 / NG-exponential-kernel draws are
// of this kind, one normal and one uniform instead of a rejection loop.
__host__ __device__ inline double inverse_gaussian(RngStream& g, double mu, double sh) {
  const double z = g.normal();
  const double w = mu * z * z / (2.0 * sh);
  const double y = mu / (1.0 + w + sqrt(w * (2.0 + w)));
  return (g.uniform() * (mu + y) <= mu) ? y : mu * mu / y;
}

// One GIG(lambda, chi, psi) variate.  Returns NaN for invalid parameters (the
// reference raises an R error there).
__host__ __device__ inline double gig_draw(RngStream& g, double lambda, double chi, double psi) {
  if (!(isfinite(lambda) && isfinite(chi) && isfinite(psi)) || chi < 0.0 || psi < 0.0 ||
      (chi == 0.0 && lambda <= 0.0) || (psi == 0.0 && lambda >= 0.0))
    return nan("");
  if (chi < BVB_GIG_ZTOL) {
    // gamma limit (scale 2/psi); for lambda <= 0 GIGrvg returns the reciprocal form
    return (lambda > 0.0) ? g.gamma(lambda) * (2.0 / psi) : 1.0 / (g.gamma(-lambda) * (2.0 / psi));
  }
  if (psi < BVB_GIG_ZTOL) {
    // inverse-gamma limit (X = 1 / Gamma(|lambda|, scale 2/chi))
    return 1.0 / (g.gamma(fabs(lambda)) * (2.0 / chi));
  }
  if (lambda == -0.5) return inverse_gaussian(g, sqrt(chi / psi), chi);
  if (lambda == 0.5) return 1.0 / inverse_gaussian(g, sqrt(psi / chi), psi);
  const double lambda_old = lambda;
  if (lambda < 0.0) lambda = -lambda;
  const double alpha = sqrt(chi / psi), omega = sqrt(psi * chi);
  double X;
  if (lambda > 2.0 || omega > 3.0) X = gig_rou_shift(g, lambda, omega);
  else if (lambda >= 1.0 - 2.25 * omega * omega || omega > 0.2) X = gig_rou_noshift(g, lambda, omega);
  else X = gig_newapproach(g, lambda, omega);
  return (lambda_old < 0.0) ? alpha / X : alpha * X;
}

#ifdef __CUDACC__
// gig_draw called by the 32 lanes of a warp together, every lane with the same stream state and arguments: the
// rejection loop of the three-part hat (whose worst lane would otherwise set the pace of the whole kernel) runs 32
// attempts at a time; the other branches are simply evaluated redundantly.  Same variate as gig_draw.
__device__ inline double gig_draw_warp(RngStream& g, double lambda, double chi, double psi, int lane) {
  if (!(isfinite(lambda) && isfinite(chi) && isfinite(psi)) || chi < 0.0 || psi < 0.0 ||
      (chi == 0.0 && lambda <= 0.0) || (psi == 0.0 && lambda >= 0.0))
    return nan("");
  if (chi < BVB_GIG_ZTOL || psi < BVB_GIG_ZTOL || lambda == -0.5 || lambda == 0.5) return gig_draw(g, lambda, chi, psi);
  const double la = fabs(lambda);
  const double alpha = sqrt(chi / psi), omega = sqrt(psi * chi);
  if (la > 2.0 || omega > 3.0 || la >= 1.0 - 2.25 * omega * omega || omega > 0.2) return gig_draw(g, lambda, chi, psi);
  const double X = gig_newapproach_warp(g, la, omega, lane);
  return (lambda < 0.0) ? alpha / X : alpha * X;
}
#endif

}  // namespace bvb
