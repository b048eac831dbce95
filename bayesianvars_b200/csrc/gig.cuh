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
  for (int it = 0; it < 100000; ++it) {
    const double U = um * g.uniform();
    const double V = g.uniform();
    X = U / V;
    if (log(V) <= t * log(X) - s * (X + 1.0 / X) - nc) break;
  }
  return X;
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
  for (int it = 0; it < 100000; ++it) {
    const double U = uminus + g.uniform() * (uplus - uminus);
    const double V = g.uniform();
    X = U / V + xm;
    if (X > 0.0 && log(V) <= t * log(X) - s * (X + 1.0 / X) - nc) break;
    X = xm;
  }
  return X;
}

// Three-part hat for 0 <= lambda < 1 and small omega (Hoermann & Leydold, Alg. 3).  This is the branch the DL / NG /
// R2D2 local scales hit for coefficients shrunk to zero, i.e. almost all of them, so it is written in log space:
// x^y as exp(y log x) with the logs shared (no pow), the hat value carried as its logarithm so the acceptance test
// needs no exp, and log X produced by the inversion itself where possible.  Two uniforms per attempt as before.
__host__ __device__ inline double gig_newapproach(RngStream& g, double lambda, double omega) {
  const double xm = gig_mode(lambda, omega);
  const double x0 = omega / (1.0 - lambda);
  const double lm1 = lambda - 1.0, hw = 0.5 * omega;
  const double lk0 = lm1 * log(xm) - hw * (xm + 1.0 / xm);      // log k0
  const double A0 = exp(lk0) * x0;
  const double lx0 = log(x0), l2w = log(2.0 / omega);
  const bool two = (x0 >= 2.0 / omega);                          // no middle part
  double lk1 = 0.0, A1 = 0.0, lk2, A2, x0lam = 0.0, k1inv = 0.0;
  if (two) {
    lk2 = lm1 * lx0;
    A2 = exp(lk2 - hw * x0) * 2.0 / omega;
  } else {
    lk1 = -omega;
    const double k1 = exp(lk1);
    k1inv = 1.0 / k1;
    if (lambda == 0.0) A1 = k1 * (2.0 * l2w - 0.6931471805599453);   // k1 log(2 / omega^2)
    else { x0lam = exp(lambda * lx0); A1 = k1 / lambda * (exp(lambda * l2w) - x0lam); }
    lk2 = lm1 * l2w;
    A2 = exp(lk2 - 1.0) * 2.0 / omega;
  }
  const double Atot = A0 + A1 + A2;
  const double aa = two ? x0 : 2.0 / omega;
  const double ea = exp(-hw * aa), c2 = omega / (2.0 * exp(lk2));
  double X = xm;
  for (int it = 0; it < 100000; ++it) {
    double V = Atot * g.uniform();
    double lhx, lX;
    if (V <= A0) {
      X = x0 * V / A0;
      lX = log(X);
      lhx = lk0;
    } else {
      V -= A0;
      if (V <= A1) {
        if (lambda == 0.0) {
          lX = log(omega) + k1inv * V;                          // X = omega exp(exp(omega) V)
          X = exp(lX);
        } else {
          lX = log(x0lam + lambda * k1inv * V) / lambda;          // X = (x0^lambda + lambda V / k1)^(1/lambda)
          X = exp(lX);
        }
        lhx = lk1 + lm1 * lX;                                     // hat = k1 X^(lambda-1)
      } else {
        V -= A1;
        X = -2.0 / omega * log(ea - c2 * V);
        lX = log(X);
        lhx = lk2 - hw * X;                                       // hat = k2 exp(-omega X / 2)
      }
    }
    if (log(g.uniform()) + lhx <= lm1 * lX - hw * (X + 1.0 / X)) break;
  }
  return X;
}

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

}  // namespace bvb
