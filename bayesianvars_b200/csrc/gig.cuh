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
//   * otherwise (0 <= lambda < 1, small omega) : three-part hat ("new approach").
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

// Three-part hat for 0 <= lambda < 1 and small omega (Hoermann & Leydold, Alg. 3).
__host__ __device__ inline double gig_newapproach(RngStream& g, double lambda, double omega) {
  const double xm = gig_mode(lambda, omega);
  const double x0 = omega / (1.0 - lambda);
  const double k0 = exp((lambda - 1.0) * log(xm) - 0.5 * omega * (xm + 1.0 / xm));
  const double A0 = k0 * x0;
  double k1, A1, k2, A2;
  if (x0 >= 2.0 / omega) {
    k1 = 0.0;
    A1 = 0.0;
    k2 = pow(x0, lambda - 1.0);
    A2 = k2 * 2.0 * exp(-omega * x0 / 2.0) / omega;
  } else {
    k1 = exp(-omega);
    A1 = (lambda == 0.0) ? k1 * log(2.0 / (omega * omega))
                         : k1 / lambda * (pow(2.0 / omega, lambda) - pow(x0, lambda));
    k2 = pow(2.0 / omega, lambda - 1.0);
    A2 = k2 * 2.0 * exp(-1.0) / omega;
  }
  const double Atot = A0 + A1 + A2;
  double X = xm;
  for (int it = 0; it < 100000; ++it) {
    double V = Atot * g.uniform();
    double hx;
    if (V <= A0) {
      X = x0 * V / A0;
      hx = k0;
    } else {
      V -= A0;
      if (V <= A1) {
        if (lambda == 0.0) {
          X = omega * exp(exp(omega) * V);
          hx = k1 / X;
        } else {
          X = pow(pow(x0, lambda) + (lambda / k1 * V), 1.0 / lambda);
          hx = k1 * pow(X, lambda - 1.0);
        }
      } else {
        V -= A1;
        const double a = (x0 > 2.0 / omega) ? x0 : 2.0 / omega;
        X = -2.0 / omega * log(exp(-omega / 2.0 * a) - omega / (2.0 * k2) * V);
        hx = k2 * exp(-omega / 2.0 * X);
      }
    }
    const double U = g.uniform() * hx;
    if (log(U) <= (lambda - 1.0) * log(X) - omega / 2.0 * (X + 1.0 / X)) break;
  }
  return X;
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
