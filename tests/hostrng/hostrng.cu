// Host-side harness around the __host__ __device__ RNG code (Philox stream,
// normal/gamma/GIG samplers) so its LOGIC can be tested distributionally without
// a GPU.  Test tool only: the product draws these variates on the device.
#include "../../bayesianvars_b200/csrc/gig.cuh"
#include "../../bayesianvars_b200/csrc/sv_math.cuh"
#include <vector>

extern "C" {
void hostrng_gig(int n, double lambda, double chi, double psi, unsigned long long seed, double* out) {
  for (int i = 0; i < n; ++i) {
    bvb::RngStream g(seed, 7u, (uint32_t)i, 0u);
    out[i] = bvb::gig_draw(g, lambda, chi, psi);
  }
}
void hostrng_gamma(int n, double shape, unsigned long long seed, double* out) {
  for (int i = 0; i < n; ++i) {
    bvb::RngStream g(seed, 8u, (uint32_t)i, 0u);
    out[i] = g.gamma(shape);
  }
}
void hostrng_normal(int n, unsigned long long seed, double* out) {
  for (int i = 0; i < n; ++i) {
    bvb::RngStream g(seed, 9u, (uint32_t)(i / 2), 0u);
    double a = g.normal(), b = g.normal();
    out[i] = (i & 1) ? b : a;
  }
}
void hostrng_uniform(int n, unsigned long long seed, double* out) {
  bvb::RngStream g(seed, 10u, 0u, 0u);
  for (int i = 0; i < n; ++i) out[i] = g.uniform();
}

// n_sweeps of the univariate SV sampler on one series (indicators drawn first, as
// stochvol orders them); prior = {mu_mean, mu_sd, phi_a, phi_b, s2_rate, h0_var, mu_const}.
// para_io = {mu, phi, sigma, h0}; h_io[T]; draws_out[n_sweeps][4 + T] (mu, phi, sigma, h0, h...).
void hostrng_sv_chain(int T, const double* y, int n_sweeps, const double* prior, double* para_io, double* h_io,
                      unsigned long long seed, double* draws_out) {
  bvb::SvPrior pr;
  pr.mu_mean = prior[0]; pr.mu_sd = prior[1]; pr.phi_a = prior[2]; pr.phi_b = prior[3];
  pr.s2_shape = 0.5; pr.s2_rate = prior[4]; pr.h0_var = prior[5]; pr.mu_const = (int)prior[6];
  std::vector<double> ystar(T), eps(T + 1), scratch(4 * (T + 1));
  std::vector<unsigned char> r(T);
  for (int t = 0; t < T; ++t) ystar[t] = log(y[t] * y[t]);
  double mu = para_io[0], phi = para_io[1], sigma = para_io[2], h0 = para_io[3];
  for (int it = 0; it < n_sweeps; ++it) {
    for (int t = 0; t <= T; ++t) {
      bvb::RngStream e(seed, 1u, (uint32_t)t, (uint32_t)it);
      eps[t] = e.normal();
      if (t < T) r[t] = (unsigned char)bvb::sv_draw_indicator(ystar[t] - h_io[t], e.uniform());
    }
    bvb::RngStream g(seed, 2u, 0u, (uint32_t)it);
    bvb::sv_update_series(g, T, ystar.data(), h_io, h0, mu, phi, sigma, r.data(), eps.data(), pr, scratch.data());
    double* o = draws_out + (size_t)it * (4 + T);
    o[0] = mu; o[1] = phi; o[2] = sigma; o[3] = h0;
    for (int t = 0; t < T; ++t) o[4 + t] = h_io[t];
  }
  para_io[0] = mu; para_io[1] = phi; para_io[2] = sigma; para_io[3] = h0;
}
}
