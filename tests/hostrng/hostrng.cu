// Host-side harness around the __host__ __device__ RNG code (Philox stream,
// normal/gamma/GIG samplers) so its LOGIC can be tested distributionally without
// a GPU.  Test tool only: the product draws these variates on the device.
#include "../../bayesianvars_b200/csrc/gig.cuh"

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
}
