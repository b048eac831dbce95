// K3 - hyper-parameter draws of the hierarchical shrinkage priors (Philox
// counter-based RNG: stream = (purpose, coefficient | group, sweep, attempt)).
//
// Replaces sample_V_i_DL / _GT / _HS / _SSVS_beta / _HMP / _U_HMP
// (src/sample_coefficients.cpp:190-377): the per-coefficient loops there are
// sequential scalar calls through an R function pointer; here every coefficient
// is one thread, group statistics are fixed-order (deterministic) reductions and
// the per-group scalars (xi, zeta, varpi, p, a) are drawn by one CTA per group (any
// number of groups).  Four small kernels: local1 -> global1 -> local2 -> global2.
#include "hyper.cuh"

namespace bvb {

__device__ __forceinline__ int hyper_active(const HyperDev& hd, uint32_t rep, int burnin) {
  // bvar_cpp.cpp:585 (SSVS: rep > 0.1*burnin || SSVS_hyper), :602 (HMP: rep > 0.1*burnin)
  if (hd.always_active) return 1;
  const bool late = (double)rep > 0.1 * (double)burnin;
  if (hd.prior == BVB_PRIOR_SSVS) return (late || hd.hyper) ? 1 : 0;
  if (hd.prior == BVB_PRIOR_HMP) return late ? 1 : 0;
  return 1;
}

// Sum of v over the members of group g.  gridDim.y = hd.nsplit CTAs share a group: every CTA reduces its contiguous
// slice of the member list (threads stride over it, fixed-order block reduction) and, when the group is split, writes a
// partial; the LAST CTA of the group to arrive adds the partials in slice order and goes on alone (returns true), the
// others leave.  The result does not depend on which CTA was last.  (One CTA per group walked the 40 000 members of a
// single global group in 156 dependent strides: 19 us on the side branch that feeds the next sweep's K1.)
__device__ __forceinline__ bool group_sum(const HyperDev& hd, const double* v, int g, double* sh, double& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = (int)gridDim.y, sl = (int)blockIdx.y;
  const int beg0 = hd.gstart[g], end0 = hd.gstart[g + 1];
  const int per = (end0 - beg0 + ns - 1) / ns;
  const int beg = beg0 + sl * per, end = min(beg + per, end0);
  double x = 0.0;
  for (int k = beg + (int)threadIdx.x; k < end; k += HYP_THREADS) x += v[hd.gmem[k]];
  x = warp_sum(x);
  __syncthreads();
  if (lane == 0) sh[warp] = x;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < HYP_THREADS / 32; ++w) t += sh[w];
  if (ns == 1) { total = t; return true; }
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    hd.gpart[(size_t)g * ns + sl] = t;
    __threadfence();
    const int old = atomicAdd(hd.gcnt + g, 1);
    __threadfence();
    s_last = (old == ns - 1) ? 1 : 0;
    if (s_last) hd.gcnt[g] = 0;
  }
  __syncthreads();
  if (!s_last) return false;
  double tt = 0.0;
  for (int q = 0; q < ns; ++q) tt += __ldcg(hd.gpart + (size_t)g * ns + q);
  total = tt;
  return true;
}

__global__ void __launch_bounds__(HYP_THREADS) hyper_local1(const HyperDev hd, uint64_t seed, const uint32_t* sweep_p, int burnin) {
  const uint32_t sweep = *sweep_p;
  const int active = hyper_active(hd, sweep, burnin);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hd.n) return;
  const int g = hd.grp[i];
  double red = 0.0;
  if (g >= 0) {
    const double* gp = hd.gpar + g * HYP_GP;
    const double phi = hd.coef[i], c2 = phi * phi;
    RngStream rng(seed, hd.stream_base, (uint32_t)i, sweep);
    switch (hd.prior) {
      case BVB_PRIOR_DL: {  // sample_coefficients.cpp:276-282: lambda first, then psi
        const double lam = gig_draw(rng, gp[GP_A] - 1.0, 2.0 * fabs(phi), 2.0 * gp[GP_XI]);
        const double psi = gig_draw(rng, 0.5, c2 / (lam * lam), 1.0);
        hd.loc1[i] = lam;
        hd.loc2[i] = psi;
        red = lam;
        if (hd.status && (isnan(lam) || isnan(psi))) atomicExch(hd.status, 1);
      } break;
      case BVB_PRIOR_GT: {  // :202-210
        double psi = hd.loc2[i];
        if (hd.kernel_exp) {
          psi = 1.0 / gig_draw(rng, -0.5, 1.0, c2 / (hd.loc1[i] * hd.vs));
          hd.loc2[i] = psi;
        }
        const double lam = gig_draw(rng, gp[GP_A] - 0.5, c2 / (psi * hd.vs), 2.0 * gp[GP_XI]);
        hd.loc1[i] = lam;
        red = lam;
        if (hd.status && (isnan(lam) || isnan(psi))) atomicExch(hd.status, 1);
      } break;
      case BVB_PRIOR_HS: {  // :252-256  (R::rgamma(1, scale) = scale * Exp(1))
        const double theta = (1.0 / hd.loc2[i] + c2 / (2.0 * gp[GP_ZETA])) / rng.exponential();
        const double nu = (1.0 + 1.0 / theta) / rng.exponential();
        hd.loc1[i] = theta;
        hd.loc2[i] = nu;
        red = c2 / theta;
      } break;
      case BVB_PRIOR_SSVS: {  // :326-345
        if (active) {
          const double t0 = hd.tau0[i], t1 = hd.tau1[i], p = hd.loc2[i];
          const double u1 = -log(t1) - 0.5 * (phi / t1) * (phi / t1) + log(p);
          const double u2 = -log(t0) - 0.5 * (phi / t0) * (phi / t0) + log(1.0 - p);
          const double gst = 1.0 / (1.0 + exp(u2 - u1));
          const double gam = (rng.uniform() < gst) ? 1.0 : 0.0;
          hd.loc1[i] = gam;
          hd.V_i[i] = (gam == 1.0) ? t1 * t1 : t0 * t0;
          red = gam;
        }
      } break;
      case BVB_PRIOR_HMP: {  // :361-363
        red = c2 / hd.V_prep[i];
      } break;
      default: break;
    }
  }
  hd.red[i] = red;
}

__global__ void __launch_bounds__(HYP_THREADS) hyper_global1(const HyperDev hd, uint64_t seed, const uint32_t* sweep_p, int burnin) {
  __shared__ double sh[HYP_THREADS / 32];
  const uint32_t sweep = *sweep_p;
  const int active = hyper_active(hd, sweep, burnin);
  const int g = blockIdx.x;
  double s0;
  if (!group_sum(hd, hd.red, g, sh, s0)) return;
  if (threadIdx.x != 0) return;
  double* gp = hd.gpar + g * HYP_GP;
  hd.gsum[g * HYP_NRED + 0] = s0;
  RngStream rng(seed, hd.stream_base + 1, (uint32_t)g, sweep);
  const double n = gp[GP_N];
  switch (hd.prior) {
    case BVB_PRIOR_DL:
      gp[GP_ZETA] = s0;
      if (hd.dl_plus) gp[GP_XI] = rng.gamma(gp[GP_B] + n * gp[GP_A]) / (2.0 * gp[GP_C] / gp[GP_A] + s0);  // :284-286
      break;
    case BVB_PRIOR_GT:
      gp[GP_ZETA] = s0;
      gp[GP_XI] = rng.gamma(gp[GP_B] + n * gp[GP_A]) / (2.0 * gp[GP_C] / gp[GP_A] + s0);  // :219
      break;
    case BVB_PRIOR_HS: {  // :257-258
      const double zeta = (1.0 / gp[GP_VARPI] + 0.5 * s0) / rng.gamma(0.5 * (n + 1.0));
      gp[GP_ZETA] = zeta;
      gp[GP_VARPI] = (1.0 + 1.0 / zeta) / rng.exponential();
    } break;
    case BVB_PRIOR_SSVS:
      if (active && hd.hyper) {  // :347-351
        const double ga = rng.gamma(hd.s_a + s0), gb = rng.gamma(hd.s_b + n - s0);
        gp[GP_AUX] = ga / (ga + gb);
      }
      break;
    case BVB_PRIOR_HMP:
      if (active) {  // :361-363 / :374  (n/2 is INTEGER division in the reference)
        const double half = (double)(((long long)n) / 2);
        gp[GP_AUX] = gig_draw(rng, gp[GP_A] - half, s0, 2.0 * gp[GP_B]);
        if (hd.status && isnan(gp[GP_AUX])) atomicExch(hd.status, 1);
      }
      break;
    default: break;
  }
}

__global__ void __launch_bounds__(HYP_THREADS) hyper_local2(const HyperDev hd, const uint32_t* sweep_p, int burnin) {
  const int active = hyper_active(hd, *sweep_p, burnin);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hd.n) return;
  const int g = hd.grp[i];
  double red = 0.0;
  if (g >= 0) {
    const double* gp = hd.gpar + g * HYP_GP;
    switch (hd.prior) {
      case BVB_PRIOR_DL: {  // :288-295
        double lam = hd.loc1[i];
        if (hd.tol > 0.0 && lam / gp[GP_ZETA] < hd.tol) { lam = hd.tol * gp[GP_ZETA]; hd.loc1[i] = lam; }
        hd.V_i[i] = hd.loc2[i] * lam * lam;
        red = log(lam);
      } break;
      case BVB_PRIOR_GT: {  // :211-225
        double lam = hd.loc1[i];
        if (hd.tol > 0.0 && lam / gp[GP_ZETA] < hd.tol) { lam = hd.tol * gp[GP_ZETA]; hd.loc1[i] = lam; }
        hd.V_i[i] = hd.kernel_exp ? hd.loc2[i] * lam * hd.vs : lam * hd.vs;
        red = log(lam);
      } break;
      case BVB_PRIOR_HS:
        hd.V_i[i] = hd.loc1[i] * gp[GP_ZETA];  // :259
        break;
      case BVB_PRIOR_SSVS:
        if (active && hd.hyper) hd.loc2[i] = gp[GP_AUX];  // :351
        break;
      case BVB_PRIOR_HMP:
        hd.V_i[i] = gp[GP_AUX] * hd.V_prep[i];  // :365-366
        break;
      default: break;
    }
  }
  hd.red[(size_t)hd.n + i] = red;
}

// Discrete update of `a` (and c when c_rel_a) over the grid: :297-314 / :227-244.
__global__ void __launch_bounds__(HYP_THREADS) hyper_global2(const HyperDev hd, uint64_t seed, const uint32_t* sweep_p) {
  __shared__ double sh[HYP_THREADS / 32];
  const uint32_t sweep = *sweep_p;
  const int g = blockIdx.x, lane = threadIdx.x & 31;
  double slog;
  if (!group_sum(hd, hd.red + hd.n, g, sh, slog)) return;
  if (threadIdx.x >= 32) return;
  double* gp = hd.gpar + g * HYP_GP;
  const double n = gp[GP_N], xi = gp[GP_XI], b = gp[GP_B], c = gp[GP_C];
  const double lxi = log(xi);
  // pass 1: max of the log-probabilities (lanes stride over the grid)
  double best = -INFINITY;
  for (int k = lane; k < hd.grid_len; k += 32) {
    const double av = hd.a_vec[k];
    double lp = log(hd.a_weight[k]) + n * (av * lxi - hd.norm_consts[k]) + (av - 1.0) * slog;
    if (!hd.c_rel_a) lp += -2.0 * xi * c / av - b * log(av);
    best = fmax(best, lp);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane != 0) return;
  hd.gsum[g * HYP_NRED + 1] = slog;
  // pass 2 (serial, fixed order): normalise and invert the cdf with one uniform
  double tot = 0.0;
  for (int k = 0; k < hd.grid_len; ++k) {
    const double av = hd.a_vec[k];
    double lp = log(hd.a_weight[k]) + n * (av * lxi - hd.norm_consts[k]) + (av - 1.0) * slog;
    if (!hd.c_rel_a) lp += -2.0 * xi * c / av - b * log(av);
    tot += exp(lp - best);
  }
  RngStream rng(seed, hd.stream_base + 2, (uint32_t)g, sweep);
  const double target = rng.uniform() * tot;
  double cum = 0.0;
  int pick = hd.grid_len - 1;
  for (int k = 0; k < hd.grid_len; ++k) {
    const double av = hd.a_vec[k];
    double lp = log(hd.a_weight[k]) + n * (av * lxi - hd.norm_consts[k]) + (av - 1.0) * slog;
    if (!hd.c_rel_a) lp += -2.0 * xi * c / av - b * log(av);
    cum += exp(lp - best);
    if (target < cum) { pick = k; break; }
  }
  gp[GP_A] = hd.a_vec[pick];
  if (hd.c_rel_a) gp[GP_C] = hd.c_vec[pick];
}

cudaError_t launch_hyper_update(const HyperDev& hd, uint64_t seed, const uint32_t* sweep, int burnin, cudaStream_t s) {
  if (hd.prior == BVB_PRIOR_NORMAL || hd.prior == BVB_PRIOR_NOTHING || hd.n == 0) return cudaSuccess;
  hyper_local1<<<hd.nblocks, HYP_THREADS, 0, s>>>(hd, seed, sweep, burnin);
  const dim3 gg((unsigned)hd.G, (unsigned)(hd.nsplit > 0 ? hd.nsplit : 1));
  hyper_global1<<<gg, HYP_THREADS, 0, s>>>(hd, seed, sweep, burnin);
  hyper_local2<<<hd.nblocks, HYP_THREADS, 0, s>>>(hd, sweep, burnin);
  if ((hd.prior == BVB_PRIOR_DL || hd.prior == BVB_PRIOR_GT) && hd.hyper && hd.grid_len > 0)
    hyper_global2<<<gg, HYP_THREADS, 0, s>>>(hd, seed, sweep);
  return cudaGetLastError();
}

}  // namespace bvb
