// K6 - impulse-response recursion over stored posterior draws.
//
// Replaces irf_cpp / shift_and_insert (src/irf.cpp:457-524): per draw the impact
// response (Lambda diag(sigma) shocks for the factor model, U^-T diag(sqrt d) shocks
// for the cholesky model) followed by `ahead` products predictors x PHI with a lag
// shift.  Draws are independent, and so are the shocks of one draw (each shock is
// one row of the predictor matrix), so the grid is (draw, chunk of <= 8 shocks):
// the 8 x Kp predictor rows live in shared memory, one warp owns one output
// variable at a time and streams the matching PHI column (coalesced) once per
// horizon; consecutive CTAs work on the same draw, so its PHI slice is served
// from L2 after the first touch.  HBM-bound for factor models (few shocks).
// Quirks Q1/Q2 of SURVEY.md are reproduced on purpose.
#include "bvb_internal.cuh"

namespace bvb {

constexpr int IRF_SC_MAX = 8;   // shocks per CTA: 1, 2, 4 or 8 (template parameter: a factor model has r shocks)
constexpr int IRF_THREADS = 256;

struct IrfParams {
  const double* coef;     // Kp x M x R
  const double* facload;  // M x nf x R
  const double* Uvecs;    // nU x R
  const double* logvar;   // ld x R
  const double* shocks;   // dim x n_shocks
  const double* rotation; // dim x dim x R
  double* out;            // M x n_shocks x (ahead+1) x R
  int Kp, M, R, nf, hasU, ld, dim, n_shocks, ahead;
};

template <int IRF_SC>
__global__ void __launch_bounds__(IRF_THREADS) irf_kernel(const IrfParams p) {
  extern __shared__ double sm[];
  const int M = p.M, Kp = p.Kp;
  double* pred = sm;                       // [2][SC][Kp]
  double* rs = pred + 2 * IRF_SC * Kp;     // [dim][SC] rotated, scaled shocks
  double* ir0 = rs + p.dim * IRF_SC;       // [M][SC]
  const int r = blockIdx.x / ((p.n_shocks + IRF_SC - 1) / IRF_SC);
  const int s0 = (blockIdx.x % ((p.n_shocks + IRF_SC - 1) / IRF_SC)) * IRF_SC;
  const int ns = min(IRF_SC, p.n_shocks - s0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = IRF_THREADS / 32;
  const double* PHI = p.coef + (size_t)r * Kp * M;
  // rotated_shocks = rotation[r] * shocks | shocks, rows scaled by sqrt of the variances (irf.cpp:493-503)
  for (int i = tid; i < p.dim * IRF_SC; i += IRF_THREADS) {
    const int row = i / IRF_SC, s = i % IRF_SC;
    double v = 0.0;
    if (s < ns) {
      if (p.rotation) {
        const double* rot = p.rotation + (size_t)r * p.dim * p.dim;
        for (int k = 0; k < p.dim; ++k) v = fma(rot[(size_t)k * p.dim + row], p.shocks[(size_t)(s0 + s) * p.dim + k], v);
      } else {
        v = p.shocks[(size_t)(s0 + s) * p.dim + row];
      }
      if (p.nf > 0 || p.hasU) v *= exp(0.5 * p.logvar[(size_t)r * p.ld + row]);   // Q1: head(n_factors) of logvar_t
    }
    rs[i] = v;
  }
  __syncthreads();
  // impact responses
  if (p.nf > 0) {
    const double* L = p.facload + (size_t)r * M * p.nf;
    for (int i = tid; i < M * IRF_SC; i += IRF_THREADS) {
      const int m = i / IRF_SC, s = i % IRF_SC;
      double v = 0.0;
      for (int k = 0; k < p.nf; ++k) v = fma(L[(size_t)k * M + m], rs[k * IRF_SC + s], v);
      ir0[i] = v;
    }
  } else if (p.hasU) {
    // solve(trimatl(U'), rs): forward substitution, one thread per shock (unit diagonal)
    const double* uv = p.Uvecs + (size_t)r * ((size_t)M * (M - 1) / 2);
    if (tid < IRF_SC) {
      const int s = tid;
      for (int i = 0; i < M; ++i) {
        double v = rs[i * IRF_SC + s];
        // (U')[i, k] = U[k, i], k < i, stored at i(i-1)/2 + k
        for (int k = 0; k < i; ++k) v = fma(-uv[(size_t)i * (i - 1) / 2 + k], ir0[k * IRF_SC + s], v);
        ir0[i * IRF_SC + s] = v;
      }
    }
  } else {
    for (int i = tid; i < M * IRF_SC; i += IRF_THREADS) ir0[i] = rs[i];
  }
  __syncthreads();
  const size_t ostride_t = (size_t)M * p.n_shocks;
  double* out = p.out + (size_t)r * ostride_t * (p.ahead + 1);
  for (int i = tid; i < M * ns; i += IRF_THREADS) {
    const int m = i % M, s = i / M;
    out[(size_t)(s0 + s) * M + m] = ir0[m * IRF_SC + s];
  }
  // current_predictors = [ir0', 0 ...]
  for (int i = tid; i < IRF_SC * Kp; i += IRF_THREADS) {
    const int s = i / Kp, k = i % Kp;
    pred[i] = (k < M) ? ir0[k * IRF_SC + s] : 0.0;
  }
  __syncthreads();
  int cur = 0;
  for (int t = 1; t <= p.ahead; ++t) {
    const double* pc = pred + cur * IRF_SC * Kp;
    double* pn = pred + (cur ^ 1) * IRF_SC * Kp;
    // new[s][m] = sum_k pred[s][k] PHI[k, m]: one warp per variable m, lanes over k
    for (int m = warp; m < M; m += nw) {
      const double* col = PHI + (size_t)m * Kp;
      double acc[IRF_SC];
#pragma unroll
      for (int s = 0; s < IRF_SC; ++s) acc[s] = 0.0;
      // four independent loads in flight per lane: the PHI stream (L2 after the first horizon) is latency-bound
      int k = lane;
      for (; k + 96 < Kp; k += 128) {
        const double c0 = col[k], c1 = col[k + 32], c2 = col[k + 64], c3 = col[k + 96];
#pragma unroll
        for (int s = 0; s < IRF_SC; ++s) {
          const double* ps = pc + s * Kp + k;
          acc[s] = fma(ps[0], c0, acc[s]);
          acc[s] = fma(ps[32], c1, acc[s]);
          acc[s] = fma(ps[64], c2, acc[s]);
          acc[s] = fma(ps[96], c3, acc[s]);
        }
      }
      for (; k < Kp; k += 32) {
        const double c = col[k];
#pragma unroll
        for (int s = 0; s < IRF_SC; ++s) acc[s] = fma(pc[s * Kp + k], c, acc[s]);
      }
#pragma unroll
      for (int s = 0; s < IRF_SC; ++s) acc[s] = warp_sum(acc[s]);
      if (lane == 0) {
#pragma unroll
        for (int s = 0; s < IRF_SC; ++s) {
          pn[s * Kp + m] = acc[s];
          if (s < ns) out[(size_t)t * ostride_t + (size_t)(s0 + s) * M + m] = acc[s];
        }
      }
    }
    // shift_and_insert (irf.cpp:457-465): columns M..Kp-2 take the values M to their left (Q2)
    for (int i = tid; i < IRF_SC * Kp; i += IRF_THREADS) {
      const int s = i / Kp, k = i % Kp;
      if (k >= M) pn[i] = (k <= Kp - 2) ? pc[s * Kp + k - M] : pc[i];
    }
    __syncthreads();
    cur ^= 1;
  }
}

template <int SC>
static cudaError_t launch_irf_sc(const IrfParams& p, cudaStream_t s) {
  const size_t smem = ((size_t)2 * SC * p.Kp + (size_t)p.dim * SC + (size_t)p.M * SC) * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(irf_kernel<SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int chunks = (p.n_shocks + SC - 1) / SC;
  irf_kernel<SC><<<p.R * chunks, IRF_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_irf(const IrfParams& p, cudaStream_t s) {
  if (p.n_shocks <= 1) return launch_irf_sc<1>(p, s);
  if (p.n_shocks <= 2) return launch_irf_sc<2>(p, s);
  if (p.n_shocks <= 4) return launch_irf_sc<4>(p, s);
  return launch_irf_sc<IRF_SC_MAX>(p, s);
}

}  // namespace bvb

using namespace bvb;

extern "C" int bvb_irf(bvb_handle* h, int Kp, int M, int R, int n_factors, int has_U, int ld_logvar, int dim, int n_shocks,
                       int ahead, const double* coefficients, const double* factor_loadings, const double* U_vecs,
                       const double* logvar_t, const double* shocks, const double* rotation, double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 4; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (Kp <= 0 || M <= 0 || R < 0 || n_factors < 0 || dim <= 0 || n_shocks <= 0 || ahead < 0 || !coefficients || !shocks || !out ||
      (n_factors > 0 && !factor_loadings) || (has_U && !U_vecs) || ((n_factors > 0 || has_U) && !logvar_t) || Kp < M ||
      (n_factors > 0 ? dim != n_factors : dim != M)) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_irf: invalid argument");
    return BVB_ERR_ARG;
  }
  if (R == 0) return BVB_OK;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t d = sizeof(double);
  const size_t nU = (size_t)M * (M - 1) / 2;
  double *dc = nullptr, *dl = nullptr, *du = nullptr, *dv = nullptr, *ds = nullptr, *dr = nullptr, *dout = nullptr;
  auto fin = [&](int rc) {
    for (double* q : {dc, dl, du, dv, ds, dr, dout}) if (q) cudaFree(q);
    return rc;
  };
  const size_t out_n = (size_t)M * n_shocks * (ahead + 1) * R;
#define IRF_TRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { bvb_set_error(h, BVB_ERR_CUDA, "bvb_irf: %s", cudaGetErrorString(_e)); return fin(BVB_ERR_CUDA); } } while (0)
  IRF_TRY(cudaMalloc(&dc, d * Kp * M * R));
  IRF_TRY(cudaMalloc(&ds, d * dim * n_shocks));
  IRF_TRY(cudaMalloc(&dout, d * out_n));
  IRF_TRY(cudaMemcpyAsync(dc, coefficients, d * Kp * M * R, cudaMemcpyHostToDevice, s));
  IRF_TRY(cudaMemcpyAsync(ds, shocks, d * dim * n_shocks, cudaMemcpyHostToDevice, s));
  if (n_factors > 0) {
    IRF_TRY(cudaMalloc(&dl, d * M * n_factors * R));
    IRF_TRY(cudaMemcpyAsync(dl, factor_loadings, d * M * n_factors * R, cudaMemcpyHostToDevice, s));
  }
  if (has_U && nU > 0) {
    IRF_TRY(cudaMalloc(&du, d * nU * R));
    IRF_TRY(cudaMemcpyAsync(du, U_vecs, d * nU * R, cudaMemcpyHostToDevice, s));
  }
  if (logvar_t && ld_logvar > 0) {
    IRF_TRY(cudaMalloc(&dv, d * ld_logvar * R));
    IRF_TRY(cudaMemcpyAsync(dv, logvar_t, d * ld_logvar * R, cudaMemcpyHostToDevice, s));
  }
  if (rotation) {
    IRF_TRY(cudaMalloc(&dr, d * dim * dim * R));
    IRF_TRY(cudaMemcpyAsync(dr, rotation, d * dim * dim * R, cudaMemcpyHostToDevice, s));
  }
  IrfParams p{};
  p.coef = dc; p.facload = dl; p.Uvecs = du; p.logvar = dv; p.shocks = ds; p.rotation = dr; p.out = dout;
  p.Kp = Kp; p.M = M; p.R = R; p.nf = n_factors; p.hasU = (has_U && n_factors == 0) ? 1 : 0; p.ld = ld_logvar; p.dim = dim;
  p.n_shocks = n_shocks; p.ahead = ahead;
  IRF_TRY(cudaEventRecord(h->ev[0], s));
  IRF_TRY(launch_irf(p, s));
  IRF_TRY(cudaEventRecord(h->ev[1], s));
  IRF_TRY(cudaMemcpyAsync(out, dout, d * out_n, cudaMemcpyDeviceToHost, s));
  IRF_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = 1;
#undef IRF_TRY
  return fin(BVB_OK);
}
