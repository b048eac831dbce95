// K10 - in-sample fit: insample / predict_y (src/utilities_cpp.cpp:396-433, :69-93).
//
//   y_pred[t, :, r] = x_t' PHI_r  (+ z_tr' chol_upper(Sigma_tr)  when `prediction`)
//
// The mean is a batched GEMM (T x Kp)(Kp x M) per stored draw - FP64 DMMA, operands from L2 (X is shared by every draw,
// PHI_r is read once per draw).  The error term is where the reference spends its time: one M x M Cholesky per
// (t, draw) (build_sigma, :45-67).  Neither structure needs it:
//   cholesky structure   chol_upper(Sigma) = diag(e^{h/2}) U^-1, so  z'R = w  with  w U = (z . e^{h/2})'  - one forward
//                        substitution on the unit upper triangle (M^2/2 flops instead of M^3/3 + a matrix inverse);
//   factor structure     Sigma = D + G G' (G = Lambda diag(e^{h_f/2}), M x r).  Eliminating column j of D + G C G' leaves
//                        D + G (C - C g_j g_j' C / d_j) G' on the trailing block, d_j = D_j + g_j' C g_j: the Cholesky factor is
//                        L_jj = sqrt(d_j), L_ij = g_i' v_j (i > j) with v_j = C_j g_j / sqrt(d_j), and
//                        (z'R)_j = z_j sqrt(d_j) + g_j' sum_{i<j} z_i v_i  - O(M r^2) per (t, draw), the same numbers
//                        as the dense factorisation (parity 1e-9 against it with injected normals).
// One thread per (t, draw) for the error term (T x nsave of them), C / s in local memory (r <= 16).
#include <vector>
#include "bvb_internal.cuh"

namespace bvb {

constexpr int INS_RMAX = 16;
constexpr int INS_SPLIT = 4;

// out[t, j, r] = sum_k X[t, k] PHI[k, j, r]: one CTA per 8 x 8 output tile of one draw, four warps split the contraction
// (the shape of gibbs_resid_kernel; both operands L2-resident).
__global__ void __launch_bounds__(32 * INS_SPLIT) insample_mean_kernel(double* __restrict__ out, const double* __restrict__ X,
                                                                      const double* __restrict__ PHI, int T, int Kp, int M) {
  __shared__ double part[INS_SPLIT][64];
  const int lane = threadIdx.x & 31, wk = threadIdx.x >> 5;
  const int ntj = (M + 7) / 8;
  const int tile = blockIdx.x, r = blockIdx.y;
  const int t0 = (tile / ntj) * 8, j0 = (tile % ntj) * 8;
  const int fr = lane >> 2, fk = lane & 3;
  const int tr = min(t0 + fr, T - 1), jc = min(j0 + fr, M - 1);
  const double* xa = X + tr;
  const double* pb = PHI + ((size_t)r * M + jc) * Kp;
  const int steps = (Kp + 3) / 4, per = (steps + INS_SPLIT - 1) / INS_SPLIT;
  const int kbeg = min(Kp, 4 * per * wk), kend = min(Kp, 4 * per * (wk + 1));
  double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
  for (int k = kbeg; k < kend; k += 32) {
    double av[8], bv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = k + 4 * q + fk;
      av[q] = kk < kend ? xa[(size_t)kk * T] : 0.0;
      bv[q] = kk < kend ? pb[kk] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      dmma884(c0, c1, av[q], bv[q]);
      dmma884(d0, d1, av[q + 1], bv[q + 1]);
    }
  }
  c0 += d0; c1 += d1;
  part[wk][fr * 8 + 2 * fk] = c0;
  part[wk][fr * 8 + 2 * fk + 1] = c1;
  __syncthreads();
  if (wk == 0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = fr * 8 + 2 * fk + e;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < INS_SPLIT; ++w) v += part[w][idx];
      const int t = t0 + fr, j = j0 + 2 * fk + e;
      if (t < T && j < M) out[((size_t)r * M + j) * T + t] = v;
    }
  }
}

// error term, one thread per (t, draw); z: injected normals C-ordered [T][R][M] (the reference's loop order) or null
__global__ void __launch_bounds__(128) insample_noise_kernel(double* __restrict__ out, const double* __restrict__ U,
                                                             const double* __restrict__ facload, const double* __restrict__ logvar,
                                                             const double* __restrict__ z, int factor, int T, int M, int nf, int R,
                                                             uint64_t seed) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)T * R) return;
  const int t = (int)(idx % T), r = (int)(idx / T);
  const int S = M + nf;
  const double* lv = logvar + ((size_t)r * S) * T + t;                 // lv[s * T]
  double* y = out + ((size_t)r * M) * T + t;                           // y[j * T]
  const double* zz = z ? z + ((size_t)t * R + r) * M : nullptr;
  RngStream g(seed, 0x910u, (uint32_t)idx, (uint32_t)(idx >> 32));
  if (!factor) {
    // w U = (z . e^{h/2})' : w_j = v_j - sum_{i<j} w_i U_ij, U_ij at j(j-1)/2 + i of draw r's vector; w is kept in `y`'s place
    const double* u = U + (size_t)r * ((size_t)M * (M - 1) / 2);
    double mean_j;
    for (int j = 0; j < M; ++j) {
      const double zj = zz ? zz[j] : g.normal();
      double w = zj * exp(0.5 * lv[(size_t)j * T]);
      const double* uc = u + (size_t)j * (j - 1) / 2;
      // the already finished w_i sit in y[i] - mean_i: keep them separately to leave the mean untouched
      for (int i = 0; i < j; ++i) w -= out[((size_t)R * M) * T + ((size_t)r * M + i) * T + t] * uc[i];
      out[((size_t)R * M) * T + ((size_t)r * M + j) * T + t] = w;       // scratch plane behind the result
      mean_j = y[(size_t)j * T];
      y[(size_t)j * T] = mean_j + w;
    }
    return;
  }
  double C[INS_RMAX * INS_RMAX], s[INS_RMAX], gj[INS_RMAX], uj[INS_RMAX], ef[INS_RMAX];
  for (int a = 0; a < nf; ++a) {
    ef[a] = exp(0.5 * lv[(size_t)(M + a) * T]);
    s[a] = 0.0;
    for (int b = 0; b < nf; ++b) C[a * INS_RMAX + b] = (a == b) ? 1.0 : 0.0;
  }
  const double* fl = facload + (size_t)r * M * nf;
  for (int j = 0; j < M; ++j) {
    const double zj = zz ? zz[j] : g.normal();
    double d = exp(lv[(size_t)j * T]), gs = 0.0;
    for (int a = 0; a < nf; ++a) gj[a] = fl[(size_t)a * M + j] * ef[a];
    for (int a = 0; a < nf; ++a) {
      double v = 0.0;
      for (int b = 0; b < nf; ++b) v += C[a * INS_RMAX + b] * gj[b];
      uj[a] = v;
      d += gj[a] * v;
      gs += gj[a] * s[a];
    }
    const double sd = sqrt(d);
    y[(size_t)j * T] += zj * sd + gs;
    const double zs = zj / sd, di = 1.0 / d;
    for (int a = 0; a < nf; ++a) {
      s[a] += zs * uj[a];
      for (int b = 0; b < nf; ++b) C[a * INS_RMAX + b] -= uj[a] * uj[b] * di;
    }
  }
}

}  // namespace bvb

using namespace bvb;

extern "C" int bvb_insample(bvb_handle* h, int T, int Kp, int M, int R, int n_factors, int factor, int prediction, const double* X,
                            const double* PHI, const double* U, const double* facload, const double* logvar, const double* z,
                            double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 6; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  const int nf = factor ? n_factors : 0;
  if (T <= 0 || Kp <= 0 || M <= 0 || R < 0 || nf < 0 || !X || !PHI || !out ||
      (prediction && (!logvar || (factor && nf > 0 && !facload) || (!factor && M > 1 && !U)))) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_insample: invalid argument");
    return BVB_ERR_ARG;
  }
  if (prediction && factor && nf > INS_RMAX) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "bvb_insample: error term with more than %d factors", INS_RMAX);
    return BVB_ERR_UNSUPPORTED;
  }
  if (R == 0) return BVB_OK;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t d = sizeof(double), nU = (size_t)M * (M - 1) / 2, S = (size_t)M + nf;
  std::vector<void*> allocs;
  auto fin = [&](int rc) { for (void* q : allocs) cudaFree(q); return rc; };
#define K10_TRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { bvb_set_error(h, BVB_ERR_CUDA, "bvb_insample: %s", cudaGetErrorString(_e)); return fin(BVB_ERR_CUDA); } } while (0)
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    *dst = nullptr;
    if (!src || bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    allocs.push_back(*dst);
    return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, s);
  };
  double *dX, *dP, *dU = nullptr, *dfl = nullptr, *dl = nullptr, *dz = nullptr, *dout = nullptr;
  K10_TRY(up(X, d * (size_t)T * Kp, (void**)&dX));
  K10_TRY(up(PHI, d * (size_t)Kp * M * R, (void**)&dP));
  if (prediction) {
    K10_TRY(up(logvar, d * (size_t)T * S * R, (void**)&dl));
    if (factor && nf > 0) K10_TRY(up(facload, d * (size_t)M * nf * R, (void**)&dfl));
    if (!factor && M > 1) K10_TRY(up(U, d * nU * R, (void**)&dU));
    K10_TRY(up(z, d * (size_t)T * R * M, (void**)&dz));
  }
  const size_t nout = (size_t)T * M * R;
  // (cholesky structure: a scratch plane of the same size behind the result holds the substitution vector)
  K10_TRY(cudaMalloc(&dout, d * nout * ((prediction && !factor) ? 2 : 1))); allocs.push_back(dout);
  K10_TRY(cudaEventRecord(h->ev[0], s));
  const int tiles = ((T + 7) / 8) * ((M + 7) / 8);
  for (int r0 = 0; r0 < R; r0 += 65535) {
    const int rc = R - r0 < 65535 ? R - r0 : 65535;
    insample_mean_kernel<<<dim3((unsigned)tiles, (unsigned)rc), 32 * INS_SPLIT, 0, s>>>(dout + (size_t)r0 * T * M, dX,
                                                                                          dP + (size_t)r0 * Kp * M, T, Kp, M);
  }
  K10_TRY(cudaGetLastError());
  if (prediction) {
    const size_t n = (size_t)T * R;
    insample_noise_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(dout, dU, dfl, dl, dz, factor ? 1 : 0, T, M, nf, R, h->seed);
    K10_TRY(cudaGetLastError());
  }
  K10_TRY(cudaEventRecord(h->ev[1], s));
  K10_TRY(cudaMemcpyAsync(out, dout, d * nout, cudaMemcpyDeviceToHost, s));
  K10_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = prediction ? 2 : 1;
#undef K10_TRY
  return fin(BVB_OK);
}
