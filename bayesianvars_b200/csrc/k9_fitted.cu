// K9 - per-draw post-processing next to K7: log-variance forecasts and covariance draws.
//
// Replaces predh (src/utilities_cpp.cpp:181-212) and vcov_cpp (:436-472, build_sigma :45-67).  Both are element-wise
// over (draw, series[, t]) and bound by their OUTPUT stream: vcov writes T*M*M doubles per draw (40 MB at config 3), so
// the kernels are laid out for coalesced stores (t fastest, as the reference's cube (T, M, M) is).  insample (:396-433)
// is not built yet (oracle/fitted.py holds its restatement).
#include <vector>
#include "bvb_internal.cuh"

namespace bvb {

// one thread per (series m, draw r): the AR(1) forecast moments are a recursion over the horizon
__global__ void __launch_bounds__(256) predh_kernel(double* __restrict__ out, const double* __restrict__ lvT,
                                                    const double* __restrict__ mu, const double* __restrict__ phi,
                                                    const double* __restrict__ sig, const int* __restrict__ ahead, int na,
                                                    int M, int R, int each, const double* __restrict__ z, uint64_t seed) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * R) return;
  const int m = (int)(idx % M), r = (int)(idx / M);
  const double h = lvT[idx], mu_ = mu[idx], ph = phi[idx], sg = sig[idx];
  double sigma2 = 0.0;
  int counter = 0;
  const int max_ahead = ahead[na - 1];
  for (int f = 0; f < max_ahead && counter < na; ++f) {
    const double mean = mu_ + pow(ph, (double)(f + 1)) * (h - mu_);       // :196
    const double t = sg * pow(ph, (double)f);
    sigma2 += t * t;                                                        // :197
    if (f == ahead[counter] - 1) {
      const double sd = sqrt(sigma2);
      for (int j = 0; j < each; ++j) {
        double zz;
        if (z) zz = z[(((size_t)counter * each + j) * R + r) * M + m];      // C-ordered [na][each][R][M] == (M, R) column-major
        else { RngStream g(seed, 0x900u, (uint32_t)idx, (uint32_t)(counter * each + j)); zz = g.normal(); }
        out[(size_t)counter + (size_t)na * (m + (size_t)M * ((size_t)j * R + r))] = mean + zz * sd;   // slice j*R + r (:205)
      }
      ++counter;
    }
  }
}

// U_inv = inv(trimatu(Um)) per draw, one thread per column (back substitution on the unit upper triangle)
__global__ void __launch_bounds__(128) vcov_uinv_kernel(double* __restrict__ uinv, const double* __restrict__ U, int M, int nU, int R) {
  const int r = blockIdx.x;
  const double* u = U + (size_t)r * nU;
  double* out = uinv + (size_t)r * M * M;
  for (int c = threadIdx.x; c < M; c += blockDim.x) {
    double* col = out + (size_t)c * M;
    for (int i = M - 1; i >= 0; --i) {
      double v = (i == c) ? 1.0 : 0.0;
      if (i < c) for (int k = i + 1; k <= c; ++k) v -= u[(size_t)k * (k - 1) / 2 + i] * col[k];
      col[i] = (i <= c) ? v : 0.0;
    }
  }
}

// grid = (ceil(T / 128), M * M, R): thread t of entry (a, b) of draw i; stores are contiguous in t
__global__ void __launch_bounds__(128) vcov_kernel(double* __restrict__ out, const double* __restrict__ facload,
                                                   const double* __restrict__ logvar, const double* __restrict__ uinv, int factor,
                                                   int T, int M, int nf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int a = blockIdx.y % M, b = blockIdx.y / M, i = blockIdx.z;
  const int S = M + (factor ? nf : 0);
  const double* lv = logvar + (size_t)i * T * S + t;                       // logvar(t, s, i) = lv[s * T]
  double v = 0.0;
  if (factor) {
    const double* L = facload + (size_t)i * M * nf;
    for (int k = 0; k < nf; ++k) {
      const double e = exp(lv[(size_t)(M + k) * T] / 2.0);                 // Sigma = (L e^{h/2})(L e^{h/2})' (:50-52)
      v = fma(L[(size_t)k * M + a] * e, L[(size_t)k * M + b] * e, v);
    }
    if (a == b) v += exp(lv[(size_t)a * T]);                               // :53
  } else {
    const double* ui = uinv + (size_t)i * M * M;
    const int mm = a < b ? a : b;                                          // U_inv is upper triangular
    for (int k = 0; k <= mm; ++k) {
      const double e = exp(lv[(size_t)k * T] / 2.0);                       // Sigma_chol = U_inv % e^{h/2}, Sigma = Sc' Sc (:62-65)
      v = fma(ui[(size_t)a * M + k] * e, ui[(size_t)b * M + k] * e, v);
    }
  }
  out[(size_t)t + (size_t)T * (a + (size_t)M * (b + (size_t)M * i))] = v;  // cube (T, M, M) flattened to (T*M, M) (:448-449)
}

// Factor structure, n_factors <= 8 (the usual case): thread = one (or two consecutive) time points of draw i, CTA
// column block = BCH columns b of Sigma_t; the thread walks every row a.  The exponentials are taken ONCE per thread
// (n_factors + BCH of them) instead of once per output element - the element-wise kernel above spends 40+ FP64
// instructions on exp() per 8-byte store and is compute-bound at 1.24 TB/s - and the output leaves in 16-byte
// streaming stores, contiguous in t across the warp.  Sigma(a, b) = sum_k L(a,k) L(b,k) e^{h_fk} (+ e^{h_a} on the diagonal).
template <int NF, bool VEC2>
__global__ void __launch_bounds__(128) vcov_factor_kernel(double* __restrict__ out, const double* __restrict__ facload,
                                                          const double* __restrict__ logvar, int T, int M, int bch) {
  extern __shared__ double Ls[];                                            // [NF][M] loadings of draw i
  const int i = blockIdx.z, b0 = blockIdx.y * bch;
  const int S = M + NF;
  for (int q = threadIdx.x; q < NF * M; q += blockDim.x) Ls[q] = facload[(size_t)i * M * NF + q];
  __syncthreads();
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * (VEC2 ? 2 : 1);
  if (t >= T) return;
  const double* lv = logvar + (size_t)i * T * S + t;
  double ef0[NF], ef1[NF];
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    ef0[k] = exp(lv[(size_t)(M + k) * T]);
    ef1[k] = VEC2 ? exp(lv[(size_t)(M + k) * T + 1]) : 0.0;
  }
  double* obase = out + (size_t)t + (size_t)T * M * M * i;
  for (int b = b0; b < min(b0 + bch, M); ++b) {
    const double d0 = exp(lv[(size_t)b * T]);
    const double d1 = VEC2 ? exp(lv[(size_t)b * T + 1]) : 0.0;
    double lb[NF];
#pragma unroll
    for (int k = 0; k < NF; ++k) lb[k] = Ls[k * M + b];
    double* ocol = obase + (size_t)T * M * b;
#pragma unroll 4
    for (int a = 0; a < M; ++a) {
      double v0 = (a == b) ? d0 : 0.0, v1 = (a == b) ? d1 : 0.0;
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const double c = Ls[k * M + a] * lb[k];
        v0 = fma(c, ef0[k], v0);
        v1 = fma(c, ef1[k], v1);
      }
      if (VEC2) __stcs(reinterpret_cast<double2*>(ocol + (size_t)T * a), make_double2(v0, v1));
      else __stcs(ocol + (size_t)T * a, v0);
    }
  }
}

template <int NF>
static cudaError_t launch_vcov_factor(double* out, const double* facload, const double* logvar, int T, int M, int R, cudaStream_t s) {
  const int bch = 4;
  const bool vec2 = (T % 2) == 0;
  const int tthreads = vec2 ? T / 2 : T;
  for (int r0 = 0; r0 < R; r0 += 65535) {
    const int rc = R - r0 < 65535 ? R - r0 : 65535;
    dim3 grid((unsigned)((tthreads + 127) / 128), (unsigned)((M + bch - 1) / bch), (unsigned)rc);
    const size_t sm = sizeof(double) * NF * M;
    const size_t S = (size_t)M + NF;
    if (vec2) vcov_factor_kernel<NF, true><<<grid, 128, sm, s>>>(out + (size_t)r0 * T * M * M, facload + (size_t)r0 * M * NF, logvar + (size_t)r0 * T * S, T, M, bch);
    else vcov_factor_kernel<NF, false><<<grid, 128, sm, s>>>(out + (size_t)r0 * T * M * M, facload + (size_t)r0 * M * NF, logvar + (size_t)r0 * T * S, T, M, bch);
  }
  return cudaGetLastError();
}

}  // namespace bvb

using namespace bvb;

extern "C" int bvb_predh(bvb_handle* h, int M, int R, int each, const int* ahead, int n_ahead, const double* logvar_T,
                         const double* sv_mu, const double* sv_phi, const double* sv_sigma, const double* z, double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 6; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (M <= 0 || R < 0 || each <= 0 || !ahead || n_ahead <= 0 || !logvar_T || !sv_mu || !sv_phi || !sv_sigma || !out) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_predh: invalid argument");
    return BVB_ERR_ARG;
  }
  for (int i = 0; i < n_ahead; ++i)
    if (ahead[i] <= 0 || (i > 0 && ahead[i] <= ahead[i - 1])) { bvb_set_error(h, BVB_ERR_ARG, "bvb_predh: ahead must be sorted and positive"); return BVB_ERR_ARG; }
  if (R == 0) return BVB_OK;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t d = sizeof(double), n = (size_t)M * R, nout = (size_t)n_ahead * M * R * each;
  std::vector<void*> allocs;
  auto fin = [&](int rc) { for (void* q : allocs) cudaFree(q); return rc; };
#define K9_TRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { bvb_set_error(h, BVB_ERR_CUDA, "bvb_predh/vcov: %s", cudaGetErrorString(_e)); return fin(BVB_ERR_CUDA); } } while (0)
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    *dst = nullptr;
    if (!src || bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    allocs.push_back(*dst);
    return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, s);
  };
  double *dl, *dm, *dp, *ds, *dz, *dout = nullptr;
  int* da;
  K9_TRY(up(logvar_T, d * n, (void**)&dl)); K9_TRY(up(sv_mu, d * n, (void**)&dm)); K9_TRY(up(sv_phi, d * n, (void**)&dp));
  K9_TRY(up(sv_sigma, d * n, (void**)&ds)); K9_TRY(up(z, d * n * n_ahead * each, (void**)&dz));
  K9_TRY(up(ahead, sizeof(int) * n_ahead, (void**)&da));
  K9_TRY(cudaMalloc(&dout, d * nout)); allocs.push_back(dout);
  K9_TRY(cudaEventRecord(h->ev[0], s));
  predh_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dout, dl, dm, dp, ds, da, n_ahead, M, R, each, dz, h->seed);
  K9_TRY(cudaGetLastError());
  K9_TRY(cudaEventRecord(h->ev[1], s));
  K9_TRY(cudaMemcpyAsync(out, dout, d * nout, cudaMemcpyDeviceToHost, s));
  K9_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = 1;
  return fin(BVB_OK);
}

extern "C" int bvb_vcov(bvb_handle* h, int factor, int T, int M, int n_factors, int R, const double* facload,
                        const double* logvar, const double* U, double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 6; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  const int nf = factor ? n_factors : 0;
  if (T <= 0 || M <= 0 || R < 0 || nf < 0 || !logvar || !out || (factor && nf > 0 && !facload) || (!factor && M > 1 && !U) ||
      (size_t)M * M > 65535u) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_vcov: invalid argument");
    return BVB_ERR_ARG;
  }
  if (R == 0) return BVB_OK;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t d = sizeof(double), nU = (size_t)M * (M - 1) / 2, S = (size_t)M + nf;
  std::vector<void*> allocs;
  auto fin = [&](int rc) { for (void* q : allocs) cudaFree(q); return rc; };
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    *dst = nullptr;
    if (!src || bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    allocs.push_back(*dst);
    return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, s);
  };
  double *dl, *dfl = nullptr, *du = nullptr, *dui = nullptr, *dout = nullptr;
  K9_TRY(up(logvar, d * (size_t)T * S * R, (void**)&dl));
  if (factor && nf > 0) K9_TRY(up(facload, d * (size_t)M * nf * R, (void**)&dfl));
  if (!factor) {
    K9_TRY(up(U, d * nU * R, (void**)&du));
    K9_TRY(cudaMalloc(&dui, d * (size_t)M * M * R)); allocs.push_back(dui);
  }
  const size_t nout = (size_t)T * M * M * R;
  K9_TRY(cudaMalloc(&dout, d * nout)); allocs.push_back(dout);
  K9_TRY(cudaEventRecord(h->ev[0], s));
  if (!factor) vcov_uinv_kernel<<<R, 128, 0, s>>>(dui, du, M, (int)nU, R);
  bool done_fast = false;
  if (factor && nf >= 1 && nf <= 8 && (size_t)nf * M * sizeof(double) <= 48 * 1024 && !getenv("BVB_VCOV_ELEMENTWISE")) {
    cudaError_t fe = cudaSuccess;
    switch (nf) {
      case 1: fe = launch_vcov_factor<1>(dout, dfl, dl, T, M, R, s); break;
      case 2: fe = launch_vcov_factor<2>(dout, dfl, dl, T, M, R, s); break;
      case 3: fe = launch_vcov_factor<3>(dout, dfl, dl, T, M, R, s); break;
      case 4: fe = launch_vcov_factor<4>(dout, dfl, dl, T, M, R, s); break;
      case 5: fe = launch_vcov_factor<5>(dout, dfl, dl, T, M, R, s); break;
      case 6: fe = launch_vcov_factor<6>(dout, dfl, dl, T, M, R, s); break;
      case 7: fe = launch_vcov_factor<7>(dout, dfl, dl, T, M, R, s); break;
      default: fe = launch_vcov_factor<8>(dout, dfl, dl, T, M, R, s); break;
    }
    K9_TRY(fe);
    done_fast = true;
  }
  // draws in slabs of <= 65535 (grid.z)
  for (int r0 = 0; r0 < R && !done_fast; r0 += 65535) {
    const int rc = R - r0 < 65535 ? R - r0 : 65535;
    dim3 grid((unsigned)((T + 127) / 128), (unsigned)(M * M), (unsigned)rc);
    vcov_kernel<<<grid, 128, 0, s>>>(dout + (size_t)r0 * T * M * M, dfl ? dfl + (size_t)r0 * M * nf : nullptr,
                                     dl + (size_t)r0 * T * S, dui ? dui + (size_t)r0 * M * M : nullptr, factor ? 1 : 0, T, M, nf);
  }
  K9_TRY(cudaGetLastError());
  K9_TRY(cudaEventRecord(h->ev[1], s));
  K9_TRY(cudaMemcpyAsync(out, dout, d * nout, cudaMemcpyDeviceToHost, s));
  K9_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = factor ? 1 : 2;
#undef K9_TRY
  return fin(BVB_OK);
}
