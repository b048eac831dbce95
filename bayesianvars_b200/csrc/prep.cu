// Small streaming kernels around K1/K2: pair-product matrix builder, weight
// preparation, and the FP64 peak micro-benchmarks that give the roofline
// denominators MEASURED_PEAKS.json lacks (it has HBM and bf16 only).
#include "bvb_internal.cuh"

namespace bvb {

// ---------------------------------------------------------------------------
// Packed layout + row tables (host).
void make_pair_geom(PairGeom& g, int T, int n) {
  g.T = T;
  g.Tp = ((T + K1_BK - 1) / K1_BK) * K1_BK;
  g.n = n;
  g.P = n * (n + 1) / 2;
  // 64-row tiles, both sections padded to a multiple of 6 tiles so that 32-, 48-, 64- and 128-row CTA tiles divide them
  g.nZtiles = ((g.P + K1_BM - 1) / K1_BM + 5) / 6 * 6;
  g.nXtiles = ((n + K1_BM - 1) / K1_BM + 5) / 6 * 6;
  g.nArows = K1_BM * (g.nZtiles + g.nXtiles);
  g.colbase.assign(n > 0 ? n : 1, 0);
  // column c stores rows c..n at colbase[c] + row; colbase[c] is a multiple of 16 doubles (128 bytes), so the 32-row
  // segments the tiled factorisation (k2_tiled.cu) works on - rows 32 I .. 32 I + 31 of a column - are whole 128-byte
  // lines: a line never holds entries of two tiles, which is what makes L1-cached reads of FINISHED tiles safe while
  // other CTAs are still writing their neighbours.  (Even colbase = 16-byte aligned even rows is what K2 v1 / v2 need.)
  long start = 0;  // storage index of element (c, c)
  for (int c = 0; c < n; ++c) {
    while (((start - c) & 15) != 0) ++start;
    g.colbase[c] = (int)(start - c);
    start += (n + 1 - c);
  }
  g.ldo = ((start + 32 + 15) / 16) * 16;  // slack: K2's copies read whole 32-row segments of the last columns; 128-byte rows
  g.lut.assign(g.nArows, make_int2(-1, -1));
  g.pairs.assign(g.nArows, make_int2(-1, -1));
  int row = 0;
  for (int b = 0; b < n; ++b)
    for (int a = b; a < n; ++a, ++row) {
      g.pairs[row] = make_int2(a, b);
      g.lut[row] = make_int2(g.colbase[b] + a, a == b ? b : -1);
    }
  row = K1_BM * g.nZtiles;
  for (int a = 0; a < n; ++a, ++row) {
    g.pairs[row] = make_int2(a, -1);
    g.lut[row] = make_int2(g.colbase[a] + n, a);  // rhs entry of column a (row n)
  }
}

// A[row][t] = X[t,a] * X[t,b]  (pair rows)  |  X[t,a]  (rhs rows)  |  0 (padding)
__global__ void build_A_kernel(double* __restrict__ A, const double* __restrict__ X,
                               const int2* __restrict__ pairs, int T, int Tp, int nArows) {
  const int row = blockIdx.x;
  const int2 pr = pairs[row];
  double* out = A + (size_t)row * Tp;
  for (int t = threadIdx.x; t < Tp; t += blockDim.x) {
    double v = 0.0;
    if (pr.x >= 0 && t < T) {
      v = X[(size_t)pr.x * T + t];
      if (pr.y >= 0) v *= X[(size_t)pr.y * T + t];
    }
    out[t] = v;
  }
}

cudaError_t launch_build_A(double* A, const double* X, const int2* pairs, int T, int Tp, int n,
                           int nArows, int nZrows, cudaStream_t s) {
  (void)n; (void)nZrows;
  build_A_kernel<<<nArows, 128, 0, s>>>(A, X, pairs, T, Tp, nArows);
  return cudaGetLastError();
}

// W[j][t] = exp(-h[t,j]);  WY[j][t] = W * (Y[t,j] - sum_r facload[j,r] fac[r,t])
// (sample_coefficients.cpp:93-95,99-100 with normalizer^2 folded into one weight)
__global__ void prep_factor_kernel(double* __restrict__ W, double* __restrict__ WY,
                                   const double* __restrict__ Y, const double* __restrict__ logvar,
                                   const double* __restrict__ facload, const double* __restrict__ fac,
                                   int T, int Tp, int M, int r) {
  const int j = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double yh = Y[(size_t)j * T + t];
    for (int k = 0; k < r; ++k) yh -= facload[(size_t)k * M + j] * fac[(size_t)t * r + k];
    // exp(-h/2)^2: the reference squares the normaliser through X_norm'X_norm
    const double nrm = exp(-0.5 * logvar[(size_t)j * T + t]);
    const double w = nrm * nrm;
    W[(size_t)j * Tp + t] = w;
    WY[(size_t)j * Tp + t] = w * yh;
  }
}

cudaError_t launch_prep_factor(double* W, double* WY, const double* Y, const double* logvar,
                               const double* facload, const double* fac, int T, int Tp, int M,
                               int r, cudaStream_t s) {
  prep_factor_kernel<<<M, 256, 0, s>>>(W, WY, Y, logvar, facload, fac, T, Tp, M, r);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// FP64 peak micro-benchmarks (register resident, no memory traffic).
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* sink, int iters, long long* cycles) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* sink, int iters, long long* cycles) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
  const double a = 1.0 + 1e-12, b = 1e-12;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

cudaError_t launch_measure_peak(double* out3, cudaStream_t s, int num_sms) {
  double* sink = nullptr;
  long long* cyc = nullptr;
  cudaError_t e = cudaMalloc(&sink, 64);
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&cyc, 64);
  if (e != cudaSuccess) { cudaFree(sink); return e; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = num_sms * 4, threads = 256;
  float ms = 0;
  long long hc = 0;
  // DMMA: per warp per iteration 8 x (8*8*4*2) flop
  {
    const int iters = 20000;
    peak_dmma_kernel<<<blocks, threads, 0, s>>>(sink, 1000, cyc);
    double best = 0, clk = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0, s);
      peak_dmma_kernel<<<blocks, threads, 0, s>>>(sink, iters, cyc);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
      const double flop = (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
      const double tf = flop / (ms * 1e-3) / 1e12;
      if (tf > best) { best = tf; clk = hc / (ms * 1e-3) / 1e6; }
    }
    out3[0] = best;
    out3[2] = clk;
  }
  {
    const int iters = 20000;
    peak_dfma_kernel<<<blocks, threads, 0, s>>>(sink, 1000, cyc);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0, s);
      peak_dfma_kernel<<<blocks, threads, 0, s>>>(sink, iters, cyc);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = (double)blocks * threads * iters * 16.0 * 2.0;
      const double tf = flop / (ms * 1e-3) / 1e12;
      if (tf > best) best = tf;
    }
    out3[1] = best;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  cudaFree(cyc);
  return e;
}

}  // namespace bvb
