// Dependent-chain latencies of the FP64 operations on the critical path of a Cholesky pivot (one warp, one CTA):
// DFMA, DMUL, rsqrt(), SHFL of a double, LDS.  nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0, int n) {
  __shared__ double sm[64];
  double x = x0 + threadIdx.x * 1e-9, y = 0.999999;
  long long t0, t1;
  sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y;
  __syncwarp();
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, y, 1e-12);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = x * y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  x = fabs(x) + 1.5;
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < n; ++i) { const double v = sm[idx & 63]; idx = (int)(v * 0.0) + ((idx + 1) & 63); x += v; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = 1.0 / (x + 1.5);
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = sqrt(x + 1.5);
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // independent DFMA stream: issue rate
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
  for (int i = 0; i < n; ++i) {
    a0 = fma(a0, y, 1e-12); a1 = fma(a1, y, 1e-12); a2 = fma(a2, y, 1e-12); a3 = fma(a3, y, 1e-12);
    a4 = fma(a4, y, 1e-12); a5 = fma(a5, y, 1e-12); a6 = fma(a6, y, 1e-12); a7 = fma(a7, y, 1e-12);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = x + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + idx;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, 1.0, n);
  long long h[8];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char* nm[8] = {"DFMA dependent", "DMUL dependent", "rsqrt()+DADD dependent", "SHFL double dependent", "LDS + DADD dependent", "1/x + DADD dependent", "sqrt + DADD dependent", "8 independent DFMA (per 8)"};
  for (int i = 0; i < 8; ++i) printf("%-30s %.1f cycles\n", nm[i], (double)h[i] / n);
  return cudaGetLastError() != cudaSuccess;
}
