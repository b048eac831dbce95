// DFMA issue rate of one warp for the operand patterns of the factorisation's rank-1 update.
#include <cstdio>
#include <cuda_runtime.h>
template <int VAR>
__global__ void k(double* out, long long* cyc, int n, double l0) {
  __shared__ __align__(16) double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = 1e-3 * lane; sm[lane + 32] = 0.5;
  __syncwarp();
  double a[32], t[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = lane + i; t[i] = sm[i] ; }
  double l = l0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < n; ++it) {
    if (VAR == 0) {   // in place, multiplier registers loaded once
#pragma unroll
      for (int i = 0; i < 31; ++i) a[i] = fma(-l, t[i], a[i]);
    } else if (VAR == 1) {   // sliding window
#pragma unroll
      for (int i = 0; i < 31; ++i) a[i] = fma(-l, t[i], a[i + 1]);
      a[31] = 0.0;
    } else if (VAR == 2) {   // in place, multipliers re-read from shared memory every iteration (LDS.128)
      const double2* p = reinterpret_cast<const double2*>(sm + (it & 1) * 32);
#pragma unroll
      for (int i = 0; i < 16; ++i) { const double2 v = p[i]; t[2 * i] = v.x; t[2 * i + 1] = v.y; }
#pragma unroll
      for (int i = 0; i < 31; ++i) a[i] = fma(-l, t[i], a[i]);
    } else if (VAR == 3) {   // sliding + LDS.128
      const double2* p = reinterpret_cast<const double2*>(sm + (it & 1) * 32);
#pragma unroll
      for (int i = 0; i < 16; ++i) { const double2 v = p[i]; t[2 * i] = v.x; t[2 * i + 1] = v.y; }
#pragma unroll
      for (int i = 0; i < 31; ++i) a[i] = fma(-l, t[i], a[i + 1]);
      a[31] = 0.0;
    }
    l = l * 0.999;
  }
  const long long t1 = clock64();
  if (lane == 0) cyc[0] = (t1 - t0) / n;
  double s = 0; for (int i = 0; i < 32; ++i) s += a[i];
  out[lane] = s;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  const char* nm[4] = {"in place, register multipliers", "sliding window, register multipliers", "in place + 16 LDS.128", "sliding window + 16 LDS.128"};
  for (int v = 0; v < 4; ++v) {
    for (int r = 0; r < 2; ++r) { if (v == 0) k<0><<<1, 32>>>(out, cyc, 1000, 1e-3); if (v == 1) k<1><<<1, 32>>>(out, cyc, 1000, 1e-3); if (v == 2) k<2><<<1, 32>>>(out, cyc, 1000, 1e-3); if (v == 3) k<3><<<1, 32>>>(out, cyc, 1000, 1e-3); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %lld cycles per 31 DFMA (%.1f each)\n", nm[v], h, h / 31.0);
  }
  return cudaGetLastError() != cudaSuccess;
}
