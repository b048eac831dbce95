// Where do the cycles of the ROLLED 32 x 32 factorisation loop go?  Variants of one loop body, one warp.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 32, LS = 36;
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = y * y, e = fma(-x, t, 1.0), p = fma(0.375, e, 0.5), u = y * e;
  return fma(p, u, y);
}
// FLAGS: 1 = no update (chain only), 2 = no shuffle (every lane uses its own dpiv), 4 = no rsqrt (rs = dpiv),
//        8 = update but multipliers not re-read (no LDS), 16 = no STS of the column,
//        32 = + rsbuf store by lane c, 64 = + positivity/finite check with failure code, 128 = software-pipelined body
template <int FLAGS>
__global__ void k(const double* A, double* out, long long* cyc, int reps) {
  __shared__ __align__(16) double Lr[NB * LS];
  __shared__ double dn[NB * LS];
  __shared__ double rsbuf[2 * NB];
  const int lane = threadIdx.x;
  int fail = 0;
  long long total = 0;
  for (int i = lane; i < NB * LS; i += 32) Lr[i] = 1e-3;
  for (int r = 0; r < reps; ++r) {
    for (int c = 0; c < NB; ++c) dn[c * LS + lane] = A[c * NB + lane];
    __syncwarp();
    double w[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) w[i] = (i <= lane) ? dn[i * LS + lane] : 0.0;
    const long long t0 = clock64();
    double xx = __shfl_sync(0xffffffffu, w[0], 0);
    double rs = fast_rsqrt(xx);
    double tl[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) tl[i] = 1e-3 * i;
#pragma unroll 1
    for (int c = 0; c < NB; ++c) {
      const double l = (lane >= c) ? w[0] * rs : 0.0;
      if (!(FLAGS & 16)) { if (lane > c) Lr[c * LS + (lane - c - 1)] = l; }
      const double dpiv = fma(-l, l, w[1]);
      xx = (FLAGS & 2) ? dpiv + 40.0 : __shfl_sync(0xffffffffu, dpiv, (c + 1) & 31);
      if (FLAGS & 32) { if (lane == c) { rsbuf[c] = rs; rsbuf[NB + c] = l; } }
      if (FLAGS & 64) {
        const bool ok = (xx > 0.0) && (xx < 1.7e308);
        fail = (!ok && fail == 0) ? c + 2 : fail;
        xx = ok ? xx : 1.0;
        rs = fast_rsqrt(xx);
      } else
      rs = (FLAGS & 4) ? xx * 1e-2 : fast_rsqrt(fabs(xx) + 1.0);
      __syncwarp();
      if (!(FLAGS & 1)) {
        if (!(FLAGS & 8)) {
          const double2* lcol = reinterpret_cast<const double2*>(Lr + c * LS);
#pragma unroll
          for (int i = 0; i < NB / 2; ++i) { const double2 v = lcol[i]; tl[2 * i] = v.x; tl[2 * i + 1] = v.y; }
        }
#pragma unroll
        for (int i = 0; i < NB - 1; ++i) w[i] = fma(-l, tl[i], w[i + 1]);
        w[NB - 1] = 0.0;
      } else {
        w[0] = w[1] * 0.5 + l; w[1] = w[2] + 1.0;
      }
    }
    total += clock64() - t0;
    double s = rs + fail + rsbuf[lane];
    for (int i = 0; i < NB; ++i) s += w[i];
    out[r * NB + lane] = s;
  }
  if (lane == 0) cyc[0] = total / reps;
}
#define RUN(F, name) { for (int r = 0; r < 2; ++r) k<F><<<1, 32>>>(dA, out, cyc, 13); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-58s %6lld cycles per block (%4.0f per column)\n", name, h, h / 32.0); }
int main() {
  double h[NB * NB];
  for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 40.0 + c : 1.0 / (1 + (r > c ? r - c : c - r));
  double *dA, *out; long long* cyc;
  cudaMalloc(&dA, sizeof(h)); cudaMalloc(&out, 8 * 32 * 16); cudaMalloc(&cyc, 8);
  cudaMemcpy(dA, h, sizeof(h), cudaMemcpyHostToDevice);
  RUN(0, "full loop body");
  RUN(1, "chain only (no rank-1 update)");
  RUN(1 + 16, "chain only, no STS");
  RUN(1 + 2 + 16, "chain only, no STS, no shuffle");
  RUN(1 + 2 + 4 + 16, "chain only, no STS, no shuffle, no rsqrt");
  RUN(8, "update with stale multipliers (no LDS)");
  RUN(2, "full, no shuffle");
  RUN(4, "full, no rsqrt");
  RUN(16, "full, no STS");
  RUN(32, "full + rsbuf stores");
  RUN(64, "full + check");
  RUN(96, "full + rsbuf stores + check");
  return cudaGetLastError() != cudaSuccess;
}
