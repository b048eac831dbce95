// Does straight-line (fully unrolled) code run slower when several warps of an SM stream DIFFERENT instruction regions?
// Each warp factorises its own 32 x 32 block with its own copy (template instance) of the same unrolled code.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_multi icache_multi.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 32, RS = 36;
template <int COPY>
__device__ __forceinline__ void factor(double* Lc, const double* dn, int lane, double salt) {
  double av[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) av[c] = (c <= lane) ? dn[c * RS + lane] + salt * COPY : 0.0;
  double xx = __shfl_sync(0xffffffffu, av[0], 0);
  double rs = rsqrt(xx);
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    const double l = (lane >= c) ? av[c] * rs : 0.0;
    av[c] = l;
    Lc[c * RS + lane] = l;
    __syncwarp();
    if (c + 1 < NB) {
      av[c + 1] = fma(-l, Lc[c * RS + c + 1], av[c + 1]);
      xx = __shfl_sync(0xffffffffu, av[c + 1], c + 1);
      rs = rsqrt(xx);
    }
#pragma unroll
    for (int c2 = c + 2; c2 < NB; ++c2) av[c2] = fma(-l, Lc[c * RS + c2], av[c2]);
  }
}
template <bool DISTINCT>
__global__ void k(const double* A, double* out, long long* cyc, int reps) {
  extern __shared__ double smx[];
  double (*Lc)[NB * RS] = reinterpret_cast<double (*)[NB * RS]>(smx);
  double (*dn)[NB * RS] = reinterpret_cast<double (*)[NB * RS]>(smx + 8 * NB * RS);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = 0; c < NB; ++c) dn[w][c * RS + lane] = A[c * NB + lane];
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (!DISTINCT) factor<0>(Lc[w], dn[w], lane, 0.0);
    else switch (w) {
      case 0: factor<0>(Lc[w], dn[w], lane, 0.0); break;
      case 1: factor<1>(Lc[w], dn[w], lane, 0.0); break;
      case 2: factor<2>(Lc[w], dn[w], lane, 0.0); break;
      case 3: factor<3>(Lc[w], dn[w], lane, 0.0); break;
      case 4: factor<4>(Lc[w], dn[w], lane, 0.0); break;
      case 5: factor<5>(Lc[w], dn[w], lane, 0.0); break;
      case 6: factor<6>(Lc[w], dn[w], lane, 0.0); break;
      default: factor<7>(Lc[w], dn[w], lane, 0.0); break;
    }
    __syncwarp();
  }
  const long long t1 = clock64();
  if (lane == 0) cyc[w] = (t1 - t0) / reps;
  out[threadIdx.x] = Lc[w][lane * RS + lane];
}
int main() {
  double h[NB * NB];
  for (int c = 0; c < NB; ++c) for (int r = 0; r < NB; ++r) h[c * NB + r] = (r == c) ? 40.0 + c : 1.0 / (1 + (r > c ? r - c : c - r));
  double *dA, *out; long long* cyc;
  cudaMalloc(&dA, sizeof(h)); cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 64);
  cudaMemcpy(dA, h, sizeof(h), cudaMemcpyHostToDevice);
  const int smb = 16 * NB * RS * 8;
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb);
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb);
  for (int distinct = 0; distinct < 2; ++distinct)
    for (int nw = 1; nw <= 8; nw *= 2) {
      for (int rep = 0; rep < 2; ++rep) { if (distinct) k<true><<<1, 32 * nw, smb>>>(dA, out, cyc, 13); else k<false><<<1, 32 * nw, smb>>>(dA, out, cyc, 13); }
      long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
      printf("%s code, %d warps: %lld cycles per block (warp 0)\n", distinct ? "DISTINCT" : "shared  ", nw, c[0]);
    }
  return cudaGetLastError() != cudaSuccess;
}
