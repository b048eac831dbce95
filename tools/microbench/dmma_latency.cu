// Micro-benchmark: FP64 mma.sync m8n8k4 issue interval and dependent-chain latency on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_latency dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void chain(long long* out, double a, double b, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[63] = 1;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int NACC>
void run(long long* d, int warps) {
  const int iters = 2000;
  chain<NACC><<<1, 32 * warps>>>(d, 1.0000001, 0.9999999, iters);
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("warps=%2d independent accumulators=%2d : %7.1f cycles per DMMA per warp (%.1f cycles per DMMA per SMSP)\n", warps, NACC,
         (double)h / (iters * NACC), (double)h / (iters * NACC) / ((warps + 3) / 4));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  for (int warps : {1, 4, 8, 16}) {
    run<1>(d, warps); run<2>(d, warps); run<4>(d, warps); run<8>(d, warps); run<12>(d, warps); run<16>(d, warps);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
