// Micro-benchmark: does an L2 -> shared-memory cp.async stream slow down FP64 mma.sync on the same SM
// (and vice versa)?  Warps 0..7 stream, warps 8..15 run DMMA (operands from registers or from LDS).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_vs_dmma stream_vs_dmma.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode bit0: stream, bit1: dmma, bit2: dmma operands through LDS
__global__ void __launch_bounds__(512, 1) k(const double* __restrict__ g, size_t per_cta_doubles, int passes, int dmma_iters, int mode,
                                            long long* out) {
  extern __shared__ __align__(16) double sm[];
  double* ring = sm;                 // 8 warps x 4 stages x 512 doubles (4 KB) = 128 KB
  double* ops = sm + 8 * 4 * 512;    // 16 KB of operands
  __shared__ volatile int done;
  __shared__ __align__(8) uint64_t bars[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 2048; i += 512) ops[i] = 1.0 + 1e-9 * i;
  if (tid == 0) {
    done = 0;
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 8) {
    if (mode & 1) {
      const double* src = g + (size_t)blockIdx.x * per_cta_doubles;
      double* myring = ring + warp * 4 * 512;
      const size_t chunks = per_cta_doubles / (8 * 512);   // each warp takes every 8th 4 KB chunk
      int st = 0;
      long long nchunk = 0;
      for (int pss = 0; (mode & 2) ? (done < 8) : (pss < passes); ++pss)
        for (size_t c = 0; c < chunks; ++c) {
          ++nchunk;
          const double* s = src + (c * 8 + warp) * 512;
          double* d = myring + st * 512;
          const int flavour = (mode >> 3) & 3;
          if (flavour == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d + 2 * (lane + 32 * j))), "l"(s + 2 * (lane + 32 * j)));
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 3;");
          } else if (flavour == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s32(d + 2 * (lane + 32 * j))), "l"(s + 2 * (lane + 32 * j)));
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 3;");
          } else {
            // TMA bulk: one 4 KB copy per chunk issued by lane 0, completion on the stage's mbarrier; before reusing a
            // stage wait for the copy issued 4 chunks ago
            uint64_t* bar = bars + warp * 4 + st;
            if (nchunk > 4) {
              const uint32_t par = (uint32_t)(((nchunk - 5) >> 2) & 1);
              asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(s32(bar)), "r"(par) : "memory");
            }
            if (lane == 0) {
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(4096) : "memory");
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(4096), "r"(s32(bar)) : "memory");
            }
            __syncwarp();
          }
          st = (st + 1) & 3;
        }
      asm volatile("cp.async.wait_group 0;");
      if (((mode >> 3) & 3) == 2) {   // drain the last 4 bulk copies
        for (long long q = (nchunk > 4 ? nchunk - 4 : 0); q < nchunk; ++q) {
          uint64_t* bar = bars + warp * 4 + (int)(q & 3);
          const uint32_t par = (uint32_t)((q >> 2) & 1);
          asm volatile("{ .reg .pred p; W2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W2; }" ::"r"(s32(bar)), "r"(par) : "memory");
        }
      }
      if (lane == 0) out[blockIdx.x * 32 + 16 + warp] = nchunk;
    }
    if (lane == 0) out[blockIdx.x * 32 + warp] = clock64() - t0;
  } else {
    if (mode & 2) {
      double c[12][2];
#pragma unroll
      for (int i = 0; i < 12; ++i) c[i][0] = c[i][1] = 0.0;
      double a[3] = {1.0, 1.0, 1.0}, b[4] = {1.0, 1.0, 1.0, 1.0};
      for (int it = 0; it < dmma_iters; ++it) {
        if (mode & 4) {
          const double* o = ops + ((it & 7) * 64) + lane;
#pragma unroll
          for (int i = 0; i < 3; ++i) a[i] = o[128 * i];
#pragma unroll
          for (int i = 0; i < 4; ++i) b[i] = o[512 + 96 * i];
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma(c[q * 4 + nt][0], c[q * 4 + nt][1], a[q], b[nt]);
      }
      double s = 0;
#pragma unroll
      for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
      if (s == 123.456) out[0] = 1;
    }
    if (lane == 0) { out[blockIdx.x * 32 + warp] = clock64() - t0; atomicAdd((int*)&done, 1); }
  }
}

int main() {
  const int ctas = 100;
  const size_t per = 650 * 1024 / 8;   // 650 KB per CTA, as one equation's packed system
  double* g;
  cudaMalloc(&g, ctas * per * 8);
  { std::vector<double> hbuf(ctas * per); for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (double)((i * 2654435761u) % 1000003) * 1.3e-3; cudaMemcpy(g, hbuf.data(), hbuf.size() * 8, cudaMemcpyHostToDevice); }
  long long* out;
  cudaMalloc(&out, ctas * 32 * 8);
  const int smem = (8 * 4 * 512 + 2048) * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int passes = 16, iters = 1000;   // 2.6 MB per CTA; 8 warps x 8000 x 12 DMMA
  for (int mode : {1, 2, 6, 3, 7, 1 + 8, 7 + 8, 1 + 16, 7 + 16}) {
    k<<<ctas, 512, smem>>>(g, per, passes, iters, mode, out);   // warm L2
    k<<<ctas, 512, smem>>>(g, per, passes, iters, mode, out);
    long long h[32];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long ts = 0, td = 0;
    for (int w = 0; w < 8; ++w) ts = h[w] > ts ? h[w] : ts;
    for (int w = 8; w < 16; ++w) td = h[w] > td ? h[w] : td;
    const char* fl[] = {"cp.async.cg16", "cp.async.ca16", "TMA bulk 4KB", "?"};
    printf("mode %2d (%s%s%s%s): ", mode, mode & 1 ? "stream " : "", mode & 2 ? "dmma " : "", mode & 4 ? "lds-operands " : "", mode & 1 ? fl[(mode >> 3) & 3] : "");
    long long nch = 0; for (int w = 0; w < 8; ++w) nch += h[16 + w];
    if (mode & 1) printf("stream %.1f B/clk/SM (%lld cycles, %lld chunks)  ", (double)nch * 4096 / ts, ts, nch);
    if (mode & 2) printf("dmma %.1f cycles per DMMA per SMSP (%lld cycles)", (double)td / (2.0 * iters * 12), td);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
