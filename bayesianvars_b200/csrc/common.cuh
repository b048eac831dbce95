// Shared device/host helpers for the bvb (bayesianVARs on B200) kernels.
// sm_100a only: FP64 tensor work is warp-level mma.sync DMMA (tcgen05 has no
// f64 kind), staged through shared memory with cp.async / TMA bulk copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define BVB_CUDA_TRY(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      bvb_set_error(h, BVB_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__,   \
                    __LINE__, cudaGetErrorString(_e));                            \
      return BVB_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

namespace bvb {

// ---------------------------------------------------------------- DMMA -----
// D(8x8) += A(8x4, row) * B(4x8, col); lane l holds A[l/4][l%4], B[l%4][l/4],
// C/D[l/4][2*(l%4) + {0,1}].  Lowers to one native DMMA.8x8x4 on sm_100a.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------ cp.async -----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte cp.async, `.cg` (bypass L1).  tools/microbench/stream_vs_dmma.cu: a saturated `.cg` stream (42 B/clk/SM)
// halves the rate of DMMA warps whose operands come through LDS, `.ca` streams 76 B/clk/SM and costs LDS less, an
// aligned TMA bulk stream costs it nothing - but at the 4-9 B/clk/SM K1 and K2 actually stream, `.ca` measured 1-4%
// SLOWER than `.cg` in both kernels (-DBVB_CP16_CA to repeat the experiment).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
#ifdef BVB_CP16_CA
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
#endif
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------- mbarrier + TMA bulk copy --
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned).  SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// order generic-proxy accesses (plain ld/st) against later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// ------------------------------------------------------- live kernel timers --
// Device-wide nanosecond timer; a kernel stamps its span into ts[0] (min over CTA starts) / ts[1] (max over CTA ends).
__device__ __forceinline__ unsigned long long bvb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void bvb_stamp_begin(unsigned long long* ts) { if (ts) atomicMin(ts, bvb_globaltimer()); }
__device__ __forceinline__ void bvb_stamp_end(unsigned long long* ts) { if (ts) atomicMax(ts + 1, bvb_globaltimer()); }

// --------------------------------------------------------------- warp ------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__host__ __device__ inline double bvb_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

// -------------------------------------------------------------- Philox -----
// Philox4x32-10 (Salmon et al. 2011).  Counter-based: the 128-bit counter is
// (c0..c3), the 64-bit key is the chain seed.  One call yields 4x32 random bits.
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
  }
  __host__ __device__ inline uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t key0 = k0, key1 = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ key0, n1 = lo1, n2 = hi0 ^ c3 ^ key1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      key0 += 0x9E3779B9u;
      key1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// 53-bit uniform in (0,1) from two 32-bit words (never 0 or 1).
__host__ __device__ inline double u01(uint32_t a, uint32_t b) {
  uint64_t x = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);  // 53 bits
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// Stream of uniforms / normals / gammas tied to one (stream id, element, sweep)
// triple; successive draws bump the attempt counter, so rejection loops on
// different elements never share variates (counter = seed, chain,
// coefficient, sweep, attempt).
struct RngStream {
  Philox ph;
  uint32_t c1, c2, c3;  // element, sweep, stream-id
  uint32_t ctr;         // attempt counter (c0)
  uint4 buf;
  int have;             // number of unused 64-bit lanes in buf (0..2)
  bool has_spare;
  double spare;
  __host__ __device__ RngStream(uint64_t seed, uint32_t stream, uint32_t elem, uint32_t sweep)
      : ph(seed), c1(elem), c2(sweep), c3(stream), ctr(0), have(0), has_spare(false), spare(0.0) {
    buf = make_uint4(0, 0, 0, 0);
  }
  __host__ __device__ inline double uniform() {
    if (have == 0) {
      buf = ph(ctr++, c1, c2, c3);
      have = 2;
    }
    double u = (have == 2) ? u01(buf.x, buf.y) : u01(buf.z, buf.w);
    --have;
    return u;
  }
  __host__ __device__ inline double normal() {
    if (has_spare) {
      has_spare = false;
      return spare;
    }
    double u1 = uniform(), u2 = uniform();
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
#ifdef __CUDA_ARCH__
    sincospi(2.0 * u2, &s, &c);
#else
    s = sin(2.0 * M_PI * u2);
    c = cos(2.0 * M_PI * u2);
#endif
    spare = rad * s;
    has_spare = true;
    return rad * c;
  }
  __host__ __device__ inline double exponential() { return -log(uniform()); }
  // Gamma(shape, scale=1): Marsaglia & Tsang (2000); shape<1 via the
  // U^(1/shape) boost.
  __host__ __device__ inline double gamma(double shape) {
    double boost = 1.0;
    if (shape < 1.0) {
      boost = pow(uniform(), 1.0 / shape);
      shape += 1.0;
    }
    double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
      double x = normal();
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      double u = uniform();
      if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return boost * d * v;
      if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;  // unreachable in practice
  }
};

}  // namespace bvb
