// Internal declarations shared by the libbvb.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <string>
#include <vector>

#include "../../include/bvb.h"
#include "common.cuh"

struct bvb_handle {
  int device = 0;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  int err_code = 0, err_step = 0, err_sweep = 0, err_eq = 0;
  char err_msg[512] = {0};
  double last_ms[4] = {0, 0, 0, 0};
  int last_launches = 0;
  cudaEvent_t ev[8] = {};
  void* scratch = nullptr;  // owned by api.cu (PhiWorkspace cache)
  void* gibbs = nullptr;    // owned by gibbs.cu (bvb::Gibbs)
  int rank = 0, world = 1;
  void* comm = nullptr;     // ncclComm_t
  int inproc = 0;           // created by bvb_create_multi: the peers live in this process (direct peer pointers, no IPC)
  unsigned long long p2p_epoch = 0;   // chains set up on this handle (stamps the flags of the peer exchange)
  void* pin[2] = {nullptr, nullptr};   // pinned staging of the stored draws, kept across bvb_gibbs_init calls
  size_t pin_bytes[2] = {0, 0};
};

void bvb_set_error(bvb_handle* h, int code, const char* fmt, ...);

namespace bvb {

// Geometry of one batched regression problem family (all equations share the
// regressor matrix X; only weights / priors / right-hand sides differ).
//
// Packed "augmented" posterior system of one equation: columns c = 0..n-1 of the
// lower triangle of Omega (n x n), each followed by the right-hand side entry
// b[c] as row n.  Element (i, c), c <= i <= n, lives at colbase[c] + i.
// colbase[c] is even, so rows with even index are 16-byte aligned.
struct PairGeom {
  int T = 0, Tp = 0;      // observations, padded to a multiple of K1_BK
  int n = 0;              // regressors (Kp)
  int P = 0;              // n(n+1)/2 pair rows
  int nZtiles = 0;        // ceil(P / 64)
  int nXtiles = 0;        // ceil(n / 64)   (rows of X^T for the rhs GEMM)
  int nArows = 0;         // 64 * (nZtiles + nXtiles)
  long ldo = 0;           // doubles per equation in the packed system (+slack)
  std::vector<int> colbase;   // n entries
  std::vector<int2> lut;      // nArows entries {dest offset | -1, coef index | -1}
  std::vector<int2> pairs;    // nArows entries {a, b} (-1 for padding) for the Z builder
};
void make_pair_geom(PairGeom& g, int T, int n);

constexpr int K1_BM = 64;
constexpr int K1_BK = 20;
constexpr int K1_STAGES = 3;
constexpr int K1_THREADS = 128;
constexpr int K1_NMAX = 104;  // equations per CTA column block (13 n8 tiles)

struct K1Params {
  const double* A;     // [nArows][Tp], K-major
  const double* W;     // [Nalloc][Tp] weights w_j[t]
  const double* WY;    // [Nalloc][Tp] w_j[t] * yhat_j[t]
  const int2* lut;
  const double* V;     // [n][M] prior variances
  const double* PHI0;  // [n][M] prior means (may be null -> 0)
  double* omega;       // [M][ldo]
  int Tp, nZtiles, ntiles, M, n;
  long ldo;
  // expert_huge mode (sample_coefficients.cpp:40-52): rows are pairs (t, s) of observations, the
  // contraction runs over the coefficients, and the epilogue writes
  //   acc * rowscale[t, j] * rowscale[s, j] + (t == s)       (= X_norm V X_norm' + I_T)
  const int2* pairs = nullptr;      // [nArows] {t, s}
  const double* rowscale = nullptr; // [n][M] normaliser exp(-h/2); non-null selects the mode
  unsigned long long* tstamp = nullptr;  // optional [2]: %globaltimer of the first CTA start (atomicMin) / last CTA end (atomicMax)
};
cudaError_t launch_k1(const K1Params& p, cudaStream_t s);

// Peer exchange fused into the K2 back-solve (multi-GPU, equation-sharded sweep): the CTA that has just solved an
// equation writes its PHI column to every peer's exchange buffer as well, the last CTA of the launch stamps the flags.
struct K2Xch {
  double* const* peer_xbuf = nullptr;            // [world] (device table)
  unsigned long long* const* peer_xflag = nullptr;
  int* cnt = nullptr;                            // CTAs of this launch that are done (self-resetting)
  int rank = 0, world = 1;
  size_t blk = 0;                                // doubles per rank block
  unsigned long long epoch_base = 0;
  const unsigned int* sweep_p = nullptr;
};

struct K2Params {
  double* omega;        // [M][ldo] in: packed augmented system; out: its Cholesky factor
  const int* colbase;   // [n]
  const double* z;      // [n][M] standard normals (ld = ldz)
  double* phi;          // [n][M] out (ld = ldphi)
  int* status;          // [M] 0 ok, c+1 = non-positive pivot at column c, -1 = non-finite
  int n, M;
  long ldo;
  int ldz, ldphi;
  long long* prof = nullptr;  // optional [8] phase cycle counters (debug)
  int mode = 0;               // 0: factor + forward solve (rhs row) + back-solve/draw; 1: factor only
  int debug_mode = 0;         // 1: skip the DMMA of the update loop, 2: skip its loads (timing experiments only)
  int direct = 1;             // look-ahead kernel: operands of the update straight from L2 into registers (BVB_K2_DIRECT)
  unsigned long long* tstamp = nullptr;  // optional [2]: %globaltimer of the first CTA start / last CTA end
  K2Xch xch;                  // world > 1: push the solved columns to the peers (honoured by the tiled form's back-solve)
  bool* xch_fused = nullptr;  // host: set to true by launch_k2 when the launch it chose does the push itself
};
cudaError_t launch_k2(const K2Params& p, cudaStream_t s);
cudaError_t launch_k2_solve(const K2Params& p, const double* bvec, int eq0, int count, cudaStream_t s);
size_t k2_smem_bytes(int n);
bool k2_supported(int n);   // systems of size n can be factorised (single-tile or row-tiled panel)

// prep kernels
cudaError_t launch_build_A(double* A, const double* X, const int2* pairs, int T, int Tp, int n,
                           int nArows, int nZrows, cudaStream_t s);
cudaError_t launch_prep_factor(double* W, double* WY, const double* Y, const double* logvar,
                               const double* facload, const double* fac, int T, int Tp, int M,
                               int r, cudaStream_t s);
cudaError_t launch_measure_peak(double* out3, cudaStream_t s, int num_sms);

}  // namespace bvb
