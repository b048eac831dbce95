// C ABI of libbvb.so (see include/bvb.h).  Host orchestration only: every
// number a caller sees is produced by the sm_100a kernels; there is no CPU path.
#include "bvb_internal.cuh"
#include "gig.cuh"
#include "chol.cuh"
#include "huge.cuh"
#include "fsv.cuh"

#include <string.h>
#include <stdlib.h>
#include <new>
#include <algorithm>

using namespace bvb;

void bvb_gibbs_release(bvb_handle* h);
void bvb_comm_release(bvb_handle* h);

void bvb_set_error(bvb_handle* h, int code, const char* fmt, ...) {
  if (!h) return;
  h->err_code = code;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h->err_msg, sizeof(h->err_msg), fmt, ap);
  va_end(ap);
}

namespace {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  // `st` must be the stream the buffer is subsequently used on: a memset on the
  // legacy default stream is not ordered with a cudaStreamNonBlocking stream.
  cudaError_t ensure(size_t count, bool zero = false, cudaStream_t st = nullptr) {
    if (count > n) {
      if (p) cudaFree(p);
      p = nullptr;
      n = 0;
      cudaError_t e = cudaMalloc(&p, count * sizeof(T));
      if (e != cudaSuccess) return e;
      n = count;
      if (zero) {
        e = cudaMemsetAsync(p, 0, count * sizeof(T), st);
        if (e != cudaSuccess) return e;
        if (!st) return cudaDeviceSynchronize();
      }
    }
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// Device-resident state of one regression family (fixed T, n, M).
struct PhiWorkspace {
  PairGeom g;
  int M = 0, r = 0, Nalloc = 0;
  DevBuf<double> X, Y, A, W, WY, omega, V, PHI0, PHI, Z, logvar, lam, fac;
  DevBuf<int2> lut, pairs;
  DevBuf<int> colbase, status;
  std::vector<int> h_status;
  void release() {
    X.release(); Y.release(); A.release(); W.release(); WY.release(); omega.release();
    V.release(); PHI0.release(); PHI.release(); Z.release(); logvar.release(); lam.release();
    fac.release(); lut.release(); pairs.release(); colbase.release(); status.release();
  }
};

// Scratch of the cholesky-structure entry points (bvb_phi_draw_cholesky, bvb_sample_u).
struct CholWorkspace {
  PairGeom g;          // geometry of the PHI systems (T, Kp)
  PairGeom gu;         // geometry of the U systems (T, M-1)
  DevBuf<double> X, Y, PHI0, V, U, dsq, Zn, PHI, A, W, WYzero, omega, d2inv, resid, XPHI, Q, rvec, bvec;
  DevBuf<double> Au, Wu, WYu, omegau, Vu, Zu, phiu, ViU, zflat;
  DevBuf<int2> lut, pairs, lutu, pairsu;
  DevBuf<int> colbase, colbaseu, status;
  void release() {
    for (DevBuf<double>* b : {&X, &Y, &PHI0, &V, &U, &dsq, &Zn, &PHI, &A, &W, &WYzero, &omega, &d2inv, &resid, &XPHI, &Q,
                              &rvec, &bvec, &Au, &Wu, &WYu, &omegau, &Vu, &Zu, &phiu, &ViU, &zflat}) b->release();
    lut.release(); pairs.release(); lutu.release(); pairsu.release();
    colbase.release(); colbaseu.release(); status.release();
  }
};

// Scratch of the expert_huge branch of bvb_phi_draw_factor.
struct HugeWorkspace {
  PairGeom g;   // pairs of observations: n = T, contraction length Kp
  DevBuf<double> X, XT, Y, logvar, V, PHI0, lam, fac, z1, delta, PHI, A, Wv, WYzero, omega, u, nrm, wsol, zero_tm;
  DevBuf<int2> lut, pairs;
  DevBuf<int> colbase, status;
  void release() {
    for (DevBuf<double>* b : {&X, &XT, &Y, &logvar, &V, &PHI0, &lam, &fac, &z1, &delta, &PHI, &A, &Wv, &WYzero, &omega, &u,
                              &nrm, &wsol, &zero_tm}) b->release();
    lut.release(); pairs.release(); colbase.release(); status.release();
  }
};

struct Scratch {
  PhiWorkspace phi;
  CholWorkspace chol;
  HugeWorkspace huge;
  uint64_t call_counter = 0;
};

Scratch* scratch_of(bvb_handle* h) { return static_cast<Scratch*>(h->scratch); }

__global__ void fill_normals_kernel(double* out, size_t n, uint64_t seed, uint32_t stream) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RngStream rng(seed, stream, (uint32_t)i, (uint32_t)(i >> 32));
  out[i] = rng.normal();
}

// out[i + n*j] = GIG(lambda[j % ll], chi[j % lc], psi[j % lp])   (gig_vec.cpp:20-43)
__global__ void gig_vec_kernel(double* out, int n, int m, const double* lambda, int ll, const double* chi, int lc,
                               const double* psi, int lp, uint64_t seed, uint64_t stream) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * m) return;
  const int j = (int)(idx / n);
  RngStream rng(seed, 0x6000u + (uint32_t)(stream & 0xfffu), (uint32_t)idx, (uint32_t)(stream >> 12));
  out[idx] = gig_draw(rng, lambda[j % ll], chi[j % lc], psi[j % lp]);
}

int setup_geometry(bvb_handle* h, PhiWorkspace& ws, int T, int n, int M, int r) {
  if (ws.g.T != T || ws.g.n != n) {
    make_pair_geom(ws.g, T, n);
    BVB_CUDA_TRY(ws.lut.ensure(ws.g.nArows));
    BVB_CUDA_TRY(ws.pairs.ensure(ws.g.nArows));
    BVB_CUDA_TRY(ws.colbase.ensure(ws.g.colbase.size()));
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.lut.p, ws.g.lut.data(), sizeof(int2) * ws.g.nArows, cudaMemcpyHostToDevice, h->stream));
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.pairs.p, ws.g.pairs.data(), sizeof(int2) * ws.g.nArows, cudaMemcpyHostToDevice, h->stream));
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.colbase.p, ws.g.colbase.data(), sizeof(int) * ws.g.colbase.size(), cudaMemcpyHostToDevice, h->stream));
    BVB_CUDA_TRY(cudaStreamSynchronize(h->stream));  // host vectors may be rebuilt by the next call
    BVB_CUDA_TRY(ws.A.ensure((size_t)ws.g.nArows * ws.g.Tp));
    BVB_CUDA_TRY(ws.X.ensure((size_t)T * n));
  }
  const int Nalloc = ((M + 7) / 8) * 8 + K1_NMAX;
  if (ws.M != M || ws.Nalloc != Nalloc || ws.W.n < (size_t)Nalloc * ws.g.Tp) {
    ws.W.release();
    ws.WY.release();
    BVB_CUDA_TRY(ws.W.ensure((size_t)Nalloc * ws.g.Tp, true, h->stream));
    BVB_CUDA_TRY(ws.WY.ensure((size_t)Nalloc * ws.g.Tp, true, h->stream));
    ws.Nalloc = Nalloc;
  }
  ws.M = M;
  ws.r = r;
  BVB_CUDA_TRY(ws.omega.ensure((size_t)M * ws.g.ldo));
  BVB_CUDA_TRY(ws.Y.ensure((size_t)T * M));
  BVB_CUDA_TRY(ws.logvar.ensure((size_t)T * M));
  BVB_CUDA_TRY(ws.V.ensure((size_t)n * M));
  BVB_CUDA_TRY(ws.PHI0.ensure((size_t)n * M));
  BVB_CUDA_TRY(ws.PHI.ensure((size_t)n * M));
  BVB_CUDA_TRY(ws.Z.ensure((size_t)n * M));
  BVB_CUDA_TRY(ws.status.ensure(M));
  if (r > 0) {
    BVB_CUDA_TRY(ws.lam.ensure((size_t)M * r));
    BVB_CUDA_TRY(ws.fac.ensure((size_t)r * T));
  }
  ws.h_status.resize(M);
  return BVB_OK;
}

}  // namespace

static int phi_draw_factor_huge(bvb_handle* h, int T, int Kp, int M, int r, const double* X, const double* Y,
                                const double* logvar_idi, const double* V_prior, const double* PHI0,
                                const double* facload, const double* fac, const double* z, const double* delta,
                                double* PHI_out) {
  if (!k2_supported(T)) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "bvb_phi_draw_factor(huge): T=%d exceeds the in-SM panel of the T x T solve", T);
    return BVB_ERR_UNSUPPORTED;
  }
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  Scratch* sc = scratch_of(h);
  HugeWorkspace& ws = sc->huge;
  cudaStream_t s = h->stream;
  make_pair_geom(ws.g, Kp, T);   // contraction over the Kp coefficients, systems of order T
  const int Kp_pad = ws.g.Tp;
  const size_t tm = (size_t)T * M, km = (size_t)Kp * M, d = sizeof(double);
  const int Nalloc = ((M + 7) / 8) * 8 + K1_NMAX;
  BVB_CUDA_TRY(ws.lut.ensure(ws.g.nArows)); BVB_CUDA_TRY(ws.pairs.ensure(ws.g.nArows));
  BVB_CUDA_TRY(ws.colbase.ensure(ws.g.colbase.size()));
  BVB_CUDA_TRY(ws.A.ensure((size_t)ws.g.nArows * Kp_pad));
  ws.Wv.release(); ws.WYzero.release(); ws.zero_tm.release();
  BVB_CUDA_TRY(ws.Wv.ensure((size_t)Nalloc * Kp_pad, true, s));
  BVB_CUDA_TRY(ws.WYzero.ensure((size_t)Nalloc * Kp_pad, true, s));
  BVB_CUDA_TRY(ws.zero_tm.ensure(tm, true, s));
  BVB_CUDA_TRY(ws.omega.ensure((size_t)M * ws.g.ldo));
  BVB_CUDA_TRY(ws.X.ensure((size_t)T * Kp)); BVB_CUDA_TRY(ws.XT.ensure((size_t)T * Kp)); BVB_CUDA_TRY(ws.Y.ensure(tm));
  BVB_CUDA_TRY(ws.logvar.ensure(tm)); BVB_CUDA_TRY(ws.V.ensure(km)); BVB_CUDA_TRY(ws.PHI0.ensure(km));
  BVB_CUDA_TRY(ws.z1.ensure(km)); BVB_CUDA_TRY(ws.delta.ensure(tm)); BVB_CUDA_TRY(ws.PHI.ensure(km));
  BVB_CUDA_TRY(ws.u.ensure(km)); BVB_CUDA_TRY(ws.nrm.ensure(tm)); BVB_CUDA_TRY(ws.wsol.ensure(tm));
  BVB_CUDA_TRY(ws.status.ensure(M));
  if (r > 0) { BVB_CUDA_TRY(ws.lam.ensure((size_t)M * r)); BVB_CUDA_TRY(ws.fac.ensure((size_t)r * T)); }
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.lut.p, ws.g.lut.data(), sizeof(int2) * ws.g.nArows, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.pairs.p, ws.g.pairs.data(), sizeof(int2) * ws.g.nArows, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.colbase.p, ws.g.colbase.data(), sizeof(int) * ws.g.colbase.size(), cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.X.p, X, d * T * Kp, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.Y.p, Y, d * tm, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.logvar.p, logvar_idi, d * tm, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.V.p, V_prior, d * km, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.PHI0.p, PHI0, d * km, cudaMemcpyHostToDevice, s));
  if (r > 0) {
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.lam.p, facload, d * M * r, cudaMemcpyHostToDevice, s));
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.fac.p, fac, d * r * T, cudaMemcpyHostToDevice, s));
  }
  if (z) BVB_CUDA_TRY(cudaMemcpyAsync(ws.z1.p, z, d * km, cudaMemcpyHostToDevice, s));
  else fill_normals_kernel<<<(unsigned)((km + 255) / 256), 256, 0, s>>>(ws.z1.p, km, h->seed, (uint32_t)(0x5100u + sc->call_counter));
  if (delta) BVB_CUDA_TRY(cudaMemcpyAsync(ws.delta.p, delta, d * tm, cudaMemcpyHostToDevice, s));
  else fill_normals_kernel<<<(unsigned)((tm + 255) / 256), 256, 0, s>>>(ws.delta.p, tm, h->seed, (uint32_t)(0x5200u + sc->call_counter));
  sc->call_counter++;
  BVB_CUDA_TRY(cudaMemsetAsync(ws.status.p, 0, sizeof(int) * M, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[0], s));
  BVB_CUDA_TRY(launch_huge_setup(ws.XT.p, ws.A.p, ws.X.p, ws.pairs.p, T, Kp, Kp_pad, ws.g.nArows, s));
  HugeCtx c;
  c.T = T; c.Kp = Kp; c.Kp_pad = Kp_pad; c.M = M; c.r = r; c.X = ws.X.p; c.Y = ws.Y.p; c.logvar = ws.logvar.p; c.V = ws.V.p;
  c.PHI0 = ws.PHI0.p; c.facload = ws.lam.p; c.fac = ws.fac.p; c.z1 = ws.z1.p; c.delta = ws.delta.p; c.PHI = ws.PHI.p;
  c.A = ws.A.p; c.lut_h = ws.lut.p; c.pairs_h = ws.pairs.p; c.colbase_h = ws.colbase.p; c.nZtiles = ws.g.nZtiles;
  c.ntiles = ws.g.nZtiles + ws.g.nXtiles; c.ldo = ws.g.ldo; c.Wv = ws.Wv.p; c.WYzero = ws.WYzero.p; c.omega = ws.omega.p;
  c.u = ws.u.p; c.nrm = ws.nrm.p; c.wsol = ws.wsol.p; c.zero_tm = ws.zero_tm.p; c.status = ws.status.p;
  BVB_CUDA_TRY(launch_phi_huge(c, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[1], s));
  std::vector<int> st(M);
  BVB_CUDA_TRY(cudaMemcpyAsync(PHI_out, ws.PHI.p, d * km, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(st.data(), ws.status.p, sizeof(int) * M, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = 7;
  for (int j = 0; j < M; ++j)
    if (st[j] != 0) {
      h->err_eq = j + 1;
      bvb_set_error(h, BVB_ERR_NUMERIC, "univariate_regression_update() failed in %ith eqation: solve(): solution not found", j + 1);
      return BVB_ERR_NUMERIC;
    }
  return BVB_OK;
}

extern "C" {

const char* bvb_version(void) { return "bvb 0.1 (sm_100a)"; }

int bvb_create(bvb_handle** out, int device, uint64_t seed) {
  if (!out) return BVB_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
    return BVB_ERR_CUDA;  // no CPU fallback: the caller must fail loudly
  if (cudaSetDevice(device) != cudaSuccess) return BVB_ERR_CUDA;
  bvb_handle* h = new (std::nothrow) bvb_handle();
  if (!h) return BVB_ERR_CUDA;
  h->device = device;
  h->seed = seed;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return BVB_ERR_CUDA; }
  h->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {  // kernels are sm_100a-only
    delete h;
    return BVB_ERR_UNSUPPORTED;
  }
  {
    // highest priority: the side branch of the sweep graph (hyper-parameters) runs at the lowest, so the critical
    // path's blocks are placed first when both are pending
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi) != cudaSuccess) { delete h; return BVB_ERR_CUDA; }
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  h->scratch = new (std::nothrow) Scratch();
  *out = h;
  return BVB_OK;
}

// New Philox key for the calls that follow (a chain set up by bvb_gibbs_init keeps using the handle's key, so call it
// before bvb_gibbs_init).  The Rcpp shim keeps ONE process-wide handle and re-keys it from R's RNG at every entry point.
int bvb_set_seed(bvb_handle* h, uint64_t seed) {
  if (!h) return BVB_ERR_ARG;
  h->seed = seed;
  return BVB_OK;
}

void bvb_destroy(bvb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  bvb_gibbs_release(h);
  bvb_comm_release(h);
  if (h->scratch) {
    scratch_of(h)->phi.release();
    scratch_of(h)->chol.release();
    scratch_of(h)->huge.release();
    delete scratch_of(h);
  }
  for (int b = 0; b < 2; ++b) if (h->pin[b]) cudaFreeHost(h->pin[b]);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int bvb_last_error(const bvb_handle* h, int* step, int* sweep, int* equation, char* msg, int msg_len) {
  if (!h) return BVB_ERR_ARG;
  if (step) *step = h->err_step;
  if (sweep) *sweep = h->err_sweep;
  if (equation) *equation = h->err_eq;
  if (msg && msg_len > 0) {
    strncpy(msg, h->err_msg, msg_len - 1);
    msg[msg_len - 1] = 0;
  }
  return h->err_code;
}

int bvb_last_timing(const bvb_handle* h, double* out4) {
  if (!h || !out4) return 0;
  for (int i = 0; i < 4; ++i) out4[i] = h->last_ms[i];
  return h->last_launches;
}

int bvb_measure_fp64_peak(bvb_handle* h, double* out3) {
  if (!h || !out3) return BVB_ERR_ARG;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  BVB_CUDA_TRY(launch_measure_peak(out3, h->stream, h->num_sms));
  return BVB_OK;
}

int bvb_phi_draw_factor(bvb_handle* h, int T, int Kp, int M, int r, const double* X,
                        const double* Y, const double* logvar_idi, const double* V_prior,
                        const double* PHI0, const double* facload, const double* fac,
                        const double* z, const double* delta, int huge, double* PHI_out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 1; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (T <= 0 || Kp <= 0 || M <= 0 || r < 0 || !X || !Y || !logvar_idi || !V_prior || !PHI0 || !PHI_out ||
      (r > 0 && (!facload || !fac))) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_phi_draw_factor: invalid argument");
    return BVB_ERR_ARG;
  }
  if (huge) return phi_draw_factor_huge(h, T, Kp, M, r, X, Y, logvar_idi, V_prior, PHI0, facload, fac, z, delta, PHI_out);
  if (!k2_supported(Kp)) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "bvb_phi_draw_factor: Kp=%d exceeds the in-SM panel of the standard path; use huge", Kp);
    return BVB_ERR_UNSUPPORTED;
  }
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  Scratch* sc = scratch_of(h);
  PhiWorkspace& ws = sc->phi;
  int rc = setup_geometry(h, ws, T, Kp, M, r);
  if (rc != BVB_OK) return rc;
  cudaStream_t s = h->stream;
  const size_t szTM = sizeof(double) * (size_t)T * M, szKM = sizeof(double) * (size_t)Kp * M;
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.X.p, X, sizeof(double) * (size_t)T * Kp, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.Y.p, Y, szTM, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.logvar.p, logvar_idi, szTM, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.V.p, V_prior, szKM, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.PHI0.p, PHI0, szKM, cudaMemcpyHostToDevice, s));
  if (r > 0) {
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.lam.p, facload, sizeof(double) * (size_t)M * r, cudaMemcpyHostToDevice, s));
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.fac.p, fac, sizeof(double) * (size_t)r * T, cudaMemcpyHostToDevice, s));
  }
  if (z) {
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.Z.p, z, szKM, cudaMemcpyHostToDevice, s));
  } else {
    const size_t nz = (size_t)Kp * M;
    fill_normals_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, s>>>(ws.Z.p, nz, h->seed, (uint32_t)(0x5000u + sc->call_counter));
  }
  sc->call_counter++;

  BVB_CUDA_TRY(cudaEventRecord(h->ev[0], s));
  BVB_CUDA_TRY(launch_build_A(ws.A.p, ws.X.p, ws.pairs.p, T, ws.g.Tp, Kp, ws.g.nArows, 0, s));
  BVB_CUDA_TRY(launch_prep_factor(ws.W.p, ws.WY.p, ws.Y.p, ws.logvar.p, ws.lam.p, ws.fac.p, T, ws.g.Tp, M, r, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[1], s));
  K1Params k1{};
  k1.A = ws.A.p; k1.W = ws.W.p; k1.WY = ws.WY.p; k1.lut = ws.lut.p; k1.V = ws.V.p; k1.PHI0 = ws.PHI0.p;
  k1.omega = ws.omega.p; k1.Tp = ws.g.Tp; k1.nZtiles = ws.g.nZtiles; k1.ntiles = ws.g.nZtiles + ws.g.nXtiles;
  k1.M = M; k1.n = Kp; k1.ldo = ws.g.ldo;
  BVB_CUDA_TRY(launch_k1(k1, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[2], s));
  BVB_CUDA_TRY(cudaMemsetAsync(ws.status.p, 0, sizeof(int) * M, s));
  K2Params k2{};
  k2.omega = ws.omega.p; k2.colbase = ws.colbase.p; k2.z = ws.Z.p; k2.phi = ws.PHI.p; k2.status = ws.status.p;
  k2.n = Kp; k2.M = M; k2.ldo = ws.g.ldo; k2.ldz = Kp; k2.ldphi = Kp;
  static long long* d_prof = nullptr;
  if (getenv("BVB_K2_DEBUG")) k2.debug_mode = atoi(getenv("BVB_K2_DEBUG"));
  if (getenv("BVB_K2_PROF") || getenv("BVB_K2T_PROF")) {
    if (!d_prof) cudaMalloc(&d_prof, 64 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 64 * sizeof(long long), s);
    k2.prof = d_prof;
  }
  BVB_CUDA_TRY(launch_k2(k2, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[3], s));
  if (k2.prof) {
    long long hp[64];
    cudaMemcpyAsync(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (getenv("BVB_K2T_PROF")) {
      fprintf(stderr, "K2 tiled chain CTA 0, cycles over all block columns | warp 0: wait-diag %lld syrk %lld load+first-pivot %lld factor-loop %lld stores %lld | warp 1: to-barrier %lld wait-operands %lld "
                      "loads+product %lld wait-Linv+solve+stores %lld | warp 3: poll %lld product+solve %lld wait-stores+publish %lld poll-D %lld one-L2-read %lld next-diag(stores) %lld [issue %lld first-landed %lld]\n",
              hp[0], hp[2], hp[4], hp[3], hp[8], hp[12], hp[13], hp[14], hp[15], hp[16], hp[17], hp[18], hp[19], hp[20], hp[21], hp[22], hp[23]);
      fprintf(stderr, "K2 tiled look-ahead tasks of equation 0: %lld tasks; cycles: wait-history %lld acc %lld P+C %lld wait-Linv %lld trsm+store+publish %lld "
                      "wait-D %lld D-update %lld\n", hp[39], hp[32], hp[33], hp[34], hp[35], hp[36], hp[37], hp[38]);
    } else
    fprintf(stderr, "K2 phase cycles (CTA 0) v1: gemm fill factor invert trsm writeback backsolve | v2: partA wait-panel trsm wb+partB fill - backsolve:\n  %lld %lld %lld %lld %lld %lld %lld\n", hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6]);
    fprintf(stderr, "K2 v2 panel warp: factor+invert %lld  wait#1 %lld  wait#2 %lld  wait#3 %lld  wait#4 %lld\n", hp[24], hp[25], hp[26], hp[27], hp[28]);
    fprintf(stderr, "K2 gemm cycles per block column:");
    for (int k = 0; k < (Kp + 31) / 32 && k < 16; ++k) fprintf(stderr, " %lld", hp[8 + k]);
    fprintf(stderr, "\n");
  }
  BVB_CUDA_TRY(cudaMemcpyAsync(PHI_out, ws.PHI.p, szKM, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.h_status.data(), ws.status.p, sizeof(int) * M, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->last_ms[0] = ms;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->last_ms[1] = ms;
  cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->last_ms[2] = ms;
  h->last_ms[3] = 0;
  h->last_launches = 4 + (z ? 0 : 1);
  for (int j = 0; j < M; ++j) {
    if (ws.h_status[j] != 0) {
      h->err_eq = j + 1;
      if (ws.h_status[j] > 0)
        bvb_set_error(h, BVB_ERR_NUMERIC, "univariate_regression_update() failed in %ith eqation: chol(): decomposition failed (pivot %d)", j + 1, ws.h_status[j]);
      else
        bvb_set_error(h, BVB_ERR_NUMERIC, "univariate_regression_update() failed in %ith eqation: non-finite draw", j + 1);
      return BVB_ERR_NUMERIC;
    }
  }
  return BVB_OK;
}

int bvb_phi_draw_cholesky(bvb_handle* h, int T, int Kp, int M, const double* PHI_in,
                          const double* PHI0, const double* Y, const double* X, const double* U,
                          const double* d_sqrt, const double* V_prior, const double* z,
                          double* PHI_out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 1; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (T <= 0 || Kp <= 0 || M <= 0 || !PHI_in || !PHI0 || !Y || !X || !U || !d_sqrt || !V_prior || !PHI_out) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_phi_draw_cholesky: invalid argument");
    return BVB_ERR_ARG;
  }
  if (!k2_supported(Kp)) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "bvb_phi_draw_cholesky: Kp=%d exceeds the in-SM panel", Kp);
    return BVB_ERR_UNSUPPORTED;
  }
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  Scratch* sc = scratch_of(h);
  CholWorkspace& ws = sc->chol;
  cudaStream_t s = h->stream;
  make_pair_geom(ws.g, T, Kp);
  const size_t tm = (size_t)T * M, km = (size_t)Kp * M;
  const int Nalloc = ((M + 7) / 8) * 8 + K1_NMAX;
  BVB_CUDA_TRY(ws.lut.ensure(ws.g.nArows)); BVB_CUDA_TRY(ws.pairs.ensure(ws.g.nArows));
  BVB_CUDA_TRY(ws.colbase.ensure(ws.g.colbase.size()));
  BVB_CUDA_TRY(ws.A.ensure((size_t)ws.g.nArows * ws.g.Tp));
  ws.W.release(); ws.WYzero.release();
  BVB_CUDA_TRY(ws.W.ensure((size_t)Nalloc * ws.g.Tp, true, s));
  BVB_CUDA_TRY(ws.WYzero.ensure((size_t)Nalloc * ws.g.Tp, true, s));
  BVB_CUDA_TRY(ws.omega.ensure((size_t)M * ws.g.ldo));
  BVB_CUDA_TRY(ws.X.ensure((size_t)T * Kp)); BVB_CUDA_TRY(ws.Y.ensure(tm)); BVB_CUDA_TRY(ws.PHI0.ensure(km));
  BVB_CUDA_TRY(ws.V.ensure(km)); BVB_CUDA_TRY(ws.U.ensure((size_t)M * M)); BVB_CUDA_TRY(ws.dsq.ensure(tm));
  BVB_CUDA_TRY(ws.Zn.ensure(km)); BVB_CUDA_TRY(ws.PHI.ensure(km)); BVB_CUDA_TRY(ws.d2inv.ensure(tm));
  BVB_CUDA_TRY(ws.resid.ensure(tm)); BVB_CUDA_TRY(ws.XPHI.ensure(tm)); BVB_CUDA_TRY(ws.Q.ensure(tm));
  BVB_CUDA_TRY(ws.rvec.ensure(T)); BVB_CUDA_TRY(ws.bvec.ensure(Kp)); BVB_CUDA_TRY(ws.status.ensure(M));
  const size_t d = sizeof(double);
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.lut.p, ws.g.lut.data(), sizeof(int2) * ws.g.nArows, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.pairs.p, ws.g.pairs.data(), sizeof(int2) * ws.g.nArows, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.colbase.p, ws.g.colbase.data(), sizeof(int) * ws.g.colbase.size(), cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.X.p, X, d * T * Kp, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.Y.p, Y, d * tm, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.PHI0.p, PHI0, d * km, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.V.p, V_prior, d * km, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.U.p, U, d * M * M, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.dsq.p, d_sqrt, d * tm, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.PHI.p, PHI_in, d * km, cudaMemcpyHostToDevice, s));
  if (z) {
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.Zn.p, z, d * km, cudaMemcpyHostToDevice, s));
  } else {
    fill_normals_kernel<<<(unsigned)((km + 255) / 256), 256, 0, s>>>(ws.Zn.p, km, h->seed, (uint32_t)(0x5800u + sc->call_counter));
  }
  sc->call_counter++;
  BVB_CUDA_TRY(cudaMemsetAsync(ws.status.p, 0, sizeof(int) * M, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[0], s));
  BVB_CUDA_TRY(launch_build_A(ws.A.p, ws.X.p, ws.pairs.p, T, ws.g.Tp, Kp, ws.g.nArows, 0, s));
  CholPhiCtx c;
  c.T = T; c.Tp = ws.g.Tp; c.Kp = Kp; c.M = M;
  c.X = ws.X.p; c.Y = ws.Y.p; c.PHI0 = ws.PHI0.p; c.V = ws.V.p; c.U = ws.U.p; c.dsq = ws.dsq.p; c.Zn = ws.Zn.p; c.PHI = ws.PHI.p;
  c.A = ws.A.p; c.lut = ws.lut.p; c.colbase = ws.colbase.p; c.nZtiles = ws.g.nZtiles; c.ntiles = ws.g.nZtiles + ws.g.nXtiles;
  c.ldo = ws.g.ldo; c.W = ws.W.p; c.WYzero = ws.WYzero.p; c.omega = ws.omega.p; c.d2inv = ws.d2inv.p; c.resid = ws.resid.p;
  c.XPHI = ws.XPHI.p; c.Q = ws.Q.p; c.rvec = ws.rvec.p; c.bvec = ws.bvec.p; c.status = ws.status.p;
  BVB_CUDA_TRY(launch_chol_phi(c, s));
  BVB_CUDA_TRY(cudaEventRecord(h->ev[1], s));
  std::vector<int> st(M);
  BVB_CUDA_TRY(cudaMemcpyAsync(PHI_out, ws.PHI.p, d * km, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(st.data(), ws.status.p, sizeof(int) * M, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->last_ms[0] = 0; h->last_ms[1] = 0; h->last_ms[2] = 0; h->last_ms[3] = ms;
  h->last_launches = 8 + 3 * M;
  for (int j = 0; j < M; ++j)
    if (st[j] != 0) {
      h->err_eq = j + 1;
      bvb_set_error(h, BVB_ERR_NUMERIC, "univariate_regression_update() failed in %ith eqation:\nchol(): decomposition failed", j + 1);
      return BVB_ERR_NUMERIC;
    }
  return BVB_OK;
}

int bvb_sample_u(bvb_handle* h, int T, int M, const double* resid, const double* V_i_U,
                 const double* d_sqrt, const double* z, double* U_out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 3; h->err_sweep = 0; h->err_eq = 0; h->err_msg[0] = 0;
  if (T <= 0 || M <= 0 || !resid || !d_sqrt || !U_out || (M > 1 && !V_i_U)) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_sample_u: invalid argument");
    return BVB_ERR_ARG;
  }
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  Scratch* sc = scratch_of(h);
  CholWorkspace& ws = sc->chol;
  cudaStream_t s = h->stream;
  const int nu = M > 1 ? M - 1 : 1, ne = M - 1;
  const size_t tm = (size_t)T * M, nU = (size_t)M * (M - 1) / 2, d = sizeof(double);
  make_pair_geom(ws.gu, T, nu);
  if (!k2_supported(nu)) {
    bvb_set_error(h, BVB_ERR_UNSUPPORTED, "bvb_sample_u: M=%d exceeds the in-SM panel", M);
    return BVB_ERR_UNSUPPORTED;
  }
  const int Nalloc = ((std::max(ne, 1) + 7) / 8) * 8 + K1_NMAX;
  BVB_CUDA_TRY(ws.lutu.ensure(ws.gu.nArows)); BVB_CUDA_TRY(ws.pairsu.ensure(ws.gu.nArows));
  BVB_CUDA_TRY(ws.colbaseu.ensure(ws.gu.colbase.size()));
  BVB_CUDA_TRY(ws.Au.ensure((size_t)ws.gu.nArows * ws.gu.Tp));
  ws.Wu.release(); ws.WYu.release();
  BVB_CUDA_TRY(ws.Wu.ensure((size_t)Nalloc * ws.gu.Tp, true, s));
  BVB_CUDA_TRY(ws.WYu.ensure((size_t)Nalloc * ws.gu.Tp, true, s));
  BVB_CUDA_TRY(ws.omegau.ensure((size_t)std::max(ne, 1) * ws.gu.ldo));
  BVB_CUDA_TRY(ws.Vu.ensure((size_t)std::max(ne, 1) * nu)); BVB_CUDA_TRY(ws.Zu.ensure((size_t)std::max(ne, 1) * nu));
  BVB_CUDA_TRY(ws.phiu.ensure((size_t)std::max(ne, 1) * nu));
  BVB_CUDA_TRY(ws.ViU.ensure(std::max<size_t>(nU, 1))); BVB_CUDA_TRY(ws.zflat.ensure(std::max<size_t>(nU, 1)));
  BVB_CUDA_TRY(ws.resid.ensure(tm)); BVB_CUDA_TRY(ws.dsq.ensure(tm)); BVB_CUDA_TRY(ws.U.ensure((size_t)M * M));
  BVB_CUDA_TRY(ws.status.ensure(M));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.lutu.p, ws.gu.lut.data(), sizeof(int2) * ws.gu.nArows, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.pairsu.p, ws.gu.pairs.data(), sizeof(int2) * ws.gu.nArows, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.colbaseu.p, ws.gu.colbase.data(), sizeof(int) * ws.gu.colbase.size(), cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.resid.p, resid, d * tm, cudaMemcpyHostToDevice, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(ws.dsq.p, d_sqrt, d * tm, cudaMemcpyHostToDevice, s));
  if (nU > 0) {
    BVB_CUDA_TRY(cudaMemcpyAsync(ws.ViU.p, V_i_U, d * nU, cudaMemcpyHostToDevice, s));
    if (z) {
      BVB_CUDA_TRY(cudaMemcpyAsync(ws.zflat.p, z, d * nU, cudaMemcpyHostToDevice, s));
    } else {
      fill_normals_kernel<<<(unsigned)((nU + 255) / 256), 256, 0, s>>>(ws.zflat.p, nU, h->seed, (uint32_t)(0x5c00u + sc->call_counter));
    }
  }
  sc->call_counter++;
  BVB_CUDA_TRY(cudaMemsetAsync(ws.status.p, 0, sizeof(int) * M, s));
  CholUCtx c;
  c.T = T; c.Tp = ws.gu.Tp; c.M = M; c.nu = nu; c.resid = ws.resid.p; c.dsq = ws.dsq.p; c.V_i_U = ws.ViU.p; c.zflat = ws.zflat.p;
  c.U = ws.U.p; c.A = ws.Au.p; c.lut = ws.lutu.p; c.pairs = ws.pairsu.p; c.colbase = ws.colbaseu.p;
  c.nZtiles = ws.gu.nZtiles; c.ntiles = ws.gu.nZtiles + ws.gu.nXtiles; c.nArows = ws.gu.nArows; c.ldo = ws.gu.ldo;
  c.W = ws.Wu.p; c.WY = ws.WYu.p; c.omega = ws.omegau.p; c.Vu = ws.Vu.p; c.Zu = ws.Zu.p; c.phi = ws.phiu.p; c.status = ws.status.p;
  BVB_CUDA_TRY(launch_sample_u(c, s));
  std::vector<int> st(M, 0);
  BVB_CUDA_TRY(cudaMemcpyAsync(U_out, ws.U.p, d * M * M, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaMemcpyAsync(st.data(), ws.status.p, sizeof(int) * M, cudaMemcpyDeviceToHost, s));
  BVB_CUDA_TRY(cudaStreamSynchronize(s));
  for (int e = 0; e < ne; ++e)
    if (st[e] != 0) {
      h->err_eq = e + 2;
      bvb_set_error(h, BVB_ERR_NUMERIC, "univariate_regression_update() failed in equation %i: chol(): decomposition failed", e + 2);
      return BVB_ERR_NUMERIC;
    }
  return BVB_OK;
}

int bvb_gig(bvb_handle* h, int n, const double* lambda, int len_lambda, const double* chi,
            int len_chi, const double* psi, int len_psi, uint64_t stream, double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 2; h->err_msg[0] = 0;
  if (n < 0 || len_lambda <= 0 || len_chi <= 0 || len_psi <= 0 || !lambda || !chi || !psi || (n > 0 && !out)) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_gig: invalid argument");
    return BVB_ERR_ARG;
  }
  const int m = len_lambda > len_chi ? (len_lambda > len_psi ? len_lambda : len_psi) : (len_chi > len_psi ? len_chi : len_psi);
  const size_t tot = (size_t)n * m;
  if (tot == 0) return BVB_OK;
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  DevBuf<double> dl, dc, dp, dout;
  cudaStream_t s = h->stream;
  auto fin = [&](int rc) { dl.release(); dc.release(); dp.release(); dout.release(); return rc; };
  if (dl.ensure(len_lambda) || dc.ensure(len_chi) || dp.ensure(len_psi) || dout.ensure(tot)) {
    bvb_set_error(h, BVB_ERR_CUDA, "bvb_gig: allocation failed");
    return fin(BVB_ERR_CUDA);
  }
  cudaMemcpyAsync(dl.p, lambda, sizeof(double) * len_lambda, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dc.p, chi, sizeof(double) * len_chi, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dp.p, psi, sizeof(double) * len_psi, cudaMemcpyHostToDevice, s);
  cudaEventRecord(h->ev[0], s);
  gig_vec_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, s>>>(dout.p, n, m, dl.p, len_lambda, dc.p, len_chi, dp.p, len_psi, h->seed, stream);
  cudaEventRecord(h->ev[1], s);
  cudaMemcpyAsync(out, dout.p, sizeof(double) * tot, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    h->last_ms[0] = h->last_ms[1] = h->last_ms[2] = 0; h->last_ms[3] = ms;
    h->last_launches = 1;
  }
  if (e != cudaSuccess) {
    bvb_set_error(h, BVB_ERR_CUDA, "bvb_gig: %s", cudaGetErrorString(e));
    return fin(BVB_ERR_CUDA);
  }
  for (size_t i = 0; i < tot; ++i) {
    if (out[i] != out[i]) {  // NaN <=> invalid parameters (the reference raises an R error)
      const size_t j = i / (size_t)n;
      bvb_set_error(h, BVB_ERR_ARG, "invalid parameters for GIG distribution: lambda=%g, chi=%g, psi=%g",
                    lambda[j % len_lambda], chi[j % len_chi], psi[j % len_psi]);
      return fin(BVB_ERR_ARG);
    }
  }
  return fin(BVB_OK);
}

/* Test hook for the parallel state draw of the SV series kernel (see include/bvb.h). */
int bvb_sv_latent_draw(bvb_handle* h, int T, int n_rep, const double* ystar, const unsigned char* mixind, double mu,
                       double phi, double sigma, double h0_var, double* out) {
  if (!h) return BVB_ERR_ARG;
  h->err_code = 0; h->err_step = 2; h->err_msg[0] = 0;
  if (T < 1 || n_rep < 1 || !ystar || !mixind || !out || !(sigma > 0.0) || !(phi > -1.0 && phi < 1.0)) {
    bvb_set_error(h, BVB_ERR_ARG, "bvb_sv_latent_draw: invalid argument");
    return BVB_ERR_ARG;
  }
  for (int t = 0; t < T; ++t) if (mixind[t] > 9) { bvb_set_error(h, BVB_ERR_ARG, "bvb_sv_latent_draw: mixture indicator out of range"); return BVB_ERR_ARG; }
  BVB_CUDA_TRY(cudaSetDevice(h->device));
  DevBuf<double> dy, dout;
  DevBuf<unsigned char> dr;
  cudaStream_t s = h->stream;
  auto fin = [&](int rc) { dy.release(); dr.release(); dout.release(); return rc; };
  if (dy.ensure(T) || dr.ensure(T) || dout.ensure((size_t)n_rep * (T + 1))) {
    bvb_set_error(h, BVB_ERR_CUDA, "bvb_sv_latent_draw: allocation failed");
    return fin(BVB_ERR_CUDA);
  }
  cudaMemcpyAsync(dy.p, ystar, sizeof(double) * T, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dr.p, mixind, (size_t)T, cudaMemcpyHostToDevice, s);
  cudaError_t e = launch_sv_latent_test(T, n_rep, dy.p, dr.p, mu, phi, sigma, h0_var, h->seed, dout.p, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout.p, sizeof(double) * (size_t)n_rep * (T + 1), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    bvb_set_error(h, BVB_ERR_CUDA, "bvb_sv_latent_draw: %s", cudaGetErrorString(e));
    return fin(BVB_ERR_CUDA);
  }
  h->last_launches = 1;
  return fin(BVB_OK);
}

}  // extern "C"
