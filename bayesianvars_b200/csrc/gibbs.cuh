// Sampler state of one Gibbs chain (device buffers + host bookkeeping).
#pragma once
#include <vector>
#include "bvb_internal.cuh"
#include "hyper.cuh"
#include "fsv.cuh"
#include "chol.cuh"
#include "huge.cuh"

namespace bvb {

// Stream on which GBuf clears fresh allocations: it must be the stream the
// uploads / kernels run on (a memset on the legacy default stream is NOT ordered
// with a cudaStreamNonBlocking stream and could land after an upload).
inline cudaStream_t& gbuf_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}

template <typename T>
struct GBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t count, bool zero = false) {
    if (count == 0) count = 1;
    if (count > n) {
      if (p) cudaFree(p);
      p = nullptr; n = 0;
      cudaError_t e = cudaMalloc(&p, count * sizeof(T));
      if (e != cudaSuccess) return e;
      n = count;
      zero = true;   // fresh allocations are always cleared (padding is relied upon)
    }
    if (zero) return cudaMemsetAsync(p, 0, n * sizeof(T), gbuf_stream());
    return cudaSuccess;
  }
  ~GBuf() { if (p) cudaFree(p); }
};

struct GibbsHyper {
  HyperDev d;
  GBuf<int> grp, gstart, gmem, gcnt;
  GBuf<double> coef, V_i, loc1, loc2, gpar, red, gsum, gpart, a_vec, a_weight, norm_consts, c_vec, tau0, tau1, V_prep;
};

struct GibbsSeg {
  const double* src;
  int rows, cols, src_ld, src_row0;
  size_t dst_off;
};
struct GibbsStore {
  GibbsSeg seg[16];
  int nseg = 0;
  double* ring = nullptr;
  size_t slot_doubles = 0;
  int ring_slots = 0;
  const int* first_slot = nullptr;   // device: index of the stored draw that maps to ring slot 0
  int burnin = 0, thin = 1;
};

struct Gibbs {
  bool ready = false;
  int M = 0, T = 0, K = 0, Kp = 0, intercept = 0, r = 0, S = 0;
  int draws = 0, burnin = 0, thin = 1, nsave = 0, tvp_all = 0, rep = 0;
  double PHI_tol = 0, L_tol = 0;
  uint64_t seed = 0;
  int rank = 0, world = 1, j0 = 0, nloc = 0, nloc_max = 0;
  void* comm = nullptr;
  PairGeom geom;
  GBuf<double> X, Y, A, W, WY, omega, Vprior, PHI0, PHI, Zn, resid;
  std::vector<double> Xhost;   // the regressors the pair-product matrix A was built from (a recycled chain skips the rebuild)
  GBuf<int2> lut, pairs;
  GBuf<int> colbase, status, eqstatus, nstored, first_slot;
  GBuf<uint32_t> sweep;
  GBuf<unsigned long long> ktime;   // live kernel timers: [0..1] K1 span, [2..3] K2 span, [4] K1 ns, [5] K2 ns, [6] sweeps
  GibbsHyper hphi;
  // Sigma_t block
  FsvDev fsv;
  GBuf<double> logvar, logvar0, sv_para, facload, fac, tau2, lambda2, ystar, eps, svscratch, eidi, wexp;
  GBuf<double> offset, priorsigma2, priorh0, priorhomo, shrink_a, shrink_c, shrink_d;
  GBuf<unsigned char> mixind;
  GBuf<int> hetero, restr;
  // cholesky structure
  int sigma_type = BVB_SIGMA_FACTOR;
  int nU = 0;
  PairGeom geomU;
  GibbsHyper hU;
  GBuf<double> U, dsq, d2inv, XPHI, Q, rvec, bvec, WYzero, strres, zflat, uvec;
  GBuf<double> Au, Wu, WYu, omegau, Vu, Zu, phiu;
  GBuf<int2> lutu, pairsu;
  GBuf<int> colbaseu, ustatus;
  CholPhiCtx cphi;
  CholUCtx cu;
  size_t off_U = 0, off_uhyper = 0;
  int uhyper_rows = 0;
  // expert_huge branch (factor structure, T x T systems)
  int huge = 0;
  PairGeom geomH;
  HugeCtx chuge;
  GBuf<double> XT, Ah, Wv, WYzh, omegah, hu, hnrm, hwsol, hzero, hdelta;
  GBuf<int2> luth, pairsh;
  GBuf<int> colbaseh;
  // stored draws
  GibbsStore store;
  GBuf<double> ring;
  double* pinned[2] = {nullptr, nullptr};
  cudaStream_t side = nullptr;                 // hyper-parameter branch of the sweep graph
  cudaStream_t copy = nullptr;                 // D2H of the stored-draw ring (its two halves alternate between sub-chunks)
  static constexpr int FS_SLOTS = 65;          // pinned staging words of `first_slot` ([0]: eager sweeps, rest: sub-chunks in flight)
  int* first_slot_host = nullptr;
  cudaEvent_t ev_sub[2] = {nullptr, nullptr};  // sweeps of the sub-chunk that filled ring half b are done
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ng = nullptr, ev_series = nullptr, ev_k1z = nullptr;
  bool k1z_valid = false;                      // the Z tiles of the NEXT PHI draw are already in `omega` (see enqueue_sweep)
  cudaEvent_t ev_pin[2] = {nullptr, nullptr};
  bvb_draws out{};
  size_t off_PHI = 0, off_Vprior = 0, off_logvar = 0, off_svpara = 0, off_hyper = 0, off_facload = 0, off_fac = 0;
  int hyper_rows = 0;
  // execution
  cudaGraphExec_t graph_exec = nullptr;
  // peer exchange of the PHI columns (push + flag over NVLink peer memory instead of an NCCL allgather)
  GBuf<double> xbuf;                     // [2 chains][2 sweeps][world][nloc_max * Kp]: what the peers pushed (parity of the chain count, of the sweep)
  GBuf<unsigned long long> xflag;        // [world]: (epoch << 32 | sweep + 1) of the last block pushed by each peer
  GBuf<double*> peer_xbuf;               // device tables of the peers' xbuf / xflag
  GBuf<unsigned long long*> peer_xflag;
  GBuf<int> xcnt;                        // [world]: slices of the current push that are out (self-resetting)
  std::vector<void*> ipc_opened;
  bool p2p = false, p2p_wanted = false;
  unsigned long long epoch_base = 0;
  char* stage = nullptr;                 // pinned staging arena of the set-up uploads
  size_t stage_cap = 0, stage_off = 0;
  bool graph_stale = false;   // chain recycled: re-capture and cudaGraphExecUpdate before the next replay
  bool timing_pass = false, timed = false;
  cudaEvent_t ev[9] = {};
  cudaEvent_t ev_chunk[2] = {nullptr, nullptr};
  float phase_ms[8] = {0};
  float last_chunk_ms = 0;
  int last_chunk_sweeps = 0, launches_per_sweep = 0;
};

}  // namespace bvb

void bvb_gibbs_release(bvb_handle* h);
