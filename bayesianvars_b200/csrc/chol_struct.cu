// Cholesky error structure: corrected triangular algorithm for PHI (Carriero,
// Chan, Clark & Marcellino 2021; sample_coefficients.cpp:118-146) and sample_U
// (sample_coefficients.cpp:160-188).
//
// The reference materialises, per equation j, the T(M-j) x Kp Kronecker design
// kron(U[j, j:]', X) / d and runs dsyrk on it (~400-800 GFLOP per sweep at
// M=100, T=500).  Algebraically (checked in
// tests/test_oracle_regression.py) the normal equations collapse to
//   Omega_j = X' diag(w_j) X + diag(1/v_j),  w_j[t] = sum_{i>=j} U[j,i]^2 / d[t,i]^2
//   b_j     = X' (q_j + w_j . (X phi_j_old)) + phi0_j / v_j,
//   q_j[t]  = sum_{i>=j} U[j,i] Q[t,i] / d[t,i]^2,   Q = (Y - X PHI) U
// and neither w_j nor Omega_j depends on PHI.  So the dense work (K1 pair-GEMM
// with weights w_j, batched Cholesky) is done for ALL equations at once, exactly
// like the factor structure, and only the right-hand sides and the triangular
// solves follow the sequential Gauss-Seidel order: per equation one GEMV for b_j,
// one solve-only CTA, one GEMV for X phi_j_new and a rank-one update of Q.
#include "chol.cuh"

namespace bvb {

__global__ void chol_d2inv_kernel(double* d2inv, const double* dsq, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double d = dsq[i]; d2inv[i] = 1.0 / (d * d); }
}

// W[j][t] = sum_{i >= j} U[j,i]^2 / d[t,i]^2
__global__ void chol_w_kernel(double* W, const double* U, const double* d2inv, int T, int Tp, int M) {
  const int j = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double w = 0.0;
    for (int i = j; i < M; ++i) { const double u = U[(size_t)i * M + j]; w = fma(u * u, d2inv[(size_t)i * T + t], w); }
    W[(size_t)j * Tp + t] = w;
  }
}

// XPHI = Y - R;  Q[t,i] = sum_{m <= i} R[t,m] U[m,i]
__global__ void chol_q_kernel(double* Q, double* XPHI, const double* R, const double* Y, const double* U, int T, int M) {
  const int i = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double q = 0.0;
    for (int m = 0; m <= i; ++m) q = fma(R[(size_t)m * T + t], U[(size_t)i * M + m], q);
    Q[(size_t)i * T + t] = q;
    XPHI[(size_t)i * T + t] = Y[(size_t)i * T + t] - R[(size_t)i * T + t];
  }
}

// r_j[t] = sum_{i>=j} U[j,i] Q[t,i]/d[t,i]^2 + w_j[t] * (X phi_j_old)[t]
__global__ void chol_r_kernel(double* rvec, const double* Q, const double* XPHI, const double* U, const double* d2inv,
                              const double* W, int T, int Tp, int M, int j) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double q = 0.0;
  for (int i = j; i < M; ++i) q = fma(U[(size_t)i * M + j] * d2inv[(size_t)i * T + t], Q[(size_t)i * T + t], q);
  rvec[t] = q + W[(size_t)j * Tp + t] * XPHI[(size_t)j * T + t];
}

// b[a] = X[:,a]' r + phi0[a,j]/v[a,j]   (one warp per coefficient)
__global__ void __launch_bounds__(256) chol_b_kernel(double* bvec, const double* X, const double* rvec, const double* PHI0,
                                                     const double* V, int T, int Kp, int j) {
  const int lane = threadIdx.x & 31;
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= Kp) return;
  const double* x = X + (size_t)a * T;
  double s = 0.0;
  for (int t = lane; t < T; t += 32) s = fma(x[t], rvec[t], s);
  s = warp_sum(s);
  if (lane == 0) bvec[a] = s + PHI0[(size_t)j * Kp + a] / V[(size_t)j * Kp + a];
}

// After phi_j is drawn: u = X phi_j_new - X phi_j_old; XPHI[:,j] = X phi_j_new;
// Q[:, i] -= u U[j,i] (i >= j); and the next equation's r_{j+1}.
// Block = 32 time points x 32 slices (of the Kp coefficients, then of the M series): the kernel sits on the sequential
// critical path M times per sweep and is pure load latency, so it runs few iterations on many threads.
constexpr int UPD_SL = 32;
__global__ void __launch_bounds__(32 * UPD_SL) chol_upd_kernel(double* Q, double* XPHI, double* rvec, const double* X,
                                                               const double* PHI, const double* U, const double* d2inv,
                                                               const double* W, int T, int Tp, int Kp, int M, int j) {
  __shared__ double sh[UPD_SL][33];
  const int tx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + tx;
  const double* phi = PHI + (size_t)j * Kp;
  double s = 0.0;
  if (t < T) for (int k = sl; k < Kp; k += UPD_SL) s = fma(X[(size_t)k * T + t], phi[k], s);
  sh[sl][tx] = s;
  __syncthreads();
  double xnew = 0.0;
#pragma unroll
  for (int z = 0; z < UPD_SL; ++z) xnew += sh[z][tx];
  __syncthreads();
  double u = 0.0;
  if (t < T) u = xnew - XPHI[(size_t)j * T + t];
  // Q update for i >= j (slices split i) and partial q_{j+1}
  double q = 0.0;
  if (t < T) {
    for (int i = j + sl; i < M; i += UPD_SL) {
      const double qv = Q[(size_t)i * T + t] - u * U[(size_t)i * M + j];
      Q[(size_t)i * T + t] = qv;
      if (i >= j + 1 && j + 1 < M) q = fma(U[(size_t)i * M + (j + 1)] * d2inv[(size_t)i * T + t], qv, q);
    }
  }
  sh[sl][tx] = q;
  __syncthreads();
  if (sl == 0 && t < T) {
    XPHI[(size_t)j * T + t] = xnew;
    if (j + 1 < M) {
      double qs = 0.0;
#pragma unroll
      for (int z = 0; z < UPD_SL; ++z) qs += sh[z][tx];
      rvec[t] = qs + W[(size_t)(j + 1) * Tp + t] * XPHI[(size_t)(j + 1) * T + t];
    }
  }
}

cudaError_t launch_chol_phi(const CholPhiCtx& c, cudaStream_t s) {
  const int T = c.T, M = c.M, Kp = c.Kp;
  const size_t tm = (size_t)T * M;
  chol_d2inv_kernel<<<(unsigned)((tm + 255) / 256), 256, 0, s>>>(c.d2inv, c.dsq, tm);
  chol_w_kernel<<<M, 256, 0, s>>>(c.W, c.U, c.d2inv, T, c.Tp, M);
  K1Params k1{};
  k1.A = c.A; k1.W = c.W; k1.WY = c.WYzero; k1.lut = c.lut; k1.V = c.V; k1.PHI0 = c.PHI0; k1.omega = c.omega;
  k1.Tp = c.Tp; k1.nZtiles = c.nZtiles; k1.ntiles = c.ntiles; k1.M = M; k1.n = Kp; k1.ldo = c.ldo;
  cudaError_t e = launch_k1(k1, s);
  if (e != cudaSuccess) return e;
  K2Params k2{};
  k2.omega = c.omega; k2.colbase = c.colbase; k2.z = c.Zn; k2.phi = c.PHI; k2.status = c.status;
  k2.n = Kp; k2.M = M; k2.ldo = c.ldo; k2.ldz = Kp; k2.ldphi = Kp; k2.mode = 1;
  e = launch_k2(k2, s);
  if (e != cudaSuccess) return e;
  e = launch_resid(c.resid, c.Y, c.X, c.PHI, T, Kp, M, s);
  if (e != cudaSuccess) return e;
  chol_q_kernel<<<M, 256, 0, s>>>(c.Q, c.XPHI, c.resid, c.Y, c.U, T, M);
  chol_r_kernel<<<(T + 127) / 128, 128, 0, s>>>(c.rvec, c.Q, c.XPHI, c.U, c.d2inv, c.W, T, c.Tp, M, 0);
  for (int j = 0; j < M; ++j) {
    chol_b_kernel<<<(Kp + 7) / 8, 256, 0, s>>>(c.bvec, c.X, c.rvec, c.PHI0, c.V, T, Kp, j);
    K2Params ks = k2;
    ks.z = c.Zn + (size_t)j * Kp;
    ks.phi = c.PHI + (size_t)j * Kp;
    e = launch_k2_solve(ks, c.bvec, j, 1, s);
    if (e != cudaSuccess) return e;
    chol_upd_kernel<<<(T + 31) / 32, 32 * UPD_SL, 0, s>>>(c.Q, c.XPHI, c.rvec, c.X, c.PHI, c.U, c.d2inv, c.W, T, c.Tp, Kp, M, j);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ sample_U
// Equation e = i - 1 (column i of U, i = 1..M-1) regresses resid[:,i]/d_i on
// -resid[:,0:i]/d_i.  All M-1 systems are embedded in nu = M-1 regressors
// (identity padding), so they share one pair-GEMM over X = resid[:, 0:nu] with
// weights 1/d_i^2 and one batched Cholesky/solve.
__global__ void u_prep_kernel(double* W, double* WY, double* Vu, double* Zu, const double* resid, const double* dsq,
                              const double* V_i_U, const double* zflat, int T, int Tp, int M, int nu) {
  const int e = blockIdx.x, i = e + 1;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double d = dsq[(size_t)i * T + t];
    const double w = 1.0 / (d * d);
    W[(size_t)e * Tp + t] = w;
    WY[(size_t)e * Tp + t] = -w * resid[(size_t)i * T + t];   // regressors carry the minus sign (:169-170)
  }
  const size_t ind = (size_t)i * (i - 1) / 2;
  for (int a = threadIdx.x; a < nu; a += blockDim.x) {
    Vu[(size_t)e * nu + a] = (a < i) ? V_i_U[ind + a] : 1.0;
    Zu[(size_t)e * nu + a] = (a < i) ? zflat[ind + a] : 0.0;
  }
}

__global__ void u_mask_kernel(double* omega, const int* colbase, long ldo, int nu) {
  const int e = blockIdx.x, i = e + 1;
  double* om = omega + (size_t)e * ldo;
  // entries (a, c), c <= a <= nu (row nu = rhs): real iff a < i (rhs: c < i)
  for (int c = 0; c < nu; ++c) {
    for (int a = c + threadIdx.x; a <= nu; a += blockDim.x) {
      const bool real = (a < nu) ? (a < i) : (c < i);
      if (!real) om[colbase[c] + a] = (a == c) ? 1.0 : 0.0;
    }
  }
}

__global__ void u_scatter_kernel(double* U, const double* phi, int M, int nu) {
  const int i = blockIdx.x;   // column of U
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    double v = (a == i) ? 1.0 : 0.0;
    if (a < i) v = phi[(size_t)(i - 1) * nu + a];
    U[(size_t)i * M + a] = v;
  }
}

cudaError_t launch_sample_u(const CholUCtx& c, cudaStream_t s) {
  const int M = c.M, nu = c.nu, ne = M - 1;
  if (ne <= 0) {
    u_scatter_kernel<<<M, 32, 0, s>>>(c.U, c.phi, M, 1);
    return cudaGetLastError();
  }
  u_prep_kernel<<<ne, 256, 0, s>>>(c.W, c.WY, c.Vu, c.Zu, c.resid, c.dsq, c.V_i_U, c.zflat, c.T, c.Tp, M, nu);
  cudaError_t e = launch_build_A(c.A, c.resid, c.pairs, c.T, c.Tp, nu, c.nArows, 0, s);
  if (e != cudaSuccess) return e;
  K1Params k1{};
  k1.A = c.A; k1.W = c.W; k1.WY = c.WY; k1.lut = c.lut; k1.V = c.Vu; k1.PHI0 = nullptr; k1.omega = c.omega;
  k1.Tp = c.Tp; k1.nZtiles = c.nZtiles; k1.ntiles = c.ntiles; k1.M = ne; k1.n = nu; k1.ldo = c.ldo;
  e = launch_k1(k1, s);
  if (e != cudaSuccess) return e;
  u_mask_kernel<<<ne, 128, 0, s>>>(c.omega, c.colbase, c.ldo, nu);
  K2Params k2{};
  k2.omega = c.omega; k2.colbase = c.colbase; k2.z = c.Zu; k2.phi = c.phi; k2.status = c.status;
  k2.n = nu; k2.M = ne; k2.ldo = c.ldo; k2.ldz = nu; k2.ldphi = nu;
  e = launch_k2(k2, s);
  if (e != cudaSuccess) return e;
  u_scatter_kernel<<<M, 128, 0, s>>>(c.U, c.phi, M, nu);
  return cudaGetLastError();
}

}  // namespace bvb
