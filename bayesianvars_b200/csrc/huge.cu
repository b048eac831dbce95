// expert_huge branch of the equation-wise PHI draw (Bhattacharya, Chakraborty &
// Mallick 2016; sample_coefficients.cpp:40-52): when K > T the Kp x Kp system is
// replaced by a T x T one,
//   u = z1 . sqrt(v) + phi0;  delta ~ N(0, I_T);  vv = X_n u + delta
//   w = (X_n V X_n' + I_T)^-1 (y_n - vv);         draw = u + V X_n' w
// with X_n = diag(exp(-h/2)) X.  The same pair-GEMM trick applies with the roles
// of observations and coefficients swapped:
//   (X V_j X')[t, s] = sum_k (X[t,k] X[s,k]) v_j[k]  ==  (Z' V)[pair(t,s), j],
// Z'[pair(t,s), k] built once per chain, the prior variances as the B operand,
// the normaliser applied in the epilogue; K2 then solves the T x T SPD systems
// (z = 0).  The reference calls arma::solve (LU); the matrix is SPD, Cholesky gives
// the same solution to rounding.
#include "huge.cuh"

namespace bvb {

// XT[k + Kp t] = X[t + T k]  (once per chain)
__global__ void huge_transpose_kernel(double* XT, const double* X, int T, int Kp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)T * Kp) return;
  const int k = (int)(idx % Kp), t = (int)(idx / Kp);
  XT[idx] = X[(size_t)k * T + t];
}

// per equation: u = z1 sqrt(v) + phi0, padded copy of v (the GEMM's B operand), normaliser
__global__ void huge_prep_kernel(const HugeCtx c) {
  const int j = blockIdx.x;
  for (int k = threadIdx.x; k < c.Kp; k += blockDim.x) {
    const double v = c.V[(size_t)j * c.Kp + k];
    c.u[(size_t)j * c.Kp + k] = c.z1[(size_t)j * c.Kp + k] * sqrt(v) + c.PHI0[(size_t)j * c.Kp + k];
    c.Wv[(size_t)j * c.Kp_pad + k] = v;
  }
  for (int t = threadIdx.x; t < c.T; t += blockDim.x) c.nrm[(size_t)j * c.T + t] = exp(-0.5 * c.logvar[(size_t)j * c.T + t]);
}

// rhs[t] = nrm (yhat - X u) - delta, written into the rhs row of the packed T x T system
__global__ void __launch_bounds__(256) huge_rhs_kernel(const HugeCtx c) {
  __shared__ double sh[8][33];
  const int j = blockIdx.y;
  const int tx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + tx;
  const double* u = c.u + (size_t)j * c.Kp;
  double s = 0.0;
  if (t < c.T) for (int k = sl; k < c.Kp; k += 8) s = fma(c.X[(size_t)k * c.T + t], u[k], s);
  sh[sl][tx] = s;
  __syncthreads();
  if (sl == 0 && t < c.T) {
    double xu = 0.0;
#pragma unroll
    for (int z = 0; z < 8; ++z) xu += sh[z][tx];
    double yh = c.Y[(size_t)j * c.T + t];
    for (int q = 0; q < c.r; ++q) yh -= c.facload[(size_t)q * c.M + j] * c.fac[(size_t)t * c.r + q];
    const double nr = c.nrm[(size_t)j * c.T + t];
    c.omega[(size_t)j * c.ldo + c.colbase_h[t] + c.T] = nr * (yh - xu) - c.delta[(size_t)j * c.T + t];
  }
}

// draw[a] = u[a] + v[a] * sum_t X[t,a] nrm[t] w[t]   (one warp per coefficient)
__global__ void __launch_bounds__(256) huge_final_kernel(const HugeCtx c) {
  const int j = blockIdx.y, lane = threadIdx.x & 31;
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= c.Kp) return;
  const double* x = c.X + (size_t)a * c.T;
  const double* nr = c.nrm + (size_t)j * c.T;
  const double* w = c.wsol + (size_t)j * c.T;
  double s = 0.0;
  for (int t = lane; t < c.T; t += 32) s = fma(x[t] * nr[t], w[t], s);
  s = warp_sum(s);
  if (lane == 0) c.PHI[(size_t)j * c.Kp + a] = c.u[(size_t)j * c.Kp + a] + c.V[(size_t)j * c.Kp + a] * s;
}

cudaError_t launch_huge_setup(double* XT, double* A, const double* X, const int2* pairs, int T, int Kp, int Kp_pad,
                              int nArows, cudaStream_t s) {
  const size_t n = (size_t)T * Kp;
  huge_transpose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(XT, X, T, Kp);
  // pair-product matrix over observation pairs: "X" is XT (Kp x T), contraction length Kp
  return launch_build_A(A, XT, pairs, Kp, Kp_pad, T, nArows, 0, s);
}

cudaError_t launch_phi_huge(const HugeCtx& c, cudaStream_t s) {
  huge_prep_kernel<<<c.M, 256, 0, s>>>(c);
  K1Params k1{};
  k1.A = c.A; k1.W = c.Wv; k1.WY = c.WYzero; k1.lut = c.lut_h; k1.V = c.V; k1.PHI0 = nullptr; k1.omega = c.omega;
  k1.Tp = c.Kp_pad; k1.nZtiles = c.nZtiles; k1.ntiles = c.ntiles; k1.M = c.M; k1.n = c.T; k1.ldo = c.ldo;
  k1.pairs = c.pairs_h; k1.rowscale = c.nrm;
  cudaError_t e = launch_k1(k1, s);
  if (e != cudaSuccess) return e;
  dim3 g1((c.T + 31) / 32, c.M);
  huge_rhs_kernel<<<g1, 256, 0, s>>>(c);
  K2Params k2{};
  k2.omega = c.omega; k2.colbase = c.colbase_h; k2.z = c.zero_tm; k2.phi = c.wsol; k2.status = c.status;
  k2.n = c.T; k2.M = c.M; k2.ldo = c.ldo; k2.ldz = c.T; k2.ldphi = c.T;
  e = launch_k2(k2, s);
  if (e != cudaSuccess) return e;
  dim3 g2((c.Kp + 7) / 8, c.M);
  huge_final_kernel<<<g2, 256, 0, s>>>(c);
  return cudaGetLastError();
}

}  // namespace bvb
