// Drives r_shim/bvar_cpp_shim.cpp - compiled against the Rcpp stand-in of this directory - the way R's generated
// RcppExports wrappers would: reads the argument lists of bvar_cpp() from a small binary file written by
// tests/test_shim_cpp.py, calls bvar_cpp(), then irf_cpp(), out_of_sample(), my_gig(), vcov_cpp(), predh() on its
// result, and writes everything back.  This is the C++ (not Python / ctypes) test of the C ABI: the struct layouts,
// argument order and ownership rules are exactly the ones the shim uses inside R.
//
// File format (little endian):  obj := kind:u8 ...
//   0 NIL | 1 REAL ndim:i32 dims:i32[ndim] data:f64[] | 2 INT ... i32[] | 3 LGL ... i32[] | 4 STR len:i32 bytes
//   5 LIST n:i32 (name:STR-payload obj)*n
#include "../../r_shim/bvar_cpp_shim.cpp"

#include <cstring>
#include <fstream>
#include <iostream>

using Rcpp::Ptr;
using Rcpp::RObj;

static int32_t rd_i32(std::istream& f) { int32_t v; f.read((char*)&v, 4); return v; }
static std::string rd_str(std::istream& f) { int32_t n = rd_i32(f); std::string s((size_t)n, '\0'); f.read(&s[0], n); return s; }

static Ptr rd_obj(std::istream& f) {
  uint8_t k; f.read((char*)&k, 1);
  if (k == 0) return Rcpp::mk(RObj::NIL);
  if (k == 4) { auto o = Rcpp::mk(RObj::STR); o->s.push_back(rd_str(f)); return o; }
  if (k == 5) {
    auto o = Rcpp::mk(RObj::LIST);
    const int n = rd_i32(f);
    for (int i = 0; i < n; ++i) { o->names.push_back(rd_str(f)); o->l.push_back(rd_obj(f)); }
    return o;
  }
  auto o = Rcpp::mk(k == 1 ? RObj::REAL : (k == 2 ? RObj::INT : RObj::LGL));
  const int nd = rd_i32(f);
  std::vector<int> dims((size_t)nd);
  size_t n = 1;
  for (auto& d : dims) { d = rd_i32(f); n *= (size_t)d; }
  if (k == 1) { o->d.resize(n); f.read((char*)o->d.data(), 8 * n); }
  else { o->i.resize(n); f.read((char*)o->i.data(), 4 * n); }
  if (nd >= 2) { auto dm = Rcpp::mk(RObj::INT); dm->i = dims; o->attrs["dim"] = dm; }
  return o;
}

static void wr_i32(std::ostream& f, int32_t v) { f.write((char*)&v, 4); }
static void wr_str(std::ostream& f, const std::string& s) { wr_i32(f, (int32_t)s.size()); f.write(s.data(), s.size()); }
static void wr_obj(std::ostream& f, const Ptr& o) {
  if (!o || o->kind == RObj::NIL) { f.put(0); return; }
  if (o->kind == RObj::LIST) {
    f.put(5); wr_i32(f, (int32_t)o->l.size());
    for (size_t i = 0; i < o->l.size(); ++i) { wr_str(f, i < o->names.size() ? o->names[i] : std::to_string(i)); wr_obj(f, o->l[i]); }
    return;
  }
  if (o->kind == RObj::STR) { f.put(4); wr_str(f, o->s.empty() ? "" : o->s[0]); return; }
  const bool real = o->kind == RObj::REAL;
  f.put(real ? 1 : (o->kind == RObj::INT ? 2 : 3));
  std::vector<int> dims;
  auto it = o->attrs.find("dim");
  const size_t n = real ? o->d.size() : o->i.size();
  if (it != o->attrs.end()) dims = it->second->i; else dims = {(int)n};
  wr_i32(f, (int32_t)dims.size());
  for (int d : dims) wr_i32(f, d);
  if (real) f.write((const char*)o->d.data(), 8 * n); else f.write((const char*)o->i.data(), 4 * n);
}

int main(int argc, char** argv) {
  if (argc < 3) { std::cerr << "usage: shim_driver in.bin out.bin\n"; return 2; }
  std::ifstream fin(argv[1], std::ios::binary);
  if (!fin) { std::cerr << "cannot open " << argv[1] << "\n"; return 2; }
  Rcpp::List in(rd_obj(fin));
  try {
    // the Philox key the shim will draw from "R's RNG" (recorded so that the Python side can run the same chain)
    R::stub_set_seed(537);
    const double u1 = R::unif_rand(), u2 = R::unif_rand();
    const uint64_t seed = ((uint64_t)(u1 * 4294967296.0) << 32) | (uint64_t)(u2 * 4294967296.0);
    R::stub_set_seed(537);

    NumericMatrix Y = in["Y"], X = in["X"], PHI0 = in["PHI0"];
    IntegerMatrix i_mat = in["i_mat"];
    IntegerVector i_vec = in["i_vec"];
    NumericVector priorIntercept = in["priorIntercept"];
    const int M = in["M"], T = in["T"], K = in["K"], draws = in["draws"], burnin = in["burnin"], thin = in["thin"];
    const int intercept = in["intercept"];
    const std::string tvp_keep = in["tvp_keep"];
    const double PHI_tol = in["PHI_tol"], L_tol = in["L_tol"];
    const bool huge = in["huge"];
    Rcpp::List res = bvar_cpp(Y, X, M, T, K, draws, burnin, thin, tvp_keep, intercept, priorIntercept, PHI0,
                              Rcpp::List(in["priorPHI_in"]), Rcpp::List(in["priorSigma_in"]), Rcpp::List(in["Rstartvals_in"]),
                              i_mat, i_vec, false, PHI_tol, L_tol, huge);
    Rcpp::List out;
    out.push(Rcpp::Named("bvar") = res);
    {
      auto s = Rcpp::mk(RObj::REAL);
      s->d = {(double)(seed >> 32), (double)(seed & 0xffffffffu)};
      Rcpp::NumericVector sv(s);
      out.push(Rcpp::Named("seed_hi_lo") = sv);
    }
    const std::string type = Rcpp::as<std::string>(Rcpp::List(in["priorSigma_in"])["type"]);
    const bool factor = type == "factor";
    NumericVector PHI = res["PHI"], facload = res["facload"], logvar = res["logvar"], sv_para = res["sv_para"];
    NumericMatrix U = res["U"];
    const IntegerVector dl = logvar.attr("dim");
    const int Tl = dl[0], S = dl[1], nsave = dl[2];
    // logvar_t = x$logvar[t,,]  (last stored time point)
    NumericMatrix lv_T(S, nsave);
    for (int d = 0; d < nsave; ++d) for (int s = 0; s < S; ++s) lv_T(s, d) = logvar[(size_t)d * Tl * S + (size_t)s * Tl + (Tl - 1)];
    // irf: unit shocks (one per factor | the first two variables)
    const int r = factor ? (int)(Rcpp::IntegerVector(facload.attr("dim"))[1]) : 0;
    const int dim = factor ? r : M, n_shocks = factor ? r : 2;
    NumericMatrix shocks(dim, n_shocks);
    for (int k = 0; k < n_shocks; ++k) shocks(k, k) = 1.0;
    Rcpp::List irf = irf_cpp(PHI, factor ? facload : NumericVector(), factor ? NumericMatrix() : U, lv_T, shocks, 5);
    out.push(Rcpp::Named("irf") = irf);
    // out_of_sample: ahead = 1:2, every series follows an SV process
    NumericMatrix sv_mu(S, nsave), sv_phi(S, nsave), sv_sigma(S, nsave);
    for (int d = 0; d < nsave; ++d) for (int s = 0; s < S; ++s) {
      sv_mu(s, d) = sv_para[(size_t)d * 3 * S + 3 * s]; sv_phi(s, d) = sv_para[(size_t)d * 3 * S + 3 * s + 1];
      sv_sigma(s, d) = sv_para[(size_t)d * 3 * S + 3 * s + 2];
    }
    IntegerVector ahead = IntegerVector::create(1, 2), svi(S), VoI = IntegerVector::create(0);
    for (int s = 0; s < S; ++s) svi[s] = s;
    const int Kp = K + intercept;
    NumericVector x_next(Kp);
    for (int k = 0; k < Kp; ++k) x_next[k] = X(T - 1, k);
    NumericMatrix Y_obs(2, M);
    for (int a = 0; a < 2; ++a) for (int m = 0; m < M; ++m) Y_obs(a, m) = Y(T - 2 + a, m);
    Rcpp::List pred = out_of_sample(2, x_next, PHI, U, facload, lv_T, ahead, sv_mu, sv_phi, sv_sigma, svi, factor, true, Y_obs,
                                    true, VoI, true);
    out.push(Rcpp::Named("pred") = pred);
    NumericVector lam = NumericVector::create(0.5, -0.5, 2.0), chi = NumericVector::create(1.0), psi = NumericVector::create(1.0, 3.0);
    out.push(Rcpp::Named("gig") = my_gig(1000, lam, chi, psi));
    out.push(Rcpp::Named("vcov") = vcov_cpp(factor, facload, logvar, U, M, r));
    out.push(Rcpp::Named("predh") = predh(lv_T, ahead, 1, sv_mu, sv_phi, sv_sigma));
    if (!factor) {
      // sample_PHI_cholesky with the last draw as the current value
      NumericMatrix PHI_cur(Kp, M), Vp(Kp, M), dsq(T, M), Umat(M, M);
      NumericVector Vprior = res["V_prior"];
      for (int i = 0; i < Kp * M; ++i) { PHI_cur[i] = PHI[(size_t)(nsave - 1) * Kp * M + i]; Vp[i] = Vprior[(size_t)(nsave - 1) * Kp * M + i]; }
      for (int m = 0; m < M; ++m) for (int t = 0; t < T; ++t) dsq(t, m) = std::exp(0.5 * lv_T(m, nsave - 1));
      for (int m = 0; m < M; ++m) Umat(m, m) = 1.0;
      int q = 0;
      for (int j = 1; j < M; ++j) for (int i = 0; i < j; ++i) Umat(i, j) = U(q++, nsave - 1);
      out.push(Rcpp::Named("phi_chol") = sample_PHI_cholesky(PHI_cur, PHI0, Y, X, Umat, dsq, Vp));
    }
    std::ofstream fout(argv[2], std::ios::binary);
    wr_obj(fout, out.p);
  } catch (const std::exception& e) {
    std::cerr << "R error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
