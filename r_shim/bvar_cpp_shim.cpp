// Rcpp shim: keeps the reference's .Call symbols and R-visible signatures
//   _bayesianVARs_bvar_cpp            src/RcppExports.cpp:16-44
//   _bayesianVARs_my_gig              src/RcppExports.cpp:46-58
//   _bayesianVARs_irf_cpp             src/RcppExports.cpp:148-163   (declared at src/irf.cpp:467-476)
//   _bayesianVARs_out_of_sample       src/RcppExports.cpp:165-191   (src/utilities_cpp.cpp:216-232)
//   _bayesianVARs_sample_PHI_cholesky src/RcppExports.cpp:224-236
//   _bayesianVARs_predh / _vcov_cpp   src/utilities_cpp.cpp:181-212, 436-472
// and forwards to the C ABI of libbvb.so (include/bvb.h).  It replaces src/bvar_cpp.cpp, src/gig_vec.cpp and the
// named functions of src/irf.cpp / src/utilities_cpp.cpp / src/sample_coefficients.cpp in the package; R/ stays
// untouched (Rcpp::compileAttributes regenerates RcppExports from the [[Rcpp::export]] tags below).  PKG_LIBS gains -lbvb.
//
// No R / Rcpp in the build container: this file is compiled and RUN there against tests/r_stub/Rcpp.h, a small
// stand-in for the Rcpp classes used here (tests/test_shim_cpp.py drives bvar_cpp(), irf_cpp(), out_of_sample(), my_gig()
// through it on the GPU).  Only Rcpp API that exists with the same spelling in real Rcpp is used.
//
// One process-wide handle (one CUDA context, cached workspaces) is created on first use; every entry point re-keys
// its Philox generator from R's RNG, so set.seed() keeps runs reproducible and successive calls get fresh streams.
#include <Rcpp.h>
#include <algorithm>
#include <string>
#include <vector>
#include "bvb.h"

using namespace Rcpp;

namespace {

uint64_t g_call_seed = 0;   // the Philox key of the current entry point (drawn from R's RNG in the_handle())

bvb_handle* the_handle() {
  static bvb_handle* h = nullptr;
  if (!h) {
    int dev = 0;
    if (const char* e = std::getenv("BVB_DEVICE")) dev = std::atoi(e);
    if (bvb_create(&h, dev, 0) != BVB_OK) {
      h = nullptr;
      stop("bayesianVARs: no usable sm_100a CUDA device (there is no CPU fallback)");
    }
  }
  // the Philox key is drawn from R's RNG at every entry point (RNGScope is opened by the generated wrapper)
  const double u1 = R::unif_rand(), u2 = R::unif_rand();
  const uint64_t seed = ((uint64_t)(u1 * 4294967296.0) << 32) | (uint64_t)(u2 * 4294967296.0);
  bvb_set_seed(h, seed);
  g_call_seed = seed;
  return h;
}

// BVB_DEVICES="0,1,2,3": bvar() with the factor structure shards its equations over these GPUs from this one R process
// (bvb_create_multi: one handle per device, an in-process NCCL communicator); unset or one device = the plain path.
struct MultiGpu { int n = 0; std::vector<bvb_handle*> h; };
MultiGpu& the_gpus() {
  static MultiGpu g;
  static bool tried = false;
  if (!tried) {
    tried = true;
    std::vector<int> devs;
    if (const char* e = std::getenv("BVB_DEVICES")) {
      std::string s(e);
      size_t pos = 0;
      while (pos < s.size()) {
        const size_t c = s.find(',', pos);
        const std::string tok = s.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
        if (!tok.empty()) devs.push_back(std::atoi(tok.c_str()));
        if (c == std::string::npos) break;
        pos = c + 1;
      }
    }
    if (devs.size() > 1) {
      g.h.assign(devs.size(), nullptr);
      if (bvb_create_multi((int)devs.size(), devs.data(), 0, g.h.data()) == BVB_OK) g.n = (int)devs.size();
      else { g.h.clear(); Rcpp::warning("bayesianVARs: BVB_DEVICES could not be opened together, using one GPU"); }
    }
  }
  return g;
}

void check(bvb_handle* h, int rc) {
  if (rc == BVB_OK) return;
  int step, sweep, eq;
  char msg[512];
  bvb_last_error(h, &step, &sweep, &eq, msg, sizeof msg);
  stop("%s", msg);  // messages already follow bvar_cpp.cpp:512,522,530 / sample_coefficients.cpp:106
}

int prior_code(const std::string& p) {
  if (p == "normal") return BVB_PRIOR_NORMAL;
  if (p == "DL") return BVB_PRIOR_DL;
  if (p == "GT") return BVB_PRIOR_GT;
  if (p == "HS") return BVB_PRIOR_HS;
  if (p == "SSVS") return BVB_PRIOR_SSVS;
  if (p == "HMP") return BVB_PRIOR_HMP;
  return BVB_PRIOR_NOTHING;
}

const double* ptr_or_null(const NumericVector& v) { return v.size() > 0 ? v.begin() : nullptr; }

}  // namespace

// [[Rcpp::export]]
List bvar_cpp(const NumericMatrix& Y, const NumericMatrix& X, const int& M, const int& T, const int& K,
              const int& draws, const int& burnin, const int& thin, const std::string& tvp_keep,
              const int& intercept, const NumericVector priorIntercept, NumericMatrix& PHI0,
              const List priorPHI_in, const List priorSigma_in, const List Rstartvals_in,
              const IntegerMatrix& i_mat, const IntegerVector& i_vec, const bool& progressbar,
              const double& PHI_tol, const double& L_tol, const bool& huge) {
  bvb_handle* h = the_handle();
  const int Kp = K + intercept;
  const std::string type = as<std::string>(priorSigma_in["type"]);
  const bool fac_on = type == "factor";
  const int r = fac_on ? as<int>(priorSigma_in["factor_factors"]) : 0;
  const int S = M + r;
  const int nsave = draws / thin;
  const bool all = tvp_keep == "all";

  // ---- prior_phi_cpp (R/utility_functions.R:1585-1617) -> bvb_prior_phi
  bvb_prior_phi pp{};
  const std::string prior = as<std::string>(priorPHI_in["prior"]);
  pp.prior = prior_code(prior);
  NumericVector V_i = priorPHI_in["V_i"], a = priorPHI_in["a"], b = priorPHI_in["b"], c = priorPHI_in["c"];
  IntegerVector groups = priorPHI_in["groups"];
  NumericVector a_vec = priorPHI_in["a_vec"], a_weight = priorPHI_in["a_weight"], norm_consts = priorPHI_in["norm_consts"],
                c_vec = priorPHI_in["c_vec"], tau0 = priorPHI_in["SSVS_tau0"], tau1 = priorPHI_in["SSVS_tau1"],
                ssvs_p = priorPHI_in["SSVS_p"], V_i_prep = priorPHI_in["V_i_prep"], l1 = priorPHI_in["lambda_1"],
                l2 = priorPHI_in["lambda_2"];
  // the grid of the discrete hyper-prior on `a`; c_vec only exists when c is tied to a
  std::vector<double> c_vec_full(a_vec.size(), 0.0);
  if (c_vec.size() == a_vec.size()) std::copy(c_vec.begin(), c_vec.end(), c_vec_full.begin());
  pp.V_i = ptr_or_null(V_i); pp.n_groups = as<int>(priorPHI_in["n_groups"]); pp.groups = groups.begin();
  pp.a = ptr_or_null(a); pp.b = ptr_or_null(b); pp.c = ptr_or_null(c); pp.GL_tol = as<double>(priorPHI_in["GL_tol"]);
  pp.c_rel_a = as<bool>(priorPHI_in["c_rel_a"]); pp.DL_hyper = as<bool>(priorPHI_in["DL_hyper"]);
  pp.DL_plus = as<bool>(priorPHI_in["DL_plus"]); pp.GT_hyper = as<bool>(priorPHI_in["GT_hyper"]);
  const bool grid_on = (prior == "DL" && pp.DL_hyper) || (prior == "GT" && pp.GT_hyper);
  pp.grid_len = grid_on ? (int)a_vec.size() : 0;
  pp.a_vec = ptr_or_null(a_vec); pp.a_weight = ptr_or_null(a_weight); pp.norm_consts = ptr_or_null(norm_consts);
  pp.c_vec = c_vec_full.empty() ? nullptr : c_vec_full.data();
  pp.GT_vs = as<double>(priorPHI_in["GT_vs"]);
  pp.GT_priorkernel = as<std::string>(priorPHI_in["GT_priorkernel"]) == "exponential";
  pp.SSVS_tau0 = ptr_or_null(tau0); pp.SSVS_tau1 = ptr_or_null(tau1); pp.SSVS_p = ptr_or_null(ssvs_p);
  pp.SSVS_s_a = as<double>(priorPHI_in["SSVS_s_a"]); pp.SSVS_s_b = as<double>(priorPHI_in["SSVS_s_b"]);
  pp.SSVS_hyper = as<bool>(priorPHI_in["SSVS_hyper"]);
  pp.V_i_prep = ptr_or_null(V_i_prep);
  for (int i = 0; i < 2; ++i) { pp.lambda_1[i] = l1.size() > i ? l1[i] : 0; pp.lambda_2[i] = l2.size() > i ? l2[i] : 0; }

  // ---- prior_sigma_cpp (R/utility_functions.R:374-420) -> bvb_prior_sigma.  Every key of the list exists for both
  // structures (the constructor fills the unused ones with placeholders), exactly as src/bvar_cpp.cpp:166-340 reads them.
  bvb_prior_sigma ps{};
  ps.type = fac_on ? BVB_SIGMA_FACTOR : BVB_SIGMA_CHOLESKY;
  // cholesky_* (bvar_cpp.cpp:166-245)
  const std::string priorU = as<std::string>(priorSigma_in["cholesky_U_prior"]);
  NumericVector cV_i = priorSigma_in["cholesky_V_i"], ca_vec = priorSigma_in["cholesky_a_vec"],
                ca_weight = priorSigma_in["cholesky_a_weight"], cnorm = priorSigma_in["cholesky_norm_consts"],
                cc_vec = priorSigma_in["cholesky_c_vec"], ctau0 = priorSigma_in["cholesky_SSVS_tau0"],
                ctau1 = priorSigma_in["cholesky_SSVS_tau1"], cssvs_p = priorSigma_in["cholesky_SSVS_p"],
                clam3 = priorSigma_in["cholesky_lambda_3"], csv_offset = priorSigma_in["cholesky_sv_offset"];
  NumericMatrix chomo = priorSigma_in["cholesky_priorhomoscedastic"];
  std::vector<double> cc_vec_full(ca_vec.size(), 0.0);
  if (cc_vec.size() == ca_vec.size()) std::copy(cc_vec.begin(), cc_vec.end(), cc_vec_full.begin());
  ps.cholesky_U_prior = prior_code(priorU);
  ps.cholesky_V_i = ptr_or_null(cV_i);
  ps.cholesky_a = as<double>(priorSigma_in["cholesky_a"]); ps.cholesky_b = as<double>(priorSigma_in["cholesky_b"]);
  ps.cholesky_c = as<double>(priorSigma_in["cholesky_c"]); ps.cholesky_GL_tol = as<double>(priorSigma_in["cholesky_GL_tol"]);
  ps.cholesky_c_rel_a = as<bool>(priorSigma_in["cholesky_c_rel_a"]); ps.cholesky_DL_hyper = as<bool>(priorSigma_in["cholesky_DL_hyper"]);
  ps.cholesky_DL_plus = as<bool>(priorSigma_in["cholesky_DL_plus"]); ps.cholesky_GT_hyper = as<bool>(priorSigma_in["cholesky_GT_hyper"]);
  const bool cgrid_on = (priorU == "DL" && ps.cholesky_DL_hyper) || (priorU == "GT" && ps.cholesky_GT_hyper);
  ps.cholesky_grid_len = cgrid_on ? (int)ca_vec.size() : 0;
  ps.cholesky_a_vec = ptr_or_null(ca_vec); ps.cholesky_a_weight = ptr_or_null(ca_weight); ps.cholesky_norm_consts = ptr_or_null(cnorm);
  ps.cholesky_c_vec = cc_vec_full.empty() ? nullptr : cc_vec_full.data();
  ps.cholesky_GT_vs = as<double>(priorSigma_in["cholesky_GT_vs"]);
  ps.cholesky_GT_priorkernel = as<std::string>(priorSigma_in["cholesky_GT_priorkernel"]) == "exponential";
  ps.cholesky_SSVS_tau0 = ptr_or_null(ctau0); ps.cholesky_SSVS_tau1 = ptr_or_null(ctau1); ps.cholesky_SSVS_p = ptr_or_null(cssvs_p);
  ps.cholesky_SSVS_s_a = as<double>(priorSigma_in["cholesky_SSVS_s_a"]); ps.cholesky_SSVS_s_b = as<double>(priorSigma_in["cholesky_SSVS_s_b"]);
  ps.cholesky_SSVS_hyper = as<bool>(priorSigma_in["cholesky_SSVS_hyper"]);
  for (int i = 0; i < 2; ++i) ps.cholesky_lambda_3[i] = clam3.size() > i ? clam3[i] : 0;
  ps.cholesky_priorhomoscedastic = chomo.nrow() > 0 ? chomo.begin() : nullptr;
  ps.cholesky_sv_offset = ptr_or_null(csv_offset);
  // factor_* (bvar_cpp.cpp:258-312)
  List shr = priorSigma_in["factor_shrinkagepriors"];
  NumericVector sh_a = shr["a"], sh_c = shr["c"], sh_d = shr["d"];
  IntegerMatrix restr = priorSigma_in["factor_restrinv"];
  NumericMatrix homo = priorSigma_in["factor_priorhomoskedastic"], ps2 = priorSigma_in["sv_priorsigma2"];
  LogicalVector het_l = priorSigma_in["sv_heteroscedastic"];
  IntegerVector het(het_l.size());
  for (int i = 0; i < het_l.size(); ++i) het[i] = het_l[i] ? 1 : 0;
  NumericVector pmu = priorSigma_in["sv_priormu"], pphi = priorSigma_in["sv_priorphi"], ph0 = priorSigma_in["sv_priorh0"];
  ps.factor_factors = r;
  ps.factor_shrink_a = ptr_or_null(sh_a); ps.factor_shrink_c = ptr_or_null(sh_c); ps.factor_shrink_d = ptr_or_null(sh_d);
  ps.factor_ngprior = as<bool>(priorSigma_in["factor_ngprior"]); ps.factor_columnwise = as<bool>(priorSigma_in["factor_columnwise"]);
  ps.factor_facloadtol = as<double>(priorSigma_in["factor_facloadtol"]);
  ps.factor_restrinv = restr.nrow() > 0 ? restr.begin() : nullptr;
  ps.factor_priorhomoskedastic = homo.nrow() > 0 ? homo.begin() : nullptr;
  ps.factor_interweaving = as<int>(priorSigma_in["factor_interweaving"]);
  // sv_* (bvar_cpp.cpp:343-398)
  ps.sv_heteroscedastic = het.begin();
  for (int i = 0; i < 2; ++i) ps.sv_priormu[i] = pmu.size() > i ? pmu[i] : 0;
  for (int i = 0; i < 4; ++i) ps.sv_priorphi[i] = pphi.size() > i ? pphi[i] : (pphi.size() > 0 ? pphi[i % pphi.size()] : 0);
  ps.sv_priorsigma2 = ps2.nrow() > 0 ? ps2.begin() : nullptr; ps.sv_priorh0 = ptr_or_null(ph0);

  // ---- start values (R/bvar_wrapper.R:352-385); clone semantics: the library copies, never mutates
  bvb_startvals sv{};
  NumericMatrix PHI_in = Rstartvals_in["PHI"], U_in = Rstartvals_in["U"], sv_para = Rstartvals_in["sv_para"],
                sv_logvar = Rstartvals_in["sv_logvar"];
  NumericVector sv_logvar0 = Rstartvals_in["sv_logvar0"];
  List fst = Rstartvals_in["factor_startval"];
  NumericMatrix facload = fst["facload"], fac = fst["fac"], tau2 = fst["tau2"];
  sv.PHI = PHI_in.begin(); sv.U = U_in.nrow() > 0 ? U_in.begin() : nullptr; sv.sv_para = sv_para.begin();
  sv.sv_logvar = sv_logvar.begin(); sv.sv_logvar0 = sv_logvar0.begin();
  sv.facload = facload.nrow() > 0 ? facload.begin() : nullptr; sv.fac = fac.nrow() > 0 ? fac.begin() : nullptr;
  sv.tau2 = tau2.nrow() > 0 ? tau2.begin() : nullptr;

  // ---- R-owned outputs, same shapes as src/bvar_cpp.cpp:409-482
  const int n = K * M, G = pp.n_groups;
  int hyper_rows = 0;
  if (prior == "DL" || prior == "GT" || prior == "HS") hyper_rows = 2 * G + 2 * n;
  else if (prior == "SSVS") hyper_rows = 2 * n;
  else if (prior == "HMP") hyper_rows = 2;
  NumericVector PHI_draws((size_t)Kp * M * nsave), V_prior_draws((size_t)Kp * M * nsave),
      logvar_draws((size_t)(all ? T : 1) * S * nsave), sv_para_draws((size_t)3 * S * nsave);
  PHI_draws.attr("dim") = IntegerVector::create(Kp, M, nsave);
  V_prior_draws.attr("dim") = IntegerVector::create(Kp, M, nsave);
  logvar_draws.attr("dim") = IntegerVector::create(all ? T : 1, S, nsave);
  sv_para_draws.attr("dim") = IntegerVector::create(3, S, nsave);
  NumericMatrix phi_hyper(hyper_rows, prior != "normal" ? nsave : 0);
  NumericVector facload_draws(fac_on ? (size_t)M * r * nsave : 0), fac_draws(fac_on ? (size_t)r * (all ? T : 1) * nsave : 0);
  facload_draws.attr("dim") = IntegerVector::create(fac_on ? M : 0, fac_on ? r : 0, fac_on ? nsave : 0);
  fac_draws.attr("dim") = IntegerVector::create(fac_on ? r : 0, fac_on ? (all ? T : 1) : 0, fac_on ? nsave : 0);
  // cholesky: U (n_U x nsave, trimatu_ind order) and its hyper-parameters (bvar_cpp.cpp:424-425, 463-482)
  const int n_U = fac_on ? 0 : (M * M - M) / 2;
  int u_rows = 0;
  if (!fac_on) {
    if (priorU == "DL" || priorU == "GT" || priorU == "HS") u_rows = 2 + 2 * n_U;
    else if (priorU == "SSVS") u_rows = 2 * n_U;
    else if (priorU == "HMP") u_rows = 1;
  }
  NumericMatrix U_draws(n_U, fac_on ? 0 : nsave), u_hyper(fac_on ? 0 : u_rows, (!fac_on && priorU != "normal") ? nsave : 0);
  bvb_draws out{};
  out.PHI = PHI_draws.begin(); out.V_prior = V_prior_draws.begin(); out.logvar = logvar_draws.begin();
  out.sv_para = sv_para_draws.begin();
  out.phi_hyperparameter = (hyper_rows && nsave) ? phi_hyper.begin() : nullptr;
  out.facload = fac_on && r && nsave ? facload_draws.begin() : nullptr;
  out.fac = fac_on && r && nsave ? fac_draws.begin() : nullptr;
  out.U = (n_U && nsave) ? U_draws.begin() : nullptr;
  out.u_hyperparameter = (u_rows && nsave && priorU != "normal") ? u_hyper.begin() : nullptr;

  // several GPUs (factor structure only: the cholesky structure is sequential over the equations)
  MultiGpu& mg = the_gpus();
  const bool multi = mg.n > 1 && fac_on && !huge && K > 0;
  if (multi) {
    for (int i = 0; i < mg.n; ++i) bvb_set_seed(mg.h[i], g_call_seed);
    check(mg.h[0], bvb_gibbs_init_multi(mg.n, mg.h.data(), Y.begin(), X.begin(), M, T, K, draws, burnin, thin, all ? 1 : 0, intercept,
                                        priorIntercept.begin(), PHI0.begin(), &pp, &ps, &sv, i_vec.begin(), PHI_tol, L_tol, 0, &out));
  } else
  check(h, bvb_gibbs_init(h, Y.begin(), X.begin(), M, T, K, draws, burnin, thin, all ? 1 : 0, intercept,
                          priorIntercept.begin(), PHI0.begin(), &pp, &ps, &sv, i_vec.begin(), PHI_tol, L_tol,
                          huge ? 1 : 0, &out));
  (void)i_mat;   // (the reference only uses i_vec inside bvar_cpp as well)
  const int tot = draws + burnin;
  int done = 0;
  if (progressbar) Rprintf("\r###  %i / %i ### (%3.0f%%) ###", 0, tot, 0.);
  while (done < tot) {
    Rcpp::checkUserInterrupt();                       // every 256 sweeps, as bvar_cpp.cpp:501-503
    if (multi) check(mg.h[0], bvb_gibbs_run_multi(mg.n, mg.h.data(), 256, &done));
    else check(h, bvb_gibbs_run(h, 256, &done));
    if (progressbar) Rprintf("\r###  %i / %i ### (%3.0f%%) ###", done, tot, 100. * done / tot);
  }
  return List::create(Named("PHI") = PHI_draws, Named("U") = U_draws, Named("logvar") = logvar_draws,
                      Named("sv_para") = sv_para_draws, Named("phi_hyperparameter") = phi_hyper,
                      Named("u_hyperparameter") = u_hyper, Named("V_prior") = V_prior_draws,
                      Named("facload") = facload_draws, Named("fac") = fac_draws);
}

// [[Rcpp::export]]
NumericMatrix my_gig(int n, NumericVector lambda, NumericVector chi, NumericVector psi) {
  bvb_handle* h = the_handle();
  const int m = (int)std::max({lambda.size(), chi.size(), psi.size()});
  NumericMatrix out(n, m);
  check(h, bvb_gig(h, n, lambda.begin(), (int)lambda.size(), chi.begin(), (int)chi.size(), psi.begin(), (int)psi.size(),
                   /*stream*/ 0, out.begin()));   // the handle was just re-keyed: stream 0 of a fresh key
  return out;
}

// [[Rcpp::export]]
NumericMatrix sample_PHI_cholesky(const NumericMatrix PHI, const NumericMatrix& PHI_prior, const NumericMatrix& Y,
                                  const NumericMatrix& X, const NumericMatrix& U, const NumericMatrix& d_sqrt,
                                  const NumericMatrix& V_prior) {
  bvb_handle* h = the_handle();
  NumericMatrix out(PHI.nrow(), PHI.ncol());
  // the reference consumes R::norm_rand here (Kp per equation, equations in order); passing them keeps set.seed()
  // semantics value for value
  NumericVector z((size_t)PHI.nrow() * PHI.ncol());
  for (R_xlen_t i = 0; i < z.size(); ++i) z[i] = R::norm_rand();
  check(h, bvb_phi_draw_cholesky(h, X.nrow(), X.ncol(), Y.ncol(), PHI.begin(), PHI_prior.begin(), Y.begin(),
                                 X.begin(), U.begin(), d_sqrt.begin(), V_prior.begin(), z.begin(), out.begin()));
  return out;
}

// src/irf.cpp:467-524.  The reference returns arma::field<arma::cube> (an R list of n_draws cubes, each
// variables x shocks x (1 + ahead)); R/irf.R:275-276 immediately does unlist() + dim<-, so the list elements are views
// of one contiguous result here: element r is draw r's (M, n_shocks, ahead + 1) block.
// [[Rcpp::export]]
List irf_cpp(const NumericVector& coefficients, const NumericVector& factor_loadings, const NumericMatrix& U_vecs,
             const NumericMatrix& logvar_t, const NumericMatrix& shocks, const int ahead,
             const Nullable<NumericVector> rotation_ = R_NilValue) {
  bvb_handle* h = the_handle();
  const IntegerVector dc = coefficients.attr("dim");         // (Kp, M, draws)
  const int Kp = dc[0], M = dc[1], R = dc[2];
  int n_factors = 0;
  if (factor_loadings.size() > 0) { const IntegerVector df = factor_loadings.attr("dim"); n_factors = df[1]; }
  const int has_U = (n_factors == 0 && U_vecs.ncol() > 0) ? 1 : 0;
  const int dim = shocks.nrow(), n_shocks = shocks.ncol();
  NumericVector rot;
  const double* rot_p = nullptr;
  if (rotation_.isNotNull()) { rot = rotation_.get(); rot_p = rot.begin(); }
  const size_t per = (size_t)M * n_shocks * (ahead + 1);
  NumericVector all(per * R);
  check(h, bvb_irf(h, Kp, M, R, n_factors, has_U, logvar_t.nrow(), dim, n_shocks, ahead, coefficients.begin(),
                   n_factors ? factor_loadings.begin() : nullptr, has_U ? U_vecs.begin() : nullptr, logvar_t.begin(),
                   shocks.begin(), rot_p, all.begin()));
  List ret(R);
  for (int r = 0; r < R; ++r) {
    NumericVector cube(all.begin() + per * r, all.begin() + per * (r + 1));
    cube.attr("dim") = IntegerVector::create(M, n_shocks, ahead + 1);
    ret[r] = cube;
  }
  return ret;
}

// src/utilities_cpp.cpp:216-393.  The standard normals the reference takes from R::norm_rand (log-variance forecasts,
// :325-327; predictive draws, :371-374) are counter-based device normals here, keyed from R's RNG through the handle.
// [[Rcpp::export]]
List out_of_sample(const int& each, const NumericVector& X_T_plus_1, const NumericVector& PHI, const NumericMatrix& U,
                   const NumericVector& facload, const NumericMatrix& logvar_T, const IntegerVector& ahead,
                   const NumericMatrix& sv_mu, const NumericMatrix& sv_phi, const NumericMatrix& sv_sigma,
                   const IntegerVector& sv_indicator, const bool& factor, const bool& LPL, const NumericMatrix& Y_obs,
                   const bool& LPL_subset, const IntegerVector& VoI, const bool& simulate_predictive) {
  bvb_handle* h = the_handle();
  const IntegerVector dp = PHI.attr("dim");                   // (Kp, M, draws)
  const int Kp = dp[0], M = dp[1], R = dp[2];
  const int nsave = R * each, na = (int)ahead.size();
  int n_factors = 0;
  if (factor && facload.size() > 0) { const IntegerVector df = facload.attr("dim"); n_factors = df[1]; }
  NumericVector predictions((size_t)(simulate_predictive ? na : 0) * (simulate_predictive ? M : 0) * (simulate_predictive ? nsave : 0));
  predictions.attr("dim") = IntegerVector::create(simulate_predictive ? na : 0, simulate_predictive ? M : 0, simulate_predictive ? nsave : 0);
  NumericMatrix LPL_draws(LPL ? na : 0, LPL ? nsave : 0);
  NumericVector PL_uni((size_t)(LPL ? na : 0) * (LPL ? M : 0) * (LPL ? nsave : 0));
  PL_uni.attr("dim") = IntegerVector::create(LPL ? na : 0, LPL ? M : 0, LPL ? nsave : 0);
  NumericMatrix LPL_sub(LPL_subset ? na : 0, LPL_subset ? nsave : 0);
  const int n_sv = (int)sv_indicator.size();
  check(h, bvb_predict(h, each, Kp, M, R, n_factors, logvar_T.nrow(), factor ? 1 : 0, X_T_plus_1.begin(), PHI.begin(),
                       (!factor && U.nrow() > 0) ? U.begin() : nullptr, n_factors ? facload.begin() : nullptr,
                       logvar_T.begin(), ahead.begin(), na, n_sv ? sv_mu.begin() : nullptr, n_sv ? sv_phi.begin() : nullptr,
                       n_sv ? sv_sigma.begin() : nullptr, n_sv ? sv_indicator.begin() : nullptr, n_sv, LPL ? 1 : 0,
                       LPL ? Y_obs.begin() : nullptr, LPL_subset ? 1 : 0, VoI.size() > 0 ? VoI.begin() : nullptr, (int)VoI.size(),
                       simulate_predictive ? 1 : 0, /*z_logvar*/ nullptr, /*z_pred*/ nullptr,
                       LPL ? LPL_draws.begin() : nullptr, LPL ? PL_uni.begin() : nullptr,
                       LPL_subset ? LPL_sub.begin() : nullptr, simulate_predictive ? predictions.begin() : nullptr));
  return List::create(Named("LPL_draws") = LPL_draws, Named("PL_univariate_draws") = PL_uni,
                      Named("LPL_sub_draws") = LPL_sub, Named("predictions") = predictions);
}

// src/utilities_cpp.cpp:436-472 -> bvb_vcov (same arguments, same (T*M, M, nsave) result)
// [[Rcpp::export]]
NumericVector vcov_cpp(const bool& factor, const NumericVector& facload, const NumericVector& logvar, const NumericMatrix& U,
                       const int& M, const int& factors) {
  bvb_handle* h = the_handle();
  const IntegerVector dl = logvar.attr("dim");          // (T, M + factors | M, nsave)
  const int T = dl[0], nsave = dl[2];
  NumericVector out((size_t)T * M * M * nsave);
  out.attr("dim") = IntegerVector::create(T * M, M, nsave);
  check(h, bvb_vcov(h, factor ? 1 : 0, T, M, factors, nsave, facload.size() > 0 ? facload.begin() : nullptr, logvar.begin(),
                    U.nrow() > 0 ? U.begin() : nullptr, out.begin()));
  return out;
}

// src/utilities_cpp.cpp:181-212 -> bvb_predh (device Philox normals; the key comes from R's RNG through the handle)
// [[Rcpp::export]]
NumericVector predh(const NumericMatrix& logvar_T, const IntegerVector& ahead, const int& each, const NumericMatrix& sv_mu,
                    const NumericMatrix& sv_phi, const NumericMatrix& sv_sigma) {
  bvb_handle* h = the_handle();
  const int M = logvar_T.nrow(), R = logvar_T.ncol();
  NumericVector out((size_t)ahead.size() * M * R * each);
  out.attr("dim") = IntegerVector::create((int)ahead.size(), M, R * each);
  check(h, bvb_predh(h, M, R, each, ahead.begin(), (int)ahead.size(), logvar_T.begin(), sv_mu.begin(), sv_phi.begin(),
                     sv_sigma.begin(), /*z*/ nullptr, out.begin()));
  return out;
}
