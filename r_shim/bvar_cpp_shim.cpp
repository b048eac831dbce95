// Rcpp shim: keeps the reference's .Call symbols and R-visible signatures
// (src/RcppExports.cpp:16-44 `_bayesianVARs_bvar_cpp`, :46-58 `_bayesianVARs_my_gig`,
// :224-236 `_bayesianVARs_sample_PHI_cholesky`) and forwards to the C ABI of
// libbvb.so (include/bvb.h).  It replaces src/bvar_cpp.cpp and src/gig_vec.cpp in
// the package; R/ stays untouched.  PKG_LIBS gains `-lbvb`.
//
// This file cannot be compiled in the build container (no R, no Rcpp headers);
// it is a mechanical translation: unpack the two lists into the POD structs,
// allocate the R-owned outputs exactly as src/bvar_cpp.cpp:409-482 does, run the
// sampler in chunks of 256 sweeps so Rcpp::checkUserInterrupt() and the progress
// line keep their cadence (src/bvar_cpp.cpp:501-503, :838-843), and turn error
// codes back into Rcpp::stop() with the reference's wording.
#include <Rcpp.h>
#include "bvb.h"

using namespace Rcpp;

namespace {

struct Handle {
  bvb_handle* h = nullptr;
  Handle() {
    // the Philox key is drawn from R's RNG so set.seed() keeps runs reproducible
    const double u1 = R::unif_rand(), u2 = R::unif_rand();
    const uint64_t seed = ((uint64_t)(u1 * 4294967296.0) << 32) | (uint64_t)(u2 * 4294967296.0);
    if (bvb_create(&h, 0, seed) != BVB_OK) stop("bayesianVARs: no usable sm_100a CUDA device (there is no CPU fallback)");
  }
  ~Handle() { bvb_destroy(h); }
};

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

}  // namespace

// [[Rcpp::export]]
List bvar_cpp(const NumericMatrix& Y, const NumericMatrix& X, const int& M, const int& T, const int& K,
              const int& draws, const int& burnin, const int& thin, const std::string& tvp_keep,
              const int& intercept, const NumericVector priorIntercept, NumericMatrix& PHI0,
              const List priorPHI_in, const List priorSigma_in, const List Rstartvals_in,
              const IntegerMatrix& i_mat, const IntegerVector& i_vec, const bool& progressbar,
              const double& PHI_tol, const double& L_tol, const bool& huge) {
  Handle H;
  const int Kp = K + intercept;
  const std::string type = priorSigma_in["type"];
  const int r = priorSigma_in["factor_factors"];
  const int S = M + (type == "factor" ? r : 0);
  const int nsave = draws / thin;
  const bool all = tvp_keep == "all";

  // ---- prior_phi_cpp (R/utility_functions.R:1585-1617) -> bvb_prior_phi
  bvb_prior_phi pp{};
  const std::string prior = priorPHI_in["prior"];
  pp.prior = prior_code(prior);
  NumericVector V_i = priorPHI_in["V_i"], a = priorPHI_in["a"], b = priorPHI_in["b"], c = priorPHI_in["c"];
  IntegerVector groups = priorPHI_in["groups"];
  NumericVector a_vec = priorPHI_in["a_vec"], a_weight = priorPHI_in["a_weight"], norm_consts = priorPHI_in["norm_consts"],
                c_vec = priorPHI_in["c_vec"], tau0 = priorPHI_in["SSVS_tau0"], tau1 = priorPHI_in["SSVS_tau1"],
                ssvs_p = priorPHI_in["SSVS_p"], V_i_prep = priorPHI_in["V_i_prep"], l1 = priorPHI_in["lambda_1"],
                l2 = priorPHI_in["lambda_2"];
  pp.V_i = V_i.begin(); pp.n_groups = priorPHI_in["n_groups"]; pp.groups = groups.begin();
  pp.a = a.begin(); pp.b = b.begin(); pp.c = c.begin(); pp.GL_tol = priorPHI_in["GL_tol"];
  pp.a_vec = a_vec.begin(); pp.a_weight = a_weight.begin(); pp.norm_consts = norm_consts.begin();
  pp.c_vec = c_vec.begin(); pp.grid_len = (prior == "DL" ? as<bool>(priorPHI_in["DL_hyper"]) : as<bool>(priorPHI_in["GT_hyper"])) ? a_vec.size() : 0;
  pp.c_rel_a = as<bool>(priorPHI_in["c_rel_a"]); pp.DL_hyper = as<bool>(priorPHI_in["DL_hyper"]);
  pp.DL_plus = as<bool>(priorPHI_in["DL_plus"]); pp.GT_hyper = as<bool>(priorPHI_in["GT_hyper"]);
  pp.GT_vs = priorPHI_in["GT_vs"];
  pp.GT_priorkernel = as<std::string>(priorPHI_in["GT_priorkernel"]) == "exponential";
  pp.SSVS_tau0 = tau0.begin(); pp.SSVS_tau1 = tau1.begin(); pp.SSVS_p = ssvs_p.begin();
  pp.SSVS_s_a = priorPHI_in["SSVS_s_a"]; pp.SSVS_s_b = priorPHI_in["SSVS_s_b"];
  pp.SSVS_hyper = as<bool>(priorPHI_in["SSVS_hyper"]);
  pp.V_i_prep = V_i_prep.begin();
  for (int i = 0; i < 2; ++i) { pp.lambda_1[i] = l1.size() > i ? l1[i] : 0; pp.lambda_2[i] = l2.size() > i ? l2[i] : 0; }

  // ---- prior_sigma_cpp (R/utility_functions.R:374-420) -> bvb_prior_sigma
  bvb_prior_sigma ps{};
  ps.type = type == "factor" ? BVB_SIGMA_FACTOR : BVB_SIGMA_CHOLESKY;
  List shr = priorSigma_in["factor_shrinkagepriors"];
  NumericVector sh_a = shr["a"], sh_c = shr["c"], sh_d = shr["d"];
  IntegerMatrix restr = priorSigma_in["factor_restrinv"];
  NumericMatrix homo = priorSigma_in["factor_priorhomoskedastic"], ps2 = priorSigma_in["sv_priorsigma2"];
  LogicalVector het_l = priorSigma_in["sv_heteroscedastic"];
  IntegerVector het(het_l.begin(), het_l.end());
  NumericVector pmu = priorSigma_in["sv_priormu"], pphi = priorSigma_in["sv_priorphi"], ph0 = priorSigma_in["sv_priorh0"];
  ps.factor_factors = r;
  ps.factor_shrink_a = sh_a.begin(); ps.factor_shrink_c = sh_c.begin(); ps.factor_shrink_d = sh_d.begin();
  ps.factor_ngprior = as<bool>(priorSigma_in["factor_ngprior"]); ps.factor_columnwise = as<bool>(priorSigma_in["factor_columnwise"]);
  ps.factor_facloadtol = priorSigma_in["factor_facloadtol"]; ps.factor_restrinv = restr.begin();
  ps.factor_priorhomoskedastic = homo.begin(); ps.factor_interweaving = priorSigma_in["factor_interweaving"];
  ps.sv_heteroscedastic = het.begin();
  for (int i = 0; i < 2; ++i) ps.sv_priormu[i] = pmu.size() > i ? pmu[i] : 0;
  for (int i = 0; i < 4; ++i) ps.sv_priorphi[i] = pphi.size() > i ? pphi[i] : pphi[i % 2];
  ps.sv_priorsigma2 = ps2.begin(); ps.sv_priorh0 = ph0.begin();
  // (cholesky_* keys are filled the same way once bvb_gibbs_init's cholesky branch lands)

  // ---- start values (R/bvar_wrapper.R:352-385); clone semantics: the library copies, never mutates
  bvb_startvals sv{};
  NumericMatrix PHI_in = Rstartvals_in["PHI"], U_in = Rstartvals_in["U"], sv_para = Rstartvals_in["sv_para"],
                sv_logvar = Rstartvals_in["sv_logvar"];
  NumericVector sv_logvar0 = Rstartvals_in["sv_logvar0"];
  List fst = Rstartvals_in["factor_startval"];
  NumericMatrix facload = fst["facload"], fac = fst["fac"], tau2 = fst["tau2"];
  sv.PHI = PHI_in.begin(); sv.U = U_in.begin(); sv.sv_para = sv_para.begin(); sv.sv_logvar = sv_logvar.begin();
  sv.sv_logvar0 = sv_logvar0.begin(); sv.facload = facload.begin(); sv.fac = fac.begin(); sv.tau2 = tau2.begin();

  // ---- R-owned outputs, same shapes as src/bvar_cpp.cpp:409-482
  const int n = K * M, G = pp.n_groups;
  int hyper_rows = 0;
  if (prior == "DL" || prior == "GT" || prior == "HS") hyper_rows = 2 * G + 2 * n;
  else if (prior == "SSVS") hyper_rows = 2 * n;
  else if (prior == "HMP") hyper_rows = 2;
  NumericVector PHI_draws(Kp * M * nsave), V_prior_draws(Kp * M * nsave), logvar_draws((all ? T : 1) * S * nsave),
      sv_para_draws(3 * S * nsave);
  PHI_draws.attr("dim") = IntegerVector::create(Kp, M, nsave);
  V_prior_draws.attr("dim") = IntegerVector::create(Kp, M, nsave);
  logvar_draws.attr("dim") = IntegerVector::create(all ? T : 1, S, nsave);
  sv_para_draws.attr("dim") = IntegerVector::create(3, S, nsave);
  NumericMatrix phi_hyper(hyper_rows, prior != "normal" ? nsave : 0);
  const bool fac_on = type == "factor";
  NumericVector facload_draws(fac_on ? M * r * nsave : 0), fac_draws(fac_on ? r * (all ? T : 1) * nsave : 0);
  facload_draws.attr("dim") = IntegerVector::create(fac_on ? M : 0, fac_on ? r : 0, fac_on ? nsave : 0);
  fac_draws.attr("dim") = IntegerVector::create(fac_on ? r : 0, fac_on ? (all ? T : 1) : 0, fac_on ? nsave : 0);
  NumericMatrix U_draws(0, 0), u_hyper(0, 0);
  bvb_draws out{};
  out.PHI = PHI_draws.begin(); out.V_prior = V_prior_draws.begin(); out.logvar = logvar_draws.begin();
  out.sv_para = sv_para_draws.begin(); out.phi_hyperparameter = hyper_rows ? phi_hyper.begin() : nullptr;
  out.facload = fac_on && r ? facload_draws.begin() : nullptr; out.fac = fac_on && r ? fac_draws.begin() : nullptr;

  check(H.h, bvb_gibbs_init(H.h, Y.begin(), X.begin(), M, T, K, draws, burnin, thin, all ? 1 : 0, intercept,
                            priorIntercept.begin(), PHI0.begin(), &pp, &ps, &sv, i_vec.begin(), PHI_tol, L_tol,
                            huge ? 1 : 0, &out));
  const int tot = draws + burnin;
  int done = 0;
  if (progressbar) Rprintf("\r###  %i / %i ### (%3.0f%%) ###", 0, tot, 0.);
  while (done < tot) {
    Rcpp::checkUserInterrupt();                       // every 256 sweeps, as bvar_cpp.cpp:501-503
    check(H.h, bvb_gibbs_run(H.h, 256, &done));
    if (progressbar) Rprintf("\r###  %i / %i ### (%3.0f%%) ###", done, tot, 100. * done / tot);
  }
  return List::create(Named("PHI") = PHI_draws, Named("U") = U_draws, Named("logvar") = logvar_draws,
                      Named("sv_para") = sv_para_draws, Named("phi_hyperparameter") = phi_hyper,
                      Named("u_hyperparameter") = u_hyper, Named("V_prior") = V_prior_draws,
                      Named("facload") = facload_draws, Named("fac") = fac_draws);
}

// [[Rcpp::export]]
NumericMatrix my_gig(int n, NumericVector lambda, NumericVector chi, NumericVector psi) {
  Handle H;
  const int m = std::max({lambda.size(), chi.size(), psi.size()});
  NumericMatrix out(n, m);
  static uint64_t stream = 0;
  check(H.h, bvb_gig(H.h, n, lambda.begin(), lambda.size(), chi.begin(), chi.size(), psi.begin(), psi.size(), stream++, out.begin()));
  return out;
}

// [[Rcpp::export]]
NumericMatrix sample_PHI_cholesky(const NumericMatrix PHI, const NumericMatrix& PHI_prior, const NumericMatrix& Y,
                                  const NumericMatrix& X, const NumericMatrix& U, const NumericMatrix& d_sqrt,
                                  const NumericMatrix& V_prior) {
  Handle H;
  NumericMatrix out(PHI.nrow(), PHI.ncol());
  // the reference consumes R::norm_rand here; passing them keeps set.seed() semantics bit-for-bit
  NumericVector z(PHI.nrow() * PHI.ncol());
  for (auto& v : z) v = R::norm_rand();
  check(H.h, bvb_phi_draw_cholesky(H.h, X.nrow(), X.ncol(), Y.ncol(), PHI.begin(), PHI_prior.begin(), Y.begin(),
                                   X.begin(), U.begin(), d_sqrt.begin(), V_prior.begin(), z.begin(), out.begin()));
  return out;
}

// src/utilities_cpp.cpp:436-472 -> bvb_vcov (same arguments, same (T*M, M, nsave) result)
// [[Rcpp::export]]
NumericVector vcov_cpp(const bool& factor, const NumericVector& facload, const NumericVector& logvar, const NumericMatrix& U,
                       const int& M, const int& factors) {
  Handle H;
  const IntegerVector dl = logvar.attr("dim");          // (T, M + factors | M, nsave)
  const int T = dl[0], nsave = dl[2];
  NumericVector out((size_t)T * M * M * nsave);
  out.attr("dim") = Dimension(T * M, M, nsave);
  check(H.h, bvb_vcov(H.h, factor ? 1 : 0, T, M, factors, nsave, facload.begin(), logvar.begin(), U.begin(), out.begin()));
  return out;
}

// src/utilities_cpp.cpp:181-212 -> bvb_predh (device Philox normals; the key comes from R's RNG through Handle)
// [[Rcpp::export]]
NumericVector predh(const NumericMatrix& logvar_T, const IntegerVector& ahead, const int& each, const NumericMatrix& sv_mu,
                    const NumericMatrix& sv_phi, const NumericMatrix& sv_sigma) {
  Handle H;
  const int M = logvar_T.nrow(), R = logvar_T.ncol();
  NumericVector out((size_t)ahead.size() * M * R * each);
  out.attr("dim") = Dimension(ahead.size(), M, R * each);
  check(H.h, bvb_predh(H.h, M, R, each, ahead.begin(), ahead.size(), logvar_T.begin(), sv_mu.begin(), sv_phi.begin(),
                       sv_sigma.begin(), /*z*/ nullptr, out.begin()));
  return out;
}
