// Rcpp shim: the reference-side binding of libsaige_b200.so.
//
// Drop this file into src/SAIGE/src/ NEXT TO the reference's SAIGE_fitGLMM_fast.cpp and compile the package with
// -DUSE_SAIGE_B200 (see rcpp_shim/Makevars.b200).  Every function below keeps the name, argument list and return
// type of the [[Rcpp::export]] it replaces, so R/RcppExports.R, src/RcppExports.cpp (CallEntries[], registered with
// R_useDynamicSymbols(dll, FALSE), RcppExports.cpp:2158-2323) and all of R/SAIGE_fitGLMM_fast.R stay byte-for-byte
// unchanged.  In the reference file the same 28 definitions are fenced with `#if !defined(USE_SAIGE_B200)`; all other
// exports (sparse-GRM machinery, step-2 helpers) keep their reference bodies.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: R, Rcpp and RcppArmadillo are not available in the build container
// (SURVEY.md 8c).  The C ABI it calls is exercised end to end by tests/ through the ctypes mirror in
// saige_gpu_b200/api.py, which follows this file line by line.
//
// Boundary rules (include/saige_b200.h): fp64 column-major host buffers, status codes -> Rcpp::stop, RNG stays in R.
#if defined(USE_SAIGE_B200)
#include <RcppArmadillo.h>
#include <cstdlib>
#include <string>
#include <vector>
#include "saige_b200.h"
// [[Rcpp::depends(RcppArmadillo)]]
using namespace Rcpp;

static sgb_ctx *g_h = nullptr;          // replaces `genoClass geno` (SAIGE_fitGLMM_fast.cpp:1188)

static void ck(int rc) { if (rc) Rcpp::stop(std::string("saige_b200: ") + sgb_last_error(g_h)); }

static sgb_ctx *ctx()
{
    if (g_h) return g_h;
    // One R process per GPU, exactly like the reference's `mpirun -n G Rscript step1_fitNULLGLMM.R` (docs/installation.md:244).
    // Under pbdMPI the launcher provides rank/size; rank 0 creates the NCCL id and broadcasts it with pbdMPI::bcast.
    int rank = 0, world = 1;
    const char *dev = std::getenv("SAIGE_B200_DEVICE");
    int device = dev ? std::atoi(dev) : 0;
#if defined(USE_pbdMPI)
    Environment mpi("package:pbdMPI");
    Function comm_rank = mpi["comm.rank"], comm_size = mpi["comm.size"], bcast = mpi["bcast"];
    rank = as<int>(comm_rank()); world = as<int>(comm_size());
    if (!dev) device = rank;             // jsrun/mpirun bind one GPU per rank; override with SAIGE_B200_DEVICE
    if (world > 1) {
        RawVector id(SGB_NCCL_ID_BYTES);
        if (rank == 0 && sgb_nccl_unique_id(&id[0])) Rcpp::stop(sgb_last_error(nullptr));
        id = bcast(id, Named("rank.source", 0));
        if (sgb_create_dist(device, rank, world, &id[0], &g_h)) Rcpp::stop(sgb_last_error(nullptr));
        sgb_set_verbose(g_h, rank == 0 ? 1 : 0);      // the reference's "iter from getPCG1ofSigmaAndVector" lines (FG.cpp:2794-2798)
        return g_h;
    }
#endif
    if (sgb_create(device, &g_h)) Rcpp::stop(sgb_last_error(nullptr));
    sgb_set_verbose(g_h, 1);                          // the reference's "iter from getPCG1ofSigmaAndVector" lines (FG.cpp:2794-2798)
    return g_h;
}

// the same handle for the step-2 shim (SAIGE_step2_b200.cpp)
sgb_ctx *saige_b200_ctx() { return ctx(); }

// R's RNG is the probe source (set_seed(200) + rbinom, SAIGE_fitGLMM_fast.cpp:3040-3054, 3134-3137)
static int draw_probes(void *, int64_t n, int count, double *out)
{
    for (int j = 0; j < count; j++) {
        NumericVector u = Rcpp::rbinom(n, 1, 0.5);
        for (int64_t i = 0; i < n; i++) out[(size_t)j * n + i] = 2.0 * u[i] - 1.0;
    }
    return 0;
}
static void set_seed(unsigned int seed);
// sgb_glmmkin_ai_pcg runs several trace estimates inside one call and announces each with count = 0: re-seed there, as GetTrace
// does (FG.cpp:3114)
static int draw_probes_reseeding(void *u, int64_t n, int count, double *out)
{
    if (count == 0) { set_seed(200); return 0; }
    return draw_probes(u, n, count, out);
}
static void set_seed(unsigned int seed)
{
#if defined(USE_pbdMPI)
    Environment e("package:pbdMPI"); Function f = e["comm.set.seed"]; f(seed);
#else
    Environment e("package:base"); Function f = e["set.seed"]; f(seed);
#endif
}

// ---- configuration ----
// [[Rcpp::export]]
void setminMAFforGRM(float minMAFforGRM) { ck(sgb_set_min_maf_for_grm(ctx(), minMAFforGRM)); }
// [[Rcpp::export]]
void setmaxMissingRateforGRM(float maxMissingforGRM) { ck(sgb_set_max_missing_rate_for_grm(ctx(), maxMissingforGRM)); }
// [[Rcpp::export]]
void setminMAC_VarianceRatio(float t_minMACVarRatio, float t_maxMACVarRatio, bool t_isVarianceRatioinGeno)
{
    ck(sgb_set_min_mac_variance_ratio(ctx(), t_minMACVarRatio, t_maxMACVarRatio, t_isVarianceRatioinGeno));
}
// [[Rcpp::export]]
void setisUseSparseSigmaforNullModelFitting(bool isUseSparseSigmaforModelFitting0)
{
    if (isUseSparseSigmaforModelFitting0) Rcpp::stop("the B200 back end implements the full-GRM path only");
}

// ---- genotype store ----
// [[Rcpp::export]]
void setgeno(std::string bedfile, std::string bimfile, std::string famfile, std::vector<int> &subSampleInGeno,
             std::vector<bool> &indicatorGenoSamplesWithPheno, float memoryChunk, bool isDiagofKinSetAsOne)
{
    std::vector<uint8_t> ind(indicatorGenoSamplesWithPheno.begin(), indicatorGenoSamplesWithPheno.end());
    // g_randMarkerIndforVR = unique(randi(1000, [0, M-1])) (SAIGE_fitGLMM_fast.cpp:866-868), drawn here with R's RNG
    std::vector<int32_t> vr;
    if (sgb_get_is_var_ratio_geno(ctx())) {
        Environment base("package:base"); Function length = base["length"], readLines = base["readLines"];
        int M = as<int>(length(readLines(bimfile)));
        IntegerVector r = Rcpp::sample(M, 1000, true) - 1;
        vr.assign(r.begin(), r.end());
    }
    ck(sgb_setgeno(ctx(), bedfile.c_str(), bimfile.c_str(), famfile.c_str(), subSampleInGeno.data(),
                   (int64_t)subSampleInGeno.size(), ind.data(), (int64_t)ind.size(), isDiagofKinSetAsOne, vr.data(),
                   (int64_t)vr.size()));
}
// [[Rcpp::export]]
void closeGenoFile_plink() { if (g_h) { sgb_destroy(g_h); g_h = nullptr; } Rprintf("closed the plinkFile!\n"); }
// [[Rcpp::export]]
int gettotalMarker() { return (int)sgb_get_total_marker(ctx()); }
// [[Rcpp::export]]
arma::fvec getAlleleFreqVec()
{
    arma::vec v(sgb_get_num_qc_markers(ctx())); ck(sgb_get_allele_freq_vec(ctx(), v.memptr()));
    return arma::conv_to<arma::fvec>::from(v);
}
// [[Rcpp::export]]
arma::ivec getMACVec() { arma::Col<int32_t> v(sgb_get_num_qc_markers(ctx())); ck(sgb_get_mac_vec(ctx(), v.memptr())); return arma::conv_to<arma::ivec>::from(v); }
// [[Rcpp::export]]
arma::ivec getMACVec_forVarRatio() { arma::Col<int32_t> v(sgb_get_num_vr_markers(ctx())); ck(sgb_get_mac_vec_for_var_ratio(ctx(), v.memptr())); return arma::conv_to<arma::ivec>::from(v); }
// [[Rcpp::export]]
arma::ivec getIndexVec_forVarRatio() { arma::Col<int32_t> v(sgb_get_num_vr_markers(ctx())); ck(sgb_get_index_vec_for_var_ratio(ctx(), v.memptr())); return arma::conv_to<arma::ivec>::from(v); }
// [[Rcpp::export]]
bool getIsVarRatioGeno() { return sgb_get_is_var_ratio_geno(ctx()) != 0; }
// [[Rcpp::export]]
std::vector<bool> getQCdMarkerIndex()
{
    std::vector<uint8_t> m(sgb_get_total_marker(ctx())); ck(sgb_get_qcd_marker_index(ctx(), m.data()));
    return std::vector<bool>(m.begin(), m.end());
}
// [[Rcpp::export]]
arma::ivec Get_OneSNP_Geno(int SNPIdx) { arma::Col<int32_t> v(sgb_get_nnomissing(ctx())); ck(sgb_get_one_snp_geno(ctx(), SNPIdx, v.memptr())); return arma::conv_to<arma::ivec>::from(v); }
// [[Rcpp::export]]
arma::ivec Get_OneSNP_Geno_forVarRatio(int SNPIdx) { arma::Col<int32_t> v(sgb_get_nnomissing(ctx())); ck(sgb_get_one_snp_geno_for_var_ratio(ctx(), SNPIdx, v.memptr())); return arma::conv_to<arma::ivec>::from(v); }

// ---- LOCO ----
// [[Rcpp::export]]
void setStartEndIndex(int startIndex, int endIndex, int chromIndex) { ck(sgb_set_start_end_index(ctx(), startIndex, endIndex, chromIndex)); }
// [[Rcpp::export]]
void setStartEndIndexVec(arma::ivec &startIndex_vec, arma::ivec &endIndex_vec)
{
    std::vector<int32_t> s(startIndex_vec.begin(), startIndex_vec.end()), e(endIndex_vec.begin(), endIndex_vec.end());
    ck(sgb_set_start_end_index_vec(ctx(), s.data(), e.data(), (int)s.size()));
}
// [[Rcpp::export]]
void set_Diagof_StdGeno_LOCO() { ck(sgb_set_diag_of_stdgeno_loco(ctx())); }

// ---- AI-REML ----  (arma::fvec& in the reference; the data crosses the C ABI as fp64)
static arma::vec d(const arma::fvec &v) { return arma::conv_to<arma::vec>::from(v); }
static arma::mat d(const arma::fmat &m) { return arma::conv_to<arma::mat>::from(m); }

static Rcpp::List coefficients_impl(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, int maxiterPCG, float tolPCG, int loco)
{
    arma::vec Y = d(Yvec), w = d(wVec), tau = d(tauVec); arma::mat X = d(Xmat);
    int p = X.n_cols; arma::uword N = X.n_rows;
    arma::vec SiY(N), alpha(p), eta(N); arma::mat SiX(N, p), cov(p, p);
    ck(sgb_get_coefficients(ctx(), Y.memptr(), X.memptr(), p, w.memptr(), tau.memptr(), maxiterPCG, tolPCG, loco,
                            SiY.memptr(), SiX.memptr(), cov.memptr(), alpha.memptr(), eta.memptr()));
    return Rcpp::List::create(Named("Sigma_iY") = SiY, Named("Sigma_iX") = SiX, Named("cov") = cov, Named("alpha") = alpha, Named("eta") = eta);
}
// [[Rcpp::export]]
Rcpp::List getCoefficients(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, int maxiterPCG, float tolPCG)
{ return coefficients_impl(Yvec, Xmat, wVec, tauVec, maxiterPCG, tolPCG, 0); }
// [[Rcpp::export]]
Rcpp::List getCoefficients_LOCO(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, int maxiterPCG, float tolPCG)
{ return coefficients_impl(Yvec, Xmat, wVec, tauVec, maxiterPCG, tolPCG, 1); }

// [[Rcpp::export]]
Rcpp::List getAIScore(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &Sigma_iY,
                      arma::fmat &Sigma_iX, arma::fmat &cov, int nrun, int maxiterPCG, float tolPCG, float traceCVcutoff)
{
    arma::vec Y = d(Yvec), w = d(wVec), tau = d(tauVec), SiY = d(Sigma_iY); arma::mat X = d(Xmat), SiX = d(Sigma_iX), cv = d(cov);
    arma::vec PY(Y.n_elem); double out4[4];
    set_seed(200);                                              // GetTrace, SAIGE_fitGLMM_fast.cpp:3114
    ck(sgb_get_ai_score(ctx(), Y.memptr(), X.memptr(), X.n_cols, w.memptr(), tau.memptr(), SiY.memptr(), SiX.memptr(),
                        cv.memptr(), nrun, maxiterPCG, tolPCG, traceCVcutoff, draw_probes, nullptr, out4, PY.memptr()));
    return Rcpp::List::create(Named("YPAPY") = out4[0], Named("Trace") = out4[1], Named("PY") = PY, Named("AI") = out4[2]);
}
// [[Rcpp::export]]
Rcpp::List getAIScore_q(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &Sigma_iY,
                        arma::fmat &Sigma_iX, arma::fmat &cov, int nrun, int maxiterPCG, float tolPCG, float traceCVcutoff)
{
    arma::vec Y = d(Yvec), w = d(wVec), tau = d(tauVec), SiY = d(Sigma_iY); arma::mat X = d(Xmat), SiX = d(Sigma_iX), cv = d(cov);
    arma::vec PY(Y.n_elem); double o[8];
    set_seed(200);                                              // GetTrace_q, SAIGE_fitGLMM_fast.cpp:3410
    ck(sgb_get_ai_score_q(ctx(), Y.memptr(), X.memptr(), X.n_cols, w.memptr(), tau.memptr(), SiY.memptr(), SiX.memptr(),
                          cv.memptr(), nrun, maxiterPCG, tolPCG, traceCVcutoff, draw_probes, nullptr, o, PY.memptr()));
    arma::vec Trace = {o[2], o[3]}; arma::mat AI = {{o[4], o[5]}, {o[5], o[6]}};
    return Rcpp::List::create(Named("YPAPY") = o[0], Named("YPA0PY") = o[1], Named("Trace") = Trace, Named("PY") = PY, Named("AI") = AI);
}
static Rcpp::List fit_impl(bool q, arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &Sigma_iY,
                           arma::fmat &Sigma_iX, arma::fmat &cov, int nrun, int maxiterPCG, float tolPCG, float tol, float traceCVcutoff)
{
    arma::vec Y = d(Yvec), w = d(wVec), tau = d(tauVec), SiY = d(Sigma_iY); arma::mat X = d(Xmat), SiX = d(Sigma_iX), cv = d(cov);
    set_seed(200);
    ck((q ? sgb_fit_glmmai_rpcg_q : sgb_fit_glmmai_rpcg)(ctx(), Y.memptr(), X.memptr(), X.n_cols, w.memptr(), tau.memptr(),
                                                          SiY.memptr(), SiX.memptr(), cv.memptr(), nrun, maxiterPCG, tolPCG, tol,
                                                          traceCVcutoff, draw_probes, nullptr));
    tauVec = arma::conv_to<arma::fvec>::from(tau);              // the reference updates tauVec in place too
    return Rcpp::List::create(Named("tau") = tau);
}
// [[Rcpp::export]]
Rcpp::List fitglmmaiRPCG(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &Sigma_iY,
                         arma::fmat &Sigma_iX, arma::fmat &cov, int nrun, int maxiterPCG, float tolPCG, float tol, float traceCVcutoff)
{ return fit_impl(false, Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, tol, traceCVcutoff); }
// [[Rcpp::export]]
Rcpp::List fitglmmaiRPCG_q(arma::fvec &Yvec, arma::fmat &Xmat, arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &Sigma_iY,
                           arma::fmat &Sigma_iX, arma::fmat &cov, int nrun, int maxiterPCG, float tolPCG, float tol, float traceCVcutoff)
{ return fit_impl(true, Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, tol, traceCVcutoff); }

// [[Rcpp::export]]
arma::fmat getSigma_X(arma::fvec &wVec, arma::fvec &tauVec, arma::fmat &Xmat, int maxiterPCG, float tolPCG)
{
    arma::vec w = d(wVec), tau = d(tauVec); arma::mat X = d(Xmat), out(X.n_rows, X.n_cols);
    ck(sgb_get_sigma_x(ctx(), w.memptr(), tau.memptr(), X.memptr(), X.n_cols, maxiterPCG, tolPCG, 0, out.memptr()));
    return arma::conv_to<arma::fmat>::from(out);
}
// [[Rcpp::export]]
arma::fvec getSigma_G(arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &Gvec, int maxiterPCG, float tolPCG)
{
    arma::vec w = d(wVec), tau = d(tauVec), G = d(Gvec), out(G.n_elem);
    ck(sgb_get_sigma_g(ctx(), w.memptr(), tau.memptr(), G.memptr(), 1, maxiterPCG, tolPCG, 0, out.memptr()));
    return arma::conv_to<arma::fvec>::from(out);
}
// [[Rcpp::export]]
float calCV(arma::fvec &xVec) { arma::vec x = d(xVec); return (float)sgb_cal_cv(x.memptr(), (int)x.n_elem); }
// [[Rcpp::export]]
float innerProduct(NumericVector x, NumericVector y) { return (float)sgb_inner_product(&x[0], &y[0], x.size()); }

// exported by the reference but reached only from C++ on this path; kept so that direct callers keep working
// [[Rcpp::export]]
arma::fvec getCrossprodMatAndKin(arma::fcolvec &bVec)
{
    arma::vec b = d(bVec), y(b.n_elem); ck(sgb_get_crossprod_mat_and_kin(ctx(), b.memptr(), 1, y.memptr()));
    return arma::conv_to<arma::fvec>::from(y);
}
// [[Rcpp::export]]
arma::fvec getPCG1ofSigmaAndVector(arma::fvec &wVec, arma::fvec &tauVec, arma::fvec &bVec, int maxiterPCG, float tolPCG)
{
    arma::vec w = d(wVec), tau = d(tauVec), b = d(bVec), x(b.n_elem); int32_t it = 0;
    ck(sgb_get_pcg1_of_sigma_and_vector(ctx(), w.memptr(), tau.memptr(), b.memptr(), 1, maxiterPCG, tolPCG, 0, x.memptr(), &it));
    Rcout << "iter from getPCG1ofSigmaAndVector " << it << std::endl;        // log line scraped downstream (FG.cpp:2798)
    return arma::conv_to<arma::fvec>::from(x);
}
// [[Rcpp::export]]
arma::fvec get_DiagofKin() { arma::vec x(sgb_get_nnomissing(ctx())); ck(sgb_get_diag_of_kin(ctx(), x.memptr())); return arma::conv_to<arma::fvec>::from(x); }

// ---------------------------------------------------------------------------------------------------------------------
// OPTIONAL exports (new names: they need `Rcpp::compileAttributes()` and the R-side patch of INTEGRATION.md "Driver loops
// inside the library").  The 28 exports above are the drop-in; these run whole R functions of R/SAIGE_fitGLMM_fast.R as one
// library call each, with every N-vector resident on the device between the solves.
// ---------------------------------------------------------------------------------------------------------------------
static int family_code(const std::string &f)
{
    if (f == "binomial") return 0;
    if (f == "gaussian") return 1;
    Rcpp::stop("saige_b200: family '" + f + "' is not built into the library (binomial, gaussian); use the R loop");
}

// Get_Coef / Get_Coef_LOCO (R/SAIGE_fitGLMM_fast.R:2-35, 42-73): body of the R function becomes
//   Get_Coef_b200(y, X, tau, family$family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, FALSE)
// [[Rcpp::export]]
Rcpp::List Get_Coef_b200(arma::vec &y, arma::mat &X, arma::vec &tau, std::string family, arma::vec &alpha0, arma::vec &eta0,
                         arma::vec &offset, int maxiterPCG, double tolPCG, int maxiter, bool isLOCO)
{
    const int p = X.n_cols; const arma::uword N = X.n_rows;
    arma::vec Y(N), alpha(p), eta(N), W(N), SiY(N), mu(N); arma::mat cov(p, p), SiX(N, p); int32_t nit = 0;
    ck(sgb_get_coef(ctx(), family_code(family), y.memptr(), X.memptr(), p, offset.memptr(), tau.memptr(), alpha0.memptr(),
                    eta0.memptr(), maxiter, maxiterPCG, tolPCG, isLOCO ? 1 : 0, Y.memptr(), alpha.memptr(), eta.memptr(), W.memptr(),
                    cov.memptr(), SiY.memptr(), SiX.memptr(), mu.memptr(), &nit));
    return Rcpp::List::create(Named("Y") = Y, Named("alpha") = alpha, Named("eta") = eta, Named("W") = W, Named("cov") = cov,
                              Named("sqrtW") = arma::sqrt(W), Named("Sigma_iY") = SiY, Named("Sigma_iX") = SiX, Named("mu") = mu);
}

// glmmkin.ai_PCG_Rcpp_Binary / _Quantitative between setgeno and the result list (R/SAIGE_fitGLMM_fast.R:127-292, 340-549):
// returns theta, coefficients, linear.predictors, fitted.values, Y, cov, converged and, with LOCO, N x 22 matrices
// (Y, linear.predictors, fitted.values), p x 22 coefficients and p*p x 22 covariances of the leave-one-chromosome-out refits;
// the R function keeps its own tail (ScoreTest_NULL_Model, Covariate_Transform_Back, the LOCOResult list).
// [[Rcpp::export]]
Rcpp::List glmmkin_ai_PCG_b200(bool isQuantitative, arma::vec &y, arma::mat &X, arma::vec &offset, arma::vec &alpha_fit0,
                               arma::vec &eta_fit0, arma::vec &tauInit, int maxiter, double tol, int nrun, double tolPCG,
                               int maxiterPCG, double traceCVcutoff, bool LOCO, int nChrom)
{
    const int p = X.n_cols; const arma::uword N = X.n_rows; const int nc = LOCO ? nChrom : 1;
    arma::vec tau(2), alpha(p), eta(N), mu(N), Y(N); arma::mat cov(p, p);
    arma::mat Yl(N, nc, arma::fill::zeros), el(N, nc, arma::fill::zeros), ml(N, nc, arma::fill::zeros), al(p, nc, arma::fill::zeros),
              cl(p * p, nc, arma::fill::zeros);
    std::vector<int32_t> nit(nc, 0); int32_t conv = 0, nouter = 0;
    sgb_set_probe_stream_fixed(ctx(), 1);                       // the first nrun probes of every estimate are the same vectors
    ck(sgb_glmmkin_ai_pcg(ctx(), isQuantitative ? 1 : 0, y.memptr(), X.memptr(), p, offset.memptr(), alpha_fit0.memptr(),
                          eta_fit0.memptr(), tauInit.memptr(), maxiter, tol, nrun, tolPCG, maxiterPCG, traceCVcutoff, LOCO ? 1 : 0,
                          draw_probes_reseeding, nullptr, tau.memptr(), alpha.memptr(), eta.memptr(), mu.memptr(), Y.memptr(),
                          cov.memptr(), &conv, &nouter, Yl.memptr(), al.memptr(), el.memptr(), cl.memptr(), ml.memptr(),
                          nit.data(), nullptr, nullptr));
    return Rcpp::List::create(Named("theta") = tau, Named("coefficients") = alpha, Named("linear.predictors") = eta,
                              Named("fitted.values") = mu, Named("Y") = Y, Named("cov") = cov, Named("converged") = conv != 0,
                              Named("LOCO.Y") = Yl, Named("LOCO.coefficients") = al, Named("LOCO.linear.predictors") = el,
                              Named("LOCO.cov") = cl, Named("LOCO.fitted.values") = ml);
}

// The marker loop of extractVarianceRatio (R/SAIGE_fitGLMM_fast.R:2298-2378) for the markers of one round (numMarkers, then +10):
// markerIdx 0-based into the GRM store (isVarRatioGeno = FALSE) or the hold-out store; returns var1, var2null, AC per marker.
// [[Rcpp::export]]
Rcpp::List varianceRatioMarkers_b200(Rcpp::IntegerVector markerIdx, bool isVarRatioGeno, arma::vec &W, arma::vec &tau, arma::mat &X,
                                     arma::mat &XV, arma::mat &XXVX_inv, arma::mat &Sigma_iX, arma::vec &mu2, bool isBinary,
                                     int maxiterPCG, double tolPCG)
{
    const int n = markerIdx.size();
    std::vector<int64_t> idx(markerIdx.begin(), markerIdx.end());
    arma::vec v1(n), v2(n), ac(n);
    ck(sgb_variance_ratio_markers(ctx(), idx.data(), n, isVarRatioGeno ? 1 : 0, W.memptr(), tau.memptr(), X.memptr(), X.n_cols,
                                  XV.memptr(), XXVX_inv.memptr(), Sigma_iX.memptr(), isBinary ? mu2.memptr() : nullptr, maxiterPCG,
                                  tolPCG, v1.memptr(), v2.memptr(), ac.memptr()));
    return Rcpp::List::create(Named("var1") = v1, Named("var2null") = v2, Named("AC") = ac);
}
#endif  // USE_SAIGE_B200
