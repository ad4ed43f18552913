// TEST INFRASTRUCTURE (oracle/): the reference's own solver layer -- getDiagOfSigma, getCrossprod, getPCG1ofSigmaAndVector,
// getCoefficients, GetTrace[_q], getAIScore[_q], fitglmmaiRPCG[_q], getSigma_X / _G, calCV and the _LOCO twins of
// /root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp -- compiled UNMODIFIED (oracle/_ref/fg_extract.inc, cut out at build
// time by extract_ref.py) behind a C interface for ctypes.  What the reference takes from the rest of its file is supplied
// here: the genotype object's diagonal and marker counts (`geno`), the GRM product (a callback, answered by the oracle's own
// product, which other artefacts pin), R's rbinom (a caller-supplied 0/1 stream, restarted by set_seed as GetTrace expects).
// Built twice by oracle/Makefile: REF_REAL=float is the reference as shipped (fp32); REF_REAL=double compiles the same text
// with `float` read as `double`, i.e. the reference's algorithm in the precision the oracle and the GPU library work in, so
// that iteration counts and results can be compared to ~1e-10 instead of fp32's ~1e-4.  Never linked by the product.
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <sys/time.h>
#include "mini_arma.h"

#ifndef REF_REAL
#define REF_REAL float
#endif
typedef REF_REAL real_t;
namespace arma {
typedef Col<real_t> fvec;
typedef Col<real_t> fcolvec;
typedef Mat<real_t> fmat;
typedef Col<long long> ivec;
}

namespace Rcpp {
struct NumericVector { std::vector<double> v; };
struct Value {
    int kind = 0; double s = 0; arma::fvec v; arma::fmat m;
    Value() {}
    Value(float x) : kind(1), s(x) {}
    Value(double x) : kind(1), s(x) {}
    Value(const arma::fvec &x) : kind(2), v(x) {}
    Value(const arma::fmat &x) : kind(3), m(x) {}
    operator float() const { return (float)s; }
    operator double() const { return s; }
    operator arma::fvec() const { return v; }
    operator arma::fmat() const { return m; }
};
struct Arg { std::string name; Value val; };
struct NamedProxy { std::string name; template <typename T> Arg operator=(const T &x) const { return Arg{name, Value(x)}; } };
inline NamedProxy Named(const char *n) { return NamedProxy{n}; }
struct List {
    std::map<std::string, Value> items;
    template <typename... A> static List create(const A &...a) { List l; const Arg arr[] = {a...}; for (const Arg &x : arr) l.items[x.name] = x.val; return l; }
    Value operator[](const char *k) const { return items.at(k); }
};
template <typename T> T as(const NumericVector &x)
{
    T r(x.v.size());
    for (size_t i = 0; i < x.v.size(); i++) r[i] = (real_t)x.v[i];
    return r;
}
}  // namespace Rcpp
using namespace Rcpp;
using namespace std;

// ---- what the extracted functions take from the rest of the reference file ----
typedef void (*fg_crossprod_fn)(const double *b, double *out, int n);
static fg_crossprod_fn g_cb = nullptr, g_cb_loco = nullptr;
// the reference prints its PCG iteration counts (and ingest chatter) to std::cout: this library's copy of the stream writes into a
// buffer instead (nothing else in a test process uses C++ iostreams)
static std::ostringstream ref_log;
static struct CoutRedirect { CoutRedirect() { std::cout.rdbuf(ref_log.rdbuf()); } } g_cout_redirect;
#if defined(REF_WITH_GENOCLASS)
// ---- libfg_refcpu.so: the reference's own genotype store and CPU product ----
// genoClass (FG.cpp:37-1183: PLINK reader, QC, best-guess imputation, re-pack, standardised genotypes, diagonals), the OpenMP
// marker loop parallelCrossProdOpenMP / parallelCrossProd[_full|_LOCO] (FG.cpp:1576-1851) and the exports that configure it, as
// built WITHOUT USE_GPU / USE_RcppParallel / USE_pbdMPI (the branch the reference takes on a CPU-only build).
#pragma omp declare reduction(+: arma::fvec: omp_out += omp_in) initializer(omp_priv = omp_orig)     // FG.cpp:31
float minMAFtoConstructGRM = 0;                                                                       // FG.cpp:35
// the class reports through printf as well ("M: ..., N: ...", FG.cpp:824): into the same buffer, a harness's stdout stays its own
static int ref_printf(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    ref_log << buf;
    return n;
}
#define printf(...) ref_printf(__VA_ARGS__)
#include "fg_extract_cpu.inc"
#undef printf
#else
static struct GenoStub {
    int N = 0, M = 0, Msub_in = 0, Msub_chr = 0;
    bool setKinDiagtoOne = false;
    arma::fvec diag, diag_loco;
    int getNnomissing() { return N; }
    int getM() { return M; }
    int getMsub() { return Msub_in; }
    int getnumberofMarkerswithMAFge_minMAFtoConstructGRM() { return M; }
    int getMsub_MAFge_minMAFtoConstructGRM_in() { return Msub_in; }
    int getMsub_MAFge_minMAFtoConstructGRM_singleChr_in() { return Msub_chr; }
    arma::fvec *Get_Diagof_StdGeno() { return &diag; }
    arma::fvec *Get_Diagof_StdGeno_LOCO() { return &diag_loco; }
} geno;
#endif
static bool isUsePrecondM = false, isUseSparseSigmaforInitTau = false, isUseSparseSigmaforModelFitting = false;   // FG.cpp:1888-1890
static arma::fvec gen_spsolve_v4(arma::fvec &, arma::fvec &, arma::fvec &) { throw std::logic_error("sparse-GRM path is not part of this build"); }
static double get_wall_time() { return 0.0; }
static double get_cpu_time() { return 0.0; }
static arma::fvec product(fg_crossprod_fn cb, arma::fcolvec &b)
{
    if (!cb) throw std::logic_error("no GRM product callback set");
    std::vector<double> in(b.n_elem), out(b.n_elem);
    for (arma::uword i = 0; i < b.n_elem; i++) in[i] = (double)b[i];
    cb(in.data(), out.data(), (int)b.n_elem);
    arma::fvec r(b.n_elem);
    for (arma::uword i = 0; i < b.n_elem; i++) r[i] = (real_t)out[i];
    return r;
}
#if defined(REF_WITH_GENOCLASS)
// the non-sparse, non-GPU branch of the dispatcher (FG.cpp:1953-1981, 1989-2000): straight to the marker loop
arma::fvec getCrossprodMatAndKin(arma::fcolvec &bVec) { return parallelCrossProd(bVec); }
arma::fvec getCrossprodMatAndKin_LOCO(arma::fcolvec &bVec) { return parallelCrossProd_LOCO(bVec); }
#else
arma::fvec getCrossprodMatAndKin(arma::fcolvec &bVec) { return product(g_cb, bVec); }            // FG.cpp:1953
arma::fvec getCrossprodMatAndKin_LOCO(arma::fcolvec &bVec) { return product(g_cb_loco, bVec); }  // FG.cpp:1989
#endif
// R's generator: a caller-supplied stream of rbinom(., 1, 0.5) draws; set_seed restarts it (GetTrace calls set_seed(200) first)
static const double *g_draws = nullptr;
static long g_ndraws = 0, g_cursor = 0;
static void set_seed(unsigned int) { g_cursor = 0; }
static NumericVector rbinom(int n, int, double)
{
    if (g_cursor + n > g_ndraws) throw std::runtime_error("probe stream exhausted");
    NumericVector r;
    r.v.assign(g_draws + g_cursor, g_draws + g_cursor + n);
    g_cursor += n;
    return r;
}
#if defined(REF_REAL_IS_DOUBLE)
#define float double
#endif
#include "fg_extract.inc"
#if defined(REF_REAL_IS_DOUBLE)
#undef float
#endif

// ---- C interface -------------------------------------------------------------------------------------------------------
static std::string g_err, g_log_copy;
static arma::fvec V(const double *p, long n) { arma::fvec r(n); for (long i = 0; i < n; i++) r[i] = (real_t)p[i]; return r; }
static arma::fmat M_(const double *p, long r, long c) { arma::fmat m(r, c); for (long i = 0; i < r * c; i++) m.d[i] = (real_t)p[i]; return m; }
static void out(double *dst, const arma::fvec &v) { for (arma::uword i = 0; i < v.n_elem; i++) dst[i] = (double)v[i]; }
static void out(double *dst, const arma::fmat &m) { for (arma::uword i = 0; i < m.n_elem; i++) dst[i] = (double)m.d[i]; }
#define GUARD(...) try { __VA_ARGS__; return 0; } catch (const std::exception &e) { g_err = e.what(); return 1; }

extern "C" {
int fgref_real_bytes() { return (int)sizeof(real_t); }
const char *fgref_last_error() { return g_err.c_str(); }
const char *fgref_log() { g_log_copy = ref_log.str(); return g_log_copy.c_str(); }
void fgref_clear_log() { ref_log.str(""); }
#if defined(REF_WITH_GENOCLASS)
#include <omp.h>
static int NN() { return geno.getNnomissing(); }
// the marker loop is `#pragma omp parallel for` (FG.cpp:1582); torchrun exports OMP_NUM_THREADS=1, so the harness sets the count
void fgref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int fgref_num_threads() { return omp_get_max_threads(); }
// setminMAFforGRM, setmaxMissingRateforGRM, setminMAC_VarianceRatio, then setgeno: the call sequence of FG.R:871-895
int fgref_setgeno(const char *bed, const char *bim, const char *fam, const int *subSampleInGeno, int nsub, const unsigned char *indicator,
                  int nfam, double minMAF, double maxMissing, int isVarRatio, double minMACvr, double maxMACvr, int kinDiagOne,
                  const long long *vr_rand_idx, int nvr)
{
    GUARD(setminMAFforGRM((float)minMAF); setmaxMissingRateforGRM((float)maxMissing);
          setminMAC_VarianceRatio((float)minMACvr, (float)maxMACvr, isVarRatio != 0);
          arma::randi_supply().assign(vr_rand_idx, vr_rand_idx + nvr);
          std::vector<int> sub(subSampleInGeno, subSampleInGeno + nsub);
          std::vector<bool> ind(indicator, indicator + nfam);
          setgeno(bed, bim, fam, sub, ind, 2.0f, kinDiagOne != 0))
}
long fgref_N() { return (long)geno.getNnomissing(); }
long fgref_M_qc() { return (long)geno.getnumberofMarkerswithMAFge_minMAFtoConstructGRM(); }
long fgref_M_raw() { return (long)geno.M; }
long fgref_M_vr() { return (long)geno.numberofMarkers_varRatio; }
void fgref_allele_freq(double *o) { for (arma::uword i = 0; i < geno.alleleFreqVec.n_elem; i++) o[i] = (double)geno.alleleFreqVec[i]; }
void fgref_inv_std(double *o) { for (arma::uword i = 0; i < geno.invstdvVec.n_elem; i++) o[i] = (double)geno.invstdvVec[i]; }
void fgref_mac(long long *o) { for (arma::uword i = 0; i < geno.MACVec.n_elem; i++) o[i] = geno.MACVec[i]; }
void fgref_qc_mask(unsigned char *o) { for (size_t i = 0; i < geno.MarkerswithMAFge_minMAFtoConstructGRM_indVec.size(); i++) o[i] = geno.MarkerswithMAFge_minMAFtoConstructGRM_indVec[i]; }
void fgref_vr_index(long long *o) { for (arma::uword i = 0; i < geno.markerIndexVec_forVarRatio.n_elem; i++) o[i] = geno.markerIndexVec_forVarRatio[i]; }
void fgref_vr_mac(long long *o) { for (arma::uword i = 0; i < geno.MACVec_forVarRatio.n_elem; i++) o[i] = geno.MACVec_forVarRatio[i]; }
int fgref_one_snp_geno(int idx, int vr, long long *o)
{
    GUARD(arma::ivec g = vr ? Get_OneSNP_Geno_forVarRatio(idx) : Get_OneSNP_Geno(idx); for (arma::uword i = 0; i < g.n_elem; i++) o[i] = g[i])
}
int fgref_one_snp_stdgeno(int idx, double *o)
{
    GUARD(arma::fvec v; geno.Get_OneSNP_StdGeno((size_t)idx, &v); out(o, v))
}
int fgref_diag_stdgeno(double *o) { GUARD(out(o, *geno.Get_Diagof_StdGeno())) }
int fgref_crossprod(const double *b, int loco, double *o)
{
    GUARD(arma::fcolvec bv = V(b, NN()); out(o, loco ? getCrossprodMatAndKin_LOCO(bv) : getCrossprodMatAndKin(bv)))
}
int fgref_set_start_end_index_vec(const long long *s, const long long *e, int n)
{
    GUARD(arma::ivec sv(n), ev(n); for (int i = 0; i < n; i++) { sv[i] = s[i]; ev[i] = e[i]; } setStartEndIndexVec(sv, ev))
}
int fgref_set_start_end_index(int s, int e, int c) { GUARD(setStartEndIndex(s, e, c)) }
int fgref_set_diag_loco() { GUARD(set_Diagof_StdGeno_LOCO()) }
#else
static int NN() { return geno.N; }
void fgref_set_problem(int N, int M, const double *diag_stdgeno, int kin_diag_one, fg_crossprod_fn cb)
{
    geno.N = N; geno.M = M; geno.diag = V(diag_stdgeno, N); geno.setKinDiagtoOne = kin_diag_one != 0; g_cb = cb;
}
void fgref_set_loco(const double *diag_loco, int Msub_in, int Msub_chr, fg_crossprod_fn cb_loco)
{
    geno.diag_loco = V(diag_loco, geno.N); geno.Msub_in = Msub_in; geno.Msub_chr = Msub_chr; g_cb_loco = cb_loco;
}
#endif
void fgref_set_draws(const double *u01, long n) { g_draws = u01; g_ndraws = n; g_cursor = 0; }
long fgref_draws_used() { return g_cursor; }
double fgref_cal_cv(const double *x, int n) { arma::fvec v = V(x, n); return (double)calCV(v); }
int fgref_diag_of_sigma(const double *w, const double *tau, int loco, double *o)
{
    GUARD(arma::fvec wv = V(w, NN()), tv = V(tau, 2); out(o, loco ? getDiagOfSigma_LOCO(wv, tv) : getDiagOfSigma(wv, tv)))
}
int fgref_pcg(const double *w, const double *tau, const double *b, int maxiterPCG, double tolPCG, int loco, double *x)
{
    GUARD(arma::fvec wv = V(w, NN()), tv = V(tau, 2), bv = V(b, NN());
          out(x, loco ? getPCG1ofSigmaAndVector_LOCO(wv, tv, bv, maxiterPCG, (real_t)tolPCG) : getPCG1ofSigmaAndVector(wv, tv, bv, maxiterPCG, (real_t)tolPCG)))
}
int fgref_get_coefficients(const double *Y, const double *X, int p, const double *w, const double *tau, int maxiterPCG, double tolPCG,
                           int loco, double *SiY, double *SiX, double *cov, double *alpha, double *eta)
{
    GUARD(arma::fvec Yv = V(Y, NN()), wv = V(w, NN()), tv = V(tau, 2); arma::fmat Xm = M_(X, NN(), p);
          List r = loco ? getCoefficients_LOCO(Yv, Xm, wv, tv, maxiterPCG, (real_t)tolPCG) : getCoefficients(Yv, Xm, wv, tv, maxiterPCG, (real_t)tolPCG);
          out(SiY, r.items.at("Sigma_iY").v); out(SiX, r.items.at("Sigma_iX").m); out(cov, r.items.at("cov").m);
          out(alpha, r.items.at("alpha").v); out(eta, r.items.at("eta").v))
}
// out8 = {YPAPY, YPA0PY, Trace0, Trace1, AI00, AI01, AI11, probes drawn}
int fgref_get_ai_score(int quant, const double *Y, const double *X, int p, const double *w, const double *tau, const double *SiY,
                       const double *SiX, const double *cov, int nrun, int maxiterPCG, double tolPCG, double traceCVcutoff,
                       double *out8, double *PY)
{
    GUARD(arma::fvec Yv = V(Y, NN()), wv = V(w, NN()), tv = V(tau, 2), SiYv = V(SiY, NN());
          arma::fmat Xm = M_(X, NN(), p), SiXm = M_(SiX, NN(), p), cv = M_(cov, p, p);
          for (int i = 0; i < 8; i++) out8[i] = 0.0;
          if (quant) {
              List r = getAIScore_q(Yv, Xm, wv, tv, SiYv, SiXm, cv, nrun, maxiterPCG, (real_t)tolPCG, (real_t)traceCVcutoff);
              out8[0] = r.items.at("YPAPY").s; out8[1] = r.items.at("YPA0PY").s;
              out8[2] = (double)r.items.at("Trace").v[0]; out8[3] = (double)r.items.at("Trace").v[1];
              const arma::fmat &AI = r.items.at("AI").m;
              out8[4] = (double)AI(0, 0); out8[5] = (double)AI(0, 1); out8[6] = (double)AI(1, 1);
              out(PY, r.items.at("PY").v);
          } else {
              List r = getAIScore(Yv, Xm, wv, tv, SiYv, SiXm, cv, nrun, maxiterPCG, (real_t)tolPCG, (real_t)traceCVcutoff);
              out8[0] = r.items.at("YPAPY").s; out8[3] = r.items.at("Trace").s; out8[6] = r.items.at("AI").s;
              out(PY, r.items.at("PY").v);
          }
          out8[7] = (double)g_cursor / (double)NN())
}
int fgref_fit_glmmai_rpcg(int quant, const double *Y, const double *X, int p, const double *w, double *tau_inout, const double *SiY,
                          const double *SiX, const double *cov, int nrun, int maxiterPCG, double tolPCG, double tol, double traceCVcutoff)
{
    GUARD(arma::fvec Yv = V(Y, NN()), wv = V(w, NN()), tv = V(tau_inout, 2), SiYv = V(SiY, NN());
          arma::fmat Xm = M_(X, NN(), p), SiXm = M_(SiX, NN(), p), cv = M_(cov, p, p);
          List r = quant ? fitglmmaiRPCG_q(Yv, Xm, wv, tv, SiYv, SiXm, cv, nrun, maxiterPCG, (real_t)tolPCG, (real_t)tol, (real_t)traceCVcutoff)
                         : fitglmmaiRPCG(Yv, Xm, wv, tv, SiYv, SiXm, cv, nrun, maxiterPCG, (real_t)tolPCG, (real_t)tol, (real_t)traceCVcutoff);
          out(tau_inout, r.items.at("tau").v))
}
int fgref_get_sigma_x(const double *w, const double *tau, const double *X, int p, int maxiterPCG, double tolPCG, double *o)
{
    GUARD(arma::fvec wv = V(w, NN()), tv = V(tau, 2); arma::fmat Xm = M_(X, NN(), p); out(o, getSigma_X(wv, tv, Xm, maxiterPCG, (real_t)tolPCG)))
}
int fgref_get_sigma_g(const double *w, const double *tau, const double *G, int maxiterPCG, double tolPCG, double *o)
{
    GUARD(arma::fvec wv = V(w, NN()), tv = V(tau, 2), Gv = V(G, NN()); out(o, getSigma_G(wv, tv, Gv, maxiterPCG, (real_t)tolPCG)))
}
}
