// Rcpp shim for STEP 2 (SPAGMMATtest): the reference-side binding of the single-variant score test + SPA (+ Firth) of
// libsaige_b200.so.  Companion of SAIGE_fitGLMM_fast_b200.cpp; same rules: drop into src/SAIGE/src/, compile with
// -DUSE_SAIGE_B200, fence the two reference definitions it replaces (setSAIGEobjInCPP, Main.cpp:770-838, and the PLINK
// branch of mainMarkerInCPP, Main.cpp:149-560) with `#if !defined(USE_SAIGE_B200)`.  Names, argument lists and return
// types are the reference's, so R/RcppExports.R, src/RcppExports.cpp and R/SAIGE_SPATest.R stay unchanged.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no R / Rcpp in the build container).  The C ABI it calls is exercised by
// tests/test_step2_golden.py through saige_gpu_b200/step2.py, which follows this file.
//
// What the library covers: PLINK input (raw 2-bit rows; best-guess imputation), BGEN / VCF input (dosage rows; best_guess /
// mean / minor imputation, zeroing of small dosages), full-GRM variance ratio (t_varRatio_null[0]), binary and
// quantitative traits, SPA / SPA_fast, Firth's effect size, the exact test for MAC <= MACCutoffforER (<= 10), categorical
// variance ratios.  Sparse-GRM variance, conditional analysis and the region tests keep the reference's code path: the
// shim refuses those option combinations instead of silently ignoring them.
#if defined(USE_SAIGE_B200)
#include <RcppArmadillo.h>
#include <string>
#include <vector>
#include "saige_b200.h"
// [[Rcpp::depends(RcppArmadillo)]]
using namespace Rcpp;

sgb_ctx *saige_b200_ctx();                       // the handle owned by SAIGE_fitGLMM_fast_b200.cpp
static void ck2(int rc) { if (rc) Rcpp::stop(std::string("saige_b200: ") + sgb_last_error(saige_b200_ctx())); }

// state the marker loop needs besides the library's (PlinkClass keeps the file; Main.cpp globals keep the cut-offs)
extern double g_marker_minMAF_cutoff, g_marker_minMAC_cutoff, g_missingRate_cutoff;      // Main.cpp:60-70
extern double g_MACCutoffforER;                  // Main.cpp:68, set by setAssocTest_GlobalVarsInCPP (Main.cpp:96-110) before the model
static std::vector<int32_t> g_pos_in_fam;        // PlinkClass::m_posSampleInPlink, filled by setPLINKobjInCPP

// [[Rcpp::export]]
void setSAIGEobjInCPP(arma::mat & t_XVX, arma::mat & t_XXVX_inv, arma::mat & t_XV, arma::mat & t_XVX_inv_XV,
                      arma::mat & t_Sigma_iXXSigma_iX, arma::mat & t_X, arma::vec & t_S_a, arma::vec & t_res, arma::vec & t_mu2,
                      arma::vec & t_mu, arma::vec & t_varRatio_sparse, arma::vec & t_varRatio_null,
                      arma::vec & t_cateVarRatioMinMACVecExclude, arma::vec & t_cateVarRatioMaxMACVecInclude, double t_SPA_Cutoff,
                      arma::vec & t_tauvec, std::string t_traitType, arma::vec & t_y, std::string t_impute_method,
                      bool t_flagSparseGRM, bool t_isFastTest, double t_pval_cutoff_for_fastTest, arma::umat & t_locationMat,
                      arma::vec & t_valueVec, int t_dimNum, bool t_isCondition, std::vector<uint32_t> & t_condition_genoIndex,
                      bool t_is_Firth_beta, double t_pCutoffforFirth, arma::vec & t_offset, arma::vec & t_resout)
{
    if (t_flagSparseGRM)
        Rcpp::stop("saige_b200: sparse-GRM variance is not provided by the B200 library; build without USE_SAIGE_B200 for it");
    // conditional analysis: assign_conditionMarkers_factors (Main.cpp:2002-2179) keeps its reference body up to the call of
    // assignConditionFactors, which in the B200 build forwards P2Mat, XXVX_inv^T P2Mat, VarInvMat and TstatVec to
    // sgb_step2_set_condition; the _c columns come back in rows 22..27 of the result table
    const int64_t N = (int64_t)t_X.n_rows;
    const int p = (int)t_X.n_cols;
    arma::mat XVX_inv_XV_t = t_XVX_inv_XV;       // N x p already (readInGLMM.R:60-75 stores XVX_inv_XV as N x p)
    ck2(sgb_step2_set_model(saige_b200_ctx(), N, p, t_traitType == "binary" ? 1 : 0, t_mu.memptr(), t_res.memptr(), t_mu2.memptr(),
                            t_y.memptr(), t_X.memptr(), t_XVX.memptr(), t_XXVX_inv.memptr(), XVX_inv_XV_t.memptr(), t_S_a.memptr(),
                            t_tauvec.memptr(), t_varRatio_null[0], t_SPA_Cutoff, g_pos_in_fam.data()));
    // se_from_fit = 0: this fork's source back-calculates the SE from the p-value (SAIGE_test.cpp:632)
    ck2(sgb_step2_set_firth(saige_b200_ctx(), t_is_Firth_beta ? 1 : 0, t_pCutoffforFirth, t_offset.n_elem == (arma::uword)N ? t_offset.memptr() : nullptr, 0));
    // exact test of rare variants (Main.cpp:408-422); t_resout is empty on this path (readInGLMM.R:123: no resampled residuals)
    ck2(sgb_step2_set_er(saige_b200_ctx(), g_MACCutoffforER));
    // one ratio per MAC category (assignVarianceRatio, SAIGE_test.cpp:801-833)
    if (t_varRatio_null.n_elem > 1)
        ck2(sgb_step2_set_variance_ratios(saige_b200_ctx(), (int)t_varRatio_null.n_elem, t_varRatio_null.memptr(),
                                          t_cateVarRatioMinMACVecExclude.memptr(), t_cateVarRatioMaxMACVecInclude.memptr()));
}

// The PLINK branch of mainMarkerInCPP (Main.cpp:149-560): one call per chunk of marker indices.  `readRawRows` stands for
// the seek + read of PlinkClass::getOneMarker (PLINK.cpp:164-300) without the decode: ceil(n_fam / 4) bytes per marker.
std::vector<uint8_t> plink_read_raw_rows(const std::vector<std::string> & t_genoIndex, int64_t & n_fam);   // PLINK.cpp side
// thin wrappers over the reference's own reader objects and globals (Main.cpp:60-75, 584-700)
int64_t saige_b200_model_n();
int saige_b200_impute_method();
void saige_b200_read_dosages(const std::string & t_genoType, std::vector<std::string> & prev, std::vector<std::string> & cur, int64_t j, double *dst);
extern double g_dosage_zerod_cutoff, g_dosage_zerod_MAC_cutoff;

// [[Rcpp::export]]
Rcpp::DataFrame mainMarkerInCPP(std::string & t_genoType, std::string & t_traitType, std::vector<std::string> & t_genoIndex_prev,
                                std::vector<std::string> & t_genoIndex, bool t_isMoreOutput, bool t_isImputation, bool t_isFirth)
{
    const int64_t q = (int64_t)t_genoIndex.size();
    arma::mat out(28, q);                          // column-major 28 x q == row-major q x 28 of the C ABI (rows 22..27: conditional results)
    // se_two_sided = 0: qnorm(p, upper tail) as in this fork's source (SAIGE_test.cpp:523-526)
    if (t_genoType == "plink") {
        int64_t n_fam = 0;
        std::vector<uint8_t> rows = plink_read_raw_rows(t_genoIndex, n_fam);
        ck2(sgb_step2_test_markers(saige_b200_ctx(), rows.data(), n_fam, q, g_marker_minMAF_cutoff, g_marker_minMAC_cutoff,
                                   g_missingRate_cutoff, 0, out.memptr()));
    } else {
        // bgen / vcf: the reference's readers (BgenClass::getOneMarker, VcfClass::getOneMarker) already deliver one dosage
        // vector per marker in model-sample order with -1 for missing; they are stacked and tested as one batch.  The model
        // was set with the identity sample map in this case (the readers did the matching).
        const int64_t n = saige_b200_model_n();
        arma::mat D(n, q);                         // column-major n x q == row-major q x n
        for (int64_t j = 0; j < q; j++) saige_b200_read_dosages(t_genoType, t_genoIndex_prev, t_genoIndex, j, D.colptr(j));   // Unified_getOneMarker
        ck2(sgb_step2_test_dosages(saige_b200_ctx(), D.memptr(), n, q, g_marker_minMAF_cutoff, g_marker_minMAC_cutoff,
                                   g_missingRate_cutoff, 0, saige_b200_impute_method() /* 1 best_guess, 2 mean, 3 minor */,
                                   g_dosage_zerod_cutoff, g_dosage_zerod_MAC_cutoff, out.memptr()));
    }
    // rows with out(0, j) == 0 were filtered (Main.cpp:296 `continue`); the others fill the vectors of Main.cpp:520-560
    std::vector<double> altCounts, altFreq, missingRate, Beta, seBeta, Tstat, varT, pval, pvalNA, AF_case, AF_ctrl;
    std::vector<bool> isSPA;
    std::vector<int> keep;
    for (int64_t j = 0; j < q; j++) {
        if (out(0, j) != 1.0) continue;
        keep.push_back((int)j);
        altCounts.push_back(out(1, j)); altFreq.push_back(out(2, j)); missingRate.push_back(out(3, j));
        Beta.push_back(out(4, j)); seBeta.push_back(out(5, j)); Tstat.push_back(out(6, j)); varT.push_back(out(7, j));
        pval.push_back(out(8, j)); pvalNA.push_back(out(9, j)); isSPA.push_back(out(10, j) != 0.0);
        AF_case.push_back(out(11, j)); AF_ctrl.push_back(out(12, j));
    }
    return Rcpp::DataFrame::create(Named("keep") = keep, Named("AC_Allele2") = altCounts, Named("AF_Allele2") = altFreq,
                                   Named("MissingRate") = missingRate, Named("BETA") = Beta, Named("SE") = seBeta,
                                   Named("Tstat") = Tstat, Named("var") = varT, Named("p.value") = pval,
                                   Named("p.value.NA") = pvalNA, Named("Is.SPA") = isSPA, Named("AF_case") = AF_case,
                                   Named("AF_ctrl") = AF_ctrl, Named("stringsAsFactors") = false);
}
#endif
