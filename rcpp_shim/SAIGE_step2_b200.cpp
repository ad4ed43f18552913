// Rcpp shim for STEP 2 (SPAGMMATtest): the reference-side binding of the single-variant score test + SPA (+ Firth, exact
// test, conditional analysis) of libsaige_b200.so.  Companion of SAIGE_fitGLMM_fast_b200.cpp; same rules: drop into
// src/SAIGE/src/, compile with -DUSE_SAIGE_B200, include it at the end of Main.cpp and fence the reference definitions it
// replaces with `#if !defined(USE_SAIGE_B200)`:
//     setSAIGEobjInCPP      Main.cpp:770-838     (the reference body stays for the sparse-GRM build; see below)
//     setPLINKobjInCPP      Main.cpp:715-728     (same body + the sample map handed to the library)
//     mainMarkerInCPP       Main.cpp:149-560     (PLINK / BGEN / VCF marker loop)
// Signatures are the reference's, CHARACTER FOR CHARACTER (mainMarkerInCPP is `void` with `bool &` flags and writes the
// result file itself through the reference's writeOutfile_single, Main.cpp:2437), so R/RcppExports.R, src/RcppExports.cpp
// and R/SAIGE_SPATest.R stay unchanged.
//
// One addition to a reference header is needed ("step 2"): PlinkClass gets two public members,
//     void readRawMarker(uint64_t gIndex, unsigned char *dst);   // fseek(3 + m_numBytesofEachMarker0 * gIndex) + fread, no decode
//     void markerInfo(uint64_t gIndex, std::string &chr, uint32_t &pd, std::string &marker, std::string &ref, std::string &alt);
//     uint64_t bytesPerMarker() const; uint32_t nFam() const; const std::vector<uint32_t> & posSampleInPlink() const;
// i.e. the first lines of PlinkClass::getOneMarker (PLINK.cpp:196-215: seek + read, then m_chr / m_pd / m_MarkerInPlink /
// m_ref / m_alt with the AlleleOrder swap of :288-296) without the per-sample decode the library does on the device.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no R / Rcpp in the build container).  The C ABI it calls is exercised by
// tests/test_step2_golden.py through saige_gpu_b200/step2.py, which follows this file.
//
// What the library covers: PLINK input (raw 2-bit rows; best-guess imputation), BGEN / VCF input (dosage rows; best_guess /
// mean / minor imputation, zeroing of small dosages), full-GRM variance ratio, binary and quantitative traits, SPA /
// SPA_fast incl. log-scale p-values, Firth's effect size, the exact test for MAC <= MACCutoffforER (<= 10), categorical
// variance ratios, conditional analysis.  Sparse-GRM variance and the region tests keep the reference's code path: the shim
// refuses those option combinations instead of silently ignoring them.
#if defined(USE_SAIGE_B200)
#include <RcppArmadillo.h>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#include "PLINK.hpp"
#include "saige_b200.h"
// [[Rcpp::depends(RcppArmadillo)]]
using namespace Rcpp;

sgb_ctx *saige_b200_ctx();                       // the handle owned by SAIGE_fitGLMM_fast_b200.cpp
static void ck2(int rc) { if (rc) Rcpp::stop(std::string("saige_b200: ") + sgb_last_error(saige_b200_ctx())); }

// This file is INCLUDED AT THE END OF Main.cpp (`#if defined(USE_SAIGE_B200)` / `#include "SAIGE_step2_b200.cpp"`), not compiled
// on its own: the reference keeps its reader objects and cut-offs as file-statics / globals of that translation unit
// (ptr_gPLINKobj Main.cpp:34, g_impute_method :44, g_marker_minMAF_cutoff / g_marker_minMAC_cutoff / g_missingRate_cutoff
// :60-70, g_dosage_zerod_cutoff / g_dosage_zerod_MAC_cutoff :61-62, g_MACCutoffforER :68) and defines Unified_getOneMarker
// (:585-700) and writeOutfile_single (:2437) there; all of them are used below as they are.

// model sample -> row of the genotype file.  NOT static: filled by setPLINKobjInCPP below (PLINK input) or left as the
// identity by setSAIGEobjInCPP (BGEN / VCF: the reference's readers deliver dosages in model order already).
std::vector<int32_t> g_saige_b200_pos_in_fam;
static bool g_isCondition = false;
static int64_t g_model_n = 0;

// [[Rcpp::export]]
void setPLINKobjInCPP(std::string t_bimFile, std::string t_famFile, std::string t_bedFile, std::vector<std::string> & t_SampleInModel,
                      std::string t_AlleleOrder)
{
    ptr_gPLINKobj = new PLINK::PlinkClass(t_bimFile, t_famFile, t_bedFile, t_AlleleOrder);        // Main.cpp:721-725, unchanged
    ptr_gPLINKobj->setPosSampleInPlink(t_SampleInModel);                                          // Main.cpp:726
    const std::vector<uint32_t> & pos = ptr_gPLINKobj->posSampleInPlink();                        // already 0-based (PLINK.cpp:135)
    g_saige_b200_pos_in_fam.assign(pos.begin(), pos.end());
    if (t_AlleleOrder != "alt-first")
        Rcpp::stop("saige_b200: PLINK input is tested alt-first (the reference's default for PLINK, Geno.R:159); "
                   "use AlleleOrder = 'alt-first' or build without USE_SAIGE_B200");
}

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
    const int64_t N = (int64_t)t_X.n_rows;
    const int p = (int)t_X.n_cols;
    g_model_n = N;
    g_isCondition = t_isCondition;
    if ((int64_t)g_saige_b200_pos_in_fam.size() != N) {          // BGEN / VCF input: identity map
        g_saige_b200_pos_in_fam.resize((size_t)N);
        for (int64_t i = 0; i < N; i++) g_saige_b200_pos_in_fam[(size_t)i] = (int32_t)i;
    }
    ck2(sgb_step2_set_model(saige_b200_ctx(), N, p, t_traitType == "binary" ? 1 : 0, t_mu.memptr(), t_res.memptr(), t_mu2.memptr(),
                            t_y.memptr(), t_X.memptr(), t_XVX.memptr(), t_XXVX_inv.memptr(), t_XVX_inv_XV.memptr() /* N x p, readInGLMM.R:60-75 */,
                            t_S_a.memptr(), t_tauvec.memptr(), t_varRatio_null[0], t_SPA_Cutoff, g_saige_b200_pos_in_fam.data()));
    // se_from_fit = 0: this fork's source back-calculates the SE from the p-value (SAIGE_test.cpp:632)
    ck2(sgb_step2_set_firth(saige_b200_ctx(), t_is_Firth_beta ? 1 : 0, t_pCutoffforFirth, t_offset.n_elem == (arma::uword)N ? t_offset.memptr() : nullptr, 0));
    // exact test of rare variants (Main.cpp:408-422); t_resout is empty on this path (readInGLMM.R:123: no resampled residuals)
    ck2(sgb_step2_set_er(saige_b200_ctx(), g_MACCutoffforER));
    // one ratio per MAC category (assignVarianceRatio, SAIGE_test.cpp:801-833)
    if (t_varRatio_null.n_elem > 1)
        ck2(sgb_step2_set_variance_ratios(saige_b200_ctx(), (int)t_varRatio_null.n_elem, t_varRatio_null.memptr(),
                                          t_cateVarRatioMinMACVecExclude.memptr(), t_cateVarRatioMaxMACVecInclude.memptr()));
    // conditional analysis: assign_conditionMarkers_factors (Main.cpp:2002-2179) keeps its reference body up to the call of
    // assignConditionFactors, which in the B200 build forwards P2Mat, XXVX_inv^T P2Mat, VarInvMat and TstatVec to
    // sgb_step2_set_condition; the _c columns come back in columns 22..27 of the result table
}

// the reference's p-value strings: "%.6E", or mantissa / exponent from the log when the value underflows (SAIGE_test.cpp:268-284)
static std::string pval_string(double p, double logp)
{
    char buf[100];
    if (std::isnan(p)) return "NA";
    if (p != 0.0 || !std::isfinite(logp)) { std::snprintf(buf, sizeof buf, "%.6E", p); return buf; }
    const double log10p = logp / std::log(10.0);
    int exponent = (int)std::floor(log10p);
    double fraction = std::pow(10.0, log10p - exponent);
    if (fraction >= 9.95) { fraction = 1; exponent++; }
    std::snprintf(buf, sizeof buf, "%.1fE%d", fraction, exponent);
    return buf;
}

// [[Rcpp::export]]
void mainMarkerInCPP(std::string & t_genoType,     // "plink", "bgen", "vcf"
                     std::string & t_traitType, std::vector<std::string> & t_genoIndex_prev, std::vector<std::string> & t_genoIndex,
                     bool & t_isMoreOutput, bool & t_isImputation, bool & t_isFirth)
{
    const int q = (int)t_genoIndex.size();
    // the reference's output vectors (Main.cpp:160-205)
    std::vector<std::string> markerVec(q), chrVec(q), posVec(q), refVec(q), altVec(q);
    std::vector<double> altFreqVec(q), altCountsVec(q), imputationInfoVec(q, 1.0), missingRateVec(q);
    std::vector<double> BetaVec(q, arma::datum::nan), seBetaVec(q, arma::datum::nan), TstatVec(q, arma::datum::nan), varTVec(q, arma::datum::nan);
    std::vector<std::string> pvalVec(q, "NA"), pvalNAVec(q, "NA"), pval_cVec(q, "NA"), pvalNA_cVec(q, "NA");
    std::vector<double> Beta_cVec(q, arma::datum::nan), seBeta_cVec(q, arma::datum::nan), Tstat_cVec(q, arma::datum::nan), varT_cVec(q, arma::datum::nan);
    std::vector<bool> isSPAConvergeVec(q);
    std::vector<double> AF_caseVec(q), AF_ctrlVec(q), N_case_homVec(q), N_ctrl_hetVec(q), N_case_hetVec(q), N_ctrl_homVec(q);
    std::vector<uint32_t> N_caseVec(q), N_ctrlVec(q), N_Vec(q);

    arma::mat out(32, q);                          // column-major 32 x q == row-major q x 32 of the C ABI
    // se_two_sided = 0: qnorm(p, upper tail) as in this fork's source (SAIGE_test.cpp:523-526)
    if (t_genoType == "plink") {
        const uint64_t B0 = ptr_gPLINKobj->bytesPerMarker();
        std::vector<unsigned char> rows((size_t)B0 * (size_t)q);
        for (int i = 0; i < q; i++) {
            const uint64_t gIndex = std::strtoull(t_genoIndex[i].c_str(), nullptr, 10);          // Main.cpp:236-240
            ptr_gPLINKobj->readRawMarker(gIndex, rows.data() + (size_t)i * B0);
            uint32_t pd;
            ptr_gPLINKobj->markerInfo(gIndex, chrVec[i], pd, markerVec[i], refVec[i], altVec[i]);
            posVec[i] = std::to_string(pd);
        }
        ck2(sgb_step2_test_markers(saige_b200_ctx(), rows.data(), (int64_t)ptr_gPLINKobj->nFam(), q, g_marker_minMAF_cutoff,
                                   g_marker_minMAC_cutoff, g_missingRate_cutoff, 0, out.memptr()));
    } else {
        // bgen / vcf: the reference's readers deliver one dosage vector per marker in model-sample order with -1 for missing
        // (Unified_getOneMarker, Main.cpp:584-700); they are stacked and tested as one batch
        const int64_t n = g_model_n;
        arma::mat D(n, q);                         // column-major n x q == row-major q x n
        arma::vec gvec(n);
        std::vector<uint> idxMissing, idxNonZero;
        bool outMissing = true, onlyNonZero = false;
        for (int i = 0; i < q; i++) {
            uint64_t gIndex = std::strtoull(t_genoIndex[i].c_str(), nullptr, 10), gIndex_prev = std::strtoull(t_genoIndex_prev[i].c_str(), nullptr, 10);
            uint32_t pd; double af, ac, mr, info;
            idxMissing.clear(); idxNonZero.clear();
            bool ok = Unified_getOneMarker(t_genoType, gIndex_prev, gIndex, refVec[i], altVec[i], markerVec[i], pd, chrVec[i], af, ac, mr, info,
                                           outMissing, idxMissing, onlyNonZero, idxNonZero, gvec, t_isImputation);
            posVec[i] = std::to_string(pd);
            imputationInfoVec[i] = info;
            if (!ok) gvec.fill(-1.0);              // unreadable record: every call missing -> filtered by the missing-rate cut-off
            for (uint m : idxMissing) gvec[m] = -1.0;
            std::copy(gvec.begin(), gvec.end(), D.colptr(i));
        }
        const int impute = g_impute_method == "best_guess" ? 1 : (g_impute_method == "mean" ? 2 : 3);
        ck2(sgb_step2_test_dosages(saige_b200_ctx(), D.memptr(), n, q, g_marker_minMAF_cutoff, g_marker_minMAC_cutoff, g_missingRate_cutoff, 0,
                                   impute, g_dosage_zerod_cutoff, g_dosage_zerod_MAC_cutoff, out.memptr()));
    }
    // rows with out(0, i) == 0 were filtered (Main.cpp:296 `continue`): their pvalVec stays "NA" and writeOutfile_single skips them
    int mFirth = 0, mFirthConverge = 0;
    for (int i = 0; i < q; i++) {
        altCountsVec[i] = out(1, i); altFreqVec[i] = out(2, i); missingRateVec[i] = out(3, i);
        if (out(0, i) != 1.0) continue;
        BetaVec[i] = out(4, i); seBetaVec[i] = out(5, i); TstatVec[i] = out(6, i); varTVec[i] = out(7, i);
        pvalVec[i] = pval_string(out(8, i), out(28, i)); pvalNAVec[i] = pval_string(out(9, i), out(29, i));
        isSPAConvergeVec[i] = out(10, i) != 0.0;
        AF_caseVec[i] = out(11, i); AF_ctrlVec[i] = out(12, i);
        N_caseVec[i] = (uint32_t)out(13, i); N_ctrlVec[i] = (uint32_t)out(14, i);
        N_case_homVec[i] = out(15, i); N_case_hetVec[i] = out(16, i); N_ctrl_homVec[i] = out(17, i); N_ctrl_hetVec[i] = out(18, i);
        N_Vec[i] = (uint32_t)g_model_n;
        mFirth += out(20, i) != 0.0; mFirthConverge += out(21, i) != 0.0;
        if (g_isCondition) {
            Beta_cVec[i] = out(22, i); seBeta_cVec[i] = out(23, i); Tstat_cVec[i] = out(24, i); varT_cVec[i] = out(25, i);
            pval_cVec[i] = pval_string(out(26, i), out(30, i)); pvalNA_cVec[i] = pval_string(out(27, i), out(31, i));
        }
    }
    writeOutfile_single(t_isMoreOutput, t_isImputation, g_isCondition, t_isFirth, mFirth, mFirthConverge, t_traitType, chrVec, posVec, markerVec,
                        refVec, altVec, altCountsVec, altFreqVec, imputationInfoVec, missingRateVec, BetaVec, seBetaVec, TstatVec, varTVec,
                        pvalVec, pvalNAVec, isSPAConvergeVec, Beta_cVec, seBeta_cVec, Tstat_cVec, varT_cVec, pval_cVec, pvalNA_cVec, AF_caseVec,
                        AF_ctrlVec, N_caseVec, N_ctrlVec, N_case_homVec, N_ctrl_hetVec, N_case_hetVec, N_ctrl_homVec, N_Vec);
}
#endif
