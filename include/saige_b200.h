/*
 * saige_b200.h -- C ABI of libsaige_b200.so, the B200 (sm_100a) CUDA back end for SAIGE step 1.
 *
 * Every entry point below is what the reference's FFI layer would bind for the null-GLMM hot path:
 * the bodies of the [[Rcpp::export]] functions in /root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp
 * ("FG.cpp") forward to these (see INTEGRATION.md and rcpp_shim/).  Each declaration cites the reference
 * function it replaces.
 *
 * Conventions
 *   - all entry points return 0 on success, non-zero on failure; sgb_last_error() gives the message.
 *     Nothing in the library calls exit() (the reference does, FG.cpp:1947).
 *   - all floating-point data at the boundary is fp64, column-major, in CALLER-OWNED HOST buffers
 *     (R numeric vectors/matrices).  The library copies in and out; it never keeps caller pointers.
 *   - one context == one GPU.  Multi-GPU runs are SPMD, one process (rank) per GPU, like the reference's
 *     `mpirun -n G Rscript ...` (FG.cpp:1905-1913); markers are sharded over ranks and every GRM product
 *     ends in one NCCL sum-allreduce (replaces MPI_Allreduce, FG.cpp:1619,1652).
 *   - the library contains no RNG on the parity path: Rademacher probes and the variance-ratio marker
 *     index set are drawn by the caller (R's RNG: FG.cpp:3052-3054, 866-868) and passed in.
 *   - there is no CPU fallback: every compute entry fails if the CUDA device is unavailable.
 */
#ifndef SAIGE_B200_H
#define SAIGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgb_ctx sgb_ctx;

#define SGB_NCCL_ID_BYTES 128

/* Which arithmetic engine computes GRM products.
 *   SGB_ENGINE_TENSOR : 2-bit genotypes decoded in registers to u8, right-hand sides split into signed int8 limbs of a
 *                       55-bit fixed-point value (8 x 7 bits on the mma.sync kernel, 7 x 8 bits on the tcgen05 kernel),
 *                       exact int32 accumulation on the tensor cores, exact recombination, one rounding (default).  Batches of
 *                       k >= 3 columns run on the tcgen05 kernel (A operand decoded straight into tensor memory,
 *                       accumulator in TMEM); k <= 2 on the HBM-bound mma.sync kernel.  Both give identical bits.
 *   SGB_ENGINE_F64    : plain fp64 FMA kernels (slow; the on-device cross-check of the tensor engine).
 *   SGB_ENGINE_UMMA   : force the tcgen05 kernel for every k >= 2;  SGB_ENGINE_IMMA : force the mma.sync kernel. */
enum { SGB_ENGINE_TENSOR = 0, SGB_ENGINE_F64 = 1, SGB_ENGINE_UMMA = 2, SGB_ENGINE_IMMA = 3 };

/* ---- life cycle -------------------------------------------------------------------------------- */
/* Replaces the file-global `genoClass geno` (FG.cpp:1188) + gpuSymMatMult (gpuSymMatMult.hpp:13-36). */
int sgb_create(int device, sgb_ctx **out);
/* Multi-GPU: rank/world + a 128-byte NCCL unique id made by rank 0 with sgb_nccl_unique_id and
 * distributed by the launcher (replaces MPI_Comm_rank/size, FG.cpp:1905-1906). */
int sgb_nccl_unique_id(void *id128);
int sgb_create_dist(int device, int rank, int world, const void *id128, sgb_ctx **out);
void sgb_destroy(sgb_ctx *h);                       /* closeGenoFile_plink (FG.cpp:1191-1209) + ~gpuSymMatMult */
const char *sgb_last_error(sgb_ctx *h);             /* h may be NULL: error of the last failed sgb_create* */
int sgb_set_engine(sgb_ctx *h, int engine);
/* verbose != 0: every PCG solve prints the reference's stdout lines, one per right-hand side in the order the reference's
 * sequential solves would ("iter from getPCG1ofSigmaAndVector <n>", "pcg did not converge. You may increase maxiter
 * number." -- FG.cpp:2794-2798), so log scrapers written for the reference keep working.  Off by default. */
int sgb_set_verbose(sgb_ctx *h, int verbose);
/* Tolerance-driven precision of WIDE batches (k >= 3 columns, the tensor-bound tcgen05 kernel): every right-hand-side value
 * is the fixed-point integer round(v 2^(8 n - 3 - E)) (E = exponent of the column maximum) written with n signed 8-bit
 * digits, so the tensor work of a batch is proportional to n.
 *   n = 7 (default): 55-bit values -- the same integers as the k <= 2 kernel, results exact up to one rounding of the input
 *                    (2^-54 of the column maximum), bit-identical across engines;
 *   n = 6 / 5      : 2^-46 / 2^-38 of the column maximum per input element (measured on GRM products: <= 1e-13 / <= 4e-11 of
 *                    the largest output), inside the 1e-10 gate of the GRM product and far inside the PCG tolerance.
 * k <= 2 products (HBM-bound) always carry the full 55 bits.  rel_tol variant: the smallest n with 2^-(8n-2) * 16 <= rel_tol. */
int sgb_set_rhs_limbs(sgb_ctx *h, int n);
int sgb_set_product_tolerance(sgb_ctx *h, double rel_tol);
int sgb_device_sync(sgb_ctx *h);

/* ---- configuration set through exports before setgeno ------------------------------------------- */
int sgb_set_min_maf_for_grm(sgb_ctx *h, float minMAF);             /* setminMAFforGRM (FG.cpp:4923) */
int sgb_set_max_missing_rate_for_grm(sgb_ctx *h, float maxMiss);   /* setmaxMissingRateforGRM (FG.cpp:4928) */
int sgb_set_min_mac_variance_ratio(sgb_ctx *h, float minMAC, float maxMAC, int isVarRatioGeno); /* FG.cpp:4970 */

/* ---- genotype store ----------------------------------------------------------------------------- */
/* setgeno -> genoClass::setGenoObj (FG.cpp:2267, 739-1024): reads <bed,bim,fam>, QC (MAF / missing rate in
 * fp32 with the reference's expression order, FG.cpp:438-493), best-guess imputation (FG.cpp:447,556-558),
 * variance-ratio hold-out (FG.cpp:496-548; vr_rand_idx = g_randMarkerIndforVR drawn by the caller), re-pack
 * in phenotype-sample order (FG.cpp:551-576), then builds the device layout.
 *   subSampleInGeno[n_sub] : 1-based .fam row of each phenotyped sample
 *   indicator[n_fam]       : 1 if the .fam sample has a phenotype */
int sgb_setgeno(sgb_ctx *h, const char *bedfile, const char *bimfile, const char *famfile,
                const int32_t *subSampleInGeno, int64_t n_sub, const uint8_t *indicator, int64_t n_fam,
                int isDiagofKinSetAsOne, const int32_t *vr_rand_idx, int64_t n_vr_idx);
/* Same, from a host-resident .bed body (no magic bytes), n_fam samples x n_bim markers, SNP-major. */
int sgb_setgeno_mem(sgb_ctx *h, const uint8_t *bed_body, int64_t n_fam, int64_t n_bim,
                    const int32_t *subSampleInGeno, int64_t n_sub, const uint8_t *indicator,
                    int isDiagofKinSetAsOne, const int32_t *vr_rand_idx, int64_t n_vr_idx);
/* Bench / test input: raw PLINK .bed rows (no magic bytes) of synthetic markers [m0, m1) of the same generator, written to
 * the host buffer `out` ((m1 - m0) * ceil(N/4) bytes); t0 / t1: thresholds of those markers.  Not part of the reference. */
int sgb_synth_bed_rows(sgb_ctx *h, int64_t N, int64_t m0, int64_t m1, uint64_t seed, const uint32_t *t0, const uint32_t *t1,
                       double miss_rate, uint8_t *out);
/* Bench/test workload generator: counter-based synthetic genotypes written straight into the
 * device layout; t0/t1 are per-marker uint32 thresholds.  Bit-identical to oracle/saige_oracle.c:orc_synth_bed. */
int sgb_setgeno_synth(sgb_ctx *h, int64_t n_samples, int64_t n_markers, uint64_t seed,
                      const uint32_t *t0, const uint32_t *t1);

int64_t sgb_get_total_marker(sgb_ctx *h);           /* gettotalMarker: .bim lines (FG.cpp:1213) */
int64_t sgb_get_num_qc_markers(sgb_ctx *h);         /* getMsub_MAFge_minMAFtoConstructGRM (FG.cpp:1271) */
int64_t sgb_get_num_local_markers(sgb_ctx *h);      /* this rank's shard (gpuDistributeSNPs n_cols, FG.cpp:1913) */
int64_t sgb_get_nnomissing(sgb_ctx *h);             /* getNnomissingOut (FG.cpp:1266) */
int64_t sgb_get_num_vr_markers(sgb_ctx *h);
int sgb_get_allele_freq_vec(sgb_ctx *h, double *out);        /* getAlleleFreqVec, float values widened (FG.cpp:1219) */
int sgb_get_mac_vec(sgb_ctx *h, int32_t *out);               /* getMACVec (FG.cpp:1224) */
int sgb_get_allele_count_vec(sgb_ctx *h, int32_t *out);      /* integer A1 count after imputation (bit-exact gate) */
int sgb_get_mac_vec_for_var_ratio(sgb_ctx *h, int32_t *out); /* getMACVec_forVarRatio (FG.cpp:1230) */
int sgb_get_index_vec_for_var_ratio(sgb_ctx *h, int32_t *out);/* getIndexVec_forVarRatio (FG.cpp:1235) */
int sgb_get_is_var_ratio_geno(sgb_ctx *h);                   /* getIsVarRatioGeno (FG.cpp:1240) */
int sgb_get_qcd_marker_index(sgb_ctx *h, uint8_t *out);      /* getQCdMarkerIndex, n_bim flags (FG.cpp:1249) */
int sgb_get_one_snp_geno(sgb_ctx *h, int64_t snp_idx, int32_t *out);             /* Get_OneSNP_Geno (FG.cpp:2281) */
int sgb_get_one_snp_geno_for_var_ratio(sgb_ctx *h, int64_t snp_idx, int32_t *out); /* FG.cpp:2291 */
int sgb_get_one_snp_stdgeno(sgb_ctx *h, int64_t snp_idx, double *out);           /* Get_OneSNP_StdGeno (FG.cpp:2302) */

/* ---- LOCO bookkeeping --------------------------------------------------------------------------- */
int sgb_set_start_end_index_vec(sgb_ctx *h, const int32_t *start, const int32_t *end, int n); /* FG.cpp:3084 */
int sgb_set_start_end_index(sgb_ctx *h, int start, int end, int chromIndex);                  /* FG.cpp:3067 */
int sgb_set_diag_of_stdgeno_loco(sgb_ctx *h);                                                 /* FG.cpp:4934 */

/* ---- GRM products ------------------------------------------------------------------------------- */
int sgb_get_diag_of_kin(sgb_ctx *h, double *out);                       /* get_DiagofKin (FG.cpp:4358) */
/* getCrossprodMatAndKin[_LOCO] (FG.cpp:1953,1989): Y[N x k] = K.B, K = (1/M) sum_m z_m z_m^T; k >= 1 columns. */
int sgb_get_crossprod_mat_and_kin(sgb_ctx *h, const double *B, int k, double *Y);
int sgb_get_crossprod_mat_and_kin_loco(sgb_ctx *h, const double *B, int k, double *Y);
int sgb_get_diag_of_sigma(sgb_ctx *h, const double *w, const double *tau, int loco, double *out); /* FG.cpp:2322,2368 */
int sgb_get_crossprod(sgb_ctx *h, const double *B, int k, const double *w, const double *tau, int loco,
                      double *Y);                                       /* getCrossprod[_LOCO] (FG.cpp:2397,2430) */

/* ---- PCG ---------------------------------------------------------------------------------------- */
/* getPCG1ofSigmaAndVector[_LOCO] (FG.cpp:2593,2943) for k right-hand sides at once.  Each column follows
 * exactly the sequential recurrence and stops on its own ||r||^2 <= tolPCG (absolute, FG.cpp:2696);
 * iters_out[k] (may be NULL) receives the per-column iteration counts the reference prints (FG.cpp:2798). */
int sgb_get_pcg1_of_sigma_and_vector(sgb_ctx *h, const double *w, const double *tau, const double *B, int k,
                                     int maxiterPCG, double tolPCG, int loco, double *X, int32_t *iters_out);

/* ---- AI-REML ------------------------------------------------------------------------------------ */
/* Probe source: fill out[N x count] (column-major) with the next `count` Rademacher vectors of the caller's
 * RNG stream, i.e. 2*rbinom(N,1,0.5)-1 (FG.cpp:3134-3137).  Return 0 on success. */
typedef int (*sgb_probe_fn)(void *user, int64_t n, int count, double *out);

/* getCoefficients[_LOCO] (FG.cpp:3167,3203): Sigma_iY[N], Sigma_iX[N x p], cov[p x p], alpha[p], eta[N]. */
int sgb_get_coefficients(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, const double *tau,
                         int maxiterPCG, double tolPCG, int loco,
                         double *Sigma_iY, double *Sigma_iX, double *cov, double *alpha, double *eta);
/* getAIScore (FG.cpp:3279): out4 = {YPAPY, Trace, AI, nrun actually used}; PY[N]. */
int sgb_get_ai_score(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, const double *tau,
                     const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun,
                     int maxiterPCG, double tolPCG, double traceCVcutoff, sgb_probe_fn probes, void *user,
                     double *out4, double *PY);
/* getAIScore_q (FG.cpp:3479): out8 = {YPAPY, YPA0PY, Trace0, Trace1, AI00, AI01, AI11, nrun used}. */
int sgb_get_ai_score_q(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, const double *tau,
                       const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun,
                       int maxiterPCG, double tolPCG, double traceCVcutoff, sgb_probe_fn probes, void *user,
                       double *out8, double *PY);
/* fitglmmaiRPCG / fitglmmaiRPCG_q (FG.cpp:3302,3610): tau_inout[2] updated in place. */
int sgb_fit_glmmai_rpcg(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, double *tau_inout,
                        const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun,
                        int maxiterPCG, double tolPCG, double tol, double traceCVcutoff,
                        sgb_probe_fn probes, void *user);
int sgb_fit_glmmai_rpcg_q(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, double *tau_inout,
                          const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun,
                          int maxiterPCG, double tolPCG, double tol, double traceCVcutoff,
                          sgb_probe_fn probes, void *user);
/* getSigma_X[_LOCO] / getSigma_G[_LOCO] (FG.cpp:3347-3405): k columns solved as one batch. */
int sgb_get_sigma_x(sgb_ctx *h, const double *w, const double *tau, const double *X, int p,
                    int maxiterPCG, double tolPCG, int loco, double *Sigma_iX);
int sgb_get_sigma_g(sgb_ctx *h, const double *w, const double *tau, const double *Gmat, int k,
                    int maxiterPCG, double tolPCG, int loco, double *Sigma_iG);
/* ---- driver loops of SAIGE_fitGLMM_fast.R ("FG.R") as single calls (optional: the exports above stay the drop-in) ----
 * The R functions below call the exports above in a loop and do O(N) vector algebra in between (IRLS working vector,
 * covariate adjustment of a marker); with 8 GPUs that host-side algebra and the N-vector round trips cost more than the
 * products.  These entry points run the same loop with every N-vector resident on the device; an R driver that wants them
 * replaces the body of the named R function by one .Call, nothing else changes.
 *
 * Get_Coef / Get_Coef_LOCO (FG.R:2-35, 42-73).  family 0 = binomial(logit) with R's own clamping (logit_linkinv /
 * logit_mu_eta of stats/src/family.c: |eta| > 30), 1 = gaussian(identity).  eta0 and the returned eta INCLUDE the offset.
 * Outputs as the R list: Y, W, eta, mu, Sigma_iY (N), Sigma_iX (N x p), alpha (p), cov (p x p); n_iter = getCoefficients
 * calls made (<= maxiter).  Any output pointer except alpha / cov may be NULL. */
int sgb_get_coef(sgb_ctx *h, int family, const double *y, const double *X, int p, const double *offset, const double *tau,
                 const double *alpha0, const double *eta0, int maxiter, int maxiterPCG, double tolPCG, int loco,
                 double *Y, double *alpha, double *eta, double *W, double *cov, double *Sigma_iY, double *Sigma_iX,
                 double *mu, int32_t *n_iter);
/* The leave-one-chromosome-out refit loop (FG.R:255-292): for every chromosome c of setStartEndIndexVec that has a range,
 * setStartEndIndex(start_c, end_c, c) and Get_Coef_LOCO starting from the previous chromosome's (alpha, eta).  Column / slot c
 * of Y, eta, mu (N x nchr), alpha (p x nchr), cov (p*p x nchr), n_iter (nchr) is written for those chromosomes only.
 * Requires sgb_set_diag_of_stdgeno_loco. */
/* on_chrom (may be NULL) is called from the calling thread as soon as the outputs of chromosome `chrom` are complete in the
 * caller's arrays (chrom = -1: the genome-wide fit of sgb_glmmkin_ai_pcg), so the caller can start its own per-chromosome work
 * (ScoreTest_NULL_Model, FG.R:283-288) on another thread while the next refit runs on the GPU. */
typedef void (*sgb_chrom_done_fn)(void *user, int chrom);
int sgb_get_coef_loco_all(sgb_ctx *h, int family, const double *y, const double *X, int p, const double *offset,
                          const double *tau, const double *alpha0, const double *eta0, int maxiter, int maxiterPCG,
                          double tolPCG, double *Y, double *alpha, double *eta, double *cov, double *mu, int32_t *n_iter,
                          sgb_chrom_done_fn on_chrom, void *chrom_user);
/* glmmkin.ai_PCG_Rcpp_Binary / glmmkin.ai_PCG_Rcpp_Quantitative after setgeno (FG.R:127-304, 340-549) as ONE call: the first
 * Get_Coef + getAIScore[_q], the outer loop of Get_Coef + fitglmmaiRPCG[_q] with its three stopping rules, the final Get_Coef
 * and, with loco != 0, set_Diagof_StdGeno_LOCO and the refit loop of sgb_get_coef_loco_all.  quantitative = 0: binomial family,
 * tau[0] fixed at 1; 1: gaussian.  alpha_fit0 / eta_fit0: coefficients and linear predictors of the glm fit0 (eta includes the
 * offset); tauInit[2] as the R argument.  Outputs: tau_out[2], alpha (p), eta, mu, Y (N), cov (p x p), converged (i < maxiter),
 * n_outer (iterations of the outer loop); the *_loco outputs are laid out as in sgb_get_coef_loco_all and may be NULL when
 * loco = 0.  Probes come from `probes` as in sgb_get_ai_score (sgb_set_probe_stream_fixed applies); because this call holds
 * several trace estimates, `probes(user, N, 0, NULL)` is called before each one: the callback restarts its stream there, as
 * GetTrace's set_seed(200) does (FG.cpp:3114). */
int sgb_glmmkin_ai_pcg(sgb_ctx *h, int quantitative, const double *y, const double *X, int p, const double *offset,
                       const double *alpha_fit0, const double *eta_fit0, const double *tauInit, int maxiter, double tol,
                       int nrun, double tolPCG, int maxiterPCG, double traceCVcutoff, int loco, sgb_probe_fn probes,
                       void *user, double *tau_out, double *alpha, double *eta, double *mu, double *Y, double *cov,
                       int32_t *converged, int32_t *n_outer, double *Y_loco, double *alpha_loco, double *eta_loco,
                       double *cov_loco, double *mu_loco, int32_t *n_iter_loco, sgb_chrom_done_fn on_chrom, void *chrom_user);
/* The marker loop of extractVarianceRatio (FG.R:2298-2378) for a batch of markers: Get_OneSNP_Geno[_forVarRatio], flip to
 * the minor allele, AC, G = G0 - XXVX_inv (XV G0), Sigma^-1 G as one multi-column solve, and
 *   var1[j] = (G' Sigma_iG - G' Sigma_iX (X' Sigma_iX)^-1 X' Sigma_iG) / AC,   var2null[j] = sum mu2 g^2 (g = G / sqrt(AC);
 * mu2 = NULL: sum g^2, the quantitative trait).  marker_idx: 0-based indices into the GRM store (from_vr_store = 0) or the
 * variance-ratio hold-out store (1).  XV is p x N, XXVX_inv and Sigma_iX N x p, all column-major as R holds them. */
int sgb_variance_ratio_markers(sgb_ctx *h, const int64_t *marker_idx, int nmark, int from_vr_store, const double *w,
                               const double *tau, const double *X, int p, const double *XV, const double *XXVX_inv,
                               const double *Sigma_iX, const double *mu2, int maxiterPCG, double tolPCG, double *var1,
                               double *var2null, double *AC);
/* GetTrace re-seeds R's generator to 200 before drawing its probes (FG.cpp:3114), so the first nrun probe vectors are the same
 * in every call for a given N.  on = 1 lets getAIScore / fitglmmaiRPCG reuse the device-resident copy of that first batch (and
 * its cached K.U) without calling `probes` for it; the callback still serves the +10 retry batches, after being asked once for
 * the skipped columns so that its stream position is right.  Default 0: every batch comes from the callback. */
int sgb_set_probe_stream_fixed(sgb_ctx *h, int on);
double sgb_cal_cv(const double *x, int n);                                   /* calCV (FG.cpp:3104) */
double sgb_inner_product(const double *x, const double *y, int64_t n);      /* innerProduct */

/* ---- step 2: single-variant score test + SPA, batched over variants ------------------- */
/* Null-model state of SAIGEClass (setSAIGEobjInCPP, SAIGE_test.cpp:30-120; fields chosen by ReadModel,
 * R/readInGLMM.R:39-170): N model samples, p covariate columns; X, XXVX_inv, XVX_inv_XV are N x p column-major,
 * XVX p x p; pos_in_fam[i] = 0-based .fam row of model sample i (PlinkClass::m_posSampleInPlink). */
int sgb_step2_set_model(sgb_ctx *h, int64_t N, int p, int binary, const double *mu, const double *res, const double *mu2,
                        const double *y, const double *X, const double *XVX, const double *XXVX_inv,
                        const double *XVX_inv_XV, const double *S_a, const double *tau, double varRatio, double SPAcutoff,
                        const int32_t *pos_in_fam);
/* mainMarkerInCPP loop body (Main.cpp:229-520) for n_markers raw PLINK rows (ceil(n_fam/4) bytes each, A1 = ALT):
 * getOneMarker -> filter -> imputeGenoAndFlip (best_guess) -> scoreTestFast -> SPA / SPA_fast (SAIGE_test.cpp:212-292,
 * 345-640).  out[n_markers x 32] row-major: tested(0/1), AC_Allele2, AF_Allele2, MissingRate, BETA, SE, Tstat, var,
 * p.value, p.value.NA, Is.SPA, AF_case, AF_ctrl, N_case, N_ctrl, N_case_hom, N_case_het, N_ctrl_hom, N_ctrl_het, var2,
 * Is.Firth, Firth converged, BETA_c, SE_c, Tstat_c, var_c, p.value_c, p.value.NA_c (NaN without sgb_step2_set_condition),
 * then the NATURAL LOGS of p.value, p.value.NA, p.value_c, p.value.NA_c: finite where the p-value itself underflows a double
 * (stat > ~1490).  There the reference switches to log-scale p-values and prints them as "%.1fE%d" strings
 * (SAIGE_test.cpp:255-284, 531-582); the log columns carry that case: the saddle-point branch, the standard errors and the
 * Firth cut-off use them, and the table writer prints mantissa / exponent from them when p.value == 0.
 * se_two_sided = 1: SE of SPA-adjusted variants = |BETA| / |qnorm(p/2)| (matches the reference's bundled golden tables);
 * 0: |BETA| / qnorm(p, upper) as this fork's source has it (SAIGE_test.cpp:523-526). */
/* is_Firth_beta / pCutoffforFirth (SAIGE_test.cpp:573-633; fast_logistf_fit_simple :893-986): variants of a binary trait
 * whose final p-value is <= p_cutoff get Firth's bias-reduced effect size from a two-column penalised logistic refit
 * [1, gtilde] with the null model's offset (N doubles, NULL = zeros; readInGLMM.R:99-160).  se_from_fit = 1: SE is the
 * refit's own standard error (what the reference's bundled positive-signal result holds); 0: |BETA| / |qnorm| of the
 * p-value as this fork's source computes it (:632).  Off after sgb_step2_set_model. */
int sgb_step2_set_firth(sgb_ctx *h, int enable, double p_cutoff, const double *offset, int se_from_fit);
/* max_MAC_for_ER (g_MACCutoffforER of setAssocTest_GlobalVarsInCPP, Main.cpp:68-110; R default 4): on a binary trait a
 * variant whose minor allele count after imputation is <= max_mac_for_er and whose |T|/sqrt(var) exceeds SPAcutoff gets
 * its p-value from the exact test over all case / control assignments of its carriers instead of the saddle-point
 * approximation (Main.cpp:408-422, SAIGE_test.cpp:426-431, 592-620; SKATExactBin_Work, ER_binary_func.cpp:186-278), and
 * SE = |BETA| / |qnorm(p/2)|; Is.SPA stays 0.  Negative: off (the state after sgb_step2_set_model).  Values above 10 are
 * refused: the reference sizes the test for MAC <= 10 (ER_binary_func.cpp:26) and its Monte-Carlo regime is not built. */
int sgb_step2_set_er(sgb_ctx *h, double max_mac_for_er);
/* Conditional analysis (--condition; assign_conditionMarkers_factors, Main.cpp:2002-2179, and the t_isCondition block of
 * getMarkerPval, SAIGE_test.cpp:640-790): with gtilde_k the covariate-adjusted genotype of conditioning marker k and v_k its
 * variance ratio, P2[N x n_cond] (column-major) = sqrt(v_k) gtilde_k % mu2 * tau0, VarInv = pinv(P1 P2) with P1 row k =
 * sqrt(v_k) gtilde_k^T, Tstat_cond[k] the marker's score, XtP2[p x n_cond] = XXVX_inv^T P2.  Every tested variant then also
 * reports BETA_c, SE_c, Tstat_c, var_c, p.value_c, p.value.NA_c (out columns 22..27): T_c = T - G1P2 VarInv Tstat_cond,
 * var_c = var - G1P2 VarInv G1P2^T, G1P2 = sqrt(v) gtilde^T P2, and the saddle-point approximation on T_c when
 * T_c^2 / var_c > SPAcutoff^2.  The bundled conditional golden table reports the adjusted p-value itself (this fork's
 * source prints half of it, SAIGE_test.cpp:752): the fixture wins.  n_cond = 0 switches it off; at most 4 markers. */
int sgb_step2_set_condition(sgb_ctx *h, int n_cond, const double *P2, const double *XtP2, const double *VarInv,
                            const double *Tstat_cond);
/* Categorical variance ratios (t_varRatio_null with more than one entry, t_cateVarRatioMinMACVecExclude,
 * t_cateVarRatioMaxMACVecInclude of setSAIGEobjInCPP; SAIGEClass::assignVarianceRatio, SAIGE_test.cpp:801-833): a variant
 * with min_mac_exclude[c] < MAC <= max_mac_include[c] (MAC after imputation) uses ratios[c]; max_mac_include has n_cate - 1
 * entries, the last category is open-ended, MAC below the first bound uses ratios[0].  The categories must tile the MAC axis
 * (min_mac_exclude[c] == max_mac_include[c-1]), which is what fitNULLGLMM's defaults (10, 20.5) / (20.5) do.  One deviation,
 * deliberate: MAC == min_mac_exclude[0] exactly matches no branch of the reference, which then keeps the previous marker's
 * ratio (order-dependent); here it belongs to the first category.  n_cate == 1: the single ratio of sgb_step2_set_model. */
int sgb_step2_set_variance_ratios(sgb_ctx *h, int n_cate, const double *ratios, const double *min_mac_exclude,
                                  const double *max_mac_include);
/* enable != 0 (default): the score sums of a chunk of variants are computed as ONE skinny GEMM on the tensor engine (packed
 * 2-bit variant rows x the 2p+3 model columns, exact integer accumulation) and only the variants that need a pass of their
 * own (saddle-point approximation, exact test, Firth, conditional analysis) run the per-variant kernel.  0: every variant
 * through the per-variant kernel (the round-1 path, kept as the on-device cross-check). */
int sgb_step2_set_batched(sgb_ctx *h, int enable);
/* Raw-row bytes per pipeline chunk of sgb_step2_test_markers (default 1 GB: a chunk must hold ~10^4 variants for the flagged
 * variants to fill the machine; smaller values bound the device / pinned staging memory, and let tests cross chunk borders). */
int sgb_step2_set_chunk_bytes(sgb_ctx *h, int64_t bytes);
int sgb_step2_test_markers(sgb_ctx *h, const uint8_t *bed_rows, int64_t n_fam, int64_t n_markers, double min_maf,
                           double min_mac, double max_missing, int se_two_sided, double *out);
/* The same marker loop for dosage rows (Unified_getOneMarker's VCF / BGEN branches, Main.cpp:584-700; VCF.cpp:120-256,
 * BGEN.cpp:132-345): dosages[n_markers x n_file_samples] row-major doubles = copies of the tested allele per sample of the
 * genotype file (hard calls as 0 / 1 / 2), negative or NaN = missing; pos_in_fam of sgb_step2_set_model indexes these rows.
 * impute_method 1 best_guess / 2 mean / 3 minor, dosage_zerod_cutoff / dosage_zerod_mac_cutoff as in imputeGenoAndFlip
 * (UTIL.cpp:58-135; R defaults 0.2 and 10).  N_*_hom / N_*_het count dosages in [1.5, 2] / [0.5, 1.5) (Main.cpp:510-525).
 * Same out table as sgb_step2_test_markers.  The imputation-INFO filter of BGEN input (minInfo) is applied by the caller
 * before this call, from the scores sgb_bgen_read returns. */
int sgb_step2_test_dosages(sgb_ctx *h, const double *dosages, int64_t n_file_samples, int64_t n_markers, double min_maf,
                           double min_mac, double max_missing, int se_two_sided, int impute_method,
                           double dosage_zerod_cutoff, double dosage_zerod_mac_cutoff, double *out);

/* ---- BGEN input for step 2 (host side) ------------------------------------------------------------------------- */
/* BgenClass::setBgenObj / getOneMarker / Parse2 (BGEN.cpp:25-130, 360-520, 132-345): BGEN v1.2 layout 2, zlib or plain
 * blocks, unphased diploid biallelic variants, 8- or 16-bit probabilities.  sgb_bgen_read decodes the next <= max_variants
 * variants into dosages[n x n_samples] (row-major; copies of the tested allele: the first allele when alt_first, else the
 * second, which is the reader's ref-first default; -1 = missing), ready for sgb_step2_test_dosages, and writes one
 * "CHR\tPOS\tID\tREF\tALT\n" line per variant into info_buf.  Blocks are read sequentially and inflated / decoded by
 * n_threads host threads.  info_scores (n doubles or NULL): the imputation INFO score of each variant over the samples flagged
 * in in_model (n_samples bytes; NULL = all) that are not missing (BGEN.cpp:275-345), for the minInfo filter of imputed data.
 * Errors: sgb_last_error(NULL).  No device work: these entry points run without a GPU. */
typedef struct sgb_bgen sgb_bgen;
int sgb_bgen_open(const char *path, sgb_bgen **out, int64_t *n_samples, int64_t *n_variants, int *has_sample_ids);
int sgb_bgen_sample_id(sgb_bgen *b, int64_t i, char *buf, int buflen);
int sgb_bgen_read(sgb_bgen *b, int64_t max_variants, int alt_first, int n_threads, const uint8_t *in_model,
                  double *dosages, double *info_scores, char *info_buf, int64_t info_len, int64_t *n_read);
void sgb_bgen_close(sgb_bgen *b);

/* ---- dense N x N GRM (BASELINE config 4) ---------------------------------------------------- */
/* The reference fork ships no code for this step (docs/overview.md:19-21 describe a "full GRM" option of SAIGE-GPU; the
 * GCTA-style files extdata/output/nfam_*_GRM.grm.bin give the output format).  K_ij = (1/M) sum_m z_mi z_mj with the same
 * standardisation as getCrossprodMatAndKin (FG.cpp:1445-1502), built on the tcgen05 int8 tensor cores from the 2-bit
 * store: weights s_m^2 as `weight_limbs` (2..8) base-128 digits, exact int32 accumulation, fp64 storage of the lower
 * block-trapezoid sharded by 128-sample block-rows across ranks (collective call when created with sgb_create_dist).
 * weight_limbs = 7 keeps the stored matrix within ~1e-13 of fp64 arithmetic; the build costs weight_limbs passes. */
int sgb_dense_grm_build(sgb_ctx *h, int weight_limbs);
int sgb_dense_grm_free(sgb_ctx *h);
/* out[ni x nj] column-major = K[i0 .. i0+ni) [j0 .. j0+nj)  (collective when distributed) */
int sgb_dense_grm_get_block(sgb_ctx *h, int64_t i0, int64_t ni, int64_t j0, int64_t nj, double *out);
/* out6 = {weight_limbs, fixed-point exponent S, block-rows, stored bytes on this rank, build ms, int8 ops issued} */
int sgb_dense_grm_info(sgb_ctx *h, double *out6);
/* Which GRM getCrossprodMatAndKin / PCG / AI-REML use: the packed genotypes (default) or the stored dense matrix
 * ("PCG on stored GRM"; no LOCO in that mode). */
enum { SGB_GRM_PACKED = 0, SGB_GRM_DENSE = 1 };
int sgb_set_grm_mode(sgb_ctx *h, int mode);

/* ---- device-resident benchmark hooks (bench.py `value` leg: inputs already in HBM) ---------------- */
/* Runs `reps` k-column GRM products on device-resident synthetic right-hand sides, timing with CUDA events
 * on the library's stream; ms_out[reps] per-product times, and ms_kernel_out[2*reps] (may be NULL) the device time of
 * the two genotype sweeps {sweep 1, sweep 2} of each product.  Result left in an internal buffer (sgb_bench_fetch_result). */
int sgb_bench_crossprod_device(sgb_ctx *h, int k, int reps, uint64_t seed, float *ms_out, float *ms_kernel_out);
int sgb_bench_fetch_result(sgb_ctx *h, int k, double *Y, double *B);
/* Builds only block-rows [first_block_row, first_block_row + n_block_rows) of the dense GRM (a bounded sample of the
 * build for bench.py; sgb_dense_grm_info reports time and operations). */
int sgb_bench_dense_build(sgb_ctx *h, int weight_limbs, int64_t first_block_row, int64_t n_block_rows);

/* ---- counters (SURVEY.md section 5 "metrics") ----------------------------------------------------- */
typedef struct {
    int64_t n_crossprod_calls;      /* multi-vector GRM products */
    int64_t n_crossprod_columns;    /* sum of k over them */
    int64_t n_pcg_solves;           /* columns solved */
    int64_t n_pcg_iterations;       /* sum over columns */
    int64_t n_kernel_launches;      /* kernels of this library launched */
    int64_t n_allreduce;            /* NCCL allreduces issued */
    int64_t bytes_h2d, bytes_d2h;
    int64_t n_probe_product_reuse;  /* getAIScore calls that reused the cached K.U of the Hutchinson probes */
    int64_t n_probe_batches_resident; /* first probe batches taken from the device copy (sgb_set_probe_stream_fixed) */
} sgb_counters;
int sgb_get_counters(sgb_ctx *h, sgb_counters *out);
int sgb_reset_counters(sgb_ctx *h);

#ifdef __cplusplus
}
#endif
#endif /* SAIGE_B200_H */
