/*
 * saige_oracle.c -- CPU restatement of the SAIGE step-1 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA library in saige_gpu_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never links or calls it.
 *
 * It restates, function by function, what /root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp
 * ("Fg->cpp") does on the packed-genotype store and the GRM.vector product.  The reference itself
 * cannot be built here (needs R, Rcpp, RcppArmadillo, RcppParallel, MPI, boost -- Fg->cpp:4-19).
 *
 * Parity pins (tests/test_oracle_golden.py):
 *   - decode + allele counts + MAF filter  <-> extdata/input/plinkforGRM_1000samples_10kMarkers.frq (exact)
 *   - fp32 reference-order matvec vs fp64 matvec (informational, ~1e-6)
 *   - everything above the matvec (PCG / AI-REML) is restated in numpy (oracle/oracle.py) and pinned there to the
 *     reference's own functions, compiled unmodified into oracle/_ref/libfg_ref{64,32}.so (tests/test_reference_solver.py).
 *
 * Two arithmetic modes for the matvec:
 *   mode 0 ("fp64")   : f = AC/(2N) and s = 1/sqrt(2f(1-f)) evaluated in double from the integer allele
 *                       count; accumulation in double.  This is what the GPU is graded against.
 *   mode 1 ("ref32")  : float f / invStd exactly as the reference stores them (Fg->cpp:484,916-921),
 *                       3-entry float LUT (Fg->cpp:152-159), float dot + float axpy in marker order
 *                       (Fg->cpp:1576-1598), thread-private float accumulators under OpenMP.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t N0;          /* samples in .fam                      (geno.N,  Fg->cpp:771)  */
    int64_t M0;          /* markers in .bim                      (geno.M,  Fg->cpp:786)  */
    int64_t N;           /* phenotyped samples                   (Nnomissing, Fg->cpp:744) */
    int64_t M;           /* markers passing QC                   (numberofMarkerswithMAFge_minMAFtoConstructGRM) */
    int64_t Mvr;         /* markers held out for variance ratio  (numberofMarkers_varRatio) */
    int64_t B;           /* bytes per re-packed marker = ceil(N/4) (m_size_of_esi, Fg->cpp:214) */
    uint8_t *geno;       /* M  x B, PLINK 2-bit codes, no missing (genoVecofPointers)          */
    uint8_t *geno_vr;    /* Mvr x B                              (genoVecofPointers_forVarRatio) */
    float *afreq;        /* alleleFreqVec   (float, Fg->cpp:484,923) */
    float *invstd;       /* invstdvVec      (float, Fg->cpp:916-922) */
    int32_t *mac;        /* MACVec          (Fg->cpp:490,926) */
    int32_t *ac;         /* allele count after imputation (not kept by the reference; = afreq*2N exactly) */
    float *afreq_vr, *invstd_vr;
    int32_t *mac_vr, *index_vr, *ac_vr;
    uint8_t *qc_mask;    /* M0 flags: MarkerswithMAFge_minMAFtoConstructGRM_indVec (Fg->cpp:927,930) */
} orc_geno;

orc_geno *orc_new(void) { return (orc_geno *)calloc(1, sizeof(orc_geno)); }

static void orc_clear(orc_geno *g)
{
    free(g->geno); free(g->geno_vr); free(g->afreq); free(g->invstd); free(g->mac); free(g->ac);
    free(g->afreq_vr); free(g->invstd_vr); free(g->mac_vr); free(g->index_vr); free(g->ac_vr); free(g->qc_mask);
    memset(g, 0, sizeof(*g));
}

void orc_free(orc_geno *g) { if (g) { orc_clear(g); free(g); } }

/* PLINK .bed 2-bit code -> bufferGeno of Get_OneSNP_Geno_atBeginning (Fg->cpp:343-354):
 * b = low bit, a = high bit; (b=1,a=0)->3 missing; (0,0)->2; (b=0,a=1)->1; (1,1)->0. */
static inline int bed_code_to_geno(int code)
{
    int b = code & 1, a = (code >> 1) & 1;
    if (b == 1 && a == 0) return 3;
    if (b == 0 && a == 0) return 2;
    if (b == 0 && a == 1) return 1;
    return 0;
}

/* setGenotype with HOM_ALT=0x3 (geno 0), HET=0x2 (geno 1), HOM_REF=0x0 (geno 2): Fg->cpp:41-44,203-205,559-565 */
static inline uint8_t geno_to_code(int g)
{
    return g == 0 ? 0x3 : (g == 1 ? 0x2 : 0x0);
}

/*
 * setGenoObj + Get_OneSNP_Geno_atBeginning (Fg->cpp:739-1024, 326-579) on an in-memory .bed body
 * (`bed` points just past the 3 magic bytes; marker i starts at bed + i*ceil(N0/4), Fg->cpp:902).
 *   sub_idx[N]   : subSampleInGeno, 1-based row of each phenotyped sample in the .fam (Fg->cpp:555)
 *   indicator[N0]: indicatorGenoSamplesWithPheno (Fg->cpp:376)
 *   vr_idx[n_vr] : g_randMarkerIndforVR -- drawn by the CALLER (the reference uses arma::randi, Fg->cpp:866-868)
 * Returns 0, or -1 on allocation failure.
 */
int orc_setgeno(orc_geno *g, const uint8_t *bed, int64_t N0, int64_t M0, const int32_t *sub_idx, int64_t N,
                const uint8_t *indicator, float minMAF, float maxMissing, int isVarRatio,
                float minMACvr, float maxMACvr, const int32_t *vr_idx, int64_t n_vr)
{
    orc_clear(g);
    g->N0 = N0; g->M0 = M0; g->N = N; g->B = (N + 3) / 4;
    int64_t B0 = (N0 + 3) / 4;
    g->geno = (uint8_t *)malloc((size_t)(M0 > 0 ? M0 : 1) * g->B);
    g->afreq = (float *)malloc(sizeof(float) * (M0 + 1));
    g->invstd = (float *)malloc(sizeof(float) * (M0 + 1));
    g->mac = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1));
    g->ac = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1));
    g->qc_mask = (uint8_t *)calloc(M0 + 1, 1);
    int64_t vr_cap = isVarRatio ? M0 : 0;
    g->geno_vr = (uint8_t *)malloc((size_t)(vr_cap > 0 ? vr_cap : 1) * g->B);
    g->afreq_vr = (float *)malloc(sizeof(float) * (vr_cap + 1));
    g->invstd_vr = (float *)malloc(sizeof(float) * (vr_cap + 1));
    g->mac_vr = (int32_t *)malloc(sizeof(int32_t) * (vr_cap + 1));
    g->ac_vr = (int32_t *)malloc(sizeof(int32_t) * (vr_cap + 1));
    g->index_vr = (int32_t *)malloc(sizeof(int32_t) * (vr_cap + 1));
    uint8_t *invr = (uint8_t *)calloc(M0 + 1, 1);
    /* per raw marker: decision (bit 0 = GRM, bit 1 = variance-ratio store), imputation fill, statistics.  The marker
     * loop of the reference is sequential (Fg->cpp:897-953); its iterations are independent except for the running
     * output indices, so it is restated as: statistics of every marker (parallel), prefix count, re-pack (parallel). */
    uint8_t *dec = (uint8_t *)calloc(M0 + 1, 1);
    int32_t *fillv = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1)), *acv = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1));
    int32_t *macv = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1));
    float *fv = (float *)malloc(sizeof(float) * (M0 + 1)), *isv = (float *)malloc(sizeof(float) * (M0 + 1));
    int64_t *dstq = (int64_t *)malloc(sizeof(int64_t) * (M0 + 1));
    if (!g->geno || !g->afreq || !g->invstd || !g->mac || !g->ac || !g->qc_mask || !g->geno_vr || !invr || !dec || !fillv ||
        !acv || !macv || !fv || !isv || !dstq) return -1;
    for (int64_t j = 0; j < n_vr; j++)
        if (vr_idx[j] >= 0 && vr_idx[j] < M0) invr[vr_idx[j]] = 1;

#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M0; m++) {
        const uint8_t *row = bed + m * B0;
        int alleleCount = 0, numMissing = 0;
        for (int64_t i = 0; i < N0; i++) {
            int gv = bed_code_to_geno((row[i >> 2] >> ((i & 3) << 1)) & 3);
            if (indicator[i]) { if (gv == 3) numMissing++; else alleleCount += gv; }
        }
        /* Fg->cpp:438-447 -- float arithmetic, in this order */
        float altFreq = alleleCount / (float)((N - numMissing) * 2);
        float missingRate = numMissing / (float)N;
        int fill = (int)roundf(2 * altFreq);              /* int(round(2*altFreq)) Fg->cpp:447 */
        if (numMissing > 0) alleleCount += fill * numMissing;   /* Fg->cpp:454 */
        altFreq = alleleCount / (float)(N * 2);           /* Fg->cpp:484 */
        float maf = altFreq < 1 - altFreq ? altFreq : 1 - altFreq;   /* Fg->cpp:489 */
        int mac = alleleCount < (int)N * 2 - alleleCount ? alleleCount : (int)N * 2 - alleleCount;
        int passQC = (maf >= minMAF && missingRate <= maxMissing);   /* Fg->cpp:493 */
        int passVR = 0;
        if (isVarRatio) {                                   /* Fg->cpp:496-548 */
            if (maxMACvr != -1) {
                if (mac >= minMACvr && mac < maxMACvr) passVR = 1;
                else if (mac >= maxMACvr) passVR = invr[m];
            } else {
                if (mac >= minMACvr) passVR = invr[m];
            }
            if (passVR) passQC = 0;
        }
        float Std = sqrtf(2 * altFreq * (1 - altFreq));    /* Fg->cpp:916-921 */
        dec[m] = (uint8_t)((passQC ? 1 : 0) | ((isVarRatio && passVR) ? 2 : 0));
        fillv[m] = fill; acv[m] = alleleCount; macv[m] = mac; fv[m] = altFreq; isv[m] = (Std == 0) ? 0.f : 1 / Std;
    }
    int64_t Mq = 0, Mv = 0;
    for (int64_t m = 0; m < M0; m++) {
        dstq[m] = -1;
        if (dec[m] & 1) {
            g->afreq[Mq] = fv[m]; g->invstd[Mq] = isv[m]; g->mac[Mq] = macv[m]; g->ac[Mq] = acv[m];
            g->qc_mask[m] = 1; dstq[m] = Mq++;
        }
        if (dec[m] & 2) {
            g->afreq_vr[Mv] = fv[m]; g->invstd_vr[Mv] = isv[m]; g->mac_vr[Mv] = macv[m]; g->ac_vr[Mv] = acv[m];
            g->index_vr[Mv] = (int32_t)m; dstq[m] = Mv++;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M0; m++) {
        if (!dec[m]) continue;                              /* Fg->cpp:551-576 */
        const uint8_t *row = bed + m * B0;
        uint8_t *dst = (dec[m] & 1) ? g->geno + dstq[m] * g->B : g->geno_vr + dstq[m] * g->B;
        memset(dst, 0, g->B);
        for (int64_t k = 0; k < N; k++) {
            int64_t src = sub_idx[k] - 1;
            int gv = bed_code_to_geno((row[src >> 2] >> ((src & 3) << 1)) & 3);
            if (gv == 3) gv = fillv[m];
            dst[k >> 2] |= (uint8_t)(geno_to_code(gv) << ((k & 3) << 1));
        }
    }
    g->M = Mq; g->Mvr = Mv;
    free(invr); free(dec); free(fillv); free(acv); free(macv); free(fv); free(isv); free(dstq);
    return 0;
}

int64_t orc_get_M(orc_geno *g) { return g->M; }
int64_t orc_get_M0(orc_geno *g) { return g->M0; }
int64_t orc_get_N(orc_geno *g) { return g->N; }
int64_t orc_get_Mvr(orc_geno *g) { return g->Mvr; }
int64_t orc_get_B(orc_geno *g) { return g->B; }
const float *orc_afreq(orc_geno *g, int vr) { return vr ? g->afreq_vr : g->afreq; }
const float *orc_invstd(orc_geno *g, int vr) { return vr ? g->invstd_vr : g->invstd; }
const int32_t *orc_mac(orc_geno *g, int vr) { return vr ? g->mac_vr : g->mac; }
const int32_t *orc_ac(orc_geno *g, int vr) { return vr ? g->ac_vr : g->ac; }
const int32_t *orc_index_vr(orc_geno *g) { return g->index_vr; }
const uint8_t *orc_qc_mask(orc_geno *g) { return g->qc_mask; }
const uint8_t *orc_packed(orc_geno *g, int vr) { return vr ? g->geno_vr : g->geno; }

/* Get_OneSNP_Geno / Get_OneSNP_Geno_forVarRatio (Fg->cpp:223-323): geno = 2-(a+b) */
void orc_one_snp_geno(orc_geno *g, int64_t idx, int vr, int32_t *out)
{
    const uint8_t *row = (vr ? g->geno_vr : g->geno) + idx * g->B;
    for (int64_t i = 0; i < g->N; i++) {
        int c = (row[i >> 2] >> ((i & 3) << 1)) & 3;
        out[i] = 2 - ((c & 1) + ((c >> 1) & 1));
    }
}

/* fp64 definitions used by the GPU build and by mode 0 (documented in DESIGN.md "parity definition") */
static inline void marker_fs64(orc_geno *g, int64_t m, double *f, double *s)
{
    double ff = (double)g->ac[m] / (double)(2 * g->N);
    double v = 2.0 * ff * (1.0 - ff);
    *f = ff;
    *s = v > 0 ? 1.0 / sqrt(v) : 0.0;
}

/* Get_OneSNP_StdGeno (Fg->cpp:582-662) */
void orc_one_snp_stdgeno(orc_geno *g, int64_t idx, int mode, double *out)
{
    const uint8_t *row = g->geno + idx * g->B;
    if (mode == 1) {
        float f2 = 2 * g->afreq[idx], is = g->invstd[idx];
        float lut[3] = {(0 - f2) * is, (1 - f2) * is, (2 - f2) * is};   /* Fg->cpp:152-159 */
        for (int64_t i = 0; i < g->N; i++) {
            int c = (row[i >> 2] >> ((i & 3) << 1)) & 3;
            out[i] = lut[2 - ((c & 1) + ((c >> 1) & 1))];
        }
    } else {
        double f, s; marker_fs64(g, idx, &f, &s);
        double lut[3] = {(0 - 2 * f) * s, (1 - 2 * f) * s, (2 - 2 * f) * s};
        for (int64_t i = 0; i < g->N; i++) {
            int c = (row[i >> 2] >> ((i & 3) << 1)) & 3;
            out[i] = lut[2 - ((c & 1) + ((c >> 1) & 1))];
        }
    }
}

/*
 * sum_{m in [m0,m1)} z_m (z_m^T b)  -- NOT divided by the marker count, like parallelCrossProd_full /
 * parallelCrossProdOpenMP (Fg->cpp:1576-1598, 1746-1786).  k right-hand sides, column-major b[N*k].
 */
void orc_crossprod_range(orc_geno *g, int64_t m0, int64_t m1, const double *b, int k, int mode, double *out)
{
    const int64_t N = g->N;
    memset(out, 0, sizeof(double) * N * k);
    if (mode == 1) {
        /* reference arithmetic: float vectors, float dot, float axpy; thread-private accumulators */
        float *bf = (float *)malloc(sizeof(float) * N * k);
        for (int64_t i = 0; i < N * k; i++) bf[i] = (float)b[i];
#pragma omp parallel
        {
            float *acc = (float *)calloc(N * k, sizeof(float));
            float *vec = (float *)malloc(sizeof(float) * N);
#pragma omp for schedule(static)
            for (int64_t m = m0; m < m1; m++) {
                const uint8_t *row = g->geno + m * g->B;
                float f2 = 2 * g->afreq[m], is = g->invstd[m];
                float lut[4] = {(2 - f2) * is, 0.f, (1 - f2) * is, (0 - f2) * is}; /* indexed by PLINK code */
                for (int64_t i = 0; i < N; i++) vec[i] = lut[(row[i >> 2] >> ((i & 3) << 1)) & 3];
                for (int c = 0; c < k; c++) {
                    const float *bc = bf + (int64_t)c * N; float *ac = acc + (int64_t)c * N;
                    float val = 0.f;
                    for (int64_t i = 0; i < N; i++) val += vec[i] * bc[i];
                    for (int64_t i = 0; i < N; i++) ac[i] += val * vec[i];
                }
            }
#pragma omp critical
            for (int64_t i = 0; i < N * k; i++) out[i] += acc[i];
            free(acc); free(vec);
        }
        free(bf);
    } else {
#pragma omp parallel
        {
            double *acc = (double *)calloc(N * k, sizeof(double));
            double *vec = (double *)malloc(sizeof(double) * N);
#pragma omp for schedule(static)
            for (int64_t m = m0; m < m1; m++) {
                const uint8_t *row = g->geno + m * g->B;
                double f, s; marker_fs64(g, m, &f, &s);
                double lut[4] = {(2 - 2 * f) * s, 0.0, (1 - 2 * f) * s, (0 - 2 * f) * s};
                for (int64_t i = 0; i < N; i++) vec[i] = lut[(row[i >> 2] >> ((i & 3) << 1)) & 3];
                for (int c = 0; c < k; c++) {
                    const double *bc = b + (int64_t)c * N; double *ac = acc + (int64_t)c * N;
                    double val = 0.0;
                    for (int64_t i = 0; i < N; i++) val += vec[i] * bc[i];
                    for (int64_t i = 0; i < N; i++) ac[i] += val * vec[i];
                }
            }
#pragma omp critical
            for (int64_t i = 0; i < N * k; i++) out[i] += acc[i];
            free(acc); free(vec);
        }
    }
}

/* sum_{m in [m0,m1)} z_mi^2  -- Get_Diagof_StdGeno (Fg->cpp:665-704) / set_Diagof_StdGeno_LOCO (Fg->cpp:4934-4958) */
void orc_diag_range(orc_geno *g, int64_t m0, int64_t m1, int mode, double *out)
{
    const int64_t N = g->N;
    memset(out, 0, sizeof(double) * N);
    if (mode == 1) {
        float *acc = (float *)calloc(N, sizeof(float));
        for (int64_t m = m0; m < m1; m++) {
            const uint8_t *row = g->geno + m * g->B;
            float f2 = 2 * g->afreq[m], is = g->invstd[m];
            float lut[4] = {(2 - f2) * is, 0.f, (1 - f2) * is, (0 - f2) * is};
            for (int64_t i = 0; i < N; i++) { float z = lut[(row[i >> 2] >> ((i & 3) << 1)) & 3]; acc[i] = acc[i] + z * z; }
        }
        for (int64_t i = 0; i < N; i++) out[i] = acc[i];
        free(acc);
    } else {
        /* fp64 mode: marker blocks in parallel, partial sums added as the threads finish (the result depends on
         * the schedule only through the rounding of fp64 sums, ~1e-16 relative) */
#pragma omp parallel
        {
            double *acc = (double *)calloc(N, sizeof(double));
#pragma omp for schedule(static)
            for (int64_t m = m0; m < m1; m++) {
                const uint8_t *row = g->geno + m * g->B;
                double f, s; marker_fs64(g, m, &f, &s);
                double lut[4] = {(2 - 2 * f) * s, 0.0, (1 - 2 * f) * s, (0 - 2 * f) * s};
                for (int64_t i = 0; i < N; i++) { double z = lut[(row[i >> 2] >> ((i & 3) << 1)) & 3]; acc[i] += z * z; }
            }
#pragma omp critical
            for (int64_t i = 0; i < N; i++) out[i] += acc[i];
            free(acc);
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU-baseline legs of bench.py set the count explicitly */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- synthetic genotype generator shared (bit-exactly) with the CUDA library's sgb_synth_* ----
 * Counter-based: u32 = mix(seed, marker, sample); genotype = (u >= t0) + (u >= t1) with per-marker integer
 * thresholds t0 = floor((1-f)^2 * 2^32), t1 = floor((1-f^2) * 2^32) supplied by the caller.
 * Written out in PLINK .bed coding (g copies of A1: 2->00, 1->10, 0->11), optional missing (01) when
 * mix2 < miss_thr.  Not part of the reference; it is the bench/test workload of SURVEY.md 8(d). */
static inline uint32_t mix32(uint64_t seed, uint64_t m, uint64_t i, uint64_t salt)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (m + 1) + 0xBF58476D1CE4E5B9ull * (i + 1) + 0x94D049BB133111EBull * salt;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

void orc_synth_bed(uint8_t *bed, int64_t N0, int64_t M0, uint64_t seed, const uint32_t *t0, const uint32_t *t1,
                   uint32_t miss_thr)
{
    int64_t B0 = (N0 + 3) / 4;
    memset(bed, 0, (size_t)B0 * M0);
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M0; m++)
        for (int64_t i = 0; i < N0; i++) {
            uint32_t u = mix32(seed, (uint64_t)m, (uint64_t)i, 0);
            int gv = (u >= t0[m]) + (u >= t1[m]);
            int code = gv == 2 ? 0x0 : (gv == 1 ? 0x2 : 0x3);
            if (miss_thr && mix32(seed, (uint64_t)m, (uint64_t)i, 1) < miss_thr) code = 0x1;
            bed[m * B0 + (i >> 2)] |= (uint8_t)(code << ((i & 3) << 1));
        }
}
