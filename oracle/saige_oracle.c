/*
 * saige_oracle.c -- CPU restatement of the SAIGE step-1 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA library in saige_gpu_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never links or calls it.
 *
 * It restates, function by function, what /root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp
 * ("FG.cpp") does on the packed-genotype store and the GRM.vector product.  The reference itself
 * cannot be built here (needs R, Rcpp, RcppArmadillo, RcppParallel, MPI, boost -- FG.cpp:4-19).
 *
 * Parity pins (tests/test_oracle_golden.py):
 *   - decode + allele counts + MAF filter  <-> extdata/input/plinkforGRM_1000samples_10kMarkers.frq (exact)
 *   - fp32 reference-order matvec vs fp64 matvec (informational, ~1e-6)
 *   - everything above the matvec (PCG / AI-REML) is restated in numpy (oracle/oracle.py); those
 *     call boundaries have no golden vectors in the reference => "parity unpinned" there.
 *
 * Two arithmetic modes for the matvec:
 *   mode 0 ("fp64")   : f = AC/(2N) and s = 1/sqrt(2f(1-f)) evaluated in double from the integer allele
 *                       count; accumulation in double.  This is what the GPU is graded against.
 *   mode 1 ("ref32")  : float f / invStd exactly as the reference stores them (FG.cpp:484,916-921),
 *                       3-entry float LUT (FG.cpp:152-159), float dot + float axpy in marker order
 *                       (FG.cpp:1576-1598), thread-private float accumulators under OpenMP.
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
    int64_t N0;          /* samples in .fam                      (geno.N,  FG.cpp:771)  */
    int64_t M0;          /* markers in .bim                      (geno.M,  FG.cpp:786)  */
    int64_t N;           /* phenotyped samples                   (Nnomissing, FG.cpp:744) */
    int64_t M;           /* markers passing QC                   (numberofMarkerswithMAFge_minMAFtoConstructGRM) */
    int64_t Mvr;         /* markers held out for variance ratio  (numberofMarkers_varRatio) */
    int64_t B;           /* bytes per re-packed marker = ceil(N/4) (m_size_of_esi, FG.cpp:214) */
    uint8_t *geno;       /* M  x B, PLINK 2-bit codes, no missing (genoVecofPointers)          */
    uint8_t *geno_vr;    /* Mvr x B                              (genoVecofPointers_forVarRatio) */
    float *afreq;        /* alleleFreqVec   (float, FG.cpp:484,923) */
    float *invstd;       /* invstdvVec      (float, FG.cpp:916-922) */
    int32_t *mac;        /* MACVec          (FG.cpp:490,926) */
    int32_t *ac;         /* allele count after imputation (not kept by the reference; = afreq*2N exactly) */
    float *afreq_vr, *invstd_vr;
    int32_t *mac_vr, *index_vr, *ac_vr;
    uint8_t *qc_mask;    /* M0 flags: MarkerswithMAFge_minMAFtoConstructGRM_indVec (FG.cpp:927,930) */
} orc_geno;

static orc_geno G;

void orc_free(void)
{
    free(G.geno); free(G.geno_vr); free(G.afreq); free(G.invstd); free(G.mac); free(G.ac);
    free(G.afreq_vr); free(G.invstd_vr); free(G.mac_vr); free(G.index_vr); free(G.ac_vr); free(G.qc_mask);
    memset(&G, 0, sizeof(G));
}

/* PLINK .bed 2-bit code -> bufferGeno of Get_OneSNP_Geno_atBeginning (FG.cpp:343-354):
 * b = low bit, a = high bit; (b=1,a=0)->3 missing; (0,0)->2; (b=0,a=1)->1; (1,1)->0. */
static inline int bed_code_to_geno(int code)
{
    int b = code & 1, a = (code >> 1) & 1;
    if (b == 1 && a == 0) return 3;
    if (b == 0 && a == 0) return 2;
    if (b == 0 && a == 1) return 1;
    return 0;
}

/* setGenotype with HOM_ALT=0x3 (geno 0), HET=0x2 (geno 1), HOM_REF=0x0 (geno 2): FG.cpp:41-44,203-205,559-565 */
static inline uint8_t geno_to_code(int g)
{
    return g == 0 ? 0x3 : (g == 1 ? 0x2 : 0x0);
}

/*
 * setGenoObj + Get_OneSNP_Geno_atBeginning (FG.cpp:739-1024, 326-579) on an in-memory .bed body
 * (`bed` points just past the 3 magic bytes; marker i starts at bed + i*ceil(N0/4), FG.cpp:902).
 *   sub_idx[N]   : subSampleInGeno, 1-based row of each phenotyped sample in the .fam (FG.cpp:555)
 *   indicator[N0]: indicatorGenoSamplesWithPheno (FG.cpp:376)
 *   vr_idx[n_vr] : g_randMarkerIndforVR -- drawn by the CALLER (the reference uses arma::randi, FG.cpp:866-868)
 * Returns 0, or -1 on allocation failure.
 */
int orc_setgeno(const uint8_t *bed, int64_t N0, int64_t M0, const int32_t *sub_idx, int64_t N,
                const uint8_t *indicator, float minMAF, float maxMissing, int isVarRatio,
                float minMACvr, float maxMACvr, const int32_t *vr_idx, int64_t n_vr)
{
    orc_free();
    G.N0 = N0; G.M0 = M0; G.N = N; G.B = (N + 3) / 4;
    int64_t B0 = (N0 + 3) / 4;
    G.geno = (uint8_t *)malloc((size_t)(M0 > 0 ? M0 : 1) * G.B);
    G.afreq = (float *)malloc(sizeof(float) * (M0 + 1));
    G.invstd = (float *)malloc(sizeof(float) * (M0 + 1));
    G.mac = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1));
    G.ac = (int32_t *)malloc(sizeof(int32_t) * (M0 + 1));
    G.qc_mask = (uint8_t *)calloc(M0 + 1, 1);
    int64_t vr_cap = isVarRatio ? M0 : 0;
    G.geno_vr = (uint8_t *)malloc((size_t)(vr_cap > 0 ? vr_cap : 1) * G.B);
    G.afreq_vr = (float *)malloc(sizeof(float) * (vr_cap + 1));
    G.invstd_vr = (float *)malloc(sizeof(float) * (vr_cap + 1));
    G.mac_vr = (int32_t *)malloc(sizeof(int32_t) * (vr_cap + 1));
    G.ac_vr = (int32_t *)malloc(sizeof(int32_t) * (vr_cap + 1));
    G.index_vr = (int32_t *)malloc(sizeof(int32_t) * (vr_cap + 1));
    int *tmp = (int *)malloc(sizeof(int) * (N0 + 4));
    uint8_t *invr = (uint8_t *)calloc(M0 + 1, 1);
    if (!G.geno || !G.afreq || !G.invstd || !G.mac || !G.ac || !G.qc_mask || !G.geno_vr || !tmp || !invr) return -1;
    for (int64_t j = 0; j < n_vr; j++)
        if (vr_idx[j] >= 0 && vr_idx[j] < M0) invr[vr_idx[j]] = 1;

    int64_t Mq = 0, Mv = 0;
    for (int64_t m = 0; m < M0; m++) {
        const uint8_t *row = bed + m * B0;
        int alleleCount = 0, numMissing = 0;
        for (int64_t i = 0; i < N0; i++) {
            int g = bed_code_to_geno((row[i >> 2] >> ((i & 3) << 1)) & 3);
            tmp[i] = g;
            if (indicator[i]) { if (g == 3) numMissing++; else alleleCount += g; }
        }
        /* FG.cpp:438-447 -- float arithmetic, in this order */
        float altFreq = alleleCount / (float)((N - numMissing) * 2);
        float missingRate = numMissing / (float)N;
        int fill = (int)roundf(2 * altFreq);              /* int(round(2*altFreq)) FG.cpp:447 */
        if (numMissing > 0) alleleCount += fill * numMissing;   /* FG.cpp:454 */
        altFreq = alleleCount / (float)(N * 2);           /* FG.cpp:484 */
        float maf = altFreq < 1 - altFreq ? altFreq : 1 - altFreq;   /* FG.cpp:489 */
        int mac = alleleCount < (int)N * 2 - alleleCount ? alleleCount : (int)N * 2 - alleleCount;
        int passQC = (maf >= minMAF && missingRate <= maxMissing);   /* FG.cpp:493 */
        int passVR = 0;
        if (isVarRatio) {                                   /* FG.cpp:496-548 */
            if (maxMACvr != -1) {
                if (mac >= minMACvr && mac < maxMACvr) passVR = 1;
                else if (mac >= maxMACvr) passVR = invr[m];
            } else {
                if (mac >= minMACvr) passVR = invr[m];
            }
            if (passVR) passQC = 0;
        }
        if (passQC || passVR) {                             /* FG.cpp:551-576 */
            uint8_t *dst = passQC ? G.geno + Mq * G.B : G.geno_vr + Mv * G.B;
            memset(dst, 0, G.B);
            for (int64_t k = 0; k < N; k++) {
                int g = tmp[sub_idx[k] - 1];
                if (g == 3) g = fill;
                dst[k >> 2] |= (uint8_t)(geno_to_code(g) << ((k & 3) << 1));
            }
        }
        float Std = sqrtf(2 * altFreq * (1 - altFreq));    /* FG.cpp:916-921 */
        float invStd = (Std == 0) ? 0.f : 1 / Std;
        if (passQC) {
            G.afreq[Mq] = altFreq; G.invstd[Mq] = invStd; G.mac[Mq] = mac; G.ac[Mq] = alleleCount;
            G.qc_mask[m] = 1; Mq++;
        }
        if (isVarRatio && passVR) {
            G.afreq_vr[Mv] = altFreq; G.invstd_vr[Mv] = invStd; G.mac_vr[Mv] = mac; G.ac_vr[Mv] = alleleCount;
            G.index_vr[Mv] = (int32_t)m; Mv++;
        }
    }
    G.M = Mq; G.Mvr = Mv;
    free(tmp); free(invr);
    return 0;
}

int64_t orc_get_M(void) { return G.M; }
int64_t orc_get_M0(void) { return G.M0; }
int64_t orc_get_N(void) { return G.N; }
int64_t orc_get_Mvr(void) { return G.Mvr; }
int64_t orc_get_B(void) { return G.B; }
const float *orc_afreq(int vr) { return vr ? G.afreq_vr : G.afreq; }
const float *orc_invstd(int vr) { return vr ? G.invstd_vr : G.invstd; }
const int32_t *orc_mac(int vr) { return vr ? G.mac_vr : G.mac; }
const int32_t *orc_ac(int vr) { return vr ? G.ac_vr : G.ac; }
const int32_t *orc_index_vr(void) { return G.index_vr; }
const uint8_t *orc_qc_mask(void) { return G.qc_mask; }
const uint8_t *orc_packed(int vr) { return vr ? G.geno_vr : G.geno; }

/* Get_OneSNP_Geno / Get_OneSNP_Geno_forVarRatio (FG.cpp:223-323): geno = 2-(a+b) */
void orc_one_snp_geno(int64_t idx, int vr, int32_t *out)
{
    const uint8_t *row = (vr ? G.geno_vr : G.geno) + idx * G.B;
    for (int64_t i = 0; i < G.N; i++) {
        int c = (row[i >> 2] >> ((i & 3) << 1)) & 3;
        out[i] = 2 - ((c & 1) + ((c >> 1) & 1));
    }
}

/* fp64 definitions used by the GPU build and by mode 0 (documented in DESIGN.md "parity definition") */
static inline void marker_fs64(int64_t m, double *f, double *s)
{
    double ff = (double)G.ac[m] / (double)(2 * G.N);
    double v = 2.0 * ff * (1.0 - ff);
    *f = ff;
    *s = v > 0 ? 1.0 / sqrt(v) : 0.0;
}

/* Get_OneSNP_StdGeno (FG.cpp:582-662) */
void orc_one_snp_stdgeno(int64_t idx, int mode, double *out)
{
    const uint8_t *row = G.geno + idx * G.B;
    if (mode == 1) {
        float f2 = 2 * G.afreq[idx], is = G.invstd[idx];
        float lut[3] = {(0 - f2) * is, (1 - f2) * is, (2 - f2) * is};   /* FG.cpp:152-159 */
        for (int64_t i = 0; i < G.N; i++) {
            int c = (row[i >> 2] >> ((i & 3) << 1)) & 3;
            out[i] = lut[2 - ((c & 1) + ((c >> 1) & 1))];
        }
    } else {
        double f, s; marker_fs64(idx, &f, &s);
        double lut[3] = {(0 - 2 * f) * s, (1 - 2 * f) * s, (2 - 2 * f) * s};
        for (int64_t i = 0; i < G.N; i++) {
            int c = (row[i >> 2] >> ((i & 3) << 1)) & 3;
            out[i] = lut[2 - ((c & 1) + ((c >> 1) & 1))];
        }
    }
}

/*
 * sum_{m in [m0,m1)} z_m (z_m^T b)  -- NOT divided by the marker count, like parallelCrossProd_full /
 * parallelCrossProdOpenMP (FG.cpp:1576-1598, 1746-1786).  k right-hand sides, column-major b[N*k].
 */
void orc_crossprod_range(int64_t m0, int64_t m1, const double *b, int k, int mode, double *out)
{
    const int64_t N = G.N;
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
                const uint8_t *row = G.geno + m * G.B;
                float f2 = 2 * G.afreq[m], is = G.invstd[m];
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
                const uint8_t *row = G.geno + m * G.B;
                double f, s; marker_fs64(m, &f, &s);
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

/* sum_{m in [m0,m1)} z_mi^2  -- Get_Diagof_StdGeno (FG.cpp:665-704) / set_Diagof_StdGeno_LOCO (FG.cpp:4934-4958) */
void orc_diag_range(int64_t m0, int64_t m1, int mode, double *out)
{
    const int64_t N = G.N;
    memset(out, 0, sizeof(double) * N);
    if (mode == 1) {
        float *acc = (float *)calloc(N, sizeof(float));
        for (int64_t m = m0; m < m1; m++) {
            const uint8_t *row = G.geno + m * G.B;
            float f2 = 2 * G.afreq[m], is = G.invstd[m];
            float lut[4] = {(2 - f2) * is, 0.f, (1 - f2) * is, (0 - f2) * is};
            for (int64_t i = 0; i < N; i++) { float z = lut[(row[i >> 2] >> ((i & 3) << 1)) & 3]; acc[i] = acc[i] + z * z; }
        }
        for (int64_t i = 0; i < N; i++) out[i] = acc[i];
        free(acc);
    } else {
        for (int64_t m = m0; m < m1; m++) {
            const uint8_t *row = G.geno + m * G.B;
            double f, s; marker_fs64(m, &f, &s);
            double lut[4] = {(2 - 2 * f) * s, 0.0, (1 - 2 * f) * s, (0 - 2 * f) * s};
            for (int64_t i = 0; i < N; i++) { double z = lut[(row[i >> 2] >> ((i & 3) << 1)) & 3]; out[i] += z * z; }
        }
    }
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
    for (int64_t m = 0; m < M0; m++)
        for (int64_t i = 0; i < N0; i++) {
            uint32_t u = mix32(seed, (uint64_t)m, (uint64_t)i, 0);
            int g = (u >= t0[m]) + (u >= t1[m]);
            int code = g == 2 ? 0x0 : (g == 1 ? 0x2 : 0x3);
            if (miss_thr && mix32(seed, (uint64_t)m, (uint64_t)i, 1) < miss_thr) code = 0x1;
            bed[m * B0 + (i >> 2)] |= (uint8_t)(code << ((i & 3) << 1));
        }
}
