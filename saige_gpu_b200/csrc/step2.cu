// Step-2 single-variant score test + saddle-point approximation, batched over variants.
//
// Replaces, per marker, the body of mainMarkerInCPP (Main.cpp:229-520): PlinkClass::getOneMarker (PLINK.cpp:164-300,
// alt-first), the MAF/MAC/missing-rate filter, imputeGenoAndFlip (UTIL.cpp:58-135, best_guess), scoreTestFast
// (SAIGE_test.cpp:212-292) and, when |T|/sqrt(var1) > SPAcutoff on a binary trait, getMarkerPval's SPA / SPA_fast
// branch (SAIGE_test.cpp:345-640, SPA.cpp:20-185, SPA_binary.cpp:21-330) and, when asked for, Firth's bias-reduced
// effect size of significant variants (fast_logistf_fit_simple, SAIGE_test.cpp:893-986), and the exact test of rare
// variants (MAC <= max_MAC_for_ER: Main.cpp:408-422, SAIGE_test.cpp:426-431, 592-620 -> er_exact.h).  All arithmetic fp64,
// like the reference.
//
// One CTA per variant.  The raw PLINK row (2 bits per .fam sample) is staged in shared memory; the per-sample model
// vectors (mu, mu2, res, X, XVX_inv_XV, XXVX_inv: N x (3p + 3) doubles) are read through L2 by every CTA.
// scoreTestFast's sums over the non-zero genotypes are folded into one pass:
//     Z = A^T g,  W = (mu2*X)^T g,  T1 = sum mu2 g^2,  R0 = sum res g
//     var2 = Z^T XVX Z + T1 - 2 Z.W            S = (R0 - S_a.Z) / tau0          (algebraically identical)
// HBM-bound on the genotype bytes (N/4 per variant) once N is large; no tensor cores (integer decode + fp64 sums).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "sgb_internal.h"
#include "er_exact.h"

#define S2_MAXP 16
#define S2_THREADS 256
#define S2_NOUT 32      // doubles per variant in the result table (see include/saige_b200.h)
#define S2_MAXCOND 4    // conditioning markers
#define S2_MAXCATE 8    // MAC categories of the variance ratio (the reference's default has 2)

struct s2_model {
    int64_t N; int p; int binary;
    const double *mu, *mu2, *res, *y, *X, *A /*XVX_inv_XV*/, *XXVXi /*XXVX_inv*/;
    double XVX[S2_MAXP * S2_MAXP], S_a[S2_MAXP];
    double tau0, varRatio, spa_cutoff;
    const int32_t *pos;      // model sample -> row in the .fam
    int identity;            // pos[i] == i: the model's samples are the first N rows of the .fam, in order
    const uint32_t *ycase;   // identity only: bit 2j of word w set when sample 16w + j is a case (y == 1)
    double ncase_tot;        // number of cases in the model
    // Firth's bias-reduced effect size for significant variants (is_Firth_beta, SAIGE_test.cpp:573-633)
    const double *offset;    // N, the null model's offset (zeros when absent)
    int firth, firth_se_from_fit;
    double firth_cutoff;
    // exact test of rare variants (g_MACCutoffforER, Main.cpp:68,408): off when negative
    double er_max_mac;
    // categorical variance ratios (assignVarianceRatio, SAIGE_test.cpp:801-833): category c covers cate_min[c] < MAC <=
    // cate_max[c], the last one is open-ended; n_cate == 1: the single ratio `varRatio`
    int n_cate;
    double cate_ratio[S2_MAXCATE], cate_min[S2_MAXCATE], cate_max[S2_MAXCATE];
    double mu_sum;           // sum of mu over the model's samples (mean fitted probability of a variant's non-carriers)
    // conditional analysis (assignConditionFactors, Main.cpp:2002-2179): P2 = sqrt(vr) gtilde_cond % mu2 tau0 (N x n_cond, device),
    // XtP2 = XXVX_inv^T P2 (p x n_cond), VarInv = pinv(P1 P2), VT = VarInv Tstat_cond
    int n_cond;
    const double *P2;
    double XtP2[S2_MAXP * S2_MAXCOND], VarInv[S2_MAXCOND * S2_MAXCOND], VT[S2_MAXCOND];
};

__device__ __forceinline__ double block_sum(double v, double *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < S2_THREADS / 32; i++) t += sm[i];
    return t;
}

// log of the chi-square(1) upper tail at stat = 2 x^2, i.e. log erfc(x), finite where erfc underflows (x > 26.5): the reference
// switches to R::pchisq(..., log = TRUE) there and prints "%.1fE%d" strings (SAIGE_test.cpp:255-284)
__device__ __forceinline__ double s2_log_erfc(double x) { return x < 20.0 ? log(erfc(x)) : log(erfcx(x)) - x * x; }
// |z| whose upper normal tail is exp(lp): qnorm(p, lower = FALSE, log.p = TRUE) of the reference (SAIGE_test.cpp:531)
__device__ __forceinline__ double s2_qnorm_from_logp(double lp)
{
    if (lp > -700.0) return fabs(normcdfinv(exp(lp)));
    double z = sqrt(-2.0 * lp);
    for (int it = 0; it < 8; it++) {
        const double x = z * 0.7071067811865476;
        const double lt = -0.6931471805599453 + log(erfcx(x)) - x * x;      // log of the upper tail at z
        z += (lt - lp) * 1.2533141373155001 * erfcx(x);                     // Newton: d/dz log tail = -density / tail
    }
    return z;
}
__device__ __forceinline__ double s2_logaddexp(double a, double b)
{
    const double hi = fmax(a, b), lo = fmin(a, b);
    return isinf(hi) ? hi : hi + log1p(exp(lo - hi));
}

// genotype of model sample i after flip / imputation: copies of the (possibly flipped) ALT allele
__device__ __forceinline__ int s2_geno(const uint8_t *srow, int32_t src, int flip, int imputeG)
{
    int code = (srow[src >> 2] >> ((src & 3) << 1)) & 3;
    int g = code == 0 ? 2 : (code == 2 ? 1 : (code == 3 ? 0 : -1));      // PLINK.hpp:48-56, alt-first
    if (g < 0) return imputeG;
    return flip ? 2 - g : g;
}

// dosage input (VCF DS / BGEN): one row of doubles per variant in the file's sample order, negative or NaN = missing
struct s2_dose {
    const double *d;        // n_markers x stride
    int64_t stride;         // samples per row in the file
    int impute;             // 1 best_guess, 2 mean, 3 minor (UTIL.cpp:80-93)
    double zerod_cutoff, zerod_mac_cutoff;      // dosages <= cutoff are zeroed when MAC <= mac cutoff (UTIL.cpp:105-109)
};

// dosage of model sample `src` after flip / imputation / zeroing (imputeGenoAndFlip, UTIL.cpp:58-135: flip, impute, clean)
__device__ __forceinline__ double s2_dose_geno(const double *row, int32_t src, int flip, double imputeG, double zero_below)
{
    double g = row[src];
    if (!(g >= 0.0)) g = imputeG;
    else if (flip) g = 2.0 - g;
    return fabs(g) <= zero_below ? 0.0 : g;          // zero_below < 0: zeroing off
}

struct s2_cgf {          // binomial CGF pieces over the non-zero genotypes + normal approximation of the zeros
    double NAmu, NAsigma;
    int fast;
};

// IDENT: the model's samples are the first N rows of the .fam in order (the usual case).  Then genotype classes are
// counted 16 samples at a time with popcounts on the raw PLINK words (allele / missing counts and every case-control
// tally), warps skip 32-sample groups that hold no minor allele, and no per-sample index or phenotype is read.
// DOSE: the genotypes are doubles read through L2 (s2_dose) instead of 2-bit codes staged in shared memory; always with the
// sample index (IDENT = false).
// The hard-call fast path keeps three CTAs per SM (<= 85 registers, as before the conditional / dosage code was added: the
// rarely taken branches spill, the popcount passes do not).
// Per-CTA scratch of the saddle-point branch (batched path: one slot per persistent CTA): gtilde of every sample, and for
// SPA_fast the compact list of (gtilde, mu) pairs of the samples with a non-zero genotype, in sample order.  The Newton
// passes then stream 16 bytes per evaluated sample instead of decoding the genotype and gathering p + 1 model values.
struct s2_spa_scratch { double *gt; double2 *nz; };

template <bool IDENT, bool DOSE>
__device__ __forceinline__ void step2_variant(const s2_model &M, const uint8_t *__restrict__ bed, int64_t B0, int64_t nm, int64_t m, double min_maf,
                                              double min_mac, double max_missing, int se_two_sided, double *__restrict__ out, const s2_dose &DS,
                                              s2_spa_scratch scr)
{
    static_assert(!(IDENT && DOSE), "dosage rows are always indexed");
    extern __shared__ uint8_t srow[];
    __shared__ double red[S2_THREADS / 32];
    __shared__ double Zs[S2_MAXP], Ws[S2_MAXP];
    __shared__ int er_cnt, er_idx[SGB_ER_MAXK];
    __shared__ double er_pv;
    __shared__ int nz_warp[S2_THREADS / 32];
    if (m >= nm) return;
    const int tid = threadIdx.x, p = M.p;
    const int64_t N = M.N;
    double *o = out + m * S2_NOUT;
    const double *drow = DOSE ? DS.d + m * DS.stride : nullptr;
    if (!DOSE) {
        for (int64_t b = tid; b < B0 + 8; b += S2_THREADS) srow[b] = b < B0 ? bed[m * B0 + b] : (uint8_t)0;      // 8 pad bytes: word-wise reads
        __syncthreads();
    }

    // ---- getOneMarker: counts over the model's samples ----
    double altCounts0, nMiss;
    double k2c = 0, k1c = 0, kmc = 0, k2a = 0, k1a = 0;          // IDENT: hom-alt / het / missing among cases, hom-alt / het among all
    if (IDENT) {
        const int64_t nw = (N + 15) >> 4;
        const uint32_t *wrow = reinterpret_cast<const uint32_t *>(srow);
        int c2 = 0, c1 = 0, cm = 0, d2 = 0, d1 = 0, dm = 0;
        for (int64_t w = tid; w < nw; w += S2_THREADS) {
            const uint32_t x = wrow[w];
            uint32_t valid = 0x55555555u;
            if (w == nw - 1 && (N & 15)) valid &= (1u << (2 * (N & 15))) - 1u;
            const uint32_t L = x & 0x55555555u, H = (x >> 1) & 0x55555555u;
            const uint32_t homalt = ~L & ~H & valid, het = H & ~L & valid, miss = L & ~H & valid;      // PLINK.hpp:48-56, alt-first
            const uint32_t yc = M.ycase[w];
            c2 += __popc(homalt); c1 += __popc(het); cm += __popc(miss);
            d2 += __popc(homalt & yc); d1 += __popc(het & yc); dm += __popc(miss & yc);
        }
        k2a = block_sum((double)c2, red); k1a = block_sum((double)c1, red); nMiss = block_sum((double)cm, red);
        k2c = block_sum((double)d2, red); k1c = block_sum((double)d1, red); kmc = block_sum((double)dm, red);
        altCounts0 = 2.0 * k2a + k1a;
    } else if (DOSE) {
        double c_alt = 0, c_miss = 0;
        for (int64_t i = tid; i < N; i += S2_THREADS) {
            const double g = drow[M.pos[i]];
            if (g >= 0.0) c_alt += g; else c_miss += 1.0;
        }
        altCounts0 = block_sum(c_alt, red); nMiss = block_sum(c_miss, red);
    } else {
        double c_alt = 0, c_miss = 0;
        for (int64_t i = tid; i < N; i += S2_THREADS) {
            int32_t src = M.pos[i];
            int code = (srow[src >> 2] >> ((src & 3) << 1)) & 3;
            c_alt += code == 0 ? 2.0 : (code == 2 ? 1.0 : 0.0);
            c_miss += code == 1 ? 1.0 : 0.0;
        }
        altCounts0 = block_sum(c_alt, red); nMiss = block_sum(c_miss, red);
    }
    const double cnt = (double)N - nMiss;
    double altFreq = cnt > 0 ? altCounts0 / cnt / 2.0 : 0.0;
    const double missingRate = nMiss / (double)N;
    const double MAF = fmin(altFreq, 1.0 - altFreq);
    const double MAC0 = MAF * (double)N * (1.0 - missingRate) * 2.0;
    if (missingRate > max_missing || MAF < min_maf || MAC0 < min_mac) {       // Main.cpp:296
        if (tid == 0) { for (int c = 0; c < S2_NOUT; c++) o[c] = nan(""); o[0] = 0.0; o[2] = altFreq; o[3] = missingRate; }
        return;
    }
    // ---- imputeGenoAndFlip (best_guess) ----
    const int flip = altFreq > 0.5;
    if (flip) altFreq = 1.0 - altFreq;
    const int imputeG = nMiss > 0 ? (int)round(2.0 * altFreq) : 0;
    // dosage rows: the three imputation methods and the zeroing of small dosages of rare variants (UTIL.cpp:80-109)
    double imputeGd = 0.0, zero_below = -1.0;
    if (DOSE) {
        if (nMiss > 0) imputeGd = DS.impute == 1 ? round(2.0 * altFreq) : (DS.impute == 2 ? 2.0 * altFreq : 0.0);
        if (DS.zerod_cutoff > 0.0 && MAC0 + imputeGd * nMiss <= DS.zerod_mac_cutoff) zero_below = DS.zerod_cutoff;
    }
    // genotype of model sample i in the tested coding
    auto GENO = [&](int64_t i) -> double {
        if (DOSE) return s2_dose_geno(drow, M.pos[i], flip, imputeGd, zero_below);
        return (double)s2_geno(srow, IDENT ? (int32_t)i : M.pos[i], flip, imputeG);
    };

    // ---- one pass over the samples: every sum scoreTestFast needs + case/control tallies ----
    double zs[S2_MAXP], ws[S2_MAXP];
#pragma unroll
    for (int j = 0; j < S2_MAXP; j++) { zs[j] = 0; ws[j] = 0; }
    double t1 = 0, r0 = 0, gsum = 0, nz = 0, gcase = 0, ncase = 0, gctrl = 0, nctrl = 0, case_hom = 0, case_het = 0, ctrl_hom = 0, ctrl_het = 0;
    if (IDENT) {
        // one warp per 32 consecutive samples (two PLINK words); groups without a minor allele are skipped
        const int lane = tid & 31, warp = tid >> 5;
        const uint32_t *wrow = reinterpret_cast<const uint32_t *>(srow);
        const int64_t ng = (N + 31) >> 5, nw = (N + 15) >> 4;
        const uint32_t allzero = flip ? 0u : 0xFFFFFFFFu;          // 16 x "no copy of the tested allele"
        for (int64_t gI = warp; gI < ng; gI += S2_THREADS / 32) {
            const uint32_t x0 = wrow[2 * gI], x1 = (2 * gI + 1 < nw) ? wrow[2 * gI + 1] : allzero;
            if (x0 == allzero && x1 == allzero) continue;
            const int64_t i = (gI << 5) + lane;
            const uint32_t x = lane < 16 ? x0 : x1;
            const int code = (int)((x >> ((lane & 15) << 1)) & 3u);
            int g = code == 0 ? 2 : (code == 2 ? 1 : (code == 3 ? 0 : -1));
            g = g < 0 ? imputeG : (flip ? 2 - g : g);
            if (i < N && g) {
                const double gd = (double)g, m2 = M.mu2[i];
                t1 += m2 * gd * gd;
                r0 += M.res[i] * gd;
                for (int j = 0; j < p; j++) {
                    zs[j] += M.A[i + (int64_t)j * N] * gd;
                    ws[j] += m2 * M.X[i + (int64_t)j * N] * gd;
                }
            }
        }
        t1 = block_sum(t1, red); r0 = block_sum(r0, red);
        // every tally follows from the class counts of the first pass (in the flipped / imputed coding)
        ncase = M.ncase_tot; nctrl = (double)N - ncase;
        const double k2t = k2a - k2c, k1t = k1a - k1c, kmt = nMiss - kmc;            // controls
        const double h2c = flip ? ncase - k2c - k1c - kmc : k2c, h2t = flip ? nctrl - k2t - k1t - kmt : k2t;   // two copies after the flip
        case_hom = h2c + (imputeG == 2 ? kmc : 0.0); case_het = k1c + (imputeG == 1 ? kmc : 0.0);
        ctrl_hom = h2t + (imputeG == 2 ? kmt : 0.0); ctrl_het = k1t + (imputeG == 1 ? kmt : 0.0);
        gcase = 2.0 * case_hom + case_het; gctrl = 2.0 * ctrl_hom + ctrl_het;
        gsum = gcase + gctrl; nz = case_hom + case_het + ctrl_hom + ctrl_het;
    } else if (DOSE) {
        for (int64_t i = tid; i < N; i += S2_THREADS) {
            const double gd = GENO(i);
            const double yi = M.y[i];
            const double hom = (gd >= 1.5 && gd <= 2.0) ? 1.0 : 0.0, het = (gd >= 0.5 && gd < 1.5) ? 1.0 : 0.0;      // Main.cpp:510-520
            if (yi == 1.0) { ncase += 1; gcase += gd; case_hom += hom; case_het += het; }
            else { nctrl += 1; gctrl += gd; ctrl_hom += hom; ctrl_het += het; }
            if (gd != 0.0) {
                const double m2 = M.mu2[i];
                gsum += gd; nz += 1;
                t1 += m2 * gd * gd;
                r0 += M.res[i] * gd;
                for (int j = 0; j < p; j++) {
                    zs[j] += M.A[i + (int64_t)j * N] * gd;
                    ws[j] += m2 * M.X[i + (int64_t)j * N] * gd;
                }
            }
        }
        t1 = block_sum(t1, red); r0 = block_sum(r0, red); gsum = block_sum(gsum, red); nz = block_sum(nz, red);
        gcase = block_sum(gcase, red); ncase = block_sum(ncase, red); gctrl = block_sum(gctrl, red); nctrl = block_sum(nctrl, red);
        case_hom = block_sum(case_hom, red); case_het = block_sum(case_het, red);
        ctrl_hom = block_sum(ctrl_hom, red); ctrl_het = block_sum(ctrl_het, red);
    } else {
        for (int64_t i = tid; i < N; i += S2_THREADS) {
            const int g = s2_geno(srow, M.pos[i], flip, imputeG);
            const double yi = M.y[i];
            if (yi == 1.0) { ncase += 1; gcase += g; case_hom += g == 2; case_het += g == 1; }
            else { nctrl += 1; gctrl += g; ctrl_hom += g == 2; ctrl_het += g == 1; }
            if (g) {
                const double gd = (double)g, m2 = M.mu2[i];
                gsum += gd; nz += 1;
                t1 += m2 * gd * gd;
                r0 += M.res[i] * gd;
                for (int j = 0; j < p; j++) {
                    zs[j] += M.A[i + (int64_t)j * N] * gd;
                    ws[j] += m2 * M.X[i + (int64_t)j * N] * gd;
                }
            }
        }
        t1 = block_sum(t1, red); r0 = block_sum(r0, red); gsum = block_sum(gsum, red); nz = block_sum(nz, red);
        gcase = block_sum(gcase, red); ncase = block_sum(ncase, red); gctrl = block_sum(gctrl, red); nctrl = block_sum(nctrl, red);
        case_hom = block_sum(case_hom, red); case_het = block_sum(case_het, red);
        ctrl_hom = block_sum(ctrl_hom, red); ctrl_het = block_sum(ctrl_het, red);
    }
    for (int j = 0; j < p; j++) {
        double z = block_sum(zs[j], red), w = block_sum(ws[j], red);
        if (tid == 0) { Zs[j] = z; Ws[j] = w; }
    }
    __syncthreads();
    // altFreq / altCounts after imputation (UTIL.cpp:112-118)
    double altCount = gsum;
    altFreq = altCount / (2.0 * (double)N);
    if (flip) { altFreq = 1.0 - altFreq; altCount = 2.0 * (double)N - altCount; }

    // ---- scoreTestFast ----
    double zxz = 0, zw = 0, saz = 0;
    for (int a = 0; a < p; a++) {
        double acc = 0;
        for (int b = 0; b < p; b++) acc += M.XVX[a + b * p] * Zs[b];
        zxz += Zs[a] * acc;
        zw += Zs[a] * Ws[a];
        saz += M.S_a[a] * Zs[a];
    }
    double var2;
    if (M.binary) var2 = zxz + t1 - 2.0 * zw;
    else {
        // quantitative (SAIGE_test.cpp:246-248): ZtXVXZ*tau0 + g.g - 2 g.B ; mu2 = 1/tau0 constant => g.g = t1*tau0, g.B = zw*tau0
        var2 = zxz * M.tau0 + t1 * M.tau0 - 2.0 * zw * M.tau0;
    }
    // variance ratio of this variant's MAC category (Main.cpp:395-404 -> assignVarianceRatio; MAC after imputation, Main.cpp:367)
    const double MACafter = fmin(altCount, 2.0 * (double)N - altCount);
    double varRatio = M.varRatio;
    if (M.n_cate > 1) {
        varRatio = M.cate_ratio[M.n_cate - 1];                          // above the last bound
        for (int c = M.n_cate - 2; c >= 0; c--)
            if (MACafter <= M.cate_max[c]) varRatio = M.cate_ratio[c];  // also MAC <= cate_min[0] -> first category
    }
    const double var1 = var2 * varRatio;
    const double S = (r0 - saz) / M.tau0;
    double stat = S * S / var1;
    double pval_noadj, lp_noadj = 0.0;                                       // p-value and its natural log (finite when p underflows)
    if (var1 <= 2.2250738585072014e-308) pval_noadj = 1.0;
    else if (isfinite(stat)) { pval_noadj = erfc(sqrt(stat * 0.5)); lp_noadj = s2_log_erfc(sqrt(stat * 0.5)); }   // chi-square(1) upper tail
    else { pval_noadj = 1.0; stat = 0.0; }
    const double Beta = S / var1;
    double seBeta = fabs(Beta) / sqrt(fabs(stat));
    double pval = pval_noadj, lp = lp_noadj, isSPA = 0.0;

    const double StdStat = fabs(S) / sqrt(var1);
    // ---- exact test of rare variants (binary traits): MAC after imputation <= max_MAC_for_ER and a score beyond the SPA
    // cutoff (Main.cpp:408-422, SAIGE_test.cpp:426-431).  The carriers (<= SGB_ER_MAXK of them, since every one holds at
    // least one minor allele) are collected in sample order; one thread enumerates their case / control assignments ----
    const bool isER = M.binary && MACafter <= M.er_max_mac && nz <= (double)SGB_ER_MAXK && (StdStat > M.spa_cutoff || isnan(StdStat));
    if (isER) {
        if (tid == 0) er_cnt = 0;
        __syncthreads();
        for (int64_t i = tid; i < N; i += S2_THREADS) {
            if (GENO(i) != 0.0) {
                const int slot = atomicAdd(&er_cnt, 1);
                if (slot < SGB_ER_MAXK) er_idx[slot] = (int)i;
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int k = er_cnt < SGB_ER_MAXK ? er_cnt : SGB_ER_MAXK;
            for (int a = 1; a < k; a++) {                          // ascending sample index (iIndex of the reference)
                const int v = er_idx[a];
                int b = a - 1;
                while (b >= 0 && er_idx[b] > v) { er_idx[b + 1] = er_idx[b]; b--; }
                er_idx[b + 1] = v;
            }
            double g1[SGB_ER_MAXK], p1[SGB_ER_MAXK], r1[SGB_ER_MAXK], musum = 0.0;
            for (int a = 0; a < k; a++) {
                const int i = er_idx[a];
                g1[a] = GENO(i);
                p1[a] = M.mu[i]; r1[a] = M.res[i];
                musum += p1[a];
            }
            const double p2mean = (M.mu_sum - musum) / ((double)N - (double)k);
            // NResampling 2e6, ExactMax 1e4, epsilon 1e-6 (SAIGE_test.cpp:599): 2^k <= 1024 assignments, all enumerated
            er_pv = sgb_er_exact_pvalue(k, g1, p1, r1, p2mean, (double)N, M.ncase_tot, 1e-6);
        }
        __syncthreads();
        pval = er_pv; lp = log(er_pv);
        // SE from the exact p-value, |qnorm(p/2)| (SAIGE_test.cpp:606-614; quantile(0) overflows there -> 0)
        seBeta = pval * 0.5 > 0.0 ? fabs(Beta) / fabs(normcdfinv(pval * 0.5)) : 0.0;
    }
    // ---- conditional analysis (t_isCondition, SAIGE_test.cpp:640-660): score and variance after projecting out the
    // conditioning markers.  gtilde^T P2 = g^T P2 - W^T (XXVX_inv^T P2): one pass over the non-zero genotypes ----
    double Tc = nan(""), vc = nan(""), Beta_c = nan(""), se_c = nan(""), pval_c = nan(""), pval_noadj_c = nan(""), stat_c = 0.0;
    double lp_c = nan(""), lp_noadj_c = nan("");
    if (M.n_cond > 0) {
        double cp[S2_MAXCOND];
#pragma unroll
        for (int c = 0; c < S2_MAXCOND; c++) cp[c] = 0.0;
        for (int64_t i = tid; i < N; i += S2_THREADS) {
            const double g = GENO(i);
            if (g != 0.0)
                for (int c = 0; c < M.n_cond; c++) cp[c] += M.P2[i + (int64_t)c * N] * g;
        }
        double g1p2[S2_MAXCOND];
        const double srv = sqrt(varRatio);
        for (int c = 0; c < M.n_cond; c++) {
            double v = block_sum(cp[c], red);
            for (int j = 0; j < p; j++) v -= Ws[j] * M.XtP2[j + c * p];
            g1p2[c] = srv * v;
        }
        Tc = S; vc = var1;
        for (int c = 0; c < M.n_cond; c++) {
            Tc -= g1p2[c] * M.VT[c];
            double acc = 0.0;
            for (int d = 0; d < M.n_cond; d++) acc += M.VarInv[c + d * M.n_cond] * g1p2[d];
            vc -= g1p2[c] * acc;
        }
        stat_c = Tc * Tc / vc;
        lp_noadj_c = 0.0;
        if (vc <= 2.2250738585072014e-308) { pval_noadj_c = 1.0; stat_c = 0.0; }
        else if (isfinite(stat_c)) { pval_noadj_c = erfc(sqrt(stat_c * 0.5)); lp_noadj_c = s2_log_erfc(sqrt(stat_c * 0.5)); }
        else { pval_noadj_c = 1.0; stat_c = 0.0; }
        Beta_c = Tc / vc;
        se_c = fabs(Beta_c) / sqrt(stat_c);
        pval_c = pval_noadj_c; lp_c = lp_noadj_c;
    }
    // ---- saddle-point approximation (binary traits): the marginal test when |T|/sqrt(var) > cutoff and the exact test did
    // not take the variant, the conditional test when its own statistic exceeds cutoff^2 (SAIGE_test.cpp:699).  The
    // reference runs the conditional SPA on whatever the marginal block left behind (its NAmu / NAsigma are only set
    // there); here the shared quantities are computed whenever either test needs them ----
    const bool spa_u = !isER && M.binary && isfinite(StdStat) && StdStat > M.spa_cutoff;
    const bool spa_c = M.n_cond > 0 && M.binary && stat_c > M.spa_cutoff * M.spa_cutoff;
    if (spa_u || spa_c) {
        // gtilde_i = g_i - XXVX_inv[i,:] . (XV g),  XV g = W  (getadjGFast, SAIGE_test.cpp:306-315)
        double m1p = 0, gpos = 0, gneg = 0, gmuNB = 0, sigNB = 0;
        const int fast = ((double)N - nz) / (double)N >= 0.5;
        const bool use_scr = scr.gt != nullptr;
        int nnz = 0;                                     // entries of scr.nz (fast mode with scratch)
        // the loop runs over whole 256-sample blocks so that the compaction's block scan sees uniform control flow
        for (int64_t i0 = 0; i0 < N; i0 += S2_THREADS) {
            const int64_t i = i0 + tid;
            double g = 0.0, gt = 0.0, mu = 0.0;
            if (i < N) {
                g = GENO(i);
                gt = g;
                for (int j = 0; j < p; j++) gt -= M.XXVXi[i + (int64_t)j * N] * Ws[j];
                mu = M.mu[i];
                m1p += mu * gt;
                if (gt > 0) gpos += gt; else if (gt < 0) gneg += gt;
                if (g != 0.0) { gmuNB += gt * mu; sigNB += mu * (1.0 - mu) * gt * gt; }
                if (use_scr && !fast) scr.gt[i] = gt;
            }
            if (use_scr && fast) {
                // deterministic compaction in sample order: ballot inside the warp, exclusive scan over the 8 warps
                const bool keep = i < N && g != 0.0;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if ((tid & 31) == 0) nz_warp[tid >> 5] = __popc(bal);
                __syncthreads();
                int base = nnz, total = 0;
#pragma unroll
                for (int w = 0; w < S2_THREADS / 32; w++) { if (w < (tid >> 5)) base += nz_warp[w]; total += nz_warp[w]; }
                if (keep) scr.nz[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = make_double2(gt, mu);
                nnz += total;
                __syncthreads();
            }
        }
        const double m1 = block_sum(m1p, red);
        gpos = block_sum(gpos, red); gneg = block_sum(gneg, red); gmuNB = block_sum(gmuNB, red); sigNB = block_sum(sigNB, red);
        const double NAmu = m1 - gmuNB, NAsigma = var2 - sigNB;
        const double tol = 1.220703125e-4;          // eps^0.25 (SAIGE_test.cpp:515-516)

        // CGF sums at t over the samples that enter exactly: all samples (SPA) or the non-zero genotypes (SPA_fast)
        // K0 is only needed at the root (Lugannani-Rice), the Newton iterations use K1 and K2: one exponential and one
        // reciprocal per sample there (fp64 exp / log / divide are long software sequences)
        auto cgf = [&](double t, double &k0, double &k1, double &k2, bool want0) {
            double a0 = 0, a1 = 0, a2 = 0;
            const int64_t nit = use_scr && fast ? (int64_t)nnz : N;
            for (int64_t i = tid; i < nit; i += S2_THREADS) {
                double gt, mu;
                if (use_scr) {
                    if (fast) { const double2 v = scr.nz[i]; gt = v.x; mu = v.y; }
                    else { gt = scr.gt[i]; mu = M.mu[i]; }
                } else {
                    const double g = GENO(i);
                    if (fast && g == 0.0) continue;
                    gt = g;
                    for (int j = 0; j < p; j++) gt -= M.XXVXi[i + (int64_t)j * N] * Ws[j];
                    mu = M.mu[i];
                }
                const double x = gt * t;
                const double e = exp(-x);
                const double den = (1.0 - mu) * e + mu;
                const double r = 1.0 / den;
                if (want0) a0 += x > 0 ? x + log(den) : log(1.0 - mu + mu / e);      // log(1 - mu + mu exp(x)), overflow-safe
                a1 += mu * gt * r;
                a2 += (1.0 - mu) * mu * gt * gt * e * (r * r);
            }
            k0 = want0 ? block_sum(a0, red) : 0.0; k1 = block_sum(a1, red); k2 = block_sum(a2, red);
            if (fast) { k0 += NAmu * t + 0.5 * NAsigma * t * t; k1 += NAmu + NAsigma * t; k2 += NAsigma; }
        };
        // SPA / SPA_fast (SPA.cpp:20-185) for the statistic q: both tails; false when a root or a saddle point fails
        // log-scale twin of every probability: the reference runs this branch with logp = TRUE when the unadjusted p-value
        // underflowed (SPA.cpp:20-110, Get_Saddle_Prob_Binom's R::pnorm(..., logp)); here both scales are always carried
        auto spa = [&](double q, double pnoadj, double lpnoadj, double &pspa, double &lpspa) -> bool {
            double qinv;
            if (q - m1 > 0) qinv = -fabs(q - m1) + m1; else if (q - m1 == 0) qinv = m1; else qinv = fabs(q - m1) + m1;
            double pside[2], lside[2]; bool conv_all = true, saddle_all = true;
            for (int side = 0; side < 2; side++) {
                const double qq = side == 0 ? q : qinv;
                double root; bool conv = true;
                if (qq >= gpos || qq <= gneg) root = INFINITY;
                else {
                    // getroot_K1[_fast]_Binom (SPA_binary.cpp:70-140, 214-270), init 0, maxiter 1000
                    double t = 0.0, k0, k1, k2, prevJump = INFINITY;
                    cgf(t, k0, k1, k2, false);
                    double K1e = k1 - qq;
                    int rep = 1;
                    while (true) {
                        double tnew = t - K1e / k2;
                        if (isnan(tnew)) { conv = false; break; }
                        if (fabs(tnew - t) < tol) { conv = true; break; }
                        if (rep == 1000) { conv = false; break; }
                        double n0, n1, n2;
                        cgf(tnew, n0, n1, n2, false);
                        double newK1 = n1 - qq;
                        const bool changed = fast ? (K1e * newK1 < 0) : ((K1e > 0) - (K1e < 0)) != ((newK1 > 0) - (newK1 < 0));
                        if (changed) {
                            if (fabs(tnew - t) > prevJump - tol) {
                                const double d = newK1 - K1e;
                                tnew = t + ((d > 0) - (d < 0)) * prevJump / 2;
                                cgf(tnew, n0, n1, n2, false);
                                newK1 = n1 - qq;
                                prevJump = prevJump / 2;
                            } else prevJump = fabs(tnew - t);
                        }
                        rep++; t = tnew; K1e = newK1; k2 = n2;
                    }
                    root = t;
                }
                if (!conv) { conv_all = false; break; }
                // Get_Saddle_Prob[_fast]_Binom (SPA_binary.cpp:146-214, 276-330): Lugannani-Rice
                double k0, k1, k2;
                double ps = 0.0, lps = -INFINITY; bool isSaddle = false;
                if (isfinite(root)) {
                    cgf(root, k0, k1, k2, true);
                    const double temp1 = root * qq - k0;
                    if (isfinite(k0) && isfinite(k2) && temp1 >= 0 && k2 >= 0) {
                        const double w = ((root > 0) - (root < 0)) * sqrt(2.0 * temp1), v = root * sqrt(k2);
                        if (w != 0) {
                            const double Zt = w + log(v / w) / w;
                            ps = Zt > 0 ? 0.5 * erfc(Zt * 0.7071067811865476) : -0.5 * erfc(-Zt * 0.7071067811865476);
                            lps = -0.6931471805599453 + s2_log_erfc(fabs(Zt) * 0.7071067811865476);      // log |ps|
                            isSaddle = true;
                        }
                    }
                }
                if (!isSaddle) { saddle_all = false; ps = pnoadj / 2; lps = lpnoadj - 0.6931471805599453; }
                pside[side] = ps; lside[side] = lps;
            }
            if (!conv_all) return false;
            pspa = fabs(pside[0]) + fabs(pside[1]);
            lpspa = s2_logaddexp(lside[0], lside[1]);
            // a vanished adjusted p-value un-converges the test only on the linear scale (SAIGE_test.cpp:541): when the
            // unadjusted p-value already underflowed the reference is on the log scale and keeps the saddle-point result
            return saddle_all && (pspa != 0 || pnoadj == 0);
        };
        if (spa_u) {
            double pspa, lpspa;
            if (spa(S / sqrt(var1 / var2) + m1, pval_noadj, lp_noadj, pspa, lpspa)) {
                isSPA = 1.0; pval = pspa; lp = lpspa;
                // SE from the SPA p-value.  se_two_sided: |qnorm(p/2)| (what produced the reference's bundled golden tables);
                // otherwise qnorm(p, upper tail) as written in this fork's source (SAIGE_test.cpp:523-526).
                const double qv = s2_qnorm_from_logp(se_two_sided ? lpspa - 0.6931471805599453 : lpspa);
                seBeta = fabs(Beta) / qv;
            }
        }
        if (spa_c) {
            // the reference's bundled conditional table reports the adjusted p-value itself and SE = |BETA_c| / |qnorm(p/2)|
            // (this fork's source prints half of it, SAIGE_test.cpp:752: the fixture wins)
            double pspa, lpspa;
            if (spa(Tc / sqrt(vc / var2) + m1, pval_noadj_c, lp_noadj_c, pspa, lpspa)) {
                pval_c = pspa; lp_c = lpspa;
                se_c = fabs(Beta_c) / s2_qnorm_from_logp(lpspa - 0.6931471805599453);
            }
        }
    }
    // ---- Firth's penalised-likelihood refit of the effect size (binary traits, p <= pCutoffforFirth) ----
    // x = [1, gtilde], offset = the null model's; Newton steps with the modified score x^T((y - pi) + h (0.5 - pi)), h the hat
    // values of sqrt(W) x; every quantity is a 2 x 2 / 2-vector block reduction over the samples (SAIGE_test.cpp:893-986)
    double BetaOut = Beta, isFirth = 0.0, firthConv = 0.0;
    if (M.firth && M.binary && lp <= log(M.firth_cutoff)) {            // on the log scale: also right when p underflowed (SAIGE_test.cpp:572-582)
        isFirth = 1.0;
        double b0 = 0.0, b1 = 0.0, c00 = nan(""), c01 = nan(""), c11 = nan("");
        int iter = 0;
        while (iter <= 50) {
            double f00 = 0, f01 = 0, f11 = 0;
            for (int64_t i = tid; i < N; i += S2_THREADS) {
                double gt = GENO(i);
                for (int j = 0; j < p; j++) gt -= M.XXVXi[i + (int64_t)j * N] * Ws[j];
                const double pi = 1.0 / (exp(-(b0 + b1 * gt) - M.offset[i]) + 1.0), w = pi * (1.0 - pi);
                f00 += w; f01 += w * gt; f11 += w * gt * gt;
            }
            f00 = block_sum(f00, red); f01 = block_sum(f01, red); f11 = block_sum(f11, red);
            const double det = f00 * f11 - f01 * f01;
            if (!(det > 0.0) || !(f00 > 0.0)) break;                      // inv_sympd fails
            c00 = f11 / det; c01 = -f01 / det; c11 = f00 / det;
            double u0 = 0, u1 = 0;
            for (int64_t i = tid; i < N; i += S2_THREADS) {
                double gt = GENO(i);
                for (int j = 0; j < p; j++) gt -= M.XXVXi[i + (int64_t)j * N] * Ws[j];
                const double pi = 1.0 / (exp(-(b0 + b1 * gt) - M.offset[i]) + 1.0), w = pi * (1.0 - pi);
                const double hat = w * (c00 + 2.0 * c01 * gt + c11 * gt * gt);
                const double r = (M.y[i] - pi) + hat * (0.5 - pi);
                u0 += r; u1 += gt * r;
            }
            u0 = block_sum(u0, red); u1 = block_sum(u1, red);
            double d0 = c00 * u0 + c01 * u1, d1 = c01 * u0 + c11 * u1;
            const double mx = fmax(fabs(d0), fabs(d1)) / 15.0;
            if (mx > 1.0) { d0 /= mx; d1 /= mx; }
            iter++;
            b0 += d0; b1 += d1;
            if (iter == 50 || (fmax(fabs(d0), fabs(d1)) <= 1e-5 && fabs(u0) <= 1e-5 && fabs(u1) <= 1e-5)) { firthConv = 1.0; break; }
        }
        if (isnan(c00) || isnan(c01) || isnan(c11)) { BetaOut = nan(""); seBeta = nan(""); }
        else {
            BetaOut = b1;
            // SE: the fit's own (what the reference's bundled positive-signal result holds) or |beta| / |qnorm| of the p-value
            // as this fork's source has it (SAIGE_test.cpp:632)
            seBeta = M.firth_se_from_fit ? sqrt(c11) : fabs(b1) / s2_qnorm_from_logp((se_two_sided || isER) ? lp - 0.6931471805599453 : lp);
        }
    }
    if (tid == 0) {
        const double sgn = flip ? -1.0 : 1.0;
        double afc = ncase > 0 ? gcase / ncase / 2.0 : nan(""), aft = nctrl > 0 ? gctrl / nctrl / 2.0 : nan("");
        if (flip) { afc = 1.0 - afc; aft = 1.0 - aft; case_hom = ncase - case_het - case_hom; ctrl_hom = nctrl - ctrl_het - ctrl_hom; }
        o[0] = 1.0;            // tested
        o[1] = altCount; o[2] = altFreq; o[3] = missingRate;
        o[4] = sgn * BetaOut; o[5] = seBeta; o[6] = sgn * S; o[7] = var1; o[8] = pval; o[9] = pval_noadj; o[10] = isSPA;
        o[11] = afc; o[12] = aft; o[13] = ncase; o[14] = nctrl; o[15] = case_hom; o[16] = case_het; o[17] = ctrl_hom; o[18] = ctrl_het;
        o[19] = var2; o[20] = isFirth; o[21] = firthConv;
        o[22] = sgn * Beta_c; o[23] = se_c; o[24] = sgn * Tc; o[25] = vc; o[26] = pval_c; o[27] = pval_noadj_c;
        o[28] = lp; o[29] = lp_noadj; o[30] = lp_c; o[31] = lp_noadj_c;
    }
}

// one CTA per variant (per-variant path, dosage rows)
template <bool IDENT, bool DOSE>
__global__ void __launch_bounds__(S2_THREADS, IDENT ? 3 : 2)
step2_kernel(s2_model M, const uint8_t *__restrict__ bed, int64_t B0, int64_t nm, double min_maf, double min_mac,
             double max_missing, int se_two_sided, double *__restrict__ out, s2_dose DS)
{
    step2_variant<IDENT, DOSE>(M, bed, B0, nm, (int64_t)blockIdx.x, min_maf, min_mac, max_missing, se_two_sided, out, DS, s2_spa_scratch{nullptr, nullptr});
}

// batched path: persistent CTAs pull the flagged variants (saddle point / exact test / Firth / conditional) off a list; a
// saddle-point variant keeps a CTA busy for milliseconds, so dynamic assignment also balances the tail of a chunk
__global__ void __launch_bounds__(S2_THREADS, 3)
step2_flagged_kernel(s2_model M, const uint8_t *__restrict__ bed, int64_t B0, int64_t nm, double min_maf, double min_mac,
                     double max_missing, int se_two_sided, double *__restrict__ out, const int *__restrict__ list,
                     const int *__restrict__ list_count, int *__restrict__ next, double *__restrict__ scratch, int64_t slot_doubles)
{
    __shared__ int cur;
    s2_spa_scratch scr;
    scr.gt = scratch + (int64_t)blockIdx.x * slot_doubles;
    scr.nz = reinterpret_cast<double2 *>(scr.gt + ((M.N + 1) & ~(int64_t)1));
    const int count = *list_count;
    for (;;) {
        __syncthreads();                                  // the previous variant's shared state is no longer read
        if (threadIdx.x == 0) cur = atomicAdd(next, 1);
        __syncthreads();
        const int idx = cur;
        if (idx >= count) break;
        step2_variant<true, false>(M, bed, B0, nm, (int64_t)list[idx], min_maf, min_mac, max_missing, se_two_sided, out, s2_dose{}, scr);
    }
}

// ---------------------------------------------------------------------------------------------------
// Batched score sums on the tensor engine ("the score statistics are a skinny GEMM over a variant batch").
//
// For a chunk of variants the sums scoreTestFast needs,
//     Z = A^T g (p)   W = (mu2 X)^T g (p)   R0 = res.g   T1 = sum mu2 g^2 = mu2.g + 2 mu2.[g == 2]
// are "packed 2-bit rows x (2p + 2) fp64 columns" on the VALUE plane plus one column on the [g == 2] plane: the same
// contraction as a GRM sweep, with variants as rows and samples as the contraction index.  The chunk is re-packed on the
// device into the pair-ternary tiled store (missing calls filled with the variant's best-guess genotype, raw alt-allele
// coding), pk2_umma_kernel / pk2_stream_kernel contract it against limb images of the model columns that are built once
// per model, and one thread per variant finishes the test.  Allele flips are applied algebraically to the sums
// (g' = 2 - g is linear; [g' == 2] = 1 - g + [g == 2]).  Only variants that need a pass of their own -- saddle-point
// approximation (|z| > SPAcutoff), exact test, Firth, conditional analysis -- go through step2_kernel afterwards.
// ---------------------------------------------------------------------------------------------------
// identity-ordered raw rows for a model whose samples are a subset / permutation of the .fam
__global__ void s2_gather_kernel(const uint8_t *__restrict__ bed, int64_t B0, const int32_t *__restrict__ pos, int64_t N,
                                 int64_t B, uint8_t *__restrict__ out)
{
    const int64_t m = blockIdx.y;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint8_t *row = bed + m * B0;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t i = 4 * b + j;
        uint32_t code = 3u;                                  // padding: hom ref = genotype 0 of the tested (first) allele
        if (i < N) { const int32_t src = pos[i]; code = (row[src >> 2] >> ((src & 3) << 1)) & 3u; }
        v |= code << (2 * j);
    }
    out[m * B + b] = (uint8_t)v;
}

// class counts of one variant over the model's samples: hom-alt / het / missing among all and among cases (popcounts)
__global__ void __launch_bounds__(256) s2_count_kernel(const uint8_t *__restrict__ bed, int64_t B0, int64_t N,
                                                       const uint32_t *__restrict__ ycase, int32_t *__restrict__ cnt)
{
    __shared__ int sm[6][8];
    const int64_t m = blockIdx.x;
    const int64_t nw = (N + 15) >> 4;
    const uint8_t *row = bed + m * B0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 3) == 0);
    int c[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t w = threadIdx.x; w < nw; w += 256) {
        uint32_t x;
        if (aligned && 4 * w + 4 <= B0) x = *reinterpret_cast<const uint32_t *>(row + 4 * w);
        else { x = 0; for (int j = 0; j < 4; j++) if (4 * w + j < B0) x |= (uint32_t)row[4 * w + j] << (8 * j); }
        uint32_t valid = 0x55555555u;
        if (w == nw - 1 && (N & 15)) valid &= (1u << (2 * (N & 15))) - 1u;
        const uint32_t L = x & 0x55555555u, H = (x >> 1) & 0x55555555u;
        const uint32_t homalt = ~L & ~H & valid, het = H & ~L & valid, miss = L & ~H & valid;      // PLINK.hpp:48-56, alt-first
        const uint32_t yc = ycase[w];
        c[0] += __popc(homalt); c[1] += __popc(het); c[2] += __popc(miss);
        c[3] += __popc(homalt & yc); c[4] += __popc(het & yc); c[5] += __popc(miss & yc);
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c[q] += __shfl_xor_sync(0xffffffffu, c[q], o);
        if ((threadIdx.x & 31) == 0) sm[q][threadIdx.x >> 5] = c[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        int t = 0;
        for (int i = 0; i < 8; i++) t += sm[threadIdx.x][i];
        cnt[m * 6 + threadIdx.x] = t;
    }
}

struct s2_head { double altFreq0, missingRate, altFreq; int pass, flip, imputeG; };
// getOneMarker's frequencies, the filter of Main.cpp:296 and imputeGenoAndFlip's decisions from the class counts
__device__ __forceinline__ s2_head s2_head_of(const int32_t *c, int64_t N, double min_maf, double min_mac, double max_missing)
{
    s2_head h;
    const double k2a = c[0], k1a = c[1], nMiss = c[2];
    const double altCounts0 = 2.0 * k2a + k1a, cntv = (double)N - nMiss;
    h.altFreq0 = cntv > 0 ? altCounts0 / cntv / 2.0 : 0.0;
    h.missingRate = nMiss / (double)N;
    const double MAF = fmin(h.altFreq0, 1.0 - h.altFreq0);
    const double MAC0 = MAF * (double)N * (1.0 - h.missingRate) * 2.0;
    h.pass = !(h.missingRate > max_missing || MAF < min_maf || MAC0 < min_mac);
    h.flip = h.altFreq0 > 0.5;
    h.altFreq = h.flip ? 1.0 - h.altFreq0 : h.altFreq0;
    h.imputeG = nMiss > 0 ? (int)round(2.0 * h.altFreq) : 0;
    return h;
}

// raw PLINK rows -> pair-ternary tiled store (rows = variants, contraction index = model samples), raw alt-allele coding,
// missing calls = the best-guess genotype of the variant; rows beyond nm and samples beyond N are genotype 0.
// One thread per 32-bit word (16 samples); byte translation through a shared-memory table [fill][byte].
__global__ void __launch_bounds__(256) s2_repack_kernel(const uint8_t *__restrict__ bed, int64_t B0, int64_t N, int64_t nm,
                                                        const int32_t *__restrict__ cnt, double min_maf, double min_mac,
                                                        double max_missing, uint8_t *__restrict__ out, int64_t stride)
{
    __shared__ uint8_t lut[3][256];
    for (int idx = threadIdx.x; idx < 768; idx += 256) {
        const int fill = idx >> 8, byte = idx & 255;
        int g[4];
        for (int j = 0; j < 4; j++) {
            const int code = (byte >> (2 * j)) & 3;
            g[j] = code == 0 ? 2 : (code == 2 ? 1 : (code == 3 ? 0 : fill));
        }
        lut[fill][byte] = (uint8_t)((g[0] + 3 * g[1]) | ((g[2] + 3 * g[3]) << 4));
    }
    __syncthreads();
    const int64_t m = blockIdx.y;
    const int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (4 * w >= stride) return;
    uint32_t o = 0;
    if (m < nm) {
        const s2_head hd = s2_head_of(cnt + m * 6, N, min_maf, min_mac, max_missing);
        const int fill_raw = hd.flip ? 2 - hd.imputeG : hd.imputeG;
        const uint8_t *row = bed + m * B0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t b = 4 * w + j;
            uint32_t byte = b < B0 ? row[b] : 0xFFu;
            const int64_t i0 = 4 * b;
            if (i0 + 4 > N) {                                     // samples beyond N: code 11 = genotype 0
                for (int q = 0; q < 4; q++) if (i0 + q >= N) byte |= 3u << (2 * q);
            }
            o |= (uint32_t)lut[fill_raw][byte] << (8 * j);
        }
    }
    *reinterpret_cast<uint32_t *>(out + sgb_tiled_off(m, 4 * w, stride)) = o;
}

// One thread per variant: scoreTestFast from the batched sums.  rawV [rows_pad x kv] (ld rows_pad): columns
// A (p) | mu2 X (p) | res | mu2 over the value plane, rawI [rows_pad]: mu2 over the [g == 2] plane.  csum[kv]: the
// columns' sums over all samples (flip algebra).  Variants that need their own pass are appended to `list`.
__global__ void __launch_bounds__(128) s2_finish_kernel(s2_model M, const int32_t *__restrict__ cnt, const double *__restrict__ rawV,
                                                        const double *__restrict__ rawI, int64_t ld, const double *__restrict__ csum,
                                                        int64_t nm, double min_maf, double min_mac, double max_missing,
                                                        double *__restrict__ out, int *__restrict__ list, int *__restrict__ list_count)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nm) return;
    const int p = M.p;
    const int64_t N = M.N;
    double *o = out + m * S2_NOUT;
    const int32_t *c = cnt + m * 6;
    const s2_head hd = s2_head_of(c, N, min_maf, min_mac, max_missing);
    if (!hd.pass) {
        for (int q = 0; q < S2_NOUT; q++) o[q] = nan("");
        o[0] = 0.0; o[2] = hd.altFreq0; o[3] = hd.missingRate;
        return;
    }
    const int flip = hd.flip, imputeG = hd.imputeG;
    const double k2a = c[0], k1a = c[1], nMiss = c[2], k2c = c[3], k1c = c[4], kmc = c[5];
    // tallies in the flipped / imputed coding (same algebra as step2_kernel's identity path)
    const double ncase = M.ncase_tot, nctrl = (double)N - ncase;
    const double k2t = k2a - k2c, k1t = k1a - k1c, kmt = nMiss - kmc;
    const double h2c = flip ? ncase - k2c - k1c - kmc : k2c, h2t = flip ? nctrl - k2t - k1t - kmt : k2t;
    double case_hom = h2c + (imputeG == 2 ? kmc : 0.0), case_het = k1c + (imputeG == 1 ? kmc : 0.0);
    double ctrl_hom = h2t + (imputeG == 2 ? kmt : 0.0), ctrl_het = k1t + (imputeG == 1 ? kmt : 0.0);
    const double gcase = 2.0 * case_hom + case_het, gctrl = 2.0 * ctrl_hom + ctrl_het;
    const double gsum = gcase + gctrl, nz = case_hom + case_het + ctrl_hom + ctrl_het;
    // sums in the tested coding
    double Z[S2_MAXP], W[S2_MAXP];
    for (int j = 0; j < p; j++) {
        const double zr = rawV[m + (int64_t)j * ld], wr = rawV[m + (int64_t)(p + j) * ld];
        Z[j] = flip ? 2.0 * csum[j] - zr : zr;
        W[j] = flip ? 2.0 * csum[p + j] - wr : wr;
    }
    const double rr = rawV[m + (int64_t)(2 * p) * ld], mr = rawV[m + (int64_t)(2 * p + 1) * ld], ir = rawI[m];
    const double r0 = flip ? 2.0 * csum[2 * p] - rr : rr;
    const double t1 = flip ? 4.0 * csum[2 * p + 1] - 3.0 * mr + 2.0 * ir : mr + 2.0 * ir;
    double altCount = gsum, altFreq = altCount / (2.0 * (double)N);
    if (flip) { altFreq = 1.0 - altFreq; altCount = 2.0 * (double)N - altCount; }
    double zxz = 0, zw = 0, saz = 0;
    for (int a = 0; a < p; a++) {
        double acc = 0;
        for (int b = 0; b < p; b++) acc += M.XVX[a + b * p] * Z[b];
        zxz += Z[a] * acc;
        zw += Z[a] * W[a];
        saz += M.S_a[a] * Z[a];
    }
    const double var2 = M.binary ? zxz + t1 - 2.0 * zw : zxz * M.tau0 + t1 * M.tau0 - 2.0 * zw * M.tau0;
    const double MACafter = fmin(altCount, 2.0 * (double)N - altCount);
    double varRatio = M.varRatio;
    if (M.n_cate > 1) {
        varRatio = M.cate_ratio[M.n_cate - 1];
        for (int q = M.n_cate - 2; q >= 0; q--)
            if (MACafter <= M.cate_max[q]) varRatio = M.cate_ratio[q];
    }
    const double var1 = var2 * varRatio;
    const double S = (r0 - saz) / M.tau0;
    double stat = S * S / var1;
    double pval_noadj, lp_noadj = 0.0;
    if (var1 <= 2.2250738585072014e-308) pval_noadj = 1.0;
    else if (isfinite(stat)) { pval_noadj = erfc(sqrt(stat * 0.5)); lp_noadj = s2_log_erfc(sqrt(stat * 0.5)); }
    else { pval_noadj = 1.0; stat = 0.0; }
    const double Beta = S / var1;
    const double seBeta = fabs(Beta) / sqrt(fabs(stat));
    const double StdStat = fabs(S) / sqrt(var1);
    const bool isER = M.binary && MACafter <= M.er_max_mac && nz <= (double)SGB_ER_MAXK && (StdStat > M.spa_cutoff || isnan(StdStat));
    const bool spa_u = !isER && M.binary && isfinite(StdStat) && StdStat > M.spa_cutoff;
    const bool firth = M.firth && M.binary && lp_noadj <= log(M.firth_cutoff);
    if (isER || spa_u || firth || M.n_cond > 0) {
        list[atomicAdd(list_count, 1)] = (int)m;          // step2_kernel writes this row
        return;
    }
    const double sgn = flip ? -1.0 : 1.0;
    double afc = ncase > 0 ? gcase / ncase / 2.0 : nan(""), aft = nctrl > 0 ? gctrl / nctrl / 2.0 : nan("");
    if (flip) { afc = 1.0 - afc; aft = 1.0 - aft; case_hom = ncase - case_het - case_hom; ctrl_hom = nctrl - ctrl_het - ctrl_hom; }
    o[0] = 1.0; o[1] = altCount; o[2] = altFreq; o[3] = hd.missingRate;
    o[4] = sgn * Beta; o[5] = seBeta; o[6] = sgn * S; o[7] = var1; o[8] = pval_noadj; o[9] = pval_noadj; o[10] = 0.0;
    o[11] = afc; o[12] = aft; o[13] = ncase; o[14] = nctrl; o[15] = case_hom; o[16] = case_het; o[17] = ctrl_hom; o[18] = ctrl_het;
    o[19] = var2; o[20] = 0.0; o[21] = 0.0;
    for (int q = 22; q < S2_NOUT; q++) o[q] = nan("");
    o[28] = lp_noadj; o[29] = lp_noadj;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct sgb_step2 {
    s2_model M;
    double *d_vec = nullptr;      // mu | mu2 | res | y | X | A | XXVXi
    int32_t *d_pos = nullptr;
    uint32_t *d_ycase = nullptr;
    double *d_offset = nullptr;
    double *d_P2 = nullptr;
    uint8_t *d_bed = nullptr; size_t bed_bytes = 0;
    double *d_out = nullptr; size_t out_elems = 0;
    uint8_t *pin[2] = {nullptr, nullptr}; size_t pin_bytes = 0;
    double *pout[2] = {nullptr, nullptr}; size_t pout_bytes = 0;      // capacities tracked apart: the chunk length depends on n_fam
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;               // H2D of chunk c+1 runs here while h->stream works on chunk c
    cudaEvent_t ev_up[2] = {nullptr, nullptr};        // "rows of buffer i have arrived" (recorded on copy_stream)
    // batched score sums (tensor engine): limb images of the model columns, built once per model
    int kv = 0;                           // value-plane columns: A (p) | mu2 X (p) | res | mu2
    int64_t stride = 0;                   // packed bytes per variant row of the tiled chunk store (multiple of 64)
    double *d_V = nullptr;                // N x kv model columns (device)
    double *d_csum = nullptr;             // kv column sums over all samples
    int8_t *d_Lv = nullptr, *d_Li = nullptr;          // limb images: tcgen05 (kv columns, value plane) and mma.sync (mu2, [g==2] plane)
    double *d_multv = nullptr, *d_multi = nullptr;    // fixed-point multipliers of the images
    int32_t *d_lsv = nullptr, *d_lsi = nullptr;       // exact limb column sums
    uint8_t *d_gath = nullptr; size_t gath_bytes = 0; // identity-ordered raw rows (models on a subset / permutation of the .fam)
    uint8_t *d_tiled = nullptr; size_t tiled_bytes = 0;
    int32_t *d_accv = nullptr; size_t accv_elems = 0; // int32 limb accumulators (zero between chunks)
    int32_t *d_acci = nullptr; size_t acci_elems = 0;
    double *d_raw = nullptr; size_t raw_elems = 0;    // recombined sums: [rows_pad x kv] | [rows_pad]
    int32_t *d_cnt = nullptr; size_t cnt_elems = 0;   // class counts, 6 per variant
    int *d_list = nullptr; size_t list_elems = 0;     // [count | next | variant indices] of the flagged variants
    double *d_spa = nullptr; size_t spa_bytes = 0;    // saddle-point scratch: one slot (2 N doubles) per persistent CTA
    bool batched = true;                  // sgb_step2_set_batched(0): every variant through step2_kernel (cross-check)
    int64_t chunk_bytes = (int64_t)1 << 30;   // raw rows per chunk (sgb_step2_set_chunk_bytes)
};

extern "C" int sgb_step2_set_model(sgb_ctx *h, int64_t N, int p, int binary, const double *mu, const double *res,
                                   const double *mu2, const double *y, const double *X, const double *XVX,
                                   const double *XXVX_inv, const double *XVX_inv_XV, const double *S_a, const double *tau,
                                   double varRatio, double SPAcutoff, const int32_t *pos_in_fam)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    if (p < 1 || p > S2_MAXP) return sgb_fail(h, "step2: p=%d out of range [1,%d]", p, S2_MAXP);
    if (N < 1) return sgb_fail(h, "step2: empty model");
    if (!h->step2) h->step2 = new sgb_step2();
    sgb_step2 *s = h->step2;
    if (s->d_vec) { cudaFree(s->d_vec); s->d_vec = nullptr; }
    if (s->d_pos) { cudaFree(s->d_pos); s->d_pos = nullptr; }
    if (s->d_P2) { cudaFree(s->d_P2); s->d_P2 = nullptr; }
    const size_t nvec = (size_t)N * (4 + 3 * p);
    CUDA_OK(h, cudaMalloc((void **)&s->d_vec, sizeof(double) * nvec));
    CUDA_OK(h, cudaMalloc((void **)&s->d_pos, sizeof(int32_t) * N));
    double *d = s->d_vec;
    const double *src[7] = {mu, mu2, res, y, X, XVX_inv_XV, XXVX_inv};
    const size_t len[7] = {(size_t)N, (size_t)N, (size_t)N, (size_t)N, (size_t)N * p, (size_t)N * p, (size_t)N * p};
    const double *dev[7];
    for (int i = 0; i < 7; i++) {
        CUDA_OK(h, cudaMemcpyAsync(d, src[i], sizeof(double) * len[i], cudaMemcpyHostToDevice, h->stream));
        dev[i] = d; d += len[i];
    }
    CUDA_OK(h, cudaMemcpyAsync(s->d_pos, pos_in_fam, sizeof(int32_t) * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    s2_model &M = s->M;
    M.N = N; M.p = p; M.binary = binary;
    M.mu = dev[0]; M.mu2 = dev[1]; M.res = dev[2]; M.y = dev[3]; M.X = dev[4]; M.A = dev[5]; M.XXVXi = dev[6];
    for (int i = 0; i < p * p; i++) M.XVX[i] = XVX[i];
    for (int i = 0; i < p; i++) M.S_a[i] = S_a[i];
    M.tau0 = tau[0]; M.varRatio = varRatio; M.spa_cutoff = SPAcutoff; M.pos = s->d_pos;
    // identity fast path: pos[i] == i; the case mask lives in the bit positions of the PLINK low bits
    M.identity = 1;
    for (int64_t i = 0; i < N; i++) if (pos_in_fam[i] != (int32_t)i) { M.identity = 0; break; }
    std::vector<uint32_t> yc((size_t)((N + 15) / 16) + 2, 0u);
    double ncase = 0;
    for (int64_t i = 0; i < N; i++) if (y[i] == 1.0) { yc[(size_t)(i >> 4)] |= 1u << (2 * (i & 15)); ncase += 1; }
    if (s->d_ycase) { cudaFree(s->d_ycase); s->d_ycase = nullptr; }
    CUDA_OK(h, cudaMalloc((void **)&s->d_ycase, sizeof(uint32_t) * yc.size()));
    CUDA_OK(h, cudaMemcpy(s->d_ycase, yc.data(), sizeof(uint32_t) * yc.size(), cudaMemcpyHostToDevice));
    M.ycase = s->d_ycase; M.ncase_tot = ncase;
    if (s->d_offset) { cudaFree(s->d_offset); s->d_offset = nullptr; }
    CUDA_OK(h, cudaMalloc((void **)&s->d_offset, sizeof(double) * N));
    CUDA_OK(h, cudaMemset(s->d_offset, 0, sizeof(double) * N));
    M.offset = s->d_offset; M.firth = 0; M.firth_se_from_fit = 1; M.firth_cutoff = 0.01;
    M.er_max_mac = -1.0; M.mu_sum = 0.0; M.n_cate = 1; M.n_cond = 0; M.P2 = nullptr;
    for (int64_t i = 0; i < N; i++) M.mu_sum += mu[i];

    // ---- batched score sums: model columns A | mu2 X | res | mu2, their sums, and their limb images (once per model) ----
    {
        const int kv = 2 * p + 2;
        s->kv = kv;
        s->stride = ((N + 3) / 4 + SGB_KSTEP_BYTES - 1) / SGB_KSTEP_BYTES * SGB_KSTEP_BYTES;
        void **olds[] = {(void **)&s->d_V, (void **)&s->d_csum, (void **)&s->d_Lv, (void **)&s->d_Li, (void **)&s->d_multv,
                         (void **)&s->d_multi, (void **)&s->d_lsv, (void **)&s->d_lsi};
        for (auto q : olds) { if (*q) cudaFree(*q); *q = nullptr; }
        std::vector<double> V((size_t)N * kv), cs(kv, 0.0);
        for (int j = 0; j < p; j++)
            for (int64_t i = 0; i < N; i++) {
                V[(size_t)j * N + i] = XVX_inv_XV[(size_t)j * N + i];
                V[(size_t)(p + j) * N + i] = mu2[i] * X[(size_t)j * N + i];
            }
        for (int64_t i = 0; i < N; i++) { V[(size_t)(2 * p) * N + i] = res[i]; V[(size_t)(2 * p + 1) * N + i] = mu2[i]; }
        for (int c = 0; c < kv; c++) { double t = 0.0; for (int64_t i = 0; i < N; i++) t += V[(size_t)c * N + i]; cs[c] = t; }
        const size_t img = k_umma_image_bytes(k_umma_npad(kv, 7), s->stride), frag = (size_t)(s->stride / SGB_KSTEP_BYTES) * 2048;
        CUDA_OK(h, cudaMalloc((void **)&s->d_V, sizeof(double) * V.size()));
        CUDA_OK(h, cudaMalloc((void **)&s->d_csum, sizeof(double) * kv));
        CUDA_OK(h, cudaMalloc((void **)&s->d_Lv, img));
        CUDA_OK(h, cudaMalloc((void **)&s->d_Li, frag));
        CUDA_OK(h, cudaMalloc((void **)&s->d_multv, sizeof(double) * kv));
        CUDA_OK(h, cudaMalloc((void **)&s->d_multi, sizeof(double)));
        CUDA_OK(h, cudaMalloc((void **)&s->d_lsv, sizeof(int32_t) * 8 * kv));
        CUDA_OK(h, cudaMalloc((void **)&s->d_lsi, sizeof(int32_t) * 8));
        CUDA_OK(h, cudaMemcpyAsync(s->d_V, V.data(), sizeof(double) * V.size(), cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(h, cudaMemcpyAsync(s->d_csum, cs.data(), sizeof(double) * kv, cudaMemcpyHostToDevice, h->stream));
        SGB_TRY(k_split_limbs_umma(h, s->d_V, N, N, kv, s->d_Lv, s->stride, s->d_multv, s->d_lsv, 7, 0));
        SGB_TRY(k_split_limbs(h, s->d_V + (size_t)(2 * p + 1) * N, N, N, 1, s->d_Li, s->stride / SGB_KSTEP_BYTES, s->d_multi, s->d_lsi, 0));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

extern "C" int sgb_step2_set_chunk_bytes(sgb_ctx *h, int64_t bytes)
{
    if (bytes < 1) return sgb_fail(h, "step2: chunk size must be positive");
    if (!h->step2) h->step2 = new sgb_step2();
    h->step2->chunk_bytes = bytes;
    return 0;
}

extern "C" int sgb_step2_set_batched(sgb_ctx *h, int enable)
{
    if (!h->step2) h->step2 = new sgb_step2();
    h->step2->batched = enable != 0;
    return 0;
}

extern "C" int sgb_step2_set_variance_ratios(sgb_ctx *h, int n_cate, const double *ratios, const double *min_mac_exclude,
                                             const double *max_mac_include)
{
    sgb_step2 *s = h->step2;
    if (!s || !s->d_vec) return sgb_fail(h, "step2: call sgb_step2_set_model first");
    if (n_cate < 1 || n_cate > S2_MAXCATE) return sgb_fail(h, "step2: %d variance-ratio categories, supported 1..%d", n_cate, S2_MAXCATE);
    s2_model &M = s->M;
    if (n_cate == 1) { M.n_cate = 1; M.varRatio = ratios[0]; return 0; }
    for (int c = 0; c < n_cate; c++) {
        if (c + 1 < n_cate && !(max_mac_include[c] > min_mac_exclude[c]))
            return sgb_fail(h, "step2: variance-ratio category %d is empty (%g, %g]", c + 1, min_mac_exclude[c], max_mac_include[c]);
        if (c > 0 && min_mac_exclude[c] != max_mac_include[c - 1])
            return sgb_fail(h, "step2: variance-ratio categories must tile the MAC axis (category %d starts at %g, the one before ends at %g)",
                            c + 1, min_mac_exclude[c], max_mac_include[c - 1]);
        M.cate_ratio[c] = ratios[c]; M.cate_min[c] = min_mac_exclude[c];
        M.cate_max[c] = c + 1 < n_cate ? max_mac_include[c] : INFINITY;
    }
    M.n_cate = n_cate; M.varRatio = ratios[0];
    return 0;
}

extern "C" int sgb_step2_set_condition(sgb_ctx *h, int n_cond, const double *P2, const double *XtP2, const double *VarInv,
                                       const double *Tstat_cond)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    sgb_step2 *s = h->step2;
    if (!s || !s->d_vec) return sgb_fail(h, "step2: call sgb_step2_set_model first");
    if (n_cond < 0 || n_cond > S2_MAXCOND) return sgb_fail(h, "step2: %d conditioning markers, supported 0..%d", n_cond, S2_MAXCOND);
    s2_model &M = s->M;
    if (s->d_P2) { cudaFree(s->d_P2); s->d_P2 = nullptr; }
    M.n_cond = 0; M.P2 = nullptr;
    if (n_cond == 0) return 0;
    CUDA_OK(h, cudaMalloc((void **)&s->d_P2, sizeof(double) * (size_t)M.N * n_cond));
    CUDA_OK(h, cudaMemcpy(s->d_P2, P2, sizeof(double) * (size_t)M.N * n_cond, cudaMemcpyHostToDevice));
    for (int i = 0; i < M.p * n_cond; i++) M.XtP2[i] = XtP2[i];
    for (int i = 0; i < n_cond * n_cond; i++) M.VarInv[i] = VarInv[i];
    for (int c = 0; c < n_cond; c++) {
        double v = 0.0;
        for (int d = 0; d < n_cond; d++) v += VarInv[c + d * n_cond] * Tstat_cond[d];
        M.VT[c] = v;
    }
    M.n_cond = n_cond; M.P2 = s->d_P2;
    return 0;
}

extern "C" int sgb_step2_set_er(sgb_ctx *h, double max_mac_for_er)
{
    sgb_step2 *s = h->step2;
    if (!s || !s->d_vec) return sgb_fail(h, "step2: call sgb_step2_set_model first");
    if (max_mac_for_er > (double)SGB_ER_MAXK)
        return sgb_fail(h, "step2: max_MAC_for_ER=%g is above %d (the exact test enumerates 2^carriers assignments)", max_mac_for_er, SGB_ER_MAXK);
    s->M.er_max_mac = max_mac_for_er < 0 ? -1.0 : max_mac_for_er;
    return 0;
}

extern "C" int sgb_step2_set_firth(sgb_ctx *h, int enable, double p_cutoff, const double *offset, int se_from_fit)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    sgb_step2 *s = h->step2;
    if (!s || !s->d_vec) return sgb_fail(h, "step2: call sgb_step2_set_model first");
    if (enable && !(p_cutoff >= 0.0 && p_cutoff <= 1.0)) return sgb_fail(h, "step2: pCutoffforFirth=%g out of [0,1]", p_cutoff);
    if (offset) CUDA_OK(h, cudaMemcpy(s->d_offset, offset, sizeof(double) * s->M.N, cudaMemcpyHostToDevice));
    else CUDA_OK(h, cudaMemset(s->d_offset, 0, sizeof(double) * s->M.N));
    s->M.firth = enable ? 1 : 0; s->M.firth_cutoff = p_cutoff; s->M.firth_se_from_fit = se_from_fit ? 1 : 0;
    return 0;
}

// pageable -> pinned staging copy on 4 host threads (one thread tops out near 10 GB/s, below what the kernel consumes)
static void s2_par_copy(uint8_t *dst, const uint8_t *src, size_t n)
{
    const int nt = n > ((size_t)8 << 20) ? 8 : 1;
    if (nt == 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const size_t o = t * per, len = o < n ? std::min(per, n - o) : 0;
        if (len) th.emplace_back([=] { memcpy(dst + o, src + o, len); });
    }
    for (auto &x : th) x.join();
}

extern "C" int sgb_step2_test_markers(sgb_ctx *h, const uint8_t *bed_rows, int64_t n_fam, int64_t n_markers, double min_maf,
                                      double min_mac, double max_missing, int se_two_sided, double *out)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    sgb_step2 *s = h->step2;
    if (!s || !s->d_vec) return sgb_fail(h, "step2: call sgb_step2_set_model first");
    if (n_markers <= 0) return 0;
    const int64_t B0 = (n_fam + 3) / 4;
    if (B0 > 200 * 1024) return sgb_fail(h, "step2: more than 819,200 samples in the .fam are not supported yet");
    // chunks of <= 1 GB of raw rows, double-buffered: the H2D of chunk c+1 (on its own copy stream, one event per buffer) overlaps
    // the kernels of chunk c on h->stream.  The chunk is this large for the flagged variants: a saddle-point variant keeps one CTA busy for milliseconds
    // (~20 passes over all samples), so the per-variant kernel only fills the machine (444 resident CTAs) when a chunk holds
    // >= ~10^4 variants of which ~5 % are flagged (measured: 43 us per flagged variant with 256 MB chunks at N = 200k)
    // at most 65,024 variants per chunk: the re-pack / gather kernels index the variant with blockIdx.y (<= 65,535)
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_markers, s->chunk_bytes / B0), 65024));
    const size_t cbytes = (size_t)chunk * B0, obytes = sizeof(double) * (size_t)chunk * S2_NOUT;
    if (!s->pin[0] || s->pin_bytes < cbytes) {
        for (int i = 0; i < 2; i++) { if (s->pin[i]) cudaFreeHost(s->pin[i]); s->pin[i] = nullptr; }
        s->pin_bytes = 0;
        for (int i = 0; i < 2; i++) CUDA_OK(h, cudaMallocHost((void **)&s->pin[i], cbytes));
        s->pin_bytes = cbytes;
    }
    if (!s->pout[0] || s->pout_bytes < obytes) {      // a later call with fewer samples per row has MORE markers per chunk
        for (int i = 0; i < 2; i++) { if (s->pout[i]) cudaFreeHost(s->pout[i]); s->pout[i] = nullptr; }
        s->pout_bytes = 0;
        for (int i = 0; i < 2; i++) CUDA_OK(h, cudaMallocHost((void **)&s->pout[i], obytes));
        s->pout_bytes = obytes;
    }
    for (int i = 0; i < 2; i++) if (!s->ev[i]) CUDA_OK(h, cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) if (!s->ev_up[i]) CUDA_OK(h, cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming));
    if (!s->copy_stream) CUDA_OK(h, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    SGB_TRY(sgb_ensure(h, (void **)&s->d_bed, &s->bed_bytes, 2 * cbytes));
    size_t ob = s->out_elems * sizeof(double);
    SGB_TRY(sgb_ensure(h, (void **)&s->d_out, &ob, 2 * obytes));
    s->out_elems = ob / sizeof(double);
    if (sgb_first_on_device(h->device, SGB_SITE_STEP2)) {
        CUDA_OK(h, cudaFuncSetAttribute(step2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
        CUDA_OK(h, cudaFuncSetAttribute(step2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
        CUDA_OK(h, cudaFuncSetAttribute(step2_flagged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
    }
    // batched path: chunk = a multiple of the row alignment of the tiled store; scratch for the tensor-engine sums
    const bool batched = s->batched;
    const int64_t N = s->M.N, Bm = (N + 3) / 4;
    const int64_t rows_pad = (chunk + SGB_ROW_ALIGN - 1) / SGB_ROW_ALIGN * SGB_ROW_ALIGN;
    const int kv = s->kv, npadv = k_umma_npad(kv, 7);
    const int flag_ctas = h->sm_count * 3;                               // persistent CTAs of the flagged-variant kernel (3 per SM)
    const int64_t slot_doubles = 2 * ((N + 1) & ~(int64_t)1) + 2;        // gtilde of N samples + N/2 (gtilde, mu) pairs
    if (batched) {
        if (!s->M.identity) SGB_TRY(sgb_ensure(h, (void **)&s->d_gath, &s->gath_bytes, (size_t)chunk * Bm));
        SGB_TRY(sgb_ensure(h, (void **)&s->d_tiled, &s->tiled_bytes, (size_t)rows_pad * s->stride));
        size_t b;
        b = s->accv_elems * 4; if (b < (size_t)rows_pad * npadv * 4) { SGB_TRY(sgb_ensure(h, (void **)&s->d_accv, &b, (size_t)rows_pad * npadv * 4)); s->accv_elems = b / 4; CUDA_OK(h, cudaMemsetAsync(s->d_accv, 0, b, h->stream)); }
        b = s->acci_elems * 4; if (b < (size_t)rows_pad * 8 * 4) { SGB_TRY(sgb_ensure(h, (void **)&s->d_acci, &b, (size_t)rows_pad * 8 * 4)); s->acci_elems = b / 4; CUDA_OK(h, cudaMemsetAsync(s->d_acci, 0, b, h->stream)); }
        SGB_TRY(sgb_ensure_f64(h, &s->d_raw, &s->raw_elems, (size_t)rows_pad * (kv + 1)));
        b = s->cnt_elems * 4; SGB_TRY(sgb_ensure(h, (void **)&s->d_cnt, &b, (size_t)chunk * 6 * 4)); s->cnt_elems = b / 4;
        b = s->list_elems * 4; SGB_TRY(sgb_ensure(h, (void **)&s->d_list, &b, (size_t)(chunk + 2) * 4)); s->list_elems = b / 4;
        SGB_TRY(sgb_ensure(h, (void **)&s->d_spa, &s->spa_bytes, (size_t)flag_ctas * slot_doubles * sizeof(double)));
    }
    // rows in page-locked memory (cudaHostAlloc / cudaHostRegister by the caller) go to the device straight from the caller's
    // buffer; pageable rows pass through the pinned double buffer (a host memcpy at ~13 GB/s: the e2e limit of that case)
    cudaPointerAttributes pattr;
    const bool rows_pinned = cudaPointerGetAttributes(&pattr, bed_rows) == cudaSuccess && pattr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    int cur = 0;
    int64_t held_m0[2] = {-1, -1}, held_nm[2] = {0, 0};
    for (int64_t m0 = 0; m0 < n_markers; m0 += chunk, cur ^= 1) {
        const int64_t nm = std::min(chunk, n_markers - m0);
        CUDA_OK(h, cudaEventSynchronize(s->ev[cur]));                      // buffers `cur` are free again (chunk c-2 done)
        if (held_m0[cur] >= 0) memcpy(out + (size_t)held_m0[cur] * S2_NOUT, s->pout[cur], sizeof(double) * held_nm[cur] * S2_NOUT);
        uint8_t *db = s->d_bed + cur * cbytes;
        double *dout = s->d_out + cur * (size_t)chunk * S2_NOUT;
        // db was last read by the kernels of chunk c-2 (finished: ev[cur] above), so the copy engine may refill it while
        // chunk c-1's kernels still run on h->stream
        // pageable rows are staged as ONE piece per chunk: 64 MB pieces with a copy per piece were measured slower
        // (455k against 540k variants/s), the staging threads being cheaper to start once per GB
        if (!rows_pinned) s2_par_copy(s->pin[cur], bed_rows + (size_t)m0 * B0, (size_t)nm * B0);
        CUDA_OK(h, cudaMemcpyAsync(db, rows_pinned ? bed_rows + (size_t)m0 * B0 : s->pin[cur], (size_t)nm * B0, cudaMemcpyHostToDevice, s->copy_stream));
        CUDA_OK(h, cudaEventRecord(s->ev_up[cur], s->copy_stream));
        CUDA_OK(h, cudaStreamWaitEvent(h->stream, s->ev_up[cur], 0));
        h->cnt.bytes_h2d += nm * B0;
        if (batched) {
            SGB_RANGE("step2_chunk_batched");
            const uint8_t *rows = db;
            int64_t Brow = B0;
            if (!s->M.identity) {          // model samples gathered into identity order once; every later pass is the identity path
                s2_gather_kernel<<<dim3((unsigned)((Bm + 255) / 256), (unsigned)nm), 256, 0, h->stream>>>(db, B0, s->M.pos, N, Bm, s->d_gath);
                h->cnt.n_kernel_launches++;
                rows = s->d_gath; Brow = Bm;
            }
            const int64_t rp = (nm + SGB_ROW_ALIGN - 1) / SGB_ROW_ALIGN * SGB_ROW_ALIGN;
            s2_count_kernel<<<(unsigned)nm, 256, 0, h->stream>>>(rows, Brow, N, s->M.ycase, s->d_cnt);
            s2_repack_kernel<<<dim3((unsigned)((s->stride / 4 + 255) / 256), (unsigned)rp), 256, 0, h->stream>>>(rows, Brow, N, nm, s->d_cnt, min_maf, min_mac,
                                                                                                              max_missing, s->d_tiled, s->stride);
            h->cnt.n_kernel_launches += 2;
            CUDA_OK(h, cudaGetLastError());
            double *rawV = s->d_raw, *rawI = s->d_raw + (size_t)rows_pad * kv;
            SGB_TRY(k_pk2_umma(h, s->d_tiled, s->stride, rp, s->stride, s->d_Lv, kv, s->d_accv, SGB_PLANE_VALUE, 7));
            SGB_TRY(k_recombine_umma(h, s->d_accv, rp, kv, s->d_multv, s->d_lsv, SGB_PLANE_VALUE, rawV, rows_pad));
            SGB_TRY(k_pk2_gemm(h, s->d_tiled, s->stride, rp, s->stride, s->d_Li, 1, s->d_acci, SGB_PLANE_IS2));
            SGB_TRY(k_recombine(h, s->d_acci, rp, 1, 1, s->d_multi, s->d_lsi, SGB_PLANE_IS2, rawI, rows_pad));
            CUDA_OK(h, cudaMemsetAsync(s->d_list, 0, 2 * sizeof(int), h->stream));
            s2_model Mi = s->M; Mi.identity = 1;
            s2_finish_kernel<<<(unsigned)((nm + 127) / 128), 128, 0, h->stream>>>(Mi, s->d_cnt, rawV, rawI, rows_pad, s->d_csum, nm, min_maf, min_mac,
                                                                                max_missing, dout, s->d_list + 2, s->d_list);
            // the flagged variants (saddle point, exact test, Firth, conditional): persistent CTAs pull them off the list
            step2_flagged_kernel<<<(unsigned)std::min<int64_t>(flag_ctas, nm), S2_THREADS, (size_t)Brow + 8, h->stream>>>(
                Mi, rows, Brow, nm, min_maf, min_mac, max_missing, se_two_sided, dout, s->d_list + 2, s->d_list, s->d_list + 1, s->d_spa, slot_doubles);
            h->cnt.n_kernel_launches += 2;
        } else if (s->M.identity)
            step2_kernel<true, false><<<(unsigned)nm, S2_THREADS, (size_t)B0 + 8, h->stream>>>(s->M, db, B0, nm, min_maf, min_mac, max_missing, se_two_sided, dout, s2_dose{});
        else
            step2_kernel<false, false><<<(unsigned)nm, S2_THREADS, (size_t)B0 + 8, h->stream>>>(s->M, db, B0, nm, min_maf, min_mac, max_missing, se_two_sided, dout, s2_dose{});
        h->cnt.n_kernel_launches++;
        CUDA_OK(h, cudaGetLastError());
        CUDA_OK(h, cudaMemcpyAsync(s->pout[cur], dout, sizeof(double) * nm * S2_NOUT, cudaMemcpyDeviceToHost, h->stream));
        h->cnt.bytes_d2h += sizeof(double) * nm * S2_NOUT;
        CUDA_OK(h, cudaEventRecord(s->ev[cur], h->stream));
        held_m0[cur] = m0; held_nm[cur] = nm;
    }
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 2; i++)
        if (held_m0[i] >= 0) memcpy(out + (size_t)held_m0[i] * S2_NOUT, s->pout[i], sizeof(double) * held_nm[i] * S2_NOUT);
    return 0;
}

// Dosage rows (VCF DS / GT, BGEN): same marker loop, genotypes read as doubles.  Chunks of <= 256 MB of rows go to the
// device from the caller's (pageable) buffer; the transfer dominates (8 bytes per sample against 1/4 byte for hard calls),
// so this first version keeps one buffer and no overlap.
extern "C" int sgb_step2_test_dosages(sgb_ctx *h, const double *dosages, int64_t n_file_samples, int64_t n_markers,
                                      double min_maf, double min_mac, double max_missing, int se_two_sided, int impute_method,
                                      double dosage_zerod_cutoff, double dosage_zerod_mac_cutoff, double *out)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    sgb_step2 *s = h->step2;
    if (!s || !s->d_vec) return sgb_fail(h, "step2: call sgb_step2_set_model first");
    if (impute_method < 1 || impute_method > 3) return sgb_fail(h, "step2: impute_method %d (1 best_guess, 2 mean, 3 minor)", impute_method);
    if (n_file_samples < 1) return sgb_fail(h, "step2: empty dosage rows");
    if (n_markers <= 0) return 0;
    const size_t row_bytes = sizeof(double) * (size_t)n_file_samples;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_markers, (int64_t)(((size_t)256 << 20) / row_bytes)));
    SGB_TRY(sgb_ensure(h, (void **)&s->d_bed, &s->bed_bytes, (size_t)chunk * row_bytes));
    size_t ob = s->out_elems * sizeof(double);
    SGB_TRY(sgb_ensure(h, (void **)&s->d_out, &ob, sizeof(double) * (size_t)chunk * S2_NOUT));
    s->out_elems = ob / sizeof(double);
    for (int64_t m0 = 0; m0 < n_markers; m0 += chunk) {
        const int64_t nm = std::min(chunk, n_markers - m0);
        CUDA_OK(h, cudaMemcpyAsync(s->d_bed, dosages + (size_t)m0 * n_file_samples, (size_t)nm * row_bytes, cudaMemcpyHostToDevice, h->stream));
        h->cnt.bytes_h2d += nm * row_bytes;
        s2_dose ds;
        ds.d = reinterpret_cast<const double *>(s->d_bed); ds.stride = n_file_samples; ds.impute = impute_method;
        ds.zerod_cutoff = dosage_zerod_cutoff; ds.zerod_mac_cutoff = dosage_zerod_mac_cutoff;
        step2_kernel<false, true><<<(unsigned)nm, S2_THREADS, 0, h->stream>>>(s->M, nullptr, 0, nm, min_maf, min_mac, max_missing, se_two_sided, s->d_out, ds);
        h->cnt.n_kernel_launches++;
        CUDA_OK(h, cudaGetLastError());
        CUDA_OK(h, cudaMemcpyAsync(out + (size_t)m0 * S2_NOUT, s->d_out, sizeof(double) * nm * S2_NOUT, cudaMemcpyDeviceToHost, h->stream));
        h->cnt.bytes_d2h += sizeof(double) * nm * S2_NOUT;
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

void sgb_step2_free(sgb_ctx *h)
{
    sgb_step2 *s = h->step2;
    if (!s) return;
    if (s->d_vec) cudaFree(s->d_vec);
    if (s->d_pos) cudaFree(s->d_pos);
    if (s->d_ycase) cudaFree(s->d_ycase);
    if (s->d_offset) cudaFree(s->d_offset);
    if (s->d_P2) cudaFree(s->d_P2);
    if (s->d_bed) cudaFree(s->d_bed);
    if (s->d_out) cudaFree(s->d_out);
    void *more[] = {s->d_V, s->d_csum, s->d_Lv, s->d_Li, s->d_multv, s->d_multi, s->d_lsv, s->d_lsi, s->d_gath, s->d_tiled, s->d_accv, s->d_acci,
                    s->d_raw, s->d_cnt, s->d_list, s->d_spa};
    for (auto q : more) if (q) cudaFree(q);
    for (int i = 0; i < 2; i++) { if (s->pin[i]) cudaFreeHost(s->pin[i]); if (s->pout[i]) cudaFreeHost(s->pout[i]); if (s->ev[i]) cudaEventDestroy(s->ev[i]); if (s->ev_up[i]) cudaEventDestroy(s->ev_up[i]); }
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
    h->step2 = nullptr;
}
