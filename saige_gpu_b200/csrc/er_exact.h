// Exact test of rare variants in step 2 ("efficient resampling", variants with MAC <= max_MAC_for_ER on a binary trait).
//
// Replaces SKATExactBin_Work and its helpers (src/SAIGE/src/ER_binary_func.cpp:23-85, 113-143, 186-278) together with
// the two classes they drive: HyperGeo::Run / Get_lprob (Binary_HyperGeo.cpp:37-150, lCombinations :172-190) and
// ComputeExact::Init / Run / GetPvalues (Binary_ComputeExact.cpp:296-470), for one variant (m = 1) in the all-exact
// regime (2^k assignments <= NResampling = 2e6; k = number of carriers <= SGB_ER_MAXK).
//
// Same arithmetic, different organisation: no recursion and no tables of 2^k entries.
//   * P(j carriers are cases): the reference walks every allocation of cases to the carrier classes recursively; the
//     same sum is the j-th coefficient of prod_classes (1 + w x)^size, built by polynomial multiplication, times
//     C(n - k, ncase - j).
//   * the per-size normaliser sum_{|S| = j} prod_{i in S} odds_i is the elementary symmetric polynomial e_j(odds).
//   * one pass over the 2^k bit masks accumulates P(stat >= observed) and P(stat == observed) (ties within epsilon).
// Written once for host and device: step2.cu calls it from one thread of the variant's CTA; tests/test_step2_rare_exact.py
// compiles this header with g++ and checks it against the outputs of the reference's own compiled code
// (tests/golden/er_golden.json).  Nothing here is reachable without the CUDA kernel in the product.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SGB_ER_FN __host__ __device__ __noinline__
#else
#define SGB_ER_FN static
#endif

#define SGB_ER_MAXK 10          // the reference sizes its probability classes for "MAC <= 10" (ER_binary_func.cpp:26)

// HyperGeo::lCombinations: log C(n, r) as R's lchoose, except r > n -> 0 (the reference's own convention) and r < 0 -> -inf
SGB_ER_FN double sgb_er_lchoose(double n, double r)
{
    if (r > n) return 0.0;
    if (r < 0) return -INFINITY;
    return lgamma(n + 1.0) - lgamma(r + 1.0) - lgamma(n - r + 1.0);
}

// g1 / p1 / res1: genotype, fitted probability and residual (y - mu) of the k carriers; p2mean: mean fitted probability of
// the n - k non-carriers; ncase: cases among all n samples.  Returns pval - pval_same / 2 (ER_binary_func.cpp:274).
SGB_ER_FN double sgb_er_exact_pvalue(int k, const double *g1, const double *p1, const double *res1, double p2mean, double n,
                                     double ncase, double epsilon)
{
    double poly[SGB_ER_MAXK + 1], tmp[SGB_ER_MAXK + 1], prob[SGB_ER_MAXK + 1], esym[SGB_ER_MAXK + 1], odds[SGB_ER_MAXK];
    if (k < 0 || k > SGB_ER_MAXK) return NAN;
    // ---- SKATExactBin_ComputeProb_Group: ten classes of fitted probability, class-mean odds relative to the non-carriers' ----
    const double p2odd = p2mean / (1.0 - p2mean);
    for (int j = 0; j <= k; j++) poly[j] = j == 0 ? 1.0 : 0.0;
    for (int b = 0; b < 10; b++) {
        const double a1 = (double)b / 10, a2 = (double)(b + 1) / 10;
        int c = 0;
        double s = 0.0;
        for (int i = 0; i < k; i++) {
            const double p = p1[i] >= 1.0 ? 0.999 : p1[i];
            if (p >= a1 && (b + 1 < 10 ? p < a2 : p <= a2)) { c++; s += p; }
        }
        if (!c) continue;
        const double pm = s / c, w = pm / (1.0 - pm) / p2odd;
        // poly *= (1 + w x)^c, i.e. sum_i C(c, i) w^i x^i  (HyperGeo's table lCombinations(c, i) + i log w)
        for (int j = 0; j <= k; j++) tmp[j] = 0.0;
        double term = 1.0;                       // C(c, i) w^i
        for (int i = 0; i <= c; i++) {
            for (int j = i; j <= k; j++) tmp[j] += poly[j - i] * term;
            term = term * w * (double)(c - i) / (double)(i + 1);
        }
        for (int j = 0; j <= k; j++) poly[j] = tmp[j];
    }
    // ---- last class (n - k non-carriers, weight 1): C(n - k, ncase - j), scaled by the largest (m_ref, never below 0) ----
    double ref = 0.0;
    for (int j = 0; j <= k; j++) {
        tmp[j] = sgb_er_lchoose(n - k, ncase - j);
        if (tmp[j] > ref) ref = tmp[j];
    }
    double tot = 0.0;
    for (int j = 0; j <= k; j++) {
        prob[j] = (double)j <= ncase ? poly[j] * exp(tmp[j] - ref) : 0.0;
        tot += prob[j];
    }
    for (int j = 0; j <= k; j++) prob[j] /= tot;
    // ---- SKATExactBin_Work + ComputeExact: every case / control assignment of the carriers ----
    double z0sum = 0.0, gobs = 0.0;
    for (int j = 0; j <= k; j++) esym[j] = j == 0 ? 1.0 : 0.0;
    for (int i = 0; i < k; i++) {
        odds[i] = p1[i] / (1.0 - p1[i]);
        z0sum += g1[i] * (-p1[i]);
        if (res1[i] > 0) gobs += g1[i];
        for (int j = i + 1; j >= 1; j--) esym[j] += esym[j - 1] * odds[i];
    }
    const double Q = (z0sum + gobs) * (z0sum + gobs);
    double all = 0.0, pv = 0.0, same = 0.0;
    for (unsigned mask = 0; mask < (1u << k); mask++) {
        double s = 0.0, w = 1.0;
        int j = 0;
        for (int i = 0; i < k; i++)
            if (mask >> i & 1u) { s += g1[i]; w *= odds[i]; j++; }
        const double stat = (z0sum + s) * (z0sum + s);
        const double fp = w / esym[j] * prob[j];
        all += fp;
        double d = Q - stat;
        if (fabs(d) <= epsilon) d = 0.0;
        if (d <= 0) { pv += fp; if (d == 0) same += fp; }
    }
    return (pv - same / 2) / all;
}
