"""CPU oracle for the NEXT row of SURVEY.md 8(f): step-2 single-variant score test with saddle-point approximation
(binary and quantitative traits, PLINK input, full-GRM variance ratio).  TEST INFRASTRUCTURE ONLY.

Restates, in numpy fp64:
  PlinkClass::getOneMarker          /root/reference/src/SAIGE/src/PLINK.cpp:164-300   (alt-first: A1 of the .bim is ALT)
  mainMarkerInCPP (marker loop)     src/Main.cpp:149-560
  imputeGenoAndFlip                 src/UTIL.cpp:58-135
  SAIGEClass::scoreTestFast         src/SAIGE_test.cpp:212-292
  SAIGEClass::getMarkerPval         src/SAIGE_test.cpp:345-640   (SPA / SPA_fast dispatch, SE from the SPA p-value)
  SPA, SPA_fast                     src/SPA.cpp:20-185
  Korg/K1/K2/getroot/Get_Saddle_Prob (+ _fast) for the binomial CGF   src/SPA_binary.cpp:21-330
  ReadModel                         src/SAIGE/R/readInGLMM.R:39-170 (which fields of the .rda step 2 consumes, LOCO swap)

Pinned by the reference's own golden table extdata/output/genotype_100markers_marker_plink.txt (32 variants, produced
from extdata/output/example_binary.rda + extdata/input/genotype_100markers.{bed,bim,fam}) -- tests/test_step2_golden.py.
Firth's bias-reduced effect size (is_Firth_beta, SAIGE_test.cpp:573-633, fast_logistf_fit_simple :893-986) is pinned by the
BETA / SE of extdata/output/example_binary_positive_signal.assoc.step2.txt.
Efficient resampling ("ER", the exact test of rare variants, MAC <= max_MAC_for_ER = 4; SAIGE_test.cpp:426-431, 592-620,
ER_binary_func.cpp:23-278, Binary_HyperGeo.cpp:37-193, Binary_ComputeExact.cpp:26-470) is restated in `er_pvalue` and pinned on
the reference's OWN compiled code: oracle/Makefile builds those Binary_*.cpp files (they carry a _STAND_ALONE_ switch) into
oracle/_ref/libskat_exact_ref.so, and tests/golden/er_golden.json holds its outputs (tests/golden/make_er_golden.py).
Not implemented: conditional analysis, sparse-GRM variance, categorical variance ratios.
"""
import math

import numpy as np
from scipy import special, stats

EPS25 = np.finfo(np.float64).eps ** 0.25          # tol = eps^0.25 (SAIGE_test.cpp:515-516)


def read_model(modglmm, chrom=None, LOCO=True):
    """ReadModel (readInGLMM.R:39-170): returns the dict of arrays step 2 uses."""
    m = dict(modglmm)
    mu = None if m.get("fitted.values") is None else np.asarray(m["fitted.values"], dtype=np.float64).ravel()
    res = None if m.get("residuals") is None else np.asarray(m["residuals"], dtype=np.float64).ravel()
    noK = m.get("obj.noK")
    if LOCO and chrom is not None and bool(np.asarray(m["LOCO"]).ravel()[0]):
        lr = m["LOCOResult"][int(chrom) - 1]
        if lr is not None and isinstance(lr, dict) and "fitted.values" in lr:
            mu = np.asarray(lr["fitted.values"], dtype=np.float64).ravel()
            res = np.asarray(lr["residuals"], dtype=np.float64).ravel()
            noK = lr["obj.noK"]
    trait = m["traitType"][0] if isinstance(m["traitType"], list) else str(m["traitType"])
    tau = np.asarray(m["theta"], dtype=np.float64).ravel()
    mu2 = mu * (1 - mu) if trait == "binary" else np.full(len(mu), 1.0 / tau[0])
    # offset of the Firth refit (readInGLMM.R:99-101,134-160): the chromosome's own when LOCO stored one, else the model's, else 0
    offset = m.get("offset")
    if LOCO and chrom is not None and bool(np.asarray(m["LOCO"]).ravel()[0]):
        lr = m["LOCOResult"][int(chrom) - 1]
        if isinstance(lr, dict) and lr.get("offset") is not None:
            offset = lr["offset"]
    offset = np.zeros(len(mu)) if offset is None else np.asarray(offset, dtype=np.float64).ravel()
    return dict(mu=mu, res=res, mu2=mu2, tau=tau, trait=trait, offset=offset, y=np.asarray(m["y"], dtype=np.float64).ravel(),
                X=np.asarray(m["X"], dtype=np.float64), XV=np.asarray(noK["XV"]), XVX=np.asarray(noK["XVX"]),
                XXVX_inv=np.asarray(noK["XXVX_inv"]), XVX_inv_XV=np.asarray(noK["XVX_inv_XV"]),
                S_a=np.asarray(noK["S_a"]).ravel(), sampleID=list(m["sampleID"]))


def plink_marker(bed_body, n_fam, marker, pos_in_plink):
    """getOneMarker, alt-first: genotype = copies of A1; -1 = missing (PLINK.hpp:48-56)."""
    B0 = (n_fam + 3) // 4
    row = bed_body[marker * B0:(marker + 1) * B0]
    codes = ((row[:, None] >> np.array([0, 2, 4, 6])) & 3).reshape(-1)[:n_fam]
    g = np.array([2, -1, 1, 0])[codes]             # 0b00 HOM_ALT=2, 0b01 missing, 0b10 HET=1, 0b11 HOM_REF=0
    return g[pos_in_plink].astype(np.float64)


# ---- binomial cumulant generating function and its saddle point (SPA_binary.cpp) ----
def _K0(t, mu, g):
    return np.sum(np.log(1 - mu + mu * np.exp(g * t)))


def _K1(t, mu, g, q):
    return np.sum(mu * g / ((1 - mu) * np.exp(-g * t) + mu)) - q


def _K2(t, mu, g):
    e = np.exp(-g * t)
    return np.sum((1 - mu) * mu * g * g * e / ((1 - mu) * e + mu) ** 2)


def _getroot(K1f, K2f, q, gpos, gneg, tol, maxiter=1000, fast=False):
    """getroot_K1_Binom / getroot_K1_fast_Binom (SPA_binary.cpp:70-140, 214-270): safeguarded Newton."""
    if q >= gpos or q <= gneg:
        return np.inf, 0, True
    t, K1e, prev = 0.0, K1f(0.0), np.inf
    rep, conv = 1, True
    while rep <= maxiter:
        K2e = K2f(t)
        tnew = t - K1e / K2e
        if np.isnan(tnew):
            conv = False
            break
        if abs(tnew - t) < tol:
            conv = True
            break
        if rep == maxiter:
            conv = False
            break
        newK1 = K1f(tnew)
        changed = (K1e * newK1 < 0) if fast else (np.sign(K1e) != np.sign(newK1))
        if changed:
            if abs(tnew - t) > prev - tol:
                tnew = t + np.sign(newK1 - K1e) * prev / 2
                newK1 = K1f(tnew)
                prev = prev / 2
            else:
                prev = abs(tnew - t)
        rep += 1
        t, K1e = tnew, newK1
    return t, rep, conv


def _saddle_prob(zeta, K0f, K2f, q):
    """Get_Saddle_Prob[_fast]_Binom (SPA_binary.cpp:146-214, 276-330), logp = False: Lugannani-Rice."""
    k1, k2 = K0f(zeta), K2f(zeta)
    temp1 = zeta * q - k1
    if np.isfinite(k1) and np.isfinite(k2) and temp1 >= 0 and k2 >= 0:
        w = np.sign(zeta) * np.sqrt(2 * temp1)
        v = zeta * np.sqrt(k2)
        if w != 0:
            Z = w + np.log(v / w) / w
            _saddle_prob.last_log = float(special.log_ndtr(-abs(Z)))       # log |p| (R::pnorm(..., logp), SPA_binary.cpp:190-200)
            return (stats.norm.sf(Z) if Z > 0 else -stats.norm.cdf(Z)), True
    _saddle_prob.last_log = -np.inf
    return 0.0, False


def log_erfc(x):
    """log erfc(x), finite where erfc underflows: log of the chi-square(1) upper tail at stat = 2 x^2 (R::pchisq(..., log = TRUE),
    SAIGE_test.cpp:274)."""
    return float(np.log(special.erfc(x))) if x < 20 else float(np.log(special.erfcx(x)) - x * x)


def qnorm_from_logp(logp):
    """|qnorm(p, lower = FALSE, log.p = TRUE)| (SAIGE_test.cpp:531)."""
    return abs(float(special.ndtri_exp(logp)))


def spa_pvalue(mu, gt, q, qinv, pval_noadj, idx_nz, fast, var2=None, logp_noadj=None):
    """SPA / SPA_fast (SPA.cpp:20-185)."""
    gpos, gneg = gt[gt > 0].sum(), gt[gt < 0].sum()
    if fast:
        gNB, muNB = gt[idx_nz], mu[idx_nz]
        m1 = mu @ gt
        NAmu = m1 - gNB @ muNB
        NAsigma = var2 - np.sum(muNB * (1 - muNB) * gNB ** 2)
        K0f = lambda t: _K0(t, muNB, gNB) + NAmu * t + 0.5 * NAsigma * t * t
        K1f = lambda qq: (lambda t: _K1(t, muNB, gNB, qq) + NAmu + NAsigma * t)
        K2f = lambda t: _K2(t, muNB, gNB) + NAsigma
    else:
        K0f = lambda t: _K0(t, mu, gt)
        K1f = lambda qq: (lambda t: _K1(t, mu, gt, qq))
        K2f = lambda t: _K2(t, mu, gt)
    r1, _, c1 = _getroot(K1f(q), K2f, q, gpos, gneg, EPS25, fast=fast)
    r2, _, c2 = _getroot(K1f(qinv), K2f, qinv, gpos, gneg, EPS25, fast=fast)
    if logp_noadj is None:
        with np.errstate(divide="ignore"):
            logp_noadj = float(np.log(pval_noadj))
    spa_pvalue.last_log = logp_noadj
    if not (c1 and c2):
        return pval_noadj, False
    conv = True
    p1, s1 = _saddle_prob(r1, K0f, K2f, q)
    l1 = _saddle_prob.last_log
    p2, s2 = _saddle_prob(r2, K0f, K2f, qinv)
    l2 = _saddle_prob.last_log
    if not s1:
        conv, p1, l1 = False, pval_noadj / 2, logp_noadj - np.log(2.0)
    if not s2:
        conv, p2, l2 = False, pval_noadj / 2, logp_noadj - np.log(2.0)
    spa_pvalue.last_log = float(np.logaddexp(l1, l2))                  # add_logp (SPA.cpp:95)
    return abs(p1) + abs(p2), conv


def _lchoose(n, k):
    """HyperGeo::lCombinations (Binary_HyperGeo.cpp:172-190): R's lchoose, except that k > n gives 0 (not -inf)."""
    if k > n:
        return 0.0
    if k < 0:
        return -math.inf
    return math.lgamma(n + 1.0) - math.lgamma(k + 1.0) - math.lgamma(n - k + 1.0)


def er_group_prob(p1, p2mean, n, ncase):
    """SKATExactBin_ComputeProb_Group (ER_binary_func.cpp:23-85) + HyperGeo::Run / Get_lprob (Binary_HyperGeo.cpp:37-150):
    P(j of the k carriers are cases | ncase cases among n), carriers binned into ten fitted-probability classes with the
    class-mean odds as non-central hypergeometric weights, everybody else one class of weight 1."""
    k = len(p1)
    p1 = np.where(p1 >= 1, 0.999, p1)
    p2odd = p2mean / (1 - p2mean)
    groups, weights = [], []
    for b in range(10):
        a1, a2 = b / 10.0, (b + 1) / 10.0
        sel = (p1 >= a1) & ((p1 < a2) if b < 9 else (p1 <= a2))
        if sel.any():
            pm = p1[sel].mean()
            weights.append(pm / (1 - pm) / p2odd)
            groups.append(int(sel.sum()))
    last = [_lchoose(n - k, ncase - j) for j in range(k + 1)]          # weight of the last class is p2odd / p2odd = 1
    ref = max([0.0] + [v for v in last if v > -math.inf])
    kprob = [0.0] * (k + 1)

    def rec(lprob, idx, used):
        if idx == len(groups):
            kprob[used] += math.exp(lprob + last[used] - ref)
            return
        for i in range(groups[idx] + 1):
            if used + i <= ncase:
                rec(lprob + _lchoose(groups[idx], i) + math.log(weights[idx]) * i, idx + 1, used + i)
    rec(0.0, 0, 0)
    tot = sum(kprob)
    return np.array([v / tot for v in kprob])


def er_pvalue(g1, p1, res1, p2mean, n, ncase, epsilon=1e-6):
    """SKATExactBin_Work (ER_binary_func.cpp:186-278) for one variant (m = 1) + ComputeExact::Init / Run / GetPvalues
    (Binary_ComputeExact.cpp:296-470) in the all-exact regime (2^k <= NResampling): every case/control assignment of the k
    carriers is enumerated; statistic (sum_i g_i([i case] - p_i))^2; probability of an assignment with j cases =
    prod(odds of its cases) / e_j(odds) * P(j); result = P(stat >= observed) - P(stat == observed) / 2 (ties within epsilon).
    g1, p1, res1: genotype, fitted probability and residual of the carriers (observed cases: res1 > 0)."""
    k = len(g1)
    prob = er_group_prob(np.asarray(p1, dtype=np.float64), p2mean, n, ncase)
    odds = p1 / (1 - p1)
    z0sum = float(np.sum(-g1 * p1))
    obs = [i for i in range(k) if res1[i] > 0]
    Q = (z0sum + sum(g1[i] for i in obs)) ** 2
    stat, fprob, size = [], [], []
    denom = [0.0] * (k + 1)
    for j in range(k + 1):
        for S in _subsets(k, j):
            stat.append((z0sum + sum(g1[i] for i in S)) ** 2)
            w = 1.0
            for i in S:
                w *= odds[i]
            fprob.append(w)
            size.append(j)
            denom[j] += w
    fprob = [fprob[i] / denom[size[i]] * prob[size[i]] for i in range(len(fprob))]
    tot = sum(fprob)
    pval = same = 0.0
    for s, w in zip(stat, fprob):
        d = Q - s
        if abs(d) <= epsilon:
            d = 0.0
        if d <= 0:
            pval += w / tot
            if d == 0:
                same += w / tot
    return pval - same / 2


def _subsets(k, j):
    """size-j subsets of range(k) in lexicographic order (SKAT_Exact_Recurse, Binary_ComputeExact.cpp:112-129)."""
    import itertools
    return itertools.combinations(range(k), j)


def assign_variance_ratio(M, mac):
    """assignVarianceRatio (SAIGE_test.cpp:801-833); M["varRatio"] a float, or one value per MAC category with
    M["cateVarRatioMinMACVecExclude"] / M["cateVarRatioMaxMACVecInclude"].  MAC == the first bound exactly matches no branch
    in the reference (it keeps the previous marker's value); by convention here: first category."""
    vr = np.asarray(M["varRatio"], dtype=np.float64).reshape(-1)
    if len(vr) == 1:
        return float(vr[0])
    lo, hi = list(M["cateVarRatioMinMACVecExclude"]), list(M["cateVarRatioMaxMACVecInclude"])
    for i in range(len(hi)):
        if lo[i] < mac <= hi[i]:
            return float(vr[i])
    if mac <= lo[0]:
        return float(vr[0])
    return float(vr[-1])


def impute_and_flip(Graw, impute_method="best_guess", dosage_zerod_cutoff=0.0, dosage_zerod_MAC_cutoff=0.0):
    """getOneMarker's counts + imputeGenoAndFlip (UTIL.cpp:58-135) for one marker: (G in the tested coding, flip, MAC)."""
    n = len(Graw)
    miss = Graw < 0
    cnt = n - int(miss.sum())
    alt_freq = float(Graw[~miss].sum()) / cnt / 2 if cnt > 0 else 0.0
    mac = min(alt_freq, 1 - alt_freq) * n * (1 - miss.sum() / n) * 2
    G = Graw.copy()
    flip = alt_freq > 0.5
    if flip:
        G = 2 - G
        alt_freq = 1 - alt_freq
    if miss.any():
        impute_g = {"best_guess": math.floor(2 * alt_freq + 0.5), "mean": 2 * alt_freq, "minor": 0.0}[impute_method]
        G[miss] = impute_g
        mac = mac + impute_g * int(miss.sum())
    if dosage_zerod_cutoff > 0 and mac <= dosage_zerod_MAC_cutoff:
        G[np.abs(G) <= dosage_zerod_cutoff] = 0.0
    return G, flip, min(G.sum(), 2 * n - G.sum())


def condition_factors(M, cond_Graw, **impute_kw):
    """assign_conditionMarkers_factors (Main.cpp:2002-2179): for every conditioning marker the covariate-adjusted genotype
    gtilde (tested coding), P1 row = sqrt(vr) gtilde, P2 column = sqrt(vr) gtilde % mu2 tau0 (the is_region branch of
    getMarkerPval, SAIGE_test.cpp:780-783), its score; VarInv = pinv(P1 P2)."""
    P1, P2, T = [], [], []
    for Graw in cond_Graw:
        G, _, mac = impute_and_flip(np.asarray(Graw, dtype=np.float64), **impute_kw)
        vr = assign_variance_ratio(M, mac)
        gt = G - M["XXVX_inv"] @ (M["XV"] @ G)
        P1.append(np.sqrt(vr) * gt)
        P2.append(np.sqrt(vr) * gt * M["mu2"] * M["tau"][0])
        T.append(score_test_fast(M, G, np.nonzero(G != 0)[0], vr)["Tstat"])
    P1, P2 = np.array(P1), np.array(P2).T
    return dict(P2=P2, VarInv=np.linalg.pinv(P1 @ P2), Tstat=np.array(T))


def score_test_fast(M, G, idx, var_ratio=None):
    """scoreTestFast (SAIGE_test.cpp:212-292)."""
    g1, X1, A1, res1 = G[idx], M["X"][idx], M["XVX_inv_XV"][idx], M["res"][idx]
    Z = A1.T @ g1
    Bv = X1 @ Z
    gt1 = g1 - Bv
    if M["trait"] == "binary":
        mu21 = M["mu2"][idx]
        var2 = float(Z @ M["XVX"] @ Z) - float(Bv ** 2 @ mu21) + float(gt1 ** 2 @ mu21)
    else:
        var2 = float(Z @ M["XVX"] @ Z) * M["tau"][0] + float(g1 @ g1) - 2 * float(g1 @ Bv)
    var1 = var2 * (float(np.asarray(M["varRatio"]).reshape(-1)[0]) if var_ratio is None else var_ratio)
    S = (float(res1 @ gt1) - float((M["S_a"] - res1 @ X1) @ Z)) / M["tau"][0]
    stat = S * S / var1
    ok = not (var1 <= np.finfo(float).tiny or not np.isfinite(stat))
    pval = float(stats.chi2.sf(stat, 1)) if ok else 1.0
    logp = log_erfc(np.sqrt(stat / 2)) if ok else 0.0                  # the log-scale p-value of SAIGE_test.cpp:273-283
    beta = S / var1
    return dict(Beta=beta, seBeta=abs(beta) / np.sqrt(abs(stat)), pval=pval, logp=logp, Tstat=S, var1=var1, var2=var2)


def firth_fit(gt, y, offset, maxit=50, maxstep=15, xconv=1e-5, gconv=1e-5):
    """fast_logistf_fit_simple (SAIGE_test.cpp:893-986) for x = [1, gtilde], init 0: Firth's penalised-likelihood Newton
    iteration with the hat values of sqrt(W) x.  Returns (beta_G, sebeta_G, converged)."""
    x = np.column_stack([np.ones(len(gt)), gt])
    beta = np.zeros(2)
    pi = 1.0 / (np.exp(-x @ beta - offset) + 1.0)
    it, conv, cov = 0, False, np.full((2, 2), np.nan)
    while it <= maxit:
        w = pi * (1 - pi)
        xw = x * np.sqrt(w)[:, None]
        fisher = xw.T @ xw
        try:
            cov = np.linalg.inv(fisher)
        except np.linalg.LinAlgError:
            break
        if not np.all(np.isfinite(cov)) or np.linalg.det(fisher) <= 0:
            break
        h = np.einsum("ij,jk,ik->i", xw, cov, xw)
        u = x.T @ ((y - pi) + h * (0.5 - pi))
        delta = cov @ u
        mx = np.max(np.abs(delta)) / maxstep
        if mx > 1:
            delta = delta / mx
        it += 1
        beta = beta + delta
        pi = 1.0 / (np.exp(-x @ beta - offset) + 1.0)
        if it == maxit or (np.max(np.abs(delta)) <= xconv and np.all(np.abs(u) <= gconv)):
            conv = True
            break
    if np.any(np.isnan(cov)):
        return np.nan, np.nan, conv
    return float(beta[1]), float(np.sqrt(cov[1, 1])), conv


def test_marker(M, Graw, min_maf=0.0, min_mac=0.5, max_missing=0.15, spa_cutoff=2.0, se_two_sided=True, is_Firth_beta=False,
                pCutoffforFirth=0.01, firth_se_from_fit=True, max_MAC_for_ER=-1.0, impute_method="best_guess",
                dosage_zerod_cutoff=0.0, dosage_zerod_MAC_cutoff=0.0, cond=None):
    """One pass of the mainMarkerInCPP loop body (Main.cpp:229-520).  Returns None when the marker is filtered.
    Graw: copies of the ALT allele per model sample, hard calls (0/1/2) or dosages in [0, 2]; negative = missing."""
    test_marker.__test__ = False
    n = len(Graw)
    miss = Graw < 0
    cnt = n - int(miss.sum())
    alt_counts = float(Graw[~miss].sum())
    alt_freq = alt_counts / cnt / 2 if cnt > 0 else 0.0
    missing_rate = miss.sum() / n
    maf = min(alt_freq, 1 - alt_freq)
    mac = maf * n * (1 - missing_rate) * 2
    if missing_rate > max_missing or maf < min_maf or mac < min_mac:
        return None
    # imputeGenoAndFlip (UTIL.cpp:58-135), best_guess
    G = Graw.copy()
    flip = alt_freq > 0.5
    if flip:
        G = 2 - G
        alt_freq = 1 - alt_freq
    if miss.any():
        # best_guess: std::round (half away from zero); mean: 2 f; minor: 0 (UTIL.cpp:80-93)
        impute_g = {"best_guess": math.floor(2 * alt_freq + 0.5), "mean": 2 * alt_freq, "minor": 0.0}[impute_method]
        G[miss] = impute_g
        mac = mac + impute_g * int(miss.sum())
    if dosage_zerod_cutoff > 0 and mac <= dosage_zerod_MAC_cutoff:
        G[np.abs(G) <= dosage_zerod_cutoff] = 0.0                  # arma clean() (UTIL.cpp:105-109): small dosages of rare variants
    alt_count = G.sum()
    alt_freq = alt_count / (2 * n)
    if flip:
        alt_freq, alt_count = 1 - alt_freq, 2 * n - alt_count
    idx = np.nonzero(G != 0)[0]
    st = score_test_fast(M, G, idx, assign_variance_ratio(M, min(alt_count, 2 * n - alt_count)))
    std_stat = abs(st["Tstat"]) / np.sqrt(st["var1"])
    pval, se, is_spa = st["pval"], st["seBeta"], False
    logp = st["logp"]
    islog = st["pval"] == 0                                             # ispvallog: the p-value underflowed (SAIGE_test.cpp:270-284)
    # exact test of rare variants (Main.cpp:408-422: MAC after imputation <= g_MACCutoffforER; SAIGE_test.cpp:426-431, 592-620)
    mac_after = min(alt_count, 2 * n - alt_count)
    is_er = (M["trait"] == "binary" and mac_after <= max_MAC_for_ER and (std_stat > spa_cutoff or np.isnan(std_stat)))
    if is_er:
        mu = M["mu"]
        p2mean = float(np.delete(mu, idx).mean())
        pval = er_pvalue(G[idx], mu[idx], M["res"][idx], p2mean, n, int((M["y"] == 1).sum()))
        se = 0.0 if pval / 2 <= 0 else abs(st["Beta"]) / abs(stats.norm.ppf(pval / 2))
        logp = float(np.log(pval))
    elif np.isfinite(std_stat) and std_stat > spa_cutoff and M["trait"] == "binary":
        gt = G - M["XXVX_inv"] @ (M["XV"] @ G)                      # getadjGFast
        m1 = float(M["mu"] @ gt)
        q = st["Tstat"] / np.sqrt(st["var1"] / st["var2"]) + m1
        qinv = -abs(q - m1) + m1 if q - m1 > 0 else (m1 if q == m1 else abs(q - m1) + m1)
        fast = (n - len(idx)) / n >= 0.5
        pspa, conv = spa_pvalue(M["mu"], gt, q, qinv, st["pval"], idx, fast, st["var2"], st["logp"])
        if conv and (pspa != 0 or islog):                              # SAIGE_test.cpp:541: only the linear scale un-converges on 0
            is_spa = True
            pval, logp = pspa, spa_pvalue.last_log
            # SE from the SPA p-value: the reference's bundled golden table corresponds to |qnorm(p/2)| (upstream SAIGE);
            # this fork's source has qnorm(p, upper tail) (SAIGE_test.cpp:523-526) -> se_two_sided=False
            se = abs(st["Beta"]) / qnorm_from_logp(logp - np.log(2.0) if se_two_sided else logp)
    beta, is_firth, firth_conv = st["Beta"], False, False
    if is_Firth_beta and M["trait"] == "binary" and logp <= np.log(pCutoffforFirth):
        # SAIGE_test.cpp:573-633.  firth_se_from_fit: SE = the fit's own sqrt(cov[1,1]) (what the reference's bundled
        # positive-signal result holds); otherwise |beta| / |qnorm| of the p-value as this fork's source has it (:632)
        gt = G - M["XXVX_inv"] @ (M["XV"] @ G)
        beta, se_fit, firth_conv = firth_fit(gt, M["y"], M["offset"])
        is_firth = True
        se = se_fit if firth_se_from_fit else abs(beta) / qnorm_from_logp(logp - np.log(2.0) if (se_two_sided or is_er) else logp)
    sgn = -1.0 if flip else 1.0
    extra = {}
    if cond is not None:
        # t_isCondition (SAIGE_test.cpp:640-790).  p.value_c = the adjusted p-value itself, SE_c = |BETA_c| / |qnorm(p/2)|: what
        # the reference's bundled conditional table holds (this fork's source prints half the SPA p-value, :752)
        vr_val = assign_variance_ratio(M, min(alt_count, 2 * n - alt_count))
        gt = G - M["XXVX_inv"] @ (M["XV"] @ G)
        g1p2 = np.sqrt(vr_val) * (gt @ cond["P2"])
        Tc = st["Tstat"] - float(g1p2 @ cond["VarInv"] @ cond["Tstat"])
        vc = st["var1"] - float(g1p2 @ cond["VarInv"] @ g1p2)
        stat_c = Tc * Tc / vc
        if vc <= np.finfo(float).tiny or not np.isfinite(stat_c):
            p_na_c, stat_c = 1.0, 0.0
        else:
            p_na_c = float(stats.chi2.sf(stat_c, 1))
        beta_c = Tc / vc
        with np.errstate(divide="ignore", invalid="ignore"):
            se_c = abs(beta_c) / np.sqrt(stat_c)
        p_c = p_na_c
        if M["trait"] == "binary" and stat_c > spa_cutoff ** 2:
            m1 = float(M["mu"] @ gt)
            q_c = Tc / np.sqrt(vc / st["var2"]) + m1
            qinv_c = -abs(q_c - m1) + m1 if q_c - m1 > 0 else (m1 if q_c == m1 else abs(q_c - m1) + m1)
            fast = (n - len(idx)) / n >= 0.5
            pspa_c, conv_c = spa_pvalue(M["mu"], gt, q_c, qinv_c, p_na_c, idx, fast, st["var2"])
            if conv_c and pspa_c != 0:
                p_c = pspa_c
                se_c = abs(beta_c) / abs(stats.norm.isf(pspa_c / 2))
        extra = dict(BETA_c=sgn * beta_c, SE_c=se_c, Tstat_c=sgn * Tc, var_c=vc, p_value_c=p_c, p_value_NA_c=p_na_c)
    y = M["y"]
    case, ctrl = y == 1, y == 0
    afc = G[case].mean() / 2 if case.any() else np.nan
    aft = G[ctrl].mean() / 2 if ctrl.any() else np.nan
    if flip:
        afc, aft = 1 - afc, 1 - aft
    # is_output_moreDetails (Main.cpp:510-525): dosage ranges [1.5, 2] / [0.5, 1.5) of the flipped, imputed vector
    hom, het = (G >= 1.5) & (G <= 2), (G >= 0.5) & (G < 1.5)
    n_case_hom, n_case_het = int((hom & case).sum()), int((het & case).sum())
    n_ctrl_hom, n_ctrl_het = int((hom & ctrl).sum()), int((het & ctrl).sum())
    if flip:
        n_case_hom, n_ctrl_hom = int(case.sum()) - n_case_het - n_case_hom, int(ctrl.sum()) - n_ctrl_het - n_ctrl_hom
    return dict(**extra, N_case_hom=n_case_hom, N_case_het=n_case_het, N_ctrl_hom=n_ctrl_hom, N_ctrl_het=n_ctrl_het, AC_Allele2=alt_count, AF_Allele2=alt_freq, MissingRate=missing_rate, BETA=sgn * beta, SE=se,
                Tstat=sgn * st["Tstat"], var=st["var1"], p_value=pval, p_value_NA=st["pval"], log_p_value=logp, log_p_value_NA=st["logp"],
                Is_SPA=is_spa, Is_ER=is_er,
                Is_Firth=is_firth, Firth_converged=firth_conv,
                AF_case=afc, AF_ctrl=aft, N_case=int(case.sum()), N_ctrl=int(ctrl.sum()))
