"""Step-1 driver: the part of src/SAIGE/R/SAIGE_fitGLMM_fast.R ("FG.R") that sits between `fitNULLGLMM` and the
native exports, restated in Python so that the whole null-GLMM fit can be run and timed without R.

In a real deployment this layer is the UNCHANGED R code; it is mirrored here, function by
function and with the same names, because R is not available in the build/bench environment.  It only calls
the export mirror in api.py (i.e. the C ABI); no numerical work on genotypes happens here.

  Get_Coef / Get_Coef_LOCO              FG.R:2-35, 42-73
  glmmkin_ai_PCG_Rcpp_Binary            FG.R:79-304
  glmmkin_ai_PCG_Rcpp_Quantitative      FG.R:309-549
  ScoreTest_NULL_Model                  FG.R:579-591
  Covariate_Transform[_Back]            FG.R:1612-1659
  extractVarianceRatio                  FG.R:2152-2423 (single ratio, full-GRM path)
  updateChrStartEndIndexVec             R/Util.R:29-65
"""
import time

from concurrent.futures import ThreadPoolExecutor

import numpy as np


class Binomial:
    name = "binomial"

    @staticmethod
    def linkinv(eta):
        return 1.0 / (1.0 + np.exp(-eta))

    @staticmethod
    def mu_eta(eta):
        e = np.exp(-np.abs(eta))
        return np.maximum(e / (1.0 + e) ** 2, np.finfo(float).eps)

    @staticmethod
    def variance(mu):
        return mu * (1.0 - mu)


class Gaussian:
    name = "gaussian"
    linkinv = staticmethod(lambda eta: eta)
    mu_eta = staticmethod(lambda eta: np.ones_like(eta))
    variance = staticmethod(lambda mu: np.ones_like(mu))


def glm_fit(y, X, family, offset=None, maxit=25, epsilon=1e-8):
    """R's glm.fit IRLS; provides fit0 (FG.R:1119)."""
    n = len(y)
    offset = np.zeros(n) if offset is None else offset
    if family.name == "binomial":
        mu = (y + 0.5) / 2.0
        eta = np.log(mu / (1 - mu))
    else:
        mu = y.astype(np.float64).copy()
        eta = mu.copy()
    devold = np.inf
    coef = np.zeros(X.shape[1])
    for _ in range(maxit):
        me = family.mu_eta(eta)
        z = (eta - offset) + (y - mu) / me
        sw = np.sqrt(me ** 2 / family.variance(mu))
        coef, *_ = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=None)
        eta = X @ coef + offset
        mu = family.linkinv(eta)
        if family.name == "binomial":
            with np.errstate(divide="ignore", invalid="ignore"):
                d = 2 * (np.where(y > 0, y * np.log(y / mu), 0) + np.where(y < 1, (1 - y) * np.log((1 - y) / (1 - mu)), 0))
            dev = d.sum()
        else:
            dev = ((y - mu) ** 2).sum()
        if abs(dev - devold) / (abs(dev) + 0.1) < epsilon:
            break
        devold = dev
    return dict(coef=coef, eta=eta, mu=mu, y=y, offset=offset, family=family, X=X)


def Covariate_Transform(X1):
    Q, R = np.linalg.qr(X1)
    return Q * np.sqrt(X1.shape[0]), dict(qrr=R, N=X1.shape[0])


def Covariate_Transform_Back(coef, param):
    return np.linalg.solve(param["qrr"], coef * np.sqrt(param["N"]))


def ScoreTest_NULL_Model(mu, mu2, y, X):
    V = np.asarray(mu2, dtype=np.float64)
    res = y - mu
    XVt = X * V[:, None]                      # N x p once; XV is its transpose (a view)
    XVX = X.T @ XVt
    XVX_inv = np.linalg.inv(XVX)
    XXVX_inv = X @ XVX_inv
    return dict(XV=XVt.T, XVX=XVX, XXVX_inv=XXVX_inv, XVX_inv=XVX_inv, S_a=res @ X,
                XVX_inv_XV=XXVX_inv * V[:, None], V=V)


def updateChrStartEndIndexVec(geno, chrVec):
    chrVec = np.asarray(chrVec)
    start, end = [], []
    for c in range(1, 23):
        idx = np.nonzero(chrVec == c)[0]
        if len(idx):
            if start and start[-1] != -1 and (idx.min() <= start[-1] or idx.max() <= end[-1]):
                raise ValueError("ERROR! chromosomes need to be ordered from 1 to 22 in the bim file.")
            start.append(int(idx.min())); end.append(int(idx.max()))
        else:
            start.append(-1); end.append(-1)
    LOCO = sum(s != -1 for s in start) > 1
    geno.setStartEndIndexVec(start, end)
    return LOCO, start, end


def r_unif_rand(seed, n):
    """n successive unif_rand() values of R's default generator (Mersenne-Twister) after set.seed(seed): the initial
    scrambling of RNG_Init (50 + 625 steps of seed = 69069 seed + 1), the standard MT19937 stream scaled by
    2.3283064365386963e-10, and R's fixup into the open interval.  set.seed(1); runif(3) = 0.2655087 0.3721239 0.5728534."""
    s = np.uint32(seed)
    words = np.empty(625, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(50):
            s = np.uint32(np.uint32(69069) * s + np.uint32(1))
        for j in range(625):
            s = np.uint32(np.uint32(69069) * s + np.uint32(1))
            words[j] = s
    bg = np.random.MT19937()
    st = bg.state
    st["state"]["key"] = words[1:].copy()          # i_seed[0] is mti (forced to 624 by FixupSeeds), the rest is mt[]
    st["state"]["pos"] = 624
    bg.state = st
    u = bg.random_raw(n).astype(np.float64) * 2.3283064365386963e-10
    i2 = 2.328306437080797e-10
    u = np.where(u <= 0.0, 0.5 * i2, u)
    return np.where(1.0 - u <= 0.0, 1.0 - 0.5 * i2, u)


class ProbeStream:
    """The probe sequence of the trace estimator: GetTrace re-seeds to 200 on every call (FG.cpp:3114) and then draws
    2*rbinom(N,1,0.5)-1 per run (`nb`, FG.cpp:3052), so every call sees the same sequence.  Here the sequence is a fixed
    N x nmax Rademacher matrix; `fresh()` returns a draw(n) callable that starts again at column 0.
    rng="numpy": numpy's default generator (what the parity tests share with the oracle);
    rng="R": bit-for-bit the stream R produces -- rbinom(1, 0.5) consumes one unif_rand() and returns u >= 0.5
    (nmath/rbinom.c, inversion branch), run r / sample i is draw r*N + i."""

    def __init__(self, N, nmax=130, seed=200, rng="numpy"):
        if rng == "R":
            u = r_unif_rand(seed, N * nmax)
            self.U = np.asfortranarray((2.0 * (u >= 0.5) - 1.0).reshape(nmax, N).T)
        else:
            gen = np.random.default_rng(seed)
            self.U = np.asfortranarray(gen.integers(0, 2, size=(N, nmax)).astype(np.float64) * 2.0 - 1.0)

    def fresh(self):
        pos = [0]

        def draw(n):
            out = self.U[:, pos[0]:pos[0] + n]
            if out.shape[1] != n:
                raise RuntimeError("probe matrix exhausted")
            pos[0] += n
            return out
        return draw


def Get_Coef(geno, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, loco=False, native=False):
    """FG.R:2-35 (loco: 42-73).  native=True: the same loop inside the library (sgb_get_coef), N-vectors device resident."""
    if native:
        return geno.Get_Coef(y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, loco=loco)
    tol_coef = 0.1
    mu = family.linkinv(eta0)
    me = family.mu_eta(eta0)
    Y = eta0 - offset + (y - mu) / me
    sqrtW = me / np.sqrt(family.variance(mu))
    W = sqrtW ** 2
    for _ in range(maxiter):
        rc = geno.getCoefficients(Y, X, W, tau, maxiterPCG, tolPCG, loco)
        alpha = rc["alpha"]
        eta = rc["eta"] + offset
        mu = family.linkinv(eta)
        me = family.mu_eta(eta)
        Y = eta - offset + (y - mu) / me
        sqrtW = me / np.sqrt(family.variance(mu))
        W = sqrtW ** 2
        if np.max(np.abs(alpha - alpha0) / (np.abs(alpha) + np.abs(alpha0) + tol_coef)) < tol_coef:
            break
        alpha0 = alpha
    return dict(Y=Y, alpha=alpha, eta=eta, W=W, cov=rc["cov"], sqrtW=sqrtW, Sigma_iY=rc["Sigma_iY"],
                Sigma_iX=rc["Sigma_iX"], mu=mu)


def glmmkin_ai_PCG(geno, fit0, probes, trait="binary", tauInit=(0.0, 0.0), maxiter=20, tol=0.02, nrun=30,
                   tolPCG=1e-5, maxiterPCG=500, traceCVcutoff=0.0025, LOCO=False, verbose=False, timings=None, native_loops=False):
    """glmmkin.ai_PCG_Rcpp_Binary / _Quantitative after setgeno (FG.R:127-304, 340-549).
    native_loops=True: Get_Coef, the LOCO refit loop and the first probe batch run inside the library (sgb_get_coef,
    sgb_get_coef_loco_all, sgb_set_probe_stream_fixed) instead of in this mirror of the R code; same results (the IRLS
    update uses the device's exp instead of numpy's: differences at the 1e-15 level).  native_loops=True runs the whole function
    as ONE library call (sgb_glmmkin_ai_pcg), native_loops="calls" keeps this loop and calls the library once per R loop."""
    nat = dict(native=True) if native_loops else {}
    set_fixed = getattr(geno, "setProbeStreamFixed", None)      # a device context has it; the oracle-backed test double does not
    if set_fixed is not None:
        set_fixed(bool(native_loops))
    if native_loops is True:
        return _glmmkin_ai_PCG_one_call(geno, fit0, probes, trait, tauInit, maxiter, tol, nrun, tolPCG, maxiterPCG, traceCVcutoff, LOCO,
                                        verbose, timings)
    y, X, offset, family = fit0["y"], fit0["X"], fit0["offset"], fit0["family"]
    X = np.asfortranarray(X, dtype=np.float64)      # column-major once: every ABI call takes it without another copy
    n = len(y)
    quant = trait == "quantitative"
    eta = fit0["eta"]
    alpha0, eta0 = fit0["coef"], eta
    tau = np.array([0.0, 0.0])
    tauInit = np.asarray(tauInit, dtype=np.float64)
    if not quant:
        tau[0] = 1.0
        tau[1] = 0.1 if tauInit[1] == 0 else tauInit[1]
    elif tauInit.sum() == 0:
        tau[:] = (1.0, 0.0)
    else:
        tau[:] = tauInit
    tau0 = tau.copy()
    t0 = time.time()
    rc = Get_Coef(geno, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, **nat)
    if not quant:
        re = geno.getAIScore(rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], nrun, maxiterPCG, tolPCG,
                             traceCVcutoff, probes.fresh())
        tau[1] = max(0.0, tau0[1] + tau0[1] ** 2 * (re["YPAPY"] - re["Trace"]) / n)
    else:
        re = geno.getAIScore_q(rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], nrun, maxiterPCG,
                               tolPCG, traceCVcutoff, probes.fresh())
        tau[1] = max(0.0, tau0[1] + tau0[1] ** 2 * (re["YPAPY"] - re["Trace"][1]) / n)
        tau[0] = max(0.0, tau0[0] + tau0[0] ** 2 * (re["YPA0PY"] - re["Trace"][0]) / n)
    if verbose:
        print("Variance component estimates:", tau)
    tau_path = [tau.copy()]
    alpha = fit0["coef"] if quant else rc["alpha"]
    i = 0
    for i in range(1, maxiter + 1):
        alpha0 = alpha if quant else rc["alpha"]
        tau0 = tau.copy()
        eta0 = eta
        rc = Get_Coef(geno, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, **nat)
        fit = (geno.fitglmmaiRPCG_q if quant else geno.fitglmmaiRPCG)(
            rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], nrun, maxiterPCG, tolPCG, tol,
            traceCVcutoff, probes.fresh())
        tau = np.asarray(fit["tau"], dtype=np.float64)
        alpha, eta = rc["alpha"], rc["eta"]
        tau_path.append(tau.copy())
        if verbose:
            print("Iteration", i, "tau:", tau)
        if quant and tau[0] <= 0:
            raise RuntimeError("ERROR! The first variance component parameter estimate is 0")
        if (not quant and tau[1] == 0) or (quant and tau[1] <= 0):
            break
        if np.max(np.abs(tau - tau0) / (np.abs(tau) + np.abs(tau0) + tol)) < tol:
            break
        if np.max(tau) > tol ** (-2):
            i = maxiter
            break
    rc = Get_Coef(geno, y, X, tau, family, alpha, eta, offset, maxiterPCG, tolPCG, maxiter, **nat)
    alpha, eta, mu = rc["alpha"], rc["eta"], rc["mu"]
    mu2 = mu * (1 - mu) if not quant else np.full(n, 1.0 / tau[0])
    out = dict(theta=tau, coefficients=alpha, linear_predictors=eta, fitted_values=mu, Y=rc["Y"], residuals=y - mu,
               cov=rc["cov"], converged=i < maxiter, obj_noK=ScoreTest_NULL_Model(mu, mu2, y, X), y=y, X=X,
               traitType=trait, LOCO=LOCO, tau_path=tau_path, offset=offset)
    if timings is not None:
        timings["fit_s"] = time.time() - t0
    if LOCO:
        t1 = time.time()
        geno.set_Diagof_StdGeno_LOCO()
        out["LOCOResult"] = []
        # The refits are sequential (chromosome j starts from chromosome j-1's alpha / eta, FG.R:1253-1275), but the
        # score-test matrices of chromosome j are output only: they are computed on a host thread while the GPU already
        # solves chromosome j+1 (numpy and the ctypes call both release the GIL).
        pending = []
        # several workers: at 8 GPUs a chromosome's solves take ~25 ms, about what one score-test matrix set costs on a host
        # core, so a single worker would become the critical path of the refit phase
        with ThreadPoolExecutor(max_workers=4) as pool:
            early = {}

            def on_chrom(c, rl_c):          # chromosome c is done: its score-test matrices start while the GPU refits c + 1
                mu_c = rl_c["mu"]
                early[c] = pool.submit(ScoreTest_NULL_Model, mu_c, mu_c * (1 - mu_c) if not quant else np.full(n, 1.0 / tau[0]), y, X)
            allchr = (geno.Get_Coef_LOCO_all(y, X, tau, family, alpha, eta, offset, maxiterPCG, tolPCG, maxiter, on_chrom=on_chrom)
                      if native_loops else None)
            for j, (s, e) in enumerate(zip(geno_start_vec(geno), geno_end_vec(geno))):
                if s == -1 or e == -1:
                    out["LOCOResult"].append(dict(isLOCO=False))
                    continue
                if native_loops:
                    rl = allchr[j]
                else:
                    geno.setStartEndIndex(s, e, j)
                    rl = Get_Coef(geno, y, X, tau, family, alpha, eta, offset, maxiterPCG, tolPCG, maxiter, loco=True)
                alpha, eta, mu = rl["alpha"], rl["eta"], rl["mu"]
                mu2 = mu * (1 - mu) if not quant else np.full(n, 1.0 / tau[0])
                entry = dict(isLOCO=True, coefficients=alpha, linear_predictors=eta, fitted_values=mu,
                             Y=rl["Y"], residuals=y - mu, cov=rl["cov"])
                pending.append((entry, early[j] if native_loops else pool.submit(ScoreTest_NULL_Model, mu, mu2, y, X)))
                out["LOCOResult"].append(entry)
            for entry, fut in pending:
                entry["obj_noK"] = fut.result()
        if timings is not None:
            timings["loco_s"] = time.time() - t1
    return out


def _glmmkin_ai_PCG_one_call(geno, fit0, probes, trait, tauInit, maxiter, tol, nrun, tolPCG, maxiterPCG, traceCVcutoff, LOCO, verbose,
                             timings):
    """glmmkin_ai_PCG through sgb_glmmkin_ai_pcg; what stays here is what the R function does after its loops: the result list and
    the score-test matrices (ScoreTest_NULL_Model, FG.R:231-238, 283-288).  Those are started on host threads from the library's
    per-chromosome callback, i.e. while the GPU already refits the next chromosome."""
    y, X, offset = fit0["y"], np.asfortranarray(fit0["X"], dtype=np.float64), fit0["offset"]
    quant = trait == "quantitative"
    n = len(y)
    t0 = time.time()
    futures = {}
    with ThreadPoolExecutor(max_workers=4) as pool:
        def on_chrom(c, r):
            mu = r["fitted_values"] if c < 0 else r["mu"]
            tau0 = r["theta"][0] if c < 0 else tau_box[0]
            if c < 0:
                tau_box[0] = float(r["theta"][0])
            mu2 = mu * (1 - mu) if not quant else np.full(n, 1.0 / tau0)
            futures[c] = pool.submit(ScoreTest_NULL_Model, mu, mu2, y, X)
        tau_box = [1.0]
        r = geno.glmmkin_ai_PCG(trait, y, X, offset, fit0["coef"], fit0["eta"], tauInit, maxiter, tol, nrun, tolPCG, maxiterPCG,
                                traceCVcutoff, LOCO, probes.fresh, on_chrom=on_chrom)
        tau, mu = r["theta"], r["fitted_values"]
        if verbose:
            print("Final", tau)
        out = dict(theta=tau, coefficients=r["coefficients"], linear_predictors=r["linear_predictors"], fitted_values=mu, Y=r["Y"],
                   residuals=y - mu, cov=r["cov"], converged=r["converged"], y=y, X=X, traitType=trait, LOCO=LOCO, offset=offset,
                   n_outer=r["n_outer"], obj_noK=futures[-1].result())
        if LOCO:
            out["LOCOResult"] = []
            for c, rl in enumerate(r["loco"]):
                if rl is None:
                    out["LOCOResult"].append(dict(isLOCO=False))
                    continue
                out["LOCOResult"].append(dict(isLOCO=True, coefficients=rl["alpha"], linear_predictors=rl["eta"], fitted_values=rl["mu"],
                                              Y=rl["Y"], residuals=y - rl["mu"], cov=rl["cov"], obj_noK=futures[c].result()))
    if timings is not None:
        timings["fit_s"] = time.time() - t0
        timings["loco_s"] = 0.0
    return out


def geno_start_vec(geno):
    return geno._loco_start


def geno_end_vec(geno):
    return geno._loco_end


def set_loco_ranges(geno, chr_of_qc_marker):
    """FG.R:116-125: chromosome ranges over QC-passing markers -> setStartEndIndexVec."""
    LOCO, start, end = updateChrStartEndIndexVec(geno, chr_of_qc_marker)
    geno._loco_start, geno._loco_end = start, end
    return LOCO


def extractVarianceRatio(geno, model, family, marker_order, numMarkers=30, maxiterPCG=500, tolPCG=1e-5,
                         ratioCVcutoff=0.001, batch=True, native_loops=False):
    """FG.R:2152-2423 (one ratio).  `marker_order` = the caller's `sample(MACindex)` permutation (FG.R:2242).
    With batch=True the getSigma_G solves of a round (numMarkers, then +10 per CV retry) run as one multi-column
    PCG; results are identical to the sequential loop because each column follows its own recurrence."""
    mu, eta, y, X = model["fitted_values"], model["linear_predictors"], model["y"], model["X"]
    tau, noK = model["theta"], model["obj_noK"]
    me = family.mu_eta(eta)
    W = (me / np.sqrt(family.variance(mu))) ** 2
    Sigma_iX = geno.getSigma_X(W, tau, X, maxiterPCG, tolPCG)
    use_vr = geno.getIsVarRatioGeno() and geno.Mvr > 0
    N = geno.N
    ratios, pos, target = [], 0, numMarkers
    XtSiX = X.T @ Sigma_iX
    while True:
        take = marker_order[pos:pos + (target - len(ratios))]
        pos += len(take)
        if len(take) and native_loops:
            # the marker loop inside the library (sgb_variance_ratio_markers): genotype columns, covariate adjustment, the
            # multi-column solve and the quadratic forms stay on the device
            mu2 = mu * (1 - mu) if model["traitType"] == "binary" else None
            for c0 in range(0, len(take), 128):
                v1, v2, _ = geno.varianceRatioMarkers(take[c0:c0 + 128], use_vr, W, tau, X, noK["XV"], noK["XXVX_inv"], Sigma_iX,
                                                      mu2, maxiterPCG, tolPCG)
                ratios.extend((v1 / v2).tolist())
        elif len(take):
            Gs, ACs = [], []
            for i in take:
                G0 = (geno.Get_OneSNP_Geno_forVarRatio(i) if use_vr else geno.Get_OneSNP_Geno(i)).astype(np.float64)
                if G0.sum() / (2 * N) > 0.5:
                    G0 = 2 - G0
                ACs.append(G0.sum())
                Gs.append(G0 - noK["XXVX_inv"] @ (noK["XV"] @ G0))
            Gm = np.column_stack(Gs)
            if batch:
                SiG = geno.getSigma_G(W, tau, Gm, maxiterPCG, tolPCG)
            else:
                SiG = np.column_stack([geno.getSigma_G(W, tau, Gm[:, j], maxiterPCG, tolPCG) for j in range(Gm.shape[1])])
            for j in range(Gm.shape[1]):
                Gt, AC, Sg = Gm[:, j], ACs[j], SiG[:, j]
                gn = Gt / np.sqrt(AC)
                var1 = (Gt @ Sg - Gt @ Sigma_iX @ np.linalg.solve(XtSiX, X.T @ Sg)) / AC
                var2 = float((mu * (1 - mu)) @ (gn * gn)) if model["traitType"] == "binary" else float(gn @ gn)
                ratios.append(var1 / var2)
        cv = geno.calCV(np.array(ratios))
        if cv > ratioCVcutoff and pos < len(marker_order):
            target += 10
        else:
            break
    return float(np.mean(ratios)), ratios


def extractVarianceRatio_cate(geno, model, family, mac_of_marker, marker_ok, rng, cateVarRatioMinMACVecExclude=(10, 20.5),
                              cateVarRatioMaxMACVecInclude=(20.5,), cateVarRatioIndexVec=None, numMarkers=30, **kw):
    """The categorical branch of extractVarianceRatio (FG.R:2244-2280 + the per-category loop :2287-2411): one ratio per MAC
    category, category k = (min[k], max[k]], the last one open-ended when max is one entry shorter; categories switched off
    in cateVarRatioIndexVec get the ratio 1.  `mac_of_marker`: MAC of every candidate marker (the hold-out store or the GRM
    store), `marker_ok`: which of them may be used (autosomes), `rng`: numpy generator standing for R's sample()."""
    lo, hi = list(cateVarRatioMinMACVecExclude), list(cateVarRatioMaxMACVecInclude)
    idxvec = [1] * len(lo) if cateVarRatioIndexVec is None else list(cateVarRatioIndexVec)
    ncat = len(idxvec)
    mac = np.asarray(mac_of_marker, dtype=np.float64)
    out = []
    for k in range(ncat):
        if k < ncat - 1 or len(hi) == ncat:
            sel, label = (mac > lo[k]) & (mac <= hi[k]), "%g< MAC <= %g" % (lo[k], hi[k])
        else:
            sel, label = mac > lo[k], "%g< MAC" % lo[k]
        if idxvec[k] != 1:
            out.append((1.0, []))
            continue
        members = np.nonzero(sel)[0]
        if len(members) < numMarkers:
            raise ValueError("ERROR! number of genetic variants in %s is lower than %d\nPlease include more markers in this MAC "
                             "category in the plink file" % (label, numMarkers))
        order = rng.permutation(members)
        order = order[np.asarray(marker_ok)[order]]
        out.append(extractVarianceRatio(geno, model, family, order, numMarkers=numMarkers, **kw))
    return out
