"""`fitNULLGLMM`: the user-facing entry of step 1, same arguments and same output files as the reference's R function
(/root/reference/src/SAIGE/R/SAIGE_fitGLMM_fast.R:650-1610), full-GRM path.

In a deployment this layer is the reference's UNCHANGED R code calling the Rcpp shim.  It is mirrored in
Python because R is not available in the build / bench environment, so that the whole path phenotype file + PLINK files ->
`<prefix>.rda` + `<prefix>.varianceRatio.txt` can be run, tested and timed end to end, and so that the files it writes can
be read by the reference's own step 2 (`load()` in readInGLMM.R:39-45; writer: rdata.save_rda).  It does no numerical work
on genotypes: everything goes through the export mirror in api.py, i.e. the C ABI.

What follows the reference, block by block:
  output-file checks, overwrite rule of the variance-ratio file      FG.R:703-760
  .fam / .bim / phenotype reading, complete cases, sample include file, sex filter, categorical covariates
  (treatment contrasts of model.matrix), merge on the sample ID in .fam order                 FG.R:812-957
  inverse-normal transform of a quantitative trait                                           FG.R:959-967
  checkPerfectSep (binary trait, sparse two-level covariates)                                FG.R:969-984, 2110-2136
  Covariate_Transform: collinear columns dropped, X = Q sqrt(N)                              FG.R:1020-1043, 1612-1650
  glm with all covariates, isCovariateOffset (covariate effects as a fixed offset)           FG.R:1046-1063
  setminMAC_VarianceRatio(20, -1, TRUE), setminMAFforGRM, setmaxMissingRateforGRM            FG.R:1080-1106
  glmmkin.ai_PCG_Rcpp_Binary / _Quantitative (step1.glmmkin_ai_PCG)                          FG.R:79-549
  modglmm$offset, per-chromosome offsets, Covariate_Transform_Back, save(modglmm)            FG.R:1159-1301
  extractVarianceRatio (autosomal markers only unless includeNonautoMarkersforVarRatio)      FG.R:2152-2423

Not provided (the reference's sparse-GRM machinery is out of scope): useSparseGRMtoFitNULL,
useSparseGRMforVarRatio.  They raise, they are never ignored.

Randomness: the reference draws the Hutchinson probes, the variance-ratio hold-out set and the marker order from R's RNG
(set.seed(1) in the CLI, set_seed(200) in GetTrace).  The probes reproduce R's stream bit for bit (step1.ProbeStream,
rng="R"); the two index draws come from numpy's generator seeded with `seed` (arma::randi / sample() under R's RNG state
are not reproducible outside an R session)."""
import gzip
import os

import numpy as np

from . import step1
from .rdata import RList, load_rda, save_rda


class SaigeInputError(ValueError):
    pass


def _read_table(path, id_col=None):
    """data.table::fread on a phenotype file: header line, separator sniffed (tab, comma or blanks), gz transparently."""
    opener = gzip.open if path.endswith((".gz", ".bgz")) else open
    with opener(path, "rt") as f:
        lines = [l.rstrip("\n").rstrip("\r") for l in f if l.strip()]
    sep = "\t" if "\t" in lines[0] else ("," if "," in lines[0] else None)
    rows = [l.split(sep) for l in lines]
    hdr = [h.strip().strip('"') for h in rows[0]]
    cols = {h: [r[i].strip().strip('"') if i < len(r) else "" for r in rows[1:]] for i, h in enumerate(hdr)}
    return hdr, cols


def _is_na(s):
    return s in ("", "NA", "NaN", "nan", "na", ".")


def _numeric(vals, name):
    try:
        return np.array([float(v) for v in vals], dtype=np.float64)
    except ValueError:
        raise SaigeInputError("ERROR! column %s of the phenoFile is not numeric; list it in qCovarCol if it is categorical" % name)


def _factor_levels(vals):
    """Levels of as.factor(): sorted numerically when every value is a number, else as strings."""
    u = sorted(set(vals))
    try:
        return sorted(u, key=float)
    except ValueError:
        return u


def _design(cols, rows, covarColList, qCovarCol):
    """model.matrix(pheno ~ covariates): intercept, numeric covariates as they are, categorical ones as treatment contrasts
    (first level is the baseline, column name = covariate name followed by the level)."""
    X, names = [np.ones(len(rows))], ["(Intercept)"]
    for c in covarColList:
        v = [cols[c][i] for i in rows]
        if c in qCovarCol:
            for lev in _factor_levels(v)[1:]:
                X.append(np.array([1.0 if x == lev else 0.0 for x in v]))
                names.append(c + lev)
        else:
            X.append(_numeric(v, c))
            names.append(c)
    return np.column_stack(X), names


def checkPerfectSep(X, names, y, minCovariateCount):
    """FG.R:2110-2136: two-valued covariates whose 2 x 2 table against the phenotype has a cell below minCovariateCount."""
    drop = []
    for j in range(1, X.shape[1]):
        u = np.unique(X[:, j])
        if len(u) == 2:
            cells = [np.sum((y == a) & (X[:, j] == b)) for a in np.unique(y) for b in u]
            if any(c < minCovariateCount for c in cells):
                drop.append(names[j])
    return drop


def Covariate_Transform(X1, names):
    """FG.R:1612-1650: columns that lm() would report as NA (linearly dependent on the ones before them) are dropped, the
    rest is replaced by Q sqrt(N) of its QR decomposition."""
    keep, R = [], np.zeros((0, 0))
    for j in range(X1.shape[1]):
        trial = X1[:, keep + [j]]
        if np.linalg.matrix_rank(trial) == len(keep) + 1:
            keep.append(j)
    idx_na = [j for j in range(X1.shape[1]) if j not in keep]
    Xk = X1[:, keep]
    Q, R = np.linalg.qr(Xk)
    sgn = np.where(np.diag(R) < 0, -1.0, 1.0)            # LINPACK's dqrdc2 (R's qr) and LAPACK differ by column signs only;
    Q, R = Q * sgn, R * sgn[:, None]                     # a positive diagonal fixes one representative
    N = X1.shape[0]
    new_names = [names[j] for j in keep]
    if 0 not in idx_na:
        new_names[0] = "minus1"
    return Q * np.sqrt(N), dict(qrr=R, N=N, X_name=new_names, idx_na=idx_na)


def _chr_number(c):
    try:
        return int(str(c).replace("chr", ""))
    except ValueError:
        return -1


def write_variance_ratio(path, ratio):
    """write.table(varRatioTable, quote = F, col.names = F, row.names = F) of FG.R:2413-2417: `<ratio> null <category>`,
    one line per MAC category."""
    with open(path, "w") as f:
        for k, r in enumerate(np.asarray(ratio, dtype=np.float64).reshape(-1)):
            f.write("%.15g null %d\n" % (float(r), k + 1))


def fitNULLGLMM(geno=None, plinkFile="", bedFile="", bimFile="", famFile="", phenoFile="", phenoCol="", traitType="binary",
                invNormalize=False, covarColList=None, qCovarCol=None, sampleIDColinphenoFile="", tol=0.02, maxiter=20,
                tolPCG=1e-5, maxiterPCG=500, nThreads=1, SPAcutoff=2, numMarkersForVarRatio=30, skipModelFitting=False,
                memoryChunk=2, tauInit=(0, 0), LOCO=True, isLowMemLOCO=False, traceCVcutoff=0.0025, ratioCVcutoff=0.001,
                outputPrefix="", outputPrefix_varRatio="", IsOverwriteVarianceRatioFile=False, sparseGRMFile="",
                sparseGRMSampleIDFile="", numRandomMarkerforSparseKin=1000, relatednessCutoff=0.125,
                isCateVarianceRatio=False, cateVarRatioIndexVec=None, cateVarRatioMinMACVecExclude=(10, 20.5),
                cateVarRatioMaxMACVecInclude=(20.5,), isCovariateTransform=True, isDiagofKinSetAsOne=False,
                minCovariateCount=-1, minMAFforGRM=0.01, maxMissingRateforGRM=0.15, useSparseGRMtoFitNULL=False,
                useSparseGRMforVarRatio=False, includeNonautoMarkersforVarRatio=False, sexCol="", FemaleCode=1,
                FemaleOnly=False, MaleCode=0, MaleOnly=False, SampleIDIncludeFile="", isCovariateOffset=False,
                skipVarianceRatioEstimation=False, nrun=30, probe_rng="R", seed=1, verbose=False):
    """Returns dict(modglmm=..., varianceRatio=..., modelFile=..., varRatioFile=...); writes the two files."""
    covarColList = list(covarColList or [])
    qCovarCol = list(qCovarCol or [])
    say = print if verbose else (lambda *a, **k: None)
    if useSparseGRMtoFitNULL or useSparseGRMforVarRatio:
        raise NotImplementedError("sparse-GRM fitting / variance ratios are not provided by the B200 back end (full-GRM path only)")
    if nThreads > 1:
        raise SaigeInputError("setting threads via RcppParallel is not allowed")          # FG.R:785
    if traitType not in ("binary", "quantitative"):
        raise SaigeInputError("traitType must be 'binary' or 'quantitative'")
    if FemaleOnly and MaleOnly:
        raise SaigeInputError("Both FemaleOnly and MaleOnly are TRUE. Please specify only one of them as TRUE to run the sex-specific job")
    # ---- output files (FG.R:703-760, 790-800) ----
    if FemaleOnly:
        outputPrefix += "_FemaleOnly"
    elif MaleOnly:
        outputPrefix += "_MaleOnly"
    modelOut = outputPrefix + ("_noLOCO.rda" if (LOCO and isLowMemLOCO and not skipModelFitting) else ".rda")     # FG.R:711-714
    if skipModelFitting and not os.path.exists(modelOut):
        raise SaigeInputError("skipModelFitting=TRUE but %s does not exist" % modelOut)
    varRatioFile = None
    if not skipVarianceRatioEstimation:
        varRatioFile = (outputPrefix_varRatio or outputPrefix) + ".varianceRatio.txt"
        if os.path.exists(varRatioFile) and os.path.getsize(varRatioFile) > 0 and not IsOverwriteVarianceRatioFile:
            raise SaigeInputError("WARNING: The variance ratio file %s already exists. Remove it, give outputPrefix_varRatio, or "
                                  "specify IsOverwriteVarianceRatioFile=TRUE" % varRatioFile)
    if plinkFile:
        bedFile, bimFile, famFile = plinkFile + ".bed", plinkFile + ".bim", plinkFile + ".fam"
    for f, what in ((bedFile, "bed"), (bimFile, "bim"), (famFile, "fam")):
        if not os.path.exists(f):
            raise SaigeInputError("ERROR! %s file does not exsit" % what)
    if not os.path.exists(phenoFile):
        raise SaigeInputError("ERROR! phenoFile %s does not exsit" % phenoFile)
    bim_chr = [_chr_number(l.split()[0]) for l in open(bimFile) if l.strip()]
    fam_iid = [l.split()[1] for l in open(famFile) if l.strip()]
    say(len(fam_iid), " samples have genotypes")
    # ---- phenotype file (FG.R:858-957) ----
    hdr, cols = _read_table(phenoFile)
    for c in [phenoCol] + covarColList + [sampleIDColinphenoFile]:
        if c not in cols:
            raise SaigeInputError("ERROR! column for %s does not exist in the phenoFile" % c)
    if not all(q in covarColList for q in qCovarCol):
        raise SaigeInputError("ERROR! all covariates in qCovarCol must be in covarColList")
    need = [phenoCol] + covarColList + [sampleIDColinphenoFile] + ([sexCol] if (FemaleOnly or MaleOnly) else [])
    if (FemaleOnly or MaleOnly) and sexCol not in cols:
        raise SaigeInputError("ERROR! column for sex %s does not exist in the phenoFile" % sexCol)
    nrow = len(cols[phenoCol])
    rows = [i for i in range(nrow) if not any(_is_na(cols[c][i]) for c in need)]
    if SampleIDIncludeFile:
        if not os.path.exists(SampleIDIncludeFile):
            raise SaigeInputError("ERROR! SampleIDIncludeFile %s does not exsit" % SampleIDIncludeFile)
        inc = {l.split()[0] for l in open(SampleIDIncludeFile) if l.strip()}
        rows = [i for i in rows if cols[sampleIDColinphenoFile][i] in inc]
    if FemaleOnly or MaleOnly:
        code = float(FemaleCode if FemaleOnly else MaleCode)
        rows = [i for i in rows if float(cols[sexCol][i]) == code]
        if not rows:
            raise SaigeInputError("ERROR! no samples in the phenotype are coded as %s in the column %s" % (code, sexCol))
    say(len(rows), " samples have non-missing phenotypes")
    # merge with the .fam on the sample ID, genotype order (FG.R:939-957)
    where = {}
    for i in rows:
        where.setdefault(cols[sampleIDColinphenoFile][i], i)
    geno_index = [j for j, iid in enumerate(fam_iid) if iid in where]
    rows = [where[fam_iid[j]] for j in geno_index]
    indicator = np.zeros(len(fam_iid), dtype=np.uint8)
    indicator[geno_index] = 1
    subSampleInGeno = np.array(geno_index, dtype=np.int32) + 1          # IndexGeno, 1-based
    sampleID = [fam_iid[j] for j in geno_index]
    N = len(rows)
    if N == 0:
        raise SaigeInputError("no sample of the phenotype file is in the .fam")
    say(N, " samples will be used for analysis")
    y = _numeric([cols[phenoCol][i] for i in rows], phenoCol)
    if traitType == "quantitative" and invNormalize:
        from scipy import stats
        y = stats.norm.ppf((stats.rankdata(y, method="average") - 0.5) / N)
    X, xnames = _design(cols, rows, covarColList, qCovarCol)
    hasCovariate = len(covarColList) > 0
    if traitType == "binary":
        u = np.unique(y)
        if len(u) != 2 or u[0] != 0 or u[1] != 1:
            raise SaigeInputError("ERROR! phenotype value needs to be 0 or 1")
        if hasCovariate:
            drop = checkPerfectSep(X, xnames, y, minCovariateCount)
            if drop:
                keep = [j for j, nme in enumerate(xnames) if nme not in drop]
                X, xnames = X[:, keep], [xnames[j] for j in keep]
            hasCovariate = X.shape[1] > 1
    if not hasCovariate:
        isCovariateOffset = False
    out_transform = None
    if isCovariateTransform and hasCovariate:
        X, out_transform = Covariate_Transform(X, xnames)
        xnames = out_transform["X_name"]
    family = step1.Binomial if traitType == "binary" else step1.Gaussian
    modwitcov = step1.glm_fit(y, X, family)
    Xorig = None
    if isCovariateOffset:
        covoffset = X[:, 1:] @ modwitcov["coef"][1:]
        Xorig = X
        fit0 = step1.glm_fit(y, np.ones((N, 1)), family, offset=covoffset)
        hasCovariate = False
    else:
        covoffset = np.zeros(N)
        fit0 = modwitcov
    fit0_has_offset = isCovariateOffset
    # ---- genotype store (FG.R:1080-1106; glmmkin.ai_PCG_Rcpp_*: setgeno, LOCO ranges over the QC-passing markers) ----
    own_geno = geno is None
    if own_geno:
        from .api import SaigeB200
        geno = SaigeB200(device=0)
    # marker-sharded multi-GPU runs (one process per GPU, the same call on every rank): all ranks compute, rank 0 writes --
    # the reference's `if (comm.rank(comm = 0) == 0) save(...)` (FG.R:1297-1301)
    writer = getattr(geno, "rank", 0) == 0
    save = save_rda if writer else (lambda *a, **k: None)
    rng = np.random.default_rng(seed)
    vr_idx = None
    if not skipVarianceRatioEstimation:
        if isCateVarianceRatio:       # FG.R:1082-1086: every marker inside the MAC categories is held out, random ones above
            geno.setminMAC_VarianceRatio(min(cateVarRatioMinMACVecExclude), max(cateVarRatioMaxMACVecInclude), True)
        else:
            geno.setminMAC_VarianceRatio(20, -1, True)
        vr_idx = np.unique(rng.integers(0, len(bim_chr), size=1000)).astype(np.int32)       # FG.cpp:866-868: 1000 draws, unique
    geno.setminMAFforGRM(minMAFforGRM)
    geno.setmaxMissingRateforGRM(maxMissingRateforGRM)
    geno.setgeno(bedFile, bimFile, famFile, subSampleInGeno, indicator, memoryChunk, isDiagofKinSetAsOne, vr_rand_idx=vr_idx)
    qc = np.asarray(geno.getQCdMarkerIndex()).astype(bool)
    if LOCO:
        LOCO = step1.set_loco_ranges(geno, np.asarray(bim_chr)[qc])
    # ---- the fit ----
    if not skipModelFitting:
        probes = step1.ProbeStream(N, nmax=nrun + 100, seed=200, rng=probe_rng)
        m = step1.glmmkin_ai_PCG(geno, fit0, probes, trait=traitType, tauInit=tauInit, maxiter=maxiter, tol=tol, nrun=nrun,
                                 tolPCG=tolPCG, maxiterPCG=maxiterPCG, traceCVcutoff=traceCVcutoff,
                                 LOCO=LOCO and not isLowMemLOCO, verbose=verbose)
        Xfit = np.asarray(fit0["X"])
        Xout = Xorig if isCovariateOffset else Xfit
        tau = np.asarray(m["theta"], dtype=np.float64)

        def noK(mu):
            mu2 = mu * (1 - mu) if traitType == "binary" else np.full(N, 1.0 / tau[0])
            o = step1.ScoreTest_NULL_Model(mu, mu2, y, Xout)
            return RList([(k, np.ascontiguousarray(o[k])) for k in ("XV", "XVX", "XXVX_inv", "XVX_inv", "S_a", "XVX_inv_XV", "V")],
                         r_class=["SA_NULL"])

        def back(alpha):
            if out_transform is not None and not fit0_has_offset:
                return step1.Covariate_Transform_Back(alpha, out_transform)
            return alpha

        col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1, 1)
        alpha0 = np.asarray(m["coefficients"], dtype=np.float64)
        if isCovariateOffset or not hasCovariate:
            offset = covoffset
        else:
            offset = Xfit[:, 1:] @ alpha0[1:]
        modglmm = dict([
            ("theta", tau), ("coefficients", col(back(alpha0))), ("linear.predictors", col(m["linear_predictors"])),
            ("fitted.values", col(m["fitted_values"])), ("Y", col(m["Y"])), ("residuals", col(m["residuals"])),
            ("cov", np.asarray(m["cov"])), ("converged", bool(m["converged"])), ("sampleID", list(sampleID)),
            ("obj.noK", noK(np.asarray(m["fitted_values"]))), ("y", y), ("X", Xout), ("traitType", traitType),
            ("isCovariateOffset", bool(isCovariateOffset)), ("LOCO", bool(LOCO and not isLowMemLOCO))])
        if LOCO and not isLowMemLOCO:
            lres = []
            for e in m["LOCOResult"]:
                if not e.get("isLOCO"):
                    lres.append(dict(isLOCO=False))
                    continue
                a = np.asarray(e["coefficients"], dtype=np.float64)
                d = dict([("isLOCO", True), ("coefficients", col(back(a))), ("linear.predictors", col(e["linear_predictors"])),
                          ("fitted.values", col(e["fitted_values"])), ("Y", col(e["Y"])), ("residuals", col(e["residuals"])),
                          ("cov", np.asarray(e["cov"])), ("obj.noK", noK(np.asarray(e["fitted_values"]))), ("alpha0", col(a))])
                if not isCovariateOffset and hasCovariate:
                    d["offset"] = col(Xfit[:, 1:] @ a[1:])
                lres.append(d)
            modglmm["LOCOResult"] = lres
        modglmm["offset"] = col(offset)
        modglmm["useSparseGRMtoFitNULL"] = False
        save(modelOut, {"modglmm": modglmm})
        if LOCO and isLowMemLOCO:
            # FG.R:1205-1290: the model without LOCO is on disk; every chromosome is refitted FROM THE MAIN FIT's alpha / eta
            # (not from the previous chromosome's) and saved to its own <prefix>_chr<j>.rda, which holds that chromosome only
            slim = dict(modglmm)
            slim["LOCO"] = True
            for k in ("Y", "linear.predictors", "coefficients", "cov", "fitted.values", "residuals", "obj.noK", "offset"):
                del slim[k]                                   # `modglmm$Y = NULL` removes the element
            geno.set_Diagof_StdGeno_LOCO()
            eta0 = np.asarray(m["linear_predictors"], dtype=np.float64)
            off0 = np.asarray(fit0["offset"], dtype=np.float64)
            state = []
            for j, (s0, e0) in enumerate(zip(step1.geno_start_vec(geno), step1.geno_end_vec(geno))):
                if s0 == -1 or e0 == -1:
                    state.append(dict(isLOCO=False))
                    continue
                geno.setStartEndIndex(s0, e0, j)
                rl = step1.Get_Coef(geno, y, np.asfortranarray(Xfit), tau, family, alpha0, eta0, off0, maxiterPCG, tolPCG, maxiter, loco=True)
                a = np.asarray(rl["alpha"], dtype=np.float64)
                d = dict([("isLOCO", True), ("coefficients", col(back(a))), ("linear.predictors", col(rl["eta"])),
                          ("fitted.values", col(rl["mu"])), ("Y", col(rl["Y"])), ("residuals", col(y - rl["mu"])),
                          ("cov", np.asarray(rl["cov"])), ("obj.noK", noK(np.asarray(rl["mu"])))])
                if not isCovariateOffset and hasCovariate:
                    d["offset"] = col(Xfit[:, 1:] @ a[1:])
                slim["LOCOResult"] = state + [d] + [[None]] * (21 - j)
                save("%s_chr%d.rda" % (outputPrefix, j + 1), {"modglmm": slim})
                state.append([None])
    else:
        modglmm = load_rda(modelOut)["modglmm"]
        if modglmm.get("LOCO") is None:
            modglmm["LOCO"] = False
        if LOCO:
            geno.set_Diagof_StdGeno_LOCO()
    # ---- variance ratio (FG.R:1325-1352, 2152-2423) ----
    ratio = None
    if not skipVarianceRatioEstimation:
        model = dict(fitted_values=np.asarray(modglmm["fitted.values"], dtype=np.float64).ravel(),
                     linear_predictors=np.asarray(modglmm["linear.predictors"], dtype=np.float64).ravel(),
                     y=np.asarray(modglmm["y"], dtype=np.float64).ravel(), X=np.asarray(modglmm["X"], dtype=np.float64),
                     theta=np.asarray(modglmm["theta"], dtype=np.float64).ravel(),
                     obj_noK={k: np.asarray(v) for k, v in modglmm["obj.noK"].items()},
                     traitType=modglmm["traitType"][0] if isinstance(modglmm["traitType"], list) else modglmm["traitType"])
        use_vr = geno.getIsVarRatioGeno() and geno.Mvr > 0
        if geno.getIsVarRatioGeno() and geno.Mvr == 0:
            raise SaigeInputError("No markers were found for variance ratio estimation")        # FG.R:2225
        if use_vr:
            chr_of = np.asarray(bim_chr)[np.asarray(geno.getIndexVec_forVarRatio())]
            n_avail = geno.Mvr
        else:
            chr_of = np.asarray(bim_chr)[qc]
            n_avail = geno.M
        auto = np.ones(n_avail, dtype=bool) if includeNonautoMarkersforVarRatio else (chr_of >= 1) & (chr_of <= 22)   # FG.R:2286
        if isCateVarianceRatio:
            mac = np.asarray(geno.getMACVec_forVarRatio() if use_vr else geno.getMACVec())
            per_cat = step1.extractVarianceRatio_cate(geno, model, family, mac, auto, rng, cateVarRatioMinMACVecExclude,
                                                      cateVarRatioMaxMACVecInclude, cateVarRatioIndexVec,
                                                      numMarkers=numMarkersForVarRatio, maxiterPCG=maxiterPCG, tolPCG=tolPCG,
                                                      ratioCVcutoff=ratioCVcutoff)
            ratio = [r for r, _ in per_cat]
        else:
            order = rng.permutation(n_avail)                                                   # sample(MACindex), FG.R:2233
            ratio, ratios = step1.extractVarianceRatio(geno, model, family, order[auto[order]], numMarkers=numMarkersForVarRatio,
                                                       maxiterPCG=maxiterPCG, tolPCG=tolPCG, ratioCVcutoff=ratioCVcutoff)
        if writer:
            write_variance_ratio(varRatioFile, ratio)
        say("varRatio_null", ratio)
    if own_geno:
        geno.closeGenoFile_plink()                    # FG.R:1352; a handle passed in stays open for the caller (e.g. step 2)
    return dict(modglmm=modglmm, varianceRatio=ratio, modelFile=modelOut, varRatioFile=varRatioFile)
