"""fitNULLGLMM (host mirror of FG.R:650-1610) and the R save() writer: phenotype file + PLINK files -> .rda + varianceRatio.txt.

CPU tests drive the mirror with a test-only adapter that answers the export calls from the CPU oracle (so the host logic,
the file formats and the hand-over to step 2 are covered without a GPU); the GPU test runs the same call through the C ABI
and compares the two results at the north_star tolerance (1e-6 on tau / coefficients / variance ratio)."""
import gzip
import os

import numpy as np
import pytest

TOL_FIT = 1e-6


class OracleBackend:
    """TEST INFRASTRUCTURE: the export surface of saige_gpu_b200.api.SaigeB200 that step1 / fitnull call, answered by
    oracle/oracle.py.  Lives in tests/ only; the product never sees it."""

    def __init__(self):
        from oracle import oracle as O
        self.O, self.o = O, O.OracleGeno()
        self._vr = False

    def setminMAFforGRM(self, v):
        self.o.minMAF = v

    def setmaxMissingRateforGRM(self, v):
        self.o.maxMissing = v

    def setminMAC_VarianceRatio(self, lo, hi, flag):
        self.o.isVarRatio, self.o.minMACvr, self.o.maxMACvr = bool(flag), lo, hi

    def setgeno(self, bed, bim, fam, sub, ind, memoryChunk=2.0, isDiagofKinSetAsOne=False, vr_rand_idx=None):
        body, N0, M0, _ = self.O.read_bed(bed[:-4])
        self.o.setgeno(body, N0, M0, sub, ind, isDiagofKinSetAsOne, vr_rand_idx=vr_rand_idx)
        self.N, self.M, self.Mvr = self.o.N, self.o.M, self.o.Mvr

    def getQCdMarkerIndex(self):
        return self.o.qc_mask

    def getIsVarRatioGeno(self):
        return bool(self.o.isVarRatio)

    def getMACVec(self):
        return self.o.MACVec

    def getMACVec_forVarRatio(self):
        return self.o.MACVec_forVarRatio

    def getIndexVec_forVarRatio(self):
        return self.o.markerIndexVec_forVarRatio

    def Get_OneSNP_Geno(self, i):
        return self.o.Get_OneSNP_Geno(i)

    def Get_OneSNP_Geno_forVarRatio(self, i):
        return self.o.Get_OneSNP_Geno(i, vr=True)

    def setStartEndIndexVec(self, s, e):
        self.o.setStartEndIndexVec(s, e)

    def setStartEndIndex(self, s, e, c):
        self.o.setStartEndIndex(s, e, c)

    def set_Diagof_StdGeno_LOCO(self):
        self.o.set_Diagof_StdGeno_LOCO()

    def getCoefficients(self, Y, X, W, tau, maxiterPCG, tolPCG, loco=False):
        return self.O.getCoefficients(self.o, Y, X, W, tau, maxiterPCG, tolPCG, loco)

    def getAIScore(self, *a):
        return self.O.getAIScore(self.o, *a)

    def getAIScore_q(self, *a):
        return self.O.getAIScore_q(self.o, *a)

    def fitglmmaiRPCG(self, *a):
        return dict(tau=self.O.fitglmmaiRPCG(self.o, *a))

    def fitglmmaiRPCG_q(self, *a):
        return dict(tau=self.O.fitglmmaiRPCG_q(self.o, *a))

    def getSigma_X(self, W, tau, X, maxiterPCG, tolPCG):
        return self.O.getSigma_X(self.o, W, tau, X, maxiterPCG, tolPCG)

    def getSigma_G(self, W, tau, G, maxiterPCG, tolPCG):
        G = np.asarray(G)
        if G.ndim == 1:
            return self.O.getSigma_G(self.o, W, tau, G, maxiterPCG, tolPCG)
        return np.column_stack([self.O.getSigma_G(self.o, W, tau, G[:, j], maxiterPCG, tolPCG) for j in range(G.shape[1])])

    def calCV(self, x):
        return self.O.calCV(x)

    def closeGenoFile_plink(self):
        pass


# ---- helpers of the mirror -------------------------------------------------------------------------------------------------
def test_design_matrix_and_covariate_transform():
    from saige_gpu_b200 import fitnull
    cols = {"y": ["1", "0", "1", "0", "1", "0"], "age": ["30", "41", "52", "63", "74", "35"], "site": ["b", "a", "c", "a", "b", "c"],
            "batch": ["10", "2", "2", "10", "1", "1"]}
    X, names = fitnull._design(cols, [0, 1, 2, 3, 4, 5], ["age", "site", "batch"], ["site", "batch"])
    assert names == ["(Intercept)", "age", "siteb", "sitec", "batch2", "batch10"]          # levels: a < b < c; 1 < 2 < 10 (numeric order)
    assert X[:, 2].tolist() == [1, 0, 0, 0, 1, 0] and X[:, 5].tolist() == [1, 0, 0, 1, 0, 0]
    # a collinear column (copy of age, scaled) is dropped like lm()'s NA coefficient; X'X = N I afterwards
    Xc = np.column_stack([X[:, :2], 2 * X[:, 1], X[:, 2]])
    Xt, par = fitnull.Covariate_Transform(Xc, ["(Intercept)", "age", "age2", "siteb"])
    assert par["idx_na"] == [2] and par["X_name"] == ["minus1", "age", "siteb"]
    assert np.allclose(Xt.T @ Xt, 6 * np.eye(3), atol=1e-10)
    from saige_gpu_b200 import step1
    beta = np.array([0.3, -0.02, 0.5])
    coef_t = par["qrr"] @ beta / np.sqrt(6)                        # coefficients on the transformed scale
    assert np.allclose(step1.Covariate_Transform_Back(coef_t, par), beta)
    # checkPerfectSep: a two-level covariate with an empty phenotype x level cell goes
    y = np.array([1, 0, 1, 0, 1, 0.0])
    sparse = np.array([1, 0, 1, 0, 1, 0.0])
    assert fitnull.checkPerfectSep(np.column_stack([np.ones(6), sparse, X[:, 1]]), ["i", "s", "age"], y, 1) == ["s"]
    assert fitnull.checkPerfectSep(np.column_stack([np.ones(6), sparse, X[:, 1]]), ["i", "s", "age"], y, -1) == []


def test_rda_writer_round_trip_and_r_bytes(golden_dir, tmp_path):
    """save_rda -> load_rda is the identity on every type a model file holds, and re-serialising the reference's own model
    reproduces R's bytes up to the first attribute the reader does not keep (dimnames of XV): the pairlist / symbol
    reference / vector encodings are therefore exactly R's."""
    from saige_gpu_b200.rdata import RList, load_rda, save_rda
    obj = dict(theta=np.array([1.0, 0.25]), n=np.array([3], dtype=np.int32), flag=True, name="binary", ids=["a", "b", None],
               M=np.arange(6.0).reshape(2, 3), nested=RList([("XV", np.ones((2, 2))), ("S_a", np.array([1.5]))], r_class=["SA_NULL"]),
               lst=[dict(isLOCO=False), dict(isLOCO=True, v=np.array([np.nan, np.inf]))], nothing=None,
               lg=np.array([1, 0, -1], dtype=np.int8))
    p = str(tmp_path / "t.rda")
    save_rda(p, {"modglmm": obj, "second": np.array([7.0])})
    back = load_rda(p)
    assert list(back.keys()) == ["modglmm", "second"]
    b = back["modglmm"]
    assert list(b.keys()) == list(obj.keys())
    assert np.array_equal(b["theta"], obj["theta"]) and b["n"].dtype == np.int32 and b["flag"].tolist() == [1]
    assert b["name"] == ["binary"] and b["ids"] == ["a", "b", None] and np.array_equal(b["M"], obj["M"])
    assert isinstance(b["nested"], RList) and b["nested"].r_class == ["SA_NULL"] and b["nested"]["XV"].shape == (2, 2)
    assert b["lst"][0]["isLOCO"].tolist() == [0] and np.isnan(b["lst"][1]["v"][0]) and np.isinf(b["lst"][1]["v"][1])
    assert b["nothing"] is None and b["lg"].tolist() == [1, 0, -1]
    # R's own bytes
    src = os.path.join(golden_dir, "example_binary.rda")
    a = gzip.decompress(open(src, "rb").read())
    assert a[:5] == b"RDX3\n"
    import struct
    ha = 23 + struct.unpack(">i", a[19:23])[0]                      # RDX3 header: 7 + 3 ints + native-encoding string
    save_rda(p, load_rda(src))
    mine = gzip.decompress(open(p, "rb").read())[19:]               # RDX2 header: 7 + 3 ints
    same = next((i for i, (x, y) in enumerate(zip(a[ha:], mine)) if x != y), None)
    assert same is not None and same > 60000                        # theta ... sampleID, obj.noK's header and XV's data
    assert a[ha + same - 200:ha + same + 40].find(b"dimnames") >= 0


# ---- the whole entry point, oracle-backed (CPU) ----------------------------------------------------------------------------
@pytest.fixture(scope="module")
def bim22(golden_dir, tmp_path_factory):
    """The bundled 10k-marker set with its markers dealt to 22 chromosomes in contiguous blocks (its own .bim has 9,988
    markers on chromosome 1), so that LOCO has something to leave out; .bed / .fam are the committed fixtures."""
    p = str(tmp_path_factory.mktemp("bim22") / "grm10k_22chr.bim")
    rows = [l.split() for l in open(os.path.join(golden_dir, "grm10k.bim"))]
    with open(p, "w") as f:
        for i, r in enumerate(rows):
            f.write("\t".join([str(1 + (i * 22) // len(rows))] + r[1:]) + "\n")
    return p


def _run(geno, golden_dir, bim, out, **kw):
    from saige_gpu_b200 import fitnull
    args = dict(bedFile=os.path.join(golden_dir, "grm10k.bed"), bimFile=bim, famFile=os.path.join(golden_dir, "grm10k.fam"),
                phenoFile=os.path.join(golden_dir, "pheno_1000samples.txt"),
                phenoCol="y_binary", covarColList=["x1", "x2"], sampleIDColinphenoFile="IID", traitType="binary",
                outputPrefix=out, nrun=30, LOCO=True, minMAFforGRM=0.01, probe_rng="numpy", IsOverwriteVarianceRatioFile=True)
    args.update(kw)
    return fitnull.fitNULLGLMM(geno, **args)


@pytest.fixture(scope="module")
def oracle_fit(golden_dir, bim22, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fitnull") / "oracle_binary")
    return _run(OracleBackend(), golden_dir, bim22, out), out


def test_fitnullglmm_writes_the_reference_files(oracle_fit, golden_dir):
    from saige_gpu_b200.rdata import RList, load_rda
    r, out = oracle_fit
    assert os.path.exists(out + ".rda") and os.path.exists(out + ".varianceRatio.txt")
    m = load_rda(out + ".rda")["modglmm"]
    ref = load_rda(os.path.join(golden_dir, "example_binary.rda"))["modglmm"]
    # same fields, in the reference's order (obj.glm.null is R's glm object: not produced, step 2 does not read it)
    assert list(m.keys()) == [k for k in ref.keys() if k != "obj.glm.null"]
    N, p = 1000, 3
    for k, shape in (("theta", (2,)), ("coefficients", (p, 1)), ("linear.predictors", (N, 1)), ("fitted.values", (N, 1)),
                     ("Y", (N, 1)), ("residuals", (N, 1)), ("cov", (p, p)), ("y", (N,)), ("X", (N, p)), ("offset", (N, 1))):
        assert m[k].shape == shape and m[k].dtype == ref[k].dtype, k
    assert m["traitType"] == ["binary"] and m["LOCO"].tolist() == [1] and m["converged"].tolist() == [1]
    assert m["sampleID"][:3] == ref["sampleID"][:3] and len(m["sampleID"]) == N
    assert isinstance(m["obj.noK"], RList) and m["obj.noK"].r_class == ["SA_NULL"]
    assert list(m["obj.noK"].keys()) == list(ref["obj.noK"].keys())
    for k in ref["obj.noK"]:
        assert m["obj.noK"][k].shape == ref["obj.noK"][k].shape, k
    assert len(m["LOCOResult"]) == 22
    lr = m["LOCOResult"][0]
    assert lr["isLOCO"].tolist() == [1] and set(ref["LOCOResult"][0].keys()) <= set(lr.keys()) and lr["offset"].shape == (N, 1)
    # the numbers: tau[0] fixed at 1 for a binary trait, coefficients on the ORIGINAL covariate scale, mu in (0, 1)
    assert m["theta"][0] == 1.0 and m["theta"][1] > 0
    mu = m["fitted.values"].ravel()
    assert np.all((mu > 0) & (mu < 1)) and np.allclose(m["residuals"].ravel(), m["y"] - mu)
    assert np.allclose(m["X"].T @ m["X"], N * np.eye(p), atol=1e-8)               # covariates are stored transformed
    line = open(out + ".varianceRatio.txt").read().split()
    assert line[1:] == ["null", "1"] and abs(float(line[0]) - r["varianceRatio"]) < 1e-14 * r["varianceRatio"]
    assert 0.5 < r["varianceRatio"] < 1.5


def test_covariate_transform_does_not_change_the_model(oracle_fit, golden_dir, bim22, tmp_path):
    """Coefficients are reported on the original scale (Covariate_Transform_Back): fitting without the QR transform gives
    the same tau, coefficients and fitted values up to the PCG tolerance."""
    r, _ = oracle_fit
    r2 = _run(OracleBackend(), golden_dir, bim22, str(tmp_path / "noqr"), isCovariateTransform=False)
    a, b = r["modglmm"], r2["modglmm"]
    assert abs(a["theta"][1] - b["theta"][1]) < 2e-3 * b["theta"][1]
    assert np.allclose(a["coefficients"], b["coefficients"], rtol=2e-3)
    assert np.allclose(a["fitted.values"], b["fitted.values"], rtol=2e-3)
    assert abs(r2["varianceRatio"] - r["varianceRatio"]) < 2e-3 * r["varianceRatio"]
    # without the variance-ratio step no marker is held out of the GRM and no ratio file is written
    r3 = _run(OracleBackend(), golden_dir, bim22, str(tmp_path / "novr"), skipVarianceRatioEstimation=True, LOCO=False)
    assert r3["varianceRatio"] is None and r3["varRatioFile"] is None and not os.path.exists(str(tmp_path / "novr.varianceRatio.txt"))
    assert r3["modglmm"]["LOCO"] is False and abs(r3["modglmm"]["theta"][1] - 0.3137) < 1e-3      # the value DESIGN.md quotes for this set


def test_skip_model_fitting_reuses_the_written_model(oracle_fit, golden_dir, bim22):
    """skipModelFitting = TRUE (FG.R:1303-1313): the .rda on disk is loaded and only the variance ratio is estimated again;
    same hold-out set and marker order (same seed) -> the same ratio."""
    r, out = oracle_fit
    before = open(out + ".rda", "rb").read()
    r2 = _run(OracleBackend(), golden_dir, bim22, out, skipModelFitting=True)
    assert open(out + ".rda", "rb").read() == before                       # the model file is not rewritten
    assert abs(r2["varianceRatio"] - r["varianceRatio"]) < 1e-12 * r["varianceRatio"]        # (the oracle's OpenMP sums are not ordered)
    assert np.array_equal(r2["modglmm"]["theta"], r["modglmm"]["theta"])


def test_step2_consumes_the_written_model(oracle_fit, golden_dir):
    """The hand-over: the oracle's step 2 reads the .rda / varianceRatio.txt this run wrote (ReadModel's fields, LOCO swap)."""
    from oracle import oracle as O
    from oracle import step2_oracle as S2
    from saige_gpu_b200.rdata import load_rda
    from saige_gpu_b200.step2 import Get_Variance_Ratio
    _, out = oracle_fit
    M = S2.read_model(load_rda(out + ".rda")["modglmm"], chrom=1, LOCO=True)
    M["varRatio"] = Get_Variance_Ratio(out + ".varianceRatio.txt")
    bed, N0, M0, _ = O.read_bed(os.path.join(golden_dir, "step2_100markers"))
    fam = [l.split()[1] for l in open(os.path.join(golden_dir, "step2_100markers.fam"))]
    pos = np.array([fam.index(s) for s in M["sampleID"]])
    n = 0
    for mk in range(M0):
        r = S2.test_marker(M, S2.plink_marker(bed, N0, mk, pos), min_mac=20)
        if r is not None:
            n += 1
            assert 0 < r["p_value"] <= 1 and np.isfinite(r["BETA"]) and r["SE"] > 0
    assert n == 32                                                    # the reference's golden table has 32 rows at minMAC 20


def test_options_and_refusals(golden_dir, tmp_path):
    from saige_gpu_b200 import fitnull
    base = dict(plinkFile=os.path.join(golden_dir, "chr22_1000"), phenoFile=os.path.join(golden_dir, "pheno_1000samples.txt"),
                phenoCol="y_binary", covarColList=["x1", "x2"], sampleIDColinphenoFile="IID", outputPrefix=str(tmp_path / "x"))
    for bad in (dict(useSparseGRMtoFitNULL=True), dict(useSparseGRMforVarRatio=True)):
        with pytest.raises(NotImplementedError):
            fitnull.fitNULLGLMM(OracleBackend(), **{**base, **bad})
    with pytest.raises(fitnull.SaigeInputError):
        fitnull.fitNULLGLMM(OracleBackend(), **{**base, "phenoCol": "nope"})
    with pytest.raises(fitnull.SaigeInputError):
        fitnull.fitNULLGLMM(OracleBackend(), **{**base, "nThreads": 4})
    with pytest.raises(fitnull.SaigeInputError):
        fitnull.fitNULLGLMM(OracleBackend(), **{**base, "skipModelFitting": True})       # no .rda yet
    open(str(tmp_path / "x.varianceRatio.txt"), "w").write("0.9 null 1\n")
    with pytest.raises(fitnull.SaigeInputError):
        fitnull.fitNULLGLMM(OracleBackend(), **base)                                   # would overwrite the ratio file
    # a rank other than 0 of a marker-sharded run computes but writes nothing (FG.R:1297-1301: rank 0 saves)
    be = OracleBackend()
    be.rank = 1
    r1 = fitnull.fitNULLGLMM(be, **{**base, "outputPrefix": str(tmp_path / "rank1"), "LOCO": False, "probe_rng": "numpy",
                                    "bedFile": os.path.join(golden_dir, "grm10k.bed"), "bimFile": os.path.join(golden_dir, "grm10k.bim"),
                                    "famFile": os.path.join(golden_dir, "grm10k.fam"), "plinkFile": ""})
    assert r1["varianceRatio"] > 0 and not os.path.exists(str(tmp_path / "rank1.rda")) and not os.path.exists(str(tmp_path / "rank1.varianceRatio.txt"))
    # quantitative trait with inverse normalisation, covariates as offset, a sample include file, no LOCO
    inc = str(tmp_path / "inc.txt")
    ids = [l.split()[4] for l in open(os.path.join(golden_dir, "pheno_1000samples.txt"))][1:]
    open(inc, "w").write("\n".join(ids[:600]) + "\n")
    r = fitnull.fitNULLGLMM(OracleBackend(), **{**base, "phenoCol": "y_quantitative", "traitType": "quantitative", "invNormalize": True,
                                                "isCovariateOffset": True, "SampleIDIncludeFile": inc, "LOCO": False,
                                                "outputPrefix": str(tmp_path / "q"), "probe_rng": "numpy"})
    m = r["modglmm"]
    assert len(m["sampleID"]) == 600 and m["X"].shape == (600, 3) and m["coefficients"].shape == (1, 1)
    assert m["isCovariateOffset"] is True and m["LOCO"] is False and "LOCOResult" not in m
    assert abs(np.mean(m["y"])) < 1e-12 and abs(np.std(m["y"]) - 1) < 0.01              # rank-based inverse normal scores
    assert m["theta"][0] > 0 and r["varianceRatio"] > 0


def test_low_memory_loco_writes_one_model_per_chromosome(oracle_fit, golden_dir, bim22, tmp_path):
    """isLowMemLOCO (FG.R:1205-1290): <prefix>_noLOCO.rda without LOCO results + <prefix>_chr<j>.rda holding chromosome j's
    refit only; step 2 reads the chromosome's file.  Same model as the in-memory LOCO run up to the IRLS stopping rule
    (each chromosome restarts from the main fit instead of from the previous chromosome)."""
    from saige_gpu_b200 import step2
    from saige_gpu_b200.rdata import load_rda
    ref, _ = oracle_fit
    out = str(tmp_path / "lowmem")
    r = _run(OracleBackend(), golden_dir, bim22, out, isLowMemLOCO=True)
    assert r["modelFile"] == out + "_noLOCO.rda" and not os.path.exists(out + ".rda")
    main = load_rda(out + "_noLOCO.rda")["modglmm"]
    assert main["LOCO"].tolist() == [0] and "LOCOResult" not in main
    assert np.allclose(main["theta"], ref["modglmm"]["theta"], rtol=1e-12)
    assert abs(r["varianceRatio"] - ref["varianceRatio"]) < 1e-9 * ref["varianceRatio"]
    for j in (1, 7, 22):
        m = load_rda("%s_chr%d.rda" % (out, j))["modglmm"]
        assert m["LOCO"].tolist() == [1] and "fitted.values" not in m and "obj.noK" not in m and len(m["LOCOResult"]) == 22
        assert [isinstance(x, dict) and "fitted.values" in x for x in m["LOCOResult"]] == [k == j - 1 for k in range(22)]
        a, b = m["LOCOResult"][j - 1], ref["modglmm"]["LOCOResult"][j - 1]
        assert np.allclose(a["fitted.values"], b["fitted.values"], rtol=5e-3)
        assert set(b.keys()) - {"alpha0"} <= set(a.keys())
        M = step2.ReadModel("%s_chr%d.rda" % (out, j), chrom=str(j), LOCO=True)
        assert np.array_equal(M["mu"], a["fitted.values"].ravel()) and M["XVX"].shape == (3, 3)
        with pytest.raises(ValueError):
            step2.ReadModel("%s_chr%d.rda" % (out, j), chrom=str(j % 22 + 1), LOCO=True)


def test_phenotype_file_variants_sex_filter_and_categorical_covariate(golden_dir, tmp_path):
    """A gzipped comma-separated phenotype file with missing entries, a categorical covariate (qCovarCol -> treatment
    contrasts), FemaleOnly (output prefix gets _FemaleOnly, FG.R:790-800) and a covariate that fails checkPerfectSep."""
    from saige_gpu_b200 import fitnull
    rows = [l.split() for l in open(os.path.join(golden_dir, "pheno_1000samples.txt"))]
    hdr, body = rows[0], rows[1:]
    rng = np.random.default_rng(7)
    site = rng.choice(["north", "south", "west"], size=len(body))
    yb = np.array([r[hdr.index("y_binary")] for r in body])
    rare = np.where((rng.uniform(size=len(body)) < 0.01) & (yb == "0"), "1", "0")          # only controls carry it: an empty 2 x 2 cell
    path = str(tmp_path / "pheno.csv.gz")
    with gzip.open(path, "wt") as f:
        f.write(",".join(hdr + ["site", "rare"]) + "\n")
        for i, r in enumerate(body):
            r = list(r)
            if i % 97 == 0:
                r[hdr.index("x1")] = "NA"                                                   # incomplete case: dropped
            f.write(",".join(r + [site[i], rare[i]]) + "\n")
    out = str(tmp_path / "sex")
    r = fitnull.fitNULLGLMM(OracleBackend(), plinkFile=os.path.join(golden_dir, "grm10k"), phenoFile=path, phenoCol="y_binary",
                            covarColList=["x1", "site", "rare"], qCovarCol=["site"], sampleIDColinphenoFile="IID", traitType="binary",
                            outputPrefix=out, LOCO=False, sexCol="x2", FemaleCode=1, FemaleOnly=True, minCovariateCount=1,
                            probe_rng="numpy", skipVarianceRatioEstimation=True)
    assert r["modelFile"] == out + "_FemaleOnly.rda" and os.path.exists(r["modelFile"])
    m = r["modglmm"]
    x2 = {row[hdr.index("IID")]: row[hdr.index("x2")] for row in body}
    assert all(x2[s] == "1" for s in m["sampleID"]) and 400 < len(m["sampleID"]) < 600
    assert not any(body[i][hdr.index("IID")] in m["sampleID"] for i in range(0, len(body), 97))
    assert m["X"].shape[1] == 4 and m["coefficients"].shape == (4, 1)          # intercept, x1, sitesouth, sitewest; `rare` dropped
    assert m["theta"][0] == 1.0 and np.all(np.isfinite(m["coefficients"]))
    with pytest.raises(fitnull.SaigeInputError):
        fitnull.fitNULLGLMM(OracleBackend(), plinkFile=os.path.join(golden_dir, "grm10k"), phenoFile=path, phenoCol="y_binary",
                            covarColList=["x1"], sampleIDColinphenoFile="IID", outputPrefix=out, FemaleOnly=True, MaleOnly=True)
    with pytest.raises(fitnull.SaigeInputError):
        fitnull.fitNULLGLMM(OracleBackend(), plinkFile=os.path.join(golden_dir, "grm10k"), phenoFile=path, phenoCol="y_binary",
                            covarColList=["x1", "site"], sampleIDColinphenoFile="IID", outputPrefix=out)       # `site` is not numeric


def test_categorical_variance_ratios(golden_dir, bim22, tmp_path):
    """isCateVarianceRatio: every marker with 10 <= MAC < 20.5 is held out of the GRM (FG.cpp:497-501), one ratio per MAC
    category is estimated and written as `<ratio> null <k>`; step 2 picks the ratio by the variant's MAC."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import step1
    from saige_gpu_b200.rdata import load_rda
    from saige_gpu_b200.step2 import Get_Variance_Ratio
    out = str(tmp_path / "cate")
    be = OracleBackend()
    r = _run(be, golden_dir, bim22, out, isCateVarianceRatio=True, LOCO=False)
    lines = [l.split() for l in open(out + ".varianceRatio.txt")]
    assert [l[1:] for l in lines] == [["null", "1"], ["null", "2"]]
    vr = Get_Variance_Ratio(out + ".varianceRatio.txt")
    assert isinstance(vr, list) and len(vr) == 2 and np.allclose(vr, r["varianceRatio"], rtol=1e-14)
    assert all(0.5 < v < 1.5 for v in vr) and vr[0] != vr[1]
    mac_vr = np.asarray(be.getMACVec_forVarRatio())
    assert ((mac_vr >= 10) & (mac_vr < 20.5)).sum() > 300 and (mac_vr < 10).sum() == 0          # the whole category is in the hold-out store
    assert (np.asarray(be.getMACVec()) >= 20.5).all() or (np.asarray(be.getMACVec()) < 10).any()  # ... and none of it in the GRM
    with pytest.raises(ValueError):
        Get_Variance_Ratio(out + ".varianceRatio.txt", (10, 20.5, 30), (20.5, 30))
    # a category switched off gets 1; too few markers in a category is an error, as in the reference
    m = load_rda(out + ".rda")["modglmm"]
    model = dict(fitted_values=m["fitted.values"].ravel(), linear_predictors=m["linear.predictors"].ravel(), y=m["y"], X=m["X"],
                 theta=m["theta"], obj_noK=dict(m["obj.noK"]), traitType="binary")
    rng = np.random.default_rng(3)
    pc = step1.extractVarianceRatio_cate(be, model, step1.Binomial, mac_vr, np.ones(len(mac_vr), bool), rng, (10, 20.5), (20.5,), [0, 1])
    assert pc[0] == (1.0, []) and 0.5 < pc[1][0] < 1.5
    with pytest.raises(ValueError):
        step1.extractVarianceRatio_cate(be, model, step1.Binomial, mac_vr, np.ones(len(mac_vr), bool), rng, (10, 20.5), (20.5,),
                                        numMarkers=5000)
    # step 2 (oracle) with the two ratios: the ratio follows the variant's MAC category
    M = S2.read_model(m, LOCO=False)
    M.update(varRatio=vr, cateVarRatioMinMACVecExclude=(10, 20.5), cateVarRatioMaxMACVecInclude=(20.5,))
    assert [S2.assign_variance_ratio(M, x) for x in (3, 10, 10.5, 20.5, 21, 900)] == [vr[0], vr[0], vr[0], vr[0], vr[1], vr[1]]


# ---- through the C ABI ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_fitnullglmm_matches_oracle_backed_run_and_feeds_step2(oracle_fit, golden_dir, bim22, tmp_path):
    from saige_gpu_b200 import SaigeB200, step2
    from saige_gpu_b200.rdata import load_rda
    ro, _ = oracle_fit
    out = str(tmp_path / "gpu_binary")
    g = SaigeB200(device=0)
    rg = _run(g, golden_dir, bim22, out)
    a, b = rg["modglmm"], ro["modglmm"]
    rel = lambda x, y: float(np.max(np.abs(np.asarray(x) - np.asarray(y))) / np.max(np.abs(np.asarray(y))))
    assert rel(a["theta"], b["theta"]) < TOL_FIT
    assert rel(a["coefficients"], b["coefficients"]) < TOL_FIT
    assert rel(a["fitted.values"], b["fitted.values"]) < TOL_FIT
    for j in range(22):
        assert rel(a["LOCOResult"][j]["fitted.values"], b["LOCOResult"][j]["fitted.values"]) < TOL_FIT
        assert rel(a["LOCOResult"][j]["obj.noK"]["XVX_inv_XV"], b["LOCOResult"][j]["obj.noK"]["XVX_inv_XV"]) < TOL_FIT
    assert abs(rg["varianceRatio"] - ro["varianceRatio"]) < TOL_FIT * ro["varianceRatio"]
    # step 2 on the GPU from the files step 1 just wrote (the handle passed in stays open)
    p = os.path.join(golden_dir, "step2_100markers")
    rows = step2.SPAGMMATtest(g, p + ".bed", p + ".bim", p + ".fam", out + ".rda", out + ".varianceRatio.txt", chrom="1", LOCO=True,
                              min_MAC=20)
    assert len(rows) == 32 and all(0 < r["p.value"] <= 1 for r in rows)
    assert load_rda(out + ".rda")["modglmm"]["theta"][1] == a["theta"][1]
    g.close()


def test_rda_writer_round_trip_property():
    """Property test (hypothesis): any nesting of the value types a model file holds survives save_rda -> load_rda."""
    import tempfile
    from hypothesis import given, settings, strategies as st
    from saige_gpu_b200.rdata import load_rda, save_rda

    names = st.text(alphabet="abcXYZ._0123456789é", min_size=1, max_size=8).filter(lambda s: s not in ("",))
    floats = st.floats(allow_nan=False, allow_infinity=True, width=64)
    leaf = st.one_of(
        st.lists(floats, min_size=0, max_size=6).map(lambda v: np.array(v, dtype=np.float64)),
        st.lists(st.integers(-2 ** 31 + 1, 2 ** 31 - 1), min_size=0, max_size=6).map(lambda v: np.array(v, dtype=np.int32)),
        st.lists(st.sampled_from([1, 0, -1]), min_size=1, max_size=5).map(lambda v: np.array(v, dtype=np.int8)),
        st.lists(st.one_of(st.none(), st.text(alphabet="abc é\t", max_size=5)), min_size=1, max_size=4).filter(
            lambda v: any(isinstance(x, str) for x in v)),
        st.tuples(st.integers(1, 3), st.integers(1, 3)).flatmap(
            lambda s: st.lists(floats, min_size=s[0] * s[1], max_size=s[0] * s[1]).map(lambda v: np.array(v).reshape(s))),
        st.none())
    tree = st.recursive(leaf, lambda kids: st.one_of(st.dictionaries(names, kids, min_size=1, max_size=4),
                                                     st.lists(st.dictionaries(names, kids, min_size=1, max_size=2), min_size=1, max_size=3)),
                        max_leaves=12)

    def same(a, b):
        if isinstance(a, dict):
            return isinstance(b, dict) and list(a) == list(b) and all(same(a[k], b[k]) for k in a)
        if isinstance(a, list):
            return isinstance(b, list) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        if isinstance(a, np.ndarray):
            return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
        if isinstance(a, str):
            return a == b
        return a is None and b is None

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "p.rda")

        @settings(max_examples=60, deadline=None)
        @given(st.dictionaries(names, tree, min_size=1, max_size=3))
        def check(obj):
            save_rda(path, obj)
            back = load_rda(path)
            assert same(obj, back), (obj, back)
        check()
