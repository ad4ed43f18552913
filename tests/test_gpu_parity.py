"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star):
  decode / allele counts / MAC / QC mask / VR hold-out ....... bit-exact
  GRM products, LOCO products, GRM diagonal .................. <= 1e-10 relative (max-norm over the vector)
  PCG solutions, coefficients, AI-REML scalars, tau, variance ratio ... <= 1e-6 relative
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_MATVEC = 1e-10
TOL_FIT = 1e-6


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def gpu():
    from saige_gpu_b200 import SaigeB200
    g = SaigeB200(device=0)
    yield g
    g.close()


@pytest.fixture()
def gpu2():
    """A second, throw-away context for tests that load other genotype sets."""
    from saige_gpu_b200 import SaigeB200
    g = SaigeB200(device=0)
    yield g
    g.close()


@pytest.fixture(scope="module")
def pair10k(gpu, grm10k):
    """GPU context and oracle both loaded with the bundled 10k-marker set, MAF >= 0.01, missing <= 0.15."""
    from oracle import oracle as O
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno()
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    gpu.setminMAFforGRM(0.01)
    gpu.setmaxMissingRateforGRM(0.15)
    gpu.setminMAC_VarianceRatio(20, -1, False)
    p = grm10k["prefix"]
    gpu.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    return gpu, o


def test_ingest_bit_exact_from_file(pair10k):
    g, o = pair10k
    assert (g.N, g.M, g.M0) == (o.N, o.M, o.M0) == (1000, 9650, 10000)
    assert np.array_equal(g.getAlleleCountVec(), o.ACVec)
    assert np.array_equal(g.getMACVec(), o.MACVec)
    assert np.array_equal(g.getQCdMarkerIndex(), o.qc_mask)
    assert np.array_equal(g.getAlleleFreqVec().astype(np.float32), o.alleleFreqVec)
    for idx in (0, 1, 17, 4096, o.M - 1):
        assert np.array_equal(g.Get_OneSNP_Geno(idx), o.Get_OneSNP_Geno(idx))
        assert rel(g.Get_OneSNP_StdGeno(idx), o.Get_OneSNP_StdGeno(idx)) < 1e-14


def test_frq_golden_through_gpu(pair10k, golden_dir):
    """The reference's own .frq fixture pins the allele counting convention."""
    g, _ = pair10k
    frq = np.array([float(l.split()[4]) for l in open(os.path.join(golden_dir, "grm10k.frq")).readlines()[1:]])
    mask = g.getQCdMarkerIndex()
    mine = g.getAlleleCountVec() / 2000.0
    assert all(float("%.4g" % a) == b for a, b in zip(mine, frq[mask]))


@pytest.mark.parametrize("engine", ["tensor", "f64", "imma", "umma"])
@pytest.mark.parametrize("k", [1, 2, 3, 7, 16, 31])
def test_crossprod_matches_oracle(pair10k, engine, k):
    g, o = pair10k
    g.set_engine(engine)
    rng = np.random.default_rng(100 + k)
    B = rng.normal(size=(o.N, k)) * 10.0 ** rng.integers(-3, 4, size=k)
    if k >= 2:
        B[:, 1] = rng.integers(0, 2, size=o.N) * 2.0 - 1.0          # a Rademacher probe
    Y = g.getCrossprodMatAndKin(B)
    Yo = o.getCrossprodMatAndKin(B)
    for c in range(k):
        assert rel(Y[:, c], Yo[:, c]) < TOL_MATVEC, (engine, k, c)
    g.set_engine("tensor")


def test_crossprod_special_vectors(pair10k):
    g, o = pair10k
    ones = np.ones(o.N)
    y = g.getCrossprodMatAndKin(ones)
    assert np.max(np.abs(y)) < 1e-9                     # columns are centred with the exact allele frequency
    z = g.getCrossprodMatAndKin(np.zeros(o.N))
    assert np.all(z == 0)
    e = np.zeros(o.N); e[3] = 1.0
    assert rel(g.getCrossprodMatAndKin(e), o.getCrossprodMatAndKin(e)) < TOL_MATVEC
    # column maxima just below a power of two exercise the top of the fixed-point limb range
    rng = np.random.default_rng(11)
    for top in (1.984375, 1.9999999999999998, 1.0, 3.999):
        v = rng.uniform(-1, 1, size=o.N)
        v[rng.integers(0, o.N, 5)] = top
        v[rng.integers(0, o.N, 5)] = -top
        assert rel(g.getCrossprodMatAndKin(v), o.getCrossprodMatAndKin(v)) < TOL_MATVEC, top


def test_crossprod_is_linear_and_symmetric(pair10k):
    g, _ = pair10k
    rng = np.random.default_rng(5)
    a, b = rng.normal(size=g.N), rng.normal(size=g.N)
    Ka, Kb, Kab = g.getCrossprodMatAndKin(a), g.getCrossprodMatAndKin(b), g.getCrossprodMatAndKin(2 * a - 3 * b)
    assert rel(Kab, 2 * Ka - 3 * Kb) < 1e-10
    assert abs(a @ Kb - b @ Ka) / abs(a @ Kb) < 1e-10
    assert a @ Ka > 0


def test_batched_equals_sequential(pair10k):
    """Exact integer accumulation: a column's result is bit-identical whatever the batch width and whichever tensor
    kernel (mma.sync or tcgen05) computed it."""
    g, _ = pair10k
    rng = np.random.default_rng(6)
    B = rng.normal(size=(g.N, 5))
    Y = g.getCrossprodMatAndKin(B)                                   # k = 5 -> tcgen05 kernel
    for c in range(5):
        assert np.array_equal(Y[:, c], g.getCrossprodMatAndKin(B[:, c])), "column result must not depend on the batch"
    g.set_engine("imma")
    assert np.array_equal(Y, g.getCrossprodMatAndKin(B))
    g.set_engine("umma")
    assert np.array_equal(Y[:, :2], g.getCrossprodMatAndKin(B[:, :2]))
    g.set_engine("tensor")


@pytest.mark.parametrize("engine", ["tensor", "f64"])
def test_diag_of_kin(pair10k, engine):
    g, o = pair10k
    g.set_engine(engine)
    d = g.get_DiagofKin()
    assert rel(d, o.get_DiagofKin()) < TOL_MATVEC
    assert abs(d[:4] - np.array([1.00114237, 1.08651917, 1.03021335, 1.05902383])).max() < 1e-6
    g.set_engine("tensor")


def test_plink2_allele_counts_through_gpu(gpu2, chr22, golden_dir):
    """plink2's allele counts of the 22-chromosome set (shipped by the reference) against the GPU ingest."""
    rows = [l.split() for l in open(os.path.join(golden_dir, "chr22_1000.acount"))][1:]
    g = gpu2
    g.setminMAFforGRM(0.0); g.setmaxMissingRateforGRM(1.0); g.setminMAC_VarianceRatio(20, -1, False)
    p = chr22["prefix"]
    g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, chr22["N0"] + 1), np.ones(chr22["N0"], np.uint8))
    assert g.M == 1000 and np.array_equal(g.getAlleleCountVec(), np.array([int(r[4]) for r in rows]))


def test_loco_products_and_diag(gpu2, chr22):
    gpu = gpu2
    from oracle import oracle as O
    from saige_gpu_b200 import step1
    N0, M0 = chr22["N0"], chr22["M0"]
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(chr22["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    gpu.setminMAFforGRM(0.01); gpu.setmaxMissingRateforGRM(0.15); gpu.setminMAC_VarianceRatio(20, -1, False)
    gpu.setgeno_mem(chr22["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    assert np.array_equal(gpu.getQCdMarkerIndex(), o.qc_mask)
    chrq = chr22["chrs"][o.qc_mask]
    LOCO, s, e = O.updateChrStartEndIndexVec(chrq)
    o.setStartEndIndexVec(s, e)
    assert step1.set_loco_ranges(gpu, chrq) == LOCO
    o.set_Diagof_StdGeno_LOCO(); gpu.set_Diagof_StdGeno_LOCO()
    rng = np.random.default_rng(3)
    B = rng.normal(size=(N0, 3))
    w = rng.uniform(0.05, 0.25, size=N0); tau = np.array([1.0, 0.4])
    for j in (0, 7, 21):
        o.setStartEndIndex(s[j], e[j], j); gpu.setStartEndIndex(s[j], e[j], j)
        assert rel(gpu.getCrossprodMatAndKin_LOCO(B), o.getCrossprodMatAndKin_LOCO(B)) < TOL_MATVEC
        assert rel(gpu.getDiagOfSigma_LOCO(w, tau), o.getDiagOfSigma(w, tau, loco=True)) < TOL_MATVEC
        x, it = gpu.getPCG1ofSigmaAndVector(w, tau, B[:, 0], 500, 1e-5, loco=True, return_iter=True)
        xo, ito = o.getPCG1ofSigmaAndVector(w, tau, B[:, 0], 500, 1e-5, loco=True, return_iter=True)
        assert it == ito and rel(x, xo) < TOL_FIT


def test_ingest_missing_subset_and_vr_holdout(gpu2):
    gpu = gpu2
    """Synthetic .bed with 2% missing calls, a phenotyped subset in shuffled order and a VR hold-out index set."""
    from oracle import oracle as O
    N0, M0 = 1237, 3000
    bed = O.synth_bed(N0, M0, seed=77, miss_rate=0.02)
    rng = np.random.default_rng(4)
    keep = np.sort(rng.choice(N0, size=1001, replace=False))
    sub = rng.permutation(keep) + 1
    ind = np.zeros(N0, np.uint8); ind[keep] = 1
    vr = np.unique(rng.integers(0, M0, size=200))
    o = O.OracleGeno(); o.minMAF, o.maxMissing, o.isVarRatio = 0.06, 0.03, True
    o.setgeno(bed, N0, M0, sub, ind, vr_rand_idx=vr)
    gpu.setminMAFforGRM(0.06); gpu.setmaxMissingRateforGRM(0.03); gpu.setminMAC_VarianceRatio(20, -1, True)
    gpu.setgeno_mem(bed, N0, M0, sub, ind, vr_rand_idx=vr)
    assert (gpu.N, gpu.M, gpu.Mvr) == (o.N, o.M, o.Mvr) and o.Mvr > 50 and 0 < o.M < M0 - o.Mvr
    assert np.array_equal(gpu.getAlleleCountVec(), o.ACVec)
    assert np.array_equal(gpu.getMACVec(), o.MACVec)
    assert np.array_equal(gpu.getQCdMarkerIndex(), o.qc_mask)
    assert np.array_equal(gpu.getMACVec_forVarRatio(), o.MACVec_forVarRatio)
    assert np.array_equal(gpu.getIndexVec_forVarRatio(), o.markerIndexVec_forVarRatio)
    for idx in range(0, o.M, 97):
        assert np.array_equal(gpu.Get_OneSNP_Geno(idx), o.Get_OneSNP_Geno(idx))
    for idx in range(0, o.Mvr, 11):
        assert np.array_equal(gpu.Get_OneSNP_Geno_forVarRatio(idx), o.Get_OneSNP_Geno(idx, vr=True))
    b = rng.normal(size=o.N)
    assert rel(gpu.getCrossprodMatAndKin(b), o.getCrossprodMatAndKin(b)) < TOL_MATVEC
    gpu.setminMAC_VarianceRatio(20, -1, False)


def test_pcg_multi_rhs_equals_sequential_oracle(pair10k):
    g, o = pair10k
    rng = np.random.default_rng(9)
    w = rng.uniform(0.02, 0.25, size=o.N)
    tau = np.array([1.0, 0.35])
    B = np.column_stack([rng.normal(size=o.N), rng.integers(0, 2, size=o.N) * 2.0 - 1, np.ones(o.N), np.zeros(o.N),
                         1e-4 * rng.normal(size=o.N)])
    X, it = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    Xo, ito = o.pcg_multi(w, tau, B, 500, 1e-5)
    assert list(it) == list(ito)
    for c in range(B.shape[1]):
        assert rel(X[:, c], Xo[:, c]) < TOL_FIT or np.all(Xo[:, c] == 0) and np.all(X[:, c] == 0)
    # tau1 == 0 short-circuit (FG.cpp:2401-2404)
    x0 = g.getPCG1ofSigmaAndVector(w, np.array([1.0, 0.0]), B[:, 0], 500, 1e-5)
    assert rel(x0, o.getPCG1ofSigmaAndVector(w, np.array([1.0, 0.0]), B[:, 0], 500, 1e-5)) < TOL_FIT
    assert rel(g.getDiagOfSigma(w, tau), o.getDiagOfSigma(w, tau)) < TOL_MATVEC
    assert rel(g.getCrossprod(B[:, 0], w, tau), o.getCrossprod(B[:, 0], w, tau)) < TOL_MATVEC


def _pheno(golden_dir):
    rows = [l.split() for l in open(os.path.join(golden_dir, "pheno_1000samples.txt")).readlines()]
    hdr, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    yb = np.array([float(r[col["y_binary"]]) for r in rows])
    yq = np.array([float(r[col["y_quantitative"]]) for r in rows])
    X = np.column_stack([np.ones(len(rows)), [float(r[col["x1"]]) for r in rows], [float(r[col["x2"]]) for r in rows]])
    return yb, yq, X


def test_ai_reml_pieces(pair10k, golden_dir):
    from oracle import oracle as O
    g, o = pair10k
    yb, _, X = _pheno(golden_dir)
    fit0 = O.glm_fit(yb, X, O.Binomial)
    mu = fit0["mu"]; W = mu * (1 - mu); Y = fit0["eta"] + (yb - mu) / W
    tau = np.array([1.0, 0.1])
    rg = g.getCoefficients(Y, X, W, tau, 500, 1e-5)
    ro = O.getCoefficients(o, Y, X, W, tau, 500, 1e-5)
    for key in ("Sigma_iY", "Sigma_iX", "cov", "alpha", "eta"):
        assert rel(rg[key], ro[key]) < TOL_FIT, key
    U = np.random.default_rng(200).integers(0, 2, size=(o.N, 60)) * 2.0 - 1.0
    draws = O.make_draw(U)
    ag = g.getAIScore(Y, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.0025, draws())
    ao = O.getAIScore(o, Y, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.0025, draws())
    assert ag["nrun_used"] == ao["nrun_used"]
    for key in ("YPAPY", "Trace", "AI", "PY"):
        assert rel(ag[key], ao[key]) < TOL_FIT, key
    tg = g.fitglmmaiRPCG(Y, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.02, 0.0025, draws())["tau"]
    to = O.fitglmmaiRPCG(o, Y, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.02, 0.0025, draws())
    assert rel(tg, to) < TOL_FIT
    # forcing CV retries: an absurdly small cutoff makes both sides extend by 10 probes the same number of times
    ag = g.getAIScore(Y, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.0011, draws())
    ao = O.getAIScore(o, Y, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.0011, draws())
    assert ag["nrun_used"] == ao["nrun_used"] and rel(ag["Trace"], ao["Trace"]) < TOL_FIT


def test_probe_product_cache_is_transparent(pair10k, golden_dir):
    """K.U of the Hutchinson probes is reused while U stays bitwise the same (GetTrace re-seeds before every call,
    FG.cpp:3114); the reuse must not change a single bit, and any other U, engine or genotype set must miss."""
    from oracle import oracle as O
    g, o = pair10k
    yb, _, X = _pheno(golden_dir)
    fit0 = O.glm_fit(yb, X, O.Binomial)
    mu = fit0["mu"]; W = mu * (1 - mu); Y = fit0["eta"] + (yb - mu) / W
    tau = np.array([1.0, 0.2])
    rc = g.getCoefficients(Y, X, W, tau, 500, 1e-5)
    U = np.random.default_rng(7).integers(0, 2, size=(o.N, 40)) * 2.0 - 1.0
    draws = O.make_draw(U)
    args = (Y, X, W, tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], 30, 500, 1e-5, 0.0025)
    g.set_engine("tensor")                       # also drops whatever an earlier test cached
    g.reset_counters()
    a1 = g.getAIScore(*args, draws())
    assert g.counters()["n_probe_product_reuse"] == 0
    a2 = g.getAIScore(*args, draws())
    assert g.counters()["n_probe_product_reuse"] == 1
    for key in ("YPAPY", "Trace", "AI", "PY"):
        assert np.array_equal(np.asarray(a1[key]), np.asarray(a2[key])), key
    U2 = U.copy(); U2[17, 3] = -U2[17, 3]        # one flipped sign: a different probe set
    a3 = g.getAIScore(*args, O.make_draw(U2)())
    assert g.counters()["n_probe_product_reuse"] == 1
    assert a3["Trace"] != a1["Trace"]
    ao = O.getAIScore(o, *args, O.make_draw(U2)())
    assert rel(a3["Trace"], ao["Trace"]) < TOL_FIT


def test_ai_reml_quantitative_pieces(pair10k, golden_dir):
    from oracle import oracle as O
    g, o = pair10k
    _, yq, X = _pheno(golden_dir)
    W = np.ones(o.N); tau = np.array([0.6, 0.3])
    ro = O.getCoefficients(o, yq, X, W, tau, 500, 1e-5)
    U = np.random.default_rng(200).integers(0, 2, size=(o.N, 80)) * 2.0 - 1.0
    draws = O.make_draw(U)
    ag = g.getAIScore_q(yq, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.0025, draws())
    ao = O.getAIScore_q(o, yq, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.0025, draws())
    assert ag["nrun_used"] == ao["nrun_used"]
    for key in ("YPAPY", "YPA0PY", "Trace", "AI", "PY"):
        assert rel(ag[key], ao[key]) < TOL_FIT, key
    tg = g.fitglmmaiRPCG_q(yq, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.02, 0.0025, draws())["tau"]
    to = O.fitglmmaiRPCG_q(o, yq, X, W, tau, ro["Sigma_iY"], ro["Sigma_iX"], ro["cov"], 30, 500, 1e-5, 0.02, 0.0025, draws())
    assert rel(tg, to) < TOL_FIT


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_full_step1_matches_oracle(pair10k, golden_dir, trait):
    """Whole null-GLMM fit (config 1 of BASELINE.json on the bundled 10k-marker set): tau, alpha, fitted values and
    the variance ratio against the fp64 oracle run with identical probes."""
    from oracle import oracle as O
    from saige_gpu_b200 import step1
    g, o = pair10k
    yb, yq, X = _pheno(golden_dir)
    probes = step1.ProbeStream(o.N, nmax=130, seed=200)
    if trait == "binary":
        fam_o, fam_g, y = O.Binomial, step1.Binomial, yb
    else:
        fam_o, fam_g, y = O.Gaussian, step1.Gaussian, yq
    fit0o = O.glm_fit(y, X, fam_o)
    fit0g = step1.glm_fit(y, X, fam_g)
    mo = O.glmmkin_ai_PCG(o, fit0o, (0, 0), probes.U, trait=trait)
    mg = step1.glmmkin_ai_PCG(g, fit0g, probes, trait=trait)
    assert mg["converged"] == mo["converged"]
    assert rel(mg["theta"], mo["theta"]) < TOL_FIT
    assert rel(mg["coefficients"], mo["coefficients"]) < TOL_FIT
    assert rel(mg["fitted_values"], mo["fitted_values"]) < TOL_FIT
    assert rel(mg["obj_noK"]["XVX_inv_XV"], mo["obj_noK"]["XVX_inv_XV"]) < TOL_FIT
    order = np.random.default_rng(1).permutation(o.M)[:200]
    vo, _ = O.extractVarianceRatio(o, mo, fam_o, order)
    vg, _ = step1.extractVarianceRatio(g, mg, fam_g, order)
    assert rel(vg, vo) < TOL_FIT
    if trait == "binary":
        # the reference's own result for this cohort / phenotype / covariates (extdata/output/example.rda):
        # theta = (1, 0.32472724), alpha = (-2.97337569, 0.7511719, 0.91698671).  alpha is RNG-free up to the PCG
        # tolerance; tau carries the Monte-Carlo error of R's 30 probes vs ours (tests/test_oracle_golden.py).
        ref_alpha = np.array([-2.97337569, 0.7511719, 0.91698671])
        assert mg["theta"][0] == 1.0 and abs(mg["theta"][1] - 0.32472724) / 0.32472724 < 0.08
        assert np.max(np.abs(mg["coefficients"] - ref_alpha) / np.abs(ref_alpha)) < 5e-3


@pytest.mark.parametrize("shape", [(5, 9), (37, 50), (255, 257), (256, 1024), (1025, 777), (3001, 130)])
def test_ragged_and_tiny_shapes(gpu2, shape):
    """Sizes that are not multiples of the packing / tile units (4 samples per byte, 256-genotype k-steps, 128/512-row
    tiles), down to a handful of samples: every engine against the oracle."""
    from oracle import oracle as O
    N0, M0 = shape
    bed = O.synth_bed(N0, M0, seed=1000 + N0, miss_rate=0.03)
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.0, 1.0
    o.setgeno(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    g = gpu2
    g.setminMAFforGRM(0.0); g.setmaxMissingRateforGRM(1.0); g.setminMAC_VarianceRatio(20, -1, False)
    g.setgeno_mem(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    assert (g.N, g.M) == (o.N, o.M) and np.array_equal(g.getAlleleCountVec(), o.ACVec)
    for idx in (0, o.M // 2, o.M - 1):
        assert np.array_equal(g.Get_OneSNP_Geno(idx), o.Get_OneSNP_Geno(idx))
    rng = np.random.default_rng(N0)
    B = rng.normal(size=(N0, 5))
    want = o.getCrossprodMatAndKin(B)
    for eng in ("tensor", "imma", "umma", "f64"):
        g.set_engine(eng)
        assert rel(g.getCrossprodMatAndKin(B), want) < TOL_MATVEC, eng
        assert rel(g.getCrossprodMatAndKin(B[:, 0]), want[:, 0]) < TOL_MATVEC, eng
        assert rel(g.get_DiagofKin(), o.get_DiagofKin()) < TOL_MATVEC, eng
    g.set_engine("tensor")


def test_pcg_limits_and_flags(pair10k):
    g, o = pair10k
    rng = np.random.default_rng(21)
    w = rng.uniform(0.02, 0.25, size=o.N); tau = np.array([1.0, 2.5]); b = rng.normal(size=o.N)
    # maxiterPCG reached before convergence (FG.cpp:2794-2796 prints "pcg did not converge"): same truncated iterate
    x, it = g.getPCG1ofSigmaAndVector(w, tau, b, 2, 1e-12, return_iter=True)
    xo, ito = o.getPCG1ofSigmaAndVector(w, tau, b, 2, 1e-12, return_iter=True)
    assert it == ito == 2 and rel(x, xo) < TOL_FIT
    # diag(Sigma) floor at 1e-4 (FG.cpp:2355-2357)
    wbig = np.full(o.N, 1e9)
    d = g.getDiagOfSigma(wbig, np.array([1.0, 0.0]))
    assert np.all(d == 1e-4) and np.array_equal(d, o.getDiagOfSigma(wbig, np.array([1.0, 0.0])))


def test_kin_diag_set_as_one(gpu2, grm10k):
    """isDiagofKinSetAsOne (FG.cpp:2334-2340, 4372): diag(K) := 1 in the preconditioner and in get_DiagofKin."""
    from oracle import oracle as O
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8), isDiagofKinSetAsOne=True)
    g = gpu2
    g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
    g.setgeno_mem(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8), isDiagofKinSetAsOne=True)
    assert np.all(g.get_DiagofKin() == 1.0)
    rng = np.random.default_rng(2)
    w = rng.uniform(0.05, 0.25, size=N0); tau = np.array([1.0, 0.4]); b = rng.normal(size=N0)
    assert rel(g.getDiagOfSigma(w, tau), o.getDiagOfSigma(w, tau)) < 1e-14
    x, it = g.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, return_iter=True)
    xo, ito = o.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, return_iter=True)
    assert it == ito and rel(x, xo) < TOL_FIT


def test_error_paths(gpu2, golden_dir):
    from saige_gpu_b200 import SaigeB200Error
    g = gpu2
    with pytest.raises(SaigeB200Error, match="not loaded"):
        g.N = 10; g.getCrossprodMatAndKin(np.zeros(10))
    with pytest.raises(SaigeB200Error, match="bed file not open|fam file not open|bim file not open"):
        g.setgeno("/nonexistent.bed", "/nonexistent.bim", "/nonexistent.fam", [1], [1])
    p = os.path.join(golden_dir, "grm10k")
    with pytest.raises(SaigeB200Error, match="out of range"):
        g.setgeno(p + ".bed", p + ".bim", p + ".fam", [0, 5], np.ones(1000, np.uint8))
    with pytest.raises(SaigeB200Error, match="indicator length"):
        g.setgeno(p + ".bed", p + ".bim", p + ".fam", [1, 2], np.ones(10, np.uint8))
    g.setminMAFforGRM(0.01)
    g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, 1001), np.ones(1000, np.uint8))
    with pytest.raises(SaigeB200Error, match="out of range"):
        g.Get_OneSNP_Geno(g.M)
    with pytest.raises(SaigeB200Error, match="setStartEndIndex"):
        g.getCrossprodMatAndKin_LOCO(np.zeros(g.N))


def test_reference_pcg_log_lines(pair10k, capfd):
    """sgb_set_verbose: every solve prints the reference's stdout lines (FG.cpp:2794-2798), one per right-hand side, so that
    log scrapers written for the reference keep working although the solves are batched."""
    g, o = pair10k
    rng = np.random.default_rng(3)
    B = rng.normal(size=(o.N, 3))
    w = rng.uniform(0.05, 0.25, size=o.N)
    g.set_verbose(True)
    try:
        _, it = g.getPCG1ofSigmaAndVector(w, np.array([1.0, 0.3]), B, 500, 1e-5, return_iter=True)
        _, it2 = g.getPCG1ofSigmaAndVector(w, np.array([1.0, 0.3]), B[:, 0], 2, 1e-5, return_iter=True)      # hits maxiter
    finally:
        g.set_verbose(False)
    lines = [l for l in capfd.readouterr().out.splitlines() if l.strip()]
    want = ["iter from getPCG1ofSigmaAndVector %d" % v for v in it]
    want += ["pcg did not converge. You may increase maxiter number.", "iter from getPCG1ofSigmaAndVector 2"]
    assert lines == want, lines
