"""The CUDA library against the REFERENCE'S OWN solver code (oracle/_ref/libfg_ref64.so: getPCG1ofSigmaAndVector,
getCoefficients, GetTrace, getAIScore, fitglmmaiRPCG and the _q / _LOCO variants compiled unmodified from
SAIGE_fitGLMM_fast.cpp:2322-3662 with `float` read as `double`; see tests/test_reference_solver.py and oracle/Makefile).
The reference code gets its GRM product and diagonal from the CPU oracle; everything above the product is the reference's text.

north_star tolerances: PCG solutions / coefficients / tau <= 1e-6 relative, PCG iteration counts identical."""
import os

import numpy as np
import pytest

from oracle import ref_solver as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(64), reason="oracle/_ref/libfg_ref64.so not built")]
TOL = 1e-6


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def trio(grm10k, golden_dir):
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200, step1
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    g = SaigeB200(device=0)
    g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
    p = grm10k["prefix"]
    g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    probes = step1.ProbeStream(o.N, nmax=200, seed=200)
    rows = [l.split() for l in open(os.path.join(golden_dir, "pheno_1000samples.txt")).readlines()]
    col = {h: i for i, h in enumerate(rows[0])}
    yb = np.array([float(r[col["y_binary"]]) for r in rows[1:]])
    yq = np.array([float(r[col["y_quantitative"]]) for r in rows[1:]])
    X = np.column_stack([np.ones(N0), [float(r[col["x1"]]) for r in rows[1:]], [float(r[col["x2"]]) for r in rows[1:]]])
    chrq = np.array([int(c) for c in grm10k["chrs"]])[o.qc_mask]
    _, s, e = O.updateChrStartEndIndexVec(chrq)
    o.setStartEndIndexVec(s, e); o.set_Diagof_StdGeno_LOCO()
    step1.set_loco_ranges(g, chrq); g.set_Diagof_StdGeno_LOCO()
    yield g, o, R.RefSolver(o, 64, probes.U), probes, yb, yq, X
    g.close()


def test_pcg_and_coefficients_vs_reference_code(trio):
    g, o, r, probes, yb, yq, X = trio
    rng = np.random.default_rng(11)
    w = rng.uniform(0.02, 0.25, size=o.N); tau = np.array([1.0, 0.45]); B = rng.normal(size=(o.N, 4))
    Xg, itg = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    for c in range(B.shape[1]):
        xr, itr = r.getPCG1ofSigmaAndVector(w, tau, B[:, c], 500, 1e-5, return_iter=True)
        assert int(itg[c]) == itr and rel(Xg[:, c], xr) < TOL
    Y = rng.normal(size=o.N) + 1.0
    a, b = g.getCoefficients(Y, X, w, tau, 500, 1e-5), r.getCoefficients(Y, X, w, tau, 500, 1e-5)
    for key in ("Sigma_iY", "Sigma_iX", "cov", "alpha", "eta"):
        assert rel(a[key], b[key]) < TOL, key
    have = [c for c in range(len(g._loco_start)) if g._loco_start[c] != -1]
    for c in (have[1], have[-1]):                      # leave-one-chromosome-out: getCoefficients_LOCO
        r.set_loco_chromosome(c); g.setStartEndIndex(g._loco_start[c], g._loco_end[c], c)
        a, b = g.getCoefficients(Y, X, w, tau, 500, 1e-5, loco=True), r.getCoefficients(Y, X, w, tau, 500, 1e-5, loco=True)
        for key in ("Sigma_iY", "Sigma_iX", "cov", "alpha", "eta"):
            assert rel(a[key], b[key]) < TOL, (c, key)


@pytest.mark.parametrize("cvcut", [0.0025, 3e-4])
def test_ai_score_and_tau_update_vs_reference_code(trio, cvcut):
    g, o, r, probes, yb, yq, X = trio
    W = np.random.default_rng(12).uniform(0.05, 0.25, size=o.N); tau = np.array([1.0, 0.3])
    c = r.getCoefficients(yb, X, W, tau, 500, 1e-5)
    a = g.getAIScore(yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, cvcut, probes.fresh())
    b = r.getAIScore(yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, cvcut)
    assert a["nrun_used"] == b["nrun_used"]
    for key in ("YPAPY", "Trace", "AI", "PY"):
        assert rel(a[key], b[key]) < TOL, key
    tg = g.fitglmmaiRPCG(yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, cvcut, probes.fresh())["tau"]
    assert rel(tg, r.fitglmmaiRPCG(yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, cvcut)) < TOL
    # quantitative trait
    Wq = np.ones(o.N); tq = np.array([0.6, 0.3])
    cq = r.getCoefficients(yq, X, Wq, tq, 500, 1e-5)
    a = g.getAIScore_q(yq, X, Wq, tq, cq["Sigma_iY"], cq["Sigma_iX"], cq["cov"], 30, 500, 1e-5, cvcut, probes.fresh())
    b = r.getAIScore_q(yq, X, Wq, tq, cq["Sigma_iY"], cq["Sigma_iX"], cq["cov"], 30, 500, 1e-5, cvcut)
    assert a["nrun_used"] == b["nrun_used"]
    for key in ("YPAPY", "YPA0PY", "Trace", "AI", "PY"):
        assert rel(a[key], b[key]) < TOL, key
    tg = g.fitglmmaiRPCG_q(yq, X, Wq, tq, cq["Sigma_iY"], cq["Sigma_iX"], cq["cov"], 30, 500, 1e-5, 0.02, cvcut, probes.fresh())["tau"]
    assert rel(tg, r.fitglmmaiRPCG(yq, X, Wq, tq, cq["Sigma_iY"], cq["Sigma_iX"], cq["cov"], 30, 500, 1e-5, 0.02, cvcut, quant=True)) < TOL


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("native", [False, True])
def test_whole_fit_vs_reference_code(trio, trait, native):
    """tau and alpha of the whole null-GLMM fit: the GPU library (per-export driver and one-call driver) against the R-level loop
    run over the reference's compiled C++ exports."""
    from oracle import oracle as O
    from saige_gpu_b200 import step1
    g, o, r, probes, yb, yq, X = trio
    fam_o, fam_g, y = (O.Binomial, step1.Binomial, yb) if trait == "binary" else (O.Gaussian, step1.Gaussian, yq)
    want = R.fit_through_reference(o, r, O.glm_fit(y, X, fam_o), probes.U, trait)
    got = step1.glmmkin_ai_PCG(g, step1.glm_fit(y, X, fam_g), probes, trait=trait, native_loops=native)
    g.setProbeStreamFixed(False)
    assert got["converged"] == want["converged"]
    assert rel(got["theta"], want["theta"]) < TOL
    assert rel(got["coefficients"], want["coefficients"]) < TOL
    assert rel(got["fitted_values"], want["fitted_values"]) < TOL


@pytest.mark.skipif(not R.available("cpu"), reason="oracle/_ref/libfg_refcpu.so not built")
def test_ingest_and_product_vs_reference_cpu_path(tmp_path, grm10k):
    """The CUDA library against the reference's own CPU path (oracle/_ref/libfg_refcpu.so: genoClass + parallelCrossProd compiled
    unmodified), both reading the same PLINK files: QC mask / MAC / genotypes / variance-ratio hold-out bit-exact, allele
    frequencies equal as the floats the reference stores, products to the reference's fp32 accuracy."""
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200
    from tests_support import write_plink
    N0, M0 = 1237, 3000
    bed = O.synth_bed(N0, M0, seed=77, miss_rate=0.02)
    rng = np.random.default_rng(4)
    keep = np.sort(rng.choice(N0, size=1001, replace=False))
    sub = rng.permutation(keep) + 1
    ind = np.zeros(N0, np.uint8); ind[keep] = 1
    vr = np.unique(rng.integers(0, M0, size=200))
    prefix = str(tmp_path / "cohort")
    write_plink(prefix, bed, N0, M0)
    cases = [(prefix, sub, ind, 0.06, 0.03, True, vr), (grm10k["prefix"], np.arange(1, 1001), np.ones(1000, np.uint8), 0.01, 0.15, False, None)]
    for pre, s_, i_, maf, miss, isvr, vridx in cases:
        r = R.RefCPU()
        r.setgeno(pre + ".bed", pre + ".bim", pre + ".fam", s_, i_, minMAF=maf, maxMissing=miss, isVarRatio=isvr, vr_rand_idx=vridx)
        g = SaigeB200(device=0)
        try:
            g.setminMAFforGRM(maf); g.setmaxMissingRateforGRM(miss); g.setminMAC_VarianceRatio(20, -1, isvr)
            g.setgeno(pre + ".bed", pre + ".bim", pre + ".fam", s_, i_, vr_rand_idx=vridx)
            assert (g.N, g.M, g.Mvr) == (r.N, r.M, r.Mvr)
            assert np.array_equal(g.getQCdMarkerIndex(), r.getQCdMarkerIndex())
            assert np.array_equal(g.getMACVec(), r.getMACVec())
            assert np.array_equal(g.getAlleleFreqVec().astype(np.float32), r.getAlleleFreqVec().astype(np.float32))
            for i in range(0, g.M, 211):
                assert np.array_equal(g.Get_OneSNP_Geno(i), r.Get_OneSNP_Geno(i)), i
            if isvr:
                assert np.array_equal(g.getIndexVec_forVarRatio(), r.getIndexVec_forVarRatio())
                assert np.array_equal(g.getMACVec_forVarRatio(), r.getMACVec_forVarRatio())
                for i in range(0, g.Mvr, 9):
                    assert np.array_equal(g.Get_OneSNP_Geno_forVarRatio(i), r.Get_OneSNP_Geno(i, vr=True)), i
            b = rng.normal(size=g.N)
            assert rel(g.getCrossprodMatAndKin(b), r.getCrossprodMatAndKin(b)) < 2e-6
            assert rel(g.get_DiagofKin() * g.M, r.Get_Diagof_StdGeno()) < 1e-5
        finally:
            g.close()
