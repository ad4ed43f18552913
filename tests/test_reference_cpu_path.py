"""The oracle (and, through it, the CUDA library) against the reference's OWN CPU path, end to end.

oracle/_ref/libfg_refcpu.so = the reference's genoClass (PLINK reader, QC, best-guess imputation, re-pack, standardised
genotypes, diagonals: SAIGE_fitGLMM_fast.cpp:37-1183), its OpenMP marker loop parallelCrossProd[_LOCO] (:1576-1851), the exports
that configure them and the solver layer (:2322-3662), cut out of the reference tree at build time and compiled UNMODIFIED, in
the reference's own precision (fp32), as a CPU-only build of the reference takes them.  Files in, tau out: nothing below comes
from oracle/ except the thing being checked.

Bars: genotype decode, allele counts, MAC, the QC mask, the variance-ratio hold-out: bit-exact.  Products, diagonals, PCG, tau:
the oracle computes in fp64, the reference in fp32 => agreement to float accuracy (the tolerances are written at each check)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_solver as R
from tests_support import write_plink

pytestmark = pytest.mark.skipif(not R.available("cpu"), reason="oracle/_ref/libfg_refcpu.so not built (python __graft_entry__.py)")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def bundled(grm10k):
    """Reference and oracle both loaded with the bundled 1000 x 10k set, MAF >= 0.01, missing rate <= 0.15."""
    N0, M0, p = grm10k["N0"], grm10k["M0"], grm10k["prefix"]
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    U = np.random.default_rng(200).integers(0, 2, size=(N0, 130)) * 2.0 - 1.0
    r = R.RefCPU(U)
    r.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8), minMAF=0.01, maxMissing=0.15)
    return o, r, U


def test_ingest_qc_and_decode_are_bit_exact(bundled):
    o, r, _ = bundled
    assert (r.N, r.M, r.M0) == (o.N, o.M, o.M0) == (1000, 9650, 10000)
    assert np.array_equal(r.getQCdMarkerIndex(), o.qc_mask)
    assert np.array_equal(r.getMACVec(), o.MACVec)
    # the reference keeps f and 1/sd as floats; the oracle's float getters are those very numbers
    assert np.array_equal(r.getAlleleFreqVec().astype(np.float32), o.alleleFreqVec)
    assert np.array_equal(r.getInvStdVec().astype(np.float32), o.invstdvVec)
    for i in range(0, o.M, 193):
        assert np.array_equal(r.Get_OneSNP_Geno(i), o.Get_OneSNP_Geno(i)), i
    for i in (0, 4321, o.M - 1):
        assert rel(r.Get_OneSNP_StdGeno(i), o.Get_OneSNP_StdGeno(i)) < 1e-6


def test_product_and_diagonal(bundled, grm10k):
    o, r, _ = bundled
    # the oracle's fp32 reference-order mode (what bench.py timed as the "port" baseline) against the loop it restates
    o32 = O.OracleGeno(mode=O.REF32); o32.minMAF, o32.maxMissing = 0.01, 0.15
    o32.setgeno(grm10k["bed"], grm10k["N0"], grm10k["M0"], np.arange(1, o.N + 1), np.ones(o.N, np.uint8))
    rng = np.random.default_rng(2)
    for _ in range(3):
        b = rng.normal(size=o.N)
        y = r.getCrossprodMatAndKin(b)
        assert rel(y, o.getCrossprodMatAndKin(b)) < 2e-6          # fp32 marker loop vs fp64
        assert rel(y, o32.getCrossprodMatAndKin(b)) < 2e-6        # fp32 vs fp32: the same sums in another thread order
    assert rel(r.Get_Diagof_StdGeno(), o.Get_Diagof_StdGeno()) < 1e-5


def test_loco_product_and_diagonal(bundled, grm10k):
    o, r, _ = bundled
    chrq = np.array([int(c) for c in grm10k["chrs"]])[o.qc_mask]
    LOCO, s, e = O.updateChrStartEndIndexVec(chrq)
    o.setStartEndIndexVec(s, e); o.set_Diagof_StdGeno_LOCO()
    r.setStartEndIndexVec(s, e); r.set_Diagof_StdGeno_LOCO()
    rng = np.random.default_rng(3)
    b = rng.normal(size=o.N); w = rng.uniform(0.05, 0.25, size=o.N); tau = np.array([1.0, 0.4])
    have = [c for c in range(len(s)) if s[c] != -1]
    for c in (have[0], have[len(have) // 2], have[-1]):
        o.setStartEndIndex(s[c], e[c], c); r.setStartEndIndex(s[c], e[c], c)
        # the reference subtracts two UN-normalised fp32 sums (all markers - the chromosome's, FG.cpp:1839-1848): float cancellation
        assert rel(r.getCrossprodMatAndKin_LOCO(b), o.getCrossprodMatAndKin_LOCO(b)) < 1e-4
        assert rel(r.getDiagOfSigma(w, tau, loco=True), o.getDiagOfSigma(w, tau, loco=True)) < 1e-4
        x, it = r.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, loco=True, return_iter=True)
        xo, ito = o.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, loco=True, return_iter=True)
        assert abs(it - ito) <= 1 and rel(x, xo) < 1e-3


def test_missing_calls_sample_subset_and_variance_ratio_holdout(tmp_path):
    """Everything the QC path has: 2 % missing calls (best-guess imputation), a phenotyped subset in shuffled order, the missing-rate
    and MAF filters, and the variance-ratio hold-out drawn from a supplied index set (arma::randi in the reference, FG.cpp:866-868)."""
    N0, M0 = 1237, 3000
    bed = O.synth_bed(N0, M0, seed=77, miss_rate=0.02)
    rng = np.random.default_rng(4)
    keep = np.sort(rng.choice(N0, size=1001, replace=False))
    sub = rng.permutation(keep) + 1
    ind = np.zeros(N0, np.uint8); ind[keep] = 1
    vr = np.unique(rng.integers(0, M0, size=200))
    o = O.OracleGeno(); o.minMAF, o.maxMissing, o.isVarRatio = 0.06, 0.03, True
    o.setgeno(bed, N0, M0, sub, ind, vr_rand_idx=vr)
    prefix = str(tmp_path / "cohort")
    write_plink(prefix, bed, N0, M0)
    r = R.RefCPU()
    r.setgeno(prefix + ".bed", prefix + ".bim", prefix + ".fam", sub, ind, minMAF=0.06, maxMissing=0.03, isVarRatio=True,
              minMACvr=20, maxMACvr=-1, vr_rand_idx=vr)
    assert (r.N, r.M, r.Mvr) == (o.N, o.M, o.Mvr) and o.Mvr > 50 and 0 < o.M < M0 - o.Mvr
    assert np.array_equal(r.getQCdMarkerIndex(), o.qc_mask)
    assert np.array_equal(r.getMACVec(), o.MACVec)
    assert np.array_equal(r.getIndexVec_forVarRatio(), o.markerIndexVec_forVarRatio)
    assert np.array_equal(r.getMACVec_forVarRatio(), o.MACVec_forVarRatio)
    for i in range(0, o.M, 97):
        assert np.array_equal(r.Get_OneSNP_Geno(i), o.Get_OneSNP_Geno(i)), i
    for i in range(0, o.Mvr, 5):
        assert np.array_equal(r.Get_OneSNP_Geno(i, vr=True), o.Get_OneSNP_Geno(i, vr=True)), i
    b = rng.normal(size=o.N)
    assert rel(r.getCrossprodMatAndKin(b), o.getCrossprodMatAndKin(b)) < 2e-6


def test_whole_fit_inside_the_reference(bundled, golden_dir):
    """BASELINE config 1 through the reference's own code from the .bed file to tau (fp32), against the fp64 oracle."""
    o, r, U = bundled
    rows = [l.split() for l in open(os.path.join(golden_dir, "pheno_1000samples.txt")).readlines()]
    col = {h: i for i, h in enumerate(rows[0])}
    yb = np.array([float(x[col["y_binary"]]) for x in rows[1:]])
    X = np.column_stack([np.ones(len(yb)), [float(x[col["x1"]]) for x in rows[1:]], [float(x[col["x2"]]) for x in rows[1:]]])
    fit0 = O.glm_fit(yb, X, O.Binomial)
    want = O.glmmkin_ai_PCG(o, fit0, (0, 0), U, trait="binary")
    got = R.fit_through_reference(o, r, fit0, U, "binary")
    assert got["converged"] == want["converged"]
    assert rel(got["theta"], want["theta"]) < 5e-3 and rel(got["coefficients"], want["coefficients"]) < 5e-3
    ref_alpha = np.array([-2.97337569, 0.7511719, 0.91698671])        # extdata/output/example.rda, the reference's fit of this cohort
    assert np.max(np.abs(got["coefficients"] - ref_alpha) / np.abs(ref_alpha)) < 2e-2
