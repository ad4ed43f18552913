"""The oracle's solver / AI-REML layer against the REFERENCE'S OWN CODE.

oracle/_ref/libfg_ref64.so holds getDiagOfSigma, getCrossprod, getPCG1ofSigmaAndVector, getCoefficients, GetTrace[_q],
getAIScore[_q], fitglmmaiRPCG[_q], getSigma_X / _G, calCV and their _LOCO twins exactly as they stand in
/root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp:2322-3662 (cut out of the reference tree at build time by
oracle/ref_fg/extract_ref.py, compiled against oracle/ref_fg/mini_arma.h with `float` read as `double`); libfg_ref32.so is the
same text in the reference's shipped precision.  They take the GRM product and diagonal from the oracle (pinned elsewhere:
.frq / .acount fixtures, the compiled gpuSymMatMult), so what these tests pin is everything above the product -- the part of
oracle/oracle.py that the GPU library's tau / alpha / variance-ratio parity (<= 1e-6) rests on.

Tolerances: fp64 build of the reference vs the oracle: identical PCG iteration counts, results <= 1e-9 (same algorithm, same
precision, different summation order); fp32 build: <= 5e-3 on tau / alpha (float PCG with an absolute 1e-5 stop on ||r||^2)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_solver as R

pytestmark = pytest.mark.skipif(not R.available(64), reason="oracle/_ref/libfg_ref64.so not built (python __graft_entry__.py)")

TOL64 = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _pheno(golden_dir):
    rows = [l.split() for l in open(os.path.join(golden_dir, "pheno_1000samples.txt")).readlines()]
    hdr, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    yb = np.array([float(r[col["y_binary"]]) for r in rows])
    yq = np.array([float(r[col["y_quantitative"]]) for r in rows])
    X = np.column_stack([np.ones(len(rows)), [float(r[col["x1"]]) for r in rows], [float(r[col["x2"]]) for r in rows]])
    return yb, yq, X


@pytest.fixture(scope="module")
def setup(grm10k, golden_dir):
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno()
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    U = np.random.default_rng(200).integers(0, 2, size=(o.N, 200)) * 2.0 - 1.0
    yb, yq, X = _pheno(golden_dir)
    return o, R.RefSolver(o, 64, U), U, yb, yq, X


def test_calcv_and_diag_of_sigma(setup):
    o, r, *_ = setup
    rng = np.random.default_rng(1)
    x = rng.uniform(0.5, 1.5, size=37)
    assert abs(r.calCV(x) - O.calCV(x)) < 1e-15
    w = rng.uniform(0.02, 0.25, size=o.N); tau = np.array([1.0, 0.7])
    assert rel(r.getDiagOfSigma(w, tau), o.getDiagOfSigma(w, tau)) < 1e-14
    wbig = np.full(o.N, 1e9)                         # the 1e-4 floor (FG.cpp:2355-2357)
    assert np.array_equal(r.getDiagOfSigma(wbig, np.array([1.0, 0.0])), o.getDiagOfSigma(wbig, np.array([1.0, 0.0])))


@pytest.mark.parametrize("tau", [(1.0, 0.4), (1.0, 2.5), (0.3, 0.05), (1.0, 0.0)])
def test_pcg_iterations_and_solution(setup, tau):
    o, r, *_ = setup
    rng = np.random.default_rng(int(100 * tau[1]) + 3)
    w = rng.uniform(0.02, 0.25, size=o.N); tau = np.array(tau)
    for _ in range(3):
        b = rng.normal(size=o.N)
        x, it = r.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, return_iter=True)
        xo, ito = o.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, return_iter=True)
        assert it == ito and rel(x, xo) < TOL64
    # maxiterPCG reached: the same truncated iterate
    x, it = r.getPCG1ofSigmaAndVector(w, tau, b, 2, 1e-14, return_iter=True)
    xo, ito = o.getPCG1ofSigmaAndVector(w, tau, b, 2, 1e-14, return_iter=True)
    assert it == ito and rel(x, xo) < TOL64


def test_get_coefficients_and_sigma_x_g(setup):
    o, r, U, yb, yq, X = setup
    rng = np.random.default_rng(5)
    W = rng.uniform(0.05, 0.25, size=o.N); tau = np.array([1.0, 0.3])
    Y = rng.normal(size=o.N) + 2.0
    a, b = r.getCoefficients(Y, X, W, tau, 500, 1e-5), O.getCoefficients(o, Y, X, W, tau, 500, 1e-5)
    for key in ("Sigma_iY", "Sigma_iX", "cov", "alpha", "eta"):
        assert rel(a[key], b[key]) < TOL64, key
    assert rel(r.getSigma_X(W, tau, X, 500, 1e-5), O.getSigma_X(o, W, tau, X, 500, 1e-5)) < TOL64
    assert rel(r.getSigma_G(W, tau, Y, 500, 1e-5), O.getSigma_G(o, W, tau, Y, 500, 1e-5)) < TOL64


@pytest.mark.parametrize("cvcut", [0.0025, 3e-4])
def test_ai_score_and_tau_update_binary(setup, cvcut):
    """getAIScore + fitglmmaiRPCG with the same probes; the tight CV cut-off forces +10 retry batches (FG.cpp:3148-3153)."""
    o, r, U, yb, yq, X = setup
    W = np.random.default_rng(6).uniform(0.05, 0.25, size=o.N); tau = np.array([1.0, 0.3])
    c = O.getCoefficients(o, yb, X, W, tau, 500, 1e-5)
    draws = O.make_draw(U)
    a = r.getAIScore(yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, cvcut)
    b = O.getAIScore(o, yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, cvcut, draws())
    assert a["nrun_used"] == b["nrun_used"] and (cvcut > 1e-3 or a["nrun_used"] > 30)
    for key in ("YPAPY", "Trace", "AI", "PY"):
        assert rel(a[key], b[key]) < TOL64, key
    ta = r.fitglmmaiRPCG(yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, cvcut)
    tb = O.fitglmmaiRPCG(o, yb, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, cvcut, draws())
    assert rel(ta, tb) < TOL64
    # the step-halving branch: a start value whose Newton step would go negative (FG.cpp:3330-3334)
    t0 = np.array([1.0, 1e-3])
    ta = r.fitglmmaiRPCG(yb, X, W, t0, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, 0.0025)
    tb = O.fitglmmaiRPCG(o, yb, X, W, t0, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, 0.0025, draws())
    assert np.allclose(ta, tb, rtol=TOL64, atol=1e-15)


def test_ai_score_and_tau_update_quantitative(setup):
    o, r, U, yb, yq, X = setup
    W = np.ones(o.N); tau = np.array([0.6, 0.3])
    c = O.getCoefficients(o, yq, X, W, tau, 500, 1e-5)
    draws = O.make_draw(U)
    a = r.getAIScore_q(yq, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.0025)
    b = O.getAIScore_q(o, yq, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.0025, draws())
    assert a["nrun_used"] == b["nrun_used"]
    for key in ("YPAPY", "YPA0PY", "Trace", "AI", "PY"):
        assert rel(a[key], b[key]) < TOL64, key
    ta = r.fitglmmaiRPCG(yq, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, 0.0025, quant=True)
    tb = O.fitglmmaiRPCG_q(o, yq, X, W, tau, c["Sigma_iY"], c["Sigma_iX"], c["cov"], 30, 500, 1e-5, 0.02, 0.0025, draws())
    assert rel(ta, tb) < TOL64


def test_loco_coefficients(setup, grm10k):
    """getCoefficients_LOCO / getPCG1ofSigmaAndVector_LOCO / getDiagOfSigma_LOCO on three chromosomes."""
    o, r, U, yb, yq, X = setup
    chrq = np.array([int(c) for c in grm10k["chrs"]])[o.qc_mask]
    LOCO, s, e = O.updateChrStartEndIndexVec(chrq)
    assert LOCO
    o.setStartEndIndexVec(s, e); o.set_Diagof_StdGeno_LOCO()
    rng = np.random.default_rng(8)
    W = rng.uniform(0.05, 0.25, size=o.N); tau = np.array([1.0, 0.35]); Y = rng.normal(size=o.N)
    have = [c for c in range(len(s)) if s[c] != -1]
    for c in (have[0], have[len(have) // 2], have[-1]):
        r.set_loco_chromosome(c)                     # also sets the oracle's current chromosome
        assert rel(r.getDiagOfSigma(W, tau, loco=True), o.getDiagOfSigma(W, tau, loco=True)) < 1e-14
        x, it = r.getPCG1ofSigmaAndVector(W, tau, Y, 500, 1e-5, loco=True, return_iter=True)
        xo, ito = o.getPCG1ofSigmaAndVector(W, tau, Y, 500, 1e-5, loco=True, return_iter=True)
        assert it == ito and rel(x, xo) < TOL64
        a, b = r.getCoefficients(Y, X, W, tau, 500, 1e-5, loco=True), O.getCoefficients(o, Y, X, W, tau, 500, 1e-5, loco=True)
        for key in ("Sigma_iY", "Sigma_iX", "cov", "alpha", "eta"):
            assert rel(a[key], b[key]) < TOL64, (c, key)


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_whole_fit_through_the_reference_code(setup, trait):
    """Whole null-GLMM fit (BASELINE config 1's cohort): tau and alpha from the reference's compiled solver layer vs the oracle's
    restatement of it -- the chain reference code -> oracle (here, 1e-9) -> GPU library (tests/test_gpu_parity.py, 1e-6)."""
    o, r, U, yb, yq, X = setup
    fam, y = (O.Binomial, yb) if trait == "binary" else (O.Gaussian, yq)
    fit0 = O.glm_fit(y, X, fam)
    want = O.glmmkin_ai_PCG(o, fit0, (0, 0), U, trait=trait)
    got = R.fit_through_reference(o, r, fit0, U, trait)
    assert got["converged"] == want["converged"]
    assert rel(got["theta"], want["theta"]) < TOL64
    assert rel(got["coefficients"], want["coefficients"]) < TOL64
    assert rel(got["fitted_values"], want["fitted_values"]) < TOL64
    if trait == "binary":
        # ... and against the reference's own bundled result for this cohort (extdata/output/example.rda): alpha is RNG-free up
        # to the PCG tolerance (tau is not: it carries the Monte-Carlo error of 30 probes, R's stream there, another one here)
        ref_alpha = np.array([-2.97337569, 0.7511719, 0.91698671])
        assert np.max(np.abs(got["coefficients"] - ref_alpha) / np.abs(ref_alpha)) < 2e-2


@pytest.mark.skipif(not R.available(32), reason="oracle/_ref/libfg_ref32.so not built")
def test_shipped_precision_build_agrees_to_float_accuracy(setup):
    """The same reference text compiled as shipped (fp32 vectors and scalars): the fp64 oracle reproduces its fit to float accuracy."""
    o, r64, U, yb, yq, X = setup
    r32 = R.RefSolver(o, 32, U)
    rng = np.random.default_rng(9)
    w = rng.uniform(0.05, 0.25, size=o.N); tau = np.array([1.0, 0.4]); b = rng.normal(size=o.N)
    x, it = r32.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, return_iter=True)
    xo, ito = o.getPCG1ofSigmaAndVector(w, tau, b, 500, 1e-5, return_iter=True)
    assert abs(it - ito) <= 1 and rel(x, xo) < 1e-4
    fit0 = O.glm_fit(yb, X, O.Binomial)
    want = O.glmmkin_ai_PCG(o, fit0, (0, 0), U, trait="binary")
    got = R.fit_through_reference(o, r32, fit0, U, "binary")
    assert rel(got["theta"], want["theta"]) < 5e-3 and rel(got["coefficients"], want["coefficients"]) < 5e-3
