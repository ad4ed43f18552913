"""The driver loops of SAIGE_fitGLMM_fast.R run inside the library (sgb_get_coef, sgb_get_coef_loco_all,
sgb_variance_ratio_markers, sgb_set_probe_stream_fixed) against the same loops run through the per-export mirror
(saige_gpu_b200/step1.py) and against the CPU oracle.

Tolerances: the native loops do the same arithmetic as the mirror except for the device's exp() in the IRLS update
(<= 1 ulp from numpy's), so native vs mirror is held to 1e-9; native vs oracle to the north-star 1e-6."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_SAME = 1e-9
TOL_FIT = 1e-6


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
def loco_pair(chr22):
    """GPU context and oracle loaded with the bundled 22-chromosome set, LOCO ranges set on both."""
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200, step1
    N0, M0 = chr22["N0"], chr22["M0"]
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(chr22["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    g = SaigeB200(device=0)
    g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15); g.setminMAC_VarianceRatio(20, -1, False)
    g.setgeno_mem(chr22["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    chrq = chr22["chrs"][o.qc_mask]
    LOCO, s, e = O.updateChrStartEndIndexVec(chrq)
    o.setStartEndIndexVec(s, e)
    assert step1.set_loco_ranges(g, chrq) == LOCO and LOCO
    yield g, o
    g.close()


@pytest.mark.parametrize("family", ["binomial", "gaussian"])
@pytest.mark.parametrize("loco", [False, True])
def test_get_coef_native_equals_mirror(loco_pair, golden_dir, family, loco):
    from saige_gpu_b200 import step1
    g, o = loco_pair
    yb, yq, X = _pheno(golden_dir)
    fam, y = (step1.Binomial, yb) if family == "binomial" else (step1.Gaussian, yq)
    fit0 = step1.glm_fit(y, X, fam)
    rng = np.random.default_rng(5)
    offset = rng.normal(scale=0.05, size=len(y))                 # a non-trivial offset: eta includes it, Y does not
    eta0 = fit0["eta"] + offset
    tau = np.array([1.0, 0.35])
    if loco:
        g.set_Diagof_StdGeno_LOCO()
        g.setStartEndIndex(g._loco_start[4], g._loco_end[4], 4)
    args = (y, X, tau, fam, fit0["coef"] * 0.5, eta0, offset, 500, 1e-5, 20)
    m = step1.Get_Coef(g, *args, loco=loco)
    n = step1.Get_Coef(g, *args, loco=loco, native=True)
    assert n["n_iter"] >= 1
    for key in ("Y", "alpha", "eta", "W", "cov", "sqrtW", "Sigma_iY", "Sigma_iX", "mu"):
        assert rel(n[key], m[key]) < TOL_SAME, key


def test_get_coef_maxiter_one_and_bad_arguments(loco_pair, golden_dir):
    from saige_gpu_b200 import step1, SaigeB200Error
    g, _ = loco_pair
    yb, _, X = _pheno(golden_dir)
    fit0 = step1.glm_fit(yb, X, step1.Binomial)
    off = np.zeros(len(yb)); tau = np.array([1.0, 0.2])
    a = (yb, X, tau, step1.Binomial, np.zeros(3), fit0["eta"], off, 500, 1e-5)
    m = step1.Get_Coef(g, *a, 1)
    n = step1.Get_Coef(g, *a, 1, native=True)
    assert n["n_iter"] == 1 and rel(n["alpha"], m["alpha"]) < TOL_SAME and rel(n["Y"], m["Y"]) < TOL_SAME
    with pytest.raises(SaigeB200Error):
        g.Get_Coef(yb, X, tau, "binomial", np.zeros(3), fit0["eta"], off, 500, 1e-5, 0)
    with pytest.raises(KeyError):
        g.Get_Coef(yb, X, tau, "poisson", np.zeros(3), fit0["eta"], off, 500, 1e-5, 5)


def test_binomial_link_clamps_like_R(loco_pair, golden_dir):
    """R's logit_linkinv / logit_mu_eta clamp at |eta| > 30 (stats/src/family.c); the IRLS kernel follows them."""
    g, _ = loco_pair
    yb, _, X = _pheno(golden_dir)
    N = len(yb)
    eta0 = np.linspace(-40.0, 40.0, N)
    r = g.Get_Coef(yb, X, np.array([1.0, 0.0]), "binomial", np.zeros(3), eta0, np.zeros(N), 500, 1e-5, 1)
    # with tau1 = 0 the first solve is diagonal; what is checked here is the start-up IRLS step through its outputs' finiteness
    assert np.all(np.isfinite(r["Y"])) and np.all(np.isfinite(r["W"])) and np.all(r["W"] > 0)
    eps = np.finfo(float).eps
    # one update from the returned eta reproduces mu / W with R's formulas
    eta = r["eta"]
    e = np.exp(eta)
    tmp = np.where(eta < -30, eps, np.where(eta > 30, 1 / eps, e))
    mu = tmp / (1 + tmp)
    me = np.where(np.abs(eta) > 30, eps, e / (1 + e) ** 2)
    assert rel(r["mu"], mu) < 1e-14
    assert rel(r["W"], (me / np.sqrt(mu * (1 - mu))) ** 2) < 1e-12


@pytest.mark.parametrize("mode", ["calls", True])
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_full_fit_with_loco_native_loops(loco_pair, golden_dir, trait, mode):
    """Whole fit incl. the 22 leave-one-chromosome-out refits: native loops (one library call per R loop / the whole R function as
    one call) == mirror (1e-9) == oracle (1e-6), PCG work equal."""
    from oracle import oracle as O
    from saige_gpu_b200 import step1
    g, o = loco_pair
    yb, yq, X = _pheno(golden_dir)
    fam_o, fam_g, y = (O.Binomial, step1.Binomial, yb) if trait == "binary" else (O.Gaussian, step1.Gaussian, yq)
    probes = step1.ProbeStream(o.N, nmax=130, seed=200)
    fit0 = step1.glm_fit(y, X, fam_g)
    g.reset_counters()
    mm = step1.glmmkin_ai_PCG(g, fit0, probes, trait=trait, LOCO=True)
    cm = g.counters()
    g.reset_counters()
    mn = step1.glmmkin_ai_PCG(g, fit0, probes, trait=trait, LOCO=True, native_loops=mode)
    cn = g.counters()
    g.setProbeStreamFixed(False)
    assert cn["n_pcg_iterations"] == cm["n_pcg_iterations"] and cn["n_pcg_solves"] == cm["n_pcg_solves"]
    assert cn["n_probe_batches_resident"] >= 1 and cm["n_probe_batches_resident"] == 0
    assert cn["bytes_h2d"] < 0.5 * cm["bytes_h2d"]
    assert mn["converged"] == mm["converged"]
    if mode is True:
        assert cn["bytes_h2d"] < 0.1 * cm["bytes_h2d"] and mn["n_outer"] == len(mm["tau_path"]) - 1
    mo = O.glmmkin_ai_PCG(o, O.glm_fit(y, X, fam_o), (0, 0), probes.U, trait=trait, LOCO=True)
    for key in ("theta", "coefficients", "linear_predictors", "fitted_values", "Y", "cov"):
        assert rel(mn[key], mm[key]) < TOL_SAME, key
        assert rel(mn[key], mo[key]) < TOL_FIT, key
    assert len(mn["LOCOResult"]) == len(mo["LOCOResult"]) == 22
    for c, (rn, rm, ro) in enumerate(zip(mn["LOCOResult"], mm["LOCOResult"], mo["LOCOResult"])):
        assert rn["isLOCO"] == rm["isLOCO"] == ro["isLOCO"]
        if not rn["isLOCO"]:
            continue
        for key in ("coefficients", "linear_predictors", "fitted_values", "Y", "residuals", "cov"):
            assert rel(rn[key], rm[key]) < TOL_SAME, (c, key)
            assert rel(rn[key], ro[key]) < TOL_FIT, (c, key)
        assert rel(rn["obj_noK"]["XVX_inv_XV"], rm["obj_noK"]["XVX_inv_XV"]) < TOL_SAME


def test_probe_stream_fixed_with_cv_retries(loco_pair, golden_dir):
    """A trace CV cut-off that forces +10 retry batches: the resident first batch must leave the callback's stream where the
    reference's would be, i.e. the result equals the all-callback run bit for bit."""
    from saige_gpu_b200 import step1
    g, o = loco_pair
    yb, _, X = _pheno(golden_dir)
    fit0 = step1.glm_fit(yb, X, step1.Binomial)
    tau = np.array([1.0, 0.3])
    rc = step1.Get_Coef(g, yb, X, tau, step1.Binomial, fit0["coef"], fit0["eta"], np.zeros(o.N), 500, 1e-5, 20)
    probes = step1.ProbeStream(o.N, nmax=1000, seed=200)
    g.setProbeStreamFixed(False)
    for cutoff in (1e-3, 7e-4, 5e-4, 3e-4, 2e-4):      # the CV of n trace terms falls like 1/n: find a cut-off that needs retries
        a = (rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], 30, 500, 1e-5, cutoff)
        want = g.getAIScore(*a, probes.fresh())
        if want["nrun_used"] > 30:
            break
    assert 30 < want["nrun_used"] < 1000
    g.setProbeStreamFixed(True)
    g.reset_counters()
    first = g.getAIScore(*a, probes.fresh())          # the cache already holds this batch from the run above
    again = g.getAIScore(*a, probes.fresh())
    g.setProbeStreamFixed(False)
    assert g.counters()["n_probe_batches_resident"] == 2
    for got in (first, again):
        assert got["nrun_used"] == want["nrun_used"]
        assert got["Trace"] == want["Trace"] and got["YPAPY"] == want["YPAPY"] and got["AI"] == want["AI"]


def test_one_call_fit_with_cv_retries_equals_mirror(loco_pair, golden_dir):
    """A trace CV cut-off that needs +10 retry batches in every AI step: inside the one-call fit the probe stream restarts at
    every trace estimate (count = 0 announcement) and the resident first batch is replayed before a retry, so the fit equals the
    mirror's, which re-seeds per call."""
    from saige_gpu_b200 import step1
    g, o = loco_pair
    yb, _, X = _pheno(golden_dir)
    fit0 = step1.glm_fit(yb, X, step1.Binomial)
    probes = step1.ProbeStream(o.N, nmax=1000, seed=200)
    g.reset_counters()
    m0 = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", traceCVcutoff=1.0)         # no estimate needs more than nrun probes
    base = g.counters()["n_pcg_solves"] / len(m0["tau_path"])
    for cutoff in (1e-3, 7e-4, 5e-4, 3e-4):
        g.reset_counters()
        mm = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", traceCVcutoff=cutoff)
        cm = g.counters()
        if cm["n_pcg_solves"] / len(mm["tau_path"]) >= base + 10:                         # columns solved per outer step grew by a retry batch
            break
    else:
        pytest.fail("no CV cut-off produced retry batches")
    g.reset_counters()
    mn = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", traceCVcutoff=cutoff, native_loops=True)
    cn = g.counters()
    g.setProbeStreamFixed(False)
    assert cn["n_pcg_solves"] == cm["n_pcg_solves"] and cn["n_pcg_iterations"] == cm["n_pcg_iterations"]
    assert rel(mn["theta"], mm["theta"]) < TOL_SAME and rel(mn["coefficients"], mm["coefficients"]) < TOL_SAME


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_variance_ratio_markers_native(loco_pair, golden_dir, trait):
    from oracle import oracle as O
    from saige_gpu_b200 import step1
    g, o = loco_pair
    yb, yq, X = _pheno(golden_dir)
    fam_o, fam_g, y = (O.Binomial, step1.Binomial, yb) if trait == "binary" else (O.Gaussian, step1.Gaussian, yq)
    probes = step1.ProbeStream(o.N, nmax=130, seed=200)
    mg = step1.glmmkin_ai_PCG(g, step1.glm_fit(y, X, fam_g), probes, trait=trait)
    mo = O.glmmkin_ai_PCG(o, O.glm_fit(y, X, fam_o), (0, 0), probes.U, trait=trait)
    mac = np.minimum(o.ACVec, 2 * o.N - o.ACVec)
    order = np.random.default_rng(1).permutation(np.nonzero(mac >= 20)[0])[:200]
    # a tight CV cut-off: several +10 rounds, i.e. several batches of different widths
    vm, lm = step1.extractVarianceRatio(g, mg, fam_g, order, ratioCVcutoff=0.0005)
    vn, ln = step1.extractVarianceRatio(g, mg, fam_g, order, ratioCVcutoff=0.0005, native_loops=True)
    vo, lo = O.extractVarianceRatio(o, mo, fam_o, order, ratioCVcutoff=0.0005)
    assert len(ln) == len(lm) == len(lo) and len(ln) >= 30
    assert rel(ln, lm) < TOL_SAME and rel(vn, vm) < TOL_SAME
    assert rel(ln, lo) < TOL_FIT and rel(vn, vo) < TOL_FIT


def test_variance_ratio_markers_from_holdout_store(golden_dir):
    """isVarRatioGeno: the markers come from the host-resident hold-out store (Get_OneSNP_Geno_forVarRatio)."""
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200, step1, SaigeB200Error
    N0, M0 = 1000, 4000
    bed = O.synth_bed(N0, M0, seed=31, miss_rate=0.01).reshape(M0, -1).copy()
    # every other marker with its alleles swapped (hom A1 <-> hom A2), so that the flip-to-minor branch (FG.R:2318-2320) is taken
    x = bed[::2]
    same = ~(x ^ (x >> 1)) & 0x55
    bed[::2] = x ^ (same | (same << 1))
    bed = bed.reshape(-1)
    vr = np.unique(np.random.default_rng(2).integers(0, M0, size=400))
    g = SaigeB200(device=0)
    try:
        g.setminMAFforGRM(0.05); g.setmaxMissingRateforGRM(0.15); g.setminMAC_VarianceRatio(20, -1, True)
        g.setgeno_mem(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8), vr_rand_idx=vr)
        assert g.getIsVarRatioGeno() and g.Mvr >= 60
        yb, _, X = _pheno(golden_dir)
        probes = step1.ProbeStream(N0, nmax=130, seed=200)
        m = step1.glmmkin_ai_PCG(g, step1.glm_fit(yb, X, step1.Binomial), probes, trait="binary")
        order = np.random.default_rng(3).permutation(g.Mvr)
        vm, lm = step1.extractVarianceRatio(g, m, step1.Binomial, order)
        vn, ln = step1.extractVarianceRatio(g, m, step1.Binomial, order, native_loops=True)
        assert len(ln) == len(lm) and rel(ln, lm) < TOL_SAME and rel(vn, vm) < TOL_SAME
        used = [int(g.Get_OneSNP_Geno_forVarRatio(i).sum()) for i in order[:len(ln)]]
        assert any(a > N0 for a in used) and any(a < N0 for a in used)
        noK = m["obj_noK"]
        W = (step1.Binomial.mu_eta(m["linear_predictors"]) / np.sqrt(step1.Binomial.variance(m["fitted_values"]))) ** 2
        SiX = g.getSigma_X(W, m["theta"], m["X"], 500, 1e-5)
        with pytest.raises(SaigeB200Error):
            g.varianceRatioMarkers([g.Mvr], True, W, m["theta"], m["X"], noK["XV"], noK["XXVX_inv"], SiX, None, 500, 1e-5)
        with pytest.raises(SaigeB200Error):
            g.varianceRatioMarkers(np.zeros(129, dtype=np.int64), True, W, m["theta"], m["X"], noK["XV"], noK["XXVX_inv"], SiX, None, 500, 1e-5)
    finally:
        g.close()


def test_one_call_fit_stopping_rules_and_errors(loco_pair, golden_dir):
    """maxiter exhausted -> converged False with the same tau path as the mirror; argument errors are reported, not crashed on."""
    import ctypes as C
    from saige_gpu_b200 import step1, SaigeB200Error, _lib
    g, o = loco_pair
    yb, _, X = _pheno(golden_dir)
    fit0 = step1.glm_fit(yb, X, step1.Binomial)
    probes = step1.ProbeStream(o.N, nmax=130, seed=200)
    mm = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", maxiter=1, tol=1e-9)
    mn = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", maxiter=1, tol=1e-9, native_loops=True)
    g.setProbeStreamFixed(False)
    assert mm["converged"] is False and mn["converged"] is False and mn["n_outer"] == 1
    assert rel(mn["theta"], mm["theta"]) < TOL_SAME and rel(mn["coefficients"], mm["coefficients"]) < TOL_SAME
    # tauInit given (FG.R:152-156): the start value is taken over
    m2 = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", tauInit=(0.0, 0.4))
    n2 = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", tauInit=(0.0, 0.4), native_loops=True)
    g.setProbeStreamFixed(False)
    assert rel(n2["theta"], m2["theta"]) < TOL_SAME and n2["n_outer"] == len(m2["tau_path"]) - 1
    # NULL probe callback, nrun out of range, family code out of range
    N, p = o.N, X.shape[1]
    z = np.zeros(N); Xf = np.asfortranarray(X); out = np.zeros((N, 30), order="F")
    pp = lambda a: a.ctypes.data_as(C.c_void_p)
    L = _lib.lib()
    args = lambda nrun, cb: (g._h, 0, pp(yb), pp(Xf), p, pp(z), pp(np.zeros(p)), pp(z), pp(np.zeros(2)), 5, 0.02, nrun, 1e-5, 500, 0.0025, 0,
                             cb, None, pp(np.zeros(2)), pp(np.zeros(p)), pp(out), pp(out), pp(out), pp(np.zeros((p, p))), None, None,
                             None, None, None, None, None, None, _lib.CHROM_FN(), None)
    assert L.sgb_glmmkin_ai_pcg(*args(30, _lib.PROBE_FN())) != 0 and b"probe callback" in L.sgb_last_error(g._h)
    assert L.sgb_glmmkin_ai_pcg(*args(1, g._probe_cb(None, probes.fresh))) != 0 and b"nrun" in L.sgb_last_error(g._h)
    with pytest.raises(SaigeB200Error):
        g.Get_Coef_LOCO_all(yb, X, np.array([1.0, 0.3]), "binomial", np.zeros(p), fit0["eta"], z, 500, 1e-5, 0)


def test_loco_all_requires_loco_diagonal(chr22, golden_dir):
    from saige_gpu_b200 import SaigeB200, SaigeB200Error, step1
    from oracle import oracle as O
    N0, M0 = chr22["N0"], chr22["M0"]
    g = SaigeB200(device=0)
    try:
        g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
        g.setgeno_mem(chr22["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
        step1.set_loco_ranges(g, chr22["chrs"][g.getQCdMarkerIndex()])
        yb, _, X = _pheno(golden_dir)
        with pytest.raises(SaigeB200Error, match="set_Diagof_StdGeno_LOCO"):
            g.Get_Coef_LOCO_all(yb, X, np.array([1.0, 0.3]), "binomial", np.zeros(3), np.zeros(N0), np.zeros(N0), 500, 1e-5, 5)
    finally:
        g.close()
