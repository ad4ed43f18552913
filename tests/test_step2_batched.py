"""The batched step-2 path (score sums of a variant chunk as one tensor-engine GEMM, per-variant kernel only for flagged
variants) against the per-variant kernel of round 1 and against the oracle's marker loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(rng, N, p, trait):
    X = np.column_stack([np.ones(N), rng.normal(size=(N, p - 1))])
    if trait == "binary":
        beta = np.concatenate([[-1.8], rng.normal(scale=0.3, size=p - 1)])
        mu = 1 / (1 + np.exp(-(X @ beta + rng.normal(scale=0.3, size=N))))
        y = (rng.uniform(size=N) < mu).astype(np.float64)
        tau = np.array([1.0, 0.4]); mu2 = mu * (1 - mu)
    else:
        y = X @ rng.normal(scale=0.3, size=p) + rng.normal(size=N)
        mu = X @ np.linalg.lstsq(X, y, rcond=None)[0]
        tau = np.array([0.7, 0.3]); mu2 = np.full(N, 1 / tau[0])
    res = y - mu
    XV = (X * mu2[:, None]).T
    XVX = X.T @ XV.T
    XVXi = np.linalg.inv(XVX)
    return dict(mu=mu, res=res, mu2=mu2, tau=tau, trait=trait, y=y, X=X, XV=XV, XVX=XVX, XXVX_inv=X @ XVXi,
                XVX_inv_XV=(X @ XVXi) * mu2[:, None], S_a=(X * res[:, None]).sum(0), varRatio=0.93)


def _bed_with_flips(n_fam, nm, seed, miss):
    from oracle import oracle as O
    bed = O.synth_bed(n_fam, nm, seed=seed, miss_rate=miss)
    B0 = (n_fam + 3) // 4
    rows = bed.reshape(nm, B0).copy()
    for m in range(0, nm, 3):                 # every third marker major-allele coded -> the flip branch
        r = rows[m]
        lo, hi = r & 0x55, (r >> 1) & 0x55
        hom = ~(lo ^ hi) & 0x55
        rows[m] = r ^ (hom | (hom << 1))
    return rows.reshape(-1)


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("identity", [True, False])
def test_batched_equals_per_variant_kernel(trait, identity):
    from saige_gpu_b200 import SaigeB200
    rng = np.random.default_rng(21)
    n_fam, nm, p = 2051, 1500, 3
    N = n_fam if identity else 1777
    bed = _bed_with_flips(n_fam, nm, 4, 0.01)
    pos = np.arange(N, dtype=np.int32) if identity else rng.permutation(n_fam)[:N].astype(np.int32)
    M = _model(rng, N, p, trait)
    g = SaigeB200(device=0)
    try:
        g.setSAIGEobjInCPP(M, 0.93, 2.0, pos)
        g.setMaxMACforER(4.0)
        g.setStep2Batched(False)
        ref = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 0.5, 0.15)
        g.setStep2Batched(True)
        out = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 0.5, 0.15)
    finally:
        g.close()
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.array_equal(out[:, 0], ref[:, 0]) and np.array_equal(out[:, 10], ref[:, 10])        # tested, Is.SPA
    a, b = np.nan_to_num(out), np.nan_to_num(ref)
    scale = np.maximum(np.abs(b), 1e-300)
    worst = np.max(np.abs(a - b) / scale)
    assert worst < 1e-9, worst
    assert (ref[:, 0] == 1).sum() > 1000
    if trait == "binary":
        assert ref[:, 10].sum() > 20


def test_batched_vs_oracle_larger_cohort():
    """20,011 samples (ragged: not a multiple of 4, 16 or 256) x 600 variants incl. rare ones, identity map."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import SaigeB200
    rng = np.random.default_rng(33)
    n_fam, nm, p = 20011, 600, 4
    bed = _bed_with_flips(n_fam, nm, 8, 0.005)
    pos = np.arange(n_fam, dtype=np.int32)
    M = _model(rng, n_fam, p, "binary")
    g = SaigeB200(device=0)
    try:
        g.setSAIGEobjInCPP(M, 0.93, 2.0, pos)
        out = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 0.5, 0.15)
    finally:
        g.close()
    nspa = 0
    for m in range(0, nm, 7):
        r = S2.test_marker(M, S2.plink_marker(bed, n_fam, m, pos), min_mac=0.5)
        assert (r is not None) == (out[m, 0] == 1.0), m
        if r is None:
            continue
        got = dict(zip(SaigeB200.STEP2_COLUMNS, out[m]))
        nspa += bool(r["Is_SPA"])
        for col, oc in (("AC_Allele2", "AC_Allele2"), ("AF_Allele2", "AF_Allele2"), ("BETA", "BETA"), ("SE", "SE"), ("Tstat", "Tstat"),
                        ("var", "var"), ("p.value", "p_value"), ("p.value.NA", "p_value_NA")):
            assert abs(got[col] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (m, col, got[col], r[oc])
        assert bool(got["Is.SPA"]) == bool(r["Is_SPA"])
    assert nspa >= 2


def _strong_signal_bed(rng, n, y, nm):
    """nm variants whose alternate allele is strongly enriched in the upper half of y (hard calls, no missing)."""
    hi = y > np.median(y)
    rows = np.zeros((nm, (n + 3) // 4), dtype=np.uint8)
    for m in range(nm):
        f = np.where(hi, 0.45, 0.05 + 0.02 * (m % 5))
        g = (rng.uniform(size=n) < f).astype(int) + (rng.uniform(size=n) < f).astype(int)
        code = np.array([3, 2, 0], dtype=np.uint8)[g]                      # 0 copies -> 11, 1 -> 10, 2 -> 00
        code = np.concatenate([code, np.full((-n) % 4, 3, dtype=np.uint8)]).reshape(-1, 4)
        rows[m] = code[:, 0] | (code[:, 1] << 2) | (code[:, 2] << 4) | (code[:, 3] << 6)
    return rows.reshape(-1)


@pytest.mark.parametrize("trait", ["quantitative", "binary"])
def test_pvalues_below_the_double_range_come_back_on_the_log_scale(trait):
    """stat > 1490: the p-value underflows a double; the reference switches to log p-values and "%.1fE%d" strings
    (SAIGE_test.cpp:255-284, 531-582).  The kernel's log columns against the oracle, SE and the Firth trigger included."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import SaigeB200, step2
    rng = np.random.default_rng(77)
    n, nm, p = 30011, 24, 3
    M = _model(rng, n, p, trait)
    bed = _strong_signal_bed(rng, n, M["y"] if trait == "quantitative" else M["y"] + 0.01 * rng.normal(size=n), nm)
    if trait == "binary":                       # cases carry the allele: |z| far beyond the cutoff, saddle-point branch on the log scale
        bed = _strong_signal_bed(rng, n, M["y"], nm)
    M["offset"] = np.zeros(n)
    pos = np.arange(n, dtype=np.int32)
    g = SaigeB200(device=0)
    try:
        g.setSAIGEobjInCPP(M, 0.93, 2.0, pos)
        g.setFirth(True, 0.01, M["offset"], se_from_fit=False)
        out = g.mainMarkerInCPP(bed, n, nm, 0.0, 0.5, 0.15)
    finally:
        g.close()
    cols = {c: i for i, c in enumerate(SaigeB200.STEP2_COLUMNS)}
    nlog = 0
    for m in range(nm):
        r = S2.test_marker(M, S2.plink_marker(bed, n, m, pos), min_mac=0.5, is_Firth_beta=True, pCutoffforFirth=0.01, firth_se_from_fit=False)
        row = out[m]
        assert row[0] == 1.0
        for gc, oc in (("log.p.value", "log_p_value"), ("log.p.value.NA", "log_p_value_NA"), ("SE", "SE"), ("BETA", "BETA"), ("Tstat", "Tstat")):
            assert abs(row[cols[gc]] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (trait, m, gc, row[cols[gc]], r[oc])
        assert bool(row[cols["Is.SPA"]]) == bool(r["Is_SPA"]) and bool(row[cols["Is.Firth"]]) == bool(r["Is_Firth"])
        if r["p_value_NA"] == 0.0:
            nlog += 1
            assert row[cols["p.value.NA"]] == 0.0 and np.isfinite(row[cols["log.p.value.NA"]]) and row[cols["log.p.value.NA"]] < -708
            assert np.isfinite(row[cols["SE"]]) and row[cols["SE"]] > 0
            txt = step2.format_logp(row[cols["log.p.value"]])
            mant, expo = txt.split("E")
            assert 1.0 <= float(mant) < 10.0 and int(expo) < -307
    assert nlog >= nm // 2, nlog


@pytest.mark.parametrize("identity", [True, False])
def test_chunk_borders_do_not_change_a_row(identity):
    """The double-buffered chunk pipeline (H2D of chunk c+1 and D2H of chunk c-1 behind the kernels of chunk c): many small
    chunks, a ragged last one and a second call on the same handle with another chunk size give the one-chunk table bit for bit."""
    from saige_gpu_b200 import SaigeB200
    rng = np.random.default_rng(5)
    n_fam, nm, p = 3001, 2500, 3
    N = n_fam if identity else 2222
    bed = _bed_with_flips(n_fam, nm, 6, 0.01)
    pos = np.arange(N, dtype=np.int32) if identity else rng.permutation(n_fam)[:N].astype(np.int32)
    M = _model(rng, N, p, "binary")
    g = SaigeB200(device=0)
    try:
        g.setSAIGEobjInCPP(M, 0.93, 2.0, pos)
        one = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 0.5, 0.15)
        B0 = (n_fam + 3) // 4
        for rows_per_chunk in (600, 37, 1024):
            g.setStep2ChunkBytes(rows_per_chunk * B0)
            out = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 0.5, 0.15)
            assert np.array_equal(np.nan_to_num(out, nan=-7.0), np.nan_to_num(one, nan=-7.0)), rows_per_chunk
    finally:
        g.close()
    assert one[:, 10].sum() > 30 and (one[:, 0] == 1).sum() > 2000


def test_more_variants_than_a_grid_dimension():
    """Small cohorts put > 65,535 variants into one 1 GB chunk: the chunk is capped so that the per-variant grid dimension of the
    re-pack kernel stays legal.  70,000 variants x 403 samples, batched = per-variant kernel on a sample of rows."""
    from saige_gpu_b200 import SaigeB200
    rng = np.random.default_rng(8)
    n_fam, nm, p = 403, 70_000, 2
    bed = _bed_with_flips(n_fam, nm, 12, 0.0)
    pos = np.arange(n_fam, dtype=np.int32)
    M = _model(rng, n_fam, p, "quantitative")
    g = SaigeB200(device=0)
    try:
        g.setSAIGEobjInCPP(M, 0.93, 2.0, pos)
        out = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 0.5, 0.15)
        g.setStep2Batched(False)
        B0 = (n_fam + 3) // 4
        tail = g.mainMarkerInCPP(bed[(nm - 3000) * B0:], n_fam, 3000, 0.0, 0.5, 0.15)
    finally:
        g.close()
    a, b = np.nan_to_num(out[nm - 3000:]), np.nan_to_num(tail)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) < 1e-9
    assert (out[:, 0] == 1).sum() > 60_000
