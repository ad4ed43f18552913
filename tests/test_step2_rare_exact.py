"""Step 2, exact test of rare variants ("efficient resampling", MAC <= max_MAC_for_ER on a binary trait).

Parity chain:  reference's own compiled code (oracle/_ref/libskat_exact_ref.so, built by oracle/Makefile from
/root/reference/src/SAIGE/src/Binary_*.cpp)  ->  tests/golden/er_golden.json (tests/golden/make_er_golden.py)
->  oracle restatement (oracle/step2_oracle.py: er_pvalue)  and  the product's arithmetic (saige_gpu_b200/csrc/er_exact.h,
compiled for the host here, for the device in step2.cu)  ->  GPU kernel against the oracle's marker loop.
Tolerance: 1e-8 relative against the golden vectors (the reference's stand-alone build sums logarithms where R's lchoose
and this code use lgamma), 1e-6 relative GPU vs oracle (north_star: step-2 p-values)."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "er_golden.json")))


@pytest.fixture(scope="module")
def host_er(tmp_path_factory):
    """er_exact.h compiled for the host: the very arithmetic the kernel runs, without the kernel."""
    d = tmp_path_factory.mktemp("er_host")
    src = d / "er_host.cpp"
    src.write_text('#include "er_exact.h"\n'
                   'extern "C" double er_host(int k, const double *g1, const double *p1, const double *res1, double p2mean,\n'
                   '                          double n, double ncase, double eps)\n'
                   '{ return sgb_er_exact_pvalue(k, g1, p1, res1, p2mean, n, ncase, eps); }\n')
    so = d / "liber_host.so"
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "saige_gpu_b200", "csrc"), "-o", str(so), str(src)])
    L = ctypes.CDLL(str(so))
    L.er_host.restype = ctypes.c_double
    L.er_host.argtypes = [ctypes.c_int, DP, DP, DP, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]

    def call(g1, p1, res1, p2mean, n, ncase, eps=1e-6):
        g1, p1, res1 = (np.ascontiguousarray(a, dtype=np.float64) for a in (g1, p1, res1))
        return L.er_host(len(g1), g1.ctypes.data_as(DP), p1.ctypes.data_as(DP), res1.ctypes.data_as(DP), p2mean, n, ncase, eps)
    return call


def test_oracle_er_matches_reference_compiled_code(golden):
    from oracle import step2_oracle as S2
    assert len(golden["cases"]) >= 150
    ks = set()
    for c in golden["cases"]:
        g1, p1, res1 = (np.array(c[k]) for k in ("g1", "p1", "res1"))
        ks.add(len(g1))
        prob = S2.er_group_prob(p1, c["p2mean"], c["n"], c["ncase"])
        assert np.allclose(prob, c["prob"], rtol=1e-8, atol=1e-300)
        pv = S2.er_pvalue(g1, p1, res1, c["p2mean"], c["n"], c["ncase"])
        assert abs(pv - c["pvalue"]) <= 1e-8 * abs(c["pvalue"]), (c["n"], c["ncase"], len(g1), pv, c["pvalue"])
    assert ks == set(range(1, 11))


def test_product_er_arithmetic_matches_reference_compiled_code(golden, host_er):
    for c in golden["cases"]:
        pv = host_er(c["g1"], c["p1"], c["res1"], c["p2mean"], c["n"], c["ncase"])
        assert abs(pv - c["pvalue"]) <= 1e-8 * abs(c["pvalue"]), (c["n"], c["ncase"], len(c["g1"]), pv, c["pvalue"])


def test_er_edge_cases(host_er):
    from oracle import step2_oracle as S2
    # a single carrier who is a case: exact mid-p = P(case) / 2 with P(case) from the one-class hypergeometric weight
    p1, p2 = np.array([0.05]), 0.1
    pv = host_er([1.0], p1, [0.95], p2, 1000, 100)
    w = (0.05 / 0.95) / (p2 / (1 - p2))
    pc = w * 100 / (w * 100 + 900)              # C(999, 99) w : C(999, 100) = w * 100 : 900
    assert abs(pv - pc / 2) < 1e-9 * pc
    assert abs(S2.er_pvalue(np.array([1.0]), p1, np.array([0.95]), p2, 1000, 100) - pv) < 1e-12
    # equal genotypes and probabilities (every assignment of a size ties with the others), no carrier is a case:
    # statistic (0 - 0.8)^2 is matched or beaten by 0, 2 (not: (2 - 0.8)^2 > 0.64 yes), 3, 4 cases
    g1, pr = np.ones(4), np.full(4, 0.2)
    pv = host_er(g1, pr, -pr, 0.2, 500, 100)
    prob = S2.er_group_prob(pr, 0.2, 500, 100)
    assert abs(pv - (prob[0] / 2 + prob[2] + prob[3] + prob[4])) < 1e-12
    assert abs(pv - S2.er_pvalue(g1, pr, -pr, 0.2, 500, 100)) < 1e-12
    # more carriers than the enumeration is sized for: refused (NaN), never a wrong number
    assert np.isnan(host_er(np.ones(11), np.full(11, 0.1), np.full(11, -0.1), 0.1, 500, 100))


def test_library_refuses_a_cutoff_beyond_the_enumeration():
    src = open(os.path.join(ROOT, "saige_gpu_b200", "csrc", "step2.cu")).read()
    assert "max_mac_for_er > (double)SGB_ER_MAXK" in src


# ---- marker loop: rare variants through the oracle (CPU) and through the kernel (GPU) -------------------------------------
def pack_bed(G):
    """G: markers x n_fam genotypes = copies of A1 (2, 1, 0) or -1 (missing) -> raw PLINK rows (PLINK.hpp:48-56)."""
    code = np.array([3, 2, 0, 1], dtype=np.uint8)[G]          # 0 -> 11, 1 -> 10, 2 -> 00, -1 -> 01
    nm, n = G.shape
    B0 = (n + 3) // 4
    pad = np.full((nm, B0 * 4), 3, dtype=np.uint8)
    pad[:, :n] = code
    q = pad.reshape(nm, B0, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8).reshape(-1)


def rare_variant_set(seed, n_fam, N, identity):
    """Binary-trait model + markers with 1..7 minor alleles, carriers enriched among the cases (so that most scores pass the
    SPA cutoff), some markers major-allele coded (flip), some with missing calls, some with a homozygous carrier."""
    rng = np.random.default_rng(seed)
    p = 3
    pos = np.arange(N, dtype=np.int32) if identity else rng.permutation(n_fam)[:N].astype(np.int32)
    X = np.column_stack([np.ones(N), rng.normal(size=(N, p - 1))])
    mu = 1 / (1 + np.exp(-(X @ np.array([-2.0, 0.5, -0.4]) + rng.normal(scale=0.4, size=N))))
    y = (rng.uniform(size=N) < mu).astype(np.float64)
    res, mu2 = y - mu, mu * (1 - mu)
    XV = (X * mu2[:, None]).T
    XVX = X.T @ XV.T
    XVX_inv = np.linalg.inv(XVX)
    M = dict(mu=mu, res=res, mu2=mu2, tau=np.array([1.0, 0.35]), trait="binary", y=y, X=X, XV=XV, XVX=XVX,
             XXVX_inv=X @ XVX_inv, XVX_inv_XV=(X @ XVX_inv) * mu2[:, None], S_a=(X * res[:, None]).sum(0), varRatio=0.91,
             offset=rng.normal(scale=0.1, size=N))
    cases, ctrls = np.nonzero(y == 1)[0], np.nonzero(y == 0)[0]
    nm = 240
    G = np.zeros((nm, n_fam), dtype=np.int64)
    for m in range(nm):
        mac = 1 + m % 7
        alleles = mac
        while alleles > 0:
            i = rng.choice(cases) if rng.uniform() < 0.8 else rng.choice(ctrls)
            f = pos[i]
            if G[m, f] == 0:
                c = 2 if (alleles >= 2 and rng.uniform() < 0.2) else 1
                G[m, f] = c
                alleles -= c
        if m % 4 == 1:
            G[m] = 2 - G[m]                                   # A1 is the major allele: the kernel flips
        if m % 5 == 2:
            G[m, rng.choice(n_fam, size=6, replace=False)] = -1
    return M, pos, pack_bed(G), nm


def test_oracle_marker_loop_takes_the_exact_branch():
    from oracle import step2_oracle as S2
    n_fam, N = 900, 800
    M, pos, bed, nm = rare_variant_set(11, n_fam, N, identity=False)
    ner = nspa = 0
    for m in range(nm):
        G = S2.plink_marker(bed, n_fam, m, pos)
        r4 = S2.test_marker(M, G, max_MAC_for_ER=4)
        r0 = S2.test_marker(M, G)
        assert r4 is not None
        mac = min(r4["AC_Allele2"], 2 * N - r4["AC_Allele2"])
        if r4["Is_ER"]:
            ner += 1
            assert mac <= 4 and not r4["Is_SPA"]           # (without the exact test the saddle point of a singleton often fails to converge)
            assert 0 < r4["p_value"] <= 1 and r4["p_value"] != r0["p_value"] and r4["p_value_NA"] == r0["p_value_NA"]
            from scipy import stats
            assert abs(r4["SE"] - abs(r4["BETA"]) / abs(stats.norm.ppf(r4["p_value"] / 2))) < 1e-12 * r4["SE"]
        else:
            nspa += bool(r4["Is_SPA"])
            assert r4["p_value"] == r0["p_value"]
            assert mac > 4 or abs(r4["Tstat"]) / np.sqrt(r4["var"]) <= 2.0
    assert ner > 60 and nspa > 40


@pytest.mark.gpu
@pytest.mark.parametrize("identity", [True, False])
def test_gpu_exact_test_of_rare_variants_vs_oracle(identity):
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import SaigeB200, SaigeB200Error
    n_fam, N = 900, 800
    M, pos, bed, nm = rare_variant_set(12 + identity, n_fam, N, identity)
    g = SaigeB200(device=0)
    g.setSAIGEobjInCPP(M, M["varRatio"], 2.0, pos)
    base = g.mainMarkerInCPP(bed, n_fam, nm)
    cols = (("AC_Allele2", "AC_Allele2"), ("AF_Allele2", "AF_Allele2"), ("BETA", "BETA"), ("SE", "SE"), ("Tstat", "Tstat"),
            ("var", "var"), ("p.value", "p_value"), ("p.value.NA", "p_value_NA"))
    for cutoff in (4.0, 10.0):
        g.setMaxMACforER(cutoff)
        out = g.mainMarkerInCPP(bed, n_fam, nm)
        ner = 0
        for m in range(nm):
            r = S2.test_marker(M, S2.plink_marker(bed, n_fam, m, pos), max_MAC_for_ER=cutoff)
            got = dict(zip(g.STEP2_COLUMNS, out[m]))
            assert got["tested"] == 1.0
            ner += bool(r["Is_ER"])
            assert bool(got["Is.SPA"]) == bool(r["Is_SPA"]), (m, cutoff)
            for col, oc in cols:
                assert abs(got[col] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (identity, cutoff, m, col, got[col], r[oc])
            if not r["Is_ER"]:
                assert np.array_equal(np.nan_to_num(out[m]), np.nan_to_num(base[m])), m      # untouched by the switch
        assert ner > (60 if cutoff == 4.0 else 100)
    # Firth's effect size on top of the exact p-value: SE from the fit, or |beta| / |qnorm(p/2)| (SAIGE_test.cpp:614-632)
    g.setMaxMACforER(4.0)
    for from_fit in (True, False):
        g.setFirth(True, 0.05, M["offset"], se_from_fit=from_fit)
        outf = g.mainMarkerInCPP(bed, n_fam, nm, se_two_sided=False)
        nf = 0
        for m in range(nm):
            r = S2.test_marker(M, S2.plink_marker(bed, n_fam, m, pos), max_MAC_for_ER=4.0, is_Firth_beta=True,
                               pCutoffforFirth=0.05, firth_se_from_fit=from_fit, se_two_sided=False)
            got = dict(zip(g.STEP2_COLUMNS, outf[m]))
            assert bool(got["Is.Firth"]) == bool(r["Is_Firth"])
            nf += bool(r["Is_Firth"] and r["Is_ER"])
            for col, oc in (("BETA", "BETA"), ("SE", "SE"), ("p.value", "p_value")):
                assert abs(got[col] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (identity, from_fit, m, col, got[col], r[oc])
        assert nf > 10
    g.setFirth(False)
    g.setMaxMACforER(-1.0)
    assert np.array_equal(np.nan_to_num(g.mainMarkerInCPP(bed, n_fam, nm)), np.nan_to_num(base))
    with pytest.raises(SaigeB200Error):
        g.setMaxMACforER(11.0)
    g.close()


@pytest.mark.gpu
def test_gpu_categorical_variance_ratios_vs_oracle():
    """assignVarianceRatio (SAIGE_test.cpp:801-833): the kernel picks the variance ratio by the variant's MAC category."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import SaigeB200, SaigeB200Error
    n_fam, N = 900, 800
    M, pos, bed, nm = rare_variant_set(21, n_fam, N, identity=True)
    g = SaigeB200(device=0)
    for ratios, lo, hi in (([0.8, 1.1], (2, 4.5), (4.5,)), ([0.7, 0.9, 1.2], (1, 3, 5), (3, 5))):
        M.update(varRatio=ratios, cateVarRatioMinMACVecExclude=lo, cateVarRatioMaxMACVecInclude=hi)
        g.setSAIGEobjInCPP(M, ratios, 2.0, pos)
        g.setMaxMACforER(4.0)                          # (the saddle point of a singleton rarely converges; the exact test is the reference's path)
        out = g.mainMarkerInCPP(bed, n_fam, nm)
        seen = set()
        for m in range(nm):
            r = S2.test_marker(M, S2.plink_marker(bed, n_fam, m, pos), max_MAC_for_ER=4.0)
            got = dict(zip(g.STEP2_COLUMNS, out[m]))
            seen.add(round(got["var"] / got["var2"], 12))
            for col, oc in (("BETA", "BETA"), ("SE", "SE"), ("Tstat", "Tstat"), ("var", "var"), ("p.value", "p_value"),
                            ("p.value.NA", "p_value_NA")):
                assert abs(got[col] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (ratios, m, col, got[col], r[oc])
        assert seen == set(ratios)
    with pytest.raises(SaigeB200Error):
        g.setVarianceRatios([0.8, 1.1], (2, 4.5), (5.0,))          # categories do not tile the MAC axis
    # back to a single ratio
    M["varRatio"] = 0.91
    g.setSAIGEobjInCPP(M, 0.91, 2.0, pos)
    out = g.mainMarkerInCPP(bed, n_fam, nm)
    assert np.allclose(out[:, 7] / out[:, 19], 0.91, rtol=1e-12)
    g.close()
