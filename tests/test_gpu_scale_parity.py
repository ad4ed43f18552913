"""GPU vs CPU oracle at the BASELINE.json shapes (L1 of the parity ladder, <= 1e-10 relative, fp64 oracle):
GRM product, LOCO product for three chromosomes and diag(K) at config 2 (50,000 samples x 500,000 markers) and at
config 3's sample count (200,000 samples x 50,000 markers).  Both sides draw the SAME genotypes from the counter-based
generator (oracle: orc_synth_bed -> orc_setgeno; GPU: sgb_setgeno_synth), so the comparison covers ingest statistics
(allele counts, bit-exact), the decode, the standardisation algebra and both sweeps at production row lengths.

Reference arithmetic: FG.cpp:1576-1598 (product), :1790-1851 (LOCO: full sweep minus the chromosome's range, divided by
M - M_chr), :665-729 (diagonal)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20260117
SHAPES = {"c2_50kx500k": (50_000, 500_000), "c3rows_200kx50k": (200_000, 50_000)}
LOCO_CHROMS = (0, 7, 21)        # first, a middle one, last (0-based chromosome index)


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.fixture(scope="module", params=sorted(SHAPES))
def pair(request):
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200, synth, step1
    N, M = SHAPES[request.param]
    _, t0, t1 = synth.thresholds(M, SEED)
    g = SaigeB200(device=0)
    g.setminMAFforGRM(0.01)
    g.setmaxMissingRateforGRM(0.15)
    g.setgeno_synth(N, M, SEED, t0, t1)
    bed = O.synth_bed(N, M, SEED)
    o = O.OracleGeno()
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(bed, N, M, np.arange(1, N + 1), np.ones(N, np.uint8))
    del bed
    chrq = synth.chromosomes(M)[g.getQCdMarkerIndex()]
    _, s, e = O.updateChrStartEndIndexVec(chrq)
    o.setStartEndIndexVec(s, e)
    step1.set_loco_ranges(g, chrq)
    yield dict(g=g, o=o, N=N, M=M, s=s, e=e)
    g.close()


def test_ingest_statistics_bit_exact(pair):
    g, o = pair["g"], pair["o"]
    assert (g.N, g.M) == (pair["N"], o.M)
    assert np.array_equal(g.getAlleleCountVec(), o.ACVec)
    assert np.array_equal(g.getMACVec(), o.MACVec)
    assert np.array_equal(g.getAlleleFreqVec(), o.alleleFreqVec)
    for idx in (0, 1, o.M // 3, o.M - 1):
        assert np.array_equal(g.Get_OneSNP_Geno(idx), o.Get_OneSNP_Geno(idx))


def test_product_matches_oracle(pair):
    g, o, N = pair["g"], pair["o"], pair["N"]
    rng = np.random.default_rng(11)
    B = np.column_stack([rng.normal(size=N), rng.integers(0, 2, size=N) * 2.0 - 1.0])
    Yo = o.getCrossprodMatAndKin(B)
    Y2 = g.getCrossprodMatAndKin(B)                    # k = 2: mma.sync engine
    y1 = g.getCrossprodMatAndKin(B[:, 0])              # k = 1: the headline kernel
    B4 = np.column_stack([B, B[:, 0] - B[:, 1], 0.5 * B[:, 1]])
    Y4 = g.getCrossprodMatAndKin(B4)                   # k = 4: tcgen05 engine
    assert rel(Y2, Yo) < 1e-10
    assert rel(y1, Yo[:, 0]) < 1e-10
    assert rel(Y4[:, :2], Yo) < 1e-10
    assert rel(Y4[:, 2], Yo[:, 0] - Yo[:, 1]) < 1e-10
    pair["B"], pair["Yo"] = B, Yo


def test_loco_products_match_oracle(pair):
    g, o, N = pair["g"], pair["o"], pair["N"]
    s, e = pair["s"], pair["e"]
    rng = np.random.default_rng(12)
    b = rng.normal(size=N)
    for j in LOCO_CHROMS:
        assert s[j] >= 0
        o.setStartEndIndex(s[j], e[j], j)
        g.setStartEndIndex(s[j], e[j], j)
        assert rel(g.getCrossprodMatAndKin_LOCO(b), o.getCrossprodMatAndKin_LOCO(b)) < 1e-10, j


def test_diag_matches_oracle(pair):
    g, o = pair["g"], pair["o"]
    assert rel(g.get_DiagofKin(), o.get_DiagofKin()) < 1e-10
    o.set_Diagof_StdGeno_LOCO()
    g.set_Diagof_StdGeno_LOCO()
    w = np.random.default_rng(13).uniform(0.05, 0.25, size=pair["N"])
    tau = np.array([1.0, 0.4])
    s, e = pair["s"], pair["e"]
    for j in LOCO_CHROMS:
        o.setStartEndIndex(s[j], e[j], j)
        g.setStartEndIndex(s[j], e[j], j)
        assert rel(g.getDiagOfSigma_LOCO(w, tau), o.getDiagOfSigma(w, tau, loco=True)) < 1e-10, j


@pytest.mark.parametrize("digits,tol", [(7, 1e-10), (5, 1e-10)])
def test_wide_batch_matches_oracle(pair, digits, tol):
    """31 columns (one N = 224 / N = 160 tcgen05 pass per sweep): every column is a known combination of the two columns the
    oracle multiplied in test_product_matches_oracle, so K's linearity gives the oracle answer for all 31 without 31 CPU products."""
    g, N = pair["g"], pair["N"]
    if "Yo" not in pair:
        pytest.skip("needs test_product_matches_oracle")
    B, Yo = pair["B"], pair["Yo"]
    rng = np.random.default_rng(14)
    C = rng.normal(size=(2, 31)) * 10.0 ** rng.integers(-2, 3, size=31)
    C[:, 0], C[:, 1] = (1.0, 0.0), (0.0, 1.0)
    g.set_rhs_limbs(digits)
    try:
        Y = g.getCrossprodMatAndKin(np.asfortranarray(B @ C))
    finally:
        g.set_rhs_limbs(7)
    want = Yo @ C
    worst = max(rel(Y[:, c], want[:, c]) for c in range(31))
    assert worst < tol, (digits, worst)
