"""Dense N x N GRM (BASELINE config 4): the tcgen05 build and the stored-GRM products / PCG against the oracle.

The reference fork has no code for this step, so the oracle is the definition itself: K = Z Z^T / M with the oracle's
fp64 standardised genotypes (the same z_m the on-the-fly product is graded on).  Tolerances: 1e-10 relative for matrix
entries and products (north_star's matvec bound) with the default 7 weight limbs; 1e-6 for PCG solutions.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def oracle_K(o):
    Z = np.stack([o.Get_OneSNP_StdGeno(m) for m in range(o.M)], axis=1)      # N x M
    return Z @ Z.T / o.M


@pytest.fixture(scope="module")
def dense10k(grm10k):
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno()
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    g = SaigeB200(device=0)
    g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15); g.setminMAC_VarianceRatio(20, -1, False)
    p = grm10k["prefix"]
    g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    info = g.buildDenseGRM()
    yield g, o, oracle_K(o), info
    g.close()


def test_dense_grm_matches_definition(dense10k):
    g, o, K, info = dense10k
    assert info["weight_limbs"] == 7 and info["block_rows"] == (o.N + 127) // 128
    # lower block-trapezoid of fp64: sum over block-rows of 128 * 128 (R + 1) entries
    nbr = info["block_rows"]
    assert info["stored_bytes"] == 8 * 128 * 128 * nbr * (nbr + 1) // 2
    got = g.getDenseGRMBlock(0, o.N, 0, o.N)
    assert rel(got, K) < TOL
    assert np.array_equal(got, got.T)                       # both triangles are served from one stored element
    assert rel(np.diag(got), o.get_DiagofKin()) < TOL       # get_DiagofKin is the diagonal of the same matrix
    # arbitrary windows, including ones that straddle block-rows and the diagonal
    for (i0, ni, j0, nj) in [(0, 1, 0, 1), (127, 3, 126, 5), (500, 300, 10, 77), (10, 77, 500, 300), (999, 1, 0, 1000)]:
        assert rel(g.getDenseGRMBlock(i0, ni, j0, nj), K[i0:i0 + ni, j0:j0 + nj]) < TOL


@pytest.mark.parametrize("limbs,tol", [(3, 1e-3), (4, 2e-6), (5, 2e-8), (8, 1e-10)])
def test_weight_limbs_set_the_precision(dense10k, limbs, tol):
    g, o, K, _ = dense10k
    try:
        g.buildDenseGRM(weight_limbs=limbs)
        assert rel(g.getDenseGRMBlock(0, o.N, 0, o.N), K) < tol
    finally:
        g.buildDenseGRM()


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 31])
def test_stored_grm_products(dense10k, k):
    g, o, K, _ = dense10k
    rng = np.random.default_rng(100 + k)
    B = rng.normal(size=(o.N, k))
    packed = g.getCrossprodMatAndKin(B)
    g.setGRMMode("dense")
    try:
        dense = g.getCrossprodMatAndKin(B)
    finally:
        g.setGRMMode("packed")
    assert rel(dense, K @ B) < TOL
    assert rel(dense, packed) < TOL
    assert rel(dense, o.getCrossprodMatAndKin(B)) < TOL


def test_pcg_on_stored_grm(dense10k):
    g, o, K, _ = dense10k
    rng = np.random.default_rng(5)
    w = rng.uniform(0.02, 0.25, size=o.N); tau = np.array([1.0, 0.7]); B = rng.normal(size=(o.N, 6))
    want, it_want = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    g.setGRMMode("dense")
    try:
        got, it_got = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
        sig = g.getCrossprod(B, w, tau)
    finally:
        g.setGRMMode("packed")
    assert np.array_equal(it_got, it_want)
    assert rel(got, want) < 1e-8
    for c in range(B.shape[1]):
        assert rel(got[:, c], o.getPCG1ofSigmaAndVector(w, tau, B[:, c], 500, 1e-5)) < 1e-6
    assert rel(sig, tau[0] * B / w[:, None] + tau[1] * (K @ B)) < TOL


def test_dense_mode_errors(dense10k, chr22):
    from saige_gpu_b200.api import SaigeB200Error
    from saige_gpu_b200 import SaigeB200
    g, o, K, _ = dense10k
    h = SaigeB200(device=0)
    try:
        with pytest.raises(SaigeB200Error, match="not loaded"):
            h.buildDenseGRM()
        h.setminMAFforGRM(0.0); h.setmaxMissingRateforGRM(1.0); h.setminMAC_VarianceRatio(20, -1, False)
        from oracle import oracle as O
        bed = O.synth_bed(64, 300, seed=9)
        h.setgeno_mem(bed, 64, 300, np.arange(1, 65), np.ones(64, np.uint8))
        with pytest.raises(SaigeB200Error, match="before sgb_dense_grm_build"):
            h.setGRMMode("dense")
        with pytest.raises(SaigeB200Error, match="limbs"):
            h.buildDenseGRM(weight_limbs=9)
        h.buildDenseGRM()
        h.setGRMMode("dense")
        h.setStartEndIndexVec(np.array([0]), np.array([99])); h.setStartEndIndex(0, 99, 0)
        with pytest.raises(SaigeB200Error, match="LOCO"):
            h.getCrossprodMatAndKin_LOCO(np.ones(64))
        # reloading genotypes drops the stored matrix and falls back to the packed mode
        h.setgeno_mem(bed, 64, 300, np.arange(1, 65), np.ones(64, np.uint8))
        with pytest.raises(SaigeB200Error, match="not built"):
            h.getDenseGRMBlock(0, 1, 0, 1)
        h.getCrossprodMatAndKin(np.ones(64))
    finally:
        h.close()


@pytest.mark.parametrize("shape", [(5, 9), (37, 500), (128, 257), (129, 1024), (1025, 777), (3001, 130)])
def test_dense_grm_ragged_shapes(shape):
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200
    N0, M0 = shape
    bed = O.synth_bed(N0, M0, seed=2000 + N0, miss_rate=0.02)
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.0, 1.0
    o.setgeno(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    g = SaigeB200(device=0)
    try:
        g.setminMAFforGRM(0.0); g.setmaxMissingRateforGRM(1.0); g.setminMAC_VarianceRatio(20, -1, False)
        g.setgeno_mem(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
        g.buildDenseGRM()
        K = oracle_K(o)
        assert rel(g.getDenseGRMBlock(0, N0, 0, N0), K) < TOL
        B = np.random.default_rng(N0).normal(size=(N0, 3))
        g.setGRMMode("dense")
        assert rel(g.getCrossprodMatAndKin(B), K @ B) < TOL
    finally:
        g.close()


def test_gcta_files(dense10k, tmp_path):
    g, o, K, _ = dense10k
    prefix = os.path.join(tmp_path, "grm")
    g.writeDenseGRM(prefix, rows_per_read=300)
    tri = np.fromfile(prefix + ".grm.bin", dtype=np.float32)
    nn = np.fromfile(prefix + ".grm.N.bin", dtype=np.float32)
    assert tri.size == nn.size == o.N * (o.N + 1) // 2
    assert np.all(nn == o.M)
    il = np.tril_indices(o.N)
    assert np.array_equal(tri, K[il].astype(np.float32)) or rel(tri, K[il]) < 1e-6
    assert len(open(prefix + ".grm.id").read().splitlines()) == o.N
