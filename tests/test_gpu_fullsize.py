"""Size-independent properties of the CUDA path at a BASELINE.json-scale shape (config 2: 50,000 samples x 500,000
markers would take the oracle hours, so correctness at that size is established through properties the domain
offers): linearity, symmetry, positive semi-definiteness, K.1 = 0, tensor engine == fp64 engine on a marker
sub-range, diag(K) == e_i^T K e_i, and PCG actually solving Sigma x = b (residual through an independent product)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, M, SEED = 50_000, 500_000, 20260117


@pytest.fixture(scope="module")
def big():
    from saige_gpu_b200 import SaigeB200, synth
    g = SaigeB200(device=0)
    _, t0, t1 = synth.thresholds(M, SEED)
    g.setminMAFforGRM(0.01)
    g.setmaxMissingRateforGRM(0.15)
    g.setgeno_synth(N, M, SEED, t0, t1)
    yield g
    g.close()


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_shape_and_allele_counts(big):
    from saige_gpu_b200 import synth
    f, _, _ = synth.thresholds(M, SEED)
    assert (big.N, big.M, big.M0) == (N, M, M)
    ac = big.getAlleleCountVec()
    assert np.max(np.abs(ac / (2.0 * N) - f)) < 0.01               # binomial sampling noise at N = 50k
    g0 = big.Get_OneSNP_Geno(123_456)
    assert g0.sum() == ac[123_456] and set(np.unique(g0)) <= {0, 1, 2}


def test_linearity_symmetry_psd(big):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(N, 3))
    A[:, 2] = 2.0 * A[:, 0] - 3.0 * A[:, 1]
    K = big.getCrossprodMatAndKin(A)
    assert rel(K[:, 2], 2.0 * K[:, 0] - 3.0 * K[:, 1]) < 1e-10
    s01, s10 = A[:, 0] @ K[:, 1], A[:, 1] @ K[:, 0]
    assert abs(s01 - s10) / max(abs(s01), 1e-300) < 1e-9
    assert A[:, 0] @ K[:, 0] > 0 and A[:, 1] @ K[:, 1] > 0
    ones = big.getCrossprodMatAndKin(np.ones(N))
    assert np.max(np.abs(ones)) < 1e-8                             # every marker column is centred exactly


def test_diag_matches_unit_vector_products(big):
    d = big.get_DiagofKin()
    assert abs(d.mean() - 1.0) < 0.01
    E = np.zeros((N, 3))
    idx = [0, 31_415, N - 1]
    for c, i in enumerate(idx):
        E[i, c] = 1.0
    K = big.getCrossprodMatAndKin(E)
    for c, i in enumerate(idx):
        assert abs(K[i, c] - d[i]) / d[i] < 1e-10


def test_engines_agree_at_scale(big):
    rng = np.random.default_rng(1)
    B = rng.normal(size=(N, 4))
    yt = big.getCrossprodMatAndKin(B)            # tcgen05 kernel (k = 4)
    big.set_engine("imma")
    yi = big.getCrossprodMatAndKin(B)            # mma.sync kernel
    big.set_engine("f64")
    yf = big.getCrossprodMatAndKin(B[:, 0])      # fp64 FMA kernels
    big.set_engine("tensor")
    assert np.array_equal(yt, yi)
    assert rel(yt[:, 0], yf) < 1e-10


def test_pcg_residual_at_scale(big):
    rng = np.random.default_rng(2)
    w = rng.uniform(0.05, 0.25, size=N)
    tau = np.array([1.0, 0.5])
    B = np.column_stack([rng.normal(size=N), rng.integers(0, 2, size=N) * 2.0 - 1.0])
    X, it = big.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    R = B - big.getCrossprod(X, w, tau)
    for c in range(2):
        assert R[:, c] @ R[:, c] <= 1.05e-5 and 1 <= it[c] < 100
    # batched == sequential, bit for bit
    x0 = big.getPCG1ofSigmaAndVector(w, tau, B[:, 0], 500, 1e-5)
    assert np.array_equal(x0, X[:, 0])


def test_dense_grm_at_scale():
    """BASELINE config 4 row at a reduced shape (20k x 100k: 1.6 GB stored, 0.25 s build): the stored matrix against the
    on-the-fly product through size-independent properties: K b equal for both representations, rows of K summing to zero
    (every marker column is centred), diag(K) == get_DiagofKin, a random window symmetric and equal to e_i^T K e_j."""
    from saige_gpu_b200 import SaigeB200, synth
    n, m = 20_000, 100_000
    g = SaigeB200(device=0)
    try:
        _, t0, t1 = synth.thresholds(m, SEED + 3)
        g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
        g.setgeno_synth(n, m, SEED + 3, t0, t1)
        info = g.buildDenseGRM()
        nbr = (n + 127) // 128
        assert info["stored_bytes"] == 8 * 128 * 128 * nbr * (nbr + 1) // 2 and info["int8_ops"] > 0
        rng = np.random.default_rng(4)
        B = rng.normal(size=(n, 5))
        packed = g.getCrossprodMatAndKin(B)
        g.setGRMMode("dense")
        dense = g.getCrossprodMatAndKin(B)
        ones = g.getCrossprodMatAndKin(np.ones(n))
        g.setGRMMode("packed")
        assert rel(dense, packed) < 1e-10
        assert np.max(np.abs(ones)) < 1e-9
        i0, j0 = 12_345, 777
        win = g.getDenseGRMBlock(i0, 200, j0, 150)
        assert np.array_equal(win, g.getDenseGRMBlock(j0, 150, i0, 200).T)
        E = np.zeros((n, 2)); E[j0 + 3, 0] = 1.0; E[j0 + 100, 1] = 1.0
        cols = g.getCrossprodMatAndKin(E)
        assert rel(win[:, 3], cols[i0:i0 + 200, 0]) < 1e-10 and rel(win[:, 100], cols[i0:i0 + 200, 1]) < 1e-10
        d = np.array([g.getDenseGRMBlock(i, 1, i, 1)[0, 0] for i in (0, 127, 128, 9_999, n - 1)])
        assert rel(d, g.get_DiagofKin()[[0, 127, 128, 9_999, n - 1]]) < 1e-10
    finally:
        g.close()
