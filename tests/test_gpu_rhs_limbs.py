"""Tolerance-driven digit count of wide batches (sgb_set_rhs_limbs): 7 digits = the exact 55-bit integers of the k <= 2 kernel,
6 / 5 digits = a coarser fixed point that must still sit inside the 1e-10 product gate and must not move a PCG iteration."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def pair(grm10k):
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno()
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    g = SaigeB200(device=0)
    g.setminMAFforGRM(0.01)
    g.setmaxMissingRateforGRM(0.15)
    p = grm10k["prefix"]
    g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    yield g, o
    g.close()


@pytest.mark.parametrize("limbs,tol", [(7, 1e-13), (6, 1e-12), (5, 1e-10)])
@pytest.mark.parametrize("k", [3, 16, 31])
def test_products_inside_the_gate(pair, limbs, tol, k):
    g, o = pair
    g.set_rhs_limbs(limbs)
    rng = np.random.default_rng(300 + k)
    B = rng.normal(size=(o.N, k)) * 10.0 ** rng.integers(-3, 4, size=k)
    B[:, 1] = rng.integers(0, 2, size=o.N) * 2.0 - 1.0
    B[:, 2] = rng.standard_t(3, size=o.N)                   # heavy tails: max / rms of the column is large
    Y, Yo = g.getCrossprodMatAndKin(B), o.getCrossprodMatAndKin(B)
    worst = max(rel(Y[:, c], Yo[:, c]) for c in range(k))
    g.set_rhs_limbs(7)
    assert worst < tol, (limbs, k, worst)


def test_seven_digits_equal_the_narrow_kernel_bit_for_bit(pair):
    g, o = pair
    rng = np.random.default_rng(9)
    B = rng.normal(size=(o.N, 5))
    g.set_rhs_limbs(7)
    wide = g.getCrossprodMatAndKin(B)
    narrow = np.column_stack([g.getCrossprodMatAndKin(B[:, c]) for c in range(5)])
    assert np.array_equal(wide, narrow)


@pytest.mark.parametrize("limbs", [5, 6])
def test_pcg_iteration_counts_do_not_move(pair, limbs):
    g, o = pair
    rng = np.random.default_rng(17)
    B = rng.normal(size=(o.N, 12))
    w = rng.uniform(0.05, 0.25, size=o.N)
    tau = np.array([1.0, 0.4])
    g.set_rhs_limbs(7)
    X7, it7 = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    g.set_rhs_limbs(limbs)
    X, it = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    g.set_rhs_limbs(7)
    Xo, ito = o.pcg_multi(w, tau, B, 500, 1e-5)
    assert list(it) == list(it7) == list(ito)
    assert rel(X, Xo) < 1e-6 and rel(X, X7) < 1e-8


def test_tolerance_setter_and_errors(pair):
    from saige_gpu_b200.api import SaigeB200Error
    g, _ = pair
    g.set_product_tolerance(1e-10)      # 5 digits: 16 * 2^-38 = 5.8e-11
    g.set_product_tolerance(1e-12)      # 6 digits
    g.set_product_tolerance(0.0)        # nothing coarser than the exact mode satisfies it
    with pytest.raises(SaigeB200Error):
        g.set_rhs_limbs(4)
    g.set_rhs_limbs(7)
