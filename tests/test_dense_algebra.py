"""CPU emulation of the dense-GRM arithmetic of csrc/dense_grm.cu (no GPU): fixed-point weights split into balanced
base-128 digits, int8 operands, exact integer accumulation, limb recombination as q_lo + 2^28 q_hi, and the 128-bit
centring  N^2 2^S M K_ij = N^2 Q_ij - N (U'_i + U'_j) + C'.  Checked against K = Z Z^T / M in fp64 and in exact rationals."""
from fractions import Fraction

import numpy as np
import pytest


def build(g, limbs):
    """g: N x M genotypes (0/1/2 copies of A1).  Returns K (fp64) through the library's integer pipeline."""
    N, M = g.shape
    ac = g.sum(0).astype(np.int64)
    f = ac / (2.0 * N)
    v = 2.0 * f * (1.0 - f)
    s = np.where(v > 0, 1.0 / np.sqrt(np.where(v > 0, v, 1.0)), 0.0)
    s2 = s * s
    e2 = int(np.frexp(s2.max())[1])
    S = 7 * limbs - 2 - e2
    W = np.array([int(round(float(x) * 2.0 ** S)) for x in s2], dtype=object)
    # balanced digits, |digit * h| <= 128 must fit int8
    digits = np.zeros((limbs, M), dtype=np.int64)
    Wr = W.copy()
    for l in range(limbs):
        d = np.array([((int(w) + 64) & 127) - 64 for w in Wr], dtype=np.int64)
        Wr = np.array([(int(w) - int(x)) >> 7 for w, x in zip(Wr, d)], dtype=object)
        digits[l] = d
    assert all(int(w) == 0 for w in Wr), "weight does not fit the digits"
    h = (2 - g).astype(np.int64)                                   # what the decode feeds the MMA
    B = digits[:, None, :] * h[None, :, :]                         # [limb][sample][marker]
    assert B.min() >= -128 and B.max() <= 127
    acc = np.einsum("im,ljm->lij", h, B)                           # int32 range on the device
    assert np.abs(acc).max() < 2 ** 31
    qlo = sum((128 ** l) * acc[l].astype(np.float64) for l in range(min(limbs, 4)))
    qhi = sum((128 ** (l - 4)) * acc[l].astype(np.float64) for l in range(4, limbs)) if limbs > 4 else np.zeros_like(qlo)
    assert np.abs(qlo).max() < 2 ** 53 and np.abs(qhi).max() < 2 ** 53
    cm = 2 * N - ac                                                # N phi_m = sum_i h_im
    Cp = sum(int(W[m]) * int(cm[m]) ** 2 for m in range(M))
    # U' from 12-bit pieces of W: every piece sum stays an exact double
    pieces = []
    for p in range(5):
        vp = np.array([((int(W[m]) >> (12 * p)) & 4095) * int(cm[m]) for m in range(M)], dtype=np.float64)
        up = h.astype(np.float64) @ vp
        assert np.abs(up).max() < 2 ** 53
        pieces.append(up)
    K = np.zeros((N, N))
    mul = 2.0 ** (-S) / (M * N * N)
    for i in range(N):
        for j in range(N):
            Q = int(qlo[i, j]) + (int(qhi[i, j]) << 28)
            Us = sum((int(pieces[p][i]) + int(pieces[p][j])) << (12 * p) for p in range(5))
            T = N * (N * Q - Us) + Cp
            K[i, j] = float(T) * mul
    return K, W, S


@pytest.mark.parametrize("limbs,tol", [(4, 2e-6), (6, 1e-10), (7, 1e-12), (8, 1e-13)])
def test_integer_pipeline_matches_definition(limbs, tol):
    rng = np.random.default_rng(limbs)
    N, M = 24, 300
    f = rng.uniform(0.02, 0.5, size=M)
    g = rng.binomial(2, f, size=(N, M))
    g[:, g.sum(0) == 0] = 1                                        # no monomorphic marker (weight 0 is tested below)
    K, _, _ = build(g, limbs)
    fm = g.sum(0) / (2.0 * N)
    Z = (g - 2 * fm) / np.sqrt(2 * fm * (1 - fm))
    Kref = Z @ Z.T / M
    assert np.max(np.abs(K - Kref)) / np.max(np.abs(Kref)) < tol
    assert np.allclose(K, K.T, rtol=0, atol=0)


def test_pipeline_is_exact_given_the_integer_weights():
    """With W_m in place of s_m^2 2^S the result is the correctly rounded rational number."""
    rng = np.random.default_rng(1)
    N, M = 10, 60
    g = rng.binomial(2, rng.uniform(0.1, 0.5, size=M), size=(N, M))
    g[:, 0] = 1                                                    # monomorphic marker: f = 0.5?  no: all het -> AC = N
    g[:, 1] = 0                                                    # AC = 0 -> weight 0
    K, W, S = build(g, 7)
    ac = g.sum(0)
    for i in range(N):
        for j in range(N):
            exact = Fraction(0)
            for m in range(M):
                phi = Fraction(2 * N - int(ac[m]), N)
                exact += int(W[m]) * (Fraction(2 - int(g[i, m])) - phi) * (Fraction(2 - int(g[j, m])) - phi)
            exact = exact / (2 ** S) / M
            assert K[i, j] == float(exact) or abs(K[i, j] - float(exact)) <= 2 * np.spacing(abs(float(exact)))
