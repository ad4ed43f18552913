"""CPU emulation of the tensor kernel's data path (saige_gpu_b200/csrc/kernels.cu): limb splitting, the fragment
layout written by split_limbs_kernel, the prmt-based 2-bit decode and the m16n8k32 A/B/C fragment ownership.
It pins the index algebra (which genotype meets which limb) without a GPU."""
import numpy as np

LIMBS = 8


def split_limbs(v, E):
    q = np.rint(np.ldexp(v, 54 - E)).astype(np.int64)
    out = np.zeros((len(v), LIMBS), dtype=np.int64)
    for l in range(LIMBS):
        d = ((q + 64) & 127) - 64
        q = (q - d) >> 7
        out[:, l] = d
    assert np.all(q == 0)
    return out


def frag_offset(r, l):
    """byte offset inside a 2048-byte block of limb l of genotype slot r (0..255) -- mirrors split_limbs_kernel"""
    t, wi, p = r >> 6, (r >> 4) & 3, r & 15
    odd, half, slot = p & 1, (p >> 3) & 1, (p & 7) >> 1
    j = 2 * wi + odd
    return (l * 4 + t) * 64 + j * 8 + half * 4 + slot


def decode16(w, pool=(0, 1, 2, 3)):
    """mirrors decode16(): returns 4 registers, each a list of 4 byte values"""
    e = w & 0x33333333
    o = (w >> 2) & 0x33333333

    def prmt(sel):
        return [pool[(sel >> (4 * s)) & 7] for s in range(4)]
    return [prmt(e & 0xFFFF), prmt(e >> 16), prmt(o & 0xFFFF), prmt(o >> 16)]


def test_recombination_is_exact_enough():
    rng = np.random.default_rng(1)
    v = rng.normal(size=1000) * 10.0 ** rng.integers(-6, 6, size=1000)
    mx = np.abs(v).max()
    E = int(np.floor(np.log2(mx)))
    limbs = split_limbs(v, E)
    assert limbs.min() >= -64 and limbs.max() <= 63
    rec = sum(limbs[:, l].astype(np.float64) * 128.0 ** l for l in range(LIMBS)) * 2.0 ** (E - 54)
    assert np.max(np.abs(rec - v)) <= 2.0 ** (E - 54)          # half an ulp of the fixed-point grid, doubled for slack


def test_fragment_offsets_are_a_bijection():
    seen = set()
    for r in range(256):
        for l in range(LIMBS):
            seen.add(frag_offset(r, l))
    assert seen == set(range(2048))


def test_warp_tile_product_matches_dense():
    """One warp, one k-step (64 bytes = 256 genotypes per row), 16 rows, one limb column group."""
    rng = np.random.default_rng(7)
    geno = rng.integers(0, 3, size=(16, 256))
    packed = np.zeros((16, 64), dtype=np.uint8)
    for i in range(256):
        packed[:, i >> 2] |= (geno[:, i] << (2 * (i & 3))).astype(np.uint8)
    b = rng.normal(size=256)
    E = int(np.floor(np.log2(np.abs(b).max())))
    limbs = split_limbs(b, E)                     # 256 x 8
    L = np.zeros(2048, dtype=np.int64)
    for r in range(256):
        for l in range(LIMBS):
            L[frag_offset(r, l)] = limbs[r, l]
    C = np.zeros((16, 8), dtype=np.int64)
    words = packed.view(np.uint32).reshape(16, 16)          # little endian, 16 words per row
    for j in range(8):                                      # the 8 MMAs of a k-step
        A = np.zeros((16, 32), dtype=np.int64)
        Bm = np.zeros((32, 8), dtype=np.int64)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            wi = j // 2
            dl = decode16(int(words[g, 4 * t + wi]))
            dh = decode16(int(words[g + 8, 4 * t + wi]))
            q0 = 0 if j % 2 == 0 else 2
            a0, a1, a2, a3 = dl[q0], dh[q0], dl[q0 + 1], dh[q0 + 1]
            b0 = [L[lane * 64 + j * 8 + s] for s in range(4)]
            b1 = [L[lane * 64 + j * 8 + 4 + s] for s in range(4)]
            for s in range(4):                              # PTX m16n8k32 fragment ownership
                A[g, 4 * t + s] = a0[s]; A[g + 8, 4 * t + s] = a1[s]
                A[g, 16 + 4 * t + s] = a2[s]; A[g + 8, 16 + 4 * t + s] = a3[s]
                Bm[4 * t + s, g] = b0[s]; Bm[16 + 4 * t + s, g] = b1[s]
        C += A @ Bm
    expect = geno @ limbs
    assert np.array_equal(C, expect)
    rec = sum(C[:, l].astype(np.float64) * 128.0 ** l for l in range(LIMBS)) * 2.0 ** (E - 54)
    assert np.allclose(rec, geno @ b, rtol=1e-12, atol=1e-12)


def test_indicator_plane_pool():
    w = 0b10_01_00_10_10_00_01_10_00_00_10_01_10_10_01_00
    val = decode16(w, pool=(0, 1, 2, 3))
    ind = decode16(w, pool=(0, 0, 1, 0))
    for a, b in zip(val, ind):
        assert [int(x == 2) for x in a] == b
