"""CPU emulation of the tensor kernel's data path (saige_gpu_b200/csrc/kernels.cu): limb splitting, the fragment
layout written by split_limbs_kernel, the prmt-based 2-bit decode and the m16n8k32 A/B/C fragment ownership.
It pins the index algebra (which genotype meets which limb) without a GPU."""
import numpy as np

LIMBS = 8


def split_limbs(v, E):
    q = np.rint(np.ldexp(v, 53 - E)).astype(np.int64)
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
    return (wi * 32 + l * 4 + t) * 16 + odd * 8 + half * 4 + slot


POOLS = {  # (ax, ay, bx, by) exactly as in k_pk2_gemm
    "value": (0x02000102, 0x01020001, 0x01020202, 0x00000101),
    "is2": (0x01000101, 0x01010001, 0x01010101, 0x00000101),
}


def prmt(a, b, sel):
    """PTX prmt.b32 generic mode: nibble bits 0-2 pick a byte of {a,b}; bit 3 replicates that byte's sign."""
    pool = [(a >> (8 * i)) & 255 for i in range(4)] + [(b >> (8 * i)) & 255 for i in range(4)]
    out = []
    for s in range(4):
        n = (sel >> (4 * s)) & 15
        v = pool[n & 7]
        out.append((255 if v & 128 else 0) if n & 8 else v)
    return out


def decode16(w, plane="value"):
    """mirrors decode16(): returns 4 registers, each a list of 4 byte values"""
    ax, ay, bx, by = POOLS[plane]
    hi = w >> 16
    return [prmt(ax, ay, w), prmt(ax, ay, hi), prmt(bx, by, w), prmt(bx, by, hi)]


def pack_row(geno):
    """pair-ternary nibble coding (sgb_pack4): nibble = A + 3B"""
    n = len(geno)
    out = np.zeros((n + 3) // 4, dtype=np.uint8)
    g = np.concatenate([geno, np.zeros((-n) % 4, dtype=geno.dtype)])
    out[:] = (g[0::4] + 3 * g[1::4]) | ((g[2::4] + 3 * g[3::4]) << 4)
    return out


def test_recombination_is_exact_enough():
    rng = np.random.default_rng(1)
    v = rng.normal(size=1000) * 10.0 ** rng.integers(-6, 6, size=1000)
    mx = np.abs(v).max()
    E = int(np.floor(np.log2(mx)))
    limbs = split_limbs(v, E)
    assert limbs.min() >= -64 and limbs.max() <= 63
    rec = sum(limbs[:, l].astype(np.float64) * 128.0 ** l for l in range(LIMBS)) * 2.0 ** (E - 53)
    assert np.max(np.abs(rec - v)) <= 2.0 ** (E - 53)          # half an ulp of the fixed-point grid, doubled for slack


def test_limb_range_covers_the_largest_mantissa():
    """The column maximum may sit just below 2^(E+1); its fixed-point image must still fit 8 balanced digits."""
    for top in (1.0, 1.5, 1.984375, 1.9999999999999998):
        v = np.array([top, -top, top * 0.37, 1e-30, 0.0]) * 2.0 ** 7
        E = int(np.floor(np.log2(np.abs(v).max())))
        limbs = split_limbs(v, E)                      # asserts that no carry is left after 8 digits
        rec = sum(limbs[:, l].astype(np.float64) * 128.0 ** l for l in range(LIMBS)) * 2.0 ** (E - 53)
        assert np.max(np.abs(rec - v)) <= 2.0 ** (E - 53)


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
    packed = np.stack([pack_row(geno[r]) for r in range(16)])
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
            off = ((j // 2) * 32 + lane) * 16 + (j % 2) * 8      # uint4 index (pair, lane); .x,.y / .z,.w
            b0 = [L[off + s] for s in range(4)]
            b1 = [L[off + 4 + s] for s in range(4)]
            for s in range(4):                              # PTX m16n8k32 fragment ownership
                A[g, 4 * t + s] = a0[s]; A[g + 8, 4 * t + s] = a1[s]
                A[g, 16 + 4 * t + s] = a2[s]; A[g + 8, 16 + 4 * t + s] = a3[s]
                Bm[4 * t + s, g] = b0[s]; Bm[16 + 4 * t + s, g] = b1[s]
        C += A @ Bm
    expect = geno @ limbs
    # the decode feeds (2 - g); recombine_kernel undoes it with the column limb sums
    C = 2 * limbs.sum(0)[None, :] - C
    assert np.array_equal(C, expect)
    rec = sum(C[:, l].astype(np.float64) * 128.0 ** l for l in range(LIMBS)) * 2.0 ** (E - 53)
    assert np.allclose(rec, geno @ b, rtol=1e-12, atol=1e-12)


def test_decode_planes_cover_all_nine_nibbles():
    rng = np.random.default_rng(3)
    geno = np.concatenate([np.array([2, 2, 0, 0, 1, 2, 2, 1, 0, 1, 1, 0, 2, 0, 1, 1]), rng.integers(0, 3, 48)])
    words = pack_row(geno).view(np.uint32)
    order = [0, 2, 4, 6, 8, 10, 12, 14, 1, 3, 5, 7, 9, 11, 13, 15]       # d[0], d[1], d[2], d[3]
    for wi, w in enumerate(words):
        val = sum(decode16(int(w), "value"), [])
        ind = sum(decode16(int(w), "is2"), [])
        g = geno[16 * wi + np.array([0, 2, 4, 6, 8, 10, 12, 14, 1, 3, 5, 7, 9, 11, 13, 15])]
        assert val == list(2 - g) and ind == list(1 - (g == 2).astype(int)), order


def test_tcgen05_engine_uses_seven_byte_limbs():
    """pk2_umma.cu splits q = round(v 2^(53-E)) into 7 balanced base-256 digits (the B operand is s8) and recombines the
    int32 column sums exactly, with one correctly rounded conversion at the end.  Emulated here with Python integers."""
    rng = np.random.default_rng(3)

    def digits(q):
        out = []
        for _ in range(7):
            d = ((q + 128) & 255) - 128
            q = (q - d) >> 8
            out.append(d)
        assert q == 0
        return out
    # range: |q| <= 2^54 (mantissa of the column maximum just below 2, plus rounding)
    for q in [0, 1, -1, 127, 128, -128, -129, 2 ** 54, -(2 ** 54), 2 ** 54 - 1, 2 ** 53 + 12345, -(2 ** 53) - 7]:
        d = digits(q)
        assert all(-128 <= x <= 127 for x in d) and sum(x * 256 ** l for l, x in enumerate(d)) == q
    qs = [int(x) for x in rng.integers(-2 ** 54, 2 ** 54, size=2000)]
    assert all(sum(x * 256 ** l for l, x in enumerate(digits(q))) == q for q in qs)

    def to_double(T):
        # 62 leading bits + sticky, as recombine_umma_kernel does
        neg, a = T < 0, abs(T)
        bits = a.bit_length()
        shift = max(bits - 62, 0)
        m = a >> shift
        if shift and (a & ((1 << shift) - 1)):
            m |= 1
        v = float(m) * 2.0 ** shift
        return -v if neg else v
    # sums of K products (2 - g) * digit over a long row, per limb, then the recombination
    for _ in range(200):
        K = int(rng.integers(1, 200000))
        x = [int(rng.integers(-256 * K, 256 * K)) for _ in range(7)]
        T = sum(v << (8 * l) for l, v in enumerate(x))
        assert to_double(T) == float(T)                       # Python's int -> float is correctly rounded
    for T in [2 ** 80 + 1, 2 ** 80 + 2 ** 27, 2 ** 80 + 2 ** 27 + 1, -(2 ** 81) - 3, 2 ** 53 + 1, 2 ** 62 + 2 ** 8 + 1]:
        assert to_double(T) == float(T)


# ---------------------------------------------------------------------------------------------------
# tcgen05-engine epilogues (recomb_post{1,2}_wide_kernel, kernels.cu) and limb split (pk2_umma.cu): index algebra of the
# column groups / tile pieces and the one-rounding recombination (recombine.cuh), emulated on the CPU
# ---------------------------------------------------------------------------------------------------
RCG = 8          # SGB_RCG: right-hand-side columns per block


def _npad(k, nl):
    return (nl * k + 15) & ~15


def _group_pieces(k, nl, gy):
    """(w0, W4, nc) of column group gy exactly as the wide kernels compute them"""
    pad = _npad(k, nl)
    cg0 = gy * RCG
    nc = min(RCG, k - cg0)
    w0 = nl * cg0
    w1 = pad if cg0 + RCG >= k else nl * (cg0 + RCG)
    W4 = min((w1 - w0) >> 2, 2 * nl)
    return w0, W4, nc, pad


def test_wide_epilogue_groups_cover_every_digit_once_and_stay_aligned():
    for nl in (5, 6, 7):
        for k in range(1, 61):
            pad = _npad(k, nl)
            ngroups = (k + RCG - 1) // RCG
            seen = np.zeros(pad, dtype=np.int32)
            for gy in range(ngroups):
                w0, W4, nc, _ = _group_pieces(k, nl, gy)
                assert 1 <= nc <= RCG
                assert (w0 * 4) % 16 == 0, "tile row starts on a 16-byte boundary"       # 128-bit loads
                assert (pad * 4) % 64 == 0                                                 # accumulator row stride
                assert 0 < W4 <= 2 * nl and w0 + 4 * W4 <= pad, "pieces stay inside the accumulator row"
                assert 4 * W4 >= nl * nc, "every digit of the group's columns is inside the tile"
                seen[w0:w0 + 4 * W4] += 1
                # the tile piece -> (row, q) map of a 256-row tile: every (row, q) exactly once
                idx = np.arange(256 * W4)
                row, q = idx // W4, idx % W4
                assert len(set(zip(row.tolist(), q.tolist()))) == 256 * W4 and row.max() == 255
            assert np.all(seen[:nl * k] == 1), (nl, k)         # real digits: loaded and zeroed exactly once
            assert np.all(seen <= 1)                           # padding words: at most once (they are always zero)


def _umma_value(digits_acc, limbsum, c0):
    """sgb_umma_value: exact integer sum, 62 leading bits + sticky bit, ONE rounding, exact power-of-two scaling"""
    x = [c0 * int(ls) - int(a) for ls, a in zip(limbsum, digits_acc)]
    T = sum(v << (8 * l) for l, v in enumerate(x))
    neg, a = T < 0, abs(T)
    bits = a.bit_length()
    shift = max(0, bits - 62)
    m = a >> shift
    if shift and (a & ((1 << shift) - 1)):
        m |= 1
    v = float(m) * float(2 ** shift)             # float(m): round-to-nearest-even of a < 2^63 integer
    return -v if neg else v


def test_umma_recombination_is_the_correctly_rounded_exact_sum():
    from fractions import Fraction
    rng = np.random.default_rng(11)
    for nl in (5, 6, 7):
        for _ in range(300):
            acc = rng.integers(-2**31, 2**31, size=nl)
            ls = rng.integers(-2**27, 2**27, size=nl)
            got = _umma_value(acc, ls, 2)
            exact = sum((2 * int(s) - int(a)) << (8 * l) for l, (s, a) in enumerate(zip(ls, acc)))
            # correctly rounded: the double nearest to the exact integer (ties cannot be hit through the sticky bit)
            want = float(Fraction(exact))
            assert got == want, (nl, exact, got, want)


def _image_offset(n, w, q):
    """byte offset of the 4-byte word (accumulator column n, TMEM column group w, slot quad q) inside the image of a k-block"""
    return (n >> 3) * 1024 + w * 128 + (n & 7) * 16 + 4 * q


def test_limb_image_words_of_a_k_block_do_not_overlap():
    for nl, k in ((7, 31), (5, 31), (7, 4), (6, 9)):
        ngroups = _npad(k, nl) // 8
        used = np.zeros(ngroups * 1024, dtype=np.int32)
        for c in range(k):
            for l in range(nl):
                for col in range(32):
                    o = _image_offset(nl * c + l, col >> 2, col & 3)
                    used[o:o + 4] += 1
        assert used.max() == 1
        assert used.sum() == nl * k * 128                      # 128 genotypes x one byte per digit and column


def _split_umma(v, nl):
    """split_limbs_umma_kernel: q = rint(v 2^(8 nl - 3 - E)) as nl balanced base-256 digits; E = exponent of max |v|"""
    mb = np.max(np.abs(v))
    E = int(np.floor(np.log2(mb))) if mb > 0 else 0
    sh = 8 * nl - 3
    q = np.rint(np.ldexp(v, sh - E)).astype(np.int64)
    q0 = q.copy()
    digits = np.zeros((len(v), nl), dtype=np.int64)
    for l in range(nl):
        d = ((q + 128) & 255) - 128
        q = (q - d) >> 8
        digits[:, l] = d
    assert np.all(q == 0), "nl balanced base-256 digits hold |q| <= 2^(8 nl - 2)"
    return digits, q0, float(np.ldexp(1.0, E - sh))


def test_umma_digits_product_and_tolerance_model():
    """The tcgen05 engine end to end on the CPU: digits -> int32 accumulators of sum (2 - g) digit -> recombination.  7 digits
    reproduce the exact fixed-point dot product (the integers of the mma.sync engine); 5 digits stay within 2^-38 of the column
    maximum per input element (sgb_set_product_tolerance)."""
    rng = np.random.default_rng(5)
    n = 4096
    g = rng.integers(0, 3, size=n)
    for nl in (5, 6, 7):
        v = rng.normal(size=n) * np.exp(rng.normal(scale=3.0, size=n))      # heavy-tailed column
        digits, q, mult = _split_umma(v, nl)
        assert np.all(np.abs(digits) <= 128)
        assert np.array_equal(sum(digits[:, l] << (8 * l) for l in range(nl)), q)        # digits are q, exactly
        acc = ((2 - g)[:, None] * digits).sum(axis=0)                                     # what the MMAs accumulate
        assert np.all(np.abs(acc) < 2**31)
        limbsum = digits.sum(axis=0)
        got = _umma_value(acc, limbsum, 2) * mult
        exact_fixed = int((g * q).sum())                                                  # sum g q, exact integer
        assert got == float(exact_fixed) * mult or abs(got - exact_fixed * mult) <= abs(exact_fixed * mult) * 2.0**-52
        # quantisation of the inputs: |v - q mult| <= mult / 2 per element, mult = 2^(E - 8 nl + 3)
        assert np.max(np.abs(v - q * mult)) <= 0.5 * mult
        mb = np.max(np.abs(v))
        assert mult <= mb * 2.0 ** -(8 * nl - 3)            # nl = 5: <= 2^-37 of the column maximum, i.e. error <= 2^-38 per element
        err = abs(got - float((g * v).sum()))
        assert err <= n * 2 * 0.5 * mult + abs(got) * 2.0**-52


def test_shipped_recombination_header_on_the_host(tmp_path):
    """recombine.cuh itself (the product's arithmetic, not a restatement) compiled for the host with the CUDA intrinsics
    shimmed: both engines' recombination against exact integer arithmetic."""
    import ctypes, os, subprocess
    from fractions import Fraction
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "recomb_host.cpp"
    src.write_text(r'''
#include <stdint.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define __restrict__
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
#include "recombine.cuh"
extern "C" double umma5(const int32_t *p, const int32_t *ls, int c0) { return sgb_umma_value<5>(p, ls, c0); }
extern "C" double umma6(const int32_t *p, const int32_t *ls, int c0) { return sgb_umma_value<6>(p, ls, c0); }
extern "C" double umma7(const int32_t *p, const int32_t *ls, int c0) { return sgb_umma_value<7>(p, ls, c0); }
extern "C" double umma7_acc(int32_t *acc, long long r, int c, int npad, const int32_t *ls, int c0) { return sgb_recombine_umma<7>(acc, r, c, npad, ls, c0); }
extern "C" double imma(int32_t *acc, long long r, int c, int kpad, const int32_t *ls, int c0) { return sgb_recombine_imma(acc, r, c, kpad, ls, c0); }
''')
    so = tmp_path / "librecomb_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(root, "saige_gpu_b200", "csrc"),
                           "-o", str(so), str(src)])
    L = ctypes.CDLL(str(so))
    for f in (L.umma5, L.umma6, L.umma7, L.umma7_acc, L.imma):
        f.restype = ctypes.c_double
    rng = np.random.default_rng(17)
    I32 = ctypes.POINTER(ctypes.c_int32)
    for nl, fn in ((5, L.umma5), (6, L.umma6), (7, L.umma7)):
        for trial in range(400):
            hi = 2**31 if trial % 2 else 2**12                              # large and small magnitudes (shift = 0 path)
            acc = rng.integers(-hi, hi, size=nl).astype(np.int32)
            ls = rng.integers(-2**27, 2**27, size=nl).astype(np.int32)
            got = fn(acc.ctypes.data_as(I32), ls.ctypes.data_as(I32), 2)
            exact = sum((2 * int(s) - int(a)) << (8 * l) for l, (s, a) in enumerate(zip(ls, acc)))
            assert got == float(Fraction(exact)), (nl, exact, got)
            assert got == _umma_value(acc, ls, 2)
    # the accumulator-walking wrappers: value of (row, column) and the digits zeroed behind it
    npad, k = 32, 4
    acc = rng.integers(-2**30, 2**30, size=(3, npad)).astype(np.int32)
    ls = rng.integers(-2**20, 2**20, size=7 * k).astype(np.int32)
    ref = acc.copy()
    L.umma7_acc.argtypes = [I32, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, I32, ctypes.c_int]
    got = L.umma7_acc(acc.ctypes.data_as(I32), 1, 2, npad, ls.ctypes.data_as(I32), 2)
    assert got == _umma_value(ref[1, 14:21], ls[14:21], 2)
    assert np.all(acc[1, 14:21] == 0) and np.array_equal(np.delete(acc, np.s_[14:21], axis=1), np.delete(ref, np.s_[14:21], axis=1))
    # mma.sync engine: 8 base-128 digits, layout [row][kpad*8 + c*8 + l]
    kpad = 2
    acc8 = rng.integers(-2**30, 2**30, size=(2, kpad * 8)).astype(np.int32)
    ls8 = rng.integers(-2**20, 2**20, size=kpad * 8).astype(np.int32)
    ref8 = acc8.copy()
    L.imma.argtypes = [I32, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, I32, ctypes.c_int]
    got = L.imma(acc8.ctypes.data_as(I32), 1, 1, kpad, ls8.ctypes.data_as(I32), 2)
    exact = sum((2 * int(ls8[8 + l]) - int(ref8[1, 8 + l])) << (7 * l) for l in range(8))
    assert got == float(Fraction(exact))
    assert np.all(acc8[1, 8:16] == 0) and np.array_equal(acc8[0], ref8[0]) and np.array_equal(acc8[1, :8], ref8[1, :8])
