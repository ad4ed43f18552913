// Exact recombination of the int32 limb accumulators of the two tensor engines (shared by the stand-alone recombine
// kernels and the fused product epilogues).  Both engines hold the same 55-bit fixed-point integer q = round(v 2^(53-E))
// per right-hand-side element and return the same bits.
#pragma once
#include <stdint.h>

// mma.sync engine: 8 balanced base-128 digits; acc layout [row][kpad*8 + c*8 + l].  Returns sum_k plane(g) * q, zeroes acc.
__device__ __forceinline__ double sgb_recombine_imma(int32_t *__restrict__ acc, int64_t r, int c, int kpad, const int32_t *__restrict__ limbsum, int c0)
{
    int4 *p = reinterpret_cast<int4 *>(acc + (r * kpad + c) * 8);
    int4 lo = p[0], hi = p[1];
    p[0] = make_int4(0, 0, 0, 0);
    p[1] = make_int4(0, 0, 0, 0);
    // acc = sum (c0 - plane) * limb  ->  sum plane * limb = c0 * (column limb sum) - acc
    const int4 *ls = reinterpret_cast<const int4 *>(limbsum + c * 8);
    int4 s0 = ls[0], s1 = ls[1];
    lo.x = c0 * s0.x - lo.x; lo.y = c0 * s0.y - lo.y; lo.z = c0 * s0.z - lo.z; lo.w = c0 * s0.w - lo.w;
    hi.x = c0 * s1.x - hi.x; hi.y = c0 * s1.y - hi.y; hi.z = c0 * s1.z - hi.z; hi.w = c0 * s1.w - hi.w;
    long long l4 = (long long)lo.x + ((long long)lo.y << 7) + ((long long)lo.z << 14) + ((long long)lo.w << 21);
    long long h4 = (long long)hi.x + ((long long)hi.y << 7) + ((long long)hi.z << 14) + ((long long)hi.w << 21);
    return (double)h4 * 268435456.0 + (double)l4;    // 128^4 = 2^28; both halves exact (< 2^53)
}

// tcgen05 engine: NL (5..7) balanced base-256 digits of q = round(v 2^(8 NL - 3 - E)) (NL = 7: the 55-bit value of the
// mma.sync engine; fewer digits = a coarser fixed point, see sgb_set_rhs_limbs); acc layout [row][npad], column of
// (c, l) = NL*c + l.  Zeroes the accumulators.
// Exact integer sum (up to ~2^81), then ONE correctly rounded conversion: keep 62 leading bits plus a sticky bit.
// sgb_umma_value: the value of ONE (row, column) from its NL digits `p` (any address space) and the column's limb sums `ls`.
template <int NL>
__device__ __forceinline__ double sgb_umma_value(const int32_t *p, const int32_t *__restrict__ ls, int c0)
{
    long long x[7];
#pragma unroll
    for (int l = 0; l < 7; l++) x[l] = 0;
#pragma unroll
    for (int l = 0; l < NL; l++) x[l] = (long long)c0 * ls[l] - (long long)p[l];
    const long long l4 = x[0] + (x[1] << 8) + (x[2] << 16) + (x[3] << 24);
    const long long h3 = x[4] + (x[5] << 8) + (x[6] << 16);
    const __int128 T = ((__int128)h3 << 32) + (__int128)l4;
    const bool neg = T < 0;
    const unsigned __int128 a = neg ? (unsigned __int128)(-T) : (unsigned __int128)T;
    const unsigned long long ahi = (unsigned long long)(a >> 64), alo = (unsigned long long)a;
    const int bits = ahi ? 128 - __clzll((long long)ahi) : (alo ? 64 - __clzll((long long)alo) : 0);
    const int shift = bits > 62 ? bits - 62 : 0;
    unsigned long long m = (unsigned long long)(a >> shift);
    if (shift && (a & ((((unsigned __int128)1) << shift) - 1))) m |= 1ull;
    // (double)m is the one rounding; 2^shift (shift <= 66) scales it exactly
    const double v = (double)m * __longlong_as_double((long long)(1023 + shift) << 52);
    return neg ? -v : v;
}

template <int NL>
__device__ __forceinline__ double sgb_recombine_umma(int32_t *__restrict__ acc, int64_t r, int c, int npad, const int32_t *__restrict__ limbsum, int c0)
{
    int32_t *p = acc + r * npad + NL * c;
    const double v = sgb_umma_value<NL>(p, limbsum + c * NL, c0);
#pragma unroll
    for (int l = 0; l < NL; l++) p[l] = 0;
    return v;
}

// right-hand-side columns per block of the tiled epilogues (recomb_post{1,2}_wide_kernel)
#define SGB_RCG 8

// NL = 8: the mma.sync engine; NL = 5..7: the tcgen05 engine with NL digits
template <int NL>
__device__ __forceinline__ double sgb_recombine(int32_t *__restrict__ acc, int64_t r, int c, int pad, const int32_t *__restrict__ limbsum, int c0)
{
    if constexpr (NL == 8) return sgb_recombine_imma(acc, r, c, pad, limbsum, c0);
    else return sgb_recombine_umma<NL>(acc, r, c, pad, limbsum, c0);
}
