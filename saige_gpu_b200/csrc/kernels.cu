// CUDA kernels of libsaige_b200.so (sm_100a).  See DESIGN.md for the data layout and the roofline of each.
//
// Hot kernel: pk2_gemm_kernel -- "packed 2-bit matrix x int8 limb matrix -> int32", used for BOTH genotype
// sweeps of a GRM product (rows = markers over the marker-major copy, rows = samples over the sample-major
// copy), for the GRM diagonal and for the LOCO per-chromosome diagonals.  Replaces, on the GPU, the work of
// Get_OneSNP_StdGeno + dot + axpy (FG.cpp:582-662, 1478-1495) and of the reference's dense-fp32
// cublasSgemv pair (gpuSymMatMult.cu:267,273).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "sgb_internal.h"
#include "recombine.cuh"

#define LAUNCH_CHECK(h)                                                                               \
    do {                                                                                              \
        (h)->cnt.n_kernel_launches++;                                                                 \
        cudaError_t e__ = cudaGetLastError();                                                         \
        if (e__ != cudaSuccess) return sgb_fail(h, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, \
                                                cudaGetErrorString(e__));                             \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Device coding of the genotype store ("pair-ternary"): one nibble holds TWO genotypes A (even sample) and
// B (odd sample) as n = A + 3*B in 0..8; a byte = nibble(samples 4b, 4b+1) | nibble(samples 4b+2, 4b+3) << 4.
// Still 2 bits per genotype, but the UNMASKED nibble is directly a prmt selector (see decode16).
__host__ __device__ __forceinline__ uint32_t sgb_pack4(int g0, int g1, int g2, int g3)
{
    return (uint32_t)(g0 + 3 * g1) | ((uint32_t)(g2 + 3 * g3) << 4);
}
__host__ __device__ __forceinline__ void sgb_unpack_nibble(uint32_t n, int &a, int &b)
{
    b = (int)((n * 11u) >> 5);      // n / 3 for n in 0..8
    a = (int)n - 3 * b;
}

int k_grid_blocks(sgb_ctx *h, int64_t n)
{
    int64_t b = cdiv(n, 256 * 4);
    if (b > SGB_PART_BLOCKS) b = SGB_PART_BLOCKS;
    if (b < 1) b = 1;
    (void)h;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a 256-thread block; result valid in every thread.
__device__ __forceinline__ double block_sum_256(double v, double *sm /* >= 8 doubles */)
{
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sm[i];
    return t;
}

// fixed-order sum of nblk partials (every thread of the block computes the same value)
__device__ __forceinline__ double sum_partials(const double *part, int nblk)
{
    double t = 0.0;
    for (int i = 0; i < nblk; i++) t += part[i];
    return t;
}

// ---------------------------------------------------------------------------------------------------
// ingest: allele / missing counts over raw PLINK rows (Get_OneSNP_Geno_atBeginning, FG.cpp:338-435)
// ---------------------------------------------------------------------------------------------------
__global__ void count_markers_kernel(const uint8_t *__restrict__ bed, int64_t B0, int64_t nmark,
                                     const uint8_t *__restrict__ indmask, int32_t *__restrict__ ac,
                                     int32_t *__restrict__ nmiss)
{
    int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= nmark) return;
    int lane = threadIdx.x & 31;
    const uint8_t *row = bed + m * B0;
    int a = 0, ms = 0;
    for (int64_t b = lane; b < B0; b += 32) {
        uint32_t x = row[b], im = indmask[b];
        uint32_t lo = x & 0x55u, hi = (x >> 1) & 0x55u;
        uint32_t miss = lo & ~hi & im;              // code 01 : missing        (b=1,a=0 -> 3, FG.cpp:346)
        uint32_t c2 = ~lo & ~hi & 0x55u & im;       // code 00 : 2 copies of A1 (FG.cpp:348)
        uint32_t c1 = ~lo & hi & im;                // code 10 : 1 copy         (FG.cpp:350)
        a += 2 * __popc(c2) + __popc(c1);
        ms += __popc(miss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        ms += __shfl_xor_sync(0xffffffffu, ms, o);
    }
    if (lane == 0) { ac[m] = a; nmiss[m] = ms; }
}

int k_count_markers(sgb_ctx *h, const uint8_t *d_bed, int64_t B0, int64_t nmark, const uint8_t *d_indmask,
                    int32_t *d_ac, int32_t *d_nmiss)
{
    if (nmark <= 0) return 0;
    count_markers_kernel<<<(unsigned)cdiv(nmark, 8), 256, 0, h->stream>>>(d_bed, B0, nmark, d_indmask, d_ac, d_nmiss);
    LAUNCH_CHECK(h);
    return 0;
}

// re-pack kept markers in phenotype-sample order with imputation, into the device coding (FG.cpp:551-576)
__global__ void repack_kernel(const uint8_t *__restrict__ bed, int64_t B0, const int32_t *__restrict__ src_rows,
                              const int32_t *__restrict__ fill, const int32_t *__restrict__ sub_idx, int identity,
                              int64_t N, uint8_t *__restrict__ out, int64_t out_row0, int64_t out_stride, int tiled)
{
    int64_t r = blockIdx.y;
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t B = (N + 3) >> 2;
    if (b >= B) return;
    const uint8_t *row = bed + (int64_t)src_rows[r] * B0;
    int fl = fill[r];
    int gq[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int64_t k = 4 * b + j;
        if (k < N) {
            int64_t src = identity ? k : (int64_t)sub_idx[k] - 1;
            int code = (row[src >> 2] >> ((src & 3) << 1)) & 3;
            gq[j] = code == 0 ? 2 : (code == 2 ? 1 : (code == 3 ? 0 : fl));
        }
    }
    out[tiled ? sgb_tiled_off(out_row0 + r, b, out_stride) : (out_row0 + r) * out_stride + b] = (uint8_t)sgb_pack4(gq[0], gq[1], gq[2], gq[3]);
}

// gather variant: the raw row is staged in shared memory once, then every output byte gathers its 4 samples from it
__global__ void __launch_bounds__(256) repack_gather_kernel(const uint8_t *__restrict__ bed, int64_t B0,
                                                            const int32_t *__restrict__ src_rows, const int32_t *__restrict__ fill,
                                                            const int32_t *__restrict__ sub_idx, int64_t N,
                                                            uint8_t *__restrict__ out, int64_t out_row0, int64_t out_stride, int tiled)
{
    extern __shared__ uint8_t srow[];
    const int64_t r = blockIdx.x;
    const uint8_t *row = bed + (int64_t)src_rows[r] * B0;
    for (int64_t b = threadIdx.x; b < B0; b += blockDim.x) srow[b] = row[b];
    __syncthreads();
    const int fl = fill[r];
    const int64_t B = (N + 3) >> 2;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
        int gq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t k = 4 * b + j;
            if (k < N) {
                int64_t src = (int64_t)sub_idx[k] - 1;
                int code = (srow[src >> 2] >> ((src & 3) << 1)) & 3;
                gq[j] = code == 0 ? 2 : (code == 2 ? 1 : (code == 3 ? 0 : fl));
            }
        }
        out[tiled ? sgb_tiled_off(out_row0 + r, b, out_stride) : (out_row0 + r) * out_stride + b] = (uint8_t)sgb_pack4(gq[0], gq[1], gq[2], gq[3]);
    }
}

int k_repack(sgb_ctx *h, const uint8_t *d_bed, int64_t B0, const int32_t *d_src_rows, const int32_t *d_fill,
             int64_t nrows, const int32_t *d_sub_idx, int identity, int64_t N, uint8_t *d_out, int64_t out_row0, int64_t out_stride,
             int tiled)
{
    if (nrows <= 0) return 0;
    int64_t B = (N + 3) / 4;
    if (!identity && B0 <= 200 * 1024) {
        if (sgb_first_on_device(h->device, SGB_SITE_REPACK))
            CUDA_OK(h, cudaFuncSetAttribute(repack_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        repack_gather_kernel<<<(unsigned)nrows, 256, (size_t)B0, h->stream>>>(d_bed, B0, d_src_rows, d_fill, d_sub_idx, N, d_out, out_row0, out_stride, tiled);
        LAUNCH_CHECK(h);
        return 0;
    }
    for (int64_t r0 = 0; r0 < nrows; r0 += 65535) {
        int64_t nr = nrows - r0 < 65535 ? nrows - r0 : 65535;
        dim3 grid((unsigned)cdiv(B, 256), (unsigned)nr);
        repack_kernel<<<grid, 256, 0, h->stream>>>(d_bed, B0, d_src_rows + r0, d_fill + r0, d_sub_idx, identity, N,
                                                    d_out, out_row0 + r0, out_stride, tiled);
        LAUNCH_CHECK(h);
    }
    return 0;
}

// marker-major -> sample-major copy.  CTA tile: 128 markers x 256 samples = one (panel, slab) block of the tiled
// marker-major store, read as 8 KB of consecutive bytes.
__global__ void __launch_bounds__(256) transpose_kernel(const uint8_t *__restrict__ G, int64_t sG, int64_t Mloc,
                                                        uint8_t *__restrict__ Gt, int64_t sT, int64_t N)
{
    __shared__ uint8_t tile[128][64 + 4];
    int64_t m0 = (int64_t)blockIdx.y * 128, b0 = (int64_t)blockIdx.x * 64;   // b0: byte offset in a marker row
    for (int idx = threadIdx.x; idx < 128 * 16; idx += 256) {
        int r = idx >> 4, q = idx & 15;
        uint32_t v = 0;
        if (m0 + r < Mloc && b0 + 4 * q < sG) v = *reinterpret_cast<const uint32_t *>(G + sgb_tiled_off(m0 + r, b0 + 4 * q, sG));
        *reinterpret_cast<uint32_t *>(&tile[r][4 * q]) = v;
    }
    __syncthreads();
    int64_t i = b0 * 4 + threadIdx.x;      // sample
    if (i >= N) return;
    const int byte = threadIdx.x >> 2, nsh = (threadIdx.x & 2) << 1, odd = threadIdx.x & 1;
    uint32_t o[8];
#pragma unroll
    for (int w = 0; w < 8; w++) {
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            int a0, b0, a1, b1;
            sgb_unpack_nibble((tile[w * 16 + j][byte] >> nsh) & 15u, a0, b0);
            sgb_unpack_nibble((tile[w * 16 + j + 1][byte] >> nsh) & 15u, a1, b1);
            acc |= (uint32_t)((odd ? b0 : a0) + 3 * (odd ? b1 : a1)) << (2 * j);
        }
        o[w] = acc;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(Gt + sgb_tiled_off(i, m0 >> 2, sT));      // 32 bytes inside one 64-byte slab row
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

int k_transpose(sgb_ctx *h)
{
    if (h->Mloc <= 0) return 0;
    dim3 grid((unsigned)cdiv(h->sG, 64), (unsigned)cdiv(h->Mloc, 128));
    transpose_kernel<<<grid, 256, 0, h->stream>>>(h->dG, h->sG, h->Mloc, h->dGt, h->sT, h->N);
    LAUNCH_CHECK(h);
    return 0;
}

// synthetic genotypes (same integer hash as oracle/saige_oracle.c:mix32)
__device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t m, uint64_t i, uint64_t salt)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (m + 1) + 0xBF58476D1CE4E5B9ull * (i + 1) + 0x94D049BB133111EBull * salt;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

// mode 0: count allele copies of raw marker (blockIdx.y + y0) into ac[];  mode 1: write local row r = blockIdx.y + y0
// of the marker-major store for raw marker rows[r].
__global__ void synth_kernel(int mode, int64_t y0, uint64_t seed, const uint32_t *__restrict__ t0,
                             const uint32_t *__restrict__ t1, const int64_t *__restrict__ rows, int64_t N,
                             int32_t *__restrict__ ac, uint8_t *__restrict__ G, int64_t sG)
{
    int64_t r = blockIdx.y + y0;
    int64_t m = mode ? rows[r] : r;
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a0 = t0[m], a1 = t1[m];
    int cnt = 0;
    if (b < ((N + 3) >> 2)) {
        int gq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t i = 4 * b + j;
            if (i < N) {
                uint32_t u = mix32(seed, (uint64_t)m, (uint64_t)i, 0);
                gq[j] = (u >= a0) + (u >= a1);
                cnt += gq[j];
            }
        }
        if (mode) G[sgb_tiled_off(r, b, sG)] = (uint8_t)sgb_pack4(gq[0], gq[1], gq[2], gq[3]);
    }
    if (!mode) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, of);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&ac[m], cnt);
    }
}

// raw PLINK .bed rows (SNP-major, 00 = two copies of A1, 10 = one, 11 = none, 01 = missing) of the synthetic markers
// [m0, m0 + gridDim.y): the same integer hash and coding as oracle/saige_oracle.c:orc_synth_bed, one byte per thread
__global__ void synth_bed_kernel(int64_t m0, uint64_t seed, const uint32_t *__restrict__ t0, const uint32_t *__restrict__ t1,
                                 uint32_t miss_thr, int64_t N, uint8_t *__restrict__ out)
{
    const int64_t r = blockIdx.y, m = m0 + r, B0 = (N + 3) >> 2;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B0) return;
    const uint32_t a0 = t0[m], a1 = t1[m];
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int64_t i = 4 * b + j;
        if (i < N) {
            const uint32_t u = mix32(seed, (uint64_t)m, (uint64_t)i, 0);
            const int gv = (u >= a0) + (u >= a1);
            uint32_t code = gv == 2 ? 0u : (gv == 1 ? 2u : 3u);
            if (miss_thr && mix32(seed, (uint64_t)m, (uint64_t)i, 1) < miss_thr) code = 1u;
            v |= code << (2 * j);
        }
    }
    out[r * B0 + b] = (uint8_t)v;
}

int k_synth_bed(sgb_ctx *h, int64_t m0, int64_t nm, uint64_t seed, const uint32_t *d_t0, const uint32_t *d_t1, uint32_t miss_thr,
                int64_t N, uint8_t *d_out)
{
    const int64_t B0 = (N + 3) / 4;
    for (int64_t y0 = 0; y0 < nm; y0 += 65535) {
        const int64_t ny = nm - y0 < 65535 ? nm - y0 : 65535;
        synth_bed_kernel<<<dim3((unsigned)cdiv(B0, 256), (unsigned)ny), 256, 0, h->stream>>>(m0 + y0, seed, d_t0, d_t1, miss_thr, N, d_out + y0 * B0);
        LAUNCH_CHECK(h);
    }
    return 0;
}

int k_synth(sgb_ctx *h, uint64_t seed, const uint32_t *d_t0, const uint32_t *d_t1, int32_t *d_ac)
{
    // mode selected by d_ac: non-null => count pass over all M0 raw markers; null => fill local rows (h->ws holds rows)
    int64_t B = (h->N + 3) / 4;
    int64_t total = d_ac ? h->M0 : h->Mloc;
    for (int64_t y0 = 0; y0 < total; y0 += 65535) {
        int64_t ny = total - y0 < 65535 ? total - y0 : 65535;
        dim3 grid((unsigned)cdiv(B, 256), (unsigned)ny);
        synth_kernel<<<grid, 256, 0, h->stream>>>(d_ac ? 0 : 1, y0, seed, d_t0, d_t1, (const int64_t *)h->ws, h->N, d_ac,
                                                   h->dG, h->sG);
        LAUNCH_CHECK(h);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// tensor engine
// ---------------------------------------------------------------------------------------------------
// Decode one 32-bit word (16 genotypes in pair-ternary nibbles) into four registers of 4 x u8 with the byte
// permute unit.  A nibble n = A + 3B (0..8) is used UNMASKED as the prmt selector: selector values 0..7 pick a
// byte of the 8-byte pool, selector 8 (A = B = 2) means "replicate the sign of pool byte 0" = 0x00.  With
//   pool_A[idx] = c0 - plane(idx % 3)      pool_B[idx] = c0 - plane(idx / 3)
// (plane(g) = g with c0 = 2, or plane(g) = [g == 2] with c0 = 1) both the in-pool cases and the sign case give
// exactly c0 - plane(genotype), so no mask / shift of the 2-bit fields is needed at all:
//   d[0] = samples {0,2,4,6}   d[1] = {8,10,12,14}   d[2] = {1,3,5,7}   d[3] = {9,11,13,15}
// The products come out as sum (c0 - plane) * limb; recombine_kernel turns them back with the column limb sums.
struct pk2_pools { uint32_t ax, ay, bx, by; };

// raw PTX prmt (generic mode): selector bit 3 of a nibble = replicate the sign of the selected byte.  The CUDA
// intrinsic __byte_perm masks that bit away, so it cannot be used here.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ void decode16(uint32_t w, const pk2_pools &pl, uint32_t (&d)[4])
{
    uint32_t hi = w >> 16;
    d[0] = prmt(pl.ax, pl.ay, w);
    d[1] = prmt(pl.ax, pl.ay, hi);
    d[2] = prmt(pl.bx, pl.by, w);
    d[3] = prmt(pl.bx, pl.by, hi);
}

__device__ __forceinline__ void mma_u8s8(int32_t (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- mbarrier + bulk-copy (TMA engine, UBLKCP) primitives of the sweep kernel ----
__device__ __forceinline__ uint32_t smem_u32k(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_k(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32k(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_k(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32k(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_k(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32k(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_k(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32k(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_k(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32k(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32k(bar))
                 : "memory");
}

// out[r][(c0+n)*8+l] += sum_k P[r][k] * L[c0+n][k][l]     -- the GRM sweep for k <= 2 right-hand sides (HBM-bound)
//   P      : packed rows in the TILED layout (sgb_tiled_off): 128-row panels, 8 KB (panel, k-slab) blocks
//   L      : limb fragments, per column c: ksteps blocks of 2048 bytes; block = 256 genotypes x 8 limbs laid out
//            [mma pair(4)][lane(32)][mma of the pair(2)][b0/b1(2)][byte(4)]: one 16-byte load per lane and MMA pair
//            brings the B fragments {b0,b1} of MMAs 2q and 2q+1
// Persistent CTAs (2 per SM): 8 consumer warps + 1 producer warp.  Work unit = (256-row tile, group of KS k-slabs); the
// units of the whole matrix are dealt to the CTAs as contiguous ranges, so a CTA walks along the slabs of a tile and then
// on to the next tile, flushing its int32 accumulators with RED.ADD when the tile changes (integers: order independent).
// Producer: one lane issues cp.async.bulk copies -- per stage KS*8 KB CONTIGUOUS bytes of each of the two panels plus the
// limb fragments -- onto an mbarrier with expect_tx; STAGES-deep ring, "empty" barriers released by one arrive per warp.
// Consumers: conflict-free 128-bit LDS of (row g / g+8, bytes 16t..) = the A fragments' packed form, prmt decode, IMMA.
template <int NT, int KS, int STAGES>
__global__ void __launch_bounds__(288, 2)
pk2_stream_kernel(const uint8_t *__restrict__ P, int64_t stride, int64_t tiles, int64_t ksteps, const int8_t *__restrict__ L,
                  int64_t Lcol_stride, int c0, int ncol_total, int32_t *__restrict__ out, pk2_pools pool)
{
    constexpr int WARPS = 8, MT = 2, RT = WARPS * MT * 16, PANELS = RT / SGB_PANEL_ROWS;
    constexpr uint32_t A_STAGE = (uint32_t)PANELS * KS * SGB_SLAB_BYTES, B_STAGE = (uint32_t)NT * KS * 2048, STAGE = A_STAGE + B_STAGE;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t KG = (ksteps + KS - 1) / KS, units = tiles * KG;
    const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init_k(&full[i], 1); mbar_init_k(&empty[i], WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (warp == WARPS) {
        // ================= producer: one lane feeds the ring through the TMA engine =================
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int64_t u = u0; u < u1; u++) {
                const int64_t tile = u / KG, kg = u - tile * KG;
                const int nks = (int)((kg + 1) * KS <= ksteps ? KS : ksteps - kg * KS);
                if (u - u0 >= STAGES) mbar_wait_k(&empty[st], ph ^ 1);
                uint8_t *dst = smem + (size_t)st * STAGE;
                mbar_expect_tx_k(&full[st], (uint32_t)nks * (PANELS * SGB_SLAB_BYTES + NT * 2048));
#pragma unroll
                for (int p = 0; p < PANELS; p++)
                    bulk_g2s_k(dst + p * (KS * SGB_SLAB_BYTES), P + ((tile * PANELS + p) * SGB_PANEL_ROWS) * stride + kg * KS * SGB_SLAB_BYTES,
                               (uint32_t)nks * SGB_SLAB_BYTES, &full[st]);
#pragma unroll
                for (int n = 0; n < NT; n++)
                    bulk_g2s_k(dst + A_STAGE + n * (KS * 2048), L + (int64_t)(c0 + n) * Lcol_stride + kg * KS * 2048, (uint32_t)nks * 2048, &full[st]);
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        }
        return;
    }
    // ================= consumers =================
    int32_t acc[MT][NT][4];
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int n = 0; n < NT; n++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][n][c] = 0;
    const int r0 = warp * MT * 16, panel = r0 >> 7, rp = r0 & 127;
    const int g = lane >> 2, t = lane & 3;
    const int ncols8 = ncol_total * 8;
    // C fragment: c0 (row g, limb 2t) c1 (row g, limb 2t+1) c2 (row g+8, limb 2t) c3 (row g+8, limb 2t+1)
    auto flush = [&](int64_t tile) {
#pragma unroll
        for (int a = 0; a < MT; a++)
#pragma unroll
            for (int n = 0; n < NT; n++) {
                int32_t *o = out + (tile * RT + r0 + 16 * a + g) * ncols8 + (c0 + n) * 8 + 2 * t;
                if (acc[a][n][0]) atomicAdd(o, acc[a][n][0]);
                if (acc[a][n][1]) atomicAdd(o + 1, acc[a][n][1]);
                if (acc[a][n][2]) atomicAdd(o + 8 * ncols8, acc[a][n][2]);
                if (acc[a][n][3]) atomicAdd(o + 8 * ncols8 + 1, acc[a][n][3]);
                acc[a][n][0] = acc[a][n][1] = acc[a][n][2] = acc[a][n][3] = 0;
            }
    };
    int st = 0;
    uint32_t ph = 0;
    int64_t cur_tile = u0 < u1 ? u0 / KG : 0;
    for (int64_t u = u0; u < u1; u++) {
        const int64_t tile = u / KG, kg = u - tile * KG;
        const int nks = (int)((kg + 1) * KS <= ksteps ? KS : ksteps - kg * KS);
        if (tile != cur_tile) { flush(cur_tile); cur_tile = tile; }
        mbar_wait_k(&full[st], ph);
        const uint8_t *sa = smem + (size_t)st * STAGE + panel * (KS * SGB_SLAB_BYTES) + rp * 64 + lane * 16;
        const uint8_t *sb = smem + (size_t)st * STAGE + A_STAGE + lane * 16;
        for (int ks = 0; ks < nks; ks++) {
            uint4 bf[NT][4];
#pragma unroll
            for (int n = 0; n < NT; n++)
#pragma unroll
                for (int j = 0; j < 4; j++) bf[n][j] = *reinterpret_cast<const uint4 *>(sb + n * (KS * 2048) + ks * 2048 + j * 512);
#pragma unroll
            for (int a = 0; a < MT; a++) {
                const uint4 lo = *reinterpret_cast<const uint4 *>(sa + ks * SGB_SLAB_BYTES + (16 * a) * 64);
                const uint4 hi = *reinterpret_cast<const uint4 *>(sa + ks * SGB_SLAB_BYTES + (16 * a + 8) * 64);
                const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w}, wh[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int wi = 0; wi < 4; wi++) {
                    uint32_t dl[4], dh[4];
                    decode16(wl[wi], pool, dl);
                    decode16(wh[wi], pool, dh);
#pragma unroll
                    for (int n = 0; n < NT; n++) {
                        mma_u8s8(acc[a][n], dl[0], dh[0], dl[1], dh[1], bf[n][wi].x, bf[n][wi].y);
                        mma_u8s8(acc[a][n], dl[2], dh[2], dl[3], dh[3], bf[n][wi].z, bf[n][wi].w);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_k(&empty[st]);
        if (++st == STAGES) { st = 0; ph ^= 1; }
    }
    if (u0 < u1) flush(cur_tile);
}

template <int NT, int KS, int STAGES>
static int launch_pk2(sgb_ctx *h, int site, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t ksteps, const int8_t *L,
                      int64_t Lcol_stride, int c0, int ncol_total, int32_t *out, pk2_pools pool)
{
    constexpr int RT = 256;
    constexpr size_t smem = (size_t)STAGES * ((RT / SGB_PANEL_ROWS) * KS * SGB_SLAB_BYTES + NT * KS * 2048);
    auto kern = pk2_stream_kernel<NT, KS, STAGES>;
    if (sgb_first_on_device(h->device, site)) CUDA_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = rows_pad / RT, KG = (ksteps + KS - 1) / KS;
    int64_t grid = (int64_t)h->sm_count * 2;            // persistent: 2 CTAs per SM
    if (grid > tiles * KG) grid = tiles * KG;
    kern<<<(unsigned)grid, 288, smem, h->stream>>>(P, stride, tiles, ksteps, L, Lcol_stride, c0, ncol_total, out, pool);
    LAUNCH_CHECK(h);
    return 0;
}

int k_pk2_gemm(sgb_ctx *h, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t kbytes, const int8_t *L,
               int ncol, int32_t *out, int plane)
{
    // pool bytes idx 0..7: A-position value f(idx % 3), B-position value f(idx / 3); f(g) = 2-g or 1-[g==2]
    pk2_pools pool;
    if (plane == SGB_PLANE_VALUE) { pool.ax = 0x02000102u; pool.ay = 0x01020001u; pool.bx = 0x01020202u; pool.by = 0x00000101u; }
    else                          { pool.ax = 0x01000101u; pool.ay = 0x01010001u; pool.bx = 0x01010101u; pool.by = 0x00000101u; }
    if (rows_pad % SGB_ROW_ALIGN || kbytes != stride || stride % SGB_KSTEP_BYTES)
        return sgb_fail(h, "k_pk2_gemm: unaligned operand (rows %lld, kbytes %lld, stride %lld)", (long long)rows_pad,
                        (long long)kbytes, (long long)stride);
    int64_t ksteps = kbytes / SGB_KSTEP_BYTES;
    if (ksteps == 0 || rows_pad == 0 || ncol == 0) return 0;
    // int32 accumulation: |sum| <= 2 * 64 * (genotypes per row)
    if (kbytes * 4 > ((int64_t)1 << 24)) return sgb_fail(h, "k_pk2_gemm: %lld genotypes per row exceed the int32 accumulation bound", (long long)(kbytes * 4));
    int64_t Lcs = ksteps * 2048;
    int c = 0;
    while (c < ncol) {
        if (ncol - c >= 2) { SGB_TRY((launch_pk2<2, 2, 2>(h, SGB_SITE_STREAM2, P, stride, rows_pad, ksteps, L, Lcs, c, ncol, out, pool))); c += 2; }
        else { SGB_TRY((launch_pk2<1, 2, 3>(h, SGB_SITE_STREAM1, P, stride, rows_pad, ksteps, L, Lcs, c, ncol, out, pool))); c += 1; }
    }
    return 0;
}

// rows of a tiled matrix -> row-major [nrows][nbytes] (variance-ratio markers, Get_OneSNP_Geno)
__global__ void gather_rows_kernel(const uint8_t *__restrict__ P, int64_t stride, const int64_t *__restrict__ rows, int64_t nbytes,
                                   uint8_t *__restrict__ out)
{
    const int64_t r = rows[blockIdx.y];
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)blockIdx.y * nbytes + b] = P[sgb_tiled_off(r, b, stride)];
}

int k_gather_rows(sgb_ctx *h, const uint8_t *P, int64_t stride, const int64_t *d_rows, int nrows, int64_t nbytes, uint8_t *d_out)
{
    if (nrows <= 0 || nbytes <= 0) return 0;
    int gx = (int)cdiv(nbytes, 256);
    if (gx > 64) gx = 64;
    gather_rows_kernel<<<dim3(gx, nrows), 256, 0, h->stream>>>(P, stride, d_rows, nbytes, d_out);
    LAUNCH_CHECK(h);
    return 0;
}

// column max of |V| as the bit pattern of a non-negative double (monotone as uint64)
__global__ void colmax_kernel(const double *__restrict__ V, int64_t len, int64_t ld, unsigned long long *__restrict__ mx)
{
    int c = blockIdx.y;
    const double *v = V + (int64_t)c * ld;
    unsigned long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v[i]));
        m = b > m ? b : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o);
        m = x > m ? x : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&mx[c], m);
}

// Limb split.  One thread per group of 4 genotype slots that end up in the same 32-bit B-fragment register
// (slots r, r+2, r+4, r+6 of a 256-slot block): 8 balanced base-128 digits of round(v * 2^(53-E)), E = exponent of the
// column max, written as 8 words into the fragment layout of pk2_gemm_kernel; plus the exact column sums of every limb.
__global__ void split_limbs_kernel(const double *__restrict__ V, int64_t len, int64_t ld, int64_t nblk,
                                   const unsigned long long *__restrict__ mx, int8_t *__restrict__ L,
                                   double *__restrict__ mult, int32_t *__restrict__ limbsum)
{
    int c = blockIdx.y;
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mb = mx[c];
    int E = (int)((mb >> 52) & 0x7FF) - 1023;
    if (E < -1000) E = -1000;
    if (E > 1000) E = 1000;            // inf/nan columns produce garbage, like any fp arithmetic would
    if (u == 0) mult[c] = mb ? scalbn(1.0, E - 53) : 0.0;
    const int64_t blk = u >> 6;
    const int v = (int)(u & 63), t = v >> 4, wi = (v >> 2) & 3, half = (v >> 1) & 1, odd = v & 1;
    const bool inb = blk < nblk;
    const int64_t i0 = blk * 256 + t * 64 + wi * 16 + half * 8 + odd;
    long long q[4];
#pragma unroll
    for (int sl = 0; sl < 4; sl++) {
        int64_t i = i0 + 2 * sl;
        // |v| < 2^(E+1)  =>  |q| <= 2^54, inside the range of 8 balanced base-128 digits (|.| <= 63*(128^8-1)/127 ~ 2^54.99)
        q[sl] = (inb && i < len && mb) ? __double2ll_rn(scalbn(V[(int64_t)c * ld + i], 53 - E)) : 0;
    }
    uint32_t *base = reinterpret_cast<uint32_t *>(L + ((int64_t)c * nblk + blk) * 2048 + (wi * 32 + t) * 16 + odd * 8 + half * 4);
#pragma unroll
    for (int l = 0; l < SGB_LIMBS; l++) {
        uint32_t word = 0;
        int ssum = 0;
#pragma unroll
        for (int sl = 0; sl < 4; sl++) {
            int d = (int)((q[sl] + 64) & 127) - 64;
            q[sl] = (q[sl] - d) >> 7;
            word |= (uint32_t)(d & 255) << (8 * sl);
            ssum += d;
        }
        if (inb) base[l * 16] = word;             // lane = l*4 + t  ->  +l*4 lanes * 16 bytes
        // column sum of this limb (exact integers): undoes the c0 - plane offset of the decode in recombine_kernel
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        if ((threadIdx.x & 31) == 0 && ssum) atomicAdd(&limbsum[c * SGB_LIMBS + l], ssum);
    }
}

int k_split_limbs(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, int8_t *L, int64_t nblk, double *d_mult,
                  int32_t *d_limbsum, int have_stats)
{
    unsigned long long *mx = reinterpret_cast<unsigned long long *>(h->d_scal + 2048);   // k <= 1024 slots
    if (k > 1024) return sgb_fail(h, "too many columns (%d)", k);
    if (!have_stats) {          // the fused product path (k_col_stats / k_recomb_post1) has left the column maxima and zeroed limb sums
        CUDA_OK(h, cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * k, h->stream));
        int gx = (int)cdiv(len, 256 * 8);
        if (gx > 1024) gx = 1024;
        if (gx < 1) gx = 1;
        colmax_kernel<<<dim3(gx, k), 256, 0, h->stream>>>(V, len, ld, mx);
        LAUNCH_CHECK(h);
        CUDA_OK(h, cudaMemsetAsync(d_limbsum, 0, sizeof(int32_t) * SGB_LIMBS * k, h->stream));
    }
    split_limbs_kernel<<<dim3((unsigned)cdiv(nblk, 4), k), 256, 0, h->stream>>>(V, len, ld, nblk, mx, L, d_mult, d_limbsum);
    LAUNCH_CHECK(h);
    return 0;
}

// raw[r + c*ld] = (sum_l acc[r][c*8+l] * 128^l) * mult[c];  acc is reset to 0 for the next product.
__global__ void recombine_kernel(int32_t *__restrict__ acc, int64_t rows, int k, int kpad, const double *__restrict__ mult,
                                 const int32_t *__restrict__ limbsum, int c0, double *__restrict__ raw, int64_t ld)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * k) return;
    int64_t r = idx / k;
    int c = (int)(idx - r * k);
    raw[r + (int64_t)c * ld] = sgb_recombine_imma(acc, r, c, kpad, limbsum, c0) * mult[c];
}

int k_recombine(sgb_ctx *h, int32_t *acc, int64_t rows, int k, int kpad, const double *d_mult, const int32_t *d_limbsum,
                int plane, double *raw, int64_t ld)
{
    if (rows * k == 0) return 0;
    recombine_kernel<<<(unsigned)cdiv(rows * k, 256), 256, 0, h->stream>>>(acc, rows, k, kpad, d_mult, d_limbsum,
                                                                           plane == SGB_PLANE_VALUE ? 2 : 1, raw, ld);
    LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// fused small kernels of a product: 7 launches per k-column product instead of ~19 (at 8 GPUs the dozen 5 us kernels
// around the two sweeps were 12 % of a product).  Column statistics finish inside the kernel that produces them: every
// block leaves its partial (sum, max |.|), takes a ticket, and the LAST block of a column reduces the partials in fixed
// block order -- the same order finish_partials_kernel uses, so results are bit-identical to the unfused path.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_max_256(unsigned long long m, unsigned long long *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o);
        m = x > m ? x : m;
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = m;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t = sm[i] > t ? sm[i] : t;
    return t;
}

// tail of a statistics kernel: returns true in the last block of column c (all threads), after which psum/pmax of the
// column are complete and visible
__device__ __forceinline__ bool stats_ticket(double s, unsigned long long m, int c, double *psum, unsigned long long *pmax,
                                             unsigned int *ticket, int *last_flag)
{
    if (threadIdx.x == 0) {
        psum[(int64_t)c * SGB_PART_BLOCKS + blockIdx.x] = s;
        pmax[(int64_t)c * SGB_PART_BLOCKS + blockIdx.x] = m;
        __threadfence();
        unsigned int t = atomicAdd(&ticket[c], 1u);
        *last_flag = (t == gridDim.x - 1);
    }
    __syncthreads();
    return *last_flag != 0;
}
// Last block of a column: the partials of all blocks come in with ONE load per thread (a single thread walking the list paid a
// full L2 round trip per element: ~0.13 ms for 256 partials, more than the rest of a narrow product's epilogue), then thread 0 adds
// them in block order (the fixed order every engine and batch width shares) and thread 32 takes the maximum.
// Returns the sum in thread 0 and the maximum in thread 32.  fin / finm: SGB_PART_BLOCKS elements of shared memory each.
__device__ __forceinline__ void finish_column_partials(const double *psum_c, const unsigned long long *pmax_c, int nblk, double *fin,
                                                       unsigned long long *finm, double *tot, unsigned long long *mxv)
{
    __syncthreads();                                                    // fin / finm may still be read by the previous column
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
        fin[i] = reinterpret_cast<const volatile double *>(psum_c)[i];
        finm[i] = reinterpret_cast<const volatile unsigned long long *>(pmax_c)[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < nblk; i++) t += fin[i];
        *tot = t;
    }
    if (threadIdx.x == 32) {
        unsigned long long t = 0;
        for (int i = 0; i < nblk; i++) t = finm[i] > t ? finm[i] : t;
        *mxv = t;
    }
}

// colsum[c] = sum V[:,c]; mx[c] = bits(max |V[:,c]|); the limb sums of the coming split are zeroed
__global__ void __launch_bounds__(256) col_stats_kernel(const double *__restrict__ V, int64_t len, int64_t ld, double *__restrict__ psum,
                                                        unsigned long long *__restrict__ pmax, unsigned int *__restrict__ ticket,
                                                        double *__restrict__ colsum, unsigned long long *__restrict__ mx,
                                                        int32_t *__restrict__ limbsum, int nlimb)
{
    __shared__ double sm[8];
    __shared__ unsigned long long smx[8];
    __shared__ int last;
    const int c = blockIdx.y;
    const double *v = V + (int64_t)c * ld;
    double s = 0.0;
    unsigned long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        double x = v[i];
        s += x;
        unsigned long long b = (unsigned long long)__double_as_longlong(fabs(x));
        m = b > m ? b : m;
    }
    s = block_sum_256(s, sm);
    m = block_max_256(m, smx);
    if (!stats_ticket(s, m, c, psum, pmax, ticket, &last)) return;
    __shared__ double fin[SGB_PART_BLOCKS];
    __shared__ unsigned long long finm[SGB_PART_BLOCKS];
    double tot = 0.0;
    unsigned long long mxv = 0;
    finish_column_partials(psum + (int64_t)c * SGB_PART_BLOCKS, pmax + (int64_t)c * SGB_PART_BLOCKS, gridDim.x, fin, finm, &tot, &mxv);
    if (threadIdx.x == 0) { colsum[c] = tot; ticket[c] = 0; }
    if (threadIdx.x == 32) mx[c] = mxv;
    if (threadIdx.x >= 64 && threadIdx.x < 64 + nlimb) limbsum[c * nlimb + threadIdx.x - 64] = 0;
}

int k_col_stats(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, double *d_colsum, int32_t *d_limbsum, int nlimb)
{
    unsigned long long *mx = reinterpret_cast<unsigned long long *>(h->d_scal + 2048);
    int nb = k_grid_blocks(h, len);
    col_stats_kernel<<<dim3(nb, k), 256, 0, h->stream>>>(V, len, ld, h->d_red, reinterpret_cast<unsigned long long *>(h->d_red) + 1024 * SGB_PART_BLOCKS,
                                                         h->d_ticket, d_colsum, mx, d_limbsum, nlimb);
    LAUNCH_CHECK(h);
    return 0;
}

// sweep-1 epilogue: D[m,c] = s_m^2 (recombine(acc[m,c]) mult_c - 2f_m sumb_c), zeroed inside [mask_lo, mask_hi) (the
// left-out chromosome); t_c = sum_m 2f_m D[m,c]; mx[c] = bits(max |D[:,c]|); limb sums of the sweep-2 split zeroed;
// the accumulators (padding rows included) are reset.  NL = 8: mma.sync engine, 5..7: tcgen05 engine.
template <int NL>
__global__ void __launch_bounds__(256) recomb_post1_kernel(int32_t *__restrict__ acc, int64_t rows_pad, int64_t Mloc, int pad,
                                                           const double *__restrict__ mult, const int32_t *__restrict__ limbsum, int c0,
                                                           const double *__restrict__ f2, const double *__restrict__ s2,
                                                           const double *__restrict__ colsum, int64_t mask_lo, int64_t mask_hi,
                                                           double *__restrict__ D, int64_t ld, double *__restrict__ psum,
                                                           unsigned long long *__restrict__ pmax, unsigned int *__restrict__ ticket,
                                                           double *__restrict__ t_out, double *__restrict__ t_out2,
                                                           unsigned long long *__restrict__ mx, int32_t *__restrict__ limbsum2, int nlimb2)
{
    __shared__ double sm[8];
    __shared__ unsigned long long smx[8];
    __shared__ int last;
    const int c = blockIdx.y;
    const double sb = colsum[c], mu = mult[c];
    double t = 0.0;
    unsigned long long m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows_pad; r += (int64_t)gridDim.x * blockDim.x) {
        double v = sgb_recombine<NL>(acc, r, c, pad, limbsum, c0) * mu;      // padding rows: zeroes their accumulators
        if (r < Mloc) {
            double d = s2[r] * (v - f2[r] * sb);
            if (r >= mask_lo && r < mask_hi) d = 0.0;
            D[r + (int64_t)c * ld] = d;
            t += f2[r] * d;
            unsigned long long b = (unsigned long long)__double_as_longlong(fabs(d));
            m = b > m ? b : m;
        }
    }
    t = block_sum_256(t, sm);
    m = block_max_256(m, smx);
    if (!stats_ticket(t, m, c, psum, pmax, ticket, &last)) return;
    __shared__ double fin[SGB_PART_BLOCKS];
    __shared__ unsigned long long finm[SGB_PART_BLOCKS];
    double tt = 0.0;
    unsigned long long mxv = 0;
    finish_column_partials(psum + (int64_t)c * SGB_PART_BLOCKS, pmax + (int64_t)c * SGB_PART_BLOCKS, gridDim.x, fin, finm, &tt, &mxv);
    if (threadIdx.x == 0) {
        t_out[c] = tt;
        if (t_out2) t_out2[c] = tt;
        ticket[c] = 0;
    }
    if (threadIdx.x == 32) mx[c] = mxv;
    if (threadIdx.x >= 64 && threadIdx.x < 64 + nlimb2) limbsum2[c * nlimb2 + threadIdx.x - 64] = 0;
}

// The same epilogue for the tcgen05 engine (NL = 5..7), TILED through shared memory: a block owns SGB_RCG consecutive columns and
// moves a 256-row x (NL SGB_RCG)-word tile of the accumulators with 128-bit loads along the rows (<= 224 contiguous bytes per row),
// zeroing it behind the loads, instead of NL scalar loads + NL scalar stores per (row, column) at a stride of one accumulator
// row -- the column-per-block version spent 1.7 ms on a 31-column batch (0.45 GB of accumulators), ten times its HBM time.
// Thread tid then finishes row tid of the tile column by column from shared memory (row stride NL SGB_RCG + 1 words: no bank
// conflicts).  Rows -> threads -> blocks and the order of every sum are those of recomb_post1_kernel, so t_c and the column
// maxima come out bit-identical.
template <int NL>
__global__ void __launch_bounds__(256) recomb_post1_wide_kernel(int32_t *__restrict__ acc, int64_t rows_pad, int64_t Mloc, int k, int pad,
                                                                const double *__restrict__ mult, const int32_t *__restrict__ limbsum, int c0,
                                                                const double *__restrict__ f2, const double *__restrict__ s2,
                                                                const double *__restrict__ colsum, int64_t mask_lo, int64_t mask_hi,
                                                                double *__restrict__ D, int64_t ld, double *__restrict__ psum,
                                                                unsigned long long *__restrict__ pmax, unsigned int *__restrict__ ticket,
                                                                double *__restrict__ t_out, double *__restrict__ t_out2,
                                                                unsigned long long *__restrict__ mx, int32_t *__restrict__ limbsum2, int nlimb2)
{
    constexpr int WP = NL * SGB_RCG + 1;
    extern __shared__ int32_t tile[];                                   // 256 rows x WP words
    __shared__ double sm[8];
    __shared__ unsigned long long smx[8];
    __shared__ int last;
    __shared__ int32_t ls[NL * SGB_RCG];
    __shared__ double sbs[SGB_RCG], mus[SGB_RCG];
    __shared__ double fin[SGB_PART_BLOCKS];
    __shared__ unsigned long long finm[SGB_PART_BLOCKS];
    const int cg0 = blockIdx.y * SGB_RCG, nc = k - cg0 < SGB_RCG ? k - cg0 : SGB_RCG;
    const int w0 = NL * cg0, w1 = cg0 + SGB_RCG >= k ? pad : NL * (cg0 + SGB_RCG);
    const int W4 = (w1 - w0) >> 2 < 2 * NL ? (w1 - w0) >> 2 : 2 * NL;   // 16-byte pieces per tile row (padding columns past NL SGB_RCG are never written)
    if (threadIdx.x < NL * SGB_RCG) ls[threadIdx.x] = threadIdx.x < NL * nc ? limbsum[w0 + threadIdx.x] : 0;
    if (threadIdx.x < SGB_RCG) {
        sbs[threadIdx.x] = threadIdx.x < nc ? colsum[cg0 + threadIdx.x] : 0.0;
        mus[threadIdx.x] = threadIdx.x < nc ? mult[cg0 + threadIdx.x] : 0.0;
    }
    double t[SGB_RCG];
    unsigned long long m[SGB_RCG];
#pragma unroll
    for (int cc = 0; cc < SGB_RCG; cc++) { t[cc] = 0.0; m[cc] = 0; }
    for (int64_t rb = (int64_t)blockIdx.x * 256; rb < rows_pad; rb += (int64_t)gridDim.x * 256) {
        __syncthreads();                                                // the previous tile has been consumed (and ls / sbs / mus are visible)
        // all loads of a thread first (independent, one DRAM round trip), then the shared-memory stores and the zeroing: with a
        // load -> zero -> stash sequence per piece the in-order issue made every piece wait for the one before (ncu: 3 % issue
        // slots used, long-scoreboard stalls, 3.2 ms for a 31-column batch)
        int4 v[2 * NL];
#pragma unroll
        for (int i = 0; i < 2 * NL; i++)
            if (i < W4) {
                const int idx = threadIdx.x + 256 * i, row = idx / W4, q = idx - row * W4;
                v[i] = *(reinterpret_cast<const int4 *>(acc + (rb + row) * pad + w0) + q);
            }
#pragma unroll
        for (int i = 0; i < 2 * NL; i++)
            if (i < W4) {
                const int idx = threadIdx.x + 256 * i, row = idx / W4, q = idx - row * W4;
                *(reinterpret_cast<int4 *>(acc + (rb + row) * pad + w0) + q) = make_int4(0, 0, 0, 0);
                int32_t *d = tile + row * WP + 4 * q;
                d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w;
            }
        __syncthreads();
        const int64_t r = rb + threadIdx.x;
        if (r < Mloc) {
            const int32_t *p = tile + threadIdx.x * WP;
            const double f = f2[r], sc = s2[r];
            const bool masked = r >= mask_lo && r < mask_hi;
#pragma unroll
            for (int cc = 0; cc < SGB_RCG; cc++)
                if (cc < nc) {
                    const double v = sgb_umma_value<NL>(p + NL * cc, ls + NL * cc, c0) * mus[cc];
                    double d = sc * (v - f * sbs[cc]);
                    if (masked) d = 0.0;
                    D[r + (int64_t)(cg0 + cc) * ld] = d;
                    t[cc] += f * d;
                    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(d));
                    m[cc] = b > m[cc] ? b : m[cc];
                }
        }
    }
#pragma unroll
    for (int cc = 0; cc < SGB_RCG; cc++) {
        if (cc >= nc) break;                                            // uniform over the block
        const int c = cg0 + cc;
        const double tt = block_sum_256(t[cc], sm);
        const unsigned long long mm = block_max_256(m[cc], smx);
        if (stats_ticket(tt, mm, c, psum, pmax, ticket, &last)) {
            double tot = 0.0;
            unsigned long long mxv = 0;
            finish_column_partials(psum + (int64_t)c * SGB_PART_BLOCKS, pmax + (int64_t)c * SGB_PART_BLOCKS, gridDim.x, fin, finm, &tot, &mxv);
            if (threadIdx.x == 0) {
                t_out[c] = tot;
                if (t_out2) t_out2[c] = tot;
                ticket[c] = 0;
            }
            if (threadIdx.x == 32) mx[c] = mxv;
            if (threadIdx.x >= 64 && threadIdx.x < 64 + nlimb2) limbsum2[c * nlimb2 + threadIdx.x - 64] = 0;
        }
    }
}

int k_recomb_post1(sgb_ctx *h, int nl, int32_t *acc, int64_t rows_pad, int k, int pad, const double *d_mult, const int32_t *d_limbsum,
                   const double *d_colsum, int64_t mask_lo, int64_t mask_hi, double *D, int64_t ld, double *d_t, double *d_t2,
                   int32_t *d_limbsum2, int nlimb2)
{
    unsigned long long *mx = reinterpret_cast<unsigned long long *>(h->d_scal + 2048);
    double *psum = h->d_red;
    unsigned long long *pmax = reinterpret_cast<unsigned long long *>(h->d_red) + 1024 * SGB_PART_BLOCKS;
    // the partial-sum pattern over the first Mloc rows must not depend on the padding: blocks as for Mloc elements
    int nb = k_grid_blocks(h, h->Mloc);
    dim3 grid(nb, k), gridw(nb, (k + SGB_RCG - 1) / SGB_RCG);
#define RP1(NLV) recomb_post1_kernel<NLV><<<grid, 256, 0, h->stream>>>(acc, rows_pad, h->Mloc, pad, d_mult, d_limbsum, 2, h->d_f2, h->d_s2, d_colsum, \
                                                                       mask_lo, mask_hi, D, ld, psum, pmax, h->d_ticket, d_t, d_t2, mx, d_limbsum2, nlimb2)
#define RP1W(NLV) do { \
        constexpr size_t sb_ = (size_t)256 * (NLV * SGB_RCG + 1) * sizeof(int32_t); \
        if (sgb_first_on_device(h->device, SGB_SITE_POST1_5 + NLV - 5)) \
            CUDA_OK(h, cudaFuncSetAttribute(recomb_post1_wide_kernel<NLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb_)); \
        recomb_post1_wide_kernel<NLV><<<gridw, 256, sb_, h->stream>>>(acc, rows_pad, h->Mloc, k, pad, d_mult, d_limbsum, 2, h->d_f2, h->d_s2, d_colsum, \
                                                                      mask_lo, mask_hi, D, ld, psum, pmax, h->d_ticket, d_t, d_t2, mx, d_limbsum2, nlimb2); \
    } while (0)
    switch (nl) { case 8: RP1(8); break; case 7: RP1W(7); break; case 6: RP1W(6); break; case 5: RP1W(5); break;
                  default: return sgb_fail(h, "recomb_post1: unsupported limb count %d", nl); }
#undef RP1
#undef RP1W
    LAUNCH_CHECK(h);
    return 0;
}

// sweep-2 epilogue.  Y != null: Y[i,c] = (recombine(acc[i,c]) mult_c - t_c) inv_m  (single GPU);  else raw[i,c] = recombine mult_c
template <int NL>
__global__ void __launch_bounds__(256) recomb_post2_kernel(int32_t *__restrict__ acc, int64_t rows_pad, int64_t N, int pad,
                                                           const double *__restrict__ mult, const int32_t *__restrict__ limbsum, int c0,
                                                           const double *__restrict__ t, double inv_m, double *__restrict__ Y, int64_t ldy,
                                                           double *__restrict__ raw, int64_t ldr)
{
    const int c = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_pad) return;
    double v = sgb_recombine<NL>(acc, i, c, pad, limbsum, c0) * mult[c];
    if (Y) { if (i < N) Y[i + (int64_t)c * ldy] = (v - t[c]) * inv_m; }
    else raw[i + (int64_t)c * ldr] = v;
}

// tiled version for the tcgen05 engine (see recomb_post1_wide_kernel): one 256-row tile per block
template <int NL>
__global__ void __launch_bounds__(256) recomb_post2_wide_kernel(int32_t *__restrict__ acc, int64_t rows_pad, int64_t N, int k, int pad,
                                                                const double *__restrict__ mult, const int32_t *__restrict__ limbsum, int c0,
                                                                const double *__restrict__ t, double inv_m, double *__restrict__ Y, int64_t ldy,
                                                                double *__restrict__ raw, int64_t ldr)
{
    constexpr int WP = NL * SGB_RCG + 1;
    extern __shared__ int32_t tile[];                                   // 256 rows x WP words
    __shared__ int32_t ls[NL * SGB_RCG];
    const int cg0 = blockIdx.y * SGB_RCG, nc = k - cg0 < SGB_RCG ? k - cg0 : SGB_RCG;
    const int w0 = NL * cg0, w1 = cg0 + SGB_RCG >= k ? pad : NL * (cg0 + SGB_RCG);
    const int W4 = (w1 - w0) >> 2 < 2 * NL ? (w1 - w0) >> 2 : 2 * NL;
    if (threadIdx.x < NL * SGB_RCG) ls[threadIdx.x] = threadIdx.x < NL * nc ? limbsum[w0 + threadIdx.x] : 0;
    const int64_t rb = (int64_t)blockIdx.x * 256;                       // rows_pad is a multiple of 256
    int4 v[2 * NL];
#pragma unroll
    for (int i = 0; i < 2 * NL; i++)
        if (i < W4) {
            const int idx = threadIdx.x + 256 * i, row = idx / W4, q = idx - row * W4;
            v[i] = *(reinterpret_cast<const int4 *>(acc + (rb + row) * pad + w0) + q);
        }
#pragma unroll
    for (int i = 0; i < 2 * NL; i++)
        if (i < W4) {
            const int idx = threadIdx.x + 256 * i, row = idx / W4, q = idx - row * W4;
            *(reinterpret_cast<int4 *>(acc + (rb + row) * pad + w0) + q) = make_int4(0, 0, 0, 0);
            int32_t *d = tile + row * WP + 4 * q;
            d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w;
        }
    __syncthreads();
    const int64_t i = rb + threadIdx.x;
    const int32_t *p = tile + threadIdx.x * WP;
#pragma unroll
    for (int cc = 0; cc < SGB_RCG; cc++)
        if (cc < nc) {
            const int c = cg0 + cc;
            const double val = sgb_umma_value<NL>(p + NL * cc, ls + NL * cc, c0) * mult[c];
            if (Y) { if (i < N) Y[i + (int64_t)c * ldy] = (val - t[c]) * inv_m; }
            else raw[i + (int64_t)c * ldr] = val;
        }
}

int k_recomb_post2(sgb_ctx *h, int nl, int32_t *acc, int64_t rows_pad, int k, int pad, const double *d_mult, const int32_t *d_limbsum,
                   const double *d_t, double inv_m, double *Y, int64_t ldy, double *raw, int64_t ldr)
{
    dim3 grid((unsigned)cdiv(rows_pad, 256), k), gridw((unsigned)cdiv(rows_pad, 256), (k + SGB_RCG - 1) / SGB_RCG);
    if (nl != 8 && rows_pad % 256) return sgb_fail(h, "recomb_post2: %lld accumulator rows are not a multiple of 256", (long long)rows_pad);
#define RP2(NLV) recomb_post2_kernel<NLV><<<grid, 256, 0, h->stream>>>(acc, rows_pad, h->N, pad, d_mult, d_limbsum, 2, d_t, inv_m, Y, ldy, raw, ldr)
#define RP2W(NLV) do { \
        constexpr size_t sb_ = (size_t)256 * (NLV * SGB_RCG + 1) * sizeof(int32_t); \
        if (sgb_first_on_device(h->device, SGB_SITE_POST2_5 + NLV - 5)) \
            CUDA_OK(h, cudaFuncSetAttribute(recomb_post2_wide_kernel<NLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb_)); \
        recomb_post2_wide_kernel<NLV><<<gridw, 256, sb_, h->stream>>>(acc, rows_pad, h->N, k, pad, d_mult, d_limbsum, 2, d_t, inv_m, Y, ldy, raw, ldr); \
    } while (0)
    switch (nl) { case 8: RP2(8); break; case 7: RP2W(7); break; case 6: RP2W(6); break; case 5: RP2W(5); break;
                  default: return sgb_fail(h, "recomb_post2: unsupported limb count %d", nl); }
#undef RP2
#undef RP2W
    LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// f64 engine (slow on-device cross-check of the tensor engine; plain fp64 FMA, no tensor cores)
// ---------------------------------------------------------------------------------------------------
// out[m + c*ldo] = sum_i g_mi B[i + c*ldb]      one warp per marker
__global__ void rowdot_f64_kernel(const uint8_t *__restrict__ G, int64_t sG, int64_t Mloc, int64_t N,
                                  const double *__restrict__ B, int64_t ldb, int k, double *__restrict__ out, int64_t ldo)
{
    int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= Mloc) return;
    int lane = threadIdx.x & 31;
    int64_t nw = (N + 15) >> 4;
    for (int c = 0; c < k; c++) {
        const double *b = B + (int64_t)c * ldb;
        double s1 = 0.0, s2 = 0.0;
        for (int64_t w = lane; w < nw; w += 32) {
            uint32_t x = *reinterpret_cast<const uint32_t *>(G + sgb_tiled_off(m, 4 * w, sG));
            int64_t i0 = w << 4;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int ga, gb;
                sgb_unpack_nibble((x >> (4 * j)) & 15u, ga, gb);
                int64_t i = i0 + 2 * j;
                if (ga && i < N) { double bv = b[i]; if (ga == 1) s1 += bv; else s2 += bv; }
                if (gb && i + 1 < N) { double bv = b[i + 1]; if (gb == 1) s1 += bv; else s2 += bv; }
            }
        }
        double s = warp_sum(s1 + 2.0 * s2);
        if (lane == 0) out[m + (int64_t)c * ldo] = s;
    }
}

int k_rowdot_f64(sgb_ctx *h, const double *B, int64_t ldb, int k, double *out, int64_t ldo)
{
    if (h->Mloc == 0) return 0;
    rowdot_f64_kernel<<<(unsigned)cdiv(h->Mloc, 8), 256, 0, h->stream>>>(h->dG, h->sG, h->Mloc, h->N, B, ldb, k, out, ldo);
    LAUNCH_CHECK(h);
    return 0;
}

// out[i + c*ldo] = sum_m (g_mi==1 ? D1[m,c] : g_mi==2 ? D2[m,c] : 0)   thread per 16 samples, marker chunks + atomics
__global__ void coldot_f64_kernel(const uint8_t *__restrict__ G, int64_t sG, int64_t Mloc, int64_t N,
                                  const double *__restrict__ D1, const double *__restrict__ D2, int64_t ldd, int c,
                                  int64_t mchunk, double *__restrict__ out, int64_t ldo)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t nw = (N + 15) >> 4;
    if (w >= nw) return;
    int64_t m0 = (int64_t)blockIdx.y * mchunk, m1 = m0 + mchunk;
    if (m1 > Mloc) m1 = Mloc;
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.0;
    const double *d1 = D1 + (int64_t)c * ldd, *d2 = D2 ? D2 + (int64_t)c * ldd : nullptr;
    for (int64_t m = m0; m < m1; m++) {
        uint32_t x = *reinterpret_cast<const uint32_t *>(G + sgb_tiled_off(m, 4 * w, sG));
        if (!x) continue;
        double a1 = d1[m], a2 = D2 ? d2[m] : 2.0 * a1;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int ga, gb;
            sgb_unpack_nibble((x >> (4 * j)) & 15u, ga, gb);
            acc[2 * j] += ga == 1 ? a1 : (ga == 2 ? a2 : 0.0);
            acc[2 * j + 1] += gb == 1 ? a1 : (gb == 2 ? a2 : 0.0);
        }
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int64_t i = (w << 4) + j;
        if (i < N && acc[j] != 0.0) atomicAdd(out + i + (int64_t)c * ldo, acc[j]);
    }
}

int k_coldot_f64(sgb_ctx *h, const double *D1, const double *D2, int64_t ldd, int k, double *out, int64_t ldo)
{
    for (int c = 0; c < k; c++) CUDA_OK(h, cudaMemsetAsync(out + (int64_t)c * ldo, 0, sizeof(double) * h->N, h->stream));
    if (h->Mloc == 0) return 0;
    int64_t nw = (h->N + 15) / 16;
    int64_t gx = cdiv(nw, 128);
    int64_t chunks = cdiv((int64_t)h->sm_count * 8, gx);
    if (chunks < 1) chunks = 1;
    if (chunks > 4096) chunks = 4096;
    int64_t mchunk = cdiv(h->Mloc, chunks);
    chunks = cdiv(h->Mloc, mchunk);
    for (int c = 0; c < k; c++) {
        coldot_f64_kernel<<<dim3((unsigned)gx, (unsigned)chunks), 128, 0, h->stream>>>(h->dG, h->sG, h->Mloc, h->N, D1, D2,
                                                                                       ldd, c, mchunk, out, ldo);
        LAUNCH_CHECK(h);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// algebra around the sweeps (standardisation folded in: z_m = (g_m - 2f_m) s_m)
// ---------------------------------------------------------------------------------------------------
__global__ void partial_colsum_kernel(const double *__restrict__ V, int64_t len, int64_t ld, double *__restrict__ part)
{
    __shared__ double sm[8];
    int c = blockIdx.y;
    const double *v = V + (int64_t)c * ld;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) s += v[i];
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) part[(int64_t)c * SGB_PART_BLOCKS + blockIdx.x] = s;
}

__global__ void finish_partials_kernel(const double *__restrict__ part, int nblk, double *__restrict__ out)
{
    int c = threadIdx.x;
    out[c] = sum_partials(part + (int64_t)c * SGB_PART_BLOCKS, nblk);
}

int k_colsum(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, double *d_out)
{
    double *part = reinterpret_cast<double *>(h->ws);
    int nb = k_grid_blocks(h, len);
    partial_colsum_kernel<<<dim3(nb, k), 256, 0, h->stream>>>(V, len, ld, part);
    LAUNCH_CHECK(h);
    finish_partials_kernel<<<1, k, 0, h->stream>>>(part, nb, d_out);
    LAUNCH_CHECK(h);
    return 0;
}

// D[m,c] = s_m^2 (raw[m,c] - 2f_m sumb_c), zeroed inside [mask_lo, mask_hi) (the left-out chromosome);
// part_t[c][blk] = partial of sum_m 2f_m D[m,c]
__global__ void sweep1_post_kernel(const double *__restrict__ raw, int64_t ld, int64_t Mloc, const double *__restrict__ f2,
                                   const double *__restrict__ s2, const double *__restrict__ colsum, int64_t mask_lo,
                                   int64_t mask_hi, double *__restrict__ D, double *__restrict__ part)
{
    __shared__ double sm[8];
    int c = blockIdx.y;
    double sb = colsum[c];
    double t = 0.0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < Mloc; m += (int64_t)gridDim.x * blockDim.x) {
        double d = s2[m] * (raw[m + (int64_t)c * ld] - f2[m] * sb);
        if (m >= mask_lo && m < mask_hi) d = 0.0;
        D[m + (int64_t)c * ld] = d;
        t += f2[m] * d;
    }
    t = block_sum_256(t, sm);
    if (threadIdx.x == 0) part[(int64_t)c * SGB_PART_BLOCKS + blockIdx.x] = t;
}

int k_sweep1_post(sgb_ctx *h, const double *raw, int64_t ld, int k, const double *d_colsum, int64_t mask_lo,
                  int64_t mask_hi, double *D, double *d_t)
{
    double *part = reinterpret_cast<double *>(h->ws);
    int nb = k_grid_blocks(h, h->Mloc);
    sweep1_post_kernel<<<dim3(nb, k), 256, 0, h->stream>>>(raw, ld, h->Mloc, h->d_f2, h->d_s2, d_colsum, mask_lo, mask_hi, D, part);
    LAUNCH_CHECK(h);
    finish_partials_kernel<<<1, k, 0, h->stream>>>(part, nb, d_t);
    LAUNCH_CHECK(h);
    return 0;
}

__global__ void sweep2_post_kernel(const double *__restrict__ raw, int64_t ldr, int64_t N, const double *__restrict__ t,
                                   double inv_m, double *__restrict__ Y, int64_t ldy)
{
    int c = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) Y[i + (int64_t)c * ldy] = (raw[i + (int64_t)c * ldr] - t[c]) * inv_m;
}

int k_sweep2_post(sgb_ctx *h, const double *raw, int64_t ldr, int k, const double *d_t, double inv_m, double *Y, int64_t ldy)
{
    sweep2_post_kernel<<<dim3((unsigned)cdiv(h->N, 256), k), 256, 0, h->stream>>>(raw, ldr, h->N, d_t, inv_m, Y, ldy);
    LAUNCH_CHECK(h);
    return 0;
}

// diag_i = sum_m z_mi^2 = sum_m [g==1] s^2(1-4f) + [g==2] s^2(4-8f) + sum_m 4 f^2 s^2, per marker range (column)
//   D1[m,c] = s^2(1-2*f2)   (weight of one copy; the "value" plane contributes g*D1)
//   D2[m,c] = 2 s^2         (extra weight of the g==2 indicator plane: 4-8f = 2(1-4f) + 2)
__global__ void diag_prep_kernel(int64_t Mloc, int64_t ld, const double *__restrict__ f2, const double *__restrict__ s2,
                                 const int64_t *__restrict__ lo, const int64_t *__restrict__ hi, double *__restrict__ D1,
                                 double *__restrict__ D2, double *__restrict__ part)
{
    __shared__ double sm[8];
    int c = blockIdx.y;
    int64_t l = lo[c], hh = hi[c];
    double t = 0.0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < Mloc; m += (int64_t)gridDim.x * blockDim.x) {
        bool in = m >= l && m < hh;
        double ss = s2[m], ff = f2[m];
        D1[m + (int64_t)c * ld] = in ? ss * (1.0 - 2.0 * ff) : 0.0;
        D2[m + (int64_t)c * ld] = in ? 2.0 * ss : 0.0;
        if (in) t += ff * ff * ss;
    }
    t = block_sum_256(t, sm);
    if (threadIdx.x == 0) part[(int64_t)c * SGB_PART_BLOCKS + blockIdx.x] = t;
}

int k_diag_prep(sgb_ctx *h, int nchr, const int64_t *d_lo, const int64_t *d_hi, double *D1, double *D2, double *d_const)
{
    double *part = reinterpret_cast<double *>(h->ws);
    int nb = k_grid_blocks(h, h->Mloc);
    diag_prep_kernel<<<dim3(nb, nchr), 256, 0, h->stream>>>(h->Mloc, h->rowsG, h->d_f2, h->d_s2, d_lo, d_hi, D1, D2, part);
    LAUNCH_CHECK(h);
    finish_partials_kernel<<<1, nchr, 0, h->stream>>>(part, nb, d_const);
    LAUNCH_CHECK(h);
    return 0;
}

__global__ void diag_post_kernel(const double *__restrict__ r1, const double *__restrict__ r2, int64_t ld, int64_t N,
                                 const double *__restrict__ cst, double *__restrict__ out, int64_t ldo)
{
    int c = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i + (int64_t)c * ldo] = r1[i + (int64_t)c * ld] + r2[i + (int64_t)c * ld] + cst[c];
}

int k_diag_post(sgb_ctx *h, const double *raw1, const double *raw2, int64_t ld, int ncol, const double *d_const,
                double *out, int64_t ldo)
{
    diag_post_kernel<<<dim3((unsigned)cdiv(h->N, 256), ncol), 256, 0, h->stream>>>(raw1, raw2, ld, h->N, d_const, out, ldo);
    LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Sigma diag, PCG, small dense helpers
// ---------------------------------------------------------------------------------------------------
// getDiagOfSigma[_LOCO] (FG.cpp:2322-2393): tau1*diag*diag_scale + tau0/w, floored at 1e-4
__global__ void sigma_diag_kernel(const double *__restrict__ diag, double diag_scale, int diag_one,
                                  const double *__restrict__ w, double tau0, double tau1, int64_t N, double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double d = (diag_one ? tau1 : tau1 * diag[i] * diag_scale) + tau0 / w[i];
    if (d < 1e-4) d = 1e-4;
    out[i] = d;
}

int k_sigma_diag(sgb_ctx *h, const double *diag, double diag_scale, int diag_one, const double *w, double tau0,
                 double tau1, double *out)
{
    sigma_diag_kernel<<<(unsigned)cdiv(h->N, 256), 256, 0, h->stream>>>(diag, diag_scale, diag_one, w, tau0, tau1, h->N, out);
    LAUNCH_CHECK(h);
    return 0;
}

// x=0; r=b; z=r/diag; p=z; rz = r.z ; r2 = r.r      (FG.cpp:2599-2684).  `dsig` is diag(Sigma) (not inverted).
__global__ void pcg_init_kernel(const double *__restrict__ B, const double *__restrict__ dsig, int64_t N,
                                double *__restrict__ X, double *__restrict__ R, double *__restrict__ Z,
                                double *__restrict__ P, double *__restrict__ part)
{
    __shared__ double sm[8];
    int c = blockIdx.y;
    double rz = 0.0, r2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = i + (int64_t)c * N;
        double r = B[o], z = (1.0 / dsig[i]) * r;
        X[o] = 0.0; R[o] = r; Z[o] = z; P[o] = z;
        rz += r * z; r2 += r * r;
    }
    rz = block_sum_256(rz, sm);
    r2 = block_sum_256(r2, sm);
    if (threadIdx.x == 0) {
        part[((int64_t)c * 2 + 0) * SGB_PART_BLOCKS + blockIdx.x] = rz;
        part[((int64_t)c * 2 + 1) * SGB_PART_BLOCKS + blockIdx.x] = r2;
    }
}

__global__ void pcg_init_finish_kernel(const double *__restrict__ part, int nblk, double *__restrict__ rz, double *__restrict__ r2)
{
    int c = threadIdx.x;
    rz[c] = sum_partials(part + ((int64_t)c * 2 + 0) * SGB_PART_BLOCKS, nblk);
    r2[c] = sum_partials(part + ((int64_t)c * 2 + 1) * SGB_PART_BLOCKS, nblk);
}

int k_pcg_init(sgb_ctx *h, const double *B, const double *dsig, int k, double *X, double *R, double *Z, double *P,
               double *d_rz, double *d_r2)
{
    double *part = reinterpret_cast<double *>(h->ws);
    int nb = k_grid_blocks(h, h->N);
    pcg_init_kernel<<<dim3(nb, k), 256, 0, h->stream>>>(B, dsig, h->N, X, R, Z, P, part);
    LAUNCH_CHECK(h);
    pcg_init_finish_kernel<<<1, k, 0, h->stream>>>(part, nb, d_rz, d_r2);
    LAUNCH_CHECK(h);
    return 0;
}

__global__ void gather_cols_kernel(const double *__restrict__ src, const int *__restrict__ cols, int64_t N, double *__restrict__ dst)
{
    int j = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) dst[i + (int64_t)j * N] = src[i + (int64_t)cols[j] * N];
}
__global__ void scatter_cols_kernel(const double *__restrict__ src, const int *__restrict__ cols, int64_t N, double *__restrict__ dst)
{
    int j = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) dst[i + (int64_t)cols[j] * N] = src[i + (int64_t)j * N];
}
int k_gather_cols(sgb_ctx *h, const double *src, const int *d_cols, int ncols, double *dst)
{
    gather_cols_kernel<<<dim3((unsigned)cdiv(h->N, 256), ncols), 256, 0, h->stream>>>(src, d_cols, h->N, dst);
    LAUNCH_CHECK(h);
    return 0;
}
int k_scatter_cols(sgb_ctx *h, const double *src, const int *d_cols, int ncols, double *dst)
{
    scatter_cols_kernel<<<dim3((unsigned)cdiv(h->N, 256), ncols), 256, 0, h->stream>>>(src, d_cols, h->N, dst);
    LAUNCH_CHECK(h);
    return 0;
}

// Ap = tau0*p/w + tau1*Kp  (getCrossprod, FG.cpp:2397-2425), in place over the packed Kp; partial p.Ap
__global__ void pcg_step1_kernel(const double *__restrict__ P, double *__restrict__ KP, const double *__restrict__ w,
                                 double tau0, double tau1, const int *__restrict__ act, int64_t N, double *__restrict__ part)
{
    __shared__ double sm[8];
    int j = blockIdx.y, c = act[j];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double p = P[i + (int64_t)c * N];
        double ap = tau0 * (p * (1.0 / w[i]));
        if (tau1 != 0.0) ap += tau1 * KP[i + (int64_t)j * N];
        KP[i + (int64_t)j * N] = ap;
        s += p * ap;
    }
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) part[(int64_t)j * SGB_PART_BLOCKS + blockIdx.x] = s;
}

int k_pcg_step1(sgb_ctx *h, const double *P, double *KP, const double *w, double tau0, double tau1, const int *d_act,
                int nact, double *d_part)
{
    int nb = k_grid_blocks(h, h->N);
    pcg_step1_kernel<<<dim3(nb, nact), 256, 0, h->stream>>>(P, KP, w, tau0, tau1, d_act, h->N, d_part);
    LAUNCH_CHECK(h);
    return 0;
}

// a = rz/pAp; x += a p; r -= a Ap; z = r/diag; partials of z.r and r.r     (FG.cpp:2710-2783)
__global__ void pcg_step2_kernel(const double *__restrict__ P, const double *__restrict__ AP, const double *__restrict__ dsig,
                                 const int *__restrict__ act, int64_t N, int nblk, double *__restrict__ X,
                                 double *__restrict__ R, double *__restrict__ Z, const double *__restrict__ rz,
                                 const double *__restrict__ part_in, double *__restrict__ part_out)
{
    __shared__ double sm[8];
    int j = blockIdx.y, c = act[j];
    double pap = sum_partials(part_in + (int64_t)j * SGB_PART_BLOCKS, nblk);
    double a = rz[c] / pap;
    double zr = 0.0, rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = i + (int64_t)c * N;
        double p = P[o];
        X[o] = X[o] + a * p;
        double r = R[o] - a * AP[i + (int64_t)j * N];
        double z = (1.0 / dsig[i]) * r;
        R[o] = r; Z[o] = z;
        zr += z * r; rr += r * r;
    }
    zr = block_sum_256(zr, sm);
    rr = block_sum_256(rr, sm);
    if (threadIdx.x == 0) {
        part_out[((int64_t)j * 2 + 0) * SGB_PART_BLOCKS + blockIdx.x] = zr;
        part_out[((int64_t)j * 2 + 1) * SGB_PART_BLOCKS + blockIdx.x] = rr;
    }
}

int k_pcg_step2(sgb_ctx *h, const double *P, const double *AP, const double *dsig, const int *d_act, int nact,
                double *X, double *R, double *Z, double *d_rz, const double *d_part_in, double *d_part_out)
{
    int nb = k_grid_blocks(h, h->N);
    pcg_step2_kernel<<<dim3(nb, nact), 256, 0, h->stream>>>(P, AP, dsig, d_act, h->N, nb, X, R, Z, d_rz, d_part_in, d_part_out);
    LAUNCH_CHECK(h);
    return 0;
}

// bet = z1.r1 / z.r ; p = z1 + bet p ; publish new rz and r2 (rz_out != rz_in: other blocks still read rz_in)
__global__ void pcg_step3_kernel(double *__restrict__ P, const double *__restrict__ Z, const int *__restrict__ act, int64_t N,
                                 int nblk, const double *__restrict__ rz_in, double *__restrict__ rz_out,
                                 double *__restrict__ r2, const double *__restrict__ part)
{
    int j = blockIdx.y, c = act[j];
    double zr = sum_partials(part + ((int64_t)j * 2 + 0) * SGB_PART_BLOCKS, nblk);
    double bet = zr / rz_in[c];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = i + (int64_t)c * N;
        P[o] = Z[o] + bet * P[o];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rz_out[c] = zr;
        r2[c] = sum_partials(part + ((int64_t)j * 2 + 1) * SGB_PART_BLOCKS, nblk);
    }
}

int k_pcg_step3(sgb_ctx *h, double *P, const double *Z, const int *d_act, int nact, const double *rz_in, double *rz_out,
                   double *d_r2, const double *d_part)
{
    int nb = k_grid_blocks(h, h->N);
    pcg_step3_kernel<<<dim3(nb, nact), 256, 0, h->stream>>>(P, Z, d_act, h->N, nb, rz_in, rz_out, d_r2, d_part);
    LAUNCH_CHECK(h);
    return 0;
}

// out[q] = A[:,pairs[2q]] . B[:,pairs[2q+1]]
__global__ void pair_dots_kernel(const double *__restrict__ A, int64_t lda, const double *__restrict__ B, int64_t ldb,
                                 const int *__restrict__ pairs, int64_t N, double *__restrict__ part)
{
    __shared__ double sm[8];
    int q = blockIdx.y;
    const double *a = A + (int64_t)pairs[2 * q] * lda, *b = B + (int64_t)pairs[2 * q + 1] * ldb;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) s += a[i] * b[i];
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) part[(int64_t)q * SGB_PART_BLOCKS + blockIdx.x] = s;
}
__global__ void finish_many_kernel(const double *__restrict__ part, int nblk, int n, double *__restrict__ out)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = sum_partials(part + (int64_t)q * SGB_PART_BLOCKS, nblk);
}

int k_pair_dots(sgb_ctx *h, const double *A, int64_t lda, const double *B, int64_t ldb, const int *d_pairs, int npairs,
                double *d_out)
{
    if (npairs == 0) return 0;
    if ((size_t)npairs * SGB_PART_BLOCKS * sizeof(double) > h->ws_bytes) return sgb_fail(h, "k_pair_dots: workspace too small");
    double *part = reinterpret_cast<double *>(h->ws);
    int nb = k_grid_blocks(h, h->N);
    pair_dots_kernel<<<dim3(nb, npairs), 256, 0, h->stream>>>(A, lda, B, ldb, d_pairs, h->N, part);
    LAUNCH_CHECK(h);
    finish_many_kernel<<<(unsigned)cdiv(npairs, 128), 128, 0, h->stream>>>(part, nb, npairs, d_out);
    LAUNCH_CHECK(h);
    return 0;
}

// Out[:,j] = In[:,j] - SiX * C[:,j]      (P u = Sigma^-1 u - Sigma^-1 X (cov (Sigma^-1 X)^T u), FG.cpp:3139)
__global__ void project_kernel(const double *__restrict__ In, const double *__restrict__ SiX, int p, const double *__restrict__ Cm,
                               int64_t N, double *__restrict__ Out)
{
    int j = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double v = In[i + (int64_t)j * N];
    for (int q = 0; q < p; q++) v -= SiX[i + (int64_t)q * N] * Cm[q + (int64_t)j * p];
    Out[i + (int64_t)j * N] = v;
}

int k_project(sgb_ctx *h, const double *In, const double *SiX, int p, const double *d_C, int ncol, double *Out)
{
    if (ncol == 0) return 0;
    project_kernel<<<dim3((unsigned)cdiv(h->N, 256), ncol), 256, 0, h->stream>>>(In, SiX, p, d_C, h->N, Out);
    LAUNCH_CHECK(h);
    return 0;
}

// eta = Y - tau0 (Sigma_iY - Sigma_iX alpha) / w        (FG.cpp:3194)
__global__ void eta_kernel(const double *__restrict__ Y, const double *__restrict__ SiY, const double *__restrict__ SiX, int p,
                           const double *__restrict__ alpha, const double *__restrict__ w, double tau0, int64_t N,
                           double *__restrict__ eta)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double v = SiY[i];
    for (int q = 0; q < p; q++) v -= SiX[i + (int64_t)q * N] * alpha[q];
    eta[i] = Y[i] - tau0 * v / w[i];
}

int k_eta(sgb_ctx *h, const double *Y, const double *SiY, const double *SiX, int p, const double *d_alpha,
          const double *w, double tau0, double *eta)
{
    eta_kernel<<<(unsigned)cdiv(h->N, 256), 256, 0, h->stream>>>(Y, SiY, SiX, p, d_alpha, w, tau0, h->N, eta);
    LAUNCH_CHECK(h);
    return 0;
}

__global__ void rademacher_kernel(double *__restrict__ B, int64_t n, uint64_t seed)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) B[i] = (mix32(seed, 0x5151, (uint64_t)i, 7) & 1u) ? 1.0 : -1.0;
}

int k_rademacher_fill(sgb_ctx *h, double *B, int64_t n, uint64_t seed)
{
    rademacher_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(B, n, seed);
    LAUNCH_CHECK(h);
    return 0;
}

__global__ void axpby_kernel(double a, const double *__restrict__ x, double b, const double *__restrict__ y, int64_t n,
                             double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + b * y[i];
}

int k_axpby(sgb_ctx *h, double a, const double *x, double b, const double *y, int64_t n, double *out)
{
    if (n == 0) return 0;
    axpby_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(a, x, b, y, n, out);
    LAUNCH_CHECK(h);
    return 0;
}

// d_count[0] += number of positions where a and b differ bitwise (probe-product cache check)
__global__ void count_diff_kernel(const unsigned long long *__restrict__ a, const unsigned long long *__restrict__ b, int64_t n, int *__restrict__ d_count)
{
    int local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) local += a[i] != b[i];
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(d_count, local);
}

int k_count_diff(sgb_ctx *h, const double *a, const double *b, int64_t n, int *d_count)
{
    CUDA_OK(h, cudaMemsetAsync(d_count, 0, sizeof(int), h->stream));
    int64_t gb = cdiv(n, 256 * 8);
    if (gb > (int64_t)h->sm_count * 16) gb = (int64_t)h->sm_count * 16;
    if (gb < 1) gb = 1;
    count_diff_kernel<<<(unsigned)gb, 256, 0, h->stream>>>(reinterpret_cast<const unsigned long long *>(a), reinterpret_cast<const unsigned long long *>(b), n, d_count);
    LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// resident driver loops (sgb_get_coef, sgb_variance_ratio_markers): the O(N) algebra the R driver does between exports
// ---------------------------------------------------------------------------------------------------
// The IRLS update of Get_Coef (FG.R:4-8, 21-26): eta = eta_in (+ offset), mu = linkinv(eta), Y = eta - offset + (y - mu) / mu.eta,
// W = (mu.eta / sqrt(variance(mu)))^2.  Binomial: R's logit_linkinv / logit_mu_eta (stats/src/family.c) with their |eta| > 30 clamps.
__global__ void irls_update_kernel(int family, const double *eta_in /* may alias eta_out */, int add_offset, const double *__restrict__ y,
                                   const double *__restrict__ offset, int64_t N, double *eta_out, double *__restrict__ mu_out,
                                   double *__restrict__ Y, double *__restrict__ W)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double off = offset[i];
    const double eta = add_offset ? eta_in[i] + off : eta_in[i];
    double mu, me, var;
    if (family == 0) {
        const double eps = 2.220446049250313e-16;
        const double e = exp(eta);
        const double tmp = eta < -30.0 ? eps : (eta > 30.0 ? 1.0 / eps : e);
        mu = tmp / (1.0 + tmp);
        const double opexp = 1.0 + e;
        me = (eta > 30.0 || eta < -30.0) ? eps : e / (opexp * opexp);
        var = mu * (1.0 - mu);
    } else {
        mu = eta; me = 1.0; var = 1.0;
    }
    const double sqrtW = me / sqrt(var);
    eta_out[i] = eta;
    mu_out[i] = mu;
    Y[i] = eta - off + (y[i] - mu) / me;
    W[i] = sqrtW * sqrtW;
}

int k_irls_update(sgb_ctx *h, int family, const double *eta_in, int add_offset, const double *y, const double *offset, double *eta_out,
                  double *mu, double *Y, double *W)
{
    irls_update_kernel<<<(unsigned)cdiv(h->N, 256), 256, 0, h->stream>>>(family, eta_in, add_offset, y, offset, h->N, eta_out, mu, Y, W);
    LAUNCH_CHECK(h);
    return 0;
}

// Out[i + j*N] = genotype of sample i in packed row rows[j] of P (device coding; tiled store or row-major rows of `stride`
// bytes), 0 for rows[j] < 0 (a marker another rank owns: the sum-allreduce that follows fills it in)
__global__ void decode_marker_cols_kernel(const uint8_t *__restrict__ P, int tiled, int64_t stride, const int64_t *__restrict__ rows,
                                          int64_t N, double *__restrict__ Out)
{
    const int j = blockIdx.y;
    const int64_t r = rows[j];
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // packed byte = samples 4b .. 4b+3
    if (4 * b >= N) return;
    int g[4] = {0, 0, 0, 0};
    if (r >= 0) {
        const uint32_t v = P[tiled ? sgb_tiled_off(r, b, stride) : r * stride + b];
        sgb_unpack_nibble(v & 15u, g[0], g[1]);
        sgb_unpack_nibble(v >> 4, g[2], g[3]);
    }
    double *o = Out + (int64_t)j * N + 4 * b;
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (4 * b + q < N) o[q] = (double)g[q];
}

int k_decode_marker_cols(sgb_ctx *h, const uint8_t *P, int tiled, int64_t stride, const int64_t *d_rows, int ncol, double *Out)
{
    if (ncol == 0) return 0;
    decode_marker_cols_kernel<<<dim3((unsigned)cdiv(cdiv(h->N, 4), 256), ncol), 256, 0, h->stream>>>(P, tiled, stride, d_rows, h->N, Out);
    LAUNCH_CHECK(h);
    return 0;
}

// G[:,j] = 2 - G[:,j] where flip[j] (FG.R:2318-2320)
__global__ void flip_cols_kernel(double *__restrict__ G, const int *__restrict__ flip, int64_t N)
{
    const int j = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && flip[j]) G[i + (int64_t)j * N] = 2.0 - G[i + (int64_t)j * N];
}

int k_flip_cols(sgb_ctx *h, double *G, const int *d_flip, int ncol)
{
    if (ncol == 0) return 0;
    flip_cols_kernel<<<dim3((unsigned)cdiv(h->N, 256), ncol), 256, 0, h->stream>>>(G, d_flip, h->N);
    LAUNCH_CHECK(h);
    return 0;
}

// Out[:,j] = v .* In[:,j]
__global__ void rowscale_cols_kernel(const double *__restrict__ v, const double *__restrict__ In, int64_t N, double *__restrict__ Out)
{
    const int j = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) Out[i + (int64_t)j * N] = v[i] * In[i + (int64_t)j * N];
}

int k_rowscale_cols(sgb_ctx *h, const double *v, const double *In, int ncol, double *Out)
{
    if (ncol == 0) return 0;
    rowscale_cols_kernel<<<dim3((unsigned)cdiv(h->N, 256), ncol), 256, 0, h->stream>>>(v, In, h->N, Out);
    LAUNCH_CHECK(h);
    return 0;
}
