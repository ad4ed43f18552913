// Dense N x N GRM on device (BASELINE config 4): build with the tcgen05 int8 kernel, store the
// lower block-trapezoid in fp64, and serve K.B products from the stored matrix (PCG "on stored GRM").
//
//   K_ij = (1/M) sum_m s_m^2 (g_im - 2 f_m)(g_jm - 2 f_m)          (same definition as the on-the-fly product,
//                                                                    FG.cpp:1445-1502 / 665-704 for the diagonal)
// With h = 2 - g (what the packed decode feeds the MMA) and phi_m = 2 - 2 f_m:
//   M K_ij = sum_m w_m h_im h_jm  -  U_i  -  U_j  +  C,    U_i = sum_m w_m phi_m h_im,   C = sum_m w_m phi_m^2
// The Gram term is the dense contraction.  It runs EXACTLY on the int8 tensor cores: the weights are fixed-point
// integers W_m = round(s_m^2 2^S) split into `limbs` balanced base-128 digits, the B operand of limb l is the int8
// image  digit_l(m) * h_jm  (|.| <= 128) of a 128-sample panel, the A operand is the 2-bit sample-major store decoded
// into TMEM by pk2_umma_kernel, accumulation is int32.  The centring terms are integers too: N phi_m = 2N - AC_m, so
//   N^2 2^S M K_ij = N^2 Q_ij - N (U'_i + U'_j) + C',   U'_i = sum_m W_m (2N - AC_m) h_im,   C' = sum_m W_m (2N - AC_m)^2
// is evaluated in 128-bit integer arithmetic (U' from one exact 5-column tensor sweep over 12-bit pieces of W) and rounded
// to fp64 ONCE, after the cancellation.  Only the weight rounding (2^-(7 limbs - 2) relative to the largest weight)
// separates the stored matrix from exact arithmetic.
//
// Storage: block-row R (samples 128R .. 128R+127) keeps columns 0 .. 128(R+1)-1 as panel[i][n] = K[128R+n][i]
// (n fastest, 1 KB per i).  Block-rows are dealt to ranks cyclically (work grows with R); every rank needs all
// markers for its block-rows, so the sample-major shards are exchanged once with ncclBroadcast.  The product
// y = K b reads every stored element once (both triangles are served from it) and ends in the same sum-allreduce as
// the on-the-fly product.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "sgb_internal.h"

#define DG_BLOCK 128            // samples per block-row = N of the MMA
#define DG_CHUNK 512            // rows of a panel handled by one CTA of the product kernel
#define DG_UPIECES 5            // 12-bit pieces of the integer weights in the exact centring sums (covers W < 2^60)

struct dg_item { int64_t off; int32_t R; int32_t i0; int32_t rows; int32_t partB; };

struct sgb_dense {
    int limbs = 0, S = 0;
    int64_t nbr = 0;                       // block-rows in total
    std::vector<int64_t> off;              // element offset of block-row R in `pool`, -1 if another rank owns it
    double *pool = nullptr; size_t pool_elems = 0;
    double *dU = nullptr;                  // U'_i as DG_UPIECES exact integer pieces [piece][N]
    dg_item *d_items = nullptr; int64_t n_items = 0, n_items_b = 0;   // n_items_b mirrored chunks first, then the diagonal blocks
    double build_ms = 0.0;
    bool partial = false;                  // bench sample: not all block-rows were built
    double tensor_ops = 0.0;               // int8 multiply-adds x 2 issued by the build
};

#define DG_LAUNCH_CHECK(h)                                                                              \
    do {                                                                                                \
        (h)->cnt.n_kernel_launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess) return sgb_fail(h, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, \
                                                cudaGetErrorString(e__));                               \
    } while (0)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up64(int64_t a, int64_t b) { return cdiv64(a, b) * b; }

void sgb_dense_free(sgb_ctx *h)
{
    if (!h->dense) return;
    cudaSetDevice(h->device);
    if (h->dense->pool) cudaFree(h->dense->pool);
    if (h->dense->dU) cudaFree(h->dense->dU);
    if (h->dense->d_items) cudaFree(h->dense->d_items);
    delete h->dense;
    h->dense = nullptr;
    h->grm_mode = SGB_GRM_PACKED;
}

__device__ __forceinline__ uint32_t dg_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// bytes of x (each 0..2) times the signed digits in dg (4 x int8) -> 4 x int8
__device__ __forceinline__ uint32_t mul_bytes(uint32_t x, uint32_t dg)
{
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int hv = (int)((x >> (8 * b)) & 255u);
        int dv = (int)(int8_t)((dg >> (8 * b)) & 255u);
        r |= (uint32_t)((hv * dv) & 255) << (8 * b);
    }
    return r;
}

// B operand of one (block-row, limb, shard): image of the UMMA smem stage (see split_limbs_umma_kernel) with
//   byte[blk][c][w][l][4q + s] = digit(p) * h(sample row0 + 8c + l, marker position p),  p = 128 blk + 16 w + {0,8,1,9}[q] + 2 s
// One thread per (blk, c, w, l): one 32-bit word of the packed row = 16 genotypes = the 16 bytes it writes.
__global__ void syrk_image_kernel(const uint8_t *__restrict__ Gt, int64_t sT, int64_t row0, const int8_t *__restrict__ dig,
                                  int64_t nblk, int8_t *__restrict__ img)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int l = (int)(t & 7), w = (int)((t >> 3) & 7), c = (int)((t >> 6) & 15);
    const int64_t blk = t >> 10;
    if (blk >= nblk) return;
    const uint32_t wv = *reinterpret_cast<const uint32_t *>(Gt + sgb_tiled_off(row0 + 8 * c + l, blk * 32 + w * 4, sT));
    const uint32_t hi = wv >> 16;
    // value plane of the pair-ternary coding: byte = 2 - genotype (kernels.cu decode16)
    const uint32_t h0 = dg_prmt(0x02000102u, 0x01020001u, wv);     // positions 0,2,4,6
    const uint32_t h1 = dg_prmt(0x02000102u, 0x01020001u, hi);     // 8,10,12,14
    const uint32_t h2 = dg_prmt(0x01020202u, 0x00000101u, wv);     // 1,3,5,7
    const uint32_t h3 = dg_prmt(0x01020202u, 0x00000101u, hi);     // 9,11,13,15
    const uint4 dv = *reinterpret_cast<const uint4 *>(dig + blk * 128 + w * 16);      // digits of positions 0..15
    // gather the digit bytes in the same order as the decode
    const uint32_t e0 = dg_prmt(dv.x, dv.y, 0x6420u), o0 = dg_prmt(dv.x, dv.y, 0x7531u);
    const uint32_t e1 = dg_prmt(dv.z, dv.w, 0x6420u), o1 = dg_prmt(dv.z, dv.w, 0x7531u);
    uint4 o;
    o.x = mul_bytes(h0, e0);
    o.y = mul_bytes(h1, e1);
    o.z = mul_bytes(h2, o0);
    o.w = mul_bytes(h3, o1);
    *reinterpret_cast<uint4 *>(img + (blk * 16 + c) * 1024 + w * 128 + l * 16) = o;
}

// Exact limb recombination: Q = qlo + 2^28 qhi with qlo = sum_{l<4} 128^l acc_l and qhi = sum_{l>=4} 128^(l-4) acc_l,
// both integers below 2^53 held in doubles.  acc is zeroed for the next limb.
__global__ void syrk_fold_kernel(int32_t *__restrict__ acc, int64_t n, int l, double *__restrict__ qlo, double *__restrict__ qhi)
{
    const double scale = (double)(1ll << (7 * (l & 3)));
    double *q = l < 4 ? qlo : qhi;
    const bool first = (l & 3) == 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = scale * (double)acc[i];
        acc[i] = 0;
        q[i] = first ? v : q[i] + v;
    }
}

__device__ __forceinline__ double i128_to_double(__int128 t)
{
    const bool neg = t < 0;
    unsigned __int128 a = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
    double v = (double)(unsigned long long)(a >> 64) * 18446744073709551616.0 + (double)(unsigned long long)a;
    return neg ? -v : v;
}

// panel[i][n] = (N^2 Q - N (U'_i + U'_j) + C') * mul, j = row0 + n, in 128-bit integers; entries outside N x N are zero
__global__ void syrk_finalize_kernel(double *__restrict__ panel, const double *__restrict__ qhi, int has_hi, int64_t rows, int64_t row0,
                                     int64_t N, const double *__restrict__ U, unsigned long long c_hi, unsigned long long c_lo, double mul)
{
    const int64_t tot = rows * DG_BLOCK;
    const __int128 C = (__int128)(((unsigned __int128)c_hi << 64) | c_lo);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e >> 7, j = row0 + (e & 127);
        if (i >= N || j >= N) { panel[e] = 0.0; continue; }
        __int128 Q = (__int128)(long long)panel[e];
        if (has_hi) Q += (__int128)(long long)qhi[e] << 28;
        __int128 Us = 0;
#pragma unroll
        for (int p = 0; p < DG_UPIECES; p++)
            Us += (__int128)(__double2ll_rn(U[(int64_t)p * N + i]) + __double2ll_rn(U[(int64_t)p * N + j])) << (12 * p);
        const __int128 T = (__int128)N * ((__int128)N * Q - Us) + C;
        panel[e] = i128_to_double(T) * mul;
    }
}

// U'_p,i = 2 sum_m v_p,m - (G^T v_p)_i   (raw = G^T v; exact: all sums are integers < 2^53)
__global__ void syrk_u_kernel(const double *__restrict__ raw, int64_t ldr, int64_t N, const double *__restrict__ sums, double *__restrict__ U)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (i < N) U[(int64_t)p * N + i] = 2.0 * sums[p] - raw[(int64_t)p * ldr + i];
}

// ---------------------------------------------------------------------------------------------------
// y += K b from the stored trapezoid.  One CTA per (block-row, 512-row chunk of its panel):
//   part A   y[128R + n] += sum_i panel[i][n] b[i]                    (rows of the block-row)
//   part B   y[i]        += sum_n panel[i][n] b[128R + n]   (i < 128R) (the mirrored upper triangle)
// A warp streams 4 panel rows at a time (lane = 4 consecutive n, one 32-byte load per row, the next group prefetched
// into a second register buffer), keeps the part-A sums of KC columns in registers and reduces the part-B partials with
// a shuffle butterfly (6 shuffles per 4 rows and column).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 ldg_f64x4(const double *p)
{
    double4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

// one group = 4 panel rows: part-A accumulation and the part-B sums of KC columns.  sbi holds b over the chunk's rows
// (staged once per CTA: the per-row values are warp-uniform, and 4 x KC uniform global loads per group kept the LSU
// pipe at 60 %).
template <int KC, bool PARTB>
__device__ __forceinline__ void symv_group(const double4 (&kv)[4], const double (*sbi)[DG_CHUNK], const double (*sbn)[DG_BLOCK], int lane,
                                           double (&accA)[KC][4], double (*sB)[DG_CHUNK], int lrow)
{
#pragma unroll
    for (int c = 0; c < KC; c++) {
        const double2 b01 = *reinterpret_cast<const double2 *>(&sbi[c][lrow]), b23 = *reinterpret_cast<const double2 *>(&sbi[c][lrow + 2]);
        const double b0 = b01.x, b1 = b01.y, b2 = b23.x, b3 = b23.y;
        accA[c][0] += kv[0].x * b0; accA[c][1] += kv[0].y * b0; accA[c][2] += kv[0].z * b0; accA[c][3] += kv[0].w * b0;
        accA[c][0] += kv[1].x * b1; accA[c][1] += kv[1].y * b1; accA[c][2] += kv[1].z * b1; accA[c][3] += kv[1].w * b1;
        accA[c][0] += kv[2].x * b2; accA[c][1] += kv[2].y * b2; accA[c][2] += kv[2].z * b2; accA[c][3] += kv[2].w * b2;
        accA[c][0] += kv[3].x * b3; accA[c][1] += kv[3].y * b3; accA[c][2] += kv[3].z * b3; accA[c][3] += kv[3].w * b3;
        if (PARTB) {
            const double4 bn = *reinterpret_cast<const double4 *>(&sbn[c][4 * lane]);
            double pr[4];
#pragma unroll
            for (int r = 0; r < 4; r++) pr[r] = kv[r].x * bn.x + kv[r].y * bn.y + kv[r].z * bn.z + kv[r].w * bn.w;
            // butterfly: 4 row partials over 32 lanes -> the lanes with (lane & 7) == 0 end with the full sum of row (lane >> 3)
            const bool up16 = lane & 16, up8 = lane & 8;
            const double s0 = up16 ? pr[0] : pr[2], k0 = up16 ? pr[2] : pr[0];
            const double s1 = up16 ? pr[1] : pr[3], k1 = up16 ? pr[3] : pr[1];
            const double q0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
            const double q1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
            const double s2 = up8 ? q0 : q1, k2 = up8 ? q1 : q0;
            double q = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
            q += __shfl_xor_sync(0xffffffffu, q, 4);
            q += __shfl_xor_sync(0xffffffffu, q, 2);
            q += __shfl_xor_sync(0xffffffffu, q, 1);
            if ((lane & 7) == 0) sB[c][lrow + (lane >> 3)] = q;          // every chunk row belongs to exactly one warp and group
        }
    }
}

template <int KC, bool PARTB>
__global__ void __launch_bounds__(256, (KC <= 2 ? 2 : 1))
dense_symv_kernel(const double *__restrict__ pool, const dg_item *__restrict__ items, int64_t item0, const double *__restrict__ B, int64_t N,
                  double *__restrict__ Y)
{
    // dynamic shared memory: sbn [KC][128] (b over the block-row's own samples; later the part-A sums) | sbi [KC][512] (b over
    // the chunk's rows) | sB [KC][512] (part-B sums of the chunk rows, flushed with coalesced atomics; PARTB only)
    extern __shared__ __align__(32) double symv_smem[];
    double (*sbn)[DG_BLOCK] = reinterpret_cast<double (*)[DG_BLOCK]>(symv_smem);
    double (*sbi)[DG_CHUNK] = reinterpret_cast<double (*)[DG_CHUNK]>(symv_smem + KC * DG_BLOCK);
    double (*sB)[DG_CHUNK] = reinterpret_cast<double (*)[DG_CHUNK]>(symv_smem + KC * DG_BLOCK + KC * DG_CHUNK);
    const dg_item it = items[item0 + blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *panel = pool + it.off;
    for (int e = threadIdx.x; e < KC * DG_BLOCK; e += 256) {
        const int c = e / DG_BLOCK, n = e % DG_BLOCK;
        const int64_t j = (int64_t)it.R * DG_BLOCK + n;
        sbn[c][n] = j < N ? B[(int64_t)c * N + j] : 0.0;
    }
    for (int e = threadIdx.x; e < KC * DG_CHUNK; e += 256) {
        const int c = e / DG_CHUNK, r = e % DG_CHUNK;
        const int64_t i = (int64_t)it.i0 + r;
        sbi[c][r] = (r < it.rows && i < N) ? B[(int64_t)c * N + i] : 0.0;
    }
    __syncthreads();
    double accA[KC][4];
#pragma unroll
    for (int c = 0; c < KC; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) accA[c][j] = 0.0;
    // it.rows is a multiple of 128 and a warp owns 64 consecutive chunk rows: all or nothing
    const int g0 = warp * (DG_CHUNK / 8);
    if (g0 < it.rows) {
        const double *src = panel + ((int64_t)it.i0 + g0) * DG_BLOCK + 4 * lane;
        double4 bufA[4], bufB[4];
#pragma unroll
        for (int r = 0; r < 4; r++) bufA[r] = ldg_f64x4(src + (int64_t)r * DG_BLOCK);
        // groups of 4 rows, double buffered in registers: the next group's loads are in flight while this one is consumed
#pragma unroll 1
        for (int g = 0; g < DG_CHUNK / 8; g += 8) {
#pragma unroll
            for (int r = 0; r < 4; r++) bufB[r] = ldg_f64x4(src + (int64_t)(g + 4 + r) * DG_BLOCK);
            symv_group<KC, PARTB>(bufA, sbi, sbn, lane, accA, sB, g0 + g);
            if (g + 8 < DG_CHUNK / 8) {
#pragma unroll
                for (int r = 0; r < 4; r++) bufA[r] = ldg_f64x4(src + (int64_t)(g + 8 + r) * DG_BLOCK);
            }
            symv_group<KC, PARTB>(bufB, sbi, sbn, lane, accA, sB, g0 + g + 4);
        }
    }
    // part-A sums of the 8 warps meet in shared memory (sbn is free now)
    double (*red)[DG_BLOCK] = sbn;
    __syncthreads();
    for (int e = threadIdx.x; e < KC * DG_BLOCK; e += 256) red[e / DG_BLOCK][e % DG_BLOCK] = 0.0;
    __syncthreads();
    if (g0 < it.rows) {
#pragma unroll
        for (int c = 0; c < KC; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) atomicAdd(&red[c][4 * lane + j], accA[c][j]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KC * DG_BLOCK; e += 256) {
        const int c = e / DG_BLOCK, n = e % DG_BLOCK;
        const int64_t j = (int64_t)it.R * DG_BLOCK + n;
        const double v = red[c][n];
        if (j < N && v != 0.0) atomicAdd(Y + (int64_t)c * N + j, v);
    }
    if (PARTB)
        for (int e = threadIdx.x; e < KC * DG_CHUNK; e += 256) {
            const int c = e / DG_CHUNK, r = e % DG_CHUNK;
            const int64_t i = (int64_t)it.i0 + r;
            if (r < it.rows && i < N) atomicAdd(Y + (int64_t)c * N + i, sB[c][r]);
        }
}

template <int KC>
static void symv_launch(sgb_ctx *h, const sgb_dense *d, const double *B, int64_t N, double *Y)
{
    // items are ordered [mirrored chunks (part A + B) ..., diagonal blocks (part A only)]
    const size_t smem_b = sizeof(double) * (size_t)KC * (DG_BLOCK + 2 * DG_CHUNK), smem_a = sizeof(double) * (size_t)KC * (DG_BLOCK + DG_CHUNK);
    if (sgb_first_on_device(h->device, SGB_SITE_SYMV_BASE + KC)) {
        cudaFuncSetAttribute(dense_symv_kernel<KC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);       // a failure here
        cudaFuncSetAttribute(dense_symv_kernel<KC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);      // surfaces at the caller's launch check
    }
    if (d->n_items_b)
        dense_symv_kernel<KC, true><<<(unsigned)d->n_items_b, 256, smem_b, h->stream>>>(d->pool, d->d_items, 0, B, N, Y);
    if (d->n_items > d->n_items_b)
        dense_symv_kernel<KC, false><<<(unsigned)(d->n_items - d->n_items_b), 256, smem_a, h->stream>>>(d->pool, d->d_items, d->n_items_b, B, N, Y);
    h->cnt.n_kernel_launches += 2;
}

// ---------------------------------------------------------------------------------------------------
// Stored-GRM product for 3..8 columns on the fp64 TENSOR pipe (mma.sync m8n8k4.f64): both triangles of a panel chunk
// are two small GEMMs against the (zero-padded to 8) right-hand sides,
//   part B   y[i][c]        += sum_n panel[i][n] b[128R + n][c]      A = panel rows (8 x 4 per MMA), B = b
//   part A   y[128R + n][c] += sum_i b[i][c] panel[i][n]             A = b^T (8 columns x 4 rows), B = panel rows (4 x 8)
// so the cross-lane reductions that bound the DFMA kernel at >= 4 columns (shuffle butterfly, 220-255 registers, one CTA of
// 8 latency-bound warps per SM: 3.1 / 1.8 TB/s at 4 / 8 columns) are done by the MMA.
// PERSISTENT WARPS: one CTA of 8 warps per SM; every warp owns whole items (<= 512 panel rows of one block-row), item
// wid, wid + W, ... of the list, and streams their groups of 8 rows (8 KB) through a private 3-stage shared-memory ring filled by
// cp.async.bulk (one 1 KB copy per row into a row stride of 130 doubles: fragment loads are bank-conflict free for part B and
// 1.5-way for part A).  The ring runs ACROSS items -- the copies of the next item are in flight while the current one is consumed
// -- and a warp never meets another one: part-B's operand b[128R + .] sits in 32 registers per lane for the whole item, part-A's
// sums leave through 32 atomics per lane and item.  (The first version gave a 512-row item to one CTA, 8 groups per warp, and
// met in shared memory at the end: 12 us of copy pipeline per CTA that the 10 us of DMMA work did not overlap -- 3.4 TB/s where the
// copies alone ran at 6.4 and the MMAs alone at 7.4.)  64 DMMAs and 64 LDS.64 per lane and group.
// ---------------------------------------------------------------------------------------------------
#define DM_STAGES 3
#define DM_RS 130               // doubles per staged panel row

__device__ __forceinline__ uint32_t dm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// WARP PAIRS: 16 warps per SM.  Pair w = (warp 2w, warp 2w + 1) shares one ring; the even warp does part A of every tile and
// owns the copies, the odd warp does part B.  (One warp doing both halves needed 222 registers, so 8 warps per SM: tensor pipe 53 %,
// DRAM 60 %, warps active 12 % under ncu -- latency-bound.  Split, each role fits 128 registers.)  full[pair][stage]: copy landed
// (expect_tx); empty[pair][stage]: both consumers are done with the tile (count 2) -- the producer lane waits for it before refilling.
__device__ __forceinline__ void dm_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(512, 1)
dense_symm_mma_kernel(const double *__restrict__ pool, const dg_item *__restrict__ items, int64_t n_items, const double *__restrict__ B, int kc,
                      int64_t N, double *__restrict__ Y)
{
    extern __shared__ __align__(128) double dm_smem[];
    __shared__ uint64_t full[8][DM_STAGES], empty[8][DM_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
    const int pair = warp >> 1, role = warp & 1;
    double *ring = dm_smem + (size_t)pair * (DM_STAGES * 8 * DM_RS);
    // neighbouring items (adjacent chunks of one panel) go to different SMs at the same time
    const int64_t W = (int64_t)gridDim.x * 8, wid = (int64_t)pair * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; w++)
            for (int s = 0; s < DM_STAGES; s++) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dm_smem_u32(&full[w][s])), "r"(1));
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dm_smem_u32(&empty[w][s])), "r"(2));
            }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const bool colok = r < kc;
    uint32_t tile_no = 0;                                                    // groups consumed so far: stage and phase of the ring
    if (role == 0) {
        // ---- producer cursor (lane 0): the next group of 8 rows to copy, running over this pair's items ----
        int64_t ti = wid;
        int gi = 0, ngi = 0;
        uint32_t issued = 0;
        const double *srci = nullptr;
        auto open_item = [&]() {
            if (ti < n_items) {
                const dg_item t = items[ti];
                ngi = t.rows >> 3; gi = 0;
                srci = pool + t.off + (int64_t)t.i0 * DG_BLOCK;
            }
        };
        auto issue = [&]() {
            while (ti < n_items && gi >= ngi) { ti += W; open_item(); }
            if (ti >= n_items) return;
            const int st = (int)(issued % DM_STAGES);
            if (issued >= DM_STAGES) dm_wait(dm_smem_u32(&empty[pair][st]), ((issued / DM_STAGES) - 1) & 1u);   // both consumers left the stage
            const uint32_t bar = dm_smem_u32(&full[pair][st]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8 * DG_BLOCK * 8) : "memory");
            const double *src = srci + (int64_t)gi * 8 * DG_BLOCK;
            double *dst = ring + (size_t)st * (8 * DM_RS);
#pragma unroll
            for (int row = 0; row < 8; row++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dm_smem_u32(dst + row * DM_RS)),
                             "l"(src + row * DG_BLOCK), "r"(DG_BLOCK * 8), "r"(bar)
                             : "memory");
            gi++; issued++;
        };
        if (lane == 0) {
            open_item();
            for (int s = 0; s < DM_STAGES; s++) issue();
        }
        // ---- part A: D[c][n] += sum over the 8 rows (two k = 4 halves) of b[i][c] * panel[i][n], 16 n-tiles ----
        for (int64_t tc = wid; tc < n_items; tc += W) {
            const dg_item it = items[tc];
            const int ng = it.rows >> 3;
            double accA[16][2];
#pragma unroll
            for (int t = 0; t < 16; t++) { accA[t][0] = 0.0; accA[t][1] = 0.0; }
            for (int g = 0; g < ng; g++, tile_no++) {
                const int st = (int)(tile_no % DM_STAGES);
                const int64_t irow0 = (int64_t)it.i0 + (int64_t)g * 8;               // panel row of tile row 0
                double bA[2];
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {                                     // A fragment: row = column c = r, k = q
                    const int64_t i = irow0 + 4 * hh + q;
                    bA[hh] = (colok && i < N) ? B[(int64_t)r * N + i] : 0.0;
                }
                dm_wait(dm_smem_u32(&full[pair][st]), (tile_no / DM_STAGES) & 1u);
                const double *tile = ring + (size_t)st * (8 * DM_RS);
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const double *trow = tile + (4 * hh + q) * DM_RS + r;            // B fragment: k = q (row 4hh + q), col = r
#pragma unroll
                    for (int t = 0; t < 16; t++) dmma_m8n8k4(accA[t][0], accA[t][1], bA[hh], trow[8 * t]);
                }
                __syncwarp();
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dm_smem_u32(&empty[pair][st])) : "memory");
                    issue();
                }
            }
            // part-A sums of this item: D fragment row = c = r, columns n = 8 t + 2 q, + 1
            if (colok) {
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    const int64_t j = (int64_t)it.R * DG_BLOCK + 8 * t + 2 * q;
                    if (j < N && accA[t][0] != 0.0) atomicAdd(Y + (int64_t)r * N + j, accA[t][0]);
                    if (j + 1 < N && accA[t][1] != 0.0) atomicAdd(Y + (int64_t)r * N + j + 1, accA[t][1]);
                }
            }
        }
    } else {
        // ---- part B: D[i][c] = sum_n panel[i][n] b[128R + n][c], 32 MMAs over n = 4 j + q ----
        for (int64_t tc = wid; tc < n_items; tc += W) {
            const dg_item it = items[tc];
            const int ng = it.rows >> 3;
            // the second operand for this block-row: b[128 R + 4 j + q][c = r], zero for padding columns / samples
            double tb[32];
            if (it.partB) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int64_t jj = (int64_t)it.R * DG_BLOCK + 4 * j + q;
                    tb[j] = (colok && jj < N) ? B[(int64_t)r * N + jj] : 0.0;
                }
            }
            for (int g = 0; g < ng; g++, tile_no++) {
                const int st = (int)(tile_no % DM_STAGES);
                // also for a diagonal item (no part B): waiting for the copy keeps this warp within one ring of its partner, so
                // that the two arrivals that free a stage always belong to the same tile
                dm_wait(dm_smem_u32(&full[pair][st]), (tile_no / DM_STAGES) & 1u);
                if (it.partB) {
                    const double *ta = ring + (size_t)st * (8 * DM_RS) + r * DM_RS + q;
                    // four independent accumulator pairs: a single chain of 32 dependent MMAs would be latency-bound
                    double dd[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int j = 0; j < 32; j++) dmma_m8n8k4(dd[j & 3][0], dd[j & 3][1], ta[4 * j], tb[j]);
                    const double d0 = (dd[0][0] + dd[1][0]) + (dd[2][0] + dd[3][0]), d1 = (dd[0][1] + dd[1][1]) + (dd[2][1] + dd[3][1]);
                    const int64_t i = (int64_t)it.i0 + (int64_t)g * 8 + r;           // D fragment: row = r, columns 2q, 2q + 1
                    if (i < N) {
                        if (2 * q < kc && d0 != 0.0) atomicAdd(Y + (int64_t)(2 * q) * N + i, d0);
                        if (2 * q + 1 < kc && d1 != 0.0) atomicAdd(Y + (int64_t)(2 * q + 1) * N + i, d1);
                    }
                    __syncwarp();
                }
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dm_smem_u32(&empty[pair][st])) : "memory");
            }
        }
    }
}

static void symm_mma_launch(sgb_ctx *h, const sgb_dense *d, const double *B, int kc, int64_t N, double *Y)
{
    const size_t smem = sizeof(double) * (size_t)(8 * DM_STAGES * 8 * DM_RS);
    if (sgb_first_on_device(h->device, SGB_SITE_SYMV_BASE + 0))
        cudaFuncSetAttribute(dense_symm_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (!d->n_items) return;
    const int64_t grid = std::min<int64_t>(h->sm_count, (d->n_items + 7) / 8);
    dense_symm_mma_kernel<<<(unsigned)grid, 512, smem, h->stream>>>(d->pool, d->d_items, d->n_items, B, kc, N, Y);
    h->cnt.n_kernel_launches += 1;
}

// out[a + b*ni] = K[i0+a][j0+b] if this rank stores it, else 0
__global__ void dense_get_block_kernel(const double *__restrict__ pool, const int64_t *__restrict__ off, int64_t i0, int64_t ni,
                                       int64_t j0, int64_t nj, double *__restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ni * nj) return;
    int64_t a = i0 + e % ni, b = j0 + e / ni;
    if ((a >> 7) < (b >> 7)) { int64_t t = a; a = b; b = t; }
    const int64_t R = a >> 7, o = off[R];
    out[e] = o < 0 ? 0.0 : pool[o + b * DG_BLOCK + (a & 127)];
}

// ---------------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------------
struct dg_shard { const uint8_t *gt; int64_t sT, Mloc; int8_t *dig; bool owned_copy; };

static void shard_markers(const sgb_ctx *h, int q, std::vector<int64_t> &glob)
{
    glob.clear();
    for (int64_t g = 0; g < h->M; g++)
        if (sgb_owner_of(h, g) == q) glob.push_back(g);
}

static int dense_build(sgb_ctx *h, int limbs, int64_t first_block_row, int64_t n_block_rows)
{
    if (!h->loaded) return sgb_fail(h, "genotypes not loaded: call setgeno first");
    if (limbs < 2 || limbs > 8) return sgb_fail(h, "dense GRM: weight limbs must be 2..8 (got %d)", limbs);
    if ((int64_t)h->ac.size() != h->M) return sgb_fail(h, "dense GRM: allele counts unavailable");
    if (h->N >= (1ll << 18)) return sgb_fail(h, "dense GRM: N = %lld exceeds the 128-bit centring bound (262143 samples)", (long long)h->N);
    if (h->M > (int64_t)8000000) return sgb_fail(h, "dense GRM: M = %lld exceeds the int32 accumulation bound (8M markers)", (long long)h->M);
    sgb_dense_free(h);
    sgb_dense *d = new sgb_dense();
    h->dense = d;
    d->limbs = limbs;
    const int64_t N = h->N;
    d->nbr = cdiv64(N, DG_BLOCK);
    // bench: a bounded sample of the build (block-rows [first, first + n)); the product needs all of them
    int64_t br0 = 0, nbr_build = d->nbr;
    if (n_block_rows > 0) {
        br0 = std::min(std::max<int64_t>(first_block_row, 0), d->nbr);
        nbr_build = std::min(d->nbr, br0 + n_block_rows);
        d->partial = true;
    }

    // ---- fixed-point weights: the same S on every rank ----
    double wmax = 0.0;
    std::vector<double> s2all((size_t)h->M);
    for (int64_t g = 0; g < h->M; g++) {
        double f = (double)h->ac[g] / (double)(2 * N), v = 2.0 * f * (1.0 - f);
        double s = v > 0 ? 1.0 / sqrt(v) : 0.0;
        s2all[g] = s * s;
        wmax = std::max(wmax, s2all[g]);
    }
    if (wmax <= 0) return sgb_fail(h, "dense GRM: no polymorphic marker");
    int e2 = 0;
    frexp(wmax, &e2);                       // wmax < 2^e2
    d->S = 7 * limbs - 2 - e2;              // W < 2^(7 limbs - 2): fits `limbs` balanced base-128 digits
    const double scaleW = ldexp(1.0, d->S);

    // ---- shards: local sample-major store + (world > 1) copies of the other ranks' ----
    std::vector<dg_shard> sh((size_t)h->world);
    std::vector<int64_t> glob;
    unsigned __int128 Csum = 0;
    std::vector<double> vloc((size_t)h->rowsG * DG_UPIECES, 0.0);
    double sumv[DG_UPIECES] = {0, 0, 0, 0, 0};
    for (int q = 0; q < h->world; q++) {
        shard_markers(h, q, glob);
        dg_shard &s = sh[(size_t)q];
        s.Mloc = (int64_t)glob.size();
        const int64_t rowsGq = round_up64(std::max<int64_t>(s.Mloc, 1), SGB_ROW_ALIGN);
        s.sT = round_up64((rowsGq + 3) / 4, SGB_KSTEP_BYTES);
        if (q == h->rank && (s.Mloc != h->Mloc || s.sT != h->sT)) return sgb_fail(h, "dense GRM: shard map mismatch");
        std::vector<int8_t> dig((size_t)limbs * s.sT * 4, 0);
        for (int64_t r = 0; r < s.Mloc; r++) {
            const int64_t g = glob[(size_t)r];
            long long W = llround(s2all[g] * scaleW);
            const long long cm = 2 * (long long)N - (long long)h->ac[g];        // N phi_m = sum_i h_im
            Csum += (unsigned __int128)W * (unsigned __int128)(cm * cm);
            if (q == h->rank)
                for (int p = 0; p < DG_UPIECES; p++) {
                    const double vp = (double)(((W >> (12 * p)) & 4095ll) * cm);
                    vloc[(size_t)p * h->rowsG + r] = vp;
                    sumv[p] += vp;
                }
            for (int l = 0; l < limbs; l++) {
                int dgt = (int)((W + 64) & 127) - 64;
                W = (W - dgt) >> 7;
                dig[(size_t)l * s.sT * 4 + r] = (int8_t)dgt;
            }
            if (W != 0) return sgb_fail(h, "dense GRM: weight digit overflow");
        }
        CUDA_OK(h, cudaMalloc((void **)&s.dig, dig.size()));
        CUDA_OK(h, cudaMemcpyAsync(s.dig, dig.data(), dig.size(), cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        if (q == h->rank) { s.gt = h->dGt; s.owned_copy = false; }
        else {
            uint8_t *p = nullptr;
            CUDA_OK(h, cudaMalloc((void **)&p, (size_t)h->rowsT * s.sT));
            s.gt = p; s.owned_copy = true;
        }
    }
    for (int q = 0; q < h->world && h->world > 1; q++)
        SGB_TRY(sgb_broadcast_bytes(h, (void *)sh[(size_t)q].gt, (size_t)h->rowsT * sh[(size_t)q].sT, q));

    // ---- U'_i over all markers: one DG_UPIECES-column sweep of the tensor engine (exact: 12-bit pieces, sums < 2^53) ----
    CUDA_OK(h, cudaMalloc((void **)&d->dU, sizeof(double) * N * DG_UPIECES));
    {
        const int64_t rowsT = h->rowsT;
        double *dv = nullptr, *raw = nullptr;
        CUDA_OK(h, cudaMalloc((void **)&dv, sizeof(double) * (size_t)(h->rowsG + rowsT + 1) * DG_UPIECES));
        raw = dv + h->rowsG * DG_UPIECES;
        CUDA_OK(h, cudaMemcpyAsync(dv, vloc.data(), sizeof(double) * h->rowsG * DG_UPIECES, cudaMemcpyHostToDevice, h->stream));
        int rcu = sgb_gt_times_cols(h, dv, DG_UPIECES, raw);
        if (!rcu) {
            cudaMemcpyAsync(raw + rowsT * DG_UPIECES, sumv, sizeof(double) * DG_UPIECES, cudaMemcpyHostToDevice, h->stream);
            if (h->world > 1) rcu = sgb_allreduce_sum(h, raw, (rowsT + 1) * DG_UPIECES);
        }
        if (!rcu) {
            syrk_u_kernel<<<dim3((unsigned)cdiv64(N, 256), DG_UPIECES), 256, 0, h->stream>>>(raw, rowsT, N, raw + rowsT * DG_UPIECES, d->dU);
            h->cnt.n_kernel_launches++;
        }
        cudaStreamSynchronize(h->stream);
        cudaFree(dv);
        if (rcu) return rcu;
    }

    // ---- storage of the block-rows this rank owns ----
    d->off.assign((size_t)d->nbr, -1);
    size_t elems = 0;
    std::vector<dg_item> items, diag_items;
    for (int64_t R = br0; R < nbr_build; R++) {
        if (R % h->world != h->rank) continue;
        d->off[(size_t)R] = (int64_t)elems;
        const int64_t rows = DG_BLOCK * (R + 1);
        for (int64_t i0 = 0; i0 < DG_BLOCK * R; i0 += DG_CHUNK) {
            dg_item it;
            it.off = (int64_t)elems; it.R = (int32_t)R; it.i0 = (int32_t)i0;
            it.rows = (int32_t)std::min<int64_t>(DG_CHUNK, DG_BLOCK * R - i0); it.partB = 1;
            items.push_back(it);
        }
        dg_item dgi;
        dgi.off = (int64_t)elems; dgi.R = (int32_t)R; dgi.i0 = (int32_t)(DG_BLOCK * R); dgi.rows = DG_BLOCK; dgi.partB = 0;
        diag_items.push_back(dgi);
        elems += (size_t)rows * DG_BLOCK;
    }
    d->n_items_b = (int64_t)items.size();
    items.insert(items.end(), diag_items.begin(), diag_items.end());
    d->pool_elems = elems;
    if (elems) {
        cudaError_t e = cudaMalloc((void **)&d->pool, sizeof(double) * elems);
        if (e != cudaSuccess) return sgb_fail(h, "dense GRM: cannot allocate %.1f GB for the stored matrix: %s", 8e-9 * (double)elems, cudaGetErrorString(e));
    }
    d->n_items = (int64_t)items.size();
    if (d->n_items) {
        CUDA_OK(h, cudaMalloc((void **)&d->d_items, sizeof(dg_item) * items.size()));
        CUDA_OK(h, cudaMemcpyAsync(d->d_items, items.data(), sizeof(dg_item) * items.size(), cudaMemcpyHostToDevice, h->stream));
    }

    // ---- the contraction: per block-row, per limb, per shard: image -> tcgen05 product -> fold ----
    int64_t sTmax = 0;
    for (auto &s : sh) sTmax = std::max(sTmax, s.sT);
    int8_t *img = nullptr;
    int32_t *acc = nullptr;
    const size_t img_bytes = k_umma_image_bytes(DG_BLOCK, sTmax);
    const size_t acc_elems = (size_t)round_up64(N, DG_BLOCK) * DG_BLOCK;
    CUDA_OK(h, cudaMalloc((void **)&img, img_bytes));
    CUDA_OK(h, cudaMalloc((void **)&acc, sizeof(int32_t) * acc_elems));
    double *qhi = nullptr;
    if (limbs > 4) CUDA_OK(h, cudaMalloc((void **)&qhi, sizeof(double) * acc_elems));
    CUDA_OK(h, cudaMemsetAsync(acc, 0, sizeof(int32_t) * acc_elems, h->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    CUDA_OK(h, cudaEventRecord(e0, h->stream));
    const double mul = ldexp(1.0, -d->S) / ((double)h->M * (double)N * (double)N);
    const unsigned long long c_hi = (unsigned long long)(Csum >> 64), c_lo = (unsigned long long)Csum;
    int rc = 0;
    for (int64_t R = br0; R < nbr_build && !rc; R++) {
        if (d->off[(size_t)R] < 0) continue;
        double *panel = d->pool + d->off[(size_t)R];
        const int64_t rows = DG_BLOCK * (R + 1);
        for (int l = 0; l < limbs && !rc; l++) {
            for (int q = 0; q < h->world && !rc; q++) {
                const dg_shard &s = sh[(size_t)q];
                if (s.Mloc == 0) continue;
                const int64_t nblk = s.sT / 32;
                syrk_image_kernel<<<(unsigned)cdiv64(nblk * 1024, 256), 256, 0, h->stream>>>(s.gt, s.sT, DG_BLOCK * R, s.dig + (size_t)l * s.sT * 4, nblk, img);
                h->cnt.n_kernel_launches++;
                h->umma_accumulate = q > 0;            // shards after the first add to the int32 sums
                rc = k_pk2_umma_rows(h, s.gt, s.sT, rows, s.sT, img, DG_BLOCK, acc, SGB_PLANE_VALUE);
                h->umma_accumulate = false;
                d->tensor_ops += 2.0 * (double)rows * DG_BLOCK * (double)(s.sT * 4);
            }
            int64_t n = rows * DG_BLOCK;
            int gb = (int)std::min<int64_t>(cdiv64(n, 256 * 4), (int64_t)h->sm_count * 8);
            syrk_fold_kernel<<<gb, 256, 0, h->stream>>>(acc, n, l, panel, qhi);
            h->cnt.n_kernel_launches++;
        }
        if (rc) break;
        int64_t n = rows * DG_BLOCK;
        int gb = (int)std::min<int64_t>(cdiv64(n, 256 * 4), (int64_t)h->sm_count * 8);
        syrk_finalize_kernel<<<gb, 256, 0, h->stream>>>(panel, qhi, limbs > 4, rows, DG_BLOCK * R, N, d->dU, c_hi, c_lo, mul);
        h->cnt.n_kernel_launches++;
    }
    cudaEventRecord(e1, h->stream);
    cudaError_t es = cudaStreamSynchronize(h->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    d->build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(img); cudaFree(acc); if (qhi) cudaFree(qhi);
    for (auto &s : sh) { cudaFree(s.dig); if (s.owned_copy) cudaFree((void *)s.gt); }
    if (rc) return rc;
    if (es != cudaSuccess) return sgb_fail(h, "dense GRM build failed: %s", cudaGetErrorString(es));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return sgb_fail(h, "dense GRM build failed: %s", cudaGetErrorString(le));
    return 0;
}

// Y = K B from the stored matrix (dB, dY device, ld = N; dY may alias dB)
int sgb_dense_crossprod_device(sgb_ctx *h, const double *dB, int k, double *dY)
{
    sgb_dense *d = h->dense;
    if (!d || d->partial) return sgb_fail(h, "stored-GRM product requested before sgb_dense_grm_build");
    const int64_t N = h->N;
    SGB_TRY(sgb_ensure_f64(h, &h->d_tmp, &h->tmp_elems, (size_t)N * k));
    double *acc = h->d_tmp;
    CUDA_OK(h, cudaMemsetAsync(acc, 0, sizeof(double) * N * k, h->stream));
    h->cnt.n_crossprod_calls++; h->cnt.n_crossprod_columns += k;
    if (d->n_items) {
        int c = 0;
        while (c < k) {
            const int left = k - c;
            const double *Bc = dB + (int64_t)c * N;
            double *Yc = acc + (int64_t)c * N;
            // >= 3 columns: fp64 tensor-pipe kernel, 8 columns per pass (the DFMA kernels stay selectable with the fp64 engine)
            if (left >= 3 && h->engine != SGB_ENGINE_F64) { const int kc = left < 8 ? left : 8; symm_mma_launch(h, d, Bc, kc, N, Yc); c += kc; }
            else if (left >= 8) { symv_launch<8>(h, d, Bc, N, Yc); c += 8; }
            else if (left >= 4) { symv_launch<4>(h, d, Bc, N, Yc); c += 4; }
            else if (left >= 2) { symv_launch<2>(h, d, Bc, N, Yc); c += 2; }
            else { symv_launch<1>(h, d, Bc, N, Yc); c += 1; }
            cudaError_t e__ = cudaGetLastError();
            if (e__ != cudaSuccess) return sgb_fail(h, "stored-GRM product launch failed: %s", cudaGetErrorString(e__));
        }
    }
    if (h->world > 1) SGB_TRY(sgb_allreduce_sum(h, acc, N * k));
    CUDA_OK(h, cudaMemcpyAsync(dY, acc, sizeof(double) * N * k, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

extern "C" int sgb_dense_grm_build(sgb_ctx *h, int weight_limbs)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    int rc = dense_build(h, weight_limbs, 0, 0);
    if (rc) sgb_dense_free(h);
    return rc;
}

extern "C" int sgb_dense_grm_free(sgb_ctx *h)
{
    sgb_dense_free(h);
    return 0;
}

extern "C" int sgb_set_grm_mode(sgb_ctx *h, int mode)
{
    if (mode != SGB_GRM_PACKED && mode != SGB_GRM_DENSE) return sgb_fail(h, "unknown GRM mode %d", mode);
    if (mode == SGB_GRM_DENSE && (!h->dense || h->dense->partial)) return sgb_fail(h, "stored-GRM mode requested before sgb_dense_grm_build");
    h->grm_mode = mode;
    h->ku_cols = 0;            // cached probe products belong to the other GRM representation
    return 0;
}

extern "C" int sgb_dense_grm_get_block(sgb_ctx *h, int64_t i0, int64_t ni, int64_t j0, int64_t nj, double *out)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    sgb_dense *d = h->dense;
    if (!d) return sgb_fail(h, "dense GRM not built");
    if (i0 < 0 || j0 < 0 || ni < 0 || nj < 0 || i0 + ni > h->N || j0 + nj > h->N) return sgb_fail(h, "dense GRM block out of range");
    if (ni * nj == 0) return 0;
    SGB_TRY(sgb_ensure_f64(h, &h->d_io, &h->io_elems, (size_t)(ni * nj) + (size_t)d->nbr));
    double *dout = h->d_io;
    int64_t *doff = reinterpret_cast<int64_t *>(h->d_io + ni * nj);
    CUDA_OK(h, cudaMemcpyAsync(doff, d->off.data(), sizeof(int64_t) * d->nbr, cudaMemcpyHostToDevice, h->stream));
    dense_get_block_kernel<<<(unsigned)cdiv64(ni * nj, 256), 256, 0, h->stream>>>(d->pool, doff, i0, ni, j0, nj, dout);
    DG_LAUNCH_CHECK(h);
    if (h->world > 1) SGB_TRY(sgb_allreduce_sum(h, dout, ni * nj));
    CUDA_OK(h, cudaMemcpyAsync(out, dout, sizeof(double) * ni * nj, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->cnt.bytes_d2h += sizeof(double) * ni * nj;
    return 0;
}

extern "C" int sgb_dense_grm_info(sgb_ctx *h, double *out6)
{
    sgb_dense *d = h->dense;
    if (!d) return sgb_fail(h, "dense GRM not built");
    out6[0] = (double)d->limbs; out6[1] = (double)d->S; out6[2] = (double)d->nbr;
    out6[3] = 8.0 * (double)d->pool_elems; out6[4] = d->build_ms; out6[5] = d->tensor_ops;
    return 0;
}

// bench: build only block-rows [first_block_row, first_block_row + n_block_rows) (the cost of one grows linearly with its index)
extern "C" int sgb_bench_dense_build(sgb_ctx *h, int weight_limbs, int64_t first_block_row, int64_t n_block_rows)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    int rc = dense_build(h, weight_limbs, first_block_row, n_block_rows);
    if (rc) sgb_dense_free(h);
    return rc;
}
