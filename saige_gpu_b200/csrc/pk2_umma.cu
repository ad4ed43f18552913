// tcgen05 ("UMMA") version of the packed-genotype x limb product for WIDE right-hand-side batches (k >= 3 columns).
//
//   out[r][n] += sum_k (c0 - plane(P[r][k])) * L[k][n]        n = column*8 + limb,  N = 16..128 per launch
//
// Blackwell-native data path, no shared-memory round trip for the genotypes:
//   * every producer thread owns ONE ROW of a 128-row tile (= one TMEM lane): it streams its row's packed bytes with
//     256-bit loads, decodes them in registers with prmt (same pair-ternary trick as pk2_gemm_kernel) and writes the u8
//     A operand straight into TENSOR MEMORY with tcgen05.st (4 k-values per 32-bit column);
//   * the int8 limb operand B is staged in shared memory in the canonical K-major no-swizzle UMMA layout; the limb
//     splitter writes the global copy as an image of that layout, so staging is one cp.async.bulk (TMA engine) per
//     128-genotype block, tracked by an mbarrier with expect_tx;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N, K=32) with A from TMEM, B from the smem
//     descriptor and the int32 accumulator in TMEM; tcgen05.commit -> mbarriers release the A stage and the B stage;
//   * the epilogue reads the accumulator with tcgen05.ld and adds it to the global int32 limb sums.
// Integer arithmetic is exact, so results are bit-identical to the mma.sync kernel (the k <= 2 path and cross-check).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include "sgb_internal.h"
#include "recombine.cuh"

#define UMMA_ROWS 128            // rows per CTA tile = TMEM lanes = MMA M
#define UMMA_KSTEP 256           // genotypes per pipeline step  (64 packed bytes per row, 8 MMAs of K=32)
#define UMMA_KBLK 128            // genotypes per block of the limb image (one bulk copy)
#define UMMA_A_COLS 64           // TMEM columns of one A stage   (256 k-values / 4 per column)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity));
}

__device__ __forceinline__ uint32_t prmt_u(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint4 ldg_stream_u(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// one 256-bit load = one full 32-byte sector per thread (sm_100 LDG.256): a row-per-thread access pattern would
// otherwise request every sector twice with 128-bit loads
struct u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg_stream_256(const uint8_t *p)
{
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}

struct umma_pools { uint32_t ax, ay, bx, by; };

// 32 registers (one per TMEM column) -> this thread's lane of the A stage
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, int32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]     M=128, N from idesc, K=32, int8 -> int32
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

#define UMMA_MAX_STAGES 3        // A (TMEM) pipeline depth: 2 stages of 64 columns, or 3 when the accumulator needs <= 64 columns
#define UMMA_MAX_BSTAGES 16      // B (smem) ring depth is chosen at launch: as many 128 x N byte stages as fit ~96 KB
#define UMMA_PF 3                // packed-row prefetch distance of the producer threads, in steps
#define UMMA_L2_AHEAD 12         // L2 prefetch distance of the producers, in k-steps of 64 bytes (8..16 measured equal)
#define UMMA_PGROUPS 2           // producer groups of 4 warps; group g owns the k-steps s = g (mod UMMA_PGROUPS)
#define UMMA_PWARPS (4 * UMMA_PGROUPS)
#define UMMA_THREADS (32 * (UMMA_PWARPS + 2))   // producers (+ epilogue), then one MMA-issuer warp and one bulk-copy warp

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


// MMA issue loop, run by a WHOLE warp: every value in it is warp-uniform, so the compiler keeps descriptors, addresses
// and barrier state on the uniform datapath, and one lane issues.  (A single thread running this loop needed ~12
// dependent instructions per tcgen05.mma, the 63-cycle "floor" per MMA measured at small N, plus two integer divisions
// per step for the ring indices.)  Ring indices and phases are counters, the 8 B descriptors of a step are the stage-0
// descriptors plus one add.
__device__ __forceinline__ bool elect_one()
{
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    return leader != 0;
}

template <int STAGES>
__device__ __forceinline__ void umma_issue_loop(int nsteps, int nb, int N, uint32_t smem_b0, uint32_t stage_bytes, uint32_t tmem_base,
                                                uint32_t tmem_d, uint64_t *full_a, uint64_t *empty, uint64_t *full_b, uint64_t *empty_b,
                                                uint64_t *done_bar, int lane, int skip_mma)
{
    // instruction descriptor: D=s32, A=u8 (K-major, TMEM), B=s8 (K-major), N, M=128
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(UMMA_ROWS >> 4) << 24);
    const uint32_t half_bytes = stage_bytes >> 1, st16 = stage_bytes >> 4;
    // K-major, no swizzle: core matrix = 8 n-rows x 16 k-bytes (128 B); LBO (next 16 k) = 128 B; SBO (next 8 n) = 1024 B; version bit 46
    const uint32_t dhi = (uint32_t)(1024 >> 4) | (1u << 14);
    uint32_t dlo[8];
#pragma unroll
    for (int j = 0; j < 8; j++) dlo[j] = (((smem_b0 + (j >> 2) * half_bytes + (j & 3) * 256) >> 4) & 0x3FFF) | ((uint32_t)(128 >> 4) << 16);
    int st = 0, sbi = 0;
    uint32_t pa = 0, pb = 0, boff = 0, ta = tmem_base;
    for (int s = 0; s < nsteps; s++) {
        mbar_wait(&full_b[sbi], pb);
        mbar_wait(&full_a[st], pa);
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (elect_one()) {
            if (!skip_mma) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    umma_i8_ts(tmem_d, ta + j * 8, ((uint64_t)dhi << 32) | (uint64_t)(dlo[j] + boff), idesc, (s > 0 || j > 0) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_b[sbi])) : "memory");
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; pa ^= 1; ta = tmem_base; } else ta += UMMA_A_COLS;
        if (++sbi == nb) { sbi = 0; pb ^= 1; boff = 0; } else boff += st16;
    }
    if (lane == 0 && nsteps > 0)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done_bar)) : "memory");
}

// grid: x = k-chunks, y = 128-row tiles.  block: 320 threads (8 producer warps, MMA warp, bulk-copy warp).
// dynamic smem: nb B stages of 256 x N bytes.
//   P        packed rows (pair-ternary) in the TILED layout (sgb_tiled_off), `stride` bytes per row (multiple of 64); rows padded to a multiple of 128
//   L        limb operand as an image of the smem stage: [k-block of 128][N/8][k/16 (8)][n%8 (8)][k%16 (16)] int8
//   out      int32 [rows][ldo]; this launch covers columns [n0, n0 + N)
// Warp-specialised pipeline, UMMA_STAGES deep:
//   producers : load 64 B of their row (register ring UMMA_PF steps ahead, L2 prefetch UMMA_L2_AHEAD steps ahead) -> prmt decode
//               -> tcgen05.st into the A stage -> one arrive per warp on full_a
//   bulk warp : cp.async.bulk (TMA engine) of the next B stage image -> full_b (expect_tx)
//   MMA warp  : wait full_a & full_b -> 8 x tcgen05.mma.kind::i8 (K = 32 each) -> tcgen05.commit -> empty (frees both stages)
template <int UMMA_STAGES>
__global__ void __launch_bounds__(UMMA_THREADS, 2)
pk2_umma_kernel(const uint8_t *__restrict__ P, int64_t stride, int64_t ksteps_total, int ksteps_per_chunk,
                const int8_t *__restrict__ L, int N, int64_t Lblk_stride, int32_t *__restrict__ out, int ldo, int n0,
                int use_atomic, umma_pools pool, int tmem_cols, int nb, int dbg_arg)
{
    // Ablation switches (drop the loads / decode + tcgen05.st / MMAs / B copies; results are then wrong): only in builds
    // with -DSGB_ABLATION (make ABLATION=1), where SGB_UMMA_DBG selects them at run time.  DESIGN.md 3.1b has the findings.
#ifdef SGB_ABLATION
    const int dbg = dbg_arg & 255;
    const int l2_ahead = (dbg_arg >> 8) ? (dbg_arg >> 8) : UMMA_L2_AHEAD;
#else
    constexpr int dbg = 0;
    constexpr int l2_ahead = UMMA_L2_AHEAD;
    (void)dbg_arg;
#endif
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full_a[UMMA_MAX_STAGES], empty[UMMA_MAX_STAGES], full_b[UMMA_MAX_BSTAGES], empty_b[UMMA_MAX_BSTAGES], done_bar;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int64_t ks0 = (int64_t)blockIdx.x * ksteps_per_chunk;
    int64_t ks1 = ks0 + ksteps_per_chunk;
    if (ks1 > ksteps_total) ks1 = ksteps_total;
    const int nsteps = (int)(ks1 - ks0);
    const uint32_t stage_bytes = (uint32_t)UMMA_KSTEP * (uint32_t)N;      // two 128-genotype limb blocks
    const uint32_t half_bytes = stage_bytes / 2;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        for (int i = 0; i < UMMA_STAGES; i++) { mbar_init(&full_a[i], 4); mbar_init(&empty[i], 1); }     // one arrive per producer warp
        for (int i = 0; i < nb; i++) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_d = tmem_base + UMMA_STAGES * UMMA_A_COLS;

    if (warp < UMMA_PWARPS) {
        // ================= producers: one row per thread, k-steps interleaved over the producer groups =================
        const int grp = warp >> 2, q4 = warp & 3;
        const int64_t row = (int64_t)blockIdx.y * UMMA_ROWS + q4 * 32 + (tid & 31);
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;         // this warp's TMEM lane quarter
        // tiled store: this CTA's 128-row tile is one panel; k-slab s of the panel is 8 KB on, the thread's row 64 bytes in
        const uint8_t *prow = P + (int64_t)blockIdx.y * UMMA_ROWS * stride + ks0 * SGB_SLAB_BYTES + (int64_t)(q4 * 32 + (tid & 31)) * 64;
        u32x8 pf[UMMA_PF][2];
#pragma unroll
        for (int i = 0; i < UMMA_PF; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) { pf[i][0].v[j] = 0; pf[i][1].v[j] = 0; }
            const int sp = grp + i * UMMA_PGROUPS;
            if (sp < nsteps) { pf[i][0] = ldg_stream_256(prow + (int64_t)sp * SGB_SLAB_BYTES); pf[i][1] = ldg_stream_256(prow + (int64_t)sp * SGB_SLAB_BYTES + 32); }
        }
        // the prefetch ring is indexed statically (loop unrolled by UMMA_PF): shifting a register queue would make every
        // step wait for ALL loads in flight
        for (int sbase = grp; sbase < nsteps; sbase += UMMA_PGROUPS * UMMA_PF) {
#pragma unroll
            for (int j = 0; j < UMMA_PF; j++) {
                const int s = sbase + j * UMMA_PGROUPS;
                if (s < nsteps) {
                    const int st = s % UMMA_STAGES, u = s / UMMA_STAGES;
                    if (u > 0) mbar_wait(&empty[st], (uint32_t)((u - 1) & 1));      // MMAs that read this A stage have completed
                    asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        if (dbg & 2) break;
                        uint32_t a[32];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t wv = pf[j][hf].v[i], hi = wv >> 16;
                            a[4 * i + 0] = prmt_u(pool.ax, pool.ay, wv);
                            a[4 * i + 1] = prmt_u(pool.ax, pool.ay, hi);
                            a[4 * i + 2] = prmt_u(pool.bx, pool.by, wv);
                            a[4 * i + 3] = prmt_u(pool.bx, pool.by, hi);
                        }
                        tmem_st_x32(tmem_base + lane_base + st * UMMA_A_COLS + hf * 32, a);
                    }
                    if (s + UMMA_PF * UMMA_PGROUPS < nsteps && !(dbg & 4)) {
                        pf[j][0] = ldg_stream_256(prow + (int64_t)(s + UMMA_PF * UMMA_PGROUPS) * SGB_SLAB_BYTES);
                        pf[j][1] = ldg_stream_256(prow + (int64_t)(s + UMMA_PF * UMMA_PGROUPS) * SGB_SLAB_BYTES + 32);
                    }
                    // the register prefetch holds ~100 KB per SM in flight, not enough to cover DRAM latency at full rate:
                    // group 0 also pulls the row's 128-byte line `l2_ahead` steps further on into L2 (one instruction)
                    if (grp == 0 && !(tid & 1) && s + l2_ahead < nsteps)          // even lanes: one 128-byte line covers two rows
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + (int64_t)(s + l2_ahead) * SGB_SLAB_BYTES));
                    asm volatile("tcgen05.wait::st.sync.aligned;");
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    // 128 per-thread arrives on one mbarrier serialise (~470 cycles per step measured): one per warp instead
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&full_a[st]);
                }
            }
        }
        // ---- epilogue: all MMAs committed -> TMEM -> registers -> global int32 sums (columns split over the groups) ----
        if (nsteps > 0) {
            mbar_wait(&done_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;");
            int32_t *o = out + row * ldo + n0;
            for (int c = grp * 16; c < N; c += 16 * UMMA_PGROUPS) {
                int32_t v[16];
                tmem_ld_x16(tmem_d + lane_base + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                if (use_atomic) {
#pragma unroll
                    for (int i = 0; i < 16; i++) if (v[i]) atomicAdd(o + c + i, v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<int4 *>(o + c + i) = make_int4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    } else if (warp == UMMA_PWARPS) {
        // ================= MMA issuer warp =================
        umma_issue_loop<UMMA_STAGES>(nsteps, nb, N, smem_u32(smem), stage_bytes, tmem_base, tmem_d, full_a, empty, full_b, empty_b, &done_bar,
                                     tid & 31, dbg & 1);
    } else {
        // ================= B stage loader: one bulk copy (TMA engine) per step =================
        if (tid == 32 * (UMMA_PWARPS + 1)) {
            int sbi = 0;
            uint32_t pb = 1;                               // parity of the previous use of the stage
            bool wrapped = false;
            const int8_t *src = L + (2 * ks0) * Lblk_stride;
            for (int s = 0; s < nsteps; s++) {
                if (wrapped) mbar_wait(&empty_b[sbi], pb);
                if (dbg & 8) { mbar_arrive(&full_b[sbi]); }
                else {
                    mbar_expect_tx(&full_b[sbi], stage_bytes);
                    bulk_g2s(smem + sbi * stage_bytes, src, half_bytes, &full_b[sbi]);
                    bulk_g2s(smem + sbi * stage_bytes + half_bytes, src + Lblk_stride, half_bytes, &full_b[sbi]);
                }
                src += 2 * Lblk_stride;
                if (++sbi == nb) { sbi = 0; pb ^= 1; wrapped = true; }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

// Limb split for the UMMA engine: V[len x k] -> stage images + per-column multiplier + exact limb column sums.
// The B operand is s8, so a limb carries 8 bits: q = round(v 2^(53-E)) (|q| <= 2^54) is written with UMMA_LIMBS = 7 balanced
// base-256 digits (7 x 8 = 56 bits, the same range as the 8 x 7 bits of the mma.sync engine; both engines hold the same
// integer q, hence bit-identical results) -- 1/8 fewer accumulator columns and tensor work per right-hand side.
// Accumulator column of (column c, limb l) is n = 7 c + l; the image keeps 8 n-rows per 1 KB core-matrix group.
// One thread per (128-genotype block, TMEM column 0..31): the 4 genotype slots whose k-values share that column.
#define UMMA_LIMBS 7             // digits of the full-precision split (the dense-GRM build and the diagonals always use it)
template <int NL>
__global__ void split_limbs_umma_kernel(const double *__restrict__ V, int64_t len, int64_t ld, int k, int ngroups, int64_t nblk,
                                        const unsigned long long *__restrict__ mx, int8_t *__restrict__ L,
                                        double *__restrict__ mult, int32_t *__restrict__ limbsum)
{
    const int c = blockIdx.y;                         // column 0..k-1
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t blk = u >> 5;
    const int col = (int)(u & 31), w = col >> 2, q = col & 3;
    const bool inb = blk < nblk;
    unsigned long long mb = mx[c];
    int E = (int)((mb >> 52) & 0x7FF) - 1023;
    if (E < -1000) E = -1000;
    if (E > 1000) E = 1000;
    // |v| < 2^(E+1)  =>  |q| <= 2^(8 NL - 2), inside the range of NL balanced base-256 digits; NL = 7: 53 - E as on the mma.sync engine
    constexpr int SH = 8 * NL - 3;
    if (u == 0) mult[c] = mb ? scalbn(1.0, E - SH) : 0.0;
    // slot s of this column is genotype 16 w + {0,8,1,9}[q] + 2 s of the block (the order decode produces)
    const int64_t i0 = blk * UMMA_KBLK + 16 * w + ((q & 1) ? 8 : 0) + ((q & 2) ? 1 : 0);
    long long qv[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        int64_t i = i0 + 2 * s;
        qv[s] = (inb && i < len && mb) ? __double2ll_rn(scalbn(V[(int64_t)c * ld + i], SH - E)) : 0;
    }
#pragma unroll
    for (int l = 0; l < NL; l++) {
        uint32_t word = 0;
        int ssum = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            int d = (int)((qv[s] + 128) & 255) - 128;
            qv[s] = (qv[s] - d) >> 8;
            word |= (uint32_t)(d & 255) << (8 * s);
            ssum += d;
        }
        const int n = NL * c + l;
        if (inb) *reinterpret_cast<uint32_t *>(L + (blk * (int64_t)ngroups + (n >> 3)) * 1024 + w * 128 + (n & 7) * 16 + 4 * q) = word;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        if ((threadIdx.x & 31) == 0 && ssum) atomicAdd(&limbsum[c * NL + l], ssum);
    }
}

// The same split with the image of a k-block assembled in SHARED memory and written out as one contiguous run of ngroups KB
// (128-bit stores): a block owns a 128-genotype k-block and walks over the columns eight at a time (one warp per column).  The
// per-column version above writes 16-byte pieces 128 bytes apart, two columns sharing every 32-byte sector from different
// blocks: 0.52 ms for the 112 MB image of a 31-column batch over 500k markers.  Padding rows of the image are zero without a
// separate memset.  Values, digits and limb sums are those of split_limbs_umma_kernel (integers: bit-identical).
template <int NL>
__global__ void __launch_bounds__(256) split_limbs_umma_wide_kernel(const double *__restrict__ V, int64_t len, int64_t ld, int k, int ngroups,
                                                                    int64_t nblk, const unsigned long long *__restrict__ mx,
                                                                    int8_t *__restrict__ L, double *__restrict__ mult, int32_t *__restrict__ limbsum)
{
    extern __shared__ __align__(16) uint8_t img[];                      // ngroups x 1024 bytes, then NL k limb sums of this block
    int32_t *lsum = reinterpret_cast<int32_t *>(img + (size_t)ngroups * 1024);
    const int nq = ngroups * 64;                                        // 16-byte pieces of the image
    const int col = threadIdx.x & 31, w = col >> 2, q = col & 3;
    constexpr int SH = 8 * NL - 3;
    for (int i = threadIdx.x; i < NL * k; i += 256) lsum[i] = 0;
    // A block walks over several k-blocks and keeps the limb sums in shared memory: one global atomic per (column, digit) and
    // BLOCK.  One per k-block (3908 x 217 atomics on 217 neighbouring words = a handful of L2 slices) took 0.3 of the 0.37 ms
    // this kernel needed for 31 columns x 500k markers.
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        __syncthreads();                                                // the previous image has been written out; lsum is zeroed
        for (int i = threadIdx.x; i < nq; i += 256) reinterpret_cast<int4 *>(img)[i] = make_int4(0, 0, 0, 0);
        __syncthreads();
        const int64_t i0 = blk * UMMA_KBLK + 16 * w + ((q & 1) ? 8 : 0) + ((q & 2) ? 1 : 0);
        for (int cb = 0; cb < k; cb += 8) {
            const int c = cb + (threadIdx.x >> 5);
            if (c >= k) continue;                                       // warp-uniform
            const unsigned long long mb = mx[c];
            int E = (int)((mb >> 52) & 0x7FF) - 1023;
            if (E < -1000) E = -1000;
            if (E > 1000) E = 1000;
            if (blk == 0 && col == 0) mult[c] = mb ? scalbn(1.0, E - SH) : 0.0;
            long long qv[4];
#pragma unroll
            for (int sl = 0; sl < 4; sl++) {
                const int64_t i = i0 + 2 * sl;
                qv[sl] = (i < len && mb) ? __double2ll_rn(scalbn(V[(int64_t)c * ld + i], SH - E)) : 0;
            }
#pragma unroll
            for (int l = 0; l < NL; l++) {
                uint32_t word = 0;
                int ssum = 0;
#pragma unroll
                for (int sl = 0; sl < 4; sl++) {
                    const int d = (int)((qv[sl] + 128) & 255) - 128;
                    qv[sl] = (qv[sl] - d) >> 8;
                    word |= (uint32_t)(d & 255) << (8 * sl);
                    ssum += d;
                }
                const int n = NL * c + l;
                *reinterpret_cast<uint32_t *>(img + (n >> 3) * 1024 + w * 128 + (n & 7) * 16 + 4 * q) = word;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                if (col == 0 && ssum) lsum[c * NL + l] += ssum;         // (column, digit) belongs to this warp alone
            }
        }
        __syncthreads();
        int4 *dst = reinterpret_cast<int4 *>(L + blk * (int64_t)ngroups * 1024);
        for (int i = threadIdx.x; i < nq; i += 256) dst[i] = reinterpret_cast<const int4 *>(img)[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NL * k; i += 256)
        if (lsum[i]) atomicAdd(&limbsum[i], lsum[i]);
}

// raw[r + c*ld] = (c0 * limbsum - sum_l acc[r][NL c+l] 256^l) * mult[c];  the accumulators are reset for the next product
template <int NL>
__global__ void recombine_umma_kernel(int32_t *__restrict__ acc, int64_t rows, int k, int npad, const double *__restrict__ mult,
                                      const int32_t *__restrict__ limbsum, int c0, double *__restrict__ raw, int64_t ld)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * k) return;
    int64_t r = idx / k;
    int c = (int)(idx - r * k);
    raw[r + (int64_t)c * ld] = sgb_recombine_umma<NL>(acc, r, c, npad, limbsum, c0) * mult[c];
}

__global__ void colmax_umma_kernel(const double *__restrict__ V, int64_t len, int64_t ld, unsigned long long *__restrict__ mx)
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

#define UMMA_LAUNCH_CHECK(h)                                                                            \
    do {                                                                                                \
        (h)->cnt.n_kernel_launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess) return sgb_fail(h, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, \
                                                cudaGetErrorString(e__));                               \
    } while (0)

static inline int umma_npad(int k, int nl) { return (nl * k + 15) & ~15; }        // accumulator columns: a multiple of 16
int k_umma_npad(int k, int nl) { return umma_npad(k, nl); }

// bytes of the UMMA limb operand for k columns (or `nrows` accumulator columns) over `kbytes` packed bytes per row
size_t k_umma_image_bytes(int nrows, int64_t kbytes)
{
    return (size_t)(kbytes / 32) * (size_t)(nrows / 8) * 1024;        // kbytes is a multiple of 64 => an even number of 128-genotype blocks
}
size_t k_umma_limb_bytes(int k, int64_t kbytes) { return k_umma_image_bytes(umma_npad(k, UMMA_LIMBS), kbytes); }

// nl = digits per value (5..7); have_stats: the column maxima are already in d_scal and the limb sums zeroed (fused product path)
int k_split_limbs_umma(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, int8_t *L, int64_t kbytes, double *d_mult,
                       int32_t *d_limbsum, int nl, int have_stats)
{
    unsigned long long *mx = reinterpret_cast<unsigned long long *>(h->d_scal + 2048);
    if (k > 1024) return sgb_fail(h, "too many columns (%d)", k);
    if (nl < 5 || nl > 7) return sgb_fail(h, "unsupported limb count %d", nl);
    const int npad = umma_npad(k, nl);
    const int64_t nblk = kbytes / 32;
    if (!have_stats) {
        CUDA_OK(h, cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * k, h->stream));
        CUDA_OK(h, cudaMemsetAsync(d_limbsum, 0, sizeof(int32_t) * nl * k, h->stream));
    }
    // images of up to 48 KB per k-block (k <= 54 at 7 digits) are assembled in shared memory and written as contiguous runs
    const bool wide = (size_t)(npad / 8) * 1024 <= 48 * 1024;
    // the padding rows (nl k .. npad-1) of the last core-matrix group(s) must read as zero limbs
    if (!wide && npad != nl * k) CUDA_OK(h, cudaMemsetAsync(L, 0, k_umma_image_bytes(npad, kbytes), h->stream));
    if (!have_stats) {
        int gx = (int)cdiv64(len, 256 * 8);
        if (gx > 1024) gx = 1024;
        if (gx < 1) gx = 1;
        colmax_umma_kernel<<<dim3(gx, k), 256, 0, h->stream>>>(V, len, ld, mx);
        UMMA_LAUNCH_CHECK(h);
    }
    if (wide) {
        if (nblk == 0) return 0;
        const size_t sb = (size_t)(npad / 8) * 1024 + sizeof(int32_t) * nl * k;        // <= 48 KB + 1.5 KB
        const unsigned gb = (unsigned)std::min<int64_t>(nblk, (int64_t)h->sm_count * 4);
        if (sb > 48 * 1024 && sgb_first_on_device(h->device, SGB_SITE_SPLIT_UMMA)) {
            CUDA_OK(h, cudaFuncSetAttribute(split_limbs_umma_wide_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
            CUDA_OK(h, cudaFuncSetAttribute(split_limbs_umma_wide_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
            CUDA_OK(h, cudaFuncSetAttribute(split_limbs_umma_wide_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
        }
        if (nl == 7) split_limbs_umma_wide_kernel<7><<<gb, 256, sb, h->stream>>>(V, len, ld, k, npad / 8, nblk, mx, L, d_mult, d_limbsum);
        else if (nl == 6) split_limbs_umma_wide_kernel<6><<<gb, 256, sb, h->stream>>>(V, len, ld, k, npad / 8, nblk, mx, L, d_mult, d_limbsum);
        else split_limbs_umma_wide_kernel<5><<<gb, 256, sb, h->stream>>>(V, len, ld, k, npad / 8, nblk, mx, L, d_mult, d_limbsum);
        UMMA_LAUNCH_CHECK(h);
        return 0;
    }
    dim3 grid((unsigned)cdiv64(nblk * 32, 256), k);
    if (nl == 7) split_limbs_umma_kernel<7><<<grid, 256, 0, h->stream>>>(V, len, ld, k, npad / 8, nblk, mx, L, d_mult, d_limbsum);
    else if (nl == 6) split_limbs_umma_kernel<6><<<grid, 256, 0, h->stream>>>(V, len, ld, k, npad / 8, nblk, mx, L, d_mult, d_limbsum);
    else split_limbs_umma_kernel<5><<<grid, 256, 0, h->stream>>>(V, len, ld, k, npad / 8, nblk, mx, L, d_mult, d_limbsum);
    UMMA_LAUNCH_CHECK(h);
    return 0;
}

int k_recombine_umma(sgb_ctx *h, int32_t *acc, int64_t rows, int k, const double *d_mult, const int32_t *d_limbsum, int plane,
                     double *raw, int64_t ld)
{
    if (rows * k == 0) return 0;
    recombine_umma_kernel<UMMA_LIMBS><<<(unsigned)cdiv64(rows * k, 256), 256, 0, h->stream>>>(acc, rows, k, umma_npad(k, UMMA_LIMBS), d_mult, d_limbsum,
                                                                                   plane == SGB_PLANE_VALUE ? 2 : 1, raw, ld);
    UMMA_LAUNCH_CHECK(h);
    return 0;
}

// out[r][n] (+)= sum_k (c0 - plane(P[r][k])) * image[k][n]   for `nrows` accumulator columns (a multiple of 16; out has
// ld = nrows), in balanced passes of N <= 128
int k_pk2_umma_rows(sgb_ctx *h, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t kbytes, const int8_t *L, int nrows,
                    int32_t *out, int plane)
{
    if (rows_pad % UMMA_ROWS || kbytes % 64 || stride % 64 || nrows % 16)
        return sgb_fail(h, "k_pk2_umma: unaligned operand (rows %lld, kbytes %lld, stride %lld, columns %d)", (long long)rows_pad,
                        (long long)kbytes, (long long)stride, nrows);
    umma_pools pool;
    if (plane == SGB_PLANE_VALUE) { pool.ax = 0x02000102u; pool.ay = 0x01020001u; pool.bx = 0x01020202u; pool.by = 0x00000101u; }
    else                          { pool.ax = 0x01000101u; pool.ay = 0x01010001u; pool.bx = 0x01010101u; pool.by = 0x00000101u; }
    const int64_t ksteps = kbytes / 64;
    if (ksteps == 0 || rows_pad == 0 || nrows == 0) return 0;
    // int32 accumulation: |sum| <= 2 * 128 * (genotypes per row)
    if (kbytes * 4 > ((int64_t)1 << 23)) return sgb_fail(h, "k_pk2_umma: %lld genotypes per row exceed the int32 accumulation bound", (long long)(kbytes * 4));
    const int64_t row_tiles = rows_pad / UMMA_ROWS;
    // K split.  All CTAs of a launch do the same work, so the launch takes ceil(CTAs / resident slots) rounds of one CTA time:
    // 1564 sample tiles on 296 slots are 5.3 rounds of work done in 6, 492 marker tiles x 2 chunks 3.3 done in 4.  Pick the
    // number of k-chunks that minimises rounds x (k-steps per chunk + a fixed per-CTA cost of ~16 k-steps: tensor-memory
    // allocation, pipeline fill, accumulator read-out); chunks of one row tile meet in `out` with atomics.
    const int first_pass_cols = nrows > 256 ? ((((nrows + (nrows + 255) / 256 - 1) / ((nrows + 255) / 256)) + 15) & ~15) : nrows;
    const int64_t slots = (int64_t)h->sm_count * (first_pass_cols <= 128 ? 2 : 1);
    int64_t kchunks = 1, best_cost = -1;
    for (int64_t kc = 1; kc <= 24; kc++) {
        int64_t per_c = cdiv64(ksteps, kc);
        if (per_c < 8) per_c = 8;
        if (per_c > ksteps) per_c = ksteps;
        const int64_t kcr = cdiv64(ksteps, per_c);
        const int64_t cost = cdiv64(row_tiles * kcr, slots) * (per_c + 16);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; kchunks = kcr; }
    }
#ifdef SGB_ABLATION
    if (getenv("SGB_UMMA_KCHUNKS")) kchunks = atoi(getenv("SGB_UMMA_KCHUNKS"));
#endif
    int64_t per = cdiv64(ksteps, kchunks);
    if (per < 8) per = 8;
    if (per > ksteps) per = ksteps;
    kchunks = cdiv64(ksteps, per);
    // k-chunks of one row tile meet in `out` with atomics; a single chunk writes every element once (out was zeroed by
    // the previous recombine), unless the caller accumulates several launches (dense-GRM build over marker shards)
    const int use_atomic = (kchunks > 1 || h->umma_accumulate) ? 1 : 0;
    if (sgb_first_on_device(h->device, SGB_SITE_UMMA)) {
        CUDA_OK(h, cudaFuncSetAttribute(pk2_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CUDA_OK(h, cudaFuncSetAttribute(pk2_umma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
#ifdef SGB_ABLATION          // timing experiments only (make ABLATION=1): the product build reads no environment
    static const int force_stages = getenv("SGB_UMMA_STAGES") ? atoi(getenv("SGB_UMMA_STAGES")) : 0;
    static const int dbg = getenv("SGB_UMMA_DBG") ? atoi(getenv("SGB_UMMA_DBG")) : 0;
#else
    const int force_stages = 0, dbg = 0;
#endif
    // Balanced passes of N <= 256 accumulator columns (multiples of 16).  Measured (profiles/r02_sweep_digits*.txt, ncu
    // r02_pk2_umma_k31_d5): with the A operand in tensor memory an M = 128, K = 32 MMA costs about N/2 + 25 cycles, so two
    // N = 112 passes (k = 31 at 7 digits) take 2 x 8.0 ms where ONE N = 224 pass needs ~ 11: wide batches run as few passes as
    // the 512 TMEM columns allow (3 A stages of 64 columns + N <= 256 accumulator columns, one CTA per SM); N <= 128 keeps
    // two CTAs per SM on 256 columns each.
    const int npass = (nrows + 255) / 256;
    const int per_pass = (((nrows + npass - 1) / npass) + 15) & ~15;
    const int ngroups = nrows / 8;
    for (int n0 = 0; n0 < nrows; n0 += per_pass) {
        const int N = nrows - n0 < per_pass ? nrows - n0 : per_pass;
        // three A stages let the producers run a full step ahead of the MMAs
        int stages = (N <= 64 || N > 128) ? 3 : 2;
        if (force_stages == 2 || force_stages == 3) stages = force_stages;
        int tmem_cols = stages * UMMA_A_COLS + N <= 256 ? 256 : 512;
        int nb = (96 * 1024) / (UMMA_KSTEP * N);                // B ring depth (stages of 256 x N bytes)
        if (nb > UMMA_MAX_BSTAGES) nb = UMMA_MAX_BSTAGES;
        if (nb < 2) nb = 2;
        if (tmem_cols == 512 && nb < 3) nb = 3;                  // > 114 KB of shared memory: exactly one CTA (one 512-column allocation) per SM
        // the image interleaves all core-matrix groups per k-block: this pass starts at group n0 / 8
        const int8_t *Lp = L + (int64_t)(n0 / 8) * 1024;
        for (int64_t y0 = 0; y0 < row_tiles; y0 += 65535) {
            int64_t ny = row_tiles - y0 < 65535 ? row_tiles - y0 : 65535;
            dim3 grid((unsigned)kchunks, (unsigned)ny);
            if (stages == 3)
                pk2_umma_kernel<3><<<grid, UMMA_THREADS, (size_t)nb * UMMA_KSTEP * N, h->stream>>>(P + y0 * UMMA_ROWS * stride, stride, ksteps, (int)per, Lp, N,
                                                                           (int64_t)ngroups * 1024, out + y0 * UMMA_ROWS * (int64_t)nrows,
                                                                           nrows, n0, use_atomic, pool, tmem_cols, nb, dbg);
            else
                pk2_umma_kernel<2><<<grid, UMMA_THREADS, (size_t)nb * UMMA_KSTEP * N, h->stream>>>(P + y0 * UMMA_ROWS * stride, stride, ksteps, (int)per, Lp, N,
                                                                           (int64_t)ngroups * 1024, out + y0 * UMMA_ROWS * (int64_t)nrows,
                                                                           nrows, n0, use_atomic, pool, tmem_cols, nb, dbg);
            UMMA_LAUNCH_CHECK(h);
        }
    }
    return 0;
}

// k right-hand sides = nl k accumulator columns (padded to a multiple of 16)
int k_pk2_umma(sgb_ctx *h, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t kbytes, const int8_t *L, int k,
               int32_t *out, int plane, int nl)
{
    return k_pk2_umma_rows(h, P, stride, rows_pad, kbytes, L, umma_npad(k, nl), out, plane);
}
