// Does concurrent TMEM traffic (tcgen05.st of the next A stage by the producer warps) slow the MMA chain down?
// One MMA thread issues M=128, N, K=32 kind::i8 MMAs (A in TMEM columns [64,128), D at [192,..)) while `stw` warps
// keep storing 2 x 32 columns into TMEM columns [0,64) (mode 1) or into the SAME columns the MMAs read (mode 2).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__global__ void __launch_bounds__(288) umma_st(int iters, int N, int mode, int st_per_8mma, unsigned *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 128 * 256 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t tb = slot;
    if (warp == 8) {
        if (threadIdx.x == 256) {
            uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
            uint64_t desc0 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((uint64_t)8 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46);
            for (int i = 0; i < iters; i++) {
                uint32_t acc = i > 0;
                uint64_t desc = desc0 + (uint64_t)(((i & 7) * 4096) >> 4);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(tb + 192),
                             "r"(tb + 64 + (i & 7) * 8), "l"(desc), "r"(idesc), "r"(acc), "r"(0u)
                             : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)));
            stop = 1;
        }
    } else if (mode > 0 && warp < 4 * (mode >= 3 ? 2 : 1)) {
        uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + ((mode == 2) ? 64 : 0);
        uint32_t r[32];
        for (int i = 0; i < 32; i++) r[i] = threadIdx.x + i;
        // pace: st_per_8mma store pairs per 8 MMAs is what the real kernel does (1); here free running unless paced by clock
        long long t0 = clock64();
        int n = 0;
        while (!stop) {
            tmem_st_x32(base, r);
            tmem_st_x32(base + 32, r);
            asm volatile("tcgen05.wait::st.sync.aligned;");
            n++;
            if (st_per_8mma > 0) { while (clock64() - t0 < (long long)n * st_per_8mma && !stop) { } }
        }
        if (out && n == -1) out[0] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256));
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    cudaFuncSetAttribute(umma_st, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int N : {32, 128})
        for (int ctas = 1; ctas <= 2; ctas++)
            for (int mode = 0; mode <= 3; mode++)
                for (int pace : {0, 1000}) {
                    if (mode == 0 && pace) continue;
                    umma_st<<<sms * ctas, 288, 64 * 1024>>>(100, N, mode, pace, nullptr); cudaDeviceSynchronize();
                    cudaEventRecord(e0); umma_st<<<sms * ctas, 288, 64 * 1024>>>(iters, N, mode, pace, nullptr); cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    printf("N=%3d ctas/sm=%d stores: %-34s pace=%4d clk/pair : %7.1f cyc/MMA/CTA @1.9GHz  %s\n", N, ctas,
                           mode == 0 ? "none" : mode == 1 ? "4 warps -> other TMEM columns" : mode == 2 ? "4 warps -> the columns MMAs read" : "8 warps -> other TMEM columns",
                           pace, ms * 1e-3 * 1.9e9 / iters, cudaGetErrorString(cudaGetLastError()));
                }
    return 0;
}
