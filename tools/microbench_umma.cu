// tcgen05.mma issue-rate microbenchmark on B200: kind::i8 vs kind::f8f6f4 (e4m3) vs kind::f16 (bf16), A in TMEM,
// B through a shared-memory descriptor, M = 128.  Decides which operand type the limb GEMM should use.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int KIND>   // 0 = i8, 1 = f8f6f4 e4m3, 2 = f16 bf16
__global__ void __launch_bounds__(128) umma_rate(int iters, int N, unsigned *out, int vary)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 128 * 256 / 4; i += 128) ((uint32_t *)smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t tb = slot;
    if (threadIdx.x == 0) {
        uint32_t afmt = KIND == 0 ? 0u : (KIND == 1 ? 0u : 1u), bfmt = KIND == 0 ? 1u : (KIND == 1 ? 0u : 1u), cfmt = KIND == 0 ? 2u : 1u;
        uint32_t idesc = (cfmt << 4) | (afmt << 7) | (bfmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        uint64_t desc = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((uint64_t)8 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46);
        const uint64_t desc0 = desc;
        for (int i = 0; i < iters; i++) {
            uint32_t acc = i > 0;
            if (vary) desc = desc0 + (uint64_t)(((i & 7) * 4096) >> 4);     // a different 4 KB B tile every MMA
            if (KIND == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(tb + 128 + ((vary & 2) ? (i & 3) * 32 : 0)), "r"(tb + (i & 7) * 8), "l"(desc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
            else if (KIND == 1)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(tb + 128 + ((vary & 2) ? (i & 3) * 32 : 0)), "r"(tb + (i & 7) * 8), "l"(desc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(tb + 128 + ((vary & 2) ? (i & 3) * 32 : 0)), "r"(tb + (i & 7) * 8), "l"(desc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)));
        if (out && iters < 0) out[0] = tb;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256));
}
template <int KIND> void run(const char *name, int N, int kper, int sms, int vary, int ctas_per_sm)
{
    cudaFuncSetAttribute(umma_rate<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    umma_rate<KIND><<<sms * ctas_per_sm, 128, 64 * 1024>>>(100, N, nullptr, vary); cudaDeviceSynchronize();
    cudaEventRecord(e0); umma_rate<KIND><<<sms * ctas_per_sm, 128, 64 * 1024>>>(iters, N, nullptr, vary); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    double macs = (double)sms * ctas_per_sm * iters * 128.0 * N * kper;
    printf("%-22s vary=%d ctas/sm=%d N=%3d : %8.3f ms  %7.1f cyc/MMA/CTA @1.9GHz  %8.1f T MAC-ops/s (x2 = TOPS)  %s\n", name, vary, ctas_per_sm, N, ms, ms * 1e-3 * 1.9e9 / iters,
           2 * macs / (ms * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int N : {16, 32, 64, 128, 256}) {
        run<0>("kind::i8 (u8 x s8)", N, 32, sms, 0, 1);
        run<1>("kind::f8f6f4 (e4m3)", N, 32, sms, 0, 1);
        run<2>("kind::f16 (bf16)", N, 16, sms, 0, 1);
    }
    for (int N : {16, 32}) {
        run<0>("i8, 1 accumulator", N, 32, sms, 1, 1);
        run<0>("i8, 4 accumulators", N, 32, sms, 3, 1);
        run<0>("i8, 4 accumulators", N, 32, sms, 3, 2);
    }
    return 0;
}
