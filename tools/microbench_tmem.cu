// TMEM store bandwidth: how fast can register data enter tensor memory with tcgen05.st (the A-operand path of
// pk2_umma_kernel)?  Compared with st.shared.v4 of the same bytes.   nvcc -gencode arch=compute_100a,code=sm_100a
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
// warps: 4 or 8 (two warps per lane quarter write different columns)
__global__ void tmem_store(int iters, int cols, unsigned *out)
{
    __shared__ uint32_t slot;
    int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[32];
    for (int i = 0; i < 32; i++) r[i] = threadIdx.x * 33 + i;
    for (int it = 0; it < iters; it++) {
        tmem_st_x32(base, r);
        tmem_st_x32(base + 32, r);
        asm volatile("tcgen05.wait::st.sync.aligned;");
        r[it & 31] ^= it;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (r[3] == 0x12345) out[0] = r[3];
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(cols));
}
__global__ void smem_store(int iters, unsigned *out)
{
    extern __shared__ uint4 buf[];
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) buf[j * blockDim.x + threadIdx.x] = v;      // 256 B per thread per iteration
        v.x ^= it;
        __syncwarp();
    }
    if (buf[threadIdx.x].x == 0x12345) out[0] = 1;
}
int main()
{
    unsigned *out; cudaMalloc(&out, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    double clk = p.clockRate * 1e3;
    int iters = 20000;
    for (int ctas = 1; ctas <= 2; ctas++)
        for (int warps = 4; warps <= 8; warps += 4) {
            int cols = ctas == 1 ? 512 : 256;
            tmem_store<<<p.multiProcessorCount * ctas, warps * 32>>>(100, cols, out); cudaDeviceSynchronize();
            cudaEventRecord(e0); tmem_store<<<p.multiProcessorCount * ctas, warps * 32>>>(iters, cols, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double bytes_sm = (double)ctas * warps * 32 * 256.0 * iters;
            printf("tcgen05.st.32x32b.x32: %d CTA/SM x %d warps: %.3f ms  %.1f B/clk/SM (at %.0f MHz nominal)  %s\n", ctas, warps, ms,
                   bytes_sm / (ms * 1e-3) / clk, clk / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (int warps = 4; warps <= 16; warps *= 2) {
        cudaFuncSetAttribute(smem_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        size_t sm = (size_t)warps * 32 * 256;
        smem_store<<<p.multiProcessorCount * 2, warps * 32, sm>>>(100, out); cudaDeviceSynchronize();
        cudaEventRecord(e0); smem_store<<<p.multiProcessorCount * 2, warps * 32, sm>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("st.shared.v4: 2 CTA/SM x %d warps: %.3f ms  %.1f B/clk/SM  %s\n", warps, ms, 2.0 * warps * 32 * 256.0 * iters / (ms * 1e-3) / clk,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
