// Pipe-rate microbenchmarks used to size the tensor-engine design (DESIGN.md "why int8 mma.sync"):
//   legacy mma.sync s8 (m16n8k32) and bf16 (m16n8k16) issue rates, fp64 FMA rate, and a streaming-read baseline.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void imma_kernel(int iters, int *out)
{
    int c[8][4] = {};
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 0x7fffffff) out[0] = s;
}

__global__ void hmma_kernel(int iters, float *out)
{
    float c[8][4] = {};
    unsigned a0 = 0x3f803f80, a1 = a0, a2 = a0, a3 = a0, b0 = a0, b1 = a0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 12345.f) out[0] = s;
}

__global__ void dfma_kernel(int iters, double *out)
{
    double a[8];
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 1e-3 + j;
    double x = 1.0000001, y = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fma(a[j], x, y);
    }
    double s = 0;
    for (int j = 0; j < 8; j++) s += a[j];
    if (s == 12345.0) out[0] = s;
}

__global__ void prmt_kernel(int iters, unsigned *out)
{
    unsigned a[8];
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 2654435761u + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = __byte_perm(0x03020100u, a[j], a[j] & 0x3333u) + (a[j] >> 2);
    }
    unsigned s = 0;
    for (int j = 0; j < 8; j++) s += a[j];
    if (s == 12345u) out[0] = s;
}

__global__ void stream_kernel(const uint4 *p, size_t n, unsigned *out)
{
    unsigned s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        s += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (s == 0x12345u) out[0] = s;
}

template <typename F> static float time_ms(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sms);
    void *out; cudaMalloc(&out, 64);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        dim3 g(sms * 2), b(warps * 16);      // 2 CTAs per SM, `warps` warps per SM in total
        float ms = time_ms([&] { imma_kernel<<<g, b>>>(iters, (int *)out); });
        double mmas = (double)g.x * (b.x / 32) * iters * 8;
        printf("imma m16n8k32 u8s8 : %2d warps/SM  %.1f MMA/clk/SM-equivalent @1.9GHz  %.1f TOPS\n", warps,
               mmas / (ms * 1e-3) / sms / 1.9e9, mmas * 16 * 8 * 32 * 2 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { hmma_kernel<<<g, b>>>(iters, (float *)out); });
        printf("hmma m16n8k16 bf16 : %2d warps/SM  %.1f TFLOPS\n", warps, mmas * 16 * 8 * 16 * 2 / (ms * 1e-3) / 1e12);
    }
    {
        dim3 g(sms * 4), b(256);
        float ms = time_ms([&] { dfma_kernel<<<g, b>>>(iters, (double *)out); });
        printf("dfma : %.2f TFLOPS fp64\n", (double)g.x * b.x * iters * 8 * 2 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { prmt_kernel<<<g, b>>>(iters, (unsigned *)out); });
        printf("prmt+lop+shf+add chain : %.2f T lane-iterations/s\n", (double)g.x * b.x * iters * 8 / (ms * 1e-3) / 1e12);
    }
    {
        size_t bytes = (size_t)8 << 30;
        void *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
        for (int mult = 4; mult <= 32; mult *= 2) {
            float ms = time_ms([&] { stream_kernel<<<sms * mult, 256>>>((const uint4 *)buf, bytes / 16, (unsigned *)out); });
            printf("stream read 8 GiB, %d CTAs/SM : %.0f GB/s\n", mult, bytes / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
