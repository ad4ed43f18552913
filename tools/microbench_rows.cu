// Access-pattern microbenchmark: every thread streams ITS OWN ROW of a row-major byte matrix (the tcgen05 producer
// pattern: TMEM lane = row), CH bytes per step, PF steps prefetched.  Compare with the coalesced stream of microbench.cu.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int CH, int PF>
__global__ void rowstream(const uint8_t *P, int64_t stride, int64_t nsteps, unsigned *out)
{
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint8_t *p = P + row * stride;
    uint32_t acc = 0;
    uint32_t buf[PF][CH / 4];
#pragma unroll
    for (int i = 0; i < PF; i++)
#pragma unroll
        for (int c = 0; c < CH / 32; c++)
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(buf[i][8*c]), "=r"(buf[i][8*c+1]), "=r"(buf[i][8*c+2]), "=r"(buf[i][8*c+3]), "=r"(buf[i][8*c+4]), "=r"(buf[i][8*c+5]), "=r"(buf[i][8*c+6]), "=r"(buf[i][8*c+7]) : "l"(p + (int64_t)i * CH + 32 * c));
    for (int64_t s0 = 0; s0 < nsteps; s0 += PF) {
#pragma unroll
        for (int j = 0; j < PF; j++) {
            int64_t s = s0 + j;
            if (s < nsteps) {
#pragma unroll
                for (int c = 0; c < CH / 4; c++) acc ^= buf[j][c];
                if (s + PF < nsteps)
#pragma unroll
                    for (int c = 0; c < CH / 32; c++)
                        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(buf[j][8*c]), "=r"(buf[j][8*c+1]), "=r"(buf[j][8*c+2]), "=r"(buf[j][8*c+3]), "=r"(buf[j][8*c+4]), "=r"(buf[j][8*c+5]), "=r"(buf[j][8*c+6]), "=r"(buf[j][8*c+7]) : "l"(p + (s + PF) * CH + 32 * c));
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int CH, int PF> void run(const uint8_t *buf, int64_t rows, int64_t stride, unsigned *out, int bs)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int64_t nsteps = stride / CH;
    rowstream<CH, PF><<<rows / bs, bs>>>(buf, stride, nsteps, out); cudaDeviceSynchronize();
    cudaEventRecord(e0); rowstream<CH, PF><<<rows / bs, bs>>>(buf, stride, nsteps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("row-per-thread stream: %3d B/step, prefetch %d, block %3d : %.3f ms  %.0f GB/s\n", CH, PF, bs, ms, rows * stride / (ms * 1e-3) / 1e9);
}
int main()
{
    int64_t rows = 500224, stride = 50048;      // the marker-major copy at 200k samples
    uint8_t *buf; unsigned *out; cudaMalloc(&buf, rows * stride); cudaMemset(buf, 1, rows * stride); cudaMalloc(&out, 64);
    run<32, 3>(buf, rows, stride, out, 128); run<32, 3>(buf, rows, stride, out, 256); run<32, 6>(buf, rows, stride, out, 128);
    run<64, 2>(buf, rows, stride, out, 128); run<64, 3>(buf, rows, stride, out, 128); run<64, 4>(buf, rows, stride, out, 256);
    run<128, 2>(buf, rows, stride, out, 128); run<128, 3>(buf, rows, stride, out, 256);
    return 0;
}
