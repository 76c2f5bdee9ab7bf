// FFMA2/FMUL2/FADD2 issue rate with three distinct 64-bit register operands (no broadcast), and a
// mix shaped like the packed Moller-Trumbore body.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(float2 *out, const float2 *in, int iters)
{
    float2 a[8], b[4], c[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = in[threadIdx.x + 300 + 32 * i]; c[i] = in[threadIdx.x + 600 + 32 * i]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = __ffma2_rn(a[i], b[(i + r) & 3], c[i & 3]);         // 3 distinct pairs
                if (MODE == 1) a[i] = __fmul2_rn(a[i], b[(i + r) & 3]);
                if (MODE == 2) a[i] = __fadd2_rn(a[i], b[(i + r) & 3]);
                if (MODE == 3) a[i] = __ffma2_rn(a[i], make_float2(b[r].x, b[r].x), c[i & 3]);   // one broadcast
                if (MODE == 4) { a[i].x = __fmaf_rn(a[i].x, b[(i + r) & 3].x, c[i & 3].x); a[i].y = __fmaf_rn(a[i].y, b[(i + r) & 3].y, c[i & 3].y); }
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
static void run(const char *name, int nsm)
{
    float2 *d, *in;
    const int blocks = nsm * 8, threads = 256;
    cudaMalloc(&d, sizeof(float2) * blocks * threads);
    cudaMalloc(&in, sizeof(float2) * 4096);
    cudaMemset(in, 0, sizeof(float2) * 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, in, 16);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(d, in, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads / 32 * ITERS * 32;   // warp-level packed instructions
    const double cyc = best * 1e-3 * 1.965e9 * nsm * 4;              // SMSP cycles
    printf("%-34s %8.3f ms  %.2f cycles per warp instruction pair-op per SMSP\n", name, best, cyc / ops);
    cudaFree(d); cudaFree(in);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    run<0>("FFMA2 3 distinct pairs", p.multiProcessorCount);
    run<1>("FMUL2", p.multiProcessorCount);
    run<2>("FADD2", p.multiProcessorCount);
    run<3>("FFMA2 one broadcast operand", p.multiProcessorCount);
    run<4>("2 x scalar FFMA (same work)", p.multiProcessorCount);
    return 0;
}
