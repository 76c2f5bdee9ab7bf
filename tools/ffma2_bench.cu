// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and
// mixed with ALU-pipe work (FSETP/SEL) -- decides whether the pairwise (two triangles per
// instruction) form of the intersection loop can beat the scalar one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int iters)
{
    float2 acc[NACC];
    float cnt = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                if (MODE == 0) {            // scalar: 2 FFMA per pair
                    acc[i].x = __fmaf_rn(acc[i].x, a, b);
                    acc[i].y = __fmaf_rn(acc[i].y, a, b);
                } else {                    // packed: 1 FFMA2 per pair
                    acc[i] = __ffma2_rn(acc[i], A, B);
                }
                if (MODE == 2 || MODE == 3) {   // + one compare/select per pair (ALU pipe)
                    cnt = (acc[i].x > cnt) ? acc[i].y : cnt;
                }
                if (MODE == 3) {
                    cnt = (acc[i].y < cnt) ? acc[i].x : cnt;
                }
            }
        }
    }
    float s = cnt;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int nsm)
{
    float *d;
    const int blocks = nsm * 8, threads = 256;
    cudaMalloc(&d, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 0.999f, 0.001f, 16);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(d, 0.999f, 0.001f, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fma = (double)blocks * threads * ITERS * 4 * NACC * 2;   // scalar-equivalent FMAs
    printf("%-28s %8.3f ms  %7.2f TFLOP/s (fp32 fma*2)\n", name, best, 2.0 * fma / best * 1e-9);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<0>("scalar FFMA", p.multiProcessorCount);
    run<1>("packed FFMA2", p.multiProcessorCount);
    run<2>("packed FFMA2 + 1 cmp/sel", p.multiProcessorCount);
    run<3>("packed FFMA2 + 2 cmp/sel", p.multiProcessorCount);
    return 0;
}
