// DRAM fetch granularity probe: random 32-byte gathers from a 4 GiB array.
//   mode 0: one random 32-B sector per access
//   mode 1: both sectors of a random 64-B block (two accesses)
//   mode 2: all four sectors of a random 128-B line (four accesses)
//   mode 3: sector 0 of a random 64-B block now, its sibling sector ~DELAY accesses later (same thread)
// Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum` and divide by the access count.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
template <int MODE>
__global__ void k(const float4 *a, unsigned long long nsec, int per_thread, float *out)
{
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    float s = 0.f;
    for (int i = 0; i < per_thread; ++i) {
        const unsigned long long h = mix(tid * 1000003ull + i);
        if (MODE == 0) { const unsigned long long sec = h % nsec; float4 x = a[2 * sec], y = a[2 * sec + 1]; s += x.x + y.w; }
        if (MODE == 1) { const unsigned long long blk = h % (nsec / 2); for (int q = 0; q < 2; ++q) { float4 x = a[4 * blk + 2 * q], y = a[4 * blk + 2 * q + 1]; s += x.x + y.w; } }
        if (MODE == 2) { const unsigned long long blk = h % (nsec / 4); for (int q = 0; q < 4; ++q) { float4 x = a[8 * blk + 2 * q], y = a[8 * blk + 2 * q + 1]; s += x.x + y.w; } }
        if (MODE == 3) {
            const unsigned long long blk = h % (nsec / 2);
            float4 x = a[4 * blk], y = a[4 * blk + 1]; s += x.x + y.w;
            if (i >= 8) { const unsigned long long old = mix(tid * 1000003ull + i - 8) % (nsec / 2); float4 u = a[4 * old + 2], v = a[4 * old + 3]; s += u.x + v.w; }
        }
    }
    out[tid] = s;
}
int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    // optional second argument: cudaLimitMaxL2FetchGranularity (32, 64 or 128 bytes; a hint)
    size_t gran = 0;
    cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
    printf("default L2 fetch granularity limit: %zu\n", gran);
    if (argc > 2) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[2]));
        cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
        printf("set to %s -> %s, now %zu\n", argv[2], cudaGetErrorString(e), gran);
    }
    const unsigned long long bytes = 4ull << 30, nsec = bytes / 32;
    float4 *a; float *out;
    cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes);
    const int blocks = 148 * 16, threads = 256, per = 64;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<blocks, threads>>>(a, nsec, per, out);
        if (mode == 1) k<1><<<blocks, threads>>>(a, nsec, per, out);
        if (mode == 2) k<2><<<blocks, threads>>>(a, nsec, per, out);
        if (mode == 3) k<3><<<blocks, threads>>>(a, nsec, per, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double acc = (double)blocks * threads * per * (mode == 1 ? 2 : mode == 2 ? 4 : mode == 3 ? 2 : 1);
    printf("mode %d: %.3f ms, %.1f M 32-B accesses, %.1f G accesses/s, useful %.0f GB/s\n", mode, ms, acc / 1e6, acc / ms / 1e6, acc * 32 / ms / 1e6);
    return 0;
}
