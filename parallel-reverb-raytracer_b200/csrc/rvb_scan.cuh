// Device-wide exclusive prefix sum of uint32 counts (hand-written; used by the microphone
// visibility map, the radix sort and the stream compactions).  Three launches, no inter-CTA spinning:
//   scan_reduce_kernel   one partial sum per block of RVB_SCAN_TILE items
//   scan_sums_kernel     ONE block turns the partial sums into exclusive block offsets (serial over
//                        chunks of 1024 with a carry: fine up to a few million blocks)
//   scan_apply_kernel    block-local exclusive scan + block offset
// out may alias in.  If total is not null it receives the sum of all items (it may point at out[n]).
#pragma once
#include "rvb_internal.cuh"

namespace rvb {

#define RVB_SCAN_THREADS 1024
#define RVB_SCAN_ITEMS 4
#define RVB_SCAN_TILE (RVB_SCAN_THREADS * RVB_SCAN_ITEMS)

// exclusive scan of one value per thread across a 1024-thread block; returns the block total in `total`
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *warp_sums /* [32] shared */, uint32_t &total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        warp_sums[lane] = winc - w;   // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    total = warp_sums[32];
    const uint32_t r = warp_sums[warp] + inc - v;
    __syncthreads();   // warp_sums may be reused by the caller
    return r;
}

static __global__ void __launch_bounds__(RVB_SCAN_THREADS) scan_reduce_kernel(const uint32_t *in, size_t n, uint32_t *block_sums)
{
    __shared__ uint32_t ws[33];
    const size_t base = (size_t)blockIdx.x * RVB_SCAN_TILE + (size_t)threadIdx.x * RVB_SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < RVB_SCAN_ITEMS; ++k) if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan_1024(s, ws, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(RVB_SCAN_THREADS) scan_sums_kernel(uint32_t *block_sums, uint32_t nb, uint32_t *total_out)
{
    __shared__ uint32_t ws[33];
    uint32_t carry = 0;
    for (uint32_t c0 = 0; c0 < nb; c0 += RVB_SCAN_THREADS) {
        const uint32_t i = c0 + threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan_1024(v, ws, total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

static __global__ void __launch_bounds__(RVB_SCAN_THREADS) scan_apply_kernel(const uint32_t *in, size_t n, const uint32_t *block_sums, uint32_t *out)
{
    __shared__ uint32_t ws[33];
    const size_t base = (size_t)blockIdx.x * RVB_SCAN_TILE + (size_t)threadIdx.x * RVB_SCAN_ITEMS;
    uint32_t v[RVB_SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < RVB_SCAN_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
    uint32_t total;
    uint32_t ex = block_exclusive_scan_1024(s, ws, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < RVB_SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

// d_block_sums: scratch of scan_blocks(n) uint32
static inline size_t scan_blocks(size_t n) { return (n + RVB_SCAN_TILE - 1) / RVB_SCAN_TILE; }
static inline int exclusive_scan_u32(const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_block_sums, uint32_t *d_total, cudaStream_t st)
{
    if (n == 0) {
        if (d_total) RVB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), st));
        return RVB_OK;
    }
    const unsigned nb = (unsigned)scan_blocks(n);
    note_launch(), scan_reduce_kernel<<<nb, RVB_SCAN_THREADS, 0, st>>>(d_in, n, d_block_sums);
    note_launch(), scan_sums_kernel<<<1, RVB_SCAN_THREADS, 0, st>>>(d_block_sums, nb, d_total);
    note_launch(), scan_apply_kernel<<<nb, RVB_SCAN_THREADS, 0, st>>>(d_in, n, d_block_sums, d_out);
    RVB_CUDA(cudaGetLastError());
    return RVB_OK;
}

// Stable stream compaction of indices: d_out[k] = the k-th i in [0, n) with pred(i), *d_count = how many.
template <class Pred>
static __global__ void compact_flag_kernel(const Pred pred, uint32_t n, uint32_t *flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = pred(i) ? 1u : 0u;
}
static __global__ void compact_scatter_kernel(const uint32_t *flags, const uint32_t *pos, uint32_t n, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = i;
}
template <class Pred>
static inline int compact_indices(const Pred &pred, uint32_t n, uint32_t *d_out, uint32_t *d_count, cudaStream_t st)
{
    if (n == 0) return exclusive_scan_u32(nullptr, nullptr, 0, nullptr, d_count, st);
    uint32_t *d_flags = nullptr, *d_pos = nullptr, *d_sums = nullptr;
    RVB_TRY(dev_alloc(&d_flags, (size_t)n)); RVB_TRY(dev_alloc(&d_pos, (size_t)n)); RVB_TRY(dev_alloc(&d_sums, scan_blocks(n)));
    const unsigned grd = (n + 255u) / 256u;
    note_launch(), compact_flag_kernel<<<grd, 256, 0, st>>>(pred, n, d_flags);
    int rc = exclusive_scan_u32(d_flags, d_pos, n, d_sums, d_count, st);
    if (rc == RVB_OK) {
        note_launch(), compact_scatter_kernel<<<grd, 256, 0, st>>>(d_flags, d_pos, n, d_out);
        if (cudaGetLastError() != cudaSuccess) rc = fail(RVB_ERR_CUDA, "compaction launch failed");
    }
    dev_free(d_flags); dev_free(d_pos); dev_free(d_sums);
    return rc;
}

}  // namespace rvb
