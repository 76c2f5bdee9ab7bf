// Correctness and timing of the library's radix sort (csrc/rvb_sort.cuh) against std::stable_sort and
// against cub::DeviceRadixSort.  Links librayverb_b200.so for the allocator.
//   tools/build_microbench.sh builds it; run: tools/sort_test.bin [quick]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include <cub/cub.cuh>

#include "../parallel-reverb-raytracer_b200/csrc/rvb_sort.cuh"

using namespace rvb;

#ifndef SORT_THREADS
#define SORT_THREADS 256
#endif

static inline uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// first-pass functor of the tests: keys from an array, element i invalid when its key is 0xffffffff
template <class KeyT>
struct TestIn {
    const uint32_t *raw;
    uint32_t shift;
    typedef uint32_t Raw;
    __device__ __forceinline__ Raw fetch(uint32_t i) const { return raw[i]; }
    __device__ __forceinline__ bool decode(uint32_t i, const Raw &s, KeyT &k, uint32_t &v) const
    {
        if (s == 0xffffffffu) return false;
        k = (KeyT)(s >> shift);
        v = i | ((s & ((1u << shift) - 1u)) << 30);
        return true;
    }
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <class KeyT>
static bool check_case(uint32_t n, int key_bits, uint32_t shift, double invalid_frac, int pattern, bool verbose)
{
    std::vector<uint32_t> raw(n);
    const uint32_t L = (uint32_t)((1ull << (key_bits + shift)) - (key_bits + shift > 3 ? 3 : 0));   // not a power of two
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t h = hash32(i * 2654435761u + 12345u + pattern);
        uint32_t s;
        if (pattern == 0) s = h % L;                                   // uniform
        else if (pattern == 1) s = (uint32_t)(((unsigned long long)i * L) / std::max(n, 1u));   // already sorted
        else if (pattern == 2) s = (L - 1) - (uint32_t)(((unsigned long long)i * L) / std::max(n, 1u));   // reversed
        else if (pattern == 3) s = (h % 7 == 0) ? h % L : (L / 2);     // one hot key
        else s = (i / 977 * 31 + h % 64) % L;                          // locally coherent
        if ((hash32(h) % 1000000) < invalid_frac * 1000000) s = 0xffffffffu;
        raw[i] = s;
    }
    const uint32_t nseg = ((L - 1) >> shift) + 1;
    // reference
    std::vector<uint32_t> ord;
    ord.reserve(n);
    for (uint32_t i = 0; i < n; ++i) if (raw[i] != 0xffffffffu) ord.push_back(i);
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return (raw[a] >> shift) < (raw[b] >> shift); });
    std::vector<uint32_t> ref_begin(nseg, 0xffffffffu), ref_end(nseg, 0u);
    for (uint32_t p = 0; p < ord.size(); ++p) {
        const uint32_t k = raw[ord[p]] >> shift;
        ref_begin[k] = std::min(ref_begin[k], p);
        ref_end[k] = std::max(ref_end[k], p + 1);
    }

    cudaStream_t st = lib_stream();
    uint32_t *d_raw = nullptr, *d_vals[2] = {nullptr, nullptr}, *d_begin = nullptr, *d_end = nullptr, *d_nvalid = nullptr;
    KeyT *d_keys[2] = {nullptr, nullptr};
    dev_alloc(&d_raw, std::max(n, 1u)); dev_alloc(&d_vals[0], std::max(n, 1u)); dev_alloc(&d_vals[1], std::max(n, 1u));
    dev_alloc(&d_keys[0], std::max(n, 1u)); dev_alloc(&d_keys[1], std::max(n, 1u));
    dev_alloc(&d_begin, nseg); dev_alloc(&d_end, nseg); dev_alloc(&d_nvalid, 1);
    CK(cudaMemcpyAsync(d_raw, raw.data(), n * 4ull, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_begin, 0xff, nseg * 4ull, st));
    CK(cudaMemsetAsync(d_end, 0, nseg * 4ull, st));
    rsort::PassPlan pl = rsort::plan_passes(0, key_bits);
    if (sizeof(KeyT) == 2 && pl.bits != 8) { printf("skip (u16 keys need 8-bit digits)\n"); return true; }
    TestIn<KeyT> in{d_raw, shift};
    int cur = -1;
    bool ok = true;
    for (int drop = 0; drop < 2 && ok; ++drop) {
        CK(cudaMemsetAsync(d_begin, 0xff, nseg * 4ull, st));
        CK(cudaMemsetAsync(d_end, 0, nseg * 4ull, st));
        int rc = n ? rsort::sort_pairs<SORT_THREADS, KeyT>(in, rsort::NoSink{}, in, n, pl, d_keys, d_vals, drop != 0, d_begin, d_end, &cur, d_nvalid, st) : 0;
        if (rc != RVB_OK) { printf("sort_pairs failed rc=%d\n", rc); return false; }
        CK(cudaStreamSynchronize(st));
        std::vector<uint32_t> vals(ord.size()), begin(nseg), end(nseg);
        std::vector<KeyT> keys(ord.size());
        uint32_t nvalid = 0;
        if (n) {
            CK(cudaMemcpy(&nvalid, d_nvalid, 4, cudaMemcpyDeviceToHost));
            if (nvalid != ord.size()) { printf("  nvalid %u != %zu\n", nvalid, ord.size()); ok = false; }
            CK(cudaMemcpy(vals.data(), d_vals[cur], ord.size() * 4ull, cudaMemcpyDeviceToHost));
            if (!drop) CK(cudaMemcpy(keys.data(), d_keys[cur], ord.size() * sizeof(KeyT), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(begin.data(), d_begin, nseg * 4ull, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(end.data(), d_end, nseg * 4ull, cudaMemcpyDeviceToHost));
        }
        size_t bad = 0;
        for (size_t p = 0; p < ord.size() && n; ++p) {
            const uint32_t s = raw[ord[p]];
            const uint32_t want_v = ord[p] | ((s & ((1u << shift) - 1u)) << 30);
            if (vals[p] != want_v || (!drop && keys[p] != (KeyT)(s >> shift))) { if (bad < 5) printf("  mismatch at %zu: val %08x want %08x\n", p, vals[p], want_v); ++bad; }
        }
        for (uint32_t k = 0; k < nseg && n && drop; ++k) {   // run bounds come with the key-dropping last pass
            const bool empty = ref_begin[k] >= ref_end[k], gempty = begin[k] >= end[k];
            if (empty != gempty || (!empty && (begin[k] != ref_begin[k] || end[k] != ref_end[k]))) { if (bad < 5) printf("  run %u: [%u,%u) want [%u,%u)\n", k, begin[k], end[k], ref_begin[k], ref_end[k]); ++bad; }
        }
        if (bad) ok = false;
        if (verbose || bad) printf("n=%u bits=%d shift=%u key%zu invalid=%.2f pattern=%d drop=%d passes=%dx%d valid=%zu: %s (%zu bad)\n", n, key_bits, shift, 8 * sizeof(KeyT),
                                   invalid_frac, pattern, drop, pl.npasses, pl.bits, ord.size(), bad ? "FAIL" : "ok", bad);
    }
    dev_free(d_raw); dev_free(d_vals[0]); dev_free(d_vals[1]); dev_free(d_keys[0]); dev_free(d_keys[1]); dev_free(d_begin); dev_free(d_end); dev_free(d_nvalid);
    return ok;
}

// impulse-like keys: slot order as in the trace (bounce-group-major), time grows with the bounce
__global__ void make_impulse_keys(uint32_t *raw, uint32_t n, uint32_t nrays, uint32_t L, uint32_t nbounce)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t g = i / (4u * nrays), r = (i / 4u) % nrays, b = 4u * g + (i & 3u);
        uint32_t h = (r * 2654435761u) ^ (b * 40503u); h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15;
        const float t = ((float)b + 1.0f + 3.0f * (float)(h & 0xffff) / 65536.0f) / (float)(nbounce + 4);
        uint32_t s = (uint32_t)(t * (float)L);
        raw[i] = s < L ? s : L - 1;
    }
}

template <class KeyT>
__global__ void make_cub_keys(const uint32_t *raw, uint32_t n, uint32_t shift, KeyT *k, uint32_t *v)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { k[i] = (KeyT)(raw[i] >> shift); v[i] = i; }
}

template <class KeyT>
static void bench_case(const char *name, uint32_t n, uint32_t nrays, uint32_t L, uint32_t shift, uint32_t nbounce)
{
    cudaStream_t st = lib_stream();
    const uint32_t nseg = ((L - 1) >> shift) + 1;
    int key_bits = 1;
    while ((1ull << key_bits) < nseg) ++key_bits;
    uint32_t *d_raw = nullptr, *d_vals[2] = {nullptr, nullptr}, *d_begin = nullptr, *d_end = nullptr;
    KeyT *d_keys[2] = {nullptr, nullptr};
    dev_alloc(&d_raw, n); dev_alloc(&d_vals[0], n); dev_alloc(&d_vals[1], n); dev_alloc(&d_keys[0], n); dev_alloc(&d_keys[1], n);
    dev_alloc(&d_begin, nseg); dev_alloc(&d_end, nseg);
    make_impulse_keys<<<148 * 8, 256, 0, st>>>(d_raw, n, nrays, L, nbounce);
    rsort::PassPlan pl = rsort::plan_passes(0, key_bits);
    TestIn<KeyT> in{d_raw, shift};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    int cur = 0;
    for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0, st));
        CK(cudaMemsetAsync(d_begin, 0xff, nseg * 4ull, st));
        CK(cudaMemsetAsync(d_end, 0, nseg * 4ull, st));
        int rc = rsort::sort_pairs<SORT_THREADS, KeyT>(in, rsort::NoSink{}, in, n, pl, d_keys, d_vals, true, d_begin, d_end, &cur, nullptr, st);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        if (rc != RVB_OK) { printf("sort failed\n"); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it) best = std::min(best, ms);
    }
    // CUB on the same keys (key generation kernel + sort, as the library did before)
    float best_cub = 1e9f, best_cub_sort = 1e9f;
    {
        size_t bytes = 0;
        cub::DoubleBuffer<KeyT> kb(d_keys[0], d_keys[1]);
        cub::DoubleBuffer<uint32_t> vb(d_vals[0], d_vals[1]);
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, (int)n, 0, key_bits, st);
        uint8_t *tmp = nullptr;
        dev_alloc(&tmp, bytes);
        cudaEvent_t em; cudaEventCreate(&em);
        for (int it = 0; it < 4; ++it) {
            cub::DoubleBuffer<KeyT> kb2(d_keys[0], d_keys[1]);
            cub::DoubleBuffer<uint32_t> vb2(d_vals[0], d_vals[1]);
            CK(cudaEventRecord(e0, st));
            make_cub_keys<KeyT><<<148 * 8, 256, 0, st>>>(d_raw, n, shift, d_keys[0], d_vals[0]);
            CK(cudaEventRecord(em, st));
            cub::DeviceRadixSort::SortPairs(tmp, bytes, kb2, vb2, (int)n, 0, key_bits, st);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms, ms2; cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, em, e1);
            if (it) { best_cub = std::min(best_cub, ms); best_cub_sort = std::min(best_cub_sort, ms2); }
        }
        dev_free(tmp);
    }
    printf("%-28s n=%u key_bits=%d (%d x %d-bit, key%zu): ours %.3f ms | cub keys+sort %.3f ms (sort alone %.3f; bounds kernel not included)\n", name, n, key_bits,
           pl.npasses, pl.bits, 8 * sizeof(KeyT), best, best_cub, best_cub_sort);
    dev_free(d_raw); dev_free(d_vals[0]); dev_free(d_vals[1]); dev_free(d_keys[0]); dev_free(d_keys[1]); dev_free(d_begin); dev_free(d_end);
}

int main(int argc, char **argv)
{
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    const bool bench_only = argc > 1 && !strcmp(argv[1], "bench");
    if (ensure_device() != RVB_OK) { printf("no device\n"); return 2; }
    bool ok = true;
    if (bench_only) {
        bench_case<uint16_t>("C2-like (79412 bins, shift 1)", 1u << 27, 1u << 20, 79412u, 1, 128);
        bench_case<uint32_t>("C3-like (154799 bins)", 1u << 27, 1u << 20, 154799u, 0, 128);
        return 0;
    }
    const uint32_t sizes[] = {0u, 1u, 31u, 32u, 33u, 4095u, 4096u, 4097u, 100000u, 1u << 20, 5000011u};
    for (uint32_t n : sizes) {
        for (int pattern = 0; pattern < 5; ++pattern) {
            ok &= check_case<uint32_t>(n, 18, 0, 0.0, pattern, false);
            ok &= check_case<uint16_t>(n, 15, 2, 0.1, pattern, false);
        }
        ok &= check_case<uint32_t>(n, 8, 0, 0.5, 0, true);     // one pass of 8
        ok &= check_case<uint32_t>(n, 10, 1, 0.0, 0, true);    // one pass of 10
        ok &= check_case<uint32_t>(n, 20, 0, 0.02, 4, true);   // two passes of 10
        ok &= check_case<uint32_t>(n, 24, 0, 0.02, 0, true);   // three passes of 8
        ok &= check_case<uint32_t>(n, 27, 0, 0.9, 3, true);    // three passes of 9
        ok &= check_case<uint16_t>(n, 16, 1, 0.3, 3, true);
        ok &= check_case<uint16_t>(n, 5, 0, 0.0, 0, true);
    }
    printf("correctness: %s\n", ok ? "ALL OK" : "FAILURES");
    if (!quick) {
        bench_case<uint16_t>("C2-like (79412 bins, shift 1)", 1u << 27, 1u << 20, 79412u, 1, 128);
        bench_case<uint32_t>("C3-like (154799 bins)", 1u << 27, 1u << 20, 154799u, 0, 128);
        bench_case<uint32_t>("C3-like, shift 1", 1u << 27, 1u << 20, 154799u, 1, 128);
        bench_case<uint16_t>("shard (16 Mi impulses)", 1u << 24, 1u << 17, 79412u, 1, 128);
    }
    return ok ? 0 : 1;
}
