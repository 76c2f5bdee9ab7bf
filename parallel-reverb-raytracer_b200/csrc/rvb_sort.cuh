// Hand-written LSD radix sort of (key, 32-bit value) pairs ("onesweep": one read and one write of
// the data per digit, tile prefixes by decoupled look-back), specialised for what this library sorts:
//
//   * the FIRST pass reads through a functor, so keys that are a cheap function of data already in
//     memory (flattenImpulses' sample index = round(time * sr), rayverb.cpp:69) are never stored: the
//     histogram kernel and the first pass both evaluate the functor;
//   * elements the functor reports as invalid (impulses the attenuators skip, kernel.cpp:524,607) are
//     neither counted nor scattered -- the sorted stream is COMPACT;
//   * digits are 8, 9 or 10 bits wide, chosen so that the sample-index keys of a long response
//     (18-20 bits) still take two passes;
//   * the LAST pass may drop the keys (only the values are written) and records, with one atomic per
//     run of equal keys inside a tile, where every key's run begins and ends in the output: the
//     segment bounds the histogram reduction needs, without re-reading the sorted keys.
//
// Stable (equal keys keep their input order), hence deterministic: inside a tile an element's rank is
// (warp, item, lane)-ordered through warp-private digit counters (match.any + popc, no atomics); across
// tiles the order is the tile order.  Tile ids come from an atomic counter, so a tile only ever waits
// for tiles that are already running.
#pragma once
#include "rvb_internal.cuh"

namespace rvb {
namespace rsort {

#define RVB_SORT_ITEMS 16
#ifndef RVB_SORT_RESIDENT_THREADS
#define RVB_SORT_RESIDENT_THREADS 1024   // threads per SM the pass kernel is compiled for (register cap)
#endif
#define RVB_SORT_MAX_PASSES 8
#ifndef RVB_SORT_MATCH_BATCH
#define RVB_SORT_MATCH_BATCH 1            // items ranked together (matches, then atomics, then broadcasts)
#endif

template <int THREADS> struct Shape {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * RVB_SORT_ITEMS;
};

// look-back status word: flag in the two top bits (0 = not there yet, 1 = this tile's count,
// 2 = count of this and all earlier tiles), count below.  32-bit words when n < 2^30.
template <class W> struct StatusBits;
template <> struct StatusBits<uint32_t> { static constexpr int SHIFT = 30; };
template <> struct StatusBits<unsigned long long> { static constexpr int SHIFT = 62; };

template <class KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, uint32_t lo_bit, uint32_t mask) { return (uint32_t)(key >> lo_bit) & mask; }

// Lanes of the warp that hold the same BITS-bit digit as the caller (and the same `ok`), as a ballot per bit:
// match.any is a single instruction but far slower than these 8-10 votes when a warp holds many distinct digits.
// COHERENT (the top digit of keys that arrive roughly in order: a warp holds one to three distinct values):
// there match.any, whose cost grows with the number of distinct values, is the faster one.
template <int BITS, bool COHERENT>
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool ok)
{
    if (COHERENT) return __match_any_sync(0xffffffffu, ok ? d : (1u << BITS));
    const unsigned v = __ballot_sync(0xffffffffu, ok);
    unsigned peers = ok ? v : ~v;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        unsigned mask;   // lanes whose bit b equals the caller's: test, vote, complement under the predicate
        asm("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.u32 p, t, 0;\nvote.ballot.sync.b32 %0, p, 0xffffffff;\n@!p not.b32 %0, %0;\n}\n"
            : "=r"(mask) : "r"(d), "r"(1u << b));
        peers &= mask;
    }
    return peers;
}

// Input functors have two halves so that a tile's loads are all in flight before the first one is used:
//   Raw fetch(i)                       unconditional read(s) of element i
//   bool decode(i, raw, key, value)    false: the element does not take part in the sort
// Functor over stored arrays (every pass after the first; first pass of a plain array sort):
template <class KeyT>
struct ArrayIn {
    const KeyT *keys;
    const uint32_t *vals;   // nullptr: value = index
    struct Raw { KeyT k; uint32_t v; };
    __device__ __forceinline__ Raw fetch(uint32_t i) const { return Raw{keys[i], vals ? vals[i] : i}; }
    __device__ __forceinline__ bool decode(uint32_t, const Raw &r, KeyT &k, uint32_t &v) const { k = r.k; v = r.v; return true; }
};

struct PassPlan {
    int npasses = 0;
    int bits = 8;                                // digit width of every pass
    uint32_t lo_bit[RVB_SORT_MAX_PASSES] = {0};  // digit p = (key >> lo_bit[p]) & (2^bits - 1)
};

// digits for keys of `key_bits` significant bits starting at bit `begin_bit`
static inline PassPlan plan_passes(int begin_bit, int key_bits)
{
    PassPlan pl;
    if (key_bits < 1) key_bits = 1;
    int np = (key_bits + 9) / 10;                // fewest passes with digits of <= 10 bits
    int bits = (key_bits + np - 1) / np;
    if (bits < 8) bits = 8;                      // the kernels are built for 8, 9 and 10
    pl.npasses = np;
    pl.bits = bits;
    for (int p = 0; p < np; ++p) pl.lo_bit[p] = (uint32_t)(begin_bit + p * bits);
    return pl;
}

// ---- histogram of every pass's digit, one sweep over the input ------------------------------------
// SINK: called with (index, valid, key, value) for every element -- the flatten's HRTF variant stores its keys
// there (they need the impulse positions, not worth evaluating twice).
struct NoSink { __device__ __forceinline__ void operator()(uint32_t, bool, uint32_t, uint32_t) const {} };

#define RVB_SORT_HIST_ITEMS 8
template <int BITS, int NP, class KeyT, class In, class Sink>
__global__ void __launch_bounds__(512) digit_histogram_kernel(const In in, const Sink sink, uint32_t n, PassPlan plan, uint32_t *hist /* [NP][2^BITS] */)
{
    constexpr uint32_t RADIX = 1u << BITS;
    constexpr int ITEMS = RVB_SORT_HIST_ITEMS;
    __shared__ uint32_t sh[NP][RADIX];
    for (uint32_t d = threadIdx.x; d < (uint32_t)NP * RADIX; d += blockDim.x) (&sh[0][0])[d] = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t warps = blockDim.x >> 5;
    // a warp takes 32 * ITEMS consecutive elements per round (whole warps, so that match.any sees 32 lanes)
    constexpr uint32_t CHUNK = 32u * ITEMS;
    const uint32_t nchunks = (uint32_t)(((unsigned long long)n + CHUNK - 1) / CHUNK);
    const uint32_t lo0 = plan.lo_bit[0], lo1 = plan.lo_bit[NP > 1 ? 1 : 0], lo2 = plan.lo_bit[NP > 2 ? 2 : 0];
    for (uint32_t c = blockIdx.x * warps + warp; c < nchunks; c += gridDim.x * warps) {
        const uint32_t base = c * CHUNK + lane;   // < n + 32: fits as long as n <= 2^32 - 33 (checked by the host)
        const uint32_t rem = n - c * CHUNK;       // elements from the chunk's start on (>= 1)
        typename In::Raw raw[ITEMS];
#pragma unroll
        for (int u = 0; u < ITEMS; ++u)
            if (lane + 32u * u < rem) raw[u] = in.fetch(base + 32u * u);
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const bool inside = lane + 32u * u < rem;
            KeyT k = 0;
            uint32_t v = 0;
            const bool ok = inside && in.decode(base + 32u * u, raw[u], k, v);
            if (inside) sink(base + 32u * u, ok, (uint32_t)k, v);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t d = digit_of(k, p == 0 ? lo0 : (p == 1 ? lo1 : lo2), RADIX - 1u);
                if (p == NP - 1 && NP > 1) {
                    // the top digit of neighbouring impulses mostly coincides: one atomic per group of equal digits
                    const unsigned m = match_digit<BITS, true>(d, ok);
                    if (ok && lane == (uint32_t)(__ffs(m) - 1)) atomicAdd(&sh[p][d], (uint32_t)__popc(m));
                } else if (ok) atomicAdd(&sh[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < (uint32_t)NP * RADIX; d += blockDim.x) {
        const uint32_t c = (&sh[0][0])[d];
        if (c) atomicAdd(hist + d, c);
    }
}

// exclusive scan of every pass's histogram (one CTA per pass); total[p] = number of valid elements
template <int BITS>
__global__ void __launch_bounds__(1 << BITS) digit_scan_kernel(uint32_t *hist, uint32_t *total)
{
    constexpr uint32_t RADIX = 1u << BITS;
    __shared__ uint32_t ws[33];
    uint32_t *h = hist + (size_t)blockIdx.x * RADIX;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31u) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < RADIX / 32u ? ws[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        ws[lane] = winc - w;
        if (lane == 31u) ws[32] = winc;
    }
    __syncthreads();
    h[threadIdx.x] = ws[warp] + inc - v;
    if (threadIdx.x == 0 && total) total[blockIdx.x] = ws[32];
}

// ---- one pass ---------------------------------------------------------------------------------------
template <class KeyT, class StatusT>
struct PassArgs {
    uint32_t n;                   // elements the input functor covers (upper bound when n_ptr is set)
    const uint32_t *n_ptr;        // device: actual element count (passes after a compacting first pass)
    uint32_t lo_bit;
    const uint32_t *digit_base;   // [RADIX] exclusive global offsets of this pass's digits
    uint32_t *tile_counter;       // zeroed
    StatusT *status;              // [tiles][RADIX], zeroed
    KeyT *out_keys;               // nullptr: the keys are dropped (last pass)
    uint32_t *out_vals;
    uint32_t *run_begin, *run_end;   // last pass of a sort that drops its keys, not nullptr: key k's run in the
                                     // output is [run_begin[k], run_end[k]) (initialised to 0xffffffff / 0)
};

// KEYS_OUT = true: keys and values are written (a pass that another pass follows, or a plain sort's last pass);
// KEYS_OUT = false: the last pass of a sort whose keys are dropped -- values only, plus the run bounds if asked for.
// TOP_DIGIT: this pass sorts by the keys' most significant digit (of a sort with more than one pass): see match_digit.
template <int BITS, int THREADS, class KeyT, class StatusT, class In, bool KEYS_OUT, bool TOP_DIGIT>
__global__ void __launch_bounds__(THREADS, RVB_SORT_RESIDENT_THREADS / THREADS) onesweep_pass_kernel(const In in, const PassArgs<KeyT, StatusT> A)
{
    constexpr uint32_t RADIX = 1u << BITS;
    constexpr int WARPS = Shape<THREADS>::WARPS;
    constexpr uint32_t TILE = Shape<THREADS>::TILE;
    constexpr int ITEMS = RVB_SORT_ITEMS;
    constexpr int SB = StatusBits<StatusT>::SHIFT;
    static_assert(RADIX % THREADS == 0 || THREADS % RADIX == 0, "digits and threads must divide");
    static_assert(sizeof(uint32_t) * WARPS * RADIX <= (sizeof(KeyT) + 4) * TILE, "the warp counters live in the staging buffers");
    static_assert(TILE <= 65536, "staged positions are kept in 16 bits");

    // staging buffers of the local sort; the warp-private digit counters use the same bytes before the
    // staging starts
    __shared__ __align__(16) unsigned char s_raw[(sizeof(KeyT) + 4) * TILE];
    __shared__ uint32_t s_count[RADIX];   // elements of this tile per digit
    __shared__ uint32_t s_excl[RADIX];    // exclusive scan of s_count: start of the digit's run in the staged tile
    __shared__ uint32_t s_base[RADIX];    // output position of staged element j with digit d = s_base[d] + j
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tile;
    KeyT *s_keys = reinterpret_cast<KeyT *>(s_raw);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_raw + sizeof(KeyT) * TILE);
    uint32_t *s_wh = reinterpret_cast<uint32_t *>(s_raw);

    if (threadIdx.x == 0) s_tile = atomicAdd(A.tile_counter, 1u);
    for (uint32_t d = threadIdx.x; d < (uint32_t)WARPS * RADIX; d += THREADS) s_wh[d] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n = A.n_ptr ? *A.n_ptr : A.n;
    if ((unsigned long long)tile * TILE >= n) return;   // block-uniform; no later tile looks at this one
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t base = tile * TILE;                  // < n: no wrap
    const uint32_t rem = n - base;                      // elements from `base` on (>= 1)
    const uint32_t off0 = warp * (32u * ITEMS) + lane;  // this thread's first element inside the tile
    const uint32_t lo_bit = A.lo_bit;

    KeyT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t okmask = 0;
    {
        typename In::Raw raw[ITEMS];
#pragma unroll
        for (int u = 0; u < ITEMS; ++u)
            if (off0 + 32u * u < rem) raw[u] = in.fetch(base + off0 + 32u * u);
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            key[u] = 0; val[u] = 0;
            if (off0 + 32u * u < rem && in.decode(base + off0 + 32u * u, raw[u], key[u], val[u])) okmask |= 1u << u;
        }
    }

    // rank inside the warp: elements with equal digits in (item, lane) order.  The lowest lane of every
    // group of equal digits bumps the warp's private counter by the group size (a shared-memory atomic, so
    // that the sixteen items pipeline instead of waiting for each other's read-modify-write; atomics of one
    // warp execute in program order) and broadcasts the old count.
    uint32_t *my = s_wh + warp * RADIX;
    const uint32_t my_saddr = (uint32_t)__cvta_generic_to_shared(my);
    uint32_t rank2[ITEMS / 2];   // two 16-bit ranks (later: staged positions) per register
#pragma unroll
    for (int u = 0; u < ITEMS / 2; ++u) rank2[u] = 0u;
    // in batches of eight items: all matches first (they are slow and independent), then the leaders'
    // atomics as single predicated instructions (no branch, so they pipeline), then the broadcasts
    constexpr int BATCH = RVB_SORT_MATCH_BATCH;
#pragma unroll
    for (int h = 0; h < ITEMS; h += BATCH) {
        unsigned m[BATCH];
        uint32_t prev[BATCH];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int u = h + q;
            const bool ok = (okmask >> u) & 1u;
            m[q] = match_digit<BITS, TOP_DIGIT>(digit_of(key[u], lo_bit, RADIX - 1u), ok);
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int u = h + q;
            const bool ok = (okmask >> u) & 1u;
            const uint32_t d = digit_of(key[u], lo_bit, RADIX - 1u);
            const uint32_t lead = (ok && (int)lane == __ffs(m[q]) - 1) ? 1u : 0u;
            prev[q] = 0u;
            asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\n@q atom.shared.add.u32 %0, [%1], %2;\n}\n"
                         : "+r"(prev[q]) : "r"(my_saddr + 4u * d), "r"((uint32_t)__popc(m[q])), "r"(lead) : "memory");
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int u = h + q;
            const uint32_t p = __shfl_sync(0xffffffffu, prev[q], __ffs(m[q]) - 1);
            rank2[u >> 1] |= (p + __popc(m[q] & lt)) << (16 * (u & 1));
        }
    }
    __syncthreads();

    // per digit: the warps' counts become the warps' offsets inside the digit's run
    for (uint32_t d = threadIdx.x; d < RADIX; d += THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_wh[w * RADIX + d];
            s_wh[w * RADIX + d] = run;
            run += c;
        }
        s_count[d] = run;
    }
    __syncthreads();

    // decoupled look-back: elements with digit d in all earlier tiles
    for (uint32_t d = threadIdx.x; d < RADIX; d += THREADS) {
        const StatusT count = s_count[d];
        volatile StatusT *st = A.status + (size_t)tile * RADIX + d;
        StatusT prev = 0;
        if (tile == 0) {
            *st = ((StatusT)2 << SB) | count;
        } else {
            *st = ((StatusT)1 << SB) | count;
            const volatile StatusT *look = st - RADIX;
            while (true) {
                const StatusT sv = *look;
                const uint32_t flag = (uint32_t)(sv >> SB);
                if (flag == 0u) { __nanosleep(20); continue; }
                prev += sv & ((((StatusT)1) << SB) - 1);
                if (flag == 2u) break;
                look -= RADIX;
            }
            *st = ((StatusT)2 << SB) | (prev + count);
        }
        s_base[d] = A.digit_base[d] + (uint32_t)prev;
    }
    // exclusive scan of the tile's digit counts (warp 0; RADIX / 32 consecutive digits per lane)
    if (warp == 0) {
        constexpr uint32_t PER = RADIX / 32u;
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < PER; ++q) sum += s_count[lane * PER + q];
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        uint32_t ex = inc - sum;
#pragma unroll
        for (uint32_t q = 0; q < PER; ++q) { s_excl[lane * PER + q] = ex; ex += s_count[lane * PER + q]; }
        if (lane == 31u) s_scan[0] = inc;   // valid elements in this tile
    }
    __syncthreads();
    const uint32_t tile_total = s_scan[0];

    // staged position of every element; then the counters' bytes are free
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
        if ((okmask >> u) & 1u) {
            const uint32_t d = digit_of(key[u], lo_bit, RADIX - 1u);
            rank2[u >> 1] += (s_excl[d] + my[d]) << (16 * (u & 1));   // < TILE <= 2^16: no carry into the neighbour
        }
    }
    // output position of staged element j with digit d: s_base[d] + j
    for (uint32_t d = threadIdx.x; d < RADIX; d += THREADS) s_base[d] -= s_excl[d];
    __syncthreads();
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
        if ((okmask >> u) & 1u) {
            const uint32_t lp = (rank2[u >> 1] >> (16 * (u & 1))) & 0xffffu;
            s_keys[lp] = key[u]; s_vals[lp] = val[u];
        }
    __syncthreads();

    // scatter: consecutive staged elements of a digit go to consecutive output positions
    KeyT *const out_keys = A.out_keys;
    uint32_t *const out_vals = A.out_vals;
    uint32_t *const run_begin = A.run_begin, *const run_end = A.run_end;
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
        const uint32_t j = (uint32_t)u * THREADS + threadIdx.x;
        if (j < tile_total) {
            const KeyT k = s_keys[j];
            const uint32_t pos = s_base[digit_of(k, lo_bit, RADIX - 1u)] + j;
            if (KEYS_OUT) out_keys[pos] = k;
            out_vals[pos] = s_vals[j];
            if (!KEYS_OUT && run_begin) {
                if (j == 0u || s_keys[j - 1u] != k) atomicMin(run_begin + (size_t)k, pos);
                if (j + 1u == tile_total || s_keys[j + 1u] != k) atomicMax(run_end + (size_t)k, pos + 1u);
            }
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------
// Scratch of one sort: digit histograms, per-pass totals, tile counters and look-back status words.
struct Scratch {
    uint32_t *d_hist = nullptr;      // [npasses][RADIX] -> exclusive offsets
    uint32_t *d_total = nullptr;     // [npasses] valid elements (identical for every pass)
    uint32_t *d_counter = nullptr;   // [npasses]
    void *d_status = nullptr;        // [npasses][tiles][RADIX]
    size_t status_bytes_per_pass = 0;
    void release()
    {
        dev_free(d_hist); dev_free(d_total); dev_free(d_counter);
        if (d_status) free_bytes(d_status);
        d_status = nullptr;
    }
};

template <int THREADS>
static inline uint32_t tiles_for(uint32_t n) { return (uint32_t)(((unsigned long long)n + Shape<THREADS>::TILE - 1) / Shape<THREADS>::TILE); }

template <int THREADS>
static inline int scratch_alloc(Scratch &sc, const PassPlan &pl, uint32_t n, bool wide_status, cudaStream_t st)
{
    const size_t radix = (size_t)1 << pl.bits;
    RVB_TRY(dev_alloc(&sc.d_hist, (size_t)pl.npasses * radix));
    RVB_TRY(dev_alloc(&sc.d_total, (size_t)pl.npasses));
    RVB_TRY(dev_alloc(&sc.d_counter, (size_t)pl.npasses));
    sc.status_bytes_per_pass = (size_t)tiles_for<THREADS>(n) * radix * (wide_status ? 8 : 4);
    RVB_TRY(alloc_bytes(&sc.d_status, sc.status_bytes_per_pass * pl.npasses));
    RVB_CUDA(cudaMemsetAsync(sc.d_hist, 0, (size_t)pl.npasses * radix * sizeof(uint32_t), st));
    RVB_CUDA(cudaMemsetAsync(sc.d_counter, 0, (size_t)pl.npasses * sizeof(uint32_t), st));
    RVB_CUDA(cudaMemsetAsync(sc.d_status, 0, sc.status_bytes_per_pass * pl.npasses, st));
    return RVB_OK;
}

template <int BITS, int NP, class KeyT, class In, class Sink>
static inline void launch_histogram_np(const In &in, const Sink &sink, uint32_t n, const PassPlan &sub, uint32_t *d_hist, cudaStream_t st)
{
    const size_t per_cta = (size_t)512 * RVB_SORT_HIST_ITEMS;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(((size_t)n + per_cta - 1) / per_cta, (size_t)sm_count() * 4));
    note_launch(), digit_histogram_kernel<BITS, NP, KeyT, In, Sink><<<grid, 512, 0, st>>>(in, sink, n, sub, d_hist);
}

template <int BITS, class KeyT, class In, class Sink>
static inline int launch_histogram(const In &in, const Sink &sink, uint32_t n, const PassPlan &pl, Scratch &sc, cudaStream_t st)
{
    // up to three passes are counted per sweep over the input; more (wide keys): further sweeps
    for (int p0 = 0; p0 < pl.npasses; p0 += 3) {
        PassPlan sub;
        sub.bits = pl.bits;
        sub.npasses = std::min(3, pl.npasses - p0);
        for (int p = 0; p < sub.npasses; ++p) sub.lo_bit[p] = pl.lo_bit[p0 + p];
        uint32_t *h = sc.d_hist + ((size_t)p0 << BITS);
        if (p0 == 0) {
            if (sub.npasses == 1) launch_histogram_np<BITS, 1, KeyT>(in, sink, n, sub, h, st);
            else if (sub.npasses == 2) launch_histogram_np<BITS, 2, KeyT>(in, sink, n, sub, h, st);
            else launch_histogram_np<BITS, 3, KeyT>(in, sink, n, sub, h, st);
        } else {
            if (sub.npasses == 1) launch_histogram_np<BITS, 1, KeyT>(in, NoSink{}, n, sub, h, st);
            else if (sub.npasses == 2) launch_histogram_np<BITS, 2, KeyT>(in, NoSink{}, n, sub, h, st);
            else launch_histogram_np<BITS, 3, KeyT>(in, NoSink{}, n, sub, h, st);
        }
    }
    note_launch(), digit_scan_kernel<BITS><<<pl.npasses, 1 << BITS, 0, st>>>(sc.d_hist, sc.d_total);
    RVB_CUDA(cudaGetLastError());
    return RVB_OK;
}

template <int BITS, int THREADS, class KeyT, class StatusT, class In>
static inline int launch_pass(const In &in, int p, const PassPlan &pl, Scratch &sc, uint32_t n, const uint32_t *n_ptr, KeyT *out_keys,
                              uint32_t *out_vals, uint32_t *run_begin, uint32_t *run_end, cudaStream_t st)
{
    PassArgs<KeyT, StatusT> A;
    A.n = n; A.n_ptr = n_ptr; A.lo_bit = pl.lo_bit[p];
    A.digit_base = sc.d_hist + ((size_t)p << BITS);
    A.tile_counter = sc.d_counter + p;
    A.status = reinterpret_cast<StatusT *>(static_cast<unsigned char *>(sc.d_status) + sc.status_bytes_per_pass * p);
    A.out_keys = out_keys; A.out_vals = out_vals; A.run_begin = run_begin; A.run_end = run_end;
    note_launch();
    const bool top = pl.npasses > 1 && p + 1 == pl.npasses;
    const unsigned grid = tiles_for<THREADS>(n);
    if (out_keys && top) onesweep_pass_kernel<BITS, THREADS, KeyT, StatusT, In, true, true><<<grid, THREADS, 0, st>>>(in, A);
    else if (out_keys) onesweep_pass_kernel<BITS, THREADS, KeyT, StatusT, In, true, false><<<grid, THREADS, 0, st>>>(in, A);
    else if (top) onesweep_pass_kernel<BITS, THREADS, KeyT, StatusT, In, false, true><<<grid, THREADS, 0, st>>>(in, A);
    else onesweep_pass_kernel<BITS, THREADS, KeyT, StatusT, In, false, false><<<grid, THREADS, 0, st>>>(in, A);
    RVB_CUDA(cudaGetLastError());
    return RVB_OK;
}

// Complete sort.  `hist_in` is swept once for the digit histograms (its elements also go to `sink`),
// `first` is what the first pass reads -- the same functor, or one that reads what the sink stored.
// Buffers: keys[2] / vals[2] of n elements each (ping-pong; keys[1] is only touched by sorts of three
// and more passes, keys[0] not by single-pass sorts that drop their keys); *result says which of the
// two holds the sorted values (and keys, unless drop_last_keys).  d_nvalid (optional, device) receives
// the number of elements that were sorted.
template <int BITS, int THREADS, class KeyT, class StatusT, class HIn, class Sink, class FIn>
static inline int sort_bits(const HIn &hist_in, const Sink &sink, const FIn &first, uint32_t n, const PassPlan &pl, KeyT *const keys[2],
                            uint32_t *const vals[2], bool drop_last_keys, uint32_t *run_begin, uint32_t *run_end, int *result,
                            uint32_t *d_nvalid, cudaStream_t st)
{
    Scratch sc;
    int rc = scratch_alloc<THREADS>(sc, pl, n, sizeof(StatusT) == 8, st);
    if (rc == RVB_OK) rc = launch_histogram<BITS, KeyT>(hist_in, sink, n, pl, sc, st);
    int cur = 0;   // pass p writes buffer p & 1; the first pass reads the functor
    for (int p = 0; p < pl.npasses && rc == RVB_OK; ++p) {
        const bool last = p + 1 == pl.npasses;
        KeyT *ok = (last && drop_last_keys) ? nullptr : keys[p & 1];
        uint32_t *rb = last ? run_begin : nullptr, *re = last ? run_end : nullptr;
        if (p == 0) rc = launch_pass<BITS, THREADS, KeyT, StatusT>(first, p, pl, sc, n, nullptr, ok, vals[0], rb, re, st);
        else {
            ArrayIn<KeyT> in{keys[(p - 1) & 1], vals[(p - 1) & 1]};
            rc = launch_pass<BITS, THREADS, KeyT, StatusT>(in, p, pl, sc, n, sc.d_total, ok, vals[p & 1], rb, re, st);
        }
        cur = p & 1;
    }
    if (rc == RVB_OK && d_nvalid && cudaMemcpyAsync(d_nvalid, sc.d_total, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = fail(RVB_ERR_CUDA, "radix sort: copy of the element count failed");
    sc.release();
    *result = cur;
    return rc;
}

template <int THREADS, class KeyT, class HIn, class Sink, class FIn>
static inline int sort_pairs(const HIn &hist_in, const Sink &sink, const FIn &first, uint32_t n, const PassPlan &pl, KeyT *const keys[2],
                             uint32_t *const vals[2], bool drop_last_keys, uint32_t *run_begin, uint32_t *run_end, int *result,
                             uint32_t *d_nvalid, cudaStream_t st)
{
    const bool wide = n >= (1u << 30);
#define RVB_SORT_CASE(B)                                                                                                                   \
    case B:                                                                                                                                \
        return wide ? sort_bits<B, THREADS, KeyT, unsigned long long>(hist_in, sink, first, n, pl, keys, vals, drop_last_keys, run_begin,   \
                                                                      run_end, result, d_nvalid, st)                                       \
                    : sort_bits<B, THREADS, KeyT, uint32_t>(hist_in, sink, first, n, pl, keys, vals, drop_last_keys, run_begin, run_end,    \
                                                            result, d_nvalid, st);
    if constexpr (sizeof(KeyT) == 2) {   // 16-bit keys: one or two passes of eight bits
        switch (pl.bits) {
            RVB_SORT_CASE(8)
        }
    } else {
        switch (pl.bits) {
            RVB_SORT_CASE(8)
            RVB_SORT_CASE(9)
            RVB_SORT_CASE(10)
        }
    }
#undef RVB_SORT_CASE
    return fail(RVB_ERR_INVALID, "radix sort: unsupported digit width");
}

// Plain sort of stored pairs by the key bits [begin_bit, begin_bit + nbits).  keys[1] / vals[1] hold the input
// (vals[1] == nullptr: the values are the indices 0..n-1) and are overwritten; *result says which pair of buffers
// holds the sorted output.
template <class KeyT>
static inline int sort_arrays(KeyT *const keys[2], uint32_t *const vals[2], uint32_t n, int begin_bit, int nbits, int *result, cudaStream_t st)
{
    constexpr int THREADS = sizeof(KeyT) == 8 ? 128 : 256;   // 64-bit keys: smaller tiles (static shared memory)
    *result = 1;
    if (n == 0) return RVB_OK;
    PassPlan pl = plan_passes(begin_bit, nbits);
    if (sizeof(KeyT) == 2 && pl.bits != 8) { pl.bits = 8; pl.npasses = (nbits + 7) / 8; for (int p = 0; p < pl.npasses; ++p) pl.lo_bit[p] = (uint32_t)(begin_bit + 8 * p); }
    ArrayIn<KeyT> in{keys[1], vals[1]};
    return sort_pairs<THREADS, KeyT>(in, NoSink{}, in, n, pl, keys, vals, false, nullptr, nullptr, result, nullptr, st);
}

}  // namespace rsort
}  // namespace rvb
