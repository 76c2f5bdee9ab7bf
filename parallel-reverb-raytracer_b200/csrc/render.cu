// Fused device-resident path from trace results to the impulse response: everything
// cmd/main.cpp:254-336 does after raytrace() --
//   getAllRaw / getRawImages / getRawDiffuse           (rayverb.cpp:687-714)
//   SpeakerAttenuator / HrtfAttenuator                 (kernel.cpp:506-625)
//   fixPredelay                                        (rayverb.h:47-97)
//   flattenImpulses                                    (rayverb.cpp:48-77)
//   RayverbFiltering::filter, mixdown, normalize, mul, trimTail   (filters.cpp:156-306, rayverb.cpp:80-149)
// -- without the impulses leaving the GPU.
//
// flattenImpulses is a serial float scatter-add in the reference.  Here it is a deterministic
// sort-by-sample-index (LSD radix sort of (bin, impulse id) pairs, stable => fixed order inside
// a bin) followed by a segmented reduction with a fixed lane assignment and a fixed shuffle
// tree: no float atomics, bit-identical from run to run.
#include <cmath>
#include <memory>
#include <type_traits>

#include "attenuate.cuh"
#include "rvb_internal.cuh"
#include "rvb_sort.cuh"

namespace rvb {

#define RVB_MAX_GROUP 4   // speaker channels reduced per pass over a bin
#define RVB_NO_IMPULSE 0xffffffffu

// The impulse list seen by the attenuators: diffuse slots first, then image sources
// (getAllRaw, rayverb.cpp:708-714).
struct Source {
    const float4 *vol, *pt;          // vol: 2 x float4 per impulse (one 32-B sector)
    const float *tkey;               // time if any(volume != 0) else -1
    uint32_t n_diffuse;
    const float4 *ivolA, *ivolB, *ipt;
    uint32_t n_images;
    // host side: extrema of the diffuse times already reduced by the bounce kernels (time passes
    // through the speaker attenuator unchanged, kernel.cpp:532)
    bool have_ext;
    float ext_min_nz, ext_max;
    __device__ __forceinline__ uint32_t size() const { return n_diffuse + n_images; }
    // returns false for impulses the attenuation kernels skip (kernel.cpp:524,607)
    __device__ __forceinline__ bool time_of(uint32_t i, float &t) const
    {
        if (i < n_diffuse) { t = tkey[i]; return t >= 0.0f; }
        uint32_t j = i - n_diffuse;
        float4 a = ivolA[j], b = ivolB[j];
        t = ipt[j].w;
        return a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f || b.x != 0.f || b.y != 0.f || b.z != 0.f || b.w != 0.f;
    }
    __device__ __forceinline__ float4 pos_time(uint32_t i) const { return i < n_diffuse ? pt[i] : ipt[i - n_diffuse]; }
    __device__ __forceinline__ void volume(uint32_t i, float4 &a, float4 &b) const
    {
        if (i < n_diffuse) { a = vol[2 * (size_t)i]; b = vol[2 * (size_t)i + 1]; } else { a = ivolA[i - n_diffuse]; b = ivolB[i - n_diffuse]; }
    }
};

struct ChanSpec {
    int model;          // 0 speakers, 1 hrtf, 2 none (gain 1: rvb_flatten on host arrays)
    int nch;            // channels in this pass (speakers: <= RVB_MAX_GROUP, hrtf/none: 1)
    f3 mic;
    f3 sdir[RVB_MAX_GROUP];
    float coef[RVB_MAX_GROUP];
    f3 pointing, up, ear;
    const float4 *table;   // one ear, padded by one row
    unsigned long long channel;
};

__device__ __forceinline__ int ordered(float f) { int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__host__ __device__ inline float unordered(int i) { i = i ^ ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &i, 4); return f; }

// attenuated time of impulse i for this channel spec (before fixPredelay)
__device__ __forceinline__ bool attenuated_time(const Source &S, const ChanSpec &C, uint32_t i, float &t)
{
    if (!S.time_of(i, t)) return false;
    if (C.model == 1) {
        float4 p = S.pos_time(i);
        t = hrtf_time(p.w, mk3(p), C.mic, C.ear);
    }
    return true;
}

// min non-zero / max attenuated time (findPredelay rayverb.h:47-76, flattenImpulses rayverb.cpp:54-57)
__global__ void times_kernel(const Source S, const ChanSpec C, int *out /* [0] ordered min nonzero, [1] ordered max */)
{
    float mn = __int_as_float(0x7f800000), mx = 0.0f;   // flatten's maxtime starts at 0
    const uint32_t n = S.size();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float t;
        if (!attenuated_time(S, C, i, t)) continue;
        if (t != 0.0f) mn = fminf(mn, t);
        mx = fmaxf(mx, t);
    }
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(out, ordered(mn)); atomicMax(out + 1, ordered(mx)); }
}

// fixPredelay (rayverb.h:85-89) + SAMPLE = round(time * samplerate) (rayverb.cpp:69)
__device__ __forceinline__ uint32_t sample_index(float t, float predelay, bool fix, float sr)
{
    if (fix) t = t > predelay ? __fsub_rn(t, predelay) : 0.0f;
    // round(x), half away from zero (rayverb.cpp:69), without the library call: rint rounds half to even, so it
    // differs only when the fraction is exactly one half and it went down -- x - r is exact and equals 0.5 then.
    // (negative x, NaN and infinities convert to the same integers either way)
    const float x = __fmul_rn(t, sr);
    float r = rintf(x);
    if (__fsub_rn(x, r) == 0.5f) r = __fadd_rn(r, 1.0f);
    return (uint32_t)r;
}

#define RVB_ID_MASK 0x3fffffffu

// ---- the flatten's own sort (rvb_sort.cuh) ----------------------------------------------------------
// Key = segment id (sample index >> shift), value = impulse id with the sub-bin inside the segment in
// bits 30-31.  The sample index is a function of the impulse time, so the speaker / pass-through models
// never store keys: the digit histogram and the first radix pass both evaluate it from the 4-byte time.
// The HRTF model's time needs the impulse position (kernel.cpp:612-619); there the histogram sweep stores
// the sample indices and the first pass reads them back.  Skipped impulses (time key -1, kernel.cpp:524)
// and samples past the end never enter the sort.
// image sources: time key as for the diffuse slots -- the time, or -1 when all eight volumes are zero (kernel.cpp:524)
__global__ void image_tkeys_kernel(const float4 *volA, const float4 *volB, const float4 *pt, uint32_t n, float *tkey)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 a = volA[j], b = volB[j];
    const bool any = a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f || b.x != 0.f || b.y != 0.f || b.z != 0.f || b.w != 0.f;
    tkey[j] = any ? pt[j].w : -1.0f;
}

enum { FLATTEN_TIME = 0, FLATTEN_TIME_POS = 1, FLATTEN_STORED = 2 };
template <class KeyT, int MODE>
struct FlattenIn {
    const float *tkey, *img_tkey;          // time keys of the diffuse slots / of the image sources
    const float4 *pt, *img_pt;             // FLATTEN_TIME_POS: (position, time)
    const uint32_t *stored;                // FLATTEN_STORED: sample index per impulse, 0xffffffff = skipped
    uint32_t n_diffuse;
    f3 mic, ear;
    float predelay, sr;
    uint32_t L, shift;
    bool fix;
    struct RawT { float t; };
    struct RawTP { float t; float4 p; };
    typedef typename std::conditional<MODE == FLATTEN_TIME_POS, RawTP, RawT>::type Raw;
    __device__ __forceinline__ Raw fetch(uint32_t i) const
    {
        Raw r{};
        if constexpr (MODE == FLATTEN_STORED) r.t = __uint_as_float(stored[i]);
        else {
            const bool img = i >= n_diffuse;
            r.t = (img ? img_tkey - n_diffuse : tkey)[i];
            if constexpr (MODE == FLATTEN_TIME_POS) r.p = (img ? img_pt - n_diffuse : pt)[i];
        }
        return r;
    }
    __device__ __forceinline__ bool decode(uint32_t i, const Raw &r, KeyT &k, uint32_t &v) const
    {
        uint32_t s;
        if constexpr (MODE == FLATTEN_STORED) s = __float_as_uint(r.t);   // 0xffffffff >= L
        else {
            float t = r.t;
            if (!(t >= 0.0f)) return false;                                                           // Source::time_of
            if constexpr (MODE == FLATTEN_TIME_POS) t = hrtf_time(r.p.w, mk3(r.p), mic, ear);         // attenuated_time
            s = sample_index(t, predelay, fix, sr);
        }
        if (s >= L) return false;
        k = (KeyT)(s >> shift);
        v = i | ((s & ((1u << shift) - 1u)) << 30);   // shift > 0 only when i < 2^30
        return true;
    }
};
struct StoreSampleIndex {
    uint32_t *out;
    uint32_t shift;
    __device__ __forceinline__ void operator()(uint32_t i, bool ok, uint32_t k, uint32_t v) const
    {
        out[i] = ok ? ((k << shift) | (shift ? v >> 30 : 0u)) : 0xffffffffu;
    }
};

// One warp per segment: lane l takes entries begin+l, begin+l+32, ... in order (four gathers in
// flight per lane), then a fixed xor-shuffle tree combines the 32 partial sums of every
// (sub-bin, channel, band).  hist layout: [channel][band][L].
// Every lane adds its (few) impulses in float, as the reference does (rayverb.cpp:67-74); the lanes'
// and warps' partial sums are then combined in DOUBLE and rounded to float once per bin.  With the
// sign of the volume flipping at every reflection (kernel.cpp:461) the high bands cancel heavily, and
// an all-float tree in a different grouping -- one GPU vs. eight partial histograms -- differed by up to
// 2e-6 relative L2.  (Accumulating every impulse in double was measured too: the conversions cost
// +0.5 ms on C2 and far more on the two-channel C3 reduction.)
// NSUB sub-bins x MAXC channels <= RVB_MAX_GROUP accumulator sets.
template <int MODEL, bool NEED_POS, int NSUB, int MAXC>
__device__ __forceinline__ void accumulate_one(const Source &S, const ChanSpec &C, uint32_t i, uint32_t sub, float4 va, float4 vb,
                                               float (&acc)[NSUB][MAXC][RVB_NBANDS])
{
    const float v[RVB_NBANDS] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
    float term[MAXC][RVB_NBANDS];
    if (MODEL == 0) {
        f3 pos = mk3(0.f, 0.f, 0.f);
        if (NEED_POS) pos = mk3(S.pos_time(i));
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            // coefficient 0 (omni): (1 - 0) + 0 * d == 1 exactly for every finite d, so the
            // position is not needed
            const float g = (NEED_POS && c < C.nch) ? speaker_gain(C.mic, pos, C.sdir[c], C.coef[c]) : 1.0f;
#pragma unroll
            for (int b = 0; b < RVB_NBANDS; ++b) term[c][b] = __fmul_rn(v[b], g);
        }
    } else if (MODEL == 1) {
        const f3 pos = mk3(S.pos_time(i));
        const long idx = hrtf_index(C.pointing, C.up, normalize(pos - C.mic));
        const float4 ga = __ldg(C.table + 2 * idx), gb = __ldg(C.table + 2 * idx + 1);
        const float g[RVB_NBANDS] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
        for (int b = 0; b < RVB_NBANDS; ++b) term[0][b] = __fmul_rn(v[b], g[b]);
    } else {
#pragma unroll
        for (int b = 0; b < RVB_NBANDS; ++b) term[0][b] = v[b];
    }
#pragma unroll
    for (int s = 0; s < NSUB; ++s)
        if (NSUB == 1 || s == (int)sub) {
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
#pragma unroll
                for (int b = 0; b < RVB_NBANDS; ++b) acc[s][c][b] = __fadd_rn(acc[s][c][b], term[c][b]);
        }
}

// Sum of 32 values per lane over the 32 lanes of a warp: lane l ends up with the total of value l.
// A butterfly that halves the values a lane holds at every step (16 + 8 + 4 + 2 + 1 = 31 exchanges
// instead of 32 five-step trees), in a fixed order => deterministic.  In double (see above).
__device__ __forceinline__ double warp_reduce_32x32(const float (&acc)[32], uint32_t lane)
{
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {   // step 16: lanes 16..31 keep values 16..31
        const bool up = lane & 16u;
        const double keep = (double)(up ? acc[16 + i] : acc[i]), send = (double)(up ? acc[i] : acc[16 + i]);
        v[i] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        const bool up = lane & (uint32_t)o;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const double keep = up ? v[o + i] : v[i], send = up ? v[i] : v[o + i];
            v[i] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, o));
        }
    }
    return v[0];
}

template <int MODEL, bool NEED_POS, int NSUB, int MAXC>
__global__ void __launch_bounds__(256) reduce_kernel(const Source S, const ChanSpec C, const uint32_t *keys, const uint32_t *vals,
                                                     const uint32_t *seg_begin, const uint32_t *seg_end, uint32_t nseg, uint32_t L,
                                                     float *hist, size_t chan_stride)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    // segments are handed out from the END of the response first: late bins hold the most impulses
    for (uint32_t w = warp; w < nseg; w += nwarps) {
        const uint32_t seg = nseg - 1 - w;
        const uint32_t b0 = seg_begin[seg], b1 = seg_end[seg];
        if (b0 >= b1) continue;   // warp-uniform (empty: begin 0xffffffff, end 0)
        float acc[NSUB][MAXC][RVB_NBANDS];
#pragma unroll
        for (int s = 0; s < NSUB; ++s)
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
#pragma unroll
                for (int b = 0; b < RVB_NBANDS; ++b) acc[s][c][b] = 0.0f;
        for (uint32_t j = b0 + lane; j < b1; j += 128) {
            uint32_t idx[4], sub[4];
            float4 va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                idx[u] = RVB_NO_IMPULSE; sub[u] = 0;
                if (j + 32 * u < b1) {
                    idx[u] = vals[j + 32 * u];
                    if (keys == nullptr) {                     // sub-bin in the value's top bits (ids < 2^30 whenever NSUB > 1)
                        if (NSUB > 1) { sub[u] = (idx[u] >> 30) & (NSUB - 1); idx[u] &= RVB_ID_MASK; }
                    } else if (NSUB > 1) {
                        const uint32_t k = keys[j + 32 * u];
                        if (k >= L) idx[u] = RVB_NO_IMPULSE;   // skipped impulses share the last segment
                        sub[u] = k & (NSUB - 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (idx[u] != RVB_NO_IMPULSE) S.volume(idx[u], va[u], vb[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (idx[u] != RVB_NO_IMPULSE) accumulate_one<MODEL, NEED_POS, NSUB, MAXC>(S, C, idx[u], sub[u], va[u], vb[u], acc);
        }
        {
            static_assert(NSUB * MAXC * RVB_NBANDS == 32, "one lane per (sub-bin, channel, band) after the butterfly");
            const double x = warp_reduce_32x32(reinterpret_cast<const float (&)[32]>(acc), lane);
            const uint32_t b = lane % RVB_NBANDS, c = (lane / RVB_NBANDS) % MAXC, sb = lane / (RVB_NBANDS * MAXC);
            const uint32_t bin = seg * NSUB + sb;
            if (c < (uint32_t)C.nch && bin < L) hist[(size_t)c * chan_stride + (size_t)b * L + bin] = (float)x;
        }
    }
}

// Dense responses (hundreds of impulses per bin): one CTA of eight warps per segment.  Each warp
// takes a contiguous eighth of the segment's sorted entries (same lane striding and shuffle tree as
// above), the eight partial sums are combined in warp order -- a fixed order, so the result is
// still deterministic.  What this buys is LOCALITY IN TIME: only ~2 CTAs per SM are resident, so at
// any moment the kernel works on a window of ~300 adjacent segments instead of ~9500, the two
// reflections of a slot group (rvb_internal.cuh: slot_of) are gathered within a few tens of
// megabytes of traffic of each other, and three of the four hit L2 instead of fetching their
// 128-byte line from HBM again.
template <int MODEL, bool NEED_POS, int NSUB, int MAXC, int RVB_REDUCE_WARPS>
__global__ void __launch_bounds__(32 * RVB_REDUCE_WARPS) reduce_cta_kernel(const Source S, const ChanSpec C, const uint32_t *keys,
                                                                          const uint32_t *vals, const uint32_t *seg_begin,
                                                                          const uint32_t *seg_end, uint32_t nseg, uint32_t L,
                                                                          float *hist, size_t chan_stride)
{
    static_assert(NSUB * MAXC * RVB_NBANDS == 32, "one lane per (sub-bin, channel, band) in the final combine");
    __shared__ double part[RVB_REDUCE_WARPS][32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t w = blockIdx.x; w < nseg; w += gridDim.x) {
        const uint32_t seg = nseg - 1 - w;   // from the end of the response (ascending order measured the same)
        const uint32_t b0 = seg_begin[seg], b1 = seg_end[seg];
        if (b0 >= b1) continue;   // block-uniform (empty: begin 0xffffffff, end 0)
        const uint32_t chunk = (b1 - b0 + RVB_REDUCE_WARPS - 1) / RVB_REDUCE_WARPS;
        const uint32_t m0 = min(b1, b0 + wid * chunk), m1 = min(b1, m0 + chunk);
        float acc[NSUB][MAXC][RVB_NBANDS];
#pragma unroll
        for (int s = 0; s < NSUB; ++s)
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
#pragma unroll
                for (int b = 0; b < RVB_NBANDS; ++b) acc[s][c][b] = 0.0f;
        for (uint32_t j = m0 + lane; j < m1; j += 128) {
            uint32_t idx[4], sub[4];
            float4 va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                idx[u] = RVB_NO_IMPULSE; sub[u] = 0;
                if (j + 32 * u < m1) {
                    idx[u] = vals[j + 32 * u];
                    if (keys == nullptr) {                     // sub-bin in the value's top bits (ids < 2^30 whenever NSUB > 1)
                        if (NSUB > 1) { sub[u] = (idx[u] >> 30) & (NSUB - 1); idx[u] &= RVB_ID_MASK; }
                    } else if (NSUB > 1) {
                        const uint32_t k = keys[j + 32 * u];
                        if (k >= L) idx[u] = RVB_NO_IMPULSE;   // skipped impulses share the last segment
                        sub[u] = k & (NSUB - 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (idx[u] != RVB_NO_IMPULSE) S.volume(idx[u], va[u], vb[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (idx[u] != RVB_NO_IMPULSE) accumulate_one<MODEL, NEED_POS, NSUB, MAXC>(S, C, idx[u], sub[u], va[u], vb[u], acc);
        }
        part[wid][lane] = warp_reduce_32x32(reinterpret_cast<const float (&)[32]>(acc), lane);   // value index = (s * MAXC + c) * 8 + b
        __syncthreads();
        if (wid == 0) {
            double x = part[0][lane];
#pragma unroll
            for (int q = 1; q < RVB_REDUCE_WARPS; ++q) x = __dadd_rn(x, part[q][lane]);
            const uint32_t b = lane % RVB_NBANDS, c = (lane / RVB_NBANDS) % MAXC, sb = lane / (RVB_NBANDS * MAXC);
            const uint32_t bin = seg * NSUB + sb;
            if (c < (uint32_t)C.nch && bin < L) hist[(size_t)c * chan_stride + (size_t)b * L + bin] = (float)x;
        }
        __syncthreads();
    }
}

template <class K>
static unsigned resident_grid(K kernel, int block)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return (unsigned)(per_sm * sm_count());
}

// `dense`: hundreds of entries per segment on average -> one CTA per segment (see reduce_cta_kernel)
template <int MODEL, bool NEED_POS, int NSUB, int MAXC, int WARPS>
static void launch_reduce_cta(cudaStream_t st, const Source &S, const ChanSpec &C, const uint32_t *keys, const uint32_t *vals,
                              const uint32_t *b, const uint32_t *e, uint32_t nseg, uint32_t L, float *dst, size_t stride)
{
    static const int forced_grid = getenv("RVB_REDUCE_GRID") ? atoi(getenv("RVB_REDUCE_GRID")) : 0;
    auto k = reduce_cta_kernel<MODEL, NEED_POS, NSUB, MAXC, WARPS>;
    // two CTAs per SM: 296 segments in flight measured best on C2 (2.54 ms; 444: 2.72, 148: 3.86 --
    // profiles/r1d_sweep_reduce.log): enough loads in flight, smallest window of bins
    unsigned grid = forced_grid > 0 ? (unsigned)forced_grid : std::min(resident_grid(k, 32 * WARPS), 2u * (unsigned)sm_count());
    k<<<std::min(nseg, grid), 32 * WARPS, 0, st>>>(S, C, keys, vals, b, e, nseg, L, dst, stride);
}
template <int MODEL, bool NEED_POS>
static void launch_reduce(int nsub, unsigned grid, bool dense, cudaStream_t st, const Source &S, const ChanSpec &C, const uint32_t *keys,
                          const uint32_t *vals, const uint32_t *b, const uint32_t *e, uint32_t nseg, uint32_t L, float *dst, size_t stride)
{
    static const int warps = getenv("RVB_REDUCE_WARPS") ? atoi(getenv("RVB_REDUCE_WARPS")) : 8;
    if (dense && warps == 16) {
        if (nsub == 4) launch_reduce_cta<MODEL, NEED_POS, 4, 1, 16>(st, S, C, keys, vals, b, e, nseg, L, dst, stride);
        else if (nsub == 2) launch_reduce_cta<MODEL, NEED_POS, 2, 2, 16>(st, S, C, keys, vals, b, e, nseg, L, dst, stride);
        else launch_reduce_cta<MODEL, NEED_POS, 1, RVB_MAX_GROUP, 16>(st, S, C, keys, vals, b, e, nseg, L, dst, stride);
        return;
    }
    if (dense) {
        if (nsub == 4) launch_reduce_cta<MODEL, NEED_POS, 4, 1, 8>(st, S, C, keys, vals, b, e, nseg, L, dst, stride);
        else if (nsub == 2) launch_reduce_cta<MODEL, NEED_POS, 2, 2, 8>(st, S, C, keys, vals, b, e, nseg, L, dst, stride);
        else launch_reduce_cta<MODEL, NEED_POS, 1, RVB_MAX_GROUP, 8>(st, S, C, keys, vals, b, e, nseg, L, dst, stride);
        return;
    }
    if (nsub == 4) reduce_kernel<MODEL, NEED_POS, 4, 1><<<grid, 256, 0, st>>>(S, C, keys, vals, b, e, nseg, L, dst, stride);
    else if (nsub == 2) reduce_kernel<MODEL, NEED_POS, 2, 2><<<grid, 256, 0, st>>>(S, C, keys, vals, b, e, nseg, L, dst, stride);
    else reduce_kernel<MODEL, NEED_POS, 1, RVB_MAX_GROUP><<<grid, 256, 0, st>>>(S, C, keys, vals, b, e, nseg, L, dst, stride);
}

// ---- band filters (filters.cpp:156-191): TDF-II biquad, double state, float store per sample ----
struct BiquadCoef { double b0, b1, b2, a1, a2; };
struct RowFilter { int npass; BiquadCoef q[2]; };   // LR: q[0] lopass.twopass, q[1] hipass.twopass

#define RVB_IIR_THREADS 1024
#define RVB_IIR_CHUNK 16
#define RVB_IIR_TILE (RVB_IIR_THREADS * RVB_IIR_CHUNK)
#define RVB_IIR_WARPS (RVB_IIR_THREADS / 32)

// Biquad::onepass (filters.cpp:156-168) over one chunk held in shared memory.
// STORE = false: only the end state is wanted.
template <bool STORE>
__device__ __forceinline__ void biquad_chunk(const BiquadCoef &q, float *x, double &z1, double &z2)
{
#pragma unroll
    for (int k = 0; k < RVB_IIR_CHUNK; ++k) {
        const double in = (double)x[k];
        const double out = __dadd_rn(__dmul_rn(in, q.b0), z1);
        z1 = __dsub_rn(__dadd_rn(__dmul_rn(in, q.b1), z2), __dmul_rn(q.a1, out));
        z2 = __dsub_rn(__dmul_rn(in, q.b2), __dmul_rn(q.a2, out));
        if (STORE) x[k] = (float)out;
    }
}

struct M2 { double a, b, c, d; };   // [[a, b], [c, d]]
__device__ __forceinline__ M2 m2_mul(const M2 &x, const M2 &y)
{
    return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}

// A thread-block CLUSTER of up to 8 CTAs (1024 threads each) per (channel, band) row.  The
// recurrence is linear in its (z1, z2) state, so a sweep is split twice:
//
//   tiles   16384 samples each, dealt round-robin to the CTAs of the cluster;
//   chunks  inside a tile, 1024 chunks of 16 samples, one per thread.
//
// With A the homogeneous transition [[-a1, 1], [-a2, 0]] (obtained by running the recurrence itself on
// the unit states), M = A^16 carries a state across a chunk, W = M^32 across a warp's 32 chunks and
// M_T = W^32 across a tile.  One sweep:
//
//   phase A  every thread runs the reference recurrence over its chunk from a ZERO state and keeps the
//            end state e_j.  The states the chunks WOULD start from if the tile started from zero,
//            Z_j = sum_{i<j} M^(j-1-i) e_i, come out of two shuffle scans (Kogge-Stone with M^1..M^16
//            over the lanes of a warp, then with W^1..W^16 over the 32 warp totals); thread j keeps
//            its part of Z_j in shared memory, the tile's zero-state end goes to global memory;
//   barrier  (cluster) the tile ends of the whole row are visible; thread 0 of every CTA folds them
//            with M_T into the true start state T of each of its own tiles (a few dozen 2x2 steps);
//   phase B  chunk j of a tile starts from S_j = M^j T + Z_j; every thread re-runs the reference
//            recurrence from S_j and stores (float)out.
//
// So every sample is run through the recurrence twice (once for the end states, once for real), the
// cross-chunk work is 10 shuffle steps instead of serial loops, and the tiles of a row run on up to 8
// SMs.  Up to RVB_IIR_SLOTS tiles per CTA keep their Z_j between the phases; longer rows go round by
// round (8 CTAs x 8 slots x 16384 samples = 23 s of audio per round), the fold carrying on.
// Phase B is the reference's own arithmetic; only the start states are associated differently
// (double, ~1e-16 relative), so the float outputs equal the serial filter's except for vanishingly
// rare last-bit rounding flips.  `twopass` = forward sweep then backward sweep (filters.cpp:185-191).
#define RVB_IIR_SLOTS 8

__device__ __forceinline__ void cluster_sync_all()
{
    // release/acquire at cluster scope: global writes of the other CTAs of the cluster are visible after it
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

// Inclusive scan over the lanes of a warp of the recurrence I_j = P I_{j-1} + e_j (I_{-1} = 0);
// pows[k] = P^(2^k).  In: e_j, out: I_j.
__device__ __forceinline__ void warp_scan_states(const M2 *pows, uint32_t lane, double &a, double &b)
{
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double pa = __shfl_up_sync(0xffffffffu, a, 1u << k), pb = __shfl_up_sync(0xffffffffu, b, 1u << k);
        if (lane >= (1u << k)) {
            const M2 M = pows[k];
            a += M.a * pa + M.b * pb;
            b += M.c * pa + M.d * pb;
        }
    }
}

// One tile of `tl` samples (a multiple of the chunk length) into the padded shared-memory rows; all 16
// loads of a thread are issued before the first store (an L2 round trip each: the previous sweep's
// samples were stored by other CTAs of the cluster).  Samples past the tile or the row read as zero
// and never feed a stored sample.
__device__ __forceinline__ void iir_load_tile(float *tile, const float *x, uint32_t n, uint32_t t0, uint32_t tl, int dirn, uint32_t tid)
{
    float v[RVB_IIR_CHUNK];
#pragma unroll
    for (int j = 0; j < RVB_IIR_CHUNK; ++j) {
        const uint32_t k = tid + j * RVB_IIR_THREADS, i = t0 + k;
        v[j] = 0.0f;
        if (k < tl && i < n) v[j] = dirn == 0 ? __ldcg(x + i) : __ldcg(x + (n - 1 - i));
    }
    float *dst = tile + (tid / RVB_IIR_CHUNK) * (RVB_IIR_CHUNK + 1) + (tid % RVB_IIR_CHUNK);
#pragma unroll
    for (int j = 0; j < RVB_IIR_CHUNK; ++j) dst[j * (RVB_IIR_THREADS / RVB_IIR_CHUNK) * (RVB_IIR_CHUNK + 1)] = v[j];
}

__global__ void __launch_bounds__(RVB_IIR_THREADS) iir_kernel(float *data, const uint32_t *row_off, const uint32_t *row_len,
                                                              const RowFilter *filters, int mode /* 1 onepass, 2 twopass */,
                                                              double *tile_ends /* [row][max_tiles][2] */, uint32_t max_tiles)
{
    extern __shared__ __align__(16) unsigned char iir_smem[];
    float *tile = reinterpret_cast<float *>(iir_smem);                        // RVB_IIR_THREADS * (RVB_IIR_CHUNK + 1), padded rows
    double *Z1 = reinterpret_cast<double *>(iir_smem + (size_t)RVB_IIR_THREADS * (RVB_IIR_CHUNK + 1) * sizeof(float));
    double *Z2 = Z1 + RVB_IIR_SLOTS * RVB_IIR_THREADS;                          // [slot][thread]
    __shared__ double w1[RVB_IIR_WARPS], w2[RVB_IIR_WARPS];       // warp totals (phase A), warp start states (phase B)
    __shared__ double Y1[RVB_IIR_SLOTS][RVB_IIR_WARPS], Y2[RVB_IIR_SLOTS][RVB_IIR_WARPS];   // zero-start state of every warp of a tile
    __shared__ double T1[RVB_IIR_SLOTS], T2[RVB_IIR_SLOTS];       // true start state of this CTA's tiles
    __shared__ double E1[8 * RVB_IIR_SLOTS], E2[8 * RVB_IIR_SLOTS];   // the round's tile ends
    __shared__ M2 pw[32];                                          // M^lane
    __shared__ M2 pwW[32];                                         // W^lane
    __shared__ M2 powM[5], powW[5];                                // M^(2^k), W^(2^k)
    __shared__ M2 MT;                                              // W^32 = A^RVB_IIR_TILE
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    const uint32_t row = blockIdx.x / csize, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *x = data + row_off[row];
    const uint32_t n = row_len[row];
    // equal tiles, the same number for every CTA of the cluster: tl samples (nch chunks) each
    const uint32_t per_cta = (n + csize * RVB_IIR_TILE - 1) / (csize * RVB_IIR_TILE);
    const uint32_t ntiles = csize * per_cta;
    const uint32_t tl = ntiles ? ((n + ntiles - 1) / ntiles + RVB_IIR_CHUNK - 1) / RVB_IIR_CHUNK * RVB_IIR_CHUNK : 0u;
    const uint32_t nch = tl / RVB_IIR_CHUNK;   // <= RVB_IIR_THREADS
    double *ends = tile_ends + (size_t)row * max_tiles * 2;
    const RowFilter F = filters[row];
    float *mine = tile + tid * (RVB_IIR_CHUNK + 1);
    const uint32_t round_tiles = csize * RVB_IIR_SLOTS;
    for (int p = 0; p < F.npass; ++p) {
        const BiquadCoef q = F.q[p];
        __syncthreads();
        if (tid < 32) {
            // M = A^CHUNK from the homogeneous response to the unit states, then its powers
            M2 Mc;
            {
                double c0z1 = 1.0, c0z2 = 0.0, c1z1 = 0.0, c1z2 = 1.0;
                for (int k = 0; k < RVB_IIR_CHUNK; ++k) {
                    double o = c0z1; c0z1 = c0z2 - q.a1 * o; c0z2 = -(q.a2 * o);
                    o = c1z1; c1z1 = c1z2 - q.a1 * o; c1z2 = -(q.a2 * o);
                }
                Mc = M2{c0z1, c1z1, c0z2, c1z2};
            }
            M2 P = M2{1.0, 0.0, 0.0, 1.0};
            for (uint32_t k = 0; k < lane; ++k) P = m2_mul(Mc, P);
            pw[lane] = P;
            if (lane == 31) pwW[1] = m2_mul(Mc, P);                 // W = M^32
            __syncwarp();
            const M2 W = pwW[1];
            __syncwarp();
            P = M2{1.0, 0.0, 0.0, 1.0};
            for (uint32_t k = 0; k < lane; ++k) P = m2_mul(W, P);
            pwW[lane] = P;
            const M2 W32 = m2_mul(W, P);                            // lane 31: W^32
            __syncwarp();
            if (lane == 31) MT = nch == RVB_IIR_THREADS ? W32 : m2_mul(pwW[nch >> 5], pw[nch & 31]);   // M^nch: across one tile
            if (lane < 5) { powM[lane] = pw[1u << lane]; powW[lane] = pwW[1u << lane]; }
        }
        __syncthreads();
        for (int dirn = 0; dirn < mode; ++dirn) {
            double fold_a = 0.0, fold_b = 0.0;   // thread 0: state at the start of the current round, folded onwards
            uint32_t in_smem = 0xffffffffu;       // the tile whose INPUT samples the shared-memory tile holds
            for (uint32_t r0 = 0; r0 < ntiles; r0 += round_tiles) {
                const uint32_t rtiles = min(ntiles - r0, round_tiles);
                // ---- phase A: zero-state chunk ends, the two scans ------------------------------
                for (uint32_t slot = 0; slot < RVB_IIR_SLOTS; ++slot) {
                    const uint32_t t = r0 + slot * csize + crank;
                    if (t >= ntiles) break;
                    const uint32_t t0 = t * tl;
                    iir_load_tile(tile, x, n, t0, tl, dirn, tid);
                    in_smem = t;
                    __syncthreads();
                    double a = 0.0, b = 0.0;
                    biquad_chunk<false>(q, mine, a, b);
                    warp_scan_states(powM, lane, a, b);
                    {
                        double za = __shfl_up_sync(0xffffffffu, a, 1), zb = __shfl_up_sync(0xffffffffu, b, 1);
                        if (lane == 0) { za = 0.0; zb = 0.0; }
                        Z1[slot * RVB_IIR_THREADS + tid] = za; Z2[slot * RVB_IIR_THREADS + tid] = zb;
                    }
                    if (lane == 31) { w1[warp] = a; w2[warp] = b; }
                    __syncthreads();
                    if (warp == 0) {
                        a = w1[lane]; b = w2[lane];
                        warp_scan_states(powW, lane, a, b);
                        double ya = __shfl_up_sync(0xffffffffu, a, 1), yb = __shfl_up_sync(0xffffffffu, b, 1);
                        if (lane == 0) { ya = 0.0; yb = 0.0; }
                        Y1[slot][lane] = ya; Y2[slot][lane] = yb;
                        if (lane == 31 && nch == RVB_IIR_THREADS) { __stcg(ends + 2 * t, a); __stcg(ends + 2 * t + 1, b); }
                    }
                    __syncthreads();   // w1 / w2 and (with several slots) the tile are reused
                    if (tid == nch) {
                        // the tile's zero-start end = the state chunk `nch` would start from
                        const M2 P = pw[lane];
                        const double a0 = Y1[slot][warp], b0 = Y2[slot][warp];
                        __stcg(ends + 2 * t, P.a * a0 + P.b * b0 + Z1[slot * RVB_IIR_THREADS + tid]);
                        __stcg(ends + 2 * t + 1, P.c * a0 + P.d * b0 + Z2[slot * RVB_IIR_THREADS + tid]);
                    }
                }
                // ---- the tile ends of the whole cluster -> start states of this CTA's tiles ------
                if (csize > 1) cluster_sync_all(); else __syncthreads();
                if (tid < rtiles) { E1[tid] = __ldcg(ends + 2 * (r0 + tid)); E2[tid] = __ldcg(ends + 2 * (r0 + tid) + 1); }
                __syncthreads();
                if (tid == 0) {
                    const M2 M = MT;
                    for (uint32_t k = 0; k < rtiles; ++k) {
                        if (k % csize == crank) { T1[k / csize] = fold_a; T2[k / csize] = fold_b; }
                        const double na = M.a * fold_a + M.b * fold_b + E1[k], nb = M.c * fold_a + M.d * fold_b + E2[k];
                        fold_a = na; fold_b = nb;
                    }
                }
                __syncthreads();
                // ---- phase B: the real thing -----------------------------------------------------
                // (last slot first: its samples are still in shared memory from phase A)
                for (int slot = RVB_IIR_SLOTS - 1; slot >= 0; --slot) {
                    const uint32_t t = r0 + (uint32_t)slot * csize + crank;
                    if (t >= ntiles) continue;
                    const uint32_t t0 = t * tl;
                    if (in_smem != t) iir_load_tile(tile, x, n, t0, tl, dirn, tid);
                    if (warp == 0) {   // start state of every warp: W^warp T + Y_warp
                        const M2 P = pwW[lane];
                        const double ta = T1[slot], tb = T2[slot];
                        w1[lane] = P.a * ta + P.b * tb + Y1[slot][lane];
                        w2[lane] = P.c * ta + P.d * tb + Y2[slot][lane];
                    }
                    __syncthreads();
                    double z1, z2;
                    {                  // S_j = M^lane * (warp start) + Z_j
                        const M2 P = pw[lane];
                        const double a = w1[warp], b = w2[warp];
                        z1 = P.a * a + P.b * b + Z1[slot * RVB_IIR_THREADS + tid];
                        z2 = P.c * a + P.d * b + Z2[slot * RVB_IIR_THREADS + tid];
                    }
                    biquad_chunk<true>(q, mine, z1, z2);
                    in_smem = 0xffffffffu;
                    __syncthreads();
                    {
                        const float *src = tile + (tid / RVB_IIR_CHUNK) * (RVB_IIR_CHUNK + 1) + (tid % RVB_IIR_CHUNK);
#pragma unroll
                        for (int j = 0; j < RVB_IIR_CHUNK; ++j) {
                            const uint32_t k = tid + j * RVB_IIR_THREADS, i = t0 + k;
                            if (k < tl && i < n) {
                                const float v = src[j * (RVB_IIR_THREADS / RVB_IIR_CHUNK) * (RVB_IIR_CHUNK + 1)];
                                if (dirn == 0) x[i] = v; else x[n - 1 - i] = v;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            // the next sweep reads samples the other CTAs stored
            if (csize > 1) cluster_sync_all(); else __syncthreads();
        }
    }
}

// mixdown (rayverb.cpp:80-92) + max |x| (generic_functions.h:8-24)
__global__ void mixdown_kernel(const float *bands, uint32_t L, float *out, int *maxabs_bits)
{
    float m = 0.0f;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) {
        float s = 0.0f;
#pragma unroll
        for (int b = 0; b < RVB_NBANDS; ++b) s = __fadd_rn(s, bands[(size_t)b * L + k]);
        out[k] = s;
        m = fmaxf(m, fabsf(s));
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxabs_bits, __float_as_int(m));   // m >= 0 (or NaN > inf bits: propagates)
}

// normalize (generic_functions.h:56-60): x * (float)(1.0 / max|x|), the maximum read on the device
__global__ void normalize_kernel(float *x, uint32_t L, const int *maxabs_bits)
{
    const float f = (float)(1.0 / (double)__int_as_float(*maxabs_bits));
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) x[k] = __fmul_rn(x[k], f);
}

__global__ void scale_kernel(float *x, uint32_t L, float f)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) x[k] = __fmul_rn(x[k], f);
}

// trimTail (rayverb.cpp:96-122): index of the last |x| >= minVol
__global__ void last_loud_kernel(const float *x, uint32_t L, float minVol, int *last)
{
    int l = -1;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x)
        if (fabsf(x[k]) >= minVol) l = (int)k;
    for (int o = 16; o; o >>= 1) l = max(l, __shfl_xor_sync(0xffffffffu, l, o));
    if ((threadIdx.x & 31) == 0 && l >= 0) atomicMax(last, l);
}

__global__ void aos_to_soa_kernel(const rvb_attenuated_impulse *in, uint32_t n, float4 *vol, float4 *pt, float *tkey)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *s = reinterpret_cast<const float4 *>(in + i);
    vol[2 * (size_t)i] = s[0]; vol[2 * (size_t)i + 1] = s[1];
    float t = s[2].x;
    pt[i] = make_float4(0.f, 0.f, 0.f, t);
    tkey[i] = t;   // flattenImpulses looks at every impulse, zero volume or not (rayverb.cpp:55-56,67)
}

__global__ void normalize_dirs_kernel(const float4 *in, int n, float4 *out)
{
    int i = threadIdx.x;
    if (i < n) { f3 v = normalize(mk3(in[i])); out[i] = make_float4(v.x, v.y, v.z, 0.f); }
}
__global__ void ear_kernel(f3 pointing, f3 up, f3 mic, float4 *out)
{
    int ch = threadIdx.x;
    if (ch < 2) { f3 e = hrtf_ear(pointing, up, mic, (unsigned long long)ch); out[ch] = make_float4(e.x, e.y, e.z, 0.f); }
}

static thread_local rvb_render_stats g_stats;

// Stage timings without host synchronisation inside the pipeline: a lap records two events on the
// stream and is resolved (cudaEventElapsedTime) once the caller has synchronised anyway.
struct Lap { cudaEvent_t a, b; float *acc; };
static thread_local std::vector<Lap> g_laps;
struct Timer {
    cudaEvent_t a, b;
    cudaStream_t st;
    explicit Timer(cudaStream_t s) : st(s) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    // accumulate this lap into *acc when the laps are resolved
    void stop_into(float *acc) { cudaEventRecord(b, st); g_laps.push_back(Lap{a, b, acc}); }
};
static void resolve_laps()
{
    for (const Lap &l : g_laps) {
        float ms = 0.f;
        if (cudaEventSynchronize(l.b) == cudaSuccess && cudaEventElapsedTime(&ms, l.a, l.b) == cudaSuccess) *l.acc += ms;
        cudaEventDestroy(l.a); cudaEventDestroy(l.b);
    }
    g_laps.clear();
}

static int pinned_stage(float **p, size_t n) { return pinned_scratch(reinterpret_cast<void **>(p), n * sizeof(float)); }

static int grid_for(size_t n, int block = 256)
{
    size_t g = (n + block - 1) / block;
    return (int)std::min<size_t>(std::max<size_t>(g, 1), 148 * 16);
}

// which impulses take part (cmd/main.cpp:255-269)
static int make_source(const rvb_trace *tr, const rvb_render_config *cfg, Source &S)
{
    S = Source{};
    const bool diffuse = cfg->output_mode != RVB_OUTPUT_IMAGE_ONLY;
    const bool images = cfg->output_mode != RVB_OUTPUT_DIFFUSE_ONLY;
    if (diffuse) {
        size_t n = slot_count(tr->nrays, tr->nrefl);   // includes the unused (skipped) slots of a partial last group
        S.vol = tr->d_vol; S.pt = tr->d_pt; S.tkey = tr->d_tkey; S.n_diffuse = (uint32_t)n;
        S.have_ext = tr->has_time_ext; S.ext_min_nz = tr->diffuse_tmin_nz; S.ext_max = tr->diffuse_tmax;
    }
    if (images) {
        size_t skip = (cfg->remove_direct && tr->first_is_direct) ? 1 : 0;
        S.ivolA = tr->d_img_volA + skip; S.ivolB = tr->d_img_volB + skip; S.ipt = tr->d_img_pt + skip;
        S.n_images = (uint32_t)(tr->n_images - skip);
    }
    if ((size_t)S.n_diffuse + S.n_images >= 0xffffffffull) return fail(RVB_ERR_UNSUPPORTED, "more than 2^32 impulses in one render");
    return RVB_OK;
}

struct ChannelSetup {
    int model = 0;
    size_t nch = 0;
    std::vector<ChanSpec> passes;          // one entry per (sort, reduce) pass
    std::vector<size_t> pass_first_channel;
    std::vector<int> pass_key_group;       // passes with the same key group share one sort
    float4 *d_table[2] = {nullptr, nullptr};
    ~ChannelSetup() { dev_free(d_table[0]); dev_free(d_table[1]); }
};

static int make_channels(const float mic[3], const rvb_render_config *cfg, ChannelSetup &cs)
{
    cudaStream_t st = lib_stream();
    cs.model = cfg->model;
    const f3 m = {mic[0], mic[1], mic[2]};
    if (cfg->model == RVB_MODEL_SPEAKER) {
        if (!cfg->speakers || cfg->nspeakers == 0) return fail(RVB_ERR_INVALID, "speaker model needs at least one speaker");
        cs.nch = cfg->nspeakers;
        // normalize(speaker.direction) with the device's arithmetic (kernel.cpp:511)
        std::vector<float4> h(cs.nch);
        for (size_t c = 0; c < cs.nch; ++c) h[c] = make_float4(cfg->speakers[c].direction.s[0], cfg->speakers[c].direction.s[1], cfg->speakers[c].direction.s[2], 0.f);
        if (cs.nch > 1024) return fail(RVB_ERR_UNSUPPORTED, "more than 1024 speakers");
        float4 *d = nullptr;
        RVB_TRY(dev_alloc(&d, 2 * cs.nch));
        RVB_CUDA(cudaMemcpyAsync(d, h.data(), cs.nch * sizeof(float4), cudaMemcpyHostToDevice, st));
        note_launch(), normalize_dirs_kernel<<<1, 1024, 0, st>>>(d, (int)cs.nch, d + cs.nch);
        RVB_TRY(read_back(h.data(), d + cs.nch, cs.nch * sizeof(float4)));
        dev_free(d);
        for (size_t c0 = 0; c0 < cs.nch; c0 += RVB_MAX_GROUP) {
            ChanSpec C{};
            C.model = 0; C.mic = m; C.nch = (int)std::min<size_t>(RVB_MAX_GROUP, cs.nch - c0);
            for (int c = 0; c < C.nch; ++c) { C.sdir[c] = mk3(h[c0 + c]); C.coef[c] = cfg->speakers[c0 + c].coefficient; }
            cs.passes.push_back(C); cs.pass_first_channel.push_back(c0); cs.pass_key_group.push_back(0);
        }
    } else if (cfg->model == RVB_MODEL_HRTF) {
        if (!cfg->hrtf_table) return fail(RVB_ERR_INVALID, "hrtf model needs a table");
        cs.nch = 2;
        const size_t cells = 360 * 180;
        float4 *d_ear = nullptr, h_ear[2];
        RVB_TRY(dev_alloc(&d_ear, 2));
        const f3 f = {cfg->facing[0], cfg->facing[1], cfg->facing[2]}, u = {cfg->up[0], cfg->up[1], cfg->up[2]};
        note_launch(), ear_kernel<<<1, 2, 0, st>>>(f, u, m, d_ear);
        RVB_TRY(read_back(h_ear, d_ear, sizeof h_ear));
        for (int ch = 0; ch < 2; ++ch) {
            RVB_TRY(dev_alloc(&cs.d_table[ch], 2 * (cells + 180)));
            RVB_CUDA(cudaMemsetAsync(cs.d_table[ch] + 2 * cells, 0, 2 * 180 * sizeof(float4), st));
            RVB_CUDA(cudaMemcpyAsync(cs.d_table[ch], cfg->hrtf_table + (size_t)ch * cells, cells * sizeof(rvb_float8), cudaMemcpyHostToDevice, st));
        }
        RVB_CUDA(cudaStreamSynchronize(st));
        dev_free(d_ear);
        for (int ch = 0; ch < 2; ++ch) {
            ChanSpec C{};
            C.model = 1; C.mic = m; C.nch = 1; C.pointing = f; C.up = u; C.ear = mk3(h_ear[ch]); C.table = cs.d_table[ch]; C.channel = ch;
            cs.passes.push_back(C); cs.pass_first_channel.push_back(ch); cs.pass_key_group.push_back(ch);
        }
    } else {
        return fail(RVB_ERR_INVALID, "unknown attenuation model");
    }
    return RVB_OK;
}

static int times_for_spec(const Source &S_all, const ChanSpec &C, float &mn, float &mx)
{
    cudaStream_t st = lib_stream();
    // speaker / pass-through model: the diffuse extrema came with the trace; only the (few) image
    // sources are scanned here
    const bool use_ext = C.model != 1 && S_all.have_ext && S_all.n_diffuse > 0;
    Source S = S_all;
    if (use_ext) { S.n_diffuse = 0; S.vol = nullptr; S.pt = nullptr; S.tkey = nullptr; }
    int *d = nullptr, h[2];
    RVB_TRY(dev_alloc(&d, 2));
    const int init[2] = {0x7f800000, 0};   // ordered(+inf), ordered(0)
    RVB_CUDA(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, st));
    if (S.n_diffuse + S.n_images) note_launch(), times_kernel<<<grid_for(S.n_diffuse + S.n_images), 256, 0, st>>>(S, C, d);
    RVB_TRY(read_back(h, d, sizeof h));
    dev_free(d);
    mn = unordered(h[0]); mx = unordered(h[1]);
    if (use_ext) { mn = std::min(mn, S_all.ext_min_nz); mx = std::max(mx, S_all.ext_max); }
    if (std::isinf(mn)) mn = 0.0f;   // findPredelay: no non-zero time
    return RVB_OK;
}

// The flatten's sort (see FlattenIn): on return d_vals[*cur] holds the impulse ids of the valid impulses grouped
// by segment, segment g occupying [d_begin[g], d_end[g]) (begin >= end: empty).
#ifndef RVB_SORT_THREADS
#define RVB_SORT_THREADS 256
#endif
template <class KeyT>
static int flatten_sort_keyed(const Source &S, const ChanSpec &C, float predelay, bool fix, float sr, uint32_t L, uint32_t shift,
                              const rsort::PassPlan &pl, uint32_t n, uint32_t *const d_vals[2], uint32_t *d_begin, uint32_t *d_end, int *cur,
                              cudaStream_t st)
{
    KeyT *d_keys[2] = {nullptr, nullptr};
    uint32_t *d_stored = nullptr;
    if (pl.npasses >= 2) RVB_TRY(dev_alloc(&d_keys[0], (size_t)n));
    if (pl.npasses >= 3) RVB_TRY(dev_alloc(&d_keys[1], (size_t)n));
    float *d_img_tkey = nullptr;
    if (S.n_images) {
        RVB_TRY(dev_alloc(&d_img_tkey, (size_t)S.n_images));
        note_launch(), image_tkeys_kernel<<<(S.n_images + 255) / 256, 256, 0, st>>>(S.ivolA, S.ivolB, S.ipt, S.n_images, d_img_tkey);
    }
    int rc;
    if (C.model == 1) {
        // HRTF: the time depends on the impulse position; evaluate it once (histogram sweep) and keep the sample indices
        RVB_TRY(dev_alloc(&d_stored, (size_t)n));
        FlattenIn<KeyT, FLATTEN_TIME_POS> hin{S.tkey, d_img_tkey, S.pt, S.ipt, nullptr, S.n_diffuse, C.mic, C.ear, predelay, sr, L, shift, fix};
        FlattenIn<KeyT, FLATTEN_STORED> first{nullptr, nullptr, nullptr, nullptr, d_stored, S.n_diffuse, C.mic, C.ear, predelay, sr, L, shift, fix};
        rc = rsort::sort_pairs<RVB_SORT_THREADS, KeyT>(hin, StoreSampleIndex{d_stored, shift}, first, n, pl, d_keys, d_vals, true, d_begin, d_end, cur,
                                                       nullptr, st);
    } else {
        FlattenIn<KeyT, FLATTEN_TIME> in{S.tkey, d_img_tkey, nullptr, nullptr, nullptr, S.n_diffuse, C.mic, C.ear, predelay, sr, L, shift, fix};
        rc = rsort::sort_pairs<RVB_SORT_THREADS, KeyT>(in, rsort::NoSink{}, in, n, pl, d_keys, d_vals, true, d_begin, d_end, cur, nullptr, st);
    }
    dev_free(d_img_tkey);
    dev_free(d_keys[0]); dev_free(d_keys[1]); dev_free(d_stored);
    return rc;
}

// sort + segmented reduce of one key group into hist rows
static int histogram_passes(const Source &S, const ChannelSetup &cs, float predelay, bool fix, float sr, const size_t *lengths,
                            float *d_hist)
{
    cudaStream_t st = lib_stream();
    const uint32_t n = S.n_diffuse + S.n_images;
    std::vector<size_t> chan_off(cs.nch + 1, 0);
    for (size_t c = 0; c < cs.nch; ++c) chan_off[c + 1] = chan_off[c] + RVB_NBANDS * lengths[c];
    RVB_CUDA(cudaMemsetAsync(d_hist, 0, chan_off[cs.nch] * sizeof(float), st));
    if (n == 0) return RVB_OK;
    uint32_t *d_vals[2] = {nullptr, nullptr}, *d_begin = nullptr, *d_end = nullptr;
    for (int i = 0; i < 2; ++i) RVB_TRY(dev_alloc(&d_vals[i], n));
    int last_group = -1;
    uint32_t L = 0, nseg = 0;
    int cur = 0, shift = 0;
    for (size_t p = 0; p < cs.passes.size(); ++p) {
        const ChanSpec &C = cs.passes[p];
        const size_t c0 = cs.pass_first_channel[p];
        if (cs.model == RVB_MODEL_SPEAKER) {
            for (int c = 1; c < C.nch; ++c)
                if (lengths[c0 + c] != lengths[c0]) return fail(RVB_ERR_INVALID, "speaker channels must share one length");
        }
        if (cs.pass_key_group[p] != last_group) {
            last_group = cs.pass_key_group[p];
            L = (uint32_t)lengths[c0];
            // A segment of the sorted stream holds 2^shift adjacent bins (the reduction has 32 accumulators per lane:
            // sub-bins x channels x bands), which shortens the keys: take the shift that needs the fewest radix passes,
            // then the narrowest digits.  The sub-bin rides in bits 30-31 of the value, so ids must stay below 2^30.
            static const int shift_env = getenv("RVB_FLATTEN_SHIFT") ? atoi(getenv("RVB_FLATTEN_SHIFT")) : -1;
            int max_shift = C.nch == 1 ? 2 : (C.nch == 2 ? 1 : 0);
            if (n >= (1u << 30)) max_shift = 0;
            rsort::PassPlan pl;
            int key_bits = 0;
            shift = -1;
            for (int sft = 0; sft <= max_shift; ++sft) {
                if (shift_env >= 0 && sft != std::min(shift_env, max_shift)) continue;
                const uint32_t ns = ((L - 1) >> sft) + 1;
                int bits = 1;
                while ((1ull << bits) < (unsigned long long)ns) ++bits;   // segment ids in [0, ns)
                const rsort::PassPlan cand = rsort::plan_passes(0, bits);
                if (shift < 0 || cand.npasses < pl.npasses || (cand.npasses == pl.npasses && cand.bits < pl.bits)) {
                    shift = sft; pl = cand; key_bits = bits;
                }
            }
            nseg = ((L - 1) >> shift) + 1;
            dev_free(d_begin); dev_free(d_end);
            RVB_TRY(dev_alloc(&d_begin, (size_t)nseg)); RVB_TRY(dev_alloc(&d_end, (size_t)nseg));
            Timer ts(st);
            RVB_CUDA(cudaMemsetAsync(d_begin, 0xff, (size_t)nseg * sizeof(uint32_t), st));
            RVB_CUDA(cudaMemsetAsync(d_end, 0, (size_t)nseg * sizeof(uint32_t), st));
            if (key_bits <= 16 && pl.bits == 8) RVB_TRY(flatten_sort_keyed<uint16_t>(S, C, predelay, fix, sr, L, (uint32_t)shift, pl, n, d_vals, d_begin, d_end, &cur, st));
            else RVB_TRY(flatten_sort_keyed<uint32_t>(S, C, predelay, fix, sr, L, (uint32_t)shift, pl, n, d_vals, d_begin, d_end, &cur, st));
            ts.stop_into(&g_stats.sort_ms);
        }
        Timer tr_(st);
        const size_t stride = RVB_NBANDS * (size_t)L;
        {
            const unsigned rgrid = (unsigned)sm_count() * 8;
            float *dst = d_hist + chan_off[c0];
            bool need_pos = false;
            for (int c = 0; c < C.nch; ++c) need_pos |= (C.coef[c] != 0.0f);
            note_launch();
            const int nsub = 1 << shift;
            const uint32_t *rkeys = nullptr;   // the values carry the sub-bin (FlattenIn); the sort keeps no key array
            static const int force = getenv("RVB_REDUCE_DENSE") ? atoi(getenv("RVB_REDUCE_DENSE")) : -1;
            const bool dense = force >= 0 ? force != 0 : (size_t)n >= (size_t)256 * nseg;
            if (C.model == 0 && need_pos) launch_reduce<0, true>(nsub, rgrid, dense, st, S, C, rkeys, d_vals[cur], d_begin, d_end, nseg, L, dst, stride);
            else if (C.model == 0) launch_reduce<0, false>(nsub, rgrid, dense, st, S, C, rkeys, d_vals[cur], d_begin, d_end, nseg, L, dst, stride);
            else if (C.model == 1) launch_reduce<1, true>(nsub, rgrid, dense, st, S, C, rkeys, d_vals[cur], d_begin, d_end, nseg, L, dst, stride);
            else launch_reduce<2, false>(nsub, rgrid, dense, st, S, C, rkeys, d_vals[cur], d_begin, d_end, nseg, L, dst, stride);
        }
        RVB_CUDA(cudaGetLastError());
        tr_.stop_into(&g_stats.reduce_ms);
    }
    for (int i = 0; i < 2; ++i) dev_free(d_vals[i]);
    dev_free(d_begin); dev_free(d_end);
    return RVB_OK;
}

// filters.cpp:193-218
static BiquadCoef bandpass_coef(float lo, float hi, float sr)
{
    const double c = std::sqrt(lo * hi);            // float overloads, as in the reference's C++
    const double omega = 2 * M_PI * c / sr;
    const double cs = std::cos(omega), sn = std::sin(omega);
    const double bandwidth = std::log2(hi / lo);    // float overload
    const double Q = sn / (std::log(2) * bandwidth * omega);
    const double alpha = sn * std::sinh(1 / (2 * Q));
    const double a0 = 1 + alpha, nrm = 1 / a0;
    return BiquadCoef{nrm * alpha, nrm * 0, nrm * -alpha, nrm * (-2 * cs), nrm * (1 - alpha)};
}
// filters.cpp:224-260
static double getC(double co, double sr) { const double wcT = M_PI * co / sr; return std::cos(wcT) / std::sin(wcT); }
static void lr_coef(float l, float h, float s, BiquadCoef &lop, BiquadCoef &hip)
{
    {
        const double c = getC(h, s), a0 = c * c + c * std::sqrt(2) + 1;
        lop = BiquadCoef{1 / a0, 2 / a0, 1 / a0, (-2 * (c * c - 1)) / a0, (c * c - c * std::sqrt(2) + 1) / a0};
    }
    {
        const double c = getC(l, s), a0 = c * c + c * std::sqrt(2) + 1;
        hip = BiquadCoef{(c * c) / a0, (-2 * c * c) / a0, (c * c) / a0, (-2 * (c * c - 1)) / a0, (c * c - c * std::sqrt(2) + 1) / a0};
    }
}

// ---- windowed-sinc crossover (filters.cpp:10-154, filters.h:44-147) --------------------------------
// The reference convolves through FFTW with FFT_LENGTH = 29 + inputLength - 1 (circular == linear
// convolution) and an un-normalised c2r, i.e. it computes FFT_LENGTH * (kernel (*) data).  A 29-tap
// kernel needs no FFT on a GPU: this is the same sum evaluated directly, in double.
#define RVB_SINC_TAPS 29

static void sinc_lopass(float sr, float cutoff, int length, float *out)   // filters.cpp:10-70
{
    const double a0 = 7938.0 / 18608.0, a1 = 9240.0 / 18608.0, a2 = 1430.0 / 18608.0;
    const double fc = cutoff / sr;   // float division, promoted (lopassKernel passes cutoff / sr)
    float m = 0;
    for (int i = 0; i != length; ++i) {
        const double offset = i / (length - 1.0);
        const float window = (float)(a0 - a1 * std::cos(2 * M_PI * offset) + a2 * std::cos(4 * M_PI * offset));
        float s = 1;
        if (i != (length - 1) / 2) {
            const double pit = M_PI * (2 * fc * (i - (length - 1) / 2.0));
            s = (float)(std::sin(pit) / pit);
        }
        out[i] = window * s;
        m = std::max(m, std::fabs(out[i]));
    }
    const float f = (float)(1.0 / m);   // normalize(): scaled by the largest magnitude
    for (int i = 0; i != length; ++i) out[i] *= f;
}
static void sinc_bandpass(float sr, float lo, float hi, float *k29)       // filters.cpp:74-124
{
    enum { HALF = 1 + RVB_SINC_TAPS / 2 };
    float lop[HALF], hip[HALF];
    sinc_lopass(sr, hi, HALF, lop);
    sinc_lopass(sr, lo, HALF, hip);
    for (int i = 0; i != HALF; ++i) hip[i] = -hip[i];
    hip[(HALF - 1) / 2] += 1;
    for (int k = 0; k < RVB_SINC_TAPS; ++k) {   // FastConvolution(29): 29 * (lop (*) hip)
        double acc = 0;
        for (int j = 0; j < HALF; ++j) {
            const int i = k - j;
            if (i >= 0 && i < HALF) acc += (double)lop[j] * (double)hip[i];
        }
        k29[k] = (float)(acc * RVB_SINC_TAPS);
    }
}

// out[row][n] = fft_len * sum_j k[band][j] * in[row][n - j]; rows of `in` have their own lengths
__global__ void sinc_conv_kernel(const float *in, const uint32_t *in_off, const uint32_t *in_len, const float *kernels /*[8][29]*/,
                                 uint32_t fft_len, float *out)
{
    const uint32_t row = blockIdx.y;
    const float *x = in + in_off[row];
    const uint32_t len = in_len[row];
    const float *k = kernels + (row % RVB_NBANDS) * RVB_SINC_TAPS;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < fft_len; n += gridDim.x * blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < RVB_SINC_TAPS; ++j)
            if (n >= (uint32_t)j && n - j < len) acc = __dadd_rn(acc, __dmul_rn((double)k[j], (double)x[n - j]));
        out[(size_t)row * fft_len + n] = (float)__dmul_rn(acc, (double)fft_len);
    }
}

// BandpassWindowedSinc over all channels: rows [nch][8][len_c] -> d_out [nch][8][fft_len],
// fft_len = 29 + lengths[0] - 1 (the convolver is sized by the first channel, filters.cpp:280-283).
static int sinc_device(const float *d_in, const size_t *lengths, size_t nch, float sr, float lo_cutoff, float *d_out, size_t fft_len)
{
    cudaStream_t st = lib_stream();
    const float EDGES[9] = {lo_cutoff, 175, 350, 700, 1400, 2800, 5600, 11200, 20000};
    std::vector<float> hk(RVB_NBANDS * RVB_SINC_TAPS);
    for (int b = 0; b < RVB_NBANDS; ++b) sinc_bandpass(sr, EDGES[b], EDGES[b + 1], hk.data() + b * RVB_SINC_TAPS);
    const size_t nrows = nch * RVB_NBANDS;
    std::vector<uint32_t> off(nrows), len(nrows);
    size_t o = 0;
    for (size_t c = 0; c < nch; ++c) {
        if (lengths[c] > fft_len)
            return fail(RVB_ERR_INVALID, "windowed-sinc: a channel longer than the first overflows the reference's FFT buffer");
        for (int b = 0; b < RVB_NBANDS; ++b) { off[c * RVB_NBANDS + b] = (uint32_t)o; len[c * RVB_NBANDS + b] = (uint32_t)lengths[c]; o += lengths[c]; }
    }
    float *d_k = nullptr; uint32_t *d_off = nullptr, *d_len = nullptr;
    RVB_TRY(dev_alloc(&d_k, hk.size())); RVB_TRY(dev_alloc(&d_off, nrows)); RVB_TRY(dev_alloc(&d_len, nrows));
    RVB_CUDA(cudaMemcpyAsync(d_k, hk.data(), hk.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    RVB_CUDA(cudaMemcpyAsync(d_off, off.data(), nrows * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    RVB_CUDA(cudaMemcpyAsync(d_len, len.data(), nrows * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    dim3 grid((unsigned)std::min<size_t>((fft_len + 255) / 256, 1024), (unsigned)nrows);
    note_launch(), sinc_conv_kernel<<<grid, 256, 0, st>>>(d_in, d_off, d_len, d_k, (uint32_t)fft_len, d_out);
    RVB_CUDA(cudaGetLastError());
    RVB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_k); dev_free(d_off); dev_free(d_len);
    return RVB_OK;
}

// RayverbFiltering::filter (filters.cpp:268-306) on device rows [nch][8][len_c]
static int filter_device(int ft, float *d_data, const size_t *lengths, size_t nch, float sr, float lo_cutoff)
{
    if (ft == RVB_FILTER_WINDOWED_SINC) return fail(RVB_ERR_INVALID, "the windowed-sinc filter changes the data length: use rvb_filter_sinc");
    if (ft < 0 || ft > RVB_FILTER_LINKWITZ_RILEY) return fail(RVB_ERR_INVALID, "unknown filter type");
    cudaStream_t st = lib_stream();
    const float EDGES[9] = {lo_cutoff, 175, 350, 700, 1400, 2800, 5600, 11200, 20000};
    const size_t nrows = nch * RVB_NBANDS;
    std::vector<RowFilter> hf(nrows);
    std::vector<uint32_t> off(nrows), len(nrows);
    size_t o = 0;
    for (size_t c = 0; c < nch; ++c)
        for (int b = 0; b < RVB_NBANDS; ++b) {
            RowFilter F{};
            if (ft == RVB_FILTER_LINKWITZ_RILEY) { F.npass = 2; lr_coef(EDGES[b], EDGES[b + 1], sr, F.q[0], F.q[1]); }
            else { F.npass = 1; F.q[0] = bandpass_coef(EDGES[b], EDGES[b + 1], sr); }
            hf[c * RVB_NBANDS + b] = F; off[c * RVB_NBANDS + b] = (uint32_t)o; len[c * RVB_NBANDS + b] = (uint32_t)lengths[c];
            o += lengths[c];
        }
    RowFilter *d_f = nullptr; uint32_t *d_off = nullptr, *d_len = nullptr;
    RVB_TRY(dev_alloc(&d_f, nrows)); RVB_TRY(dev_alloc(&d_off, nrows)); RVB_TRY(dev_alloc(&d_len, nrows));
    RVB_CUDA(cudaMemcpyAsync(d_f, hf.data(), nrows * sizeof(RowFilter), cudaMemcpyHostToDevice, st));
    RVB_CUDA(cudaMemcpyAsync(d_off, off.data(), nrows * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    RVB_CUDA(cudaMemcpyAsync(d_len, len.data(), nrows * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const size_t iir_smem = (size_t)RVB_IIR_THREADS * (RVB_IIR_CHUNK + 1) * sizeof(float) + (size_t)2 * RVB_IIR_SLOTS * RVB_IIR_THREADS * sizeof(double);
    RVB_CUDA(cudaFuncSetAttribute(iir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iir_smem));
    // cluster size: as many CTAs per row as the longest row has tiles, up to the portable maximum of 8
    size_t max_len = 0;
    for (size_t c = 0; c < nch; ++c) max_len = std::max(max_len, lengths[c]);
    const uint32_t whole_tiles = (uint32_t)std::max<size_t>(1, (max_len + RVB_IIR_TILE - 1) / RVB_IIR_TILE);
    unsigned csize = 1;
    while (csize < 8 && csize * 2 <= whole_tiles) csize *= 2;
    // the kernel cuts every row into the same number of (equal) tiles for each CTA of its cluster
    const uint32_t max_tiles = csize * (uint32_t)std::max<size_t>(1, (max_len + (size_t)csize * RVB_IIR_TILE - 1) / ((size_t)csize * RVB_IIR_TILE));
    double *d_ends = nullptr;
    RVB_TRY(dev_alloc(&d_ends, nrows * (size_t)max_tiles * 2));
    {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3((unsigned)nrows * csize, 1, 1);
        lc.blockDim = dim3(RVB_IIR_THREADS, 1, 1);
        lc.dynamicSmemBytes = iir_smem;
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        const int mode = ft == RVB_FILTER_BIQUAD_ONEPASS ? 1 : 2;
        note_launch();
        RVB_CUDA(cudaLaunchKernelEx(&lc, iir_kernel, d_data, (const uint32_t *)d_off, (const uint32_t *)d_len, (const RowFilter *)d_f, mode,
                                    d_ends, max_tiles));
    }
    RVB_CUDA(cudaGetLastError());
    // no synchronisation here: the small blocks are freed in stream order and every caller ends
    // with a device -> host copy + synchronise of its own
    dev_free(d_f); dev_free(d_off); dev_free(d_len); dev_free(d_ends);
    return RVB_OK;
}

// process (rayverb.cpp:125-149) on a device histogram; fills an Ir
static int finish_filtered(const rvb_render_config *cfg, float *d_hist, const size_t *lengths, size_t nch, Ir *ir, bool keep_bands);

static int finish_device(const rvb_render_config *cfg, float *d_hist /* modified */, const size_t *lengths, size_t nch, Ir *ir, bool keep_bands)
{
    cudaStream_t st = lib_stream();
    Timer tf(st);
    if (cfg->filter_type == RVB_FILTER_WINDOWED_SINC && nch > 0) {
        // every channel grows to 29 + lengths[0] - 1 samples
        const size_t fft_len = RVB_SINC_TAPS + lengths[0] - 1;
        float *d_out = nullptr;
        RVB_TRY(dev_alloc(&d_out, nch * RVB_NBANDS * fft_len));
        int rc = sinc_device(d_hist, lengths, nch, cfg->sample_rate, cfg->lo_cutoff, d_out, fft_len);
        tf.stop_into(&g_stats.filter_ms);
        if (rc == RVB_OK) {
            std::vector<size_t> grown(nch, fft_len);
            rc = finish_filtered(cfg, d_out, grown.data(), nch, ir, keep_bands);
        }
        dev_free(d_out);
        return rc;
    }
    RVB_TRY(filter_device(cfg->filter_type, d_hist, lengths, nch, cfg->sample_rate, cfg->lo_cutoff));
    tf.stop_into(&g_stats.filter_ms);
    return finish_filtered(cfg, d_hist, lengths, nch, ir, keep_bands);
}

// mixdown + normalize + scale + trimTail on already filtered rows (rayverb.cpp:136-148)
static int finish_filtered(const rvb_render_config *cfg, float *d_hist, const size_t *lengths, size_t nch, Ir *ir, bool keep_bands)
{
    cudaStream_t st = lib_stream();
    Timer tm(st);
    size_t total = 0;
    for (size_t c = 0; c < nch; ++c) total += lengths[c];
    float *d_mix = nullptr; int *d_red = nullptr;
    RVB_TRY(dev_alloc(&d_mix, total)); RVB_TRY(dev_alloc(&d_red, 2));
    const int init[2] = {0, -1};
    RVB_CUDA(cudaMemcpyAsync(d_red, init, sizeof init, cudaMemcpyHostToDevice, st));
    size_t ho = 0, mo = 0;
    for (size_t c = 0; c < nch; ++c) {
        if (lengths[c]) note_launch(), mixdown_kernel<<<grid_for(lengths[c]), 256, 0, st>>>(d_hist + ho, (uint32_t)lengths[c], d_mix + mo, d_red);
        ho += RVB_NBANDS * lengths[c]; mo += lengths[c];
    }
    // no host round trips until the very end: the normalisation factor is read on the device and
    // the channels are copied out whole, then trimmed on the host
    if (cfg->do_normalize && total) note_launch(), normalize_kernel<<<grid_for(total), 256, 0, st>>>(d_mix, (uint32_t)total, d_red);
    if (cfg->volume_scale != 1) { if (total) note_launch(), scale_kernel<<<grid_for(total), 256, 0, st>>>(d_mix, (uint32_t)total, cfg->volume_scale); }
    if (cfg->do_trim_tail) {
        mo = 0;
        for (size_t c = 0; c < nch; ++c) {
            if (lengths[c]) note_launch(), last_loud_kernel<<<grid_for(lengths[c]), 256, 0, st>>>(d_mix + mo, (uint32_t)lengths[c], 0.00001f, d_red + 1);
            mo += lengths[c];
        }
    }
    // The mixed channels go to the host through a pinned staging buffer: a copy into pageable memory
    // is staged by the driver in small pieces under its context lock, which is slow and stalls the
    // other host threads of a parameter sweep.
    float *h_stage = nullptr;
    RVB_TRY(pinned_stage(&h_stage, total + 2));
    int *h_red = reinterpret_cast<int *>(h_stage + total);
    h_red[0] = 0; h_red[1] = -1;
    RVB_CUDA(cudaMemcpyAsync(h_red, d_red, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (total) RVB_CUDA(cudaMemcpyAsync(h_stage, d_mix, total * sizeof(float), cudaMemcpyDeviceToHost, st));
    ir->channels.assign(nch, {});
    ir->bands.assign(nch, {});
    ir->band_len.assign(nch, 0);
    mo = 0; ho = 0;
    for (size_t c = 0; c < nch; ++c) {
        if (keep_bands) {
            ir->bands[c].resize(RVB_NBANDS * lengths[c]);
            ir->band_len[c] = lengths[c];
            if (lengths[c]) RVB_CUDA(cudaMemcpyAsync(ir->bands[c].data(), d_hist + ho, RVB_NBANDS * lengths[c] * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        mo += lengths[c]; ho += RVB_NBANDS * lengths[c];
    }
    RVB_CUDA(cudaStreamSynchronize(st));
    mo = 0;
    for (size_t c = 0; c < nch; ++c) {
        // trimTail: accumulate(..., 0, max(current, index)) -- the loud sample itself is dropped (Q12);
        // vector::resize zero-extends channels shorter than the trimmed length
        const size_t want = cfg->do_trim_tail ? (size_t)std::max(0, h_red[1]) : lengths[c];
        ir->channels[c].assign(h_stage + mo, h_stage + mo + std::min(want, lengths[c]));
        ir->channels[c].resize(want, 0.0f);
        mo += lengths[c];
    }
    dev_free(d_mix); dev_free(d_red);
    tm.stop_into(&g_stats.finish_ms);
    return RVB_OK;
}

static size_t flatten_length(float maxtime, float predelay, bool fix, float sr)
{
    float t = maxtime;
    if (fix) t = t > predelay ? t - predelay : 0.0f;
    return (size_t)(std::round(t * sr) + 1.0f);   // rayverb.cpp:57, float arithmetic
}

}  // namespace rvb

using namespace rvb;

static thread_local bool g_keep_bands = false;

extern "C" {

int rvb_render_keep_bands(int on)
{
    g_keep_bands = on != 0;
    return RVB_OK;
}

int rvb_render_get_stats(rvb_render_stats *stats)
{
    if (!stats) return fail(RVB_ERR_INVALID, "null argument");
    resolve_laps();
    *stats = g_stats;
    return RVB_OK;
}

int rvb_render_times(rvb_trace *tr, const rvb_render_config *cfg, float *min_nonzero_time, float *max_time, size_t *nchannels)
{
    if (!tr || !cfg || !min_nonzero_time || !max_time || !nchannels) return fail(RVB_ERR_INVALID, "null argument");
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    Source S;
    RVB_TRY(make_source(tr, cfg, S));
    ChannelSetup cs;
    RVB_TRY(make_channels(tr->mic, cfg, cs));
    *nchannels = cs.nch;
    int last_group = -1;
    float mn = 0, mx = 0;
    for (size_t p = 0; p < cs.passes.size(); ++p) {
        if (cs.pass_key_group[p] != last_group) {
            last_group = cs.pass_key_group[p];
            RVB_TRY(times_for_spec(S, cs.passes[p], mn, mx));
        }
        for (int c = 0; c < cs.passes[p].nch; ++c) { min_nonzero_time[cs.pass_first_channel[p] + c] = mn; max_time[cs.pass_first_channel[p] + c] = mx; }
    }
    return RVB_OK;
}

int rvb_render_histogram(rvb_trace *tr, const rvb_render_config *cfg, float predelay, const size_t *lengths, float *d_hist)
{
    if (!tr || !cfg || !lengths || !d_hist) return fail(RVB_ERR_INVALID, "null argument");
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    resolve_laps();   // laps of an earlier call that nobody asked about
    g_stats = rvb_render_stats{};
    Timer tt(lib_stream());
    Source S;
    RVB_TRY(make_source(tr, cfg, S));
    ChannelSetup cs;
    RVB_TRY(make_channels(tr->mic, cfg, cs));
    g_stats.n_impulses = (uint64_t)S.n_diffuse + S.n_images;
    RVB_TRY(histogram_passes(S, cs, predelay, cfg->trim_predelay != 0, cfg->sample_rate, lengths, d_hist));
    tt.stop_into(&g_stats.total_ms);
    RVB_CUDA(cudaStreamSynchronize(lib_stream()));
    resolve_laps();
    return RVB_OK;
}

int rvb_render_finish(const rvb_render_config *cfg, const float *d_hist, const size_t *lengths, size_t nchannels, rvb_ir **out)
{
    if (!cfg || !d_hist || !lengths || !out) return fail(RVB_ERR_INVALID, "null argument");
    *out = nullptr;
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    size_t total = 0;
    for (size_t c = 0; c < nchannels; ++c) total += RVB_NBANDS * lengths[c];
    float *d_work = nullptr;
    RVB_TRY(dev_alloc(&d_work, total));
    RVB_CUDA(cudaMemcpyAsync(d_work, d_hist, total * sizeof(float), cudaMemcpyDeviceToDevice, st));
    rvb_ir *ir = new rvb_ir();
    int rc = finish_device(cfg, d_work, lengths, nchannels, ir, g_keep_bands);
    dev_free(d_work);
    if (rc) { delete ir; return rc; }
    resolve_laps();
    *out = ir;
    return RVB_OK;
}

int rvb_render_ir(rvb_trace *tr, const rvb_render_config *cfg, rvb_ir **out)
{
    if (!tr || !cfg || !out) return fail(RVB_ERR_INVALID, "null argument");
    *out = nullptr;
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    cudaStream_t st = lib_stream();
    resolve_laps();   // laps of an earlier call that nobody asked about
    g_stats = rvb_render_stats{};
    StageTimer t_all("render.total");
    Timer tt(st);
    Source S;
    ChannelSetup cs;
    {
        StageTimer t_s("render.setup");
        RVB_TRY(make_source(tr, cfg, S));
        RVB_TRY(make_channels(tr->mic, cfg, cs));
    }
    g_stats.n_impulses = (uint64_t)S.n_diffuse + S.n_images;
    std::unique_ptr<StageTimer> t_t(new StageTimer("render.times"));
    // per-channel time extrema -> predelay (min over all channels) and flatten lengths
    std::vector<float> mn(cs.nch), mx(cs.nch);
    int last_group = -1;
    float a = 0, b = 0;
    for (size_t p = 0; p < cs.passes.size(); ++p) {
        if (cs.pass_key_group[p] != last_group) { last_group = cs.pass_key_group[p]; RVB_TRY(times_for_spec(S, cs.passes[p], a, b)); }
        for (int c = 0; c < cs.passes[p].nch; ++c) { mn[cs.pass_first_channel[p] + c] = a; mx[cs.pass_first_channel[p] + c] = b; }
    }
    float predelay = 0.0f;
    for (size_t c = 0; c < cs.nch; ++c) {   // findPredelay's fold: zeros are ignored
        if (predelay == 0) predelay = mn[c]; else if (mn[c] != 0) predelay = std::min(predelay, mn[c]);
    }
    const bool fix = cfg->trim_predelay != 0;
    std::vector<size_t> lengths(cs.nch);
    size_t total = 0;
    for (size_t c = 0; c < cs.nch; ++c) { lengths[c] = flatten_length(mx[c], predelay, fix, cfg->sample_rate); total += RVB_NBANDS * lengths[c]; }
    t_t.reset();
    float *d_hist = nullptr;
    RVB_TRY(dev_alloc(&d_hist, total));
    int rc;
    {
        StageTimer t_h("render.histogram");
        rc = histogram_passes(S, cs, predelay, fix, cfg->sample_rate, lengths.data(), d_hist);
    }
    rvb_ir *ir = nullptr;
    if (rc == RVB_OK) {
        StageTimer t_fi("render.finish(filter+mix)");
        ir = new rvb_ir();
        rc = finish_device(cfg, d_hist, lengths.data(), cs.nch, ir, g_keep_bands);
    }
    dev_free(d_hist);
    if (rc) { delete ir; return rc; }
    tt.stop_into(&g_stats.total_ms);
    resolve_laps();   // finish_device ended with a stream synchronisation
    *out = ir;
    return RVB_OK;
}

int rvb_ir_channels(const rvb_ir *ir, size_t *n) { if (!ir || !n) return fail(RVB_ERR_INVALID, "null argument"); *n = ir->channels.size(); return RVB_OK; }
int rvb_ir_length(const rvb_ir *ir, size_t c, size_t *n)
{
    if (!ir || !n || c >= ir->channels.size()) return fail(RVB_ERR_INVALID, "bad channel");
    *n = ir->channels[c].size();
    return RVB_OK;
}
int rvb_ir_copy(const rvb_ir *ir, size_t c, float *out)
{
    if (!ir || c >= ir->channels.size() || (!out && !ir->channels[c].empty())) return fail(RVB_ERR_INVALID, "bad channel");
    std::copy(ir->channels[c].begin(), ir->channels[c].end(), out);
    return RVB_OK;
}
int rvb_ir_band_length(const rvb_ir *ir, size_t c, size_t *n)
{
    if (!ir || !n || c >= ir->band_len.size()) return fail(RVB_ERR_INVALID, "bad channel");
    *n = ir->band_len[c];
    return RVB_OK;
}
int rvb_ir_copy_bands(const rvb_ir *ir, size_t c, float *out)
{
    if (!ir || c >= ir->bands.size() || (!out && !ir->bands[c].empty())) return fail(RVB_ERR_INVALID, "bad channel");
    std::copy(ir->bands[c].begin(), ir->bands[c].end(), out);
    return RVB_OK;
}
int rvb_ir_destroy(rvb_ir *ir) { delete ir; return RVB_OK; }

// ---- host-array post-processing entry points ---------------------------------------------------
int rvb_flatten(const rvb_attenuated_impulse *in, size_t n, float sample_rate, float **out, size_t *length)
{
    if (!out || !length || (n && !in)) return fail(RVB_ERR_INVALID, "null argument");
    *out = nullptr; *length = 0;
    RVB_TRY(ensure_device());
    if (n >= 0xffffffffull) return fail(RVB_ERR_UNSUPPORTED, "too many impulses");
    cudaStream_t st = lib_stream();
    resolve_laps();   // laps of an earlier call that nobody asked about
    g_stats = rvb_render_stats{};
    rvb_attenuated_impulse *d_in = nullptr;
    float4 *vA = nullptr, *pt = nullptr; float *tk = nullptr;
    RVB_TRY(dev_alloc(&d_in, n)); RVB_TRY(dev_alloc(&vA, 2 * n)); RVB_TRY(dev_alloc(&pt, n)); RVB_TRY(dev_alloc(&tk, n));
    if (n) {
        RVB_CUDA(cudaMemcpyAsync(d_in, in, n * sizeof(rvb_attenuated_impulse), cudaMemcpyHostToDevice, st));
        note_launch(), aos_to_soa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, (uint32_t)n, vA, pt, tk);
    }
    Source S{}; S.vol = vA; S.pt = pt; S.tkey = tk; S.n_diffuse = (uint32_t)n;
    ChannelSetup cs; cs.model = 2; cs.nch = 1;
    ChanSpec C{}; C.model = 2; C.nch = 1;
    cs.passes.push_back(C); cs.pass_first_channel.push_back(0); cs.pass_key_group.push_back(0);
    float mn, mx;
    int rc = times_for_spec(S, C, mn, mx);
    size_t L = flatten_length(mx, 0.f, false, sample_rate);
    float *d_hist = nullptr;
    if (rc == RVB_OK) rc = dev_alloc(&d_hist, RVB_NBANDS * L);
    if (rc == RVB_OK) rc = histogram_passes(S, cs, 0.f, false, sample_rate, &L, d_hist);
    float *h = nullptr;
    if (rc == RVB_OK) {
        h = (float *)malloc(sizeof(float) * RVB_NBANDS * L);
        if (!h) rc = fail(RVB_ERR_NOMEM, "host allocation failed");
    }
    if (rc == RVB_OK) {
        cudaMemcpyAsync(h, d_hist, sizeof(float) * RVB_NBANDS * L, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        *out = h; *length = L;
    }
    dev_free(d_hist); dev_free(d_in); dev_free(vA); dev_free(pt); dev_free(tk);
    return rc;
}

int rvb_filter(int filter_type, float *data, size_t nchannels, size_t length, float sample_rate, float lo_cutoff)
{
    if (!data && nchannels * length) return fail(RVB_ERR_INVALID, "null argument");
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    const size_t total = nchannels * RVB_NBANDS * length;
    if (total == 0) return RVB_OK;
    float *d = nullptr;
    RVB_TRY(dev_alloc(&d, total));
    RVB_CUDA(cudaMemcpyAsync(d, data, total * sizeof(float), cudaMemcpyHostToDevice, st));
    std::vector<size_t> lens(nchannels, length);
    int rc = filter_device(filter_type, d, lens.data(), nchannels, sample_rate, lo_cutoff);
    if (rc == RVB_OK) { cudaMemcpyAsync(data, d, total * sizeof(float), cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }
    dev_free(d);
    return rc;
}

int rvb_filter_sinc(const float *in, size_t nchannels, size_t length, float sample_rate, float lo_cutoff, float *out)
{
    if ((!in || !out) && nchannels) return fail(RVB_ERR_INVALID, "null argument");
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    if (nchannels == 0) return RVB_OK;
    const size_t fft_len = RVB_SINC_TAPS + length - 1;
    float *d_in = nullptr, *d_out = nullptr;
    RVB_TRY(dev_alloc(&d_in, std::max<size_t>(1, nchannels * RVB_NBANDS * length)));
    RVB_TRY(dev_alloc(&d_out, nchannels * RVB_NBANDS * fft_len));
    if (length) RVB_CUDA(cudaMemcpyAsync(d_in, in, nchannels * RVB_NBANDS * length * sizeof(float), cudaMemcpyHostToDevice, st));
    std::vector<size_t> lens(nchannels, length);
    int rc = sinc_device(d_in, lens.data(), nchannels, sample_rate, lo_cutoff, d_out, fft_len);
    if (rc == RVB_OK) {
        cudaMemcpyAsync(out, d_out, nchannels * RVB_NBANDS * fft_len * sizeof(float), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
    }
    dev_free(d_in); dev_free(d_out);
    return rc;
}

int rvb_process(int filter_type, float *const *data, const size_t *lengths, size_t nchannels, float sample_rate, int do_normalize,
                float lo_cutoff, int do_trim_tail, float volume_scale, float *const *out, int64_t *out_length)
{
    if (!data || !lengths || !out || !out_length) return fail(RVB_ERR_INVALID, "null argument");
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    size_t total = 0;
    for (size_t c = 0; c < nchannels; ++c) total += RVB_NBANDS * lengths[c];
    float *d = nullptr;
    RVB_TRY(dev_alloc(&d, total));
    size_t o = 0;
    for (size_t c = 0; c < nchannels; ++c) {
        if (lengths[c]) RVB_CUDA(cudaMemcpyAsync(d + o, data[c], RVB_NBANDS * lengths[c] * sizeof(float), cudaMemcpyHostToDevice, st));
        o += RVB_NBANDS * lengths[c];
    }
    rvb_render_config cfg{};
    cfg.filter_type = filter_type; cfg.sample_rate = sample_rate; cfg.lo_cutoff = lo_cutoff; cfg.do_normalize = do_normalize;
    cfg.do_trim_tail = do_trim_tail; cfg.volume_scale = volume_scale;
    Ir ir;
    int rc = finish_device(&cfg, d, lengths, nchannels, &ir, true);
    if (rc == RVB_OK) {
        const bool sinc = filter_type == RVB_FILTER_WINDOWED_SINC;
        *out_length = do_trim_tail ? (int64_t)(nchannels ? ir.channels[0].size() : 0) : -1;
        for (size_t c = 0; c < nchannels; ++c) {
            // like the reference, `data` ends up filtered in place (rayverb.cpp:135) -- except for the
            // windowed sinc, whose rows grow to 29 + lengths[0] - 1 and do not fit back
            if (!sinc) std::copy(ir.bands[c].begin(), ir.bands[c].end(), data[c]);
            const size_t cap = sinc ? RVB_SINC_TAPS + lengths[0] - 1 : lengths[c];
            size_t m = std::min(ir.channels[c].size(), cap);
            std::copy(ir.channels[c].begin(), ir.channels[c].begin() + m, out[c]);
        }
    }
    dev_free(d);
    return rc;
}

void rvb_free(void *p) { free(p); }

}  // extern "C"
