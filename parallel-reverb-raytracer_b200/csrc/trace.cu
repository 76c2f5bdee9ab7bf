// The bounce kernels: B200-native replacement of `kernel raytrace` (rayverb/kernel.cpp:304-503).
//
// Two kernels per trace:
//
//   image phase    reflections 0..8: the part of the reference loop that carries the image-source
//                  branch (kernel.cpp:379-457: mirrored triangles, validation queries, add_image).
//                  Heavy, divergent, 7 % of the reflections.  One thread per ray.
//   steady phase   reflections 9..: nearest hit -> visibility -> emit -> reflect.  Brute-force scenes:
//                  one thread per ray (one WARP per ray when there are few rays and many triangles),
//                  persistent grid-stride over rays, one inlined copy of each query loop so the hot code
//                  stays instruction-cache resident and the lanes of a warp stay in lock step.  BVH
//                  scenes: persistent warps whose lanes are independent state machines drawing work
//                  units -- pieces of rays -- from a queue (trace_bvh_wavefront_kernel below).
//
// The ray state (position, direction, path length, 8-band volume) passes between them through a
// 64-byte record per ray.  The scene is staged once per CTA in shared memory in pair-interleaved
// intersection form (v0, e0, e1 of two triangles component-wise, 144 bytes per four triangles, read
// as warp-wide broadcasts and tested two triangles per packed-FP32 instruction) for the brute-force
// path, or traversed through the GPU-built BVH (bvh.cu) whose nearest-hit result -- including the
// lowest-index tie-break of the reference's strict '<' loop (kernel.cpp:180-188) -- equals the
// brute-force loop.
//
// Output is a bounce-group-major SoA (slot_of, rvb_internal.cuh): the 32 lanes of a warp, which sit
// at the same bounce, write 32 whole 32-byte sectors, and four consecutive reflections of a ray
// share one 128-byte line for the benefit of the histogram reduction.
#include <cmath>
#include <cstdlib>
#include <cstring>


#include "rvb_internal.cuh"
#include "rvb_sort.cuh"

namespace rvb {

struct TraceParams {
    const float4 *verts;      // [3T]
    const float4 *tri;        // [3T] staged (v0,e0,e1)
    const float4 *quad;       // [9*nquads] pair-interleaved form for the brute-force loop
    uint32_t nquads;
    const float4 *nrm;        // [T]
    const rvb_surface *surfaces;
    uint32_t T;
    // BVH
    const BvhNode *nodes;
    const float4 *leaf_quads;
    const uint4 *leaf_ids;
    const float4 *dirs;       // [nrays] shared by all jobs of a batch
    const uint32_t *ray_perm;  // image phase: processing order (rays sorted by direction), or null
    uint32_t nrays, nrefl;    // rays PER JOB
    uint32_t first, last;     // reflections [first, last) handled by this launch
    float air[RVB_NBANDS];
    uint32_t wf_service, wf_leaf_min;          // wavefront kernel: lanes that must wait before `serve` / hold a leaf before the triangle step
    // wavefront kernel, segmented rays (launch_wavefront): a ray's reflections [first, last) are cut into
    // wf_nseg work units of wf_seglen reflections; wf_done[ray] = units of that ray completed so far
    // (RVB_WF_DEAD: the ray left the mesh), wf_done[nrays] = error flag (a unit waited too long)
    uint32_t wf_nseg, wf_seglen;
    unsigned int *wf_done;
    const struct TraceJob *jobs;               // [njobs] device array
    uint32_t njobs;
    uint32_t flags;
};

// One (microphone, source) pair of a trace launch.  A plain trace is a batch of one; a batched trace
// (rvb_trace_run_batch: parameter sweeps, cmd/main.cpp:241-297 run once per pair) traces every pair's
// rays in the same launches -- a CTA (a warp of the wavefront kernel) works for one job at a time and
// keeps that job's record in shared memory.
struct TraceJob {
    float mic[3], src[3];
    uint64_t ray_offset;
    MicMap micmap;            // visibility map around this job's microphone (N == 0: none)
    float4 *vol, *pt;         // vol: 2 x float4 per slot
    float *tkey;
    uint32_t *hit, *count;
    float4 *state;            // [4*nrays] (pos, distance) (dir, alive) volA volB
    float4 *cand_volA, *cand_volB, *cand_pt;   // [nrays*10] image-source candidates
    uint32_t *cand_idx;                        // [nrays*10] triangle+1, 0 = none (pre-zeroed)
    unsigned long long *counters;              // [0] queries [1] tri tests [2] box tests [3] bounces [5] emitted impulses
    int *time_ext;                             // [0] ordered min non-zero, [1] ordered max time of the emitted diffuse impulses
};

struct Counters {
    uint32_t queries = 0, bounces = 0, emitted = 0;
    unsigned long long tri = 0, box = 0;
    // extrema of the emitted (non-skipped) impulse times: what findPredelay / flattenImpulses need
    // (rayverb.h:47-76, rayverb.cpp:54-57), gathered here so the renderer does not re-read 4 B per impulse
    float tmin_nz = INFINITY, tmax = 0.0f;
};

#define RVB_NONE 0xffffffffu
#define RVB_INF __int_as_float(0x7f800000)

// ---- brute force over the shared-memory scene (kernel.cpp:167-192) ---------------------------
// The scene sits in shared memory in the pair-interleaved form of Scene::d_quad: every warp-wide
// LDS.128 is a broadcast, and each packed instruction (FFMA2 ...) tests the ray against TWO
// triangles.  Triangles are visited in index order and the strict '<' keeps the lowest index on
// exact ties, as the reference loop does.
// result of the nearest-hit query that rides along with a visibility query (BruteSmem::visible_and_next)
struct NextHit { bool want; float t; uint32_t prim; };

struct BruteSmem {
    static constexpr bool kFused = true;
    const float4 *s;
    uint32_t T, nquads;

    __device__ __forceinline__ void load(uint32_t q, f3x2 &v0a, f3x2 &e0a, f3x2 &e1a, f3x2 &v0b, f3x2 &e0b, f3x2 &e1b) const
    {
        const float4 *r = s + 9 * q;
        const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5], r6 = r[6], r7 = r[7], r8 = r[8];
        v0a.x = make_float2(r0.x, r0.y); v0a.y = make_float2(r0.z, r0.w); v0a.z = make_float2(r1.x, r1.y);
        e0a.x = make_float2(r1.z, r1.w); e0a.y = make_float2(r2.x, r2.y); e0a.z = make_float2(r2.z, r2.w);
        e1a.x = make_float2(r3.x, r3.y); e1a.y = make_float2(r3.z, r3.w); e1a.z = make_float2(r4.x, r4.y);
        v0b.x = make_float2(r4.z, r4.w); v0b.y = make_float2(r5.x, r5.y); v0b.z = make_float2(r5.z, r5.w);
        e0b.x = make_float2(r6.x, r6.y); e0b.y = make_float2(r6.z, r6.w); e0b.z = make_float2(r7.x, r7.y);
        e1b.x = make_float2(r7.z, r7.w); e1b.y = make_float2(r8.x, r8.y); e1b.z = make_float2(r8.z, r8.w);
    }
    __device__ __forceinline__ bool nearest(f3 pos, f3 dir, float &tbest, uint32_t &ibest, Counters &c) const
    {
        tbest = RVB_INF;
        ibest = RVB_NONE;
#pragma unroll 2
        for (uint32_t q = 0; q < nquads; ++q) {
            f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
            load(q, v0a, e0a, e1a, v0b, e0b, e1b);
            bool ok0, ok1, ok2, ok3;
            const float2 ta = mt_intersect_pair(v0a, e0a, e1a, pos, dir, ok0, ok1);
            const float2 tb = mt_intersect_pair(v0b, e0b, e1b, pos, dir, ok2, ok3);
            if (ok0 && ta.x > RVB_EPSILON && ta.x < tbest) { tbest = ta.x; ibest = 4 * q; }
            if (ok1 && ta.y > RVB_EPSILON && ta.y < tbest) { tbest = ta.y; ibest = 4 * q + 1; }
            if (ok2 && tb.x > RVB_EPSILON && tb.x < tbest) { tbest = tb.x; ibest = 4 * q + 2; }
            if (ok3 && tb.y > RVB_EPSILON && tb.y < tbest) { tbest = tb.y; ibest = 4 * q + 3; }
        }
        c.queries++; c.tri += T;
        return ibest != RVB_NONE;
    }
    // point_intersection (kernel.cpp:274-296): visible unless some hit has EPS < t <= mag
    __device__ __forceinline__ bool occluded_to_mic(f3 pos, f3 dir, float mag, Counters &c) const
    {
        bool blocked = false;
#pragma unroll 2
        for (uint32_t q = 0; q < nquads; ++q) {
            f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
            load(q, v0a, e0a, e1a, v0b, e0b, e1b);
            bool ok0, ok1, ok2, ok3;
            const float2 ta = mt_intersect_pair(v0a, e0a, e1a, pos, dir, ok0, ok1);
            const float2 tb = mt_intersect_pair(v0b, e0b, e1b, pos, dir, ok2, ok3);
            blocked |= (ok0 && ta.x > RVB_EPSILON && ta.x <= mag) | (ok1 && ta.y > RVB_EPSILON && ta.y <= mag)
                     | (ok2 && tb.x > RVB_EPSILON && tb.x <= mag) | (ok3 && tb.y > RVB_EPSILON && tb.y <= mag);
        }
        c.queries++; c.tri += T;
        return blocked;
    }
    // The visibility query of a reflection (kernel.cpp:463) and the nearest-hit query of the NEXT one
    // (:363) start at the same point X -- the reflected direction (:492-499) does not depend on the
    // visibility -- so ONE sweep over the records answers both: two queries, 2 T tests, but the
    // direction-independent third of every test and all shared-memory loads are done once.
    __device__ __forceinline__ bool visible_and_next(f3 pos, f3 dir_mic, float mag, f3 dir_next, NextHit &next, Counters &c) const
    {
        bool blocked = false;
        float tbest = RVB_INF;
        uint32_t ibest = RVB_NONE;
#pragma unroll 2
        for (uint32_t q = 0; q < nquads; ++q) {
            f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
            load(q, v0a, e0a, e1a, v0b, e0b, e1b);
            PairHit ma, na, mb, nb;
            mt_intersect_pair_2rays(v0a, e0a, e1a, pos, dir_mic, dir_next, ma, na);
            mt_intersect_pair_2rays(v0b, e0b, e1b, pos, dir_mic, dir_next, mb, nb);
            blocked |= (ma.ok_a && ma.t.x > RVB_EPSILON && ma.t.x <= mag) | (ma.ok_b && ma.t.y > RVB_EPSILON && ma.t.y <= mag)
                     | (mb.ok_a && mb.t.x > RVB_EPSILON && mb.t.x <= mag) | (mb.ok_b && mb.t.y > RVB_EPSILON && mb.t.y <= mag);
            if (na.ok_a && na.t.x > RVB_EPSILON && na.t.x < tbest) { tbest = na.t.x; ibest = 4 * q; }
            if (na.ok_b && na.t.y > RVB_EPSILON && na.t.y < tbest) { tbest = na.t.y; ibest = 4 * q + 1; }
            if (nb.ok_a && nb.t.x > RVB_EPSILON && nb.t.x < tbest) { tbest = nb.t.x; ibest = 4 * q + 2; }
            if (nb.ok_b && nb.t.y > RVB_EPSILON && nb.t.y < tbest) { tbest = nb.t.y; ibest = 4 * q + 3; }
        }
        next.t = tbest; next.prim = ibest;
        c.queries += 2; c.tri += 2ull * T;
        return blocked;
    }
};

// ---- brute force, one WARP per ray (few rays, many triangles) ----------------------------------
// With a few thousand rays the thread-per-ray loop leaves most of the GPU idle and every query costs
// the latency of T tests in one lane.  Here the 32 lanes of a warp carry the SAME ray (they execute
// the per-ray code redundantly and stay converged because they compute identical values) and split
// the triangle records: lane l tests records l, l + 32, ...  The (t, index) minimum over the lanes,
// lexicographic, is exactly what the in-order strict '<' loop keeps (kernel.cpp:180-188).
struct BruteSmemWarp {
    static constexpr bool kFused = false;
    const float4 *s;
    uint32_t T, nquads;

    __device__ __forceinline__ uint32_t my_share() const   // triangles this lane tests: the shares add up to T
    {
        uint32_t n = 0;
        for (uint32_t q = threadIdx.x & 31u; q < nquads; q += 32u) n += min(4u, T - 4u * q);
        return n;
    }
    __device__ __forceinline__ bool nearest(f3 pos, f3 dir, float &tbest, uint32_t &ibest, Counters &c) const
    {
        const BruteSmem one{s, T, nquads};
        float tb = RVB_INF;
        uint32_t ib = RVB_NONE;
        for (uint32_t q = threadIdx.x & 31u; q < nquads; q += 32u) {
            f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
            one.load(q, v0a, e0a, e1a, v0b, e0b, e1b);
            bool ok0, ok1, ok2, ok3;
            const float2 ta = mt_intersect_pair(v0a, e0a, e1a, pos, dir, ok0, ok1);
            const float2 tq = mt_intersect_pair(v0b, e0b, e1b, pos, dir, ok2, ok3);
            if (ok0 && ta.x > RVB_EPSILON && ta.x < tb) { tb = ta.x; ib = 4 * q; }
            if (ok1 && ta.y > RVB_EPSILON && ta.y < tb) { tb = ta.y; ib = 4 * q + 1; }
            if (ok2 && tq.x > RVB_EPSILON && tq.x < tb) { tb = tq.x; ib = 4 * q + 2; }
            if (ok3 && tq.y > RVB_EPSILON && tq.y < tb) { tb = tq.y; ib = 4 * q + 3; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float t2 = __shfl_xor_sync(0xffffffffu, tb, o);
            const uint32_t i2 = __shfl_xor_sync(0xffffffffu, ib, o);
            if (t2 < tb || (t2 == tb && i2 < ib)) { tb = t2; ib = i2; }
        }
        tbest = tb; ibest = ib;
        if ((threadIdx.x & 31u) == 0) c.queries++;
        c.tri += my_share();
        return ib != RVB_NONE;
    }
    __device__ __forceinline__ bool occluded_to_mic(f3 pos, f3 dir, float mag, Counters &c) const
    {
        const BruteSmem one{s, T, nquads};
        bool blocked = false;
        for (uint32_t q = threadIdx.x & 31u; q < nquads; q += 32u) {
            f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
            one.load(q, v0a, e0a, e1a, v0b, e0b, e1b);
            bool ok0, ok1, ok2, ok3;
            const float2 ta = mt_intersect_pair(v0a, e0a, e1a, pos, dir, ok0, ok1);
            const float2 tq = mt_intersect_pair(v0b, e0b, e1b, pos, dir, ok2, ok3);
            blocked |= (ok0 && ta.x > RVB_EPSILON && ta.x <= mag) | (ok1 && ta.y > RVB_EPSILON && ta.y <= mag)
                     | (ok2 && tq.x > RVB_EPSILON && tq.x <= mag) | (ok3 && tq.y > RVB_EPSILON && tq.y <= mag);
        }
        blocked = __any_sync(0xffffffffu, blocked);
        if ((threadIdx.x & 31u) == 0) c.queries++;
        c.tri += my_share();
        return blocked;
    }
};

// ---- BVH traversal ----------------------------------------------------------------------------
// Boxes are padded at build time and the slab test carries a relative slack, so no triangle the
// exact test would accept at (t, index) better than the current best is ever culled; candidates
// are compared lexicographically on (t, original index) to reproduce the in-order strict '<' loop.
// Leaves are records of up to four triangles in the pair-interleaved form of the brute-force loop
// (bvh.cu: fill_leaves_kernel): one burst of nine 16-byte loads, two packed tests.
// Visibility queries towards the microphone go through the visibility map (micmap.cuh) when the
// trace has one.
#define RVB_SENTINEL ((int)0x80000000)

// 32-byte load (LDG.E.256 on sm_100): the L1 data pipe charges a divergent load per (instruction,
// 128-byte line), not per byte, so a 128-byte node costs four of these instead of seven 16-byte loads.
struct f8 { float4 a, b; };
__device__ __forceinline__ f8 ldg256(const void *p)
{
    // (as four 64-bit halves: ptxas 12.9 crashes on ld.global.nc.v8.f32 under --fmad=false)
    unsigned long long x, y, z, w;
    asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(x), "=l"(y), "=l"(z), "=l"(w) : "l"(p));
    f8 r;
    r.a.x = __uint_as_float((unsigned)x); r.a.y = __uint_as_float((unsigned)(x >> 32));
    r.a.z = __uint_as_float((unsigned)y); r.a.w = __uint_as_float((unsigned)(y >> 32));
    r.b.x = __uint_as_float((unsigned)z); r.b.y = __uint_as_float((unsigned)(z >> 32));
    r.b.z = __uint_as_float((unsigned)w); r.b.w = __uint_as_float((unsigned)(w >> 32));
    return r;
}   // bottom of the traversal stack; leaf refs are >= -2^30, inner nodes >= 0

// The traversal stack lives in shared memory, one column per thread (entry k of thread t at
// [k * 128 + t]: conflict-free whatever the lanes' stack depths are).  In local memory every push /
// pop of a warp whose lanes sit at different depths touches up to 32 different lines, and the L1
// data pipe -- which is what bounds these kernels (ncu: l1tex data-pipe wavefronts 85-90 % of peak)
// -- serves them one after the other.  All BVH kernels run 128-thread CTAs; the number of entries
// comes from the depth of the hierarchy (Scene::bvh_stack).
#define RVB_BVH_CTA 128
struct TravStack {
    int *col;
    __device__ __forceinline__ void push(int &sp, int v) const { col[sp * RVB_BVH_CTA] = v; ++sp; }
    __device__ __forceinline__ int pop(int &sp) const { --sp; return col[sp * RVB_BVH_CTA]; }
};

struct BvhAccel {
    static constexpr bool kFused = false;
    const BvhNode *nodes;
    const float4 *leaf_quads;   // [RVB_LEAF_STRIDE * n_leaves]: 36 floats of triangle data + the four triangle indices
    const uint4 *leaf_ids;      // unused (the indices ride in the record)
    const float4 *tri;          // [3T] (v0, e0, e1) in the callers' order: what the visibility map indexes
    MicMap mm;
    TravStack stk;

    // Per-query constants of the slab test: t = plane * inv + oinv with oinv = -origin * inv (one fma
    // per plane).  |inv| is clamped to 1e30 so that an axis-parallel ray gives huge finite values of
    // the right sign instead of inf - inf; the rounding of the fma form moves a plane by ~6e-8 * |origin|,
    // three orders of magnitude less than the padding of the boxes.
    struct SlabRay { f3 inv, oinv; };
    __device__ __forceinline__ static SlabRay slab_ray(f3 pos, f3 dir)
    {
        SlabRay r;
        r.inv = mk3(fminf(fmaxf(__fdividef(1.0f, dir.x), -1e30f), 1e30f), fminf(fmaxf(__fdividef(1.0f, dir.y), -1e30f), 1e30f),
                    fminf(fmaxf(__fdividef(1.0f, dir.z), -1e30f), 1e30f));
        if (dir.x == 0.0f) r.inv.x = 1e30f;   // 1/0 = inf is clamped by fminf, but 1/-0 and NaN are not worth a thought: be explicit
        if (dir.y == 0.0f) r.inv.y = 1e30f;
        if (dir.z == 0.0f) r.inv.z = 1e30f;
        r.oinv = mk3(-pos.x * r.inv.x, -pos.y * r.inv.y, -pos.z * r.inv.z);
        return r;
    }

    // entry distance of one child box (inf when missed); the six plane distances are already computed
    __device__ __forceinline__ static float slab_hit(float tx0, float tx1, float ty0, float ty1, float tz0, float tz1, float lim)
    {
        const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
        const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), lim));
        // 1.00001f covers the rounding of the plane distances; NaNs are dropped by fminf / fmaxf, which widens the interval
        return tn <= tf * 1.00001f + 1e-6f ? tn : RVB_INF;
    }

    // One visit of a 4-wide node: decode and test the four child boxes, continue with the nearest hit
    // child and push the others so that the nearer ones are popped first.  Returns the next node.
    //
    // The node is 64 bytes (rvb_internal.cuh): plane = origin + q * cell, q a byte.  With
    // m = 1 + q * 2^-15 (the byte dropped into a float's mantissa by ONE PRMT) the plane distance is
    //   t = plane * inv + oinv = m * A + B,   A = 2^15 * cell * inv,   B = (origin * inv + oinv) - A
    // i.e. still one fma per plane (two children per packed fma); A and B cost nine operations per
    // node.  Rounding: B is rounded at magnitude ~2^15 cell |inv|, i.e. the planes move by ~2^-9 of a
    // cell -- far inside the padding every box carries (the q are rounded outwards on top of it).
    __device__ __forceinline__ static int visit(const BvhNode *nodes, int node, const SlabRay &sr, float tbest, const TravStack &stack, int &sp)
    {
        const char *np = reinterpret_cast<const char *>(nodes + node);
        const f8 n0 = ldg256(np), n1 = ldg256(np + 32);
        const uint32_t exps = __float_as_uint(n0.a.w);
        const uint32_t qlx = __float_as_uint(n0.b.x), qly = __float_as_uint(n0.b.y), qlz = __float_as_uint(n0.b.z);
        const uint32_t qhx = __float_as_uint(n0.b.w), qhy = __float_as_uint(n1.a.x), qhz = __float_as_uint(n1.a.y);
        int r0 = __float_as_int(n1.a.z), r1 = __float_as_int(n1.a.w), r2 = __float_as_int(n1.b.x), r3 = __float_as_int(n1.b.y);
        const float Ax = __uint_as_float(((exps & 0xffu) + 15u) << 23) * sr.inv.x;
        const float Ay = __uint_as_float((((exps >> 8) & 0xffu) + 15u) << 23) * sr.inv.y;
        const float Az = __uint_as_float((((exps >> 16) & 0xffu) + 15u) << 23) * sr.inv.z;
        const float Bx = __fmaf_rn(n0.a.x, sr.inv.x, sr.oinv.x) - Ax;
        const float By = __fmaf_rn(n0.a.y, sr.inv.y, sr.oinv.y) - Ay;
        const float Bz = __fmaf_rn(n0.a.z, sr.inv.z, sr.oinv.z) - Az;
        // candidates farther than the best are useless; equal t must still be visited (tie-break)
        const float lim = tbest * 1.00001f + 1e-6f;
        const float INF = RVB_INF;
#define RVB_M(w, k) __uint_as_float(__byte_perm((w), 0x3f800000u, 0x7604u | ((k) << 4)))   /* 1 + byte k * 2^-15 */
#define RVB_PLANES(w, A, B, ta, tb)                                                                              \
        const float2 ta = __ffma2_rn(make_float2(RVB_M(w, 0u), RVB_M(w, 1u)), bcast2(A), bcast2(B)),              \
                     tb = __ffma2_rn(make_float2(RVB_M(w, 2u), RVB_M(w, 3u)), bcast2(A), bcast2(B));
        RVB_PLANES(qlx, Ax, Bx, ax0, bx0) RVB_PLANES(qhx, Ax, Bx, ax1, bx1)
        RVB_PLANES(qly, Ay, By, ay0, by0) RVB_PLANES(qhy, Ay, By, ay1, by1)
        RVB_PLANES(qlz, Az, Bz, az0, bz0) RVB_PLANES(qhz, Az, Bz, az1, bz1)
#undef RVB_PLANES
#undef RVB_M
        float t0 = slab_hit(ax0.x, ax1.x, ay0.x, ay1.x, az0.x, az1.x, lim);
        float t1 = slab_hit(ax0.y, ax1.y, ay0.y, ay1.y, az0.y, az1.y, lim);
        float t2 = slab_hit(bx0.x, bx1.x, by0.x, by1.x, bz0.x, bz1.x, lim);
        float t3 = slab_hit(bx0.y, bx1.y, by0.y, by1.y, bz0.y, bz1.y, lim);
        // slots 0 and 1 always exist; unused slots are marked in the ref
        if (r2 == RVB_BVH_EMPTY) t2 = INF;
        if (r3 == RVB_BVH_EMPTY) t3 = INF;
        // sorting network on (t, ref), ascending; misses carry +inf
#define RVB_CSWAP(ta, ra, tb, rb) { const bool sw = tb < ta; const float tt = sw ? tb : ta; const int rr = sw ? rb : ra; \
                                    tb = sw ? ta : tb; rb = sw ? ra : rb; ta = tt; ra = rr; }
        RVB_CSWAP(t0, r0, t1, r1) RVB_CSWAP(t2, r2, t3, r3) RVB_CSWAP(t0, r0, t2, r2) RVB_CSWAP(t1, r1, t3, r3) RVB_CSWAP(t1, r1, t2, r2)
#undef RVB_CSWAP
        if (t3 < INF) stack.push(sp, r3);
        if (t2 < INF) stack.push(sp, r2);
        if (t1 < INF) stack.push(sp, r1);
        return t0 < INF ? r0 : stack.pop(sp);
    }

    // The (<= 4) triangles of one leaf record against the ray: the packed pair test of the
    // brute-force loop (same per-lane operation sequence as mt_intersect_pred => same bits).
    // Nearest: lexicographic (t, callers' index) minimum; ANY: is there a hit with EPS < t <= tlimit.
    template <bool ANY>
    __device__ __forceinline__ void leaf(int node, f3 pos, f3 dir, float tlimit, float &tbest, uint32_t &ibest, bool &found, uint32_t &ntri) const
    {
        const uint32_t ref = ~(uint32_t)node;
        const uint32_t rec = ref & 0x0fffffffu;
        ntri += (ref >> 28) + 1;
        const char *r = reinterpret_cast<const char *>(leaf_quads + RVB_LEAF_STRIDE * (size_t)rec);
        const f8 q0 = ldg256(r), q1 = ldg256(r + 32), q2 = ldg256(r + 64), q3 = ldg256(r + 96), q4 = ldg256(r + 128);
        const float4 r0 = q0.a, r1 = q0.b, r2 = q1.a, r3 = q1.b, r4 = q2.a, r5 = q2.b, r6 = q3.a, r7 = q3.b, r8 = q4.a;
        f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
        v0a.x = make_float2(r0.x, r0.y); v0a.y = make_float2(r0.z, r0.w); v0a.z = make_float2(r1.x, r1.y);
        e0a.x = make_float2(r1.z, r1.w); e0a.y = make_float2(r2.x, r2.y); e0a.z = make_float2(r2.z, r2.w);
        e1a.x = make_float2(r3.x, r3.y); e1a.y = make_float2(r3.z, r3.w); e1a.z = make_float2(r4.x, r4.y);
        v0b.x = make_float2(r4.z, r4.w); v0b.y = make_float2(r5.x, r5.y); v0b.z = make_float2(r5.z, r5.w);
        e0b.x = make_float2(r6.x, r6.y); e0b.y = make_float2(r6.z, r6.w); e0b.z = make_float2(r7.x, r7.y);
        e1b.x = make_float2(r7.z, r7.w); e1b.y = make_float2(r8.x, r8.y); e1b.z = make_float2(r8.z, r8.w);
        bool ok0, ok1, ok2, ok3;
        const float2 ta = mt_intersect_pair(v0a, e0a, e1a, pos, dir, ok0, ok1);
        const float2 tb = mt_intersect_pair(v0b, e0b, e1b, pos, dir, ok2, ok3);
        ok0 = ok0 && ta.x > RVB_EPSILON; ok1 = ok1 && ta.y > RVB_EPSILON;
        ok2 = ok2 && tb.x > RVB_EPSILON; ok3 = ok3 && tb.y > RVB_EPSILON;
        if (ANY) {
            found |= (ok0 && ta.x <= tlimit) | (ok1 && ta.y <= tlimit) | (ok2 && tb.x <= tlimit) | (ok3 && tb.y <= tlimit);
        } else if (ok0 | ok1 | ok2 | ok3) {
            const uint4 id = make_uint4(__float_as_uint(q4.b.x), __float_as_uint(q4.b.y), __float_as_uint(q4.b.z), __float_as_uint(q4.b.w));
            if (ok0 && (ta.x < tbest || (ta.x == tbest && id.x < ibest))) { tbest = ta.x; ibest = id.x; }
            if (ok1 && (ta.y < tbest || (ta.y == tbest && id.y < ibest))) { tbest = ta.y; ibest = id.y; }
            if (ok2 && (tb.x < tbest || (tb.x == tbest && id.z < ibest))) { tbest = tb.x; ibest = id.z; }
            if (ok3 && (tb.y < tbest || (tb.y == tbest && id.w < ibest))) { tbest = tb.y; ibest = id.w; }
        }
    }

    template <bool ANY>
    __device__ __forceinline__ bool traverse(f3 pos, f3 dir, float tlimit, float &tbest, uint32_t &ibest, Counters &c) const
    {
        const SlabRay sr = slab_ray(pos, dir);
        tbest = tlimit;
        ibest = RVB_NONE;
        // while-while traversal: every lane first walks inner nodes until it holds a leaf (or is
        // done), then the leaves are intersected together -- inner-node code and triangle code are
        // never interleaved inside a warp.
        int sp = 0;
        stk.push(sp, RVB_SENTINEL);
        int node = 0;
        uint32_t nbox = 0, ntri = 0;
        bool found = false;
        while (node != RVB_SENTINEL) {
            while (node >= 0) {
                node = visit(nodes, node, sr, tbest, stk, sp);
                nbox += 4;
            }
            while (node < 0 && node != RVB_SENTINEL) {
                leaf<ANY>(node, pos, dir, tlimit, tbest, ibest, found, ntri);
                node = (ANY && found) ? RVB_SENTINEL : stk.pop(sp);
            }
        }
        c.queries++; c.tri += ntri; c.box += nbox;
        return ANY ? found : (ibest != RVB_NONE);
    }
    __device__ __forceinline__ bool nearest(f3 pos, f3 dir, float &tbest, uint32_t &ibest, Counters &c) const
    {
        return traverse<false>(pos, dir, RVB_INF, tbest, ibest, c);
    }
    // any-hit traversal for traces without a visibility map; out of line so that the hot kernels do
    // not carry a second traversal loop
    __device__ __noinline__ static bool occluded_generic(const BvhNode *nodes, const float4 *leaf_quads, const uint4 *leaf_ids, int *stack_col,
                                                         f3 begin, f3 dir, float mag, Counters &c)
    {
        const BvhAccel a{nodes, leaf_quads, leaf_ids, nullptr, MicMap{}, TravStack{stack_col}};
        float t; uint32_t i;
        return a.traverse<true>(begin, dir, mag, t, i, c);
    }
    // point_intersection(begin, mic) (kernel.cpp:274-296): dir = normalize(mic - begin), mag = |mic - begin|
    __device__ __forceinline__ bool occluded_to_mic(f3 begin, f3 dir, float mag, Counters &c) const
    {
        // (the caller's own traversal is over when it asks for visibility, so its stack column is free)
        if (mm.N == 0) return occluded_generic(nodes, leaf_quads, leaf_ids, stk.col, begin, dir, mag, c);
        bool blocked = false;
        const f3 w = mk3(begin.x - mm.mic[0], begin.y - mm.mic[1], begin.z - mm.mic[2]);
        const uint32_t cell = micmap_cell(w, mm.half_n, mm.N);
        const uint32_t e0 = __ldg(mm.cell_off + cell), ne = __ldg(mm.cell_off + cell + 1) - e0;
        const uint32_t ntri = ne + mm.n_near;
        // Two candidates per iteration (their index loads, then their triangle loads, are issued together:
        // two memory latencies per pair).  Finer than that starves the loads, coarser wastes lanes: the
        // lanes of a warp hold 1..12 candidates each and wait for the longest list (ncu, groups of four:
        // 7.8 of 24 lanes active in this loop).  Past the end the last candidate is simply tested again --
        // the answer is an OR -- so every load is unconditional.
        for (uint32_t k0 = 0; k0 < ntri; k0 += 2) {
            const uint32_t ka = k0, kb = min(k0 + 1, ntri - 1);
            const uint32_t ia = ka < ne ? __ldg(mm.cell_ent + e0 + ka) : __ldg(mm.near_ent + (ka - ne));
            const uint32_t ib = kb < ne ? __ldg(mm.cell_ent + e0 + kb) : __ldg(mm.near_ent + (kb - ne));
            const float4 *pa = tri + 3 * (size_t)ia, *pb = tri + 3 * (size_t)ib;
            const float4 Aa = __ldg(pa), Ba = __ldg(pa + 1), Ab = __ldg(pb), Bb = __ldg(pb + 1);
            const float Ca = __ldg(reinterpret_cast<const float *>(pa + 2)), Cb = __ldg(reinterpret_cast<const float *>(pb + 2));
            bool oka, okb;
            const float ta = mt_intersect_pred(mk3(Aa.x, Aa.y, Aa.z), mk3(Aa.w, Ba.x, Ba.y), mk3(Ba.z, Ba.w, Ca), begin, dir, oka);
            const float tb = mt_intersect_pred(mk3(Ab.x, Ab.y, Ab.z), mk3(Ab.w, Bb.x, Bb.y), mk3(Bb.z, Bb.w, Cb), begin, dir, okb);
            blocked |= (oka && ta > RVB_EPSILON && ta <= mag) | (okb && tb > RVB_EPSILON && tb <= mag);
        }
        c.queries++; c.tri += ntri;
        return blocked;
    }

    // The same query for a whole warp at once (wavefront kernel).  In the per-lane form above the lanes of
    // a warp hold 1..12 candidates each and wait for the longest list: 7.1 of 25 serving lanes were active in
    // that loop, which -- two dependent load latencies per iteration -- was 16 % of the kernel's stall samples
    // (ncu source page, profiles/r2_wavefront_source_hotspots.md).  Here the (query, candidate pair) items of all
    // serving lanes are numbered through (inclusive scan of the item counts) and dealt to ALL 32 lanes, which
    // fetch their item's query from its owner by shuffle: ceil(items / 32) iterations instead of the longest
    // list's.  The test itself is the one above on the same operands, the answer an OR -- so the same bits.
    // q: this lane has a query; e0 / ne: its cell's entry range (micmap_lookup); returns blocked for q lanes.
    __device__ __forceinline__ void micmap_lookup(f3 begin, uint32_t &e0, uint32_t &ne) const
    {
        const f3 w = mk3(begin.x - mm.mic[0], begin.y - mm.mic[1], begin.z - mm.mic[2]);
        const uint32_t cell = micmap_cell(w, mm.half_n, mm.N);
        e0 = __ldg(mm.cell_off + cell);
        ne = __ldg(mm.cell_off + cell + 1) - e0;
    }
    __device__ __forceinline__ bool warp_occluded_to_mic(bool q, f3 begin, f3 dir, float mag, uint32_t e0, uint32_t ne,
                                                         uint32_t *s_inc /* [32], this warp's */, uint32_t *s_blk /* this warp's */) const
    {
        const unsigned FULL = 0xffffffffu;
        const uint32_t lane = threadIdx.x & 31u;
        const uint32_t ntri = ne + mm.n_near;
        const uint32_t items = q ? (ntri + 1u) >> 1 : 0u;
        uint32_t inc = items;
#pragma unroll
        for (uint32_t o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const uint32_t total = __shfl_sync(FULL, inc, 31);
        s_inc[lane] = inc;
        if (lane == 0) *s_blk = 0u;
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32u) {
            const uint32_t item = base + lane;
            const bool act = item < total;
            // owner: the first lane whose inclusive count exceeds the item number (it has items, by construction)
            uint32_t lo = 0;
            if (act) {
#pragma unroll
                for (uint32_t step = 16; step; step >>= 1)
                    if (s_inc[lo + step - 1] <= item) lo += step;
            }
            const uint32_t oe0 = __shfl_sync(FULL, e0, lo), one = __shfl_sync(FULL, ne, lo);
            const f3 ob = mk3(__shfl_sync(FULL, begin.x, lo), __shfl_sync(FULL, begin.y, lo), __shfl_sync(FULL, begin.z, lo));
            const f3 od = mk3(__shfl_sync(FULL, dir.x, lo), __shfl_sync(FULL, dir.y, lo), __shfl_sync(FULL, dir.z, lo));
            const float omag = __shfl_sync(FULL, mag, lo);
            if (act) {
                const uint32_t ontri = one + mm.n_near;
                const uint32_t k0 = 2u * (item - (lo ? s_inc[lo - 1] : 0u));
                const uint32_t ka = k0, kb = min(k0 + 1, ontri - 1);
                const uint32_t ia = ka < one ? __ldg(mm.cell_ent + oe0 + ka) : __ldg(mm.near_ent + (ka - one));
                const uint32_t ib = kb < one ? __ldg(mm.cell_ent + oe0 + kb) : __ldg(mm.near_ent + (kb - one));
                const float4 *pa = tri + 3 * (size_t)ia, *pb = tri + 3 * (size_t)ib;
                const float4 Aa = __ldg(pa), Ba = __ldg(pa + 1), Ab = __ldg(pb), Bb = __ldg(pb + 1);
                const float Ca = __ldg(reinterpret_cast<const float *>(pa + 2)), Cb = __ldg(reinterpret_cast<const float *>(pb + 2));
                bool oka, okb;
                const float ta = mt_intersect_pred(mk3(Aa.x, Aa.y, Aa.z), mk3(Aa.w, Ba.x, Ba.y), mk3(Ba.z, Ba.w, Ca), ob, od, oka);
                const float tb = mt_intersect_pred(mk3(Ab.x, Ab.y, Ab.z), mk3(Ab.w, Bb.x, Bb.y), mk3(Bb.z, Bb.w, Cb), ob, od, okb);
                if ((oka && ta > RVB_EPSILON && ta <= omag) | (okb && tb > RVB_EPSILON && tb <= omag)) atomicOr(s_blk, 1u << lo);
            }
        }
        __syncwarp();
        const bool blocked = ((*s_blk >> lane) & 1u) != 0u;
        __syncwarp();   // the scratch is reused by the next call
        return blocked;
    }
};

// kernel.cpp:195-214: pow(M_E, distance * AIR) * 1, with M_E narrowed to float, for two bands at
// once (packed FP32).  pow(e_f, x) = 2^(x * log2(e_f)) with log2(e_f) = L_HI + L_LO:
//   t = fma(x, L_HI, 1.5 * 2^23)      low mantissa bits of t = n = rint(x * L_HI) (two's complement)
//   n = t - 1.5 * 2^23
//   r = fma(x, L_LO, fma(x, L_HI, -n))   |r| <= 0.5 (+ 1 ulp), the reduction error is ~2^-26 absolute
//   2^r by ex2.approx (max relative error 2^-22), times 2^n through the exponent field.
// Total error <= 2.6e-7 relative, i.e. ~2 ulp -- the same class as powf itself (CUDA documents
// 4 ulp for powf); volumes are compared with a 2e-6 relative tolerance.  Every product feeds an
// fma, so there is no mul + add pair that ptxas could contract differently in different kernels.
__device__ __forceinline__ float2 pow_e_pair(float2 x)
{
    const float2 L_HI = bcast2(1.4426950216293335f);        // float(log2((float)M_E))
    const float2 L_LO = bcast2(-2.4551930266625277e-08f);   // log2((float)M_E) - L_HI
    const float MAGIC = 12582912.0f;                        // 1.5 * 2^23
    const float2 t = __ffma2_rn(x, L_HI, bcast2(MAGIC));
    const float2 n = __fadd2_rn(t, bcast2(-MAGIC));
    const float2 r = __ffma2_rn(x, L_LO, __ffma2_rn(x, L_HI, neg2(n)));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(r.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(r.y));
    const float2 scale = make_float2(__int_as_float((__float_as_int(t.x) << 23) + 0x3f800000),
                                     __int_as_float((__float_as_int(t.y) << 23) + 0x3f800000));   // 2^n
    return __fmul2_rn(e, scale);
}
__device__ __noinline__ float pow_e_slow(float x) { return powf(RVB_M_E_F, x); }
// All eight bands, as four pairs.  The exponents distance * AIR are <= 0 (distance >= 0, every
// AIR < 0, rayverb.cpp:632-641) and AIR[7] is the most negative, so one test on band 7 covers the
// fast path's range for all of them; NaN, negative distances and exponents below -83 (result <
// 2^-120) take the library powf -- band pair by band pair: after ~1400 m of path (the last of 128
// reflections in a large hall) band 7 leaves the fast path's range, and sending every band of such a
// lane through powf made its whole warp wait (2 % of the wavefront kernel's stall samples).
__device__ __forceinline__ void air_attenuation(float distance, const float *air, float2 *att)
{
    const float worst = __fmul_rn(__fmul_rn(distance, air[RVB_NBANDS - 1]), 1.4426950216293335f);
    if (worst > -120.0f && distance >= 0.0f) {
#pragma unroll
        for (int b = 0; b < RVB_NBANDS / 2; ++b) att[b] = pow_e_pair(__fmul2_rn(bcast2(distance), make_float2(air[2 * b], air[2 * b + 1])));
    } else {
#pragma unroll
        for (int b = 0; b < RVB_NBANDS / 2; ++b) {
            const float w = __fmul_rn(__fmul_rn(distance, air[2 * b + 1]), 1.4426950216293335f);   // the pair's more negative exponent
            if (w > -120.0f && distance >= 0.0f) att[b] = pow_e_pair(__fmul2_rn(bcast2(distance), make_float2(air[2 * b], air[2 * b + 1])));
            else att[b] = make_float2(pow_e_slow(__fmul_rn(distance, air[2 * b])), pow_e_slow(__fmul_rn(distance, air[2 * b + 1])));
        }
    }
}

// kernel.cpp:274-296.  Every call site of the reference passes the microphone as `point`
// (:335, :431, :463), which is what lets the BVH path answer from the visibility map.
template <class Accel>
__device__ __forceinline__ bool point_visible(const Accel &acc, f3 begin, f3 point, Counters &c)
{
    f3 btp = point - begin;
    float mag = length(btp);
    f3 dir = normalize(btp);
    return !acc.occluded_to_mic(begin, dir, mag, c);
}

// TriVerts with its normal cached: triangle_verts_normal (kernel.cpp:110-116) is a pure function
// of the three vertices, so evaluating it once per mirrored triangle gives the bits the
// reference recomputes on every mirror_point call (kernel.cpp:219).
struct MirrorTri { f3 v0, v1, v2, n; };

// kernel.cpp:217-221: *p += -n * dot(n, *p - t->v0) * 2
__device__ __forceinline__ f3 mirror_point(f3 p, const MirrorTri &t)
{
    const float d = dot(t.n, p - t.v0);
    return p + ((-t.n) * d) * 2.0f;
}

// kernel.cpp:243-265
__device__ __forceinline__ void add_image(const TraceParams &P, const TraceJob &J, f3 mic, f3 mic_reflection, f3 source, size_t slot,
                                          const float2 *volume, uint32_t object_index)
{
    f3 diff = source - mic_reflection;
    float dist = length(diff);
    float2 att[RVB_NBANDS / 2];
    air_attenuation(dist, P.air, att);
    float2 v[RVB_NBANDS / 2];
#pragma unroll
    for (int b = 0; b < RVB_NBANDS / 2; ++b) v[b] = __fmul2_rn(volume[b], att[b]);
    f3 p = mic + diff;
    J.cand_volA[slot] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    J.cand_volB[slot] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
    J.cand_pt[slot] = make_float4(p.x, p.y, p.z, __fmul_rn(RVB_SECONDS_PER_METER, dist));
    J.cand_idx[slot] = object_index;
}

struct RayState {
    f3 pos, dir;
    float distance;
    float2 volume[RVB_NBANDS / 2];   // eight bands as four packed pairs
};

// kernel.cpp:459-501: everything a reflection does after the nearest hit is known
// `next`: with a fused accelerator (Accel::kFused) and next->want, the visibility sweep also answers
// the nearest-hit query of the reflected ray (see BruteSmem::visible_and_next).
// `pre`: the visibility of the hit point was already worked out by the caller (BvhAccel::warp_occluded_to_mic);
// ntri = the candidates that query tested, for the counters.
struct PreVisibility { bool visible; uint32_t ntri; };
template <class Accel>
__device__ __forceinline__ void shade_and_reflect(const Accel &acc, const TraceParams &P, const TraceJob &J, f3 mic, uint32_t ray, uint32_t index,
                                                  float tclosest, uint32_t prim, RayState &R, Counters &c, NextHit *next = nullptr,
                                                  const PreVisibility *pre = nullptr)
{
    const f3 X = madd(R.pos, R.dir, tclosest);                               // :459
    const float newDist = __fadd_rn(R.distance, tclosest);
    const float4 nrm4 = __ldg(P.nrm + prim);
    const f3 normal = mk3(nrm4);
    const float4 *surf = reinterpret_cast<const float4 *>(P.surfaces + __float_as_uint(nrm4.w));
    const float4 s0 = __ldg(surf), s1 = __ldg(surf + 1), d0 = __ldg(surf + 2), d1 = __ldg(surf + 3);
    const float2 spec[RVB_NBANDS / 2] = {make_float2(s0.x, s0.y), make_float2(s0.z, s0.w), make_float2(s1.x, s1.y), make_float2(s1.z, s1.w)};
    const float2 diff[RVB_NBANDS / 2] = {make_float2(d0.x, d0.y), make_float2(d0.z, d0.w), make_float2(d1.x, d1.y), make_float2(d1.z, d1.w)};
    float2 newVol[RVB_NBANDS / 2];
#pragma unroll
    for (int b = 0; b < RVB_NBANDS / 2; ++b) newVol[b] = __fmul2_rn(neg2(R.volume[b]), spec[b]);   // :461

    const f3 newDir = reflect(normal, R.dir);                                // :492-499 (independent of the visibility)
    bool vis;
    if constexpr (Accel::kFused) {
        if (next && next->want) {
            const f3 btp = mic - X;                                          // point_intersection, :274-296
            const float mag = length(btp);
            vis = !acc.visible_and_next(X, normalize(btp), mag, newDir, *next, c);
        } else vis = point_visible(acc, X, mic, c);
    } else if (pre) { vis = pre->visible; c.queries++; c.tri += pre->ntri; }
    else vis = point_visible(acc, X, mic, c);                                // :463
    const float DIST = vis ? __fadd_rn(newDist, length(mic - X)) : 0.0f;     // :471
    const float DIFF = fabsf(dot(normal, R.dir));                            // :478
    float2 out[RVB_NBANDS / 2];
    bool any = false;
    if (vis) {
        float2 att[RVB_NBANDS / 2];
        air_attenuation(DIST, P.air, att);
#pragma unroll
        for (int b = 0; b < RVB_NBANDS / 2; ++b) {
            out[b] = __fmul2_rn(__fmul2_rn(__fmul2_rn(newVol[b], att[b]), diff[b]), bcast2(DIFF));   // :481-485
            any |= (out[b].x != 0.0f) | (out[b].y != 0.0f);
        }
    } else {
#pragma unroll
        for (int b = 0; b < RVB_NBANDS / 2; ++b) out[b] = make_float2(0.0f, 0.0f);
    }
    const float time = __fmul_rn(RVB_SECONDS_PER_METER, DIST);               // :489
    const size_t slot = slot_of(ray, index, P.nrays);
    // streaming stores: the impulses are read back much later (flatten), they should not push the
    // scene out of L2
    __stcs(J.vol + 2 * slot, make_float4(out[0].x, out[0].y, out[1].x, out[1].y));
    __stcs(J.vol + 2 * slot + 1, make_float4(out[2].x, out[2].y, out[3].x, out[3].y));
    __stcs(J.pt + slot, make_float4(X.x, X.y, X.z, time));
    __stcs(J.tkey + slot, any ? time : -1.0f);
    if (any) {
        c.emitted++;
        if (time != 0.0f) c.tmin_nz = fminf(c.tmin_nz, time);
        c.tmax = fmaxf(c.tmax, time);
    }
    if (J.hit) J.hit[slot] = prim;

    R.dir = newDir;
    R.pos = X;
    R.distance = newDist;
#pragma unroll
    for (int b = 0; b < RVB_NBANDS / 2; ++b) R.volume[b] = newVol[b];
}

// a ray that left the mesh keeps zero impulses in its remaining slots (rayverb.cpp:601-603)
__device__ __forceinline__ void zero_fill(const TraceParams &P, const TraceJob &J, uint32_t ray, uint32_t index)
{
    for (; index != P.nrefl; ++index) {
        const size_t slot = slot_of(ray, index, P.nrays);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        J.vol[2 * slot] = z; J.vol[2 * slot + 1] = z; J.pt[slot] = z; J.tkey[slot] = -1.0f;
        if (J.hit) J.hit[slot] = RVB_NONE;
    }
}

__device__ __forceinline__ void store_state(const TraceJob &J, uint32_t ray, const RayState &R, bool alive)
{
    float4 *s = J.state + 4 * (size_t)ray;
    s[0] = make_float4(R.pos.x, R.pos.y, R.pos.z, R.distance);
    s[1] = make_float4(R.dir.x, R.dir.y, R.dir.z, alive ? 1.0f : 0.0f);
    s[2] = make_float4(R.volume[0].x, R.volume[0].y, R.volume[1].x, R.volume[1].y);
    s[3] = make_float4(R.volume[2].x, R.volume[2].y, R.volume[3].x, R.volume[3].y);
}

// Reflections [0, P.last) WITH the image-source branch (kernel.cpp:379-457); P.last <= 9.
template <class Accel>
__device__ void image_phase_ray(const Accel &acc, const TraceParams &P, const TraceJob &J, uint32_t ray, Counters &c)
{
    const f3 mic = mk3(J.mic[0], J.mic[1], J.mic[2]);
    const f3 source = mk3(J.src[0], J.src[1], J.src[2]);
    RayState R;
    R.pos = source;
    R.dir = mk3(__ldg(P.dirs + ray));
    R.distance = 0.0f;
#pragma unroll
    for (int b = 0; b < RVB_NBANDS / 2; ++b) R.volume[b] = make_float2(1.0f, 1.0f);
    MirrorTri prev[RVB_NUM_IMAGE_SOURCE - 1];
    f3 mic_reflection = mic;

    // Direct path (kernel.cpp:335-357).  Every work-item of the reference computes the same
    // source->mic query and the host keeps the first ray's copy (rayverb.cpp:667-673); only
    // global ray 0 needs to do it.
    if (J.ray_offset + ray == 0) {
        if (point_visible(acc, source, mic_reflection, c))
            add_image(P, J, mic, mic_reflection, source, (size_t)ray * RVB_NUM_IMAGE_SOURCE, R.volume, 0u);
    }

    uint32_t index = 0;
    bool alive = true;
    for (; index != P.last; ++index) {
        float tclosest;
        uint32_t prim;
        if (!acc.nearest(R.pos, R.dir, tclosest, prim, c)) { alive = false; break; }   // :363-375

        {   // index < NUM_IMAGE_SOURCE - 1 holds for every reflection of this kernel (:379)
            MirrorTri cur;
            cur.v0 = mk3(__ldg(P.verts + 3 * prim)); cur.v1 = mk3(__ldg(P.verts + 3 * prim + 1)); cur.v2 = mk3(__ldg(P.verts + 3 * prim + 2));
            for (uint32_t k = 0; k != index; ++k) {                              // mirror_verts, :387-390
                cur.v0 = mirror_point(cur.v0, prev[k]);
                cur.v1 = mirror_point(cur.v1, prev[k]);
                cur.v2 = mirror_point(cur.v2, prev[k]);
            }
            cur.n = normalize(cross(cur.v1 - cur.v0, cur.v2 - cur.v0));
            prev[index] = cur;
            mic_reflection = mirror_point(mic_reflection, cur);                  // :394
            const f3 DIR = normalize(mic_reflection - source);                   // :396
            bool intersects = true;
            f3 prevI = source;
            for (uint32_t k = 0; k != index + 1 && intersects; ++k) {            // :401
                const MirrorTri &pk = prev[k];
                const float TO = mt_intersect(pk.v0, pk.v1 - pk.v0, pk.v2 - pk.v0, source, DIR);
                if (TO <= RVB_EPSILON) { intersects = false; break; }            // :405
                f3 Pt = madd(source, DIR, TO);                                   // :411
                for (int l = (int)k - 1; l != -1; --l) Pt = mirror_point(Pt, prev[l]);
                f3 idir = normalize(Pt - prevI);                                 // :417
                float ti; uint32_t pi;
                bool hit = acc.nearest(prevI, idir, ti, pi, c);
                if (!hit) ti = 0.0f;                                             // Intersection{0,0,false}
                f3 Q = madd(prevI, idir, ti);                                    // :425
                intersects = hit
                    && (__fsub_rn(Q.x, RVB_EPSILON) < Pt.x) && (__fsub_rn(Q.y, RVB_EPSILON) < Pt.y) && (__fsub_rn(Q.z, RVB_EPSILON) < Pt.z)
                    && (Pt.x < __fadd_rn(Q.x, RVB_EPSILON)) && (Pt.y < __fadd_rn(Q.y, RVB_EPSILON)) && (Pt.z < __fadd_rn(Q.z, RVB_EPSILON));
                prevI = Pt;
            }
            if (intersects) intersects = point_visible(acc, prevI, mic, c);      // :431
            if (intersects)   // pre-bounce volume
                add_image(P, J, mic, mic_reflection, source, (size_t)ray * RVB_NUM_IMAGE_SOURCE + index + 1, R.volume, prim + 1u);
        }
        shade_and_reflect(acc, P, J, mic, ray, index, tclosest, prim, R, c);
    }
    c.bounces += index;
    if (J.count) J.count[ray] = index;
    if (!alive) zero_fill(P, J, ray, index);
    if (P.last < P.nrefl) store_state(J, ray, R, alive);
}

// Reflections [P.first, P.last) without image sources: the steady state of the reference loop.
template <class Accel>
__device__ __forceinline__ void steady_phase_ray(const Accel &acc, const TraceParams &P, const TraceJob &J, uint32_t ray, Counters &c)
{
    const f3 mic = mk3(J.mic[0], J.mic[1], J.mic[2]);
    RayState R;
    if (P.first == 0) {
        R.pos = mk3(J.src[0], J.src[1], J.src[2]);
        R.dir = mk3(__ldg(P.dirs + ray));
        R.distance = 0.0f;
#pragma unroll
        for (int b = 0; b < RVB_NBANDS / 2; ++b) R.volume[b] = make_float2(1.0f, 1.0f);
    } else {
        const float4 *s = J.state + 4 * (size_t)ray;
        const float4 a = s[0], d = s[1], va = s[2], vb = s[3];
        if (d.w == 0.0f) return;   // left the mesh during the image phase (slots already zero-filled)
        R.pos = mk3(a); R.distance = a.w; R.dir = mk3(d);
        R.volume[0] = make_float2(va.x, va.y); R.volume[1] = make_float2(va.z, va.w);
        R.volume[2] = make_float2(vb.x, vb.y); R.volume[3] = make_float2(vb.z, vb.w);
    }
    uint32_t index = P.first;
    if constexpr (Accel::kFused) {
        // every reflection but the last asks for its successor's nearest hit together with its visibility
        float tclosest;
        uint32_t prim;
        bool hit = acc.nearest(R.pos, R.dir, tclosest, prim, c);                 // :363-375
        while (hit) {
            NextHit next{index + 1 != P.last, 0.0f, RVB_NONE};
            shade_and_reflect(acc, P, J, mic, ray, index, tclosest, prim, R, c, &next);
            ++index;
            if (index == P.last) break;
            tclosest = next.t; prim = next.prim;
            hit = prim != RVB_NONE;
        }
    } else {
        for (; index != P.last; ++index) {
            float tclosest;
            uint32_t prim;
            if (!acc.nearest(R.pos, R.dir, tclosest, prim, c)) break;            // :363-375
            shade_and_reflect(acc, P, J, mic, ray, index, tclosest, prim, R, c);
        }
    }
    c.bounces += index - P.first;
    if (J.count) J.count[ray] = (P.first ? J.count[ray] : 0u) + (index - P.first);
    if (index != P.last) zero_fill(P, J, ray, index);
}

__device__ __forceinline__ void flush_counters(const TraceJob &J, const Counters &c)
{
    // warp-aggregate, then one atomic per warp (integer adds: order-independent)
    unsigned long long q = c.queries, t = c.tri, b = c.box, n = c.bounces, em = c.emitted;
    float mn = c.tmin_nz, mx = c.tmax;
    for (int o = 16; o; o >>= 1) {
        q += __shfl_down_sync(0xffffffffu, q, o); t += __shfl_down_sync(0xffffffffu, t, o);
        b += __shfl_down_sync(0xffffffffu, b, o); n += __shfl_down_sync(0xffffffffu, n, o);
        em += __shfl_down_sync(0xffffffffu, em, o);
        mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(J.counters + 0, q); atomicAdd(J.counters + 1, t);
        atomicAdd(J.counters + 2, b); atomicAdd(J.counters + 3, n); atomicAdd(J.counters + 5, em);
        // order-preserving integer keys: min / max are order-independent, so this is deterministic
        atomicMin(J.time_ext, ordered_key(mn)); atomicMax(J.time_ext + 1, ordered_key(mx));
    }
}

// Stage the whole scene (144 B per four triangles, 16-byte aligned, a multiple of 16 bytes) into shared
// memory with the TMA bulk-copy engine: one elected thread issues cp.async.bulk in <= 32 KiB
// pieces that all complete on one mbarrier; everybody waits on its phase.  (UBLKCP in SASS.)
__device__ __forceinline__ void stage_scene(float4 *s_tri, const TraceParams &P)
{
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t bytes = 9u * P.nquads * (uint32_t)sizeof(float4);
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_tri);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
        const char *src = reinterpret_cast<const char *>(P.quad);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t n = min(32768u, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + off), "l"(src + off), "r"(n), "r"(bar_addr) : "memory");
        }
    }
    // phase 0 completes when all bytes have landed
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
    }
}

// The job a CTA works for (blockIdx.y) in shared memory: word-wise copy by the first threads.
__device__ __forceinline__ void load_job(TraceJob *dst, const TraceJob *src)
{
    static_assert(sizeof(TraceJob) % 4 == 0, "word copy");
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t k = threadIdx.x; k < sizeof(TraceJob) / 4; k += blockDim.x) d[k] = s[k];
    __syncthreads();
}

// BLOCK = 256 for small scenes (several CTAs per SM); when the shared-memory scene leaves room for
// only one CTA per SM the launcher switches to 1024 (steady phase) / 512 (image phase) threads so
// that the schedulers still have 8 / 4 warps each.
// BATCH = false: the single job's record is a kernel parameter (constant bank, no registers or shared
// memory spent on it); BATCH = true: the CTA's job (blockIdx.y) is copied to shared memory.
template <bool IMAGES, int BLOCK, bool BATCH>
__global__ void __launch_bounds__(BLOCK) trace_bruteforce_kernel(const TraceParams P, const TraceJob J0)
{
    extern __shared__ float4 s_tri[];
    __shared__ TraceJob sJ;
    if (BATCH) load_job(&sJ, P.jobs + blockIdx.y);
    const TraceJob &J = BATCH ? sJ : J0;
    stage_scene(s_tri, P);
    BruteSmem acc{s_tri, P.T, P.nquads};
    Counters c;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.nrays; i += gridDim.x * blockDim.x) {
        if (IMAGES) image_phase_ray(acc, P, J, P.ray_perm ? P.ray_perm[i] : i, c); else steady_phase_ray(acc, P, J, i, c);
    }
    flush_counters(J, c);
}

// One warp per ray (BruteSmemWarp): all lanes run the per-ray code on the same ray; identical values
// are stored to identical addresses, and only lane 0's bounce count is kept.
template <bool IMAGES, bool BATCH>
__global__ void __launch_bounds__(256) trace_bruteforce_warp_kernel(const TraceParams P, const TraceJob J0)
{
    extern __shared__ float4 s_tri[];
    __shared__ TraceJob sJ;
    if (BATCH) load_job(&sJ, P.jobs + blockIdx.y);
    const TraceJob &J = BATCH ? sJ : J0;
    stage_scene(s_tri, P);
    BruteSmemWarp acc{s_tri, P.T, P.nquads};
    Counters c;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < P.nrays; ray += nwarps) {
        const uint32_t before = c.bounces, emitted_before = c.emitted;
        if (IMAGES) image_phase_ray(acc, P, J, ray, c); else steady_phase_ray(acc, P, J, ray, c);
        if (threadIdx.x & 31u) { c.bounces = before; c.emitted = emitted_before; }
    }
    flush_counters(J, c);
}

template <bool IMAGES, int MINB, bool BATCH>
__global__ void __launch_bounds__(128, MINB) trace_bvh_kernel(const TraceParams P, const TraceJob J0)
{
    extern __shared__ int s_stack[];
    __shared__ TraceJob sJ;
    if (BATCH) load_job(&sJ, P.jobs + blockIdx.y);
    const TraceJob &J = BATCH ? sJ : J0;
    const BvhAccel acc{P.nodes, P.leaf_quads, P.leaf_ids, P.tri, J.micmap, TravStack{s_stack + threadIdx.x}};
    Counters c;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.nrays; i += gridDim.x * blockDim.x) {
        if (IMAGES) image_phase_ray(acc, P, J, P.ray_perm ? P.ray_perm[i] : i, c); else steady_phase_ray(acc, P, J, i, c);
    }
    flush_counters(J, c);
}

// Direction cell of a ray (256 x 256 cells over z and the azimuth), used only to
// decide the ORDER in which the image phase processes rays: neighbouring lanes then follow the same
// first reflections, so their image-source validation loops have the same trip counts.  The results
// do not depend on the order (every ray writes its own slots).
// (first-pass functor of the library's radix sort, rvb_sort.cuh: the 16-bit cell keys are never stored)
struct DirectionCellIn {
    const float4 *dirs;
    typedef float4 Raw;
    __device__ __forceinline__ Raw fetch(uint32_t i) const { return dirs[i]; }
    __device__ __forceinline__ bool decode(uint32_t i, const Raw &d, uint16_t &k, uint32_t &v) const
    {
        const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        const float inv = len > 0.f ? 1.0f / len : 0.f;
        const float z = fminf(fmaxf(d.z * inv * 0.5f + 0.5f, 0.f), 0.999999f);
        const float az = fminf(fmaxf(atan2f(d.y, d.x) * 0.15915494f + 0.5f, 0.f), 0.999999f);
        const uint32_t iz = (uint32_t)(z * 256.f), ia = (uint32_t)(az * 256.f);
        // boustrophedon over the azimuth so that consecutive cells stay adjacent
        k = (uint16_t)(iz * 256u + ((iz & 1u) ? 255u - ia : ia));
        v = i;
        return true;
    }
};

template <class K>
static int occupancy_of(K kernel, int block, size_t smem, int &per_sm)
{
    RVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    return RVB_OK;
}

template <class K>
static int launch_phase(K kernel, const TraceParams &P, const TraceJob &J0, int block, size_t smem, cudaStream_t st, bool full_grid = false)
{
    int per_sm = 0;
    RVB_TRY(occupancy_of(kernel, block, smem, per_sm));
    if (per_sm < 1) return fail(RVB_ERR_INVALID, "scene does not fit in shared memory for the brute-force kernel");
    // grid.y = job; the resident CTAs are shared out among the jobs
    const size_t want = ((size_t)P.nrays + block - 1) / block;
    const size_t share = std::max<size_t>(1, ((size_t)per_sm * sm_count() + P.njobs - 1) / P.njobs);
    // full_grid: one CTA per `block` rays, handed out by the hardware as SMs free up, instead of a resident
    // grid striding over the rays -- for kernels without per-CTA set-up whose rays differ a lot in cost
    const size_t grid = std::max<size_t>(1, full_grid ? want : std::min(want, share));
    note_launch(), kernel<<<dim3((unsigned)grid, P.njobs), block, smem, st>>>(P, J0);
    RVB_CUDA(cudaGetLastError());
    return RVB_OK;
}

// Brute force: pick the CTA size.  The packed loop wants ~96 registers per thread for full ILP, so
// more resident threads is not automatically better: 1024-thread CTAs are capped at 64 registers
// (spills), 768 at 80, 512 and below are uncapped.  Candidates are ranked by resident threads per SM
// weighted by a measured per-thread efficiency of the register budget (B200, C2 and the 4.4 k
// triangle room); RVB_BLOCK=<n> forces a size for experiments.
template <bool IMAGES, int BLOCK>
static int try_block(const TraceParams &P, size_t smem, int &per_sm)
{
    if (P.njobs > 1) return occupancy_of(trace_bruteforce_kernel<IMAGES, BLOCK, true>, BLOCK, smem, per_sm);
    return occupancy_of(trace_bruteforce_kernel<IMAGES, BLOCK, false>, BLOCK, smem, per_sm);
}
template <bool IMAGES, int BLOCK>
static int launch_block(const TraceParams &P, const TraceJob &J0, size_t smem, cudaStream_t st)
{
    // (one CTA per BLOCK rays instead of a resident grid, as in launch_bvh_images, changes nothing for the
    // shoebox's image phase: 1.70 ms either way)
    if (P.njobs > 1) return launch_phase(trace_bruteforce_kernel<IMAGES, BLOCK, true>, P, J0, BLOCK, smem, st);
    return launch_phase(trace_bruteforce_kernel<IMAGES, BLOCK, false>, P, J0, BLOCK, smem, st);
}
// Few rays over many triangles: one warp per ray.  Cost model in units of one pair-test latency:
// thread-per-ray needs ceil(nrays / resident threads) waves of T tests per query, warp-per-ray
// ceil(32 nrays / resident threads) waves of ceil(nquads / 32) records (4 tests each) + a shuffle
// reduction (~4 tests' worth).
static bool prefer_warp_per_ray(const TraceParams &P)
{
    static const int forced = getenv("RVB_WARP_PER_RAY") ? atoi(getenv("RVB_WARP_PER_RAY")) : -1;
    if (forced >= 0) return forced != 0;
    const double resident = 1024.0 * sm_count();
    const double thread_cost = std::ceil(P.nrays / resident) * P.T;
    const double warp_cost = std::ceil(32.0 * P.nrays / resident) * (4.0 * std::ceil(P.nquads / 32.0) + 4.0);
    return P.T >= 64 && warp_cost * 1.5 < thread_cost;
}

template <bool IMAGES>
static int launch_bruteforce(const TraceParams &P, const TraceJob &J0, size_t smem, cudaStream_t st)
{
    if (prefer_warp_per_ray(P)) {
        int per_sm = 0;
        auto wk = P.njobs > 1 ? trace_bruteforce_warp_kernel<IMAGES, true> : trace_bruteforce_warp_kernel<IMAGES, false>;
        RVB_TRY(occupancy_of(wk, 256, smem, per_sm));
        if (per_sm >= 1) {
            const size_t want = ((size_t)P.nrays * 32 + 255) / 256;
            const size_t share = std::max<size_t>(1, ((size_t)per_sm * sm_count() + P.njobs - 1) / P.njobs);
            const size_t grid = std::max<size_t>(1, std::min(want, share));
            note_launch(), wk<<<dim3((unsigned)grid, P.njobs), 256, smem, st>>>(P, J0);
            RVB_CUDA(cudaGetLastError());
            return RVB_OK;
        }
    }
    static const int forced = getenv("RVB_BLOCK") ? atoi(getenv("RVB_BLOCK")) : 0;
    const int sizes[5] = {256, 384, 512, 768, 1024};
    // steady phase (fused two-ray sweep, ~128 registers): 512 threads measured best on C2 and on the
    // 4352-triangle room (profiles/r2_sweep_bruteforce_block.log); image phase: as in round 1
    const float eff_images[5] = {1.0f, 1.0f, 1.0f, 0.9f, 0.7f}, eff_steady[5] = {0.9f, 0.7f, 1.0f, 0.6f, 0.49f};
    const float *eff = IMAGES ? eff_images : eff_steady;
    int per_sm[5] = {0, 0, 0, 0, 0};
    RVB_TRY((try_block<IMAGES, 256>(P, smem, per_sm[0])));
    RVB_TRY((try_block<IMAGES, 384>(P, smem, per_sm[1])));
    RVB_TRY((try_block<IMAGES, 512>(P, smem, per_sm[2])));
    RVB_TRY((try_block<IMAGES, 768>(P, smem, per_sm[3])));
    RVB_TRY((try_block<IMAGES, 1024>(P, smem, per_sm[4])));
    const size_t nsm = (size_t)sm_count();
    int best = 0;
    float best_score = -1.f;
    for (int i = 0; i < 5; ++i) {
        // threads that would actually be busy with this many rays (small CTAs spread few rays over more SMs)
        const size_t busy = std::min<size_t>(P.nrays, (size_t)per_sm[i] * sizes[i] * nsm);
        float score = (float)busy * eff[i];
        if (forced == sizes[i] && per_sm[i] > 0) score = 1e30f;
        if (score > best_score) { best_score = score; best = i; }
    }
    switch (best) {
    case 0: return launch_block<IMAGES, 256>(P, J0, smem, st);
    case 1: return launch_block<IMAGES, 384>(P, J0, smem, st);
    case 2: return launch_block<IMAGES, 512>(P, J0, smem, st);
    case 3: return launch_block<IMAGES, 768>(P, J0, smem, st);
    default: return launch_block<IMAGES, 1024>(P, J0, smem, st);
    }
}

// =================================================================================================
// Persistent wavefront form of the steady phase for the BVH path.
//
// Thread-per-ray traversal leaves most lanes idle: traversal lengths differ from ray to ray and the
// lanes of a warp wait for the slowest one at the end of EVERY query (ncu, round 1: 13 of 32 threads
// active per instruction).  Here a warp is a persistent worker whose lanes are independent state
// machines that only share the instruction stream:
//
//   serve     lanes whose nearest-hit query is complete shade it (visibility through the microphone
//             map, emit the impulse, reflect: kernel.cpp:459-501) and begin the next reflection's
//             query; a ray that has left the mesh (kernel.cpp:372-375) or finished its last
//             reflection frees its lane
//   refill    free lanes pull the next ray from a global counter (one atomic per warp): terminated
//             rays are compacted away and the warp stays full until the ray set is exhausted
//   traverse  while-while over the 4-wide BVH, all lanes with a pending query; the warp goes back to
//             `serve` once RVB_WF_SERVICE lanes are waiting (or nobody is traversing)
//
// The arithmetic of a ray is the same device code as in the per-ray kernel (same bits); every ray
// owns its output slots, so the order in which lanes pick up rays does not show in the results.
// =================================================================================================
// Per-lane state that only the `serve` step touches (band volumes, path length, ray id, counters)
// lives in shared memory so that the traversal loop keeps a small register set:
// [k][thread] -- conflict-free.
struct WfShared {
    float volume[RVB_NBANDS][128];
    float distance[128], tmin_nz[128], tmax[128];
    uint32_t ray[128], index[128], queries[128], bounces[128], emitted[128], tri[128];
    uint32_t end[128];        // the reflection index at which the lane's current work unit ends
    uint32_t vis_inc[4][32], vis_blk[4];   // scratch of the warp-wide visibility query (BvhAccel::warp_occluded_to_mic)
};

// Segmented rays (SEG).  With fewer than a few waves of rays per launch -- a shard of an 8-GPU run
// holds 131 k rays on 95 k resident lanes -- whole rays are too coarse a work unit: all rays are equally
// long (every one of them runs the full reflection count in a closed room), so the lanes finish their
// first ray together and the remaining 0.4 wave runs on a mostly idle GPU.  A ray's reflections are
// therefore cut into `wf_nseg` units of `wf_seglen` reflections, queued segment-major (unit u = segment
// u / nrays of ray u % nrays).  A lane that finishes a unit stores the ray state (the 64-byte record of
// the image -> steady hand-over: same floats, same bits) and publishes wf_done[ray] with release
// semantics; the lane that draws the ray's next unit reads it with acquire semantics and, if the
// predecessor is still running, keeps the unit PENDING while the rest of the warp goes on traversing
// (it never spins inside a diverged region, and the predecessor was drawn earlier and is being worked
// on by a resident warp, so the wait ends).  A warp with nothing but pending units sleeps and polls; a
// poll budget turns a would-be hang into an error the host reports.
#define RVB_WF_DEAD 0xffffffffu
#define RVB_WF_POLL_LIMIT (1u << 22)
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int MINB, bool BATCH, bool SEG>
__global__ void __launch_bounds__(128, MINB) trace_bvh_wavefront_kernel(const TraceParams P, const TraceJob J0, unsigned int *fetch_all /* [njobs] */)
{
    __shared__ WfShared S;
    __shared__ TraceJob sJ[4];                       // BATCH: the job each warp currently works for
    extern __shared__ int s_stack[];
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, tid = threadIdx.x, warp = threadIdx.x >> 5;
    const TravStack stack{s_stack + tid};
    const TraceJob &J = BATCH ? sJ[warp] : J0;
    // A warp works for ONE job at a time (a plain trace has one job).  With at least as many warps as
    // jobs the warps are dealt round-robin to the jobs, otherwise the jobs to the warps.
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
    const bool many_warps = nw >= P.njobs;
  for (uint32_t job = many_warps ? gw % P.njobs : gw; job < P.njobs; job += many_warps ? P.njobs : nw) {
    if (BATCH) {
        __syncwarp();
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.jobs + job);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sJ[warp]);
        for (uint32_t k = lane; k < sizeof(TraceJob) / 4; k += 32) dst[k] = src[k];
        __syncwarp();
    }
    unsigned int *fetch = fetch_all + job;
    const BvhAccel acc{P.nodes, P.leaf_quads, P.leaf_ids, P.tri, J.micmap, stack};
    int sp = 0, node = RVB_SENTINEL;
    bool has = false, exhausted = false;
    bool pend = false;                               // SEG: holds a unit whose predecessor has not been published yet
    uint32_t polls = 0;
    const uint32_t nunits = SEG ? P.nrays * P.wf_nseg : P.nrays;
    uint32_t ibest = RVB_NONE, nbox = 0, ntri = 0;
    float tbest = RVB_INF;
    f3 pos = mk3(0.f, 0.f, 0.f), dir = pos;
    BvhAccel::SlabRay sr;
    sr.inv = pos; sr.oinv = pos;
    S.tmin_nz[tid] = RVB_INF; S.tmax[tid] = 0.0f;
    S.queries[tid] = 0; S.bounces[tid] = 0; S.emitted[tid] = 0; S.tri[tid] = 0;

    while (true) {
        // ---- serve: completed queries -------------------------------------------------------
        // their visibility queries first, all lanes of the warp sharing the candidate tests
        PreVisibility pv{false, 0u};
        bool have_pv = false;
        if (acc.mm.N != 0u && __any_sync(FULL, has && node == RVB_SENTINEL)) {   // warp-uniform
            const bool q = has && node == RVB_SENTINEL && ibest != RVB_NONE;
            f3 X = pos, vdir = dir;
            float mag = 0.0f;
            uint32_t e0 = 0, ne = 0;
            if (q) {
                const f3 mic = mk3(J.mic[0], J.mic[1], J.mic[2]);
                X = madd(pos, dir, tbest);                       // the hit point, as shade_and_reflect computes it
                const f3 btp = mic - X;                          // point_visible's operands
                mag = length(btp);
                vdir = normalize(btp);
                acc.micmap_lookup(X, e0, ne);
            }
            const bool blocked = acc.warp_occluded_to_mic(q, X, vdir, mag, e0, ne, S.vis_inc[warp], &S.vis_blk[warp]);
            pv.visible = !blocked; pv.ntri = ne + acc.mm.n_near;
            have_pv = q;
        }
        if (has && node == RVB_SENTINEL) {
            const uint32_t ray = S.ray[tid];
            uint32_t index = S.index[tid];
            bool more = false;
            if (ibest != RVB_NONE) {
                const f3 mic = mk3(J.mic[0], J.mic[1], J.mic[2]);
                RayState R;
                R.pos = pos; R.dir = dir; R.distance = S.distance[tid];
#pragma unroll
                for (int b = 0; b < RVB_NBANDS / 2; ++b) R.volume[b] = make_float2(S.volume[2 * b][tid], S.volume[2 * b + 1][tid]);
                Counters c;
                c.tmin_nz = S.tmin_nz[tid]; c.tmax = S.tmax[tid];
                shade_and_reflect(acc, P, J, mic, ray, index, tbest, ibest, R, c, nullptr, have_pv ? &pv : nullptr);
                pos = R.pos; dir = R.dir; S.distance[tid] = R.distance;
#pragma unroll
                for (int b = 0; b < RVB_NBANDS / 2; ++b) { S.volume[2 * b][tid] = R.volume[b].x; S.volume[2 * b + 1][tid] = R.volume[b].y; }
                S.tmin_nz[tid] = c.tmin_nz; S.tmax[tid] = c.tmax;
                S.queries[tid] += 1u + c.queries; S.emitted[tid] += c.emitted; S.tri[tid] += (uint32_t)c.tri;
                S.bounces[tid] += 1u;
                ++index;
                more = index != (SEG ? S.end[tid] : P.last);
                if (SEG && !more && index != P.last) {
                    // end of a unit, not of the ray: hand the ray over to whoever draws its next unit
                    store_state(J, ray, R, true);
                    st_release_gpu(P.wf_done + ray, (index - P.first) / P.wf_seglen);
                }
            } else {
                S.queries[tid] += 1u;
                zero_fill(P, J, ray, index);                                    // left the mesh, kernel.cpp:372-375
                if (SEG) st_release_gpu(P.wf_done + ray, RVB_WF_DEAD);          // its later units are void
            }
            if (more) {
                S.index[tid] = index;
                sr = BvhAccel::slab_ray(pos, dir);
                tbest = RVB_INF; ibest = RVB_NONE;
                sp = 0; stack.push(sp, RVB_SENTINEL); node = 0;
            } else {
                // reflections executed by this ray so far (the image phase, if any, ran [0, first) in full)
                if (J.count && (index == P.last || ibest == RVB_NONE)) J.count[ray] = index;
                has = false;
            }
        }
        // ---- refill: free lanes take the next rays (SEG: the next units) ----------------------------
        bool want = !has && !pend && !exhausted;
        while (true) {
            const unsigned m = __ballot_sync(FULL, want);
            if (!m) break;
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(fetch, (unsigned)__popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (want) {
                const uint32_t u = base + __popc(m & ((1u << lane) - 1u));
                if (u >= nunits) { want = false; exhausted = true; }
                else {
                    const uint32_t seg = SEG ? u / P.nrays : 0u;
                    const uint32_t r = u - seg * P.nrays;
                    if (SEG) S.end[tid] = min(P.last, P.first + (seg + 1u) * P.wf_seglen);
                    if (seg != 0u) {
                        // a later unit of ray r: it starts from the state its predecessor publishes
                        S.ray[tid] = r; S.index[tid] = P.first + seg * P.wf_seglen;
                        pend = true; want = false;
                    } else {
                        bool live = true;
                        if (P.first == 0) {
                            pos = mk3(J.src[0], J.src[1], J.src[2]);
                            dir = mk3(__ldg(P.dirs + r));
                            S.distance[tid] = 0.0f;
#pragma unroll
                            for (int b = 0; b < RVB_NBANDS; ++b) S.volume[b][tid] = 1.0f;
                        } else {
                            const float4 *s = J.state + 4 * (size_t)r;
                            const float4 a = __ldcg(s), d = __ldcg(s + 1);
                            live = d.w != 0.0f;   // dead: left the mesh during the image phase (slots already zero-filled)
                            if (live) {
                                const float4 va = __ldcg(s + 2), vb = __ldcg(s + 3);
                                pos = mk3(a); dir = mk3(d); S.distance[tid] = a.w;
                                S.volume[0][tid] = va.x; S.volume[1][tid] = va.y; S.volume[2][tid] = va.z; S.volume[3][tid] = va.w;
                                S.volume[4][tid] = vb.x; S.volume[5][tid] = vb.y; S.volume[6][tid] = vb.z; S.volume[7][tid] = vb.w;
                            } else if (SEG) st_release_gpu(P.wf_done + r, RVB_WF_DEAD);
                        }
                        if (live) {
                            S.ray[tid] = r; S.index[tid] = P.first;
                            sr = BvhAccel::slab_ray(pos, dir);
                            tbest = RVB_INF; ibest = RVB_NONE;
                            sp = 0; stack.push(sp, RVB_SENTINEL); node = 0;
                            has = true; want = false;
                        }
                        // dead rays are skipped: keep asking
                    }
                }
            }
        }
        // ---- SEG: pending units whose predecessor has been published start now ----------------------
        if (SEG && pend) {
            const uint32_t r = S.ray[tid];
            const uint32_t need = (S.index[tid] - P.first) / P.wf_seglen;
            const uint32_t done = ld_acquire_gpu(P.wf_done + r);
            if (done >= need) {
                pend = false;
                if (done != RVB_WF_DEAD) {
                    // L2 loads: the record was written by another SM during this launch
                    const float4 *s = J.state + 4 * (size_t)r;
                    const float4 a = __ldcg(s), d = __ldcg(s + 1), va = __ldcg(s + 2), vb = __ldcg(s + 3);
                    pos = mk3(a); dir = mk3(d); S.distance[tid] = a.w;
                    S.volume[0][tid] = va.x; S.volume[1][tid] = va.y; S.volume[2][tid] = va.z; S.volume[3][tid] = va.w;
                    S.volume[4][tid] = vb.x; S.volume[5][tid] = vb.y; S.volume[6][tid] = vb.z; S.volume[7][tid] = vb.w;
                    sr = BvhAccel::slab_ray(pos, dir);
                    tbest = RVB_INF; ibest = RVB_NONE;
                    sp = 0; stack.push(sp, RVB_SENTINEL); node = 0;
                    has = true;
                }
            }
        }
        if (!__any_sync(FULL, has)) {
            if (!SEG) break;
            if (__all_sync(FULL, exhausted)) break;   // nothing held, nothing left to draw
            if (__any_sync(FULL, pend)) {
                __nanosleep(200);
                if (++polls > RVB_WF_POLL_LIMIT) {     // never observed; a hang would cost the caller the device
                    if (lane == 0) atomicExch(P.wf_done + P.nrays, 1u);
                    break;
                }
            }
            continue;
        }

        // ---- traverse ------------------------------------------------------------------------------
        // Every iteration runs ONE of the two code paths for all lanes that are in that state: the
        // triangle path once enough lanes hold a leaf (or nobody holds an inner node), the box path
        // otherwise -- the two never share an iteration, so neither executes with the other's lanes
        // masked off.  The loop ends when enough lanes wait for `serve` (or nothing is left to do).
        while (true) {
            const bool in = has && node >= 0;
            const bool lf = has && node < 0 && node != RVB_SENTINEL;
            const unsigned m_in = __ballot_sync(FULL, in), m_lf = __ballot_sync(FULL, lf);
            const unsigned m_sv = __ballot_sync(FULL, has && node == RVB_SENTINEL);
            if ((m_in | m_lf) == 0u || (uint32_t)__popc(m_sv) >= P.wf_service) break;
            if (m_in == 0u || (uint32_t)__popc(m_lf) >= P.wf_leaf_min) {
                if (lf) {
                    bool found = false;
                    acc.leaf<false>(node, pos, dir, RVB_INF, tbest, ibest, found, ntri);
                    node = stack.pop(sp);
                }
            } else if (in) {
                node = BvhAccel::visit(P.nodes, node, sr, tbest, stack, sp);
                nbox += 4;
            }
        }
    }
    Counters c;
    c.queries = S.queries[tid]; c.bounces = S.bounces[tid]; c.emitted = S.emitted[tid];
    c.tri = (unsigned long long)S.tri[tid] + ntri; c.box = nbox;
    c.tmin_nz = S.tmin_nz[tid]; c.tmax = S.tmax[tid];
    flush_counters(J, c);
  }
}

static int launch_wavefront(const TraceParams &P_in, const TraceJob &J0, size_t stack_bytes, cudaStream_t st, unsigned int **done_out)
{
    *done_out = nullptr;
    static const int minb_env = getenv("RVB_WF_MINB") ? atoi(getenv("RVB_WF_MINB")) : 0;
    static const int service = getenv("RVB_WF_SERVICE") ? atoi(getenv("RVB_WF_SERVICE")) : 24;
    static const int leaf_min = getenv("RVB_WF_LEAF_MIN") ? atoi(getenv("RVB_WF_LEAF_MIN")) : 8;
    static const int grid_env = getenv("RVB_WF_GRID") ? atoi(getenv("RVB_WF_GRID")) : 0;
    static const int nseg_env = getenv("RVB_WF_NSEG") ? atoi(getenv("RVB_WF_NSEG")) : 0;   // 1: whole rays
    TraceParams P = P_in;
    P.wf_service = (uint32_t)std::max(1, std::min(32, service));
    P.wf_leaf_min = (uint32_t)std::max(1, std::min(32, leaf_min));
    P.wf_nseg = 1; P.wf_seglen = P.last - P.first; P.wf_done = nullptr;
    // 5 CTAs of 128 threads per SM (<= 96 registers): 4 and 6 are 3-5 % slower, 7 (72 registers, spills) 23 %
    const int minb = minb_env ? minb_env : 5;
    const bool batch = P.njobs > 1;
    unsigned int *d_fetch = nullptr;   // one ray counter per job
    unsigned int *d_done = nullptr;
    const auto cleanup = on_scope_exit([&] { dev_free(d_fetch); dev_free(d_done); });   // d_done: unless handed to the caller
    RVB_TRY(dev_alloc(&d_fetch, (size_t)P.njobs));
    RVB_CUDA(cudaMemsetAsync(d_fetch, 0, P.njobs * sizeof(unsigned int), st));
    auto go = [&](auto kernel, auto kernel_seg) -> int {
        int per_sm = 0;
        RVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stack_bytes));
        RVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, stack_bytes));
        const size_t want = ((size_t)P.nrays * P.njobs + 127) / 128;
        unsigned grid = (unsigned)std::max<size_t>(1, std::min(want, (size_t)std::max(per_sm, 1) * sm_count()));
        if (grid_env > 0) grid = (unsigned)std::min<size_t>(want, (size_t)grid_env);
        // Segmented rays (see WfShared): only when the rays do not fit the resident lanes at once (otherwise
        // nothing queues) and a whole-ray tail would matter -- aim at >= 16 units per lane, units of >= 8 reflections.
        const uint32_t steps = P.last - P.first;
        const size_t lanes = (size_t)grid * 128;
        const uint32_t forced = nseg_env > 0 ? (uint32_t)nseg_env : ((P.flags >> 8) & 0xffu);
        if (!batch && steps >= 2 && J0.state != nullptr && (size_t)P.nrays * 16 < 0xffffffffull && (forced ? forced > 1 : steps >= 16 && P.nrays > lanes)) {
            uint32_t nseg = forced ? std::min(forced, steps) : (uint32_t)((16 * lanes + P.nrays - 1) / P.nrays);
            if (!forced) nseg = std::max(1u, std::min(nseg, std::min(16u, steps / 8u)));
            nseg = std::min(nseg, 16u);
            if (nseg > 1) {
                P.wf_seglen = (steps + nseg - 1) / nseg;
                P.wf_nseg = (steps + P.wf_seglen - 1) / P.wf_seglen;
            }
        }
        if (P.wf_nseg > 1) {
            RVB_TRY(dev_alloc(&d_done, (size_t)P.nrays + 1));
            RVB_CUDA(cudaMemsetAsync(d_done, 0, ((size_t)P.nrays + 1) * sizeof(unsigned int), st));
            P.wf_done = d_done;
            RVB_CUDA(cudaFuncSetAttribute(kernel_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stack_bytes));
            note_launch(), kernel_seg<<<grid, 128, stack_bytes, st>>>(P, J0, d_fetch);
        } else
            note_launch(), kernel<<<grid, 128, stack_bytes, st>>>(P, J0, d_fetch);
        RVB_CUDA(cudaGetLastError());
        return RVB_OK;
    };
    int rc;
    switch (minb) {
    case 4: rc = batch ? go(trace_bvh_wavefront_kernel<4, true, false>, trace_bvh_wavefront_kernel<4, true, false>)
                       : go(trace_bvh_wavefront_kernel<4, false, false>, trace_bvh_wavefront_kernel<4, false, true>); break;
    case 6: rc = batch ? go(trace_bvh_wavefront_kernel<6, true, false>, trace_bvh_wavefront_kernel<6, true, false>)
                       : go(trace_bvh_wavefront_kernel<6, false, false>, trace_bvh_wavefront_kernel<6, false, true>); break;
    case 7: rc = batch ? go(trace_bvh_wavefront_kernel<7, true, false>, trace_bvh_wavefront_kernel<7, true, false>)
                       : go(trace_bvh_wavefront_kernel<7, false, false>, trace_bvh_wavefront_kernel<7, false, true>); break;
    default: rc = batch ? go(trace_bvh_wavefront_kernel<5, true, false>, trace_bvh_wavefront_kernel<5, true, false>)
                        : go(trace_bvh_wavefront_kernel<5, false, false>, trace_bvh_wavefront_kernel<5, false, true>); break;
    }
    if (rc == RVB_OK) { *done_out = d_done; d_done = nullptr; }   // the caller reads the error flag with its end-of-trace read-back, then frees it
    return rc;
}

static int launch_bvh_images(const TraceParams &P, const TraceJob &J0, size_t smem, cudaStream_t st)
{
    // registers capped at 128 for 4 CTAs per SM instead of 3 (136 registers uncapped, RVB_IMG_MINB=1): 5.25 -> 4.76 ms
    // on C3's 1 Mi rays, 1.24 -> 1.20 ms on a 131 k-ray shard
    static const int minb = getenv("RVB_IMG_MINB") ? atoi(getenv("RVB_IMG_MINB")) : 4;
    // The image phase's rays differ widely in cost (how many image sources survive validation), so its CTAs are
    // dispatched one per 128 rays rather than striding (C3: 4.79 -> 4.32 ms for map + image phase; RVB_IMG_STRIDE=1
    // restores the resident grid)
    static const bool full = !(getenv("RVB_IMG_STRIDE") && *getenv("RVB_IMG_STRIDE") == '1');
    if (P.njobs > 1) return launch_phase(trace_bvh_kernel<true, 1, true>, P, J0, 128, smem, st, full);
    if (minb == 4) return launch_phase(trace_bvh_kernel<true, 4, false>, P, J0, 128, smem, st, full);
    return launch_phase(trace_bvh_kernel<true, 1, false>, P, J0, 128, smem, st, full);
}

int trace_launch(Trace *const *trs, size_t njobs, const float4 *d_dirs, const CandBuffers *cands)
{
    Trace *tr = trs[0];                  // scene, ray / reflection counts and flags are common to the batch
    Scene *sc = tr->scene;
    cudaStream_t st = lib_stream();
    TraceParams P{};
    P.verts = sc->d_verts; P.tri = sc->d_tri; P.quad = sc->d_quad; P.nquads = sc->nquads; P.nrm = sc->d_nrm; P.surfaces = sc->d_surfaces; P.T = sc->T;
    P.nodes = sc->d_nodes; P.leaf_quads = sc->d_leaf_quads; P.leaf_ids = sc->d_leaf_ids;
    P.dirs = d_dirs; P.nrays = (uint32_t)tr->nrays; P.nrefl = (uint32_t)tr->nrefl;
    P.njobs = (uint32_t)njobs;
    // AIR_COEFFICIENT: doubles narrowed to float (rayverb.cpp:632-641)
    const double air[RVB_NBANDS] = {0.001 * -0.1, 0.001 * -0.2, 0.001 * -0.5, 0.001 * -1.1,
                                    0.001 * -2.7, 0.001 * -9.4, 0.001 * -29.0, 0.001 * -60.0};
    for (int b = 0; b < RVB_NBANDS; ++b) P.air[b] = (float)air[b];
    P.flags = tr->flags;
    unsigned long long *d_counters = nullptr;   // six per job; [4] holds the two ordered time keys
    float4 *d_state = nullptr;
    unsigned int *d_wf_done = nullptr;
    TraceJob *d_jobs = nullptr;
    uint16_t *d_perm_keys[2] = {nullptr, nullptr};
    uint32_t *d_perm_ids[2] = {nullptr, nullptr};
    cudaEvent_t e0 = nullptr, e1 = nullptr, em = nullptr;
    std::vector<MicMapStorage> mms(njobs);
    // everything this call allocates is temporary: released on every way out
    const auto cleanup = on_scope_exit([&] {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (em) cudaEventDestroy(em);
        dev_free(d_counters); dev_free(d_state); dev_free(d_jobs); dev_free(d_wf_done);
        for (size_t j = 0; j < njobs; ++j) micmap_free(&mms[j]);
        for (int i = 0; i < 2; ++i) { dev_free(d_perm_keys[i]); dev_free(d_perm_ids[i]); }
    });
    RVB_TRY(dev_alloc(&d_counters, 6 * njobs));
    {
        std::vector<unsigned long long> init(6 * njobs, 0ull);
        for (size_t j = 0; j < njobs; ++j) init[6 * j + 4] = 0x000000007f800000ull;   // ordered(+inf), ordered(0)
        RVB_CUDA(cudaMemcpyAsync(d_counters, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        RVB_CUDA(cudaStreamSynchronize(st));   // `init` is pageable and goes out of scope
    }

    const bool images = !(tr->flags & RVB_TRACE_NO_IMAGES);
    const uint32_t split = images ? (uint32_t)std::min<size_t>(RVB_NUM_IMAGE_SOURCE - 1, tr->nrefl) : 0u;
    const bool brute = sc->accel == RVB_ACCEL_BRUTE_FORCE;
    // BVH scenes, steady phase: the persistent wavefront kernel (launch_wavefront); RVB_WAVEFRONT=0 selects
    // the one-thread-per-ray kernel for comparison.  Measured on the C3 hall with 64-byte nodes
    // (profiles/r2_bvh_kernels.md): 52.5 ms against 61.5 ms at 1 Mi rays, 14.6 against 16.0 at 262 k,
    // 8.1 against 8.2 at 131 k (the per-GPU share of an 8-GPU run).
    static const bool wavefront = !(getenv("RVB_WAVEFRONT") && *getenv("RVB_WAVEFRONT") == '0');
    if (split < tr->nrefl && (split > 0 || !brute)) RVB_TRY(dev_alloc(&d_state, 4 * tr->nrays * njobs));
    // brute force: the whole scene; BVH: the traversal stacks (one column of Scene::bvh_stack entries per thread)
    const size_t smem = brute ? (size_t)9 * sc->nquads * sizeof(float4) : (size_t)sc->bvh_stack * RVB_BVH_CTA * sizeof(int);

    // processing order of the image phase: rays sorted by direction cell (the jobs share the directions)
    if (split > 0 && tr->nrays >= 4096) {
        const uint32_t n = (uint32_t)tr->nrays;
        RVB_TRY(dev_alloc(&d_perm_keys[0], n));
        for (int i = 0; i < 2; ++i) RVB_TRY(dev_alloc(&d_perm_ids[i], n));
        const DirectionCellIn in{d_dirs};
        int cur = 0;
        const int rc_sort = rsort::sort_pairs<256, uint16_t>(in, rsort::NoSink{}, in, n, rsort::plan_passes(0, 16), d_perm_keys, d_perm_ids, true, nullptr,
                                                             nullptr, &cur, nullptr, st);
        RVB_TRY(rc_sort);
        P.ray_perm = d_perm_ids[cur];
    }
    RVB_CUDA(cudaEventCreate(&e0));
    RVB_CUDA(cudaEventCreate(&e1));
    RVB_CUDA(cudaEventCreate(&em));
    std::vector<unsigned long long> h_end(6 * njobs, 0ull), h_mid(6 * njobs, 0ull);
    RVB_CUDA(cudaEventRecord(e0, st));
    // The job records.  BVH scenes: the visibility map around every job's microphone (rebuilt for every
    // trace: it is part of the measured work, not a cache).
    std::vector<TraceJob> jobs(njobs);
    if (!brute && tr->nrays > 0) {
        StageTimer t_mm("trace.micmap");
        std::vector<float> mics(3 * njobs);
        for (size_t j = 0; j < njobs; ++j) for (int i = 0; i < 3; ++i) mics[3 * j + i] = trs[j]->mic[i];
        RVB_TRY(micmap_build_batch(sc, mics.data(), njobs, tr->nrays * std::max<size_t>(tr->nrefl, 1), mms.data(), (double)((tr->flags >> 16) & 0xffu)));
    }
    for (size_t j = 0; j < njobs; ++j) {
        Trace *t = trs[j];
        TraceJob &J = jobs[j];
        for (int i = 0; i < 3; ++i) { J.mic[i] = t->mic[i]; J.src[i] = t->src[i]; }
        J.ray_offset = t->ray_offset;
        J.micmap = mms[j].map;
        J.vol = t->d_vol; J.pt = t->d_pt; J.tkey = t->d_tkey; J.hit = t->d_hit; J.count = t->d_count;
        J.state = d_state ? d_state + 4 * t->nrays * j : nullptr;
        J.cand_volA = cands[j].volA; J.cand_volB = cands[j].volB; J.cand_pt = cands[j].pt; J.cand_idx = cands[j].idx;
        J.counters = d_counters + 6 * j;
        J.time_ext = reinterpret_cast<int *>(d_counters + 6 * j + 4);
        t->stats.micmap_cells = mms[j].map.N; t->stats.micmap_entries = mms[j].entries;
    }
    RVB_TRY(dev_alloc(&d_jobs, njobs));
    RVB_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), njobs * sizeof(TraceJob), cudaMemcpyHostToDevice, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    P.jobs = d_jobs;
    RVB_CUDA(cudaEventRecord(em, st));
    if (tr->nrays > 0 && tr->nrefl > 0) {
        if (split > 0) {
            StageTimer t_im("trace.image_phase");
            P.first = 0; P.last = split;
            if (brute) RVB_TRY(launch_bruteforce<true>(P, jobs[0], smem, st));
            else RVB_TRY(launch_bvh_images(P, jobs[0], smem, st));
            RVB_CUDA(cudaEventRecord(em, st));
            RVB_CUDA(cudaMemcpyAsync(h_mid.data(), d_counters, 6 * njobs * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        }
        if (split < tr->nrefl) {
            StageTimer t_st("trace.steady_phase");
            P.first = split; P.last = (uint32_t)tr->nrefl;
            if (brute) RVB_TRY(launch_bruteforce<false>(P, jobs[0], smem, st));
            else if (!wavefront) {   // one thread per ray, 6 CTAs of 128 threads per SM (80 registers)
                if (njobs > 1) RVB_TRY(launch_phase(trace_bvh_kernel<false, 6, true>, P, jobs[0], 128, smem, st));
                else RVB_TRY(launch_phase(trace_bvh_kernel<false, 6, false>, P, jobs[0], 128, smem, st));
            }
            else RVB_TRY(launch_wavefront(P, jobs[0], smem, st, &d_wf_done));
        }
    } else if (tr->nrays > 0 && images) {
        // zero reflections: the reference still computes the direct path for ray 0
        P.first = 0; P.last = 0;
        if (brute) RVB_TRY(launch_bruteforce<true>(P, jobs[0], smem, st));
        else RVB_TRY(launch_bvh_images(P, jobs[0], smem, st));
    }
    RVB_CUDA(cudaEventRecord(e1, st));
    RVB_CUDA(cudaMemcpyAsync(h_end.data(), d_counters, 6 * njobs * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    unsigned int wf_error = 0;   // segmented wavefront kernel: a pending unit ran out of its poll budget
    if (d_wf_done) RVB_CUDA(cudaMemcpyAsync(&wf_error, d_wf_done + tr->nrays, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f, ms_image = 0.f, ms_steady = 0.f;
    RVB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    RVB_CUDA(cudaEventElapsedTime(&ms_image, e0, em));
    RVB_CUDA(cudaEventElapsedTime(&ms_steady, em, e1));
    for (size_t j = 0; j < njobs; ++j) {
        Trace *t = trs[j];
        const unsigned long long *h = h_end.data() + 6 * j, *hm = h_mid.data() + 6 * j;
        // kernel times are those of the whole launch (all jobs of the batch run in the same kernels)
        t->stats.image_phase_ms = ms_image; t->stats.steady_phase_ms = ms_steady; t->stats.trace_kernel_ms = ms;
        t->stats.steady_tri_tests = h[1] - hm[1]; t->stats.steady_box_tests = h[2] - hm[2];
        int ext[2];
        memcpy(ext, &h[4], sizeof ext);
        t->diffuse_tmin_nz = unordered_key(ext[0]);   // +inf: no non-zero time
        t->diffuse_tmax = unordered_key(ext[1]);
        t->has_time_ext = true;
        t->stats.queries = h[0]; t->stats.tri_tests = h[1]; t->stats.box_tests = h[2]; t->stats.bounces = h[3];
        t->stats.emitted = h[5];
    }
    if (wf_error) return fail(RVB_ERR_CUDA, "wavefront kernel: a work unit waited for its predecessor beyond the poll budget");
    return RVB_OK;
}

}  // namespace rvb
