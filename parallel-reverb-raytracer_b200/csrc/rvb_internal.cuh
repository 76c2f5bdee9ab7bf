// Internal declarations shared by the translation units of librayverb_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cstring>
#include <vector>

#include "../../include/rayverb_b200.h"
#include "rvb_math.cuh"
#include "micmap.cuh"

namespace rvb {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
#define RVB_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return ::rvb::fail(RVB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define RVB_TRY(call)                 \
    do {                              \
        int rc_ = (call);             \
        if (rc_ != RVB_OK) return rc_; \
    } while (0)

// One library stream per device (non-blocking), created lazily; events for stage timings.
cudaStream_t lib_stream();
void note_launch();    // counts kernels launched by this library (rvb_launch_count)
int sm_count();        // multiprocessor count of the current device (cached)

// RVB_TIMING=1: host wall-clock per stage (stream-synchronised at scope exit), printed to stderr
struct StageTimer {
    const char *name;
    double t0;
    bool on;
    explicit StageTimer(const char *n);
    ~StageTimer();
};
int ensure_device();   // checks that the current device exists and is sm_100; configures the mem pool

// Device allocations.  Small blocks come from the device's stream-ordered pool; blocks of 1 MiB
// and more go through an exact-size cache (a trace of the same shape re-uses the very same
// blocks on the next call instead of paying cudaMalloc / pool growth -- the reference allocates
// its cl::Buffers once per Raytracer for the same reason, rayverb.cpp:252-274).
int alloc_bytes(void **p, size_t bytes);
void free_bytes(void *p);
template <class T>
inline int dev_alloc(T **p, size_t n)
{
    *p = nullptr;
    if (n == 0) return RVB_OK;
    return alloc_bytes((void **)p, n * sizeof(T));
}
template <class T>
inline void dev_free(T *&p)
{
    if (p) free_bytes((void *)p);
    p = nullptr;
}

// Runs `f` when the enclosing scope ends -- on every early return of RVB_TRY / RVB_CUDA too: the
// temporary device blocks and events of a call are released on its error paths as well.
template <class F>
struct ScopeExit {
    F f;
    explicit ScopeExit(F g) : f(g) {}
    ~ScopeExit() { f(); }
    ScopeExit(const ScopeExit &) = delete;
    ScopeExit &operator=(const ScopeExit &) = delete;
};
template <class F>
inline ScopeExit<F> on_scope_exit(F f) { return ScopeExit<F>(f); }

// order-preserving float <-> int keys for atomicMin / atomicMax on floats
__host__ __device__ inline int ordered_key(float f) { int i; memcpy(&i, &f, 4); return i ^ ((i >> 31) & 0x7fffffff); }
__host__ __device__ inline float unordered_key(int i) { i = i ^ ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &i, 4); return f; }

// Device -> host read-backs on the hot path go through a per-thread PINNED scratch buffer.  A
// cudaMemcpyAsync into pageable memory blocks inside the driver, holding its context lock, until the
// stream has drained -- which serialises every other host thread of the process (a parameter sweep
// with one stream per thread ran its pairs strictly one after the other because of it).
int pinned_scratch(void **p, size_t bytes);                       // grown on demand, valid until the next call of this thread
int read_back(void *dst, const void *d_src, size_t bytes);        // copy + synchronise the library stream of this thread
int trim_cache();                                                // idle cached blocks of the current device back to the driver

// ---- BVH node (see bvh.cu) ------------------------------------------------------------------
// 4-wide node in HALF a cache line (64 bytes, two 32-byte loads).  The bounce kernels are bound by the
// L1 data pipe, which moves a divergent warp's node fetches 16-32 bytes per lane and cycle (ncu:
// l1tex data-pipe wavefronts at 85-90 % of peak with 128-byte nodes), so the child boxes are stored
// quantised: 8 bits per plane on a grid spanned by the node's own box -- origin (3 floats) and one
// power-of-two cell size per axis -- rounded OUTWARDS, so a quantised box contains the exact one.
//   plane = origin + q * 2^(e - 127)
// q bytes are grouped by plane (qlo[axis] / qhi[axis] hold the four children's bytes) so that one PRMT
// per value drops a byte into a float's mantissa (trace.cu: BvhAccel::visit).
// Child ref: >= 0 -> inner node index; < 0 (and != RVB_BVH_EMPTY) -> leaf:
// ~ref = leaf record | (count - 1) << 28.  Unused slots carry RVB_BVH_EMPTY.
#define RVB_BVH_EMPTY ((int)0x80000001)
#define RVB_LEAF_STRIDE 10           // float4 per leaf record: 9 of triangle data + 1 of triangle indices (160 bytes)
#define RVB_BVH_STACK 96             // traversal stack entries per thread (bvh.cu refuses deeper hierarchies)
struct __align__(32) BvhNode {
    float origin[3];
    uint32_t exps;            // biased exponents of the cell sizes: x | y << 8 | z << 16
    uint32_t qlo[3], qhi[3];  // byte k of qlo[a] / qhi[a]: child k's lower / upper plane on axis a
    int ref[4];
    uint32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode is half a cache line");

// ---- scene ----------------------------------------------------------------------------------
struct Scene {
    int device = 0;
    uint32_t T = 0, S = 0;
    int accel = RVB_ACCEL_BRUTE_FORCE;
    float4 *d_verts = nullptr;        // [3T] v0, v1, v2 (w = 0): TriVerts of kernel.cpp:56-60, pre-gathered
    float4 *d_tri = nullptr;          // [3T] staged intersection form: (v0.xyz,e0.x) (e0.yz,e1.xy) (e1.z,0,0,0)
    float4 *d_quad = nullptr;         // [9*nquads] brute-force form: 4 triangles per 144-byte record, two component-wise
                                      //   PAIRS (v0x_a,v0x_b, v0y_a,v0y_b, ... e1z_a,e1z_b)(same for c,d); zero triangles pad
    uint32_t nquads = 0;
    float4 *d_nrm = nullptr;          // [T]  normalize(cross(e0,e1)) (kernel.cpp:110-116), w = surface index bits
    rvb_surface *d_surfaces = nullptr;
    // BVH (accel == RVB_ACCEL_BVH)
    BvhNode *d_nodes = nullptr;       // [n_nodes]
    uint32_t n_nodes = 0;
    float4 *d_leaf_quads = nullptr;   // [RVB_LEAF_STRIDE*n_leaves] leaf records: <= 4 triangles in the pair-interleaved form of
                                      //   d_quad (36 floats), then the callers' indices of the four slots (0xffffffff = unused)
    uint4 *d_leaf_ids = nullptr;      // (unused)
    uint32_t n_leaves = 0;
    int bvh_depth = 0;                // binary depth of the hierarchy
    int bvh_stack = 0;                // traversal stack entries a query can need (<= RVB_BVH_STACK)
    float diag = 0.f;                 // scene diagonal (sizes the conservative paddings)
    float build_ms = 0.f;
};

// ---- impulse slots ----------------------------------------------------------------------------
// Diffuse impulses live in a bounce-major SoA whose unit is a GROUP of four consecutive reflections
// of one ray: slot = ((bounce / 4) * nrays + ray) * 4 + (bounce & 3), so the four 32-byte volume
// sectors of a group fill exactly one 128-byte line.  Why: an L2 miss on B200 fetches the whole
// 128-byte line from HBM (tools/gather_bench.cu: 125 B of DRAM traffic per random 32-byte gather),
// and the histogram reduction gathers impulses at random while it walks the sample bins in order.
// The four reflections of a group arrive ~500 samples apart, i.e. within a few tens of megabytes of
// reduction traffic, so three of the four gathers of a line hit L2.  The 32 lanes of a warp (same
// bounce, consecutive rays) still write whole 32-byte sectors.  When the reflection count is not a
// multiple of four the unused slots of the last group are marked as skipped (tkey = -1).
#define RVB_SLOT_GROUP 4u
__host__ __device__ inline size_t slot_of(uint32_t ray, uint32_t bounce, size_t nrays)
{
    return ((size_t)(bounce / RVB_SLOT_GROUP) * nrays + ray) * RVB_SLOT_GROUP + (bounce % RVB_SLOT_GROUP);
}
__host__ __device__ inline size_t slot_count(size_t nrays, size_t nrefl) { return (nrefl + RVB_SLOT_GROUP - 1) / RVB_SLOT_GROUP * RVB_SLOT_GROUP * nrays; }

// ---- trace results ----------------------------------------------------------------------------
struct Trace {
    Scene *scene = nullptr;
    size_t nrays = 0, nrefl = 0;
    uint64_t ray_offset = 0;
    unsigned flags = 0;
    float mic[3] = {0, 0, 0}, src[3] = {0, 0, 0};
    // diffuse impulses, bounce-GROUP-major: slot_of(ray, bounce, nrays), slot_count() slots
    float4 *d_vol = nullptr;                        // [2*slots]: bands 0-3 at 2*slot, 4-7 at 2*slot+1 (one 32-B sector)
    float4 *d_pt = nullptr;                        // position xyz, time
    float *d_tkey = nullptr;                       // time if any(volume != 0) else -1 (kernel.cpp:524 skip rule)
    uint32_t *d_hit = nullptr;                     // debug: hit triangle per slot
    uint32_t *d_count = nullptr;                   // debug: reflections per ray
    // unique image-source impulses in key order (after dedupe)
    size_t n_images = 0;
    float4 *d_img_volA = nullptr, *d_img_volB = nullptr, *d_img_pt = nullptr;
    std::vector<uint32_t> img_keys;                // n_images * 10
    std::vector<int32_t> img_len;                  // n_images
    bool first_is_direct = false;                  // entry 0 has key {0}
    // extrema of the emitted diffuse impulse times, reduced inside the bounce kernels
    bool has_time_ext = false;
    float diffuse_tmin_nz = 0.f, diffuse_tmax = 0.f;
    rvb_trace_stats stats{};
};

struct Ir {
    std::vector<std::vector<float>> channels;      // mixed-down output
    std::vector<std::vector<float>> bands;         // [channel] -> [8][band_len] filtered, before mixdown
    std::vector<size_t> band_len;
};

// ---- microphone visibility map (micmap.cu) ----------------------------------------------------------
struct MicMapStorage {
    MicMap map;                       // what the kernels see (N == 0: no map, generic occlusion queries)
    uint32_t *d_cell_off = nullptr, *d_cell_ent = nullptr, *d_near = nullptr;
    uint32_t entries = 0;
};
int micmap_build(const Scene *sc, const float mic[3], size_t nqueries, MicMapStorage *out);
int micmap_build_batch(const Scene *sc, const float *mics /* [n][3] */, size_t n, size_t nqueries, MicMapStorage *out /* [n] */,
                       double forced_cells_per_tri = 0.0 /* RVB_TRACE_MAP_CELLS */);
void micmap_free(MicMapStorage *s);

// ---- stage entry points (defined in the .cu files) ----------------------------------------------
int scene_build_bvh(Scene *sc);                                                       // bvh.cu
// image-source candidate slots of one job (freed after de-duplication)
struct CandBuffers { float4 *volA = nullptr, *volB = nullptr, *pt = nullptr; uint32_t *idx = nullptr; };
// traces all jobs (same scene, directions, counts and flags; their own microphone / source) in one launch pair
int trace_launch(Trace *const *trs, size_t njobs, const float4 *d_dirs, const CandBuffers *cands);   // trace.cu
int images_dedupe(Trace *tr, const float4 *d_cand_volA, const float4 *d_cand_volB,
                  const float4 *d_cand_pt, const uint32_t *d_cand_idx);               // images.cu

}  // namespace rvb

struct rvb_scene : rvb::Scene {};
struct rvb_trace : rvb::Trace {};
struct rvb_ir : rvb::Ir {};
