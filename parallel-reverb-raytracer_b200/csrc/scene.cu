// Context, error plumbing and scene staging.
//
// rvb_scene_create replaces the Raytracer constructor (rayverb/rayverb.cpp:242-293): instead of
// uploading Triangle[] / vertices[] for per-test gathers through u64 indices
// (kernel.cpp:95-107), every triangle is gathered ONCE into
//   d_verts  (v0, v1, v2)                       -- TriVerts for the image-source mirroring
//   d_tri    (v0, e0 = v1-v0, e1 = v2-v0)       -- intersection form, float4-vectorised
//   d_nrm    normalize(cross(e0, e1)), surface  -- kernel.cpp:110-116 evaluated once
// The subtraction / normal are the same IEEE operations the reference performs per test, so the
// staged values carry the same bits.
#include <algorithm>
#include <cstring>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <map>
#include <mutex>
#include <unordered_map>

#include "rvb_internal.cuh"

namespace rvb {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) { g_error = msg; return code; }
const char *last_error() { return g_error.c_str(); }

static std::mutex g_mutex;
static cudaStream_t g_streams[64] = {};
static bool g_checked[64] = {};

static thread_local cudaStream_t g_user_stream = nullptr;
static thread_local bool g_has_user_stream = false;
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

cudaStream_t lib_stream()
{
    if (g_has_user_stream) return g_user_stream;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!g_streams[dev]) cudaStreamCreateWithFlags(&g_streams[dev], cudaStreamNonBlocking);
    return g_streams[dev];
}

// ---- exact-size block cache ----------------------------------------------------------------------
// `ready` is recorded on the freeing stream: a later owner on another stream waits for it ON THE
// DEVICE (cudaStreamWaitEvent) -- host threads that drive different streams never block each other
struct CachedBlock { void *p; cudaStream_t stream; cudaEvent_t ready; };
static std::multimap<size_t, CachedBlock> g_free_blocks[64];
static std::unordered_map<void *, size_t> g_live_blocks[64];
static size_t g_cached_bytes[64] = {};
static const size_t kCacheThreshold = (size_t)1 << 20;
// beyond this many bytes of IDLE blocks (per device), the largest go back to the driver; RVB_CACHE_LIMIT_MB
// overrides the 48 GiB default (a process that shares the GPU with torch's allocator may want less), and
// rvb_trim_cache() returns all of them at once
static size_t cache_limit()
{
    static const size_t v = getenv("RVB_CACHE_LIMIT_MB") ? (size_t)strtoull(getenv("RVB_CACHE_LIMIT_MB"), nullptr, 10) << 20 : (size_t)48 << 30;
    return v;
}

// Small blocks: a stream-ordered memory pool PER HOST THREAD (and device).  With the device's default
// pool, a block freed on one thread's stream and re-used on another's makes the driver chain the two
// streams together, and the worker threads of a parameter sweep end up running one after the other.
// A pool outlives the host thread that used it: traces made on a worker thread hold small blocks that
// another thread frees later, and destroying a pool at thread exit with such blocks outstanding crashed
// inside the runtime (found by tests/test_gpu_fullsize.py: worker threads == sequential).  A thread that
// ends hands its pools to an idle list; the next new thread adopts one from there.
static std::vector<cudaMemPool_t> g_idle_pools[64];
struct ThreadPools {
    cudaMemPool_t pool[64] = {};
    bool own[64] = {};
    ~ThreadPools()
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        for (int d = 0; d < 64; ++d) if (pool[d] && own[d]) g_idle_pools[d].push_back(pool[d]);
    }
};
static thread_local ThreadPools g_tp;
static cudaMemPool_t thread_pool(int dev)
{
    cudaMemPool_t *g_thread_pool = g_tp.pool;
    if (!g_thread_pool[dev]) {
        {
            std::lock_guard<std::mutex> lk(g_mutex);
            if (!g_idle_pools[dev].empty()) {
                g_thread_pool[dev] = g_idle_pools[dev].back();
                g_idle_pools[dev].pop_back();
                g_tp.own[dev] = true;
                return g_thread_pool[dev];
            }
        }
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&g_thread_pool[dev], &props) != cudaSuccess) {
            cudaGetLastError();
            cudaDeviceGetDefaultMemPool(&g_thread_pool[dev], dev);
        } else {
            g_tp.own[dev] = true;
            uint64_t keep = UINT64_MAX;   // keep freed blocks cached
            cudaMemPoolSetAttribute(g_thread_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return g_thread_pool[dev];
}

int alloc_bytes(void **p, size_t bytes)
{
    cudaStream_t st = lib_stream();
    if (bytes < kCacheThreshold) {
        int pool_dev = 0;
        cudaGetDevice(&pool_dev);
        cudaError_t e = cudaMallocFromPoolAsync(p, bytes, thread_pool(pool_dev), st);
        if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? RVB_ERR_NOMEM : RVB_ERR_CUDA,
                                          std::string("cudaMallocFromPoolAsync: ") + cudaGetErrorString(e));
        return RVB_OK;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        auto range = g_free_blocks[dev].equal_range(bytes);
        auto it = range.first;
        for (auto j = range.first; j != range.second; ++j)
            if (j->second.stream == st) { it = j; break; }          // a block last used on this very stream needs no wait
        if (it != range.second) {
            CachedBlock b = it->second;
            g_free_blocks[dev].erase(it);
            g_cached_bytes[dev] -= bytes;
            g_live_blocks[dev][b.p] = bytes;
            *p = b.p;
            if (b.stream != st) cudaStreamWaitEvent(st, b.ready, 0);   // freed on another stream: order after its last use
            cudaEventDestroy(b.ready);
            return RVB_OK;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {   // give cached blocks back and retry once
        cudaGetLastError();
        std::vector<void *> drop;
        {
            std::lock_guard<std::mutex> lk(g_mutex);
            for (auto &kv : g_free_blocks[dev]) { drop.push_back(kv.second.p); cudaEventDestroy(kv.second.ready); }
            g_free_blocks[dev].clear();
            g_cached_bytes[dev] = 0;
        }
        cudaDeviceSynchronize();
        for (void *q : drop) cudaFree(q);
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? RVB_ERR_NOMEM : RVB_ERR_CUDA,
                                      std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lk(g_mutex);
    g_live_blocks[dev][*p] = bytes;
    return RVB_OK;
}

void free_bytes(void *p)
{
    cudaStream_t st = lib_stream();
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<void *> drop;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        auto it = g_live_blocks[dev].find(p);
        if (it == g_live_blocks[dev].end()) {
            cudaFreeAsync(p, st);
            return;
        }
        const size_t bytes = it->second;
        g_live_blocks[dev].erase(it);
        cudaEvent_t ready;
        cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
        cudaEventRecord(ready, st);
        g_free_blocks[dev].emplace(bytes, CachedBlock{p, st, ready});
        g_cached_bytes[dev] += bytes;
        while (g_cached_bytes[dev] > cache_limit() && !g_free_blocks[dev].empty()) {
            auto big = std::prev(g_free_blocks[dev].end());
            drop.push_back(big->second.p);
            cudaEventDestroy(big->second.ready);
            g_cached_bytes[dev] -= big->first;
            g_free_blocks[dev].erase(big);
        }
    }
    if (!drop.empty()) {
        cudaDeviceSynchronize();
        for (void *q : drop) cudaFree(q);
    }
}

// Give every idle cached block of the current device back to the driver (live blocks are untouched).
int trim_cache()
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<void *> drop;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        for (auto &kv : g_free_blocks[dev]) { drop.push_back(kv.second.p); cudaEventDestroy(kv.second.ready); }
        g_free_blocks[dev].clear();
        g_cached_bytes[dev] = 0;
    }
    if (!drop.empty()) {
        RVB_CUDA(cudaDeviceSynchronize());   // their last users may still be running
        for (void *q : drop) cudaFree(q);
    }
    return RVB_OK;
}

static thread_local void *g_scratch = nullptr;
static thread_local size_t g_scratch_bytes = 0;
int pinned_scratch(void **p, size_t bytes)
{
    if (bytes > g_scratch_bytes) {
        if (g_scratch) cudaFreeHost(g_scratch);
        g_scratch = nullptr; g_scratch_bytes = 0;
        const size_t cap = std::max<size_t>(bytes, (size_t)1 << 20);
        RVB_CUDA(cudaHostAlloc(&g_scratch, cap, cudaHostAllocDefault));
        g_scratch_bytes = cap;
    }
    *p = g_scratch;
    return RVB_OK;
}
int read_back(void *dst, const void *d_src, size_t bytes)
{
    if (bytes == 0) return RVB_OK;
    void *h = nullptr;
    RVB_TRY(pinned_scratch(&h, bytes));
    cudaStream_t st = lib_stream();
    RVB_CUDA(cudaMemcpyAsync(h, d_src, bytes, cudaMemcpyDeviceToHost, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    memcpy(dst, h, bytes);
    return RVB_OK;
}

static int g_sm_count[64] = {};
int sm_count()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (!g_sm_count[dev]) cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    return g_sm_count[dev];
}

static bool timing_on()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("RVB_TIMING"); on = (e && *e && *e != '0') ? 1 : 0; }
    return on == 1;
}
static double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
StageTimer::StageTimer(const char *n) : name(n), t0(0), on(timing_on())
{
    if (on) { cudaStreamSynchronize(lib_stream()); t0 = now_ms(); }
}
StageTimer::~StageTimer()
{
    if (on) { cudaStreamSynchronize(lib_stream()); fprintf(stderr, "[rvb] %-28s %8.3f ms\n", name, now_ms() - t0); }
}

// RVB_DEBUG_SIGNALS=1: print a native backtrace on SIGFPE / SIGSEGV (debug aid; resolve the offsets
// with addr2line -e librayverb_b200.so)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void debug_signal_handler(int sig)
{
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "[rvb] fatal signal, native backtrace:\n";
    if (write(2, msg, sizeof msg - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
static void maybe_install_debug_signals()
{
    static bool done = false;
    if (done) return;
    done = true;
    const char *e = getenv("RVB_DEBUG_SIGNALS");
    if (e && *e == '1') { signal(SIGFPE, debug_signal_handler); signal(SIGSEGV, debug_signal_handler); }
}

int ensure_device()
{
    maybe_install_debug_signals();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(RVB_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    int dev = 0;
    RVB_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        if (g_checked[dev]) return RVB_OK;
    }
    cudaDeviceProp prop;
    RVB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(RVB_ERR_CUDA, std::string("device '") + prop.name + "' is not sm_100 (kernels are built for sm_100a only)");
    cudaMemPool_t pool;
    RVB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = UINT64_MAX;   // keep freed blocks cached: traces of the same shape re-use them
    RVB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    std::lock_guard<std::mutex> lk(g_mutex);
    g_checked[dev] = true;
    return RVB_OK;
}

__global__ void stage_scene_kernel(const rvb_triangle *tris, const float4 *vertices, uint32_t T, float4 *verts, float4 *tri,
                                   float4 *nrm)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    rvb_triangle t = tris[i];
    float4 a = vertices[t.v0], b = vertices[t.v1], c = vertices[t.v2];
    a.w = b.w = c.w = 0.f;
    verts[3 * i] = a; verts[3 * i + 1] = b; verts[3 * i + 2] = c;
    f3 v0 = mk3(a), e0 = mk3(b) - mk3(a), e1 = mk3(c) - mk3(a);
    tri[3 * i] = make_float4(v0.x, v0.y, v0.z, e0.x);
    tri[3 * i + 1] = make_float4(e0.y, e0.z, e1.x, e1.y);
    tri[3 * i + 2] = make_float4(e1.z, 0.f, 0.f, 0.f);
    f3 n = normalize(cross(e0, e1));
    nrm[i] = make_float4(n.x, n.y, n.z, __uint_as_float((uint32_t)t.surface));
}

// Pair-interleaved copy of the intersection form for the packed-FP32 brute-force loop: record q
// holds triangles 4q..4q+3 as two pairs of 18 floats, component-wise (x = first, y = second
// triangle of the pair).  Triangles past T are all-zero: det = 0 rejects them (kernel.cpp:72).
__global__ void pack_quads_kernel(const float4 *tri, uint32_t T, uint32_t nquads, float *quad)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;   // pair index
    if (p >= 2 * nquads) return;
    float a[9], b[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = b[k] = 0.f;
    const float *src = reinterpret_cast<const float *>(tri);
    if (2 * p < T) for (int k = 0; k < 9; ++k) a[k] = src[12 * (2 * p) + k];
    if (2 * p + 1 < T) for (int k = 0; k < 9; ++k) b[k] = src[12 * (2 * p + 1) + k];
    for (int k = 0; k < 9; ++k) { quad[18 * p + 2 * k] = a[k]; quad[18 * p + 2 * k + 1] = b[k]; }
}

}  // namespace rvb

using namespace rvb;

extern "C" {

const char *rvb_last_error(void) { return rvb::last_error(); }
int rvb_version(void) { return 100; }

int rvb_device_count(int *count)
{
    if (!count) return fail(RVB_ERR_INVALID, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(RVB_ERR_CUDA, cudaGetErrorString(e)); }
    *count = n;
    return RVB_OK;
}

int rvb_set_stream(void *cuda_stream, int use)
{
    g_user_stream = (cudaStream_t)cuda_stream;
    g_has_user_stream = use != 0;
    return RVB_OK;
}

unsigned long long rvb_launch_count(void) { return g_launches.load(); }
int rvb_trim_cache(void) { return rvb::trim_cache(); }

int rvb_set_device(int device)
{
    RVB_CUDA(cudaSetDevice(device));
    return ensure_device();
}

int rvb_scene_create(const rvb_triangle *triangles, size_t ntriangles, const rvb_float3 *vertices, size_t nvertices,
                     const rvb_surface *surfaces, size_t nsurfaces, int accel, rvb_scene **out)
{
    if (!out) return fail(RVB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!triangles || !vertices || !surfaces || ntriangles == 0 || nvertices == 0 || nsurfaces == 0)
        return fail(RVB_ERR_INVALID, "scene needs at least one triangle, vertex and surface");
    if (ntriangles >= (1u << 28)) return fail(RVB_ERR_UNSUPPORTED, "more than 2^28 triangles");
    // The reference never validates indices (SceneData::valid() is dead code, rayverb.cpp:463-502)
    // and out-of-range ones are undefined behaviour in its kernel; here they are an error.
    for (size_t i = 0; i < ntriangles; ++i) {
        const rvb_triangle &t = triangles[i];
        if (t.surface >= nsurfaces || t.v0 >= nvertices || t.v1 >= nvertices || t.v2 >= nvertices)
            return fail(RVB_ERR_INVALID, "triangle " + std::to_string(i) + " references an out-of-range vertex or surface");
    }
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    rvb_scene *sc = new rvb_scene();
    cudaGetDevice(&sc->device);
    sc->T = (uint32_t)ntriangles;
    sc->S = (uint32_t)nsurfaces;
    rvb_triangle *d_tris = nullptr;
    float4 *d_vertices = nullptr;
    int rc = RVB_OK;
    auto cleanup = [&]() { dev_free(d_tris); dev_free(d_vertices); };
    if ((rc = dev_alloc(&d_tris, ntriangles)) || (rc = dev_alloc(&d_vertices, nvertices)) ||
        (rc = dev_alloc(&sc->d_verts, 3 * ntriangles)) || (rc = dev_alloc(&sc->d_tri, 3 * ntriangles)) ||
        (rc = dev_alloc(&sc->d_nrm, ntriangles)) || (rc = dev_alloc(&sc->d_quad, 9 * (size_t)((ntriangles + 3) / 4))) || (rc = dev_alloc(&sc->d_surfaces, nsurfaces))) {
        cleanup(); rvb_scene_destroy(sc); return rc;
    }
    cudaMemcpyAsync(d_tris, triangles, ntriangles * sizeof(rvb_triangle), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_vertices, vertices, nvertices * sizeof(rvb_float3), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(sc->d_surfaces, surfaces, nsurfaces * sizeof(rvb_surface), cudaMemcpyHostToDevice, st);
    note_launch(), stage_scene_kernel<<<(unsigned)((ntriangles + 255) / 256), 256, 0, st>>>(d_tris, d_vertices, sc->T, sc->d_verts, sc->d_tri, sc->d_nrm);
    sc->nquads = (uint32_t)((ntriangles + 3) / 4);
    note_launch(), pack_quads_kernel<<<(2 * sc->nquads + 255) / 256, 256, 0, st>>>(sc->d_tri, sc->T, sc->nquads, reinterpret_cast<float *>(sc->d_quad));
    cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) { rvb_scene_destroy(sc); return fail(RVB_ERR_CUDA, std::string("scene staging: ") + cudaGetErrorString(e)); }

    const size_t smem_limit = 227 * 1024;
    const bool fits = (size_t)9 * sc->nquads * sizeof(float4) + 64 <= smem_limit;
    if (accel == RVB_ACCEL_BRUTE_FORCE && !fits) {
        rvb_scene_destroy(sc);
        return fail(RVB_ERR_INVALID, "brute-force path needs the scene in shared memory (<= 6452 triangles)");
    }
    if (accel == RVB_ACCEL_AUTO) accel = (ntriangles <= 1024) ? RVB_ACCEL_BRUTE_FORCE : RVB_ACCEL_BVH;
    sc->accel = accel;
    if (accel == RVB_ACCEL_BVH) {
        rc = scene_build_bvh(sc);
        if (rc) { rvb_scene_destroy(sc); return rc; }
    }
    *out = sc;
    return RVB_OK;
}

int rvb_scene_destroy(rvb_scene *sc)
{
    if (!sc) return RVB_OK;
    cudaSetDevice(sc->device);
    dev_free(sc->d_verts); dev_free(sc->d_tri); dev_free(sc->d_quad); dev_free(sc->d_nrm); dev_free(sc->d_surfaces);
    dev_free(sc->d_nodes); dev_free(sc->d_leaf_quads); dev_free(sc->d_leaf_ids);
    cudaStreamSynchronize(lib_stream());
    delete sc;
    return RVB_OK;
}

int rvb_scene_accel(const rvb_scene *scene, int *accel)
{
    if (!scene || !accel) return fail(RVB_ERR_INVALID, "null argument");
    *accel = scene->accel;
    return RVB_OK;
}

}  // extern "C"
