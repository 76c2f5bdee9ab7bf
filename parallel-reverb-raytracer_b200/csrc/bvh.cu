// GPU-built BVH for large scenes (the reference has no acceleration structure: its nearest-hit is
// the brute-force loop of kernel.cpp:167-192, which this structure must reproduce exactly).
//
// Build: LBVH (Karras 2012) -- 63-bit Morton codes of the triangle centroids, one radix sort, one
// thread per internal node for the hierarchy, bottom-up box refit with atomic visit counters --
// then subtrees of <= 4 triangles are collapsed into multi-triangle leaves.  Triangles are stored
// again in sorted order (v0, e0, e1 -- the same staged intersection form as the brute-force path)
// so a leaf reads consecutive memory; `order` maps back to the caller's triangle index, which is
// what hit results and the (t, index) tie-break use.
//
// Boxes are padded: the float Moller-Trumbore test can accept a hit whose computed parameters are
// off by a few 1e-7 * |origin - v0| / sin(grazing angle); a padding proportional to the scene
// diagonal, which the slab test sees magnified by the same 1/sin factor, keeps every triangle the
// exact test could accept inside its box.
#include <cub/cub.cuh>

#include "rvb_internal.cuh"

namespace rvb {

#define RVB_LEAF_MAX 4

struct Aabb { float3 lo, hi; };

__device__ __forceinline__ int ord(float f) { int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__host__ __device__ inline float unord(int i) { i = i ^ ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &i, 4); return f; }

__global__ void tri_bounds_kernel(const float4 *verts, uint32_t T, Aabb *boxes, float4 *centroids, int *scene /* lo xyz, hi xyz */)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = make_float3(3e38f, 3e38f, 3e38f), hi = make_float3(-3e38f, -3e38f, -3e38f);
    if (i < T) {
        float4 a = verts[3 * i], b = verts[3 * i + 1], c = verts[3 * i + 2];
        lo = make_float3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
        hi = make_float3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
        boxes[i].lo = lo; boxes[i].hi = hi;
        centroids[i] = make_float4(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z), 0.f);
    }
    for (int o = 16; o; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(scene + 0, ord(lo.x)); atomicMin(scene + 1, ord(lo.y)); atomicMin(scene + 2, ord(lo.z));
        atomicMax(scene + 3, ord(hi.x)); atomicMax(scene + 4, ord(hi.y)); atomicMax(scene + 5, ord(hi.z));
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void morton_kernel(const float4 *centroids, uint32_t T, float3 lo, float3 inv_extent, unsigned long long *codes, uint32_t *ids)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    float4 c = centroids[i];
    const float S = 2097151.0f;   // 2^21 - 1
    unsigned long long x = (unsigned long long)fminf(fmaxf((c.x - lo.x) * inv_extent.x * S, 0.f), S);
    unsigned long long y = (unsigned long long)fminf(fmaxf((c.y - lo.y) * inv_extent.y * S, 0.f), S);
    unsigned long long z = (unsigned long long)fminf(fmaxf((c.z - lo.z) * inv_extent.z * S, 0.f), S);
    codes[i] = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
    ids[i] = i;
}

// common-prefix length of sorted keys i and j (ties broken by position), -1 outside the array
__device__ __forceinline__ int delta(const unsigned long long *codes, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    unsigned long long a = codes[i], b = codes[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, one thread per internal node.  Children: index >= 0 internal, ~pos leaf.
__global__ void hierarchy_kernel(const unsigned long long *codes, int n, int2 *children, int2 *ranges, int *parent_inner, int *parent_leaf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(codes, n, i, i - d);
    int lmax = 2;
    while (delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2; ; t = (t + 1) / 2) {
        if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? ~gamma : gamma;
    int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    ranges[i] = make_int2(first, last);
    if (left >= 0) parent_inner[left] = i; else parent_leaf[~left] = i;
    if (right >= 0) parent_inner[right] = i; else parent_leaf[~right] = i;
    if (i == 0) parent_inner[0] = -1;
}

// bottom-up refit: the second thread to arrive at a node computes its box
__global__ void refit_kernel(const Aabb *tri_boxes, const uint32_t *ids, int n, const int2 *children, const int *parent_inner,
                             const int *parent_leaf, Aabb *node_boxes, int *visits)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = parent_leaf[p];
    while (node >= 0) {
        if (atomicAdd(visits + node, 1) == 0) return;   // first arrival: the sibling subtree is not done yet
        __threadfence();
        int2 c = children[node];
        Aabb a = c.x >= 0 ? node_boxes[c.x] : tri_boxes[ids[~c.x]];
        Aabb b = c.y >= 0 ? node_boxes[c.y] : tri_boxes[ids[~c.y]];
        Aabb m;
        m.lo = make_float3(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z));
        m.hi = make_float3(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z));
        node_boxes[node] = m;
        __threadfence();
        node = parent_inner[node];
    }
}

struct Slot { int ref; Aabb box; };

// reference + box of binary child c as seen from a 4-wide parent: a single triangle, a collapsed
// subtree (<= RVB_LEAF_MAX triangles, contiguous in sorted order) or an inner node
__device__ __forceinline__ Slot make_slot(int c, const int2 *ranges, const Aabb *tri_boxes, const uint32_t *ids, const Aabb *node_boxes)
{
    Slot s;
    if (c < 0) {
        s.box = tri_boxes[ids[~c]];
        s.ref = c;                                             // leaf: first = position, count 1
    } else {
        s.box = node_boxes[c];
        const int2 r = ranges[c];
        const int cnt = r.y - r.x + 1;
        s.ref = cnt <= RVB_LEAF_MAX ? ~(int)((uint32_t)r.x | ((uint32_t)(cnt - 1) << 28)) : c;
    }
    return s;
}

// 4-wide node of binary inner node i: its two children, each replaced by ITS two children when it
// is a (non-collapsed) inner node.  Only nodes reachable from the root this way are ever visited.
__global__ void emit_nodes_kernel(int n, const int2 *children, const int2 *ranges, const Aabb *tri_boxes, const uint32_t *ids,
                                  const Aabb *node_boxes, float pad, BvhNode *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    Slot slots[4];
    int ns = 0;
    const int2 c = children[i];
    const int kids[2] = {c.x, c.y};
    for (int k = 0; k < 2; ++k) {
        const Slot s = make_slot(kids[k], ranges, tri_boxes, ids, node_boxes);
        if (s.ref >= 0) {   // inner: expand
            const int2 g = children[s.ref];
            slots[ns++] = make_slot(g.x, ranges, tri_boxes, ids, node_boxes);
            slots[ns++] = make_slot(g.y, ranges, tri_boxes, ids, node_boxes);
        } else slots[ns++] = s;
    }
    float lo[3][4], hi[3][4];
    int ref[4];
    for (int k = 0; k < 4; ++k) {
        if (k < ns) {
            const Aabb &b = slots[k].box;
            lo[0][k] = b.lo.x - pad; lo[1][k] = b.lo.y - pad; lo[2][k] = b.lo.z - pad;
            hi[0][k] = b.hi.x + pad; hi[1][k] = b.hi.y + pad; hi[2][k] = b.hi.z + pad;
            ref[k] = slots[k].ref;
        } else {
            for (int a = 0; a < 3; ++a) { lo[a][k] = 3e38f; hi[a][k] = -3e38f; }
            ref[k] = RVB_BVH_EMPTY;
        }
    }
    BvhNode nd;
    nd.lox = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]); nd.loy = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
    nd.loz = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]); nd.hix = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
    nd.hiy = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]); nd.hiz = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
    nd.ref = make_int4(ref[0], ref[1], ref[2], ref[3]);
    nd.pad = make_float4(0.f, 0.f, 0.f, 0.f);
    out[i] = nd;
}

__global__ void permute_tris_kernel(const float4 *tri, const uint32_t *ids, uint32_t T, float4 *tri_sorted)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T) return;
    uint32_t i = ids[p];
    tri_sorted[3 * p] = tri[3 * i]; tri_sorted[3 * p + 1] = tri[3 * i + 1]; tri_sorted[3 * p + 2] = tri[3 * i + 2];
}

int scene_build_bvh(Scene *sc)
{
    const uint32_t T = sc->T;
    if (T < 2) return fail(RVB_ERR_INVALID, "BVH needs at least two triangles (use the brute-force path)");
    cudaStream_t st = lib_stream();
    cudaEvent_t e0, e1;
    RVB_CUDA(cudaEventCreate(&e0)); RVB_CUDA(cudaEventCreate(&e1));
    RVB_CUDA(cudaEventRecord(e0, st));
    const unsigned blk = 256, grd = (T + blk - 1) / blk;
    Aabb *d_tri_boxes = nullptr, *d_node_boxes = nullptr;
    float4 *d_centroids = nullptr;
    int *d_scene = nullptr, *d_parent_inner = nullptr, *d_parent_leaf = nullptr, *d_visits = nullptr;
    unsigned long long *d_codes[2] = {nullptr, nullptr};
    uint32_t *d_ids[2] = {nullptr, nullptr};
    int2 *d_children = nullptr, *d_ranges = nullptr;
    uint8_t *d_tmp = nullptr;
    RVB_TRY(dev_alloc(&d_tri_boxes, T)); RVB_TRY(dev_alloc(&d_node_boxes, T)); RVB_TRY(dev_alloc(&d_centroids, T));
    RVB_TRY(dev_alloc(&d_scene, 6)); RVB_TRY(dev_alloc(&d_parent_inner, T)); RVB_TRY(dev_alloc(&d_parent_leaf, T));
    RVB_TRY(dev_alloc(&d_visits, T)); RVB_TRY(dev_alloc(&d_children, T)); RVB_TRY(dev_alloc(&d_ranges, T));
    for (int i = 0; i < 2; ++i) { RVB_TRY(dev_alloc(&d_codes[i], T)); RVB_TRY(dev_alloc(&d_ids[i], T)); }
    const int init[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000};   // ordered(+/-FLT_MAX)
    RVB_CUDA(cudaMemcpyAsync(d_scene, init, sizeof init, cudaMemcpyHostToDevice, st));
    note_launch(), tri_bounds_kernel<<<grd, blk, 0, st>>>(sc->d_verts, T, d_tri_boxes, d_centroids, d_scene);
    int h_scene[6];
    RVB_CUDA(cudaMemcpyAsync(h_scene, d_scene, sizeof h_scene, cudaMemcpyDeviceToHost, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    float3 lo = make_float3(unord(h_scene[0]), unord(h_scene[1]), unord(h_scene[2]));
    float3 hi = make_float3(unord(h_scene[3]), unord(h_scene[4]), unord(h_scene[5]));
    float3 ext = make_float3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
    float3 inv = make_float3(ext.x > 0 ? 1.0f / ext.x : 0.f, ext.y > 0 ? 1.0f / ext.y : 0.f, ext.z > 0 ? 1.0f / ext.z : 0.f);
    const float diag = std::sqrt(ext.x * ext.x + ext.y * ext.y + ext.z * ext.z);
    const float pad = 2e-5f * diag + 1e-6f;
    note_launch(), morton_kernel<<<grd, blk, 0, st>>>(d_centroids, T, lo, inv, d_codes[0], d_ids[0]);
    cub::DoubleBuffer<unsigned long long> kb(d_codes[0], d_codes[1]);
    cub::DoubleBuffer<uint32_t> vb(d_ids[0], d_ids[1]);
    size_t bytes = 0;
    RVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, (int)T, 0, 63, st));
    RVB_TRY(dev_alloc(&d_tmp, bytes));
    RVB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, kb, vb, (int)T, 0, 63, st));
    const unsigned long long *codes = kb.Current();
    const uint32_t *ids = vb.Current();
    note_launch(), hierarchy_kernel<<<grd, blk, 0, st>>>(codes, (int)T, d_children, d_ranges, d_parent_inner, d_parent_leaf);
    RVB_CUDA(cudaMemsetAsync(d_visits, 0, T * sizeof(int), st));
    note_launch(), refit_kernel<<<grd, blk, 0, st>>>(d_tri_boxes, ids, (int)T, d_children, d_parent_inner, d_parent_leaf, d_node_boxes, d_visits);
    sc->n_nodes = T - 1;
    RVB_TRY(dev_alloc(&sc->d_nodes, sc->n_nodes));
    RVB_TRY(dev_alloc(&sc->d_order, T));
    RVB_TRY(dev_alloc(&sc->d_tri_sorted, 3 * (size_t)T));
    note_launch(), emit_nodes_kernel<<<grd, blk, 0, st>>>((int)T, d_children, d_ranges, d_tri_boxes, ids, d_node_boxes, pad, sc->d_nodes);
    note_launch(), permute_tris_kernel<<<grd, blk, 0, st>>>(sc->d_tri, ids, T, sc->d_tri_sorted);
    RVB_CUDA(cudaMemcpyAsync(sc->d_order, ids, T * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    RVB_CUDA(cudaGetLastError());
    RVB_CUDA(cudaEventRecord(e1, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&sc->build_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dev_free(d_tri_boxes); dev_free(d_node_boxes); dev_free(d_centroids); dev_free(d_scene);
    dev_free(d_parent_inner); dev_free(d_parent_leaf); dev_free(d_visits); dev_free(d_children); dev_free(d_ranges);
    for (int i = 0; i < 2; ++i) { dev_free(d_codes[i]); dev_free(d_ids[i]); }
    dev_free(d_tmp);
    return RVB_OK;
}

}  // namespace rvb
