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
#include "rvb_sort.cuh"

#include "rvb_internal.cuh"
#include "rvb_scan.cuh"

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
__device__ __forceinline__ Slot make_slot(int c, const int2 *ranges, const Aabb *tri_boxes, const uint32_t *ids, const Aabb *node_boxes,
                                          const uint32_t *rec_of)
{
    Slot s;
    if (c < 0) {
        s.box = tri_boxes[ids[~c]];
        s.ref = ~(int)rec_of[~c];                              // leaf record of a single triangle (count - 1 = 0)
    } else {
        s.box = node_boxes[c];
        const int2 r = ranges[c];
        const int cnt = r.y - r.x + 1;
        s.ref = cnt <= RVB_LEAF_MAX ? ~(int)(rec_of[r.x] | ((uint32_t)(cnt - 1) << 28)) : c;
    }
    return s;
}

// Leaves: a subtree of <= RVB_LEAF_MAX triangles whose parent holds more.  They partition the sorted
// triangle order into consecutive runs; flag[p] = 1 marks the first position of a run, cnt_at[p] its
// length.  (One thread per binary inner node; nodes inside a collapsed subtree mark nothing.)
__global__ void mark_leaves_kernel(int n, const int2 *children, const int2 *ranges, uint32_t *flag, uint32_t *cnt_at)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 r = ranges[i];
    if (i != 0 && r.y - r.x + 1 <= RVB_LEAF_MAX) return;   // inside a collapsed subtree (the root is never collapsed)
    const int2 c = children[i];
    const int kids[2] = {c.x, c.y};
    for (int k = 0; k < 2; ++k) {
        if (kids[k] < 0) { flag[~kids[k]] = 1u; cnt_at[~kids[k]] = 1u; }
        else {
            const int2 q = ranges[kids[k]];
            if (q.y - q.x + 1 <= RVB_LEAF_MAX) { flag[q.x] = 1u; cnt_at[q.x] = (uint32_t)(q.y - q.x + 1); }
        }
    }
}

// Leaf records in the pair-interleaved form of the brute-force loop (scene.cu: pack_quads_kernel):
// 36 floats = two component-wise pairs (v0, e0, e1) of up to four triangles, zero triangles pad
// (det = 0 rejects them, kernel.cpp:72); leaf_ids holds the callers' triangle indices.
__global__ void fill_leaves_kernel(uint32_t T, const uint32_t *flag, const uint32_t *rec_of, const uint32_t *cnt_at, const uint32_t *ids,
                                   const float4 *tri, float *leaf_quads, uint4 *leaf_ids)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T || !flag[p]) return;
    const uint32_t r = rec_of[p], cnt = cnt_at[p];
    const float *src = reinterpret_cast<const float *>(tri);
    uint32_t id[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    for (uint32_t k = 0; k < 4; ++k) {
        float v[9];
        for (int j = 0; j < 9; ++j) v[j] = 0.f;
        if (k < cnt) {
            id[k] = ids[p + k];
            for (int j = 0; j < 9; ++j) v[j] = src[12 * (size_t)id[k] + j];
        }
        for (int j = 0; j < 9; ++j) leaf_quads[4 * RVB_LEAF_STRIDE * (size_t)r + 18 * (k >> 1) + 2 * j + (k & 1)] = v[j];
    }
    reinterpret_cast<uint4 *>(leaf_quads)[RVB_LEAF_STRIDE * (size_t)r + 9] = make_uint4(id[0], id[1], id[2], id[3]);
}

// depth of every triangle in the binary hierarchy (the traversal stack is sized from the maximum)
__global__ void depth_kernel(int n, const int *parent_inner, const int *parent_leaf, int *max_depth)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;
    if (p < n) for (int node = parent_leaf[p]; node >= 0; node = parent_inner[node]) ++d;
    for (int o = 16; o; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}

// 4-wide node of binary inner node i: its two children, each replaced by ITS two children when it
// is a (non-collapsed) inner node.  Only nodes reachable from the root this way are ever visited.
__global__ void emit_nodes_kernel(int n, const int2 *children, const int2 *ranges, const Aabb *tri_boxes, const uint32_t *ids,
                                  const Aabb *node_boxes, const uint32_t *rec_of, float pad, BvhNode *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    Slot slots[4];
    int ns = 0;
    const int2 c = children[i];
    const int kids[2] = {c.x, c.y};
    for (int k = 0; k < 2; ++k) {
        const Slot s = make_slot(kids[k], ranges, tri_boxes, ids, node_boxes, rec_of);
        if (s.ref >= 0) {   // inner: expand
            const int2 g = children[s.ref];
            slots[ns++] = make_slot(g.x, ranges, tri_boxes, ids, node_boxes, rec_of);
            slots[ns++] = make_slot(g.y, ranges, tri_boxes, ids, node_boxes, rec_of);
        } else slots[ns++] = s;
    }
    // the node's own box = union of its (padded) child boxes: the quantisation grid
    float blo[3] = {3e38f, 3e38f, 3e38f}, bhi[3] = {-3e38f, -3e38f, -3e38f};
    float clo[4][3], chi[4][3];
    for (int k = 0; k < ns; ++k) {
        const Aabb &b = slots[k].box;
        clo[k][0] = b.lo.x - pad; clo[k][1] = b.lo.y - pad; clo[k][2] = b.lo.z - pad;
        chi[k][0] = b.hi.x + pad; chi[k][1] = b.hi.y + pad; chi[k][2] = b.hi.z + pad;
        for (int a = 0; a < 3; ++a) { blo[a] = fminf(blo[a], clo[k][a]); bhi[a] = fmaxf(bhi[a], chi[k][a]); }
    }
    BvhNode nd;
    nd.exps = 0;
    nd.pad[0] = nd.pad[1] = 0;
    for (int a = 0; a < 3; ++a) {
        nd.origin[a] = blo[a];
        // cell size: the smallest power of two with 255 cells covering the extent (with a little room
        // for the rounding of the decode); biased exponent kept in [1, 230] so that 2^15 * cell is a float
        int e;
        frexpf(fmaxf((bhi[a] - blo[a]) * 1.0001f / 255.0f, 1e-30f), &e);   // v < 2^e
        e = min(max(e + 127, 1), 230);
        nd.exps |= (uint32_t)e << (8 * a);
        const float cell = __uint_as_float((uint32_t)e << 23);
        uint32_t wlo = 0, whi = 0;
        for (int k = 0; k < 4; ++k) {
            int ql = 255, qh = 0;   // unused slot: inverted
            if (k < ns) {
                ql = (int)fminf(fmaxf(floorf((clo[k][a] - blo[a]) / cell), 0.0f), 255.0f);
                qh = (int)fminf(fmaxf(ceilf((chi[k][a] - blo[a]) / cell), 0.0f), 255.0f);
                // outward under the decode's own arithmetic
                while (ql > 0 && fmaf((float)ql, cell, blo[a]) > clo[k][a]) --ql;
                while (qh < 255 && fmaf((float)qh, cell, blo[a]) < chi[k][a]) ++qh;
            }
            wlo |= (uint32_t)ql << (8 * k);
            whi |= (uint32_t)qh << (8 * k);
        }
        nd.qlo[a] = wlo; nd.qhi[a] = whi;
    }
    for (int k = 0; k < 4; ++k) nd.ref[k] = k < ns ? slots[k].ref : RVB_BVH_EMPTY;
    out[i] = nd;
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
    note_launch(), morton_kernel<<<grd, blk, 0, st>>>(d_centroids, T, lo, inv, d_codes[1], d_ids[1]);
    int sorted = 1;
    const int rc_sort = rsort::sort_arrays<unsigned long long>(d_codes, d_ids, T, 0, 63, &sorted, st);   // input in buffer pair 1
    RVB_TRY(rc_sort);
    const unsigned long long *codes = d_codes[sorted];
    const uint32_t *ids = d_ids[sorted];
    note_launch(), hierarchy_kernel<<<grd, blk, 0, st>>>(codes, (int)T, d_children, d_ranges, d_parent_inner, d_parent_leaf);
    RVB_CUDA(cudaMemsetAsync(d_visits, 0, T * sizeof(int), st));
    note_launch(), refit_kernel<<<grd, blk, 0, st>>>(d_tri_boxes, ids, (int)T, d_children, d_parent_inner, d_parent_leaf, d_node_boxes, d_visits);
    sc->n_nodes = T - 1;
    RVB_TRY(dev_alloc(&sc->d_nodes, sc->n_nodes));
    // leaves: mark the runs, number them (exclusive scan of the start flags), pack their records
    uint32_t *d_flag = nullptr, *d_rec_of = nullptr, *d_cnt_at = nullptr, *d_scan_tmp = nullptr;
    int *d_depth = nullptr;
    RVB_TRY(dev_alloc(&d_flag, (size_t)T + 1)); RVB_TRY(dev_alloc(&d_rec_of, (size_t)T + 1)); RVB_TRY(dev_alloc(&d_cnt_at, T));
    RVB_TRY(dev_alloc(&d_scan_tmp, scan_blocks(T))); RVB_TRY(dev_alloc(&d_depth, 1));
    RVB_CUDA(cudaMemsetAsync(d_flag, 0, ((size_t)T + 1) * sizeof(uint32_t), st));
    RVB_CUDA(cudaMemsetAsync(d_depth, 0, sizeof(int), st));
    note_launch(), mark_leaves_kernel<<<grd, blk, 0, st>>>((int)T, d_children, d_ranges, d_flag, d_cnt_at);
    RVB_TRY(exclusive_scan_u32(d_flag, d_rec_of, T, d_scan_tmp, d_rec_of + T, st));
    note_launch(), depth_kernel<<<grd, blk, 0, st>>>((int)T, d_parent_inner, d_parent_leaf, d_depth);
    uint32_t n_leaves = 0;
    int depth = 0;
    RVB_TRY(read_back(&n_leaves, d_rec_of + T, sizeof n_leaves));
    RVB_TRY(read_back(&depth, d_depth, sizeof depth));
    // a 4-wide visit descends two binary levels (one when it ends in a leaf) and pushes at most three
    // entries; + the sentinel and one spare
    sc->bvh_stack = 3 * ((depth + 1) / 2) + 2;
    if (sc->bvh_stack > RVB_BVH_STACK)
        return fail(RVB_ERR_UNSUPPORTED, "BVH of depth " + std::to_string(depth) + " exceeds the traversal stack (near-coincident triangles?)");
    sc->n_leaves = n_leaves;
    sc->bvh_depth = depth;
    RVB_TRY(dev_alloc(&sc->d_leaf_quads, RVB_LEAF_STRIDE * (size_t)n_leaves));
    note_launch(), fill_leaves_kernel<<<grd, blk, 0, st>>>(T, d_flag, d_rec_of, d_cnt_at, ids, sc->d_tri, reinterpret_cast<float *>(sc->d_leaf_quads), sc->d_leaf_ids);
    note_launch(), emit_nodes_kernel<<<grd, blk, 0, st>>>((int)T, d_children, d_ranges, d_tri_boxes, ids, d_node_boxes, d_rec_of, pad, sc->d_nodes);
    RVB_CUDA(cudaGetLastError());
    RVB_CUDA(cudaEventRecord(e1, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_flag); dev_free(d_rec_of); dev_free(d_cnt_at); dev_free(d_scan_tmp); dev_free(d_depth);
    sc->diag = diag;
    cudaEventElapsedTime(&sc->build_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dev_free(d_tri_boxes); dev_free(d_node_boxes); dev_free(d_centroids); dev_free(d_scene);
    dev_free(d_parent_inner); dev_free(d_parent_leaf); dev_free(d_visits); dev_free(d_children); dev_free(d_ranges);
    for (int i = 0; i < 2; ++i) { dev_free(d_codes[i]); dev_free(d_ids[i]); }
    dev_free(d_tmp);
    return RVB_OK;
}

}  // namespace rvb
