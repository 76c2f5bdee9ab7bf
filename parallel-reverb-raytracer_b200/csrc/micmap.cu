// Build of the microphone visibility map (see micmap.cuh for what it is and why it is exact).
//
// One warp per triangle, two passes (count, fill) around a device-wide exclusive scan.  For each
// cube face the triangle is clipped against w >= wc (w = depth along the face axis), its vertices
// are grown by `delta` in world space and projected to (u, v) = (b / w, c / w); the triangle is
// registered in every cell of the bounding rectangle of the result -- plus an absolute slack for the
// rounding of the query's own cell computation -- that a separating-edge test against the projected
// polygon (grown by the same margins) cannot exclude.  Entries inside a cell are in no particular
// order: a visibility query is an OR over its cell's triangles.
//
// tools/bvh_lab/micmap_lab.cpp runs this very construction on the CPU against the brute-force loop
// (canonical float arithmetic) for traced and adversarial queries, including a microphone a few
// centimetres above a wall: no accepted hit was ever missing from its cell.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "micmap.cuh"
#include "rvb_internal.cuh"
#include "rvb_scan.cuh"

namespace rvb {

struct MicMapParams {
    const float4 *verts;   // [3T] v0, v1, v2
    uint32_t T, N;
    float half_n, delta, r_near, wc;
    float mic[3];
    uint32_t near_cap;
};

// cell rectangle of a triangle on one cube face + the projected (clipped) polygon and the margin it was
// grown by, for the per-cell test
struct FaceRect {
    uint32_t i0, i1, j0, j1;
    int np; float pu[4], pv[4], margin;
    // per-edge constants of the separating-edge test, worked out once per rectangle (prepare_edges)
    int ne;                       // -1: reject nothing (degenerate polygon)
    float nu[4], nv[4], thr[4];   // outward normal of edge k -> k+1 and the distance beyond which a cell is outside
};

__device__ __forceinline__ void grow_uv(float W, float U, float V, float delta, float &ulo, float &uhi, float &vlo, float &vhi, FaceRect &r)
{
    const float wn = W - delta, wf = W + delta;   // wn > 0: W >= wc = 20 delta
    const float u0 = U - delta, u1 = U + delta, v0 = V - delta, v1 = V + delta;
    const float a0 = fminf(u0 / wn, u0 / wf), a1 = fmaxf(u1 / wn, u1 / wf), b0 = fminf(v0 / wn, v0 / wf), b1 = fmaxf(v1 / wn, v1 / wf);
    ulo = fminf(ulo, a0); uhi = fmaxf(uhi, a1); vlo = fminf(vlo, b0); vhi = fmaxf(vhi, b1);
    // the vertex itself and how far its grown image can project from it
    const float pu = U / W, pv = V / W;
    if (r.np < 4) { r.pu[r.np] = pu; r.pv[r.np] = pv; ++r.np; }
    const float mu = fmaxf(a1 - pu, pu - a0), mv = fmaxf(b1 - pv, pv - b0);
    r.margin = fmaxf(r.margin, sqrtf(mu * mu + mv * mv));
}

// Is cell (i, j) certainly outside the projected polygon grown by its margin?  (Separating-edge test
// against the cell corner that lies deepest inside; a degenerate, edge-on polygon rejects nothing.)
// prepare_edges: the part that does not depend on the cell.
__device__ __forceinline__ void prepare_edges(FaceRect &r)
{
    r.ne = -1;
    if (r.np < 3) return;
    float area = 0.0f;
    for (int k = 0; k < r.np; ++k) { const int k1 = (k + 1) % r.np; area += r.pu[k] * r.pv[k1] - r.pu[k1] * r.pv[k]; }
    if (fabsf(area) < 1e-12f) return;
    const float sgn = area > 0.0f ? 1.0f : -1.0f;
    r.ne = r.np;
    for (int k = 0; k < r.np; ++k) {
        const int k1 = (k + 1) % r.np;
        const float du = r.pu[k1] - r.pu[k], dv = r.pv[k1] - r.pv[k];
        const float nu = sgn * dv, nv = -sgn * du;               // outward normal of edge k -> k1
        const float len = sqrtf(nu * nu + nv * nv);
        r.nu[k] = nu; r.nv[k] = nv;
        // (the cell bounds are recomputed from the index, hence the 2e-5 on top of the map's own 1e-5 slack);
        // a zero-length edge excludes nothing
        r.thr[k] = len == 0.0f ? 3e38f : (r.margin + 3e-5f) * len * 1.0001f + 1e-7f;
    }
}
__device__ __forceinline__ bool cell_outside(const FaceRect &r, uint32_t i, uint32_t j, uint32_t N)
{
    if (r.ne < 0) return false;
    const float h = 2.0f / (float)N;
    const float cu0 = -1.0f + (float)i * h, cu1 = cu0 + h, cv0 = -1.0f + (float)j * h, cv1 = cv0 + h;
    for (int k = 0; k < r.ne; ++k) {
        const float nu = r.nu[k], nv = r.nv[k];
        const float cu = nu > 0.0f ? cu0 : cu1, cv = nv > 0.0f ? cv0 : cv1;
        if (nu * (cu - r.pu[k]) + nv * (cv - r.pv[k]) > r.thr[k]) return true;
    }
    return false;
}

// p: the triangle's vertices relative to the microphone
__device__ __forceinline__ bool face_rect(const float (&p)[3][3], int face, const MicMapParams &M, FaceRect &r)
{
    const int a = face >> 1, b = (a + 1) % 3, c = (a + 2) % 3;
    const float s = (face & 1) ? -1.0f : 1.0f;
    float w[3], ub[3], vc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { w[k] = s * p[k][a]; ub[k] = p[k][b]; vc[k] = p[k][c]; }
    if (w[0] < M.wc && w[1] < M.wc && w[2] < M.wc) return false;
    float ulo = 3e38f, uhi = -3e38f, vlo = 3e38f, vhi = -3e38f;
    r.np = 0; r.margin = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3;
        if (w[k] >= M.wc) grow_uv(w[k], ub[k], vc[k], M.delta, ulo, uhi, vlo, vhi, r);
        if ((w[k] < M.wc) != (w[k1] < M.wc)) {   // the edge crosses the clip plane
            const float t = (M.wc - w[k]) / (w[k1] - w[k]);
            grow_uv(M.wc, ub[k] + t * (ub[k1] - ub[k]), vc[k] + t * (vc[k1] - vc[k]), M.delta, ulo, uhi, vlo, vhi, r);
        }
    }
    const float slack = 1e-5f;   // rounding of the query's (u, v)
    ulo -= slack; uhi += slack; vlo -= slack; vhi += slack;
    if (uhi < -1.0f || ulo > 1.0f || vhi < -1.0f || vlo > 1.0f) return false;
    r.i0 = micmap_coord(fmaxf(ulo, -1.0f), M.half_n, M.N); r.i1 = micmap_coord(fminf(uhi, 1.0f), M.half_n, M.N);
    r.j0 = micmap_coord(fmaxf(vlo, -1.0f), M.half_n, M.N); r.j1 = micmap_coord(fminf(vhi, 1.0f), M.half_n, M.N);
    return true;
}

// FILL = false: count entries per cell (counts[cell]) and near triangles (near_count)
// FILL = true : write them (cursor[cell] starts at 0; offsets from the scan)
// One WARP per triangle: lanes 0..5 work out the six faces' rectangles (one face each), then all lanes
// share the cells of every rectangle -- the fill's atomics return a value, and a single thread walking
// its ~30 cells one round trip after the other took 2 ms on the 100k-triangle hall.
template <bool FILL>
__global__ void __launch_bounds__(128) micmap_bin_kernel(const MicMapParams M, uint32_t *counts_or_cursor, const uint32_t *cell_off,
                                                         uint32_t *cell_ent, uint32_t *near_count, uint32_t *near_ent)
{
    __shared__ FaceRect s_rect[4][6];
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (i >= M.T) return;   // warp-uniform
    float p[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 v = M.verts[3 * i + k];
        p[k][0] = v.x - M.mic[0]; p[k][1] = v.y - M.mic[1]; p[k][2] = v.z - M.mic[2];
    }
    // distance from the microphone to the triangle's bounding box: a lower bound of the distance to the triangle
    float d2 = 0.0f;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float mn = fminf(p[0][ax], fminf(p[1][ax], p[2][ax])), mx = fmaxf(p[0][ax], fmaxf(p[1][ax], p[2][ax]));
        const float d = mn > 0.0f ? mn : (mx < 0.0f ? -mx : 0.0f);
        d2 += d * d;
    }
    if (!(d2 >= M.r_near * M.r_near)) {   // near (or NaN): tested by every query
        if (lane == 0) {
            const uint32_t k = atomicAdd(near_count, 1u);
            if (FILL && k < M.near_cap) near_ent[k] = i;
        }
        return;
    }
    bool ok = false;
    if (lane < 6) {
        FaceRect r;
        ok = face_rect(p, (int)lane, M, r);
        if (ok) { prepare_edges(r); s_rect[warp][lane] = r; }
    }
    const uint32_t valid = __ballot_sync(0xffffffffu, ok);
    __syncwarp();
    for (uint32_t f = 0; f < 6; ++f) {
        if (!((valid >> f) & 1u)) continue;
        const FaceRect &r = s_rect[warp][f];
        const uint32_t w = r.i1 - r.i0 + 1, n = w * (r.j1 - r.j0 + 1);
        for (uint32_t c = lane; c < n; c += 32u) {
            const uint32_t j = r.j0 + c / w, ii = r.i0 + c % w;
            if (cell_outside(r, ii, j, M.N)) continue;
            const uint32_t cell = (f * M.N + j) * M.N + ii;
            const uint32_t k = atomicAdd(counts_or_cursor + cell, 1u);
            if (FILL) cell_ent[cell_off[cell] + k] = i;
        }
    }
}

void micmap_free(MicMapStorage *s)
{
    dev_free(s->d_cell_off); dev_free(s->d_cell_ent); dev_free(s->d_near);
    s->map = MicMap{};
}

// Batched build: pass 1 (count + scan) of every job is enqueued, ONE synchronisation brings the entry
// totals to the host, then the lists are allocated and filled.  `nqueries` (rays x reflections per job)
// sizes the grid: building costs ~ entries, a visibility query ~ candidates per cell, so a job with few
// queries gets a coarser map (16 cells per triangle from 32 queries per triangle upwards, 4 at the
// other end; C5's 16 k rays x 128 per pair: 10).
int micmap_build_batch(const Scene *sc, const float *mics /* [n][3] */, size_t n, size_t nqueries, MicMapStorage *out, double forced_cells_per_tri)
{
    for (size_t j = 0; j < n; ++j) out[j] = MicMapStorage{};
    static const int enabled = getenv("RVB_MICMAP") ? atoi(getenv("RVB_MICMAP")) : 1;
    if (!enabled || sc->accel != RVB_ACCEL_BVH || !(sc->diag > 0.0f) || n == 0) return RVB_OK;
    static const double cpt_env = getenv("RVB_MICMAP_CPT") ? atof(getenv("RVB_MICMAP_CPT")) : 0.0;
    // finer than 16 cells per triangle only pays above ~40 M queries, and little since the wavefront kernel shares
    // the candidate tests of a serve round among its lanes (C3, profiles/r2_micmap_resolution.log: 1 Mi rays x 128,
    // trace 50.75 / 50.63 / 50.55 / 50.72 ms at 16 / 24 / 32 / 48 cells; a 131 k-ray shard: 7.46 / 7.41 / 7.43 at 10 / 16 / 24)
    const double cap = std::min(32.0, std::max(16.0, 16.0 * (double)nqueries / 40e6));
    const double cells_per_tri = forced_cells_per_tri > 0.0 ? forced_cells_per_tri : cpt_env > 0.0 ? cpt_env : std::min(cap, std::max(4.0, (double)nqueries / (2.0 * sc->T)));
    cudaStream_t st = lib_stream();
    MicMapParams M{};
    M.verts = sc->d_verts; M.T = sc->T;
    static const double nmax = getenv("RVB_MICMAP_NMAX") ? atof(getenv("RVB_MICMAP_NMAX")) : 1024.0;
    M.N = (uint32_t)std::min(nmax, std::max(8.0, std::round(std::sqrt(cells_per_tri * sc->T / 6.0))));
    M.half_n = 0.5f * (float)M.N;
    M.delta = 1e-4f * sc->diag + 1e-6f;
    M.r_near = 40.0f * M.delta;
    M.wc = 0.5f * M.r_near;
    M.near_cap = 4096;
    const size_t ncell = (size_t)6 * M.N * M.N;
    const unsigned grd = (M.T + 3) / 4;   // one warp per triangle, four per CTA
    std::vector<uint32_t *> cursor(n, nullptr), block(n, nullptr);
    uint32_t *d_counts = nullptr;          // per job: near triangles, total entries
    int rc = RVB_OK;
    auto done = [&](int code) {
        for (size_t j = 0; j < n; ++j) { dev_free(cursor[j]); dev_free(block[j]); if (code) micmap_free(&out[j]); }
        dev_free(d_counts);
        return code;
    };
    if ((rc = dev_alloc(&d_counts, 2 * n))) return done(rc);
    RVB_CUDA(cudaMemsetAsync(d_counts, 0, 2 * n * sizeof(uint32_t), st));
    for (size_t j = 0; j < n; ++j) {
        MicMapStorage &o = out[j];
        if ((rc = dev_alloc(&o.d_cell_off, ncell + 1)) || (rc = dev_alloc(&cursor[j], ncell)) || (rc = dev_alloc(&block[j], scan_blocks(ncell))) ||
            (rc = dev_alloc(&o.d_near, (size_t)M.near_cap)))
            return done(rc);
        RVB_CUDA(cudaMemsetAsync(o.d_cell_off, 0, (ncell + 1) * sizeof(uint32_t), st));
        RVB_CUDA(cudaMemsetAsync(cursor[j], 0, ncell * sizeof(uint32_t), st));
        for (int k = 0; k < 3; ++k) M.mic[k] = mics[3 * j + k];
        note_launch(), micmap_bin_kernel<false><<<grd, 128, 0, st>>>(M, o.d_cell_off, nullptr, nullptr, d_counts + 2 * j, nullptr);
        if ((rc = exclusive_scan_u32(o.d_cell_off, o.d_cell_off, ncell, block[j], o.d_cell_off + ncell, st))) return done(rc);
        RVB_CUDA(cudaMemcpyAsync(d_counts + 2 * j + 1, o.d_cell_off + ncell, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    std::vector<uint32_t> h(2 * n, 0u);
    if ((rc = read_back(h.data(), d_counts, 2 * n * sizeof(uint32_t)))) return done(rc);
    RVB_CUDA(cudaMemsetAsync(d_counts, 0, 2 * n * sizeof(uint32_t), st));
    for (size_t j = 0; j < n; ++j) {
        MicMapStorage &o = out[j];
        const uint32_t n_near = h[2 * j], total = h[2 * j + 1];
        if (n_near > M.near_cap) { micmap_free(&o); continue; }   // microphone inside dense geometry: generic queries
        if ((rc = dev_alloc(&o.d_cell_ent, (size_t)std::max(1u, total)))) return done(rc);
        for (int k = 0; k < 3; ++k) M.mic[k] = mics[3 * j + k];
        note_launch(), micmap_bin_kernel<true><<<grd, 128, 0, st>>>(M, cursor[j], o.d_cell_off, o.d_cell_ent, d_counts + 2 * j, o.d_near);
        o.map.cell_off = o.d_cell_off; o.map.cell_ent = o.d_cell_ent; o.map.near_ent = o.d_near;
        o.map.n_near = n_near; o.map.N = M.N; o.map.half_n = M.half_n;
        for (int k = 0; k < 3; ++k) o.map.mic[k] = mics[3 * j + k];
        o.entries = total;
    }
    RVB_CUDA(cudaGetLastError());
    return done(RVB_OK);
}

int micmap_build(const Scene *sc, const float mic[3], size_t nqueries, MicMapStorage *out)
{
    return micmap_build_batch(sc, mic, 1, nqueries, out, 0.0);
}

}  // namespace rvb
