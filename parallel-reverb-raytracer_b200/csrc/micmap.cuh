// Microphone visibility map.
//
// Half of all scene queries of the reference's bounce loop are point_intersection(X, mic)
// (kernel.cpp:274-296, called at :335, :431 and :463): "is the segment from X to the microphone
// blocked?".  Every one of those segments ends in the SAME point, so a triangle can only block X if
// the direction mic -> X falls inside the triangle's projection on a sphere around the microphone.
// The map is a cube map (6 faces x N x N cells) around the microphone whose cells list the
// triangles that overlap them (conservatively: the triangle is grown by `delta` in world space and
// by an absolute slack in cell coordinates before it is binned, csrc/micmap.cu).  A visibility
// query then looks up ONE cell and runs the reference's exact Moller-Trumbore test on its few
// triangles -- the same arithmetic on a superset of the triangles that can pass it, so the answer
// equals the brute-force loop's.  Triangles closer to the microphone than `r_near` (where the
// projection degenerates) sit in a short list that every query tests.
#pragma once
#include "rvb_math.cuh"

namespace rvb {

struct MicMap {
    const uint32_t *cell_off = nullptr;   // [6*N*N + 1]
    const uint32_t *cell_ent = nullptr;   // triangle indices (original numbering)
    const uint32_t *near_ent = nullptr;   // triangles within r_near of the microphone
    uint32_t n_near = 0;
    uint32_t N = 0;                       // 0: no map (generic occlusion query)
    float half_n = 0.f;
    float mic[3] = {0.f, 0.f, 0.f};
};

__host__ __device__ __forceinline__ uint32_t micmap_coord(float x, float half_n, uint32_t N)
{
    // monotonic in x: the build and the query use this same expression
    const float f = floorf((x + 1.0f) * half_n);
    return (uint32_t)fminf((float)(N - 1), fmaxf(0.0f, f));
}

// cell of the direction w = X - mic (w != 0)
__device__ __forceinline__ uint32_t micmap_cell(f3 w, float half_n, uint32_t N)
{
    const float ax = fabsf(w.x), ay = fabsf(w.y), az = fabsf(w.z);
    uint32_t a = 0;
    float m = ax, s = w.x, u = w.y, v = w.z;          // (a, a+1, a+2) cyclic
    if (ay > m) { a = 1; m = ay; s = w.y; u = w.z; v = w.x; }
    if (az > m) { a = 2; m = az; s = w.z; u = w.x; v = w.y; }
    const uint32_t face = 2u * a + (s < 0.0f ? 1u : 0u);
    const float inv = 1.0f / m;   // plain IEEE division: |u / m| <= 1
    return (face * N + micmap_coord(__fmul_rn(v, inv), half_n, N)) * N + micmap_coord(__fmul_rn(u, inv), half_n, N);
}

}  // namespace rvb
