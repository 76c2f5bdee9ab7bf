// Which packed form of the Moller-Trumbore loop is faster on sm_100a?
//   A  two TRIANGLES per instruction, ray scalars broadcast (the shipped loop; scene in shared memory)
//   B  two RAYS per instruction, triangle scalars broadcast from registers (scene in shared memory)
//   C  two RAYS per instruction, triangle scalars from UNIFORM registers (scene in constant memory)
// Each thread runs `iters` nearest-hit queries over T triangles; prints cycles per warp-level
// ray-triangle test per SM sub-partition and the FP32 fraction (52 FLOP per test).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pair_bench.bin pair_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../parallel-reverb-raytracer_b200/csrc/rvb_math.cuh"
using namespace rvb;

#define MAXT 1024
__constant__ float4 c_tri[3 * MAXT];   // (v0.xyz,e0.x) (e0.yz,e1.xy) (e1.z,0,0,0)

// B / C: the pair is (ray a, ray b); tri components are scalars
__device__ __forceinline__ float2 mt_ray_pair(f3 v0, f3 e0, f3 e1, const f3x2 &pos, const f3x2 &dir, bool &ok_a, bool &ok_b)
{
    f3x2 E0, E1, V0;
    E0.x = bcast2(e0.x); E0.y = bcast2(e0.y); E0.z = bcast2(e0.z);
    E1.x = bcast2(e1.x); E1.y = bcast2(e1.y); E1.z = bcast2(e1.z);
    V0.x = bcast2(v0.x); V0.y = bcast2(v0.y); V0.z = bcast2(v0.z);
    const f3x2 pvec = cross2(dir, E1);
    const float2 det = dot2(E0, pvec);
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(det.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(det.y));
    const float2 e = __ffma2_rn(det, r, bcast2(-1.0f));
    const float2 invdet = __ffma2_rn(r, neg2(e), r);
    f3x2 tvec;
    tvec.x = __fadd2_rn(pos.x, neg2(V0.x)); tvec.y = __fadd2_rn(pos.y, neg2(V0.y)); tvec.z = __fadd2_rn(pos.z, neg2(V0.z));
    const float2 u = __fmul2_rn(invdet, dot2(tvec, pvec));
    const f3x2 qvec = cross2(tvec, E0);
    const float2 v = __fmul2_rn(invdet, dot2(dir, qvec));
    const float2 t = __fmul2_rn(invdet, dot2(E1, qvec));
    ok_a = !((-RVB_EPSILON < det.x && det.x < RVB_EPSILON) | (u.x < 0.0f) | (v.x < 0.0f) | (1.0f < __fadd_rn(v.x, u.x)));
    ok_b = !((-RVB_EPSILON < det.y && det.y < RVB_EPSILON) | (u.y < 0.0f) | (v.y < 0.0f) | (1.0f < __fadd_rn(v.y, u.y)));
    return t;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const float4 *quads, const float4 *tris, int T, int iters, const float4 *rays, float *out)
{
    extern __shared__ float4 smem[];
    const int nquads = (T + 3) / 4;
    if (MODE == 0) for (int i = threadIdx.x; i < 9 * nquads; i += blockDim.x) smem[i] = quads[i];
    if (MODE == 1) for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) smem[i] = tris[i];
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    if (MODE == 0) {
        f3 pos = mk3(rays[2 * tid]), dir = mk3(rays[2 * tid + 1]);
        for (int it = 0; it < iters; ++it) {
            float tbest = __int_as_float(0x7f800000); unsigned ibest = 0xffffffffu;
#pragma unroll 2
            for (int q = 0; q < nquads; ++q) {
                const float4 *r = smem + 9 * q;
                const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5], r6 = r[6], r7 = r[7], r8 = r[8];
                f3x2 v0a, e0a, e1a, v0b, e0b, e1b;
                v0a.x = make_float2(r0.x, r0.y); v0a.y = make_float2(r0.z, r0.w); v0a.z = make_float2(r1.x, r1.y);
                e0a.x = make_float2(r1.z, r1.w); e0a.y = make_float2(r2.x, r2.y); e0a.z = make_float2(r2.z, r2.w);
                e1a.x = make_float2(r3.x, r3.y); e1a.y = make_float2(r3.z, r3.w); e1a.z = make_float2(r4.x, r4.y);
                v0b.x = make_float2(r4.z, r4.w); v0b.y = make_float2(r5.x, r5.y); v0b.z = make_float2(r5.z, r5.w);
                e0b.x = make_float2(r6.x, r6.y); e0b.y = make_float2(r6.z, r6.w); e0b.z = make_float2(r7.x, r7.y);
                e1b.x = make_float2(r7.z, r7.w); e1b.y = make_float2(r8.x, r8.y); e1b.z = make_float2(r8.z, r8.w);
                bool o0, o1, o2, o3;
                const float2 ta = mt_intersect_pair(v0a, e0a, e1a, pos, dir, o0, o1);
                const float2 tb = mt_intersect_pair(v0b, e0b, e1b, pos, dir, o2, o3);
                if (o0 && ta.x > RVB_EPSILON && ta.x < tbest) { tbest = ta.x; ibest = 4 * q; }
                if (o1 && ta.y > RVB_EPSILON && ta.y < tbest) { tbest = ta.y; ibest = 4 * q + 1; }
                if (o2 && tb.x > RVB_EPSILON && tb.x < tbest) { tbest = tb.x; ibest = 4 * q + 2; }
                if (o3 && tb.y > RVB_EPSILON && tb.y < tbest) { tbest = tb.y; ibest = 4 * q + 3; }
            }
            acc += tbest + (float)ibest;
            pos.x = __fadd_rn(pos.x, 1e-3f * (float)(ibest & 1));   // keep the loop alive, stay inside the box
            dir.y = -dir.y;
        }
    } else {
        f3x2 pos, dir;
        const float4 pa = rays[4 * tid], da = rays[4 * tid + 1], pb = rays[4 * tid + 2], db = rays[4 * tid + 3];
        pos.x = make_float2(pa.x, pb.x); pos.y = make_float2(pa.y, pb.y); pos.z = make_float2(pa.z, pb.z);
        dir.x = make_float2(da.x, db.x); dir.y = make_float2(da.y, db.y); dir.z = make_float2(da.z, db.z);
        for (int it = 0; it < iters; ++it) {
            float2 tbest = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
            unsigned ia = 0xffffffffu, ib = 0xffffffffu;
#pragma unroll 4
            for (int i = 0; i < T; ++i) {
                const float4 A = MODE == 1 ? smem[3 * i] : c_tri[3 * i], B = MODE == 1 ? smem[3 * i + 1] : c_tri[3 * i + 1],
                             C = MODE == 1 ? smem[3 * i + 2] : c_tri[3 * i + 2];
                bool oa, ob;
                const float2 t = mt_ray_pair(mk3(A.x, A.y, A.z), mk3(A.w, B.x, B.y), mk3(B.z, B.w, C.x), pos, dir, oa, ob);
                if (oa && t.x > RVB_EPSILON && t.x < tbest.x) { tbest.x = t.x; ia = i; }
                if (ob && t.y > RVB_EPSILON && t.y < tbest.y) { tbest.y = t.y; ib = i; }
            }
            acc += tbest.x + tbest.y + (float)(ia + ib);
            pos.x = __fadd2_rn(pos.x, make_float2(1e-3f * (float)(ia & 1), 1e-3f * (float)(ib & 1)));
            dir.y = neg2(dir.y);
        }
    }
    out[tid] = acc;
}

int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 12;
    // shoebox-like scene: T triangles tiling the faces of a box [0,10]x[0,4]x[0,6] (repeated)
    std::vector<float4> tris(3 * T), quads(9 * ((T + 3) / 4), make_float4(0, 0, 0, 0));
    const float box[12][9] = {
        {0,0,0, 10,0,0, 0,0,6}, {10,0,6, -10,0,0, 0,0,-6}, {0,4,0, 0,0,6, 10,0,0}, {10,4,6, 0,0,-6, -10,0,0},
        {0,0,0, 0,4,0, 10,0,0}, {10,4,0, 0,-4,0, -10,0,0}, {0,0,6, 10,0,0, 0,4,0}, {10,4,6, -10,0,0, 0,-4,0},
        {0,0,0, 0,0,6, 0,4,0}, {0,4,6, 0,0,-6, 0,-4,0}, {10,0,0, 0,4,0, 0,0,6}, {10,4,6, 0,-4,0, 0,0,-6}};
    for (int i = 0; i < T; ++i) {
        const float *b = box[i % 12];
        tris[3 * i] = make_float4(b[0], b[1], b[2], b[3]); tris[3 * i + 1] = make_float4(b[4], b[5], b[6], b[7]); tris[3 * i + 2] = make_float4(b[8], 0, 0, 0);
        float *qd = reinterpret_cast<float *>(quads.data()) + 18 * (i / 2) + (i & 1);
        for (int kx = 0; kx < 9; ++kx) qd[2 * kx] = b[kx];
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = T <= 64 ? 4096 : 256;
    std::vector<float4> rays(4 * blocks * threads);
    for (size_t i = 0; i < rays.size() / 2; ++i) {
        const float a = 0.37f * (float)(i % 1013), b = 0.11f * (float)(i % 797);
        rays[2 * i] = make_float4(2.f + 0.003f * (float)(i % 911), 1.5f, 2.f, 0.f);
        rays[2 * i + 1] = make_float4(cosf(a) * cosf(b), sinf(b), sinf(a) * cosf(b), 0.f);
    }
    float4 *d_q, *d_t, *d_r; float *d_o;
    cudaMalloc(&d_q, quads.size() * 16); cudaMalloc(&d_t, tris.size() * 16); cudaMalloc(&d_r, rays.size() * 16); cudaMalloc(&d_o, 4 * blocks * threads);
    cudaMemcpy(d_q, quads.data(), quads.size() * 16, cudaMemcpyHostToDevice); cudaMemcpy(d_t, tris.data(), tris.size() * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(d_r, rays.data(), rays.size() * 16, cudaMemcpyHostToDevice); cudaMemcpyToSymbol(c_tri, tris.data(), tris.size() * 16);
    const size_t smem = 16 * (size_t)(9 * ((T + 3) / 4) > 3 * T ? 9 * ((T + 3) / 4) : 3 * T);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[3] = {"A two triangles / instr (smem)", "B two rays / instr (smem scalars)", "C two rays / instr (uniform regs)"};
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<blocks, threads, smem>>>(d_q, d_t, T, iters, d_r, d_o); }
            if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<blocks, threads, smem>>>(d_q, d_t, T, iters, d_r, d_o); }
            if (mode == 2) k<2><<<blocks, threads, 0>>>(d_q, d_t, T, iters, d_r, d_o);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double rays_per_thread = mode == 0 ? 1 : 2;
        const double tests = (double)blocks * threads * rays_per_thread * iters * T;
        const double cyc = best * 1e-3 * p.clockRate * 1e3 * p.multiProcessorCount * 4;
        printf("T=%d %-36s %8.3f ms  %6.2f cycles per warp-test per SMSP  %5.1f %% of FP32 peak (%s)\n", T, names[mode], best,
               cyc / (tests / 32), 100.0 * tests * 52 / (best * 1e-3) / (p.multiProcessorCount * 128 * 2.0 * p.clockRate * 1e3), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
