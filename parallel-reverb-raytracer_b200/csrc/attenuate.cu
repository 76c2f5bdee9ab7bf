// Attenuators on host arrays: drop-in for SpeakerAttenuator::attenuate (rayverb.cpp:856-892) and
// HrtfAttenuator::attenuate (rayverb.cpp:765-818), one channel per call like the reference.
// The fused device-resident path (render.cu) is what the CLI and the benchmark use; these exist
// so the reference's own attenuation tests port 1:1.
#include "attenuate.cuh"
#include "rvb_internal.cuh"

namespace rvb {

__device__ __forceinline__ bool load_impulse(const rvb_impulse *in, size_t i, float4 &va, float4 &vb, f3 &pos, float &time)
{
    const float4 *s = reinterpret_cast<const float4 *>(in + i);
    va = s[0]; vb = s[1];
    float4 p = s[2];
    pos = mk3(p);
    time = s[3].x;
    return va.x != 0.f || va.y != 0.f || va.z != 0.f || va.w != 0.f || vb.x != 0.f || vb.y != 0.f || vb.z != 0.f || vb.w != 0.f;
}

__device__ __forceinline__ void store_attenuated(rvb_attenuated_impulse *out, size_t i, float4 va, float4 vb, float time)
{
    float4 *o = reinterpret_cast<float4 *>(out + i);
    o[0] = va; o[1] = vb; o[2] = make_float4(time, 0.f, 0.f, 0.f); o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// kernel attenuate (kernel.cpp:515-535).  Skipped (all-zero) impulses leave zeros.
__global__ void attenuate_speaker_kernel(const rvb_impulse *in, size_t n, f3 mic, f3 sdir, float coef, rvb_attenuated_impulse *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 va, vb; f3 pos; float time;
    if (!load_impulse(in, i, va, vb, pos, time)) { store_attenuated(out, i, make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), 0.f); return; }
    float g = speaker_gain(mic, pos, sdir, coef);
    va = make_float4(__fmul_rn(va.x, g), __fmul_rn(va.y, g), __fmul_rn(va.z, g), __fmul_rn(va.w, g));
    vb = make_float4(__fmul_rn(vb.x, g), __fmul_rn(vb.y, g), __fmul_rn(vb.z, g), __fmul_rn(vb.w, g));
    store_attenuated(out, i, va, vb, time);
}

// kernel hrtf (kernel.cpp:586-625)
__global__ void attenuate_hrtf_kernel(const rvb_impulse *in, size_t n, f3 mic, const float4 *table, f3 pointing, f3 up,
                                      unsigned long long channel, rvb_attenuated_impulse *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 va, vb; f3 pos; float time;
    if (!load_impulse(in, i, va, vb, pos, time)) { store_attenuated(out, i, make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), 0.f); return; }
    long idx = hrtf_index(pointing, up, normalize(pos - mic));
    float4 ga = __ldg(table + 2 * idx), gb = __ldg(table + 2 * idx + 1);
    f3 ear = hrtf_ear(pointing, up, mic, channel);
    va = make_float4(__fmul_rn(va.x, ga.x), __fmul_rn(va.y, ga.y), __fmul_rn(va.z, ga.z), __fmul_rn(va.w, ga.w));
    vb = make_float4(__fmul_rn(vb.x, gb.x), __fmul_rn(vb.y, gb.y), __fmul_rn(vb.z, gb.z), __fmul_rn(vb.w, gb.w));
    store_attenuated(out, i, va, vb, hrtf_time(time, pos, mic, ear));
}

__global__ void normalize_dir_kernel(f3 d, float *out)
{
    f3 n = normalize(d);
    out[0] = n.x; out[1] = n.y; out[2] = n.z;
}

}  // namespace rvb

using namespace rvb;

// chunked H2D -> kernel -> D2H so arbitrarily large host arrays work with bounded device memory
template <class Launch>
static int run_chunked(const rvb_impulse *in, size_t n, rvb_attenuated_impulse *out, Launch launch)
{
    cudaStream_t st = lib_stream();
    const size_t chunk = std::min<size_t>(n, (size_t)4 << 20);   // 4 Mi impulses = 256 MiB per buffer
    rvb_impulse *d_in = nullptr;
    rvb_attenuated_impulse *d_out = nullptr;
    RVB_TRY(dev_alloc(&d_in, chunk));
    RVB_TRY(dev_alloc(&d_out, chunk));
    for (size_t o = 0; o < n; o += chunk) {
        size_t m = std::min(chunk, n - o);
        RVB_CUDA(cudaMemcpyAsync(d_in, in + o, m * sizeof(rvb_impulse), cudaMemcpyHostToDevice, st));
        launch(d_in, m, d_out, st);
        RVB_CUDA(cudaGetLastError());
        RVB_CUDA(cudaMemcpyAsync(out + o, d_out, m * sizeof(rvb_attenuated_impulse), cudaMemcpyDeviceToHost, st));
        RVB_CUDA(cudaStreamSynchronize(st));
    }
    dev_free(d_in); dev_free(d_out);
    return RVB_OK;
}

extern "C" {

int rvb_attenuate_speaker(const rvb_impulse *in, size_t n, const float mic[3], const rvb_speaker *speaker, rvb_attenuated_impulse *out)
{
    if (!mic || !speaker || (n && (!in || !out))) return fail(RVB_ERR_INVALID, "null argument");
    RVB_TRY(ensure_device());
    if (n == 0) return RVB_OK;
    cudaStream_t st = lib_stream();
    // normalize(speaker.direction) once, with the device's own arithmetic (kernel.cpp:511)
    float *d_n = nullptr, h_n[3];
    RVB_TRY(dev_alloc(&d_n, 3));
    note_launch(), normalize_dir_kernel<<<1, 1, 0, st>>>(mk3(speaker->direction.s[0], speaker->direction.s[1], speaker->direction.s[2]), d_n);
    RVB_CUDA(cudaMemcpyAsync(h_n, d_n, sizeof h_n, cudaMemcpyDeviceToHost, st));
    RVB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_n);
    f3 m = {mic[0], mic[1], mic[2]}, sd = {h_n[0], h_n[1], h_n[2]};
    float coef = speaker->coefficient;
    return run_chunked(in, n, out, [&](const rvb_impulse *di, size_t cnt, rvb_attenuated_impulse *dout, cudaStream_t s) {
        note_launch(), attenuate_speaker_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(di, cnt, m, sd, coef, dout);
    });
}

int rvb_attenuate_hrtf(const rvb_impulse *in, size_t n, const float mic[3], const rvb_float8 *table_one_ear, const float facing[3],
                       const float up[3], uint64_t channel, rvb_attenuated_impulse *out)
{
    if (!mic || !table_one_ear || !facing || !up || (n && (!in || !out))) return fail(RVB_ERR_INVALID, "null argument");
    RVB_TRY(ensure_device());
    if (n == 0) return RVB_OK;
    cudaStream_t st = lib_stream();
    // one extra row: elevation index 180 reads into the next azimuth row; the last
    // row's overflow reads zeros here instead of running off the table.
    const size_t cells = 360 * 180;
    float4 *d_table = nullptr;
    RVB_TRY(dev_alloc(&d_table, 2 * (cells + 180)));
    RVB_CUDA(cudaMemsetAsync(d_table + 2 * cells, 0, 2 * 180 * sizeof(float4), st));
    RVB_CUDA(cudaMemcpyAsync(d_table, table_one_ear, cells * sizeof(rvb_float8), cudaMemcpyHostToDevice, st));
    f3 m = {mic[0], mic[1], mic[2]}, f = {facing[0], facing[1], facing[2]}, u = {up[0], up[1], up[2]};
    int rc = run_chunked(in, n, out, [&](const rvb_impulse *di, size_t cnt, rvb_attenuated_impulse *dout, cudaStream_t s) {
        note_launch(), attenuate_hrtf_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(di, cnt, m, d_table, f, u, (unsigned long long)channel, dout);
    });
    dev_free(d_table);
    return rc;
}

}  // extern "C"
