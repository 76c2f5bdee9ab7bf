// Device-side restatement of the reference's attenuation kernels, shared by the host-array
// entry points (attenuate.cu) and the fused render path (render.cu).
#pragma once
#include "rvb_math.cuh"

namespace rvb {

// speaker_attenuation (kernel.cpp:506-513) applied to getDirection(mic, position):
// (1 - c) + c * dot(normalize(direction), normalize(speaker.direction))
__device__ __forceinline__ float speaker_gain(f3 mic, f3 position, f3 speaker_dir_normalized, float coefficient)
{
    f3 dir = normalize(position - mic);                       // getDirection (kernel.cpp:299-302)
    float d = dot(normalize(dir), speaker_dir_normalized);    // the kernel normalises again (:511)
    return __fadd_rn(__fsub_rn(1.0f, coefficient), __fmul_rn(coefficient, d));
}

// transform (kernel.cpp:538-549)
__device__ __forceinline__ f3 hrtf_transform(f3 pointing, f3 up, f3 d)
{
    f3 x = normalize(cross(up, pointing));
    f3 y = cross(pointing, x);
    f3 z = pointing;
    return mk3(dot(x, d), dot(y, d), dot(z, d));
}

// hrtf_attenuation (kernel.cpp:563-584): table index a*180+e.  e may reach 180 (row overflow
// quirk); the index is returned as the reference computes it.
__device__ __forceinline__ long hrtf_index(f3 pointing, f3 up, f3 impulse_direction)
{
    f3 t = hrtf_transform(pointing, up, impulse_direction);
    float az = atan2f(t.x, t.z);                                                       // :552-555
    float el = atan2f(t.y, __fsqrt_rn(__fmaf_rn(t.z, t.z, __fmul_rn(t.x, t.x))));      // :558-561, length(d.xz)
    long a = (long)__fadd_rn(__fmul_rn(RVB_DEG_F, az), 180.0f);                        // :578
    a %= 360;
    long e = (long)__fmul_rn(RVB_DEG_F, el);
    e = 90 - e;
    return a * 180 + e;
}

// ear position (kernel.cpp:597-603): transform(pointing, up, (-/+0.1, 0, 0)) + mic
__device__ __forceinline__ f3 hrtf_ear(f3 pointing, f3 up, f3 mic, unsigned long long channel)
{
    const float WIDTH = 0.1f;
    return hrtf_transform(pointing, up, mk3(channel == 0 ? -WIDTH : WIDTH, 0.0f, 0.0f)) + mic;
}

// time shift of kernel hrtf (kernel.cpp:616-622)
__device__ __forceinline__ float hrtf_time(float time, f3 position, f3 mic, f3 ear)
{
    float dist0 = length(position - mic);
    float dist1 = length(position - ear);
    float diff = __fsub_rn(dist1, dist0);
    return __fadd_rn(time, __fmul_rn(diff, RVB_SECONDS_PER_METER));
}

}  // namespace rvb
