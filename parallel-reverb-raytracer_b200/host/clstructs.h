// POD types of the rayverb API (mirrors rayverb/clstructs.h:1-58 of the reference), expressed
// over the C ABI's layouts so host code written against the reference keeps compiling:
// cl_float3 is 16 bytes (.s[0..3]), cl_float8 is 32 bytes (.s[0..7]).
#pragma once
#include "../../include/rayverb_b200.h"

#define NUM_IMAGE_SOURCE RVB_NUM_IMAGE_SOURCE
#define SPEED_OF_SOUND RVB_SPEED_OF_SOUND

typedef float cl_float;
typedef uint64_t cl_ulong;
typedef rvb_float3 cl_float3;
typedef rvb_float8 cl_float8;

/// Type used for storing multiband volumes.
typedef cl_float8 VolumeType;
/// Offsets into the Surface array and the vertex array.
typedef rvb_triangle Triangle;
/// Per-band specular and diffuse coefficients.
typedef rvb_surface Surface;
/// volume, position (for attenuation), time in seconds
typedef rvb_impulse Impulse;
typedef rvb_attenuated_impulse AttenuatedImpulse;
/// direction + polar-pattern coefficient (0 omni .. 1 bidirectional)
typedef rvb_speaker Speaker;
