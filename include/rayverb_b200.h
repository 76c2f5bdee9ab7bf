/*
 * rayverb_b200.h -- C ABI of the B200-native rayverb hot path (librayverb_b200.so).
 *
 * This is the drop-in boundary: the entry points are what the reference's host classes
 * (Raytracer / SpeakerAttenuator / HrtfAttenuator, rayverb/rayverb.h) would bind instead of their
 * OpenCL kernels and cl::Buffer plumbing.  Plain pointers and sizes only; every function returns
 * 0 on success or a negative rvb_status, and rvb_last_error() returns a thread-local message.
 * There is NO CPU fallback: every compute entry point fails with RVB_ERR_CUDA when no sm_100
 * device is usable.
 *
 * File:line citations refer to the upstream tree (reuk/parallel-reverb-raytracer).
 */
#ifndef RAYVERB_B200_H
#define RAYVERB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVB_NUM_IMAGE_SOURCE 10      /* rayverb/clstructs.h:4 */
#define RVB_NBANDS 8                 /* VolumeType = cl_float8, clstructs.h:13 */
#define RVB_SPEED_OF_SOUND 340.0f    /* clstructs.h:5 */

/* ---- POD layouts, byte-identical to rayverb/clstructs.h:13-58 ------------------------------ */
typedef struct { float s[4]; } rvb_float3;                                   /* cl_float3, 16 B */
typedef struct { float s[8]; } rvb_float8;                                   /* cl_float8, 32 B */
typedef struct { uint64_t surface, v0, v1, v2; } rvb_triangle;               /* Triangle,  32 B */
typedef struct { rvb_float8 specular, diffuse; } rvb_surface;                /* Surface,   64 B */
typedef struct { rvb_float8 volume; rvb_float3 position; float time; float pad_[3]; } rvb_impulse;       /* Impulse, 64 B */
typedef struct { rvb_float8 volume; float time; float pad_[7]; } rvb_attenuated_impulse;                 /* AttenuatedImpulse, 64 B */
typedef struct { rvb_float3 direction; float coefficient; float pad_[3]; } rvb_speaker;                  /* Speaker, 32 B */

typedef enum {
    RVB_OK = 0,
    RVB_ERR_INVALID = -1,      /* bad argument (std::runtime_error in the C++ layer)      */
    RVB_ERR_CUDA = -2,         /* device failure / no device (cl::Error in the reference) */
    RVB_ERR_UNSUPPORTED = -3,
    RVB_ERR_NOMEM = -4
} rvb_status;

/* RayverbFiltering::FilterType, rayverb/filters.h:196-201 */
typedef enum { RVB_FILTER_WINDOWED_SINC = 0, RVB_FILTER_BIQUAD_ONEPASS = 1,
               RVB_FILTER_BIQUAD_TWOPASS = 2, RVB_FILTER_LINKWITZ_RILEY = 3 } rvb_filter_type;
/* OutputMode, rayverb/config.h:334-345 */
typedef enum { RVB_OUTPUT_ALL = 0, RVB_OUTPUT_IMAGE_ONLY = 1, RVB_OUTPUT_DIFFUSE_ONLY = 2 } rvb_output_mode;
/* AttenuationModel::Mode, rayverb/config.h:411-430 */
typedef enum { RVB_MODEL_SPEAKER = 0, RVB_MODEL_HRTF = 1 } rvb_attenuation_model;
/* scene acceleration structure (no reference counterpart: kernel.cpp:167-192 is brute force) */
typedef enum { RVB_ACCEL_AUTO = 0, RVB_ACCEL_BRUTE_FORCE = 1, RVB_ACCEL_BVH = 2 } rvb_accel;

typedef struct rvb_scene rvb_scene;   /* device-resident scene (replaces cl_triangles/cl_vertices/cl_surfaces, rayverb.h:184-190) */
typedef struct rvb_trace rvb_trace;   /* device-resident results (replaces storedDiffuse/imageSourceTally, rayverb.h:218-219) */
typedef struct rvb_ir rvb_ir;         /* flattened / filtered impulse response */

/* ---- context ------------------------------------------------------------------------------- */
/* Replaces ContextProvider / KernelLoader (rayverb.cpp:151-192): kernels are AOT-compiled for
 * sm_100a, so there is nothing to build at run time. */
const char *rvb_last_error(void);
int rvb_device_count(int *count);
int rvb_set_device(int device);          /* device used by the calling thread's later calls */
int rvb_version(void);
/* Run this thread's later calls on the caller's CUDA stream (e.g. torch's current stream) instead
 * of the library's own; use = 0 switches back.  The reference has one in-order queue per object
 * (rayverb.cpp:190-191). */
int rvb_set_stream(void *cuda_stream, int use);
/* number of kernels this library has launched so far in the process (every kernel it launches is its own:
 * there are no library sort / select passes) */
unsigned long long rvb_launch_count(void);
/* Device blocks of 1 MiB and more are kept in an exact-size cache when a trace / render frees them, so that
 * the next trace of the same shape re-uses them (the reference re-creates its cl::Buffers for every group
 * of rays, rayverb.cpp:560-583).  The cache holds at most RVB_CACHE_LIMIT_MB (environment, default 49152)
 * of idle blocks per device; this call returns all idle blocks of the calling thread's device to the
 * driver -- for a process that shares the GPU with another allocator. */
int rvb_trim_cache(void);

/* ---- scene: Raytracer::Raytracer (rayverb.h:142-148, rayverb.cpp:242-293) -------------------- */
int rvb_scene_create(const rvb_triangle *triangles, size_t ntriangles,
                     const rvb_float3 *vertices, size_t nvertices,
                     const rvb_surface *surfaces, size_t nsurfaces,
                     int accel /* rvb_accel */, rvb_scene **out);
int rvb_scene_destroy(rvb_scene *scene);
/* which structure the scene uses (RVB_ACCEL_BRUTE_FORCE or RVB_ACCEL_BVH) */
int rvb_scene_accel(const rvb_scene *scene, int *accel);

/* ---- trace: Raytracer::raytrace (rayverb.h:159-164, rayverb.cpp:538-685; kernel.cpp:304-503) -- */
/* flags */
#define RVB_TRACE_DEBUG 1u          /* also record per-bounce hit triangle + per-ray reflection count */
#define RVB_TRACE_NO_IMAGES 2u      /* skip the image-source branch (kernel.cpp:379-457); images = {} */
/* BVH scenes, testing: bits 8-15 = number of work units a ray's steady-phase reflections are cut into by the
   wavefront bounce kernel (0: chosen from the ray count; 1: whole rays).  The results do not depend on it. */
#define RVB_TRACE_UNITS(n) (((unsigned)(n) & 0xffu) << 8)
/* BVH scenes, testing: bits 16-23 = cells per triangle of the microphone visibility map (0: chosen from the query count).
   The map only pre-selects candidates for the reference's exact test: the results do not depend on it. */
#define RVB_TRACE_MAP_CELLS(n) (((unsigned)(n) & 0xffu) << 16)
/* testing: de-duplicate the image sources with the radix-sort path even when the few-candidates kernel applies */
#define RVB_TRACE_SORT_DEDUPE 4u
/* `ray_offset` is the global index of directions[0] (used when rays are sharded over GPUs: only the
 * shard with ray_offset == 0 emits the direct path, and "lowest ray wins" stays well defined). */
int rvb_trace_run(rvb_scene *scene, const float mic[3], const float source[3],
                  const rvb_float3 *directions, size_t nrays, size_t nreflections,
                  uint64_t ray_offset, unsigned flags, rvb_trace **out);
/* same, with `directions` already resident on the current device */
int rvb_trace_run_device(rvb_scene *scene, const float mic[3], const float source[3],
                         const rvb_float3 *d_directions, size_t nrays, size_t nreflections,
                         uint64_t ray_offset, unsigned flags, rvb_trace **out);
/* Parameter sweeps (cmd/main.cpp:241-297 executed once per source / microphone pair): `njobs` pairs
 * over the same scene and the same device-resident directions, traced together in one launch pair --
 * a few thousand rays per pair do not fill a B200 on their own.  mics / sources: njobs x 3 floats.
 * out[njobs] receives one trace per pair, each identical (bit for bit) to what rvb_trace_run_device
 * returns for that pair alone; their kernel-time statistics are those of the whole batch. */
int rvb_trace_run_batch(rvb_scene *scene, size_t njobs, const float *mics, const float *sources,
                        const rvb_float3 *d_directions, size_t nrays, size_t nreflections,
                        unsigned flags, rvb_trace **out);
int rvb_trace_destroy(rvb_trace *trace);

/* getRawDiffuse (rayverb.cpp:687-690): nrays*nreflections impulses, ray-major, zeros where a ray
 * had already left the mesh */
int rvb_trace_diffuse_count(const rvb_trace *trace, size_t *count);
int rvb_trace_copy_diffuse(const rvb_trace *trace, rvb_impulse *out);
/* getRawImages (rayverb.cpp:692-706): de-duplicated image-source impulses in key order.
 * keys (optional): count*10 entries, triangle index + 1 per reflection, zero padded;
 * key_lengths (optional): count entries. */
int rvb_trace_image_count(const rvb_trace *trace, int remove_direct, size_t *count);
int rvb_trace_copy_images(const rvb_trace *trace, int remove_direct, rvb_impulse *out,
                          uint32_t *keys, int32_t *key_lengths);
/* replace the image list (after a cross-GPU merge; n = 0 on non-root shards) */
int rvb_trace_set_images(rvb_trace *trace, const rvb_impulse *images, size_t n);
/* merge per-shard image lists given in shard order: the first occurrence of a key wins
 * (rayverb.cpp:669-673 with contiguous ray shards).  Output sorted by key; returns count in *nout. */
int rvb_images_merge(const rvb_impulse *images, const uint32_t *keys, const int32_t *key_lengths,
                     size_t n, rvb_impulse *out, uint32_t *out_keys, int32_t *out_key_lengths, size_t *nout);
/* debug (needs RVB_TRACE_DEBUG): hit_triangle[nrays*nreflections] ray-major (0xffffffff = none),
 * reflection_count[nrays] */
int rvb_trace_copy_debug(const rvb_trace *trace, uint32_t *hit_triangle, uint32_t *reflection_count);

typedef struct {
    uint64_t nrays, nreflections;
    uint64_t tri_tests;        /* ray-triangle tests executed (52 FLOP each, kernel.cpp:63-88) */
    uint64_t box_tests;        /* BVH node tests executed (0 for brute force)                   */
    uint64_t queries;          /* full-scene queries (nearest-hit + visibility)                 */
    uint64_t bounces;          /* reflections actually executed (<= nrays*nreflections)         */
    float trace_kernel_ms;     /* CUDA-event time of the bounce kernel(s)                       */
    float total_ms;            /* CUDA-event time of rvb_trace_run incl. H2D + image dedupe     */
    float image_phase_ms;      /* reflections 0..8 (image-source branch), 0 when skipped        */
    float steady_phase_ms;     /* reflections 9.. (the dominant kernel)                         */
    uint64_t steady_tri_tests; /* ray-triangle tests executed by the steady-phase kernel        */
    uint64_t steady_box_tests;
    uint64_t emitted;          /* diffuse impulses with a non-zero volume (the rest are skipped
                                  by the attenuators, kernel.cpp:524,607)                        */
    uint64_t micmap_entries;   /* BVH scenes: triangle references in the microphone visibility map */
    uint32_t micmap_cells;     /* ... and its cells per cube-face edge (0: no map)               */
    uint32_t reserved_;
} rvb_trace_stats;
int rvb_trace_get_stats(const rvb_trace *trace, rvb_trace_stats *stats);

/* ---- attenuators on host arrays (drop-in for the reference's per-channel calls) --------------- */
/* SpeakerAttenuator::attenuate (rayverb.cpp:856-892; kernel attenuate, kernel.cpp:506-535) */
int rvb_attenuate_speaker(const rvb_impulse *in, size_t n, const float mic[3],
                          const rvb_speaker *speaker, rvb_attenuated_impulse *out);
/* HrtfAttenuator::attenuate (rayverb.cpp:765-818; kernel hrtf, kernel.cpp:538-625).
 * table_one_ear: [360][180] float8 flattened a*180+e (rayverb.cpp:774-780). */
int rvb_attenuate_hrtf(const rvb_impulse *in, size_t n, const float mic[3],
                       const rvb_float8 *table_one_ear, const float facing[3], const float up[3],
                       uint64_t channel, rvb_attenuated_impulse *out);

/* ---- post-processing on host arrays ------------------------------------------------------------ */
/* flattenImpulses (rayverb.cpp:48-77) for one channel as a deterministic sort-by-sample +
 * segmented reduction on the device.  *length = round(maxtime*sr)+1; out (length*8 floats,
 * band-major [8][length]) is allocated by the library: release with rvb_free. */
int rvb_flatten(const rvb_attenuated_impulse *in, size_t n, float sample_rate, float **out, size_t *length);
/* RayverbFiltering::filter (filters.cpp:268-306) on [nchannels][8][length], in place */
int rvb_filter(int filter_type, float *data, size_t nchannels, size_t length, float sample_rate, float lo_cutoff);
/* BandpassWindowedSinc (filters.cpp:100-131, FILTER_TYPE_WINDOWED_SINC): out of place because the
 * data GROWS -- in [nchannels][8][length] -> out [nchannels][8][29 + length - 1], scaled by the FFT
 * length like the reference's un-normalised FFTW round trip. */
#define RVB_SINC_KERNEL_LENGTH 29
int rvb_filter_sinc(const float *in, size_t nchannels, size_t length, float sample_rate, float lo_cutoff, float *out);
/* process (rayverb.cpp:125-149): filter + mixdown + normalize + scale + trimTail.
 * With RVB_FILTER_WINDOWED_SINC every out[c] must hold 29 + lengths[0] - 1 samples and `data` is
 * left untouched (the filtered rows no longer fit).
 * data: nchannels pointers to [8][lengths[c]] (modified like the reference does);
 * out: nchannels pointers to lengths[c] floats; *out_length = trimmed common length or -1. */
int rvb_process(int filter_type, float *const *data, const size_t *lengths, size_t nchannels,
                float sample_rate, int do_normalize, float lo_cutoff, int do_trim_tail,
                float volume_scale, float *const *out, int64_t *out_length);
void rvb_free(void *p);

/* ---- fused device-resident path: trace results -> impulse response ------------------------------ */
/* Everything cmd/main.cpp:254-336 does between raytrace() and write_sndfile(), without the impulses
 * ever leaving the GPU. */
typedef struct {
    int model;                       /* rvb_attenuation_model                                     */
    const rvb_speaker *speakers;     /* model == SPEAKER: nspeakers channels                      */
    size_t nspeakers;
    const rvb_float8 *hrtf_table;    /* model == HRTF: [2][360][180] float8 (rayverb.h:255-257)   */
    float facing[3], up[3];
    float sample_rate;
    int filter_type;                 /* rvb_filter_type                                           */
    float lo_cutoff;                 /* "hipass" key, cmd/main.cpp:147,189                        */
    int do_normalize;
    float volume_scale;              /* "volumme_scale", cmd/main.cpp:151,191                     */
    int do_trim_tail;
    int trim_predelay;               /* fixPredelay, cmd/main.cpp:324-325                         */
    int remove_direct;
    int output_mode;                 /* rvb_output_mode                                           */
} rvb_render_config;

/* one call: attenuate -> (predelay) -> flatten -> filter -> mixdown -> normalize -> trim */
int rvb_render_ir(rvb_trace *trace, const rvb_render_config *cfg, rvb_ir **out);

/* staged form, for ray shards on several GPUs (the caller runs the collectives between stages):
 *   rvb_render_times      local min non-zero / max attenuated time per channel
 *   rvb_render_histogram  -> per-channel [8][length] histograms on the device, given the global
 *                            predelay and lengths
 *   (caller reduces the histograms over shards, e.g. torch.distributed / NCCL on d_hist)
 *   rvb_render_finish     filter/mixdown/normalize/trim on the reduced histogram (root shard)   */
int rvb_render_times(rvb_trace *trace, const rvb_render_config *cfg, float *min_nonzero_time /*[nch]*/,
                     float *max_time /*[nch]*/, size_t *nchannels);
int rvb_render_histogram(rvb_trace *trace, const rvb_render_config *cfg, float predelay,
                         const size_t *lengths /*[nch]*/, float *d_hist /* device, sum(8*lengths[c]) floats */);
int rvb_render_finish(const rvb_render_config *cfg, const float *d_hist, const size_t *lengths,
                      size_t nchannels, rvb_ir **out);

int rvb_ir_channels(const rvb_ir *ir, size_t *nchannels);
int rvb_ir_length(const rvb_ir *ir, size_t channel, size_t *length);
int rvb_ir_copy(const rvb_ir *ir, size_t channel, float *out);
/* per-band data before mixdown, [8][band_length] for `channel` (parity checks); only kept when
 * rvb_render_keep_bands(1) was called on this thread before the render (default off: it costs a
 * device-to-host copy of 8x the IR) */
int rvb_render_keep_bands(int on);
int rvb_ir_band_length(const rvb_ir *ir, size_t channel, size_t *length);
int rvb_ir_copy_bands(const rvb_ir *ir, size_t channel, float *out);
int rvb_ir_destroy(rvb_ir *ir);

/* per-stage CUDA-event timings of the last rvb_render_* call on this thread (ms) */
typedef struct { float keys_ms, sort_ms, reduce_ms, filter_ms, finish_ms, total_ms; uint64_t n_impulses; } rvb_render_stats;
int rvb_render_get_stats(rvb_render_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* RAYVERB_B200_H */
