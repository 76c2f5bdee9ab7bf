// C ABI around the bounce kernel: Raytracer::raytrace / getRawDiffuse / getRawImages / getAllRaw
// (rayverb/rayverb.cpp:538-714).
//
// Unlike the reference, which loops over groups of 4096 rays with ~64 MiB of blocking
// host<->device traffic per group (rayverb.cpp:587-684), the whole ray set is traced in one
// launch and the results stay on the device behind the rvb_trace handle; the host copies below
// exist for the drop-in getRaw* accessors and the parity tests.
#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <map>
#include <vector>

#include "rvb_internal.cuh"

namespace rvb {

// bounce-major SoA -> ray-major AoS (Impulse, clstructs.h:36-42) for a range of rays
__global__ void export_diffuse_kernel(const float4 *vol, const float4 *pt, uint32_t nrays, uint32_t nrefl,
                                      uint32_t ray0, uint32_t nr, rvb_impulse *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over nr*nrefl, bounce fastest
    if (i >= (size_t)nr * nrefl) return;
    uint32_t r = (uint32_t)(i / nrefl), b = (uint32_t)(i % nrefl);
    size_t slot = slot_of(ray0 + r, b, nrays);
    float4 a = vol[2 * slot], c = vol[2 * slot + 1], p = pt[slot];
    float4 *o = reinterpret_cast<float4 *>(out + i);
    o[0] = a; o[1] = c; o[2] = make_float4(p.x, p.y, p.z, 0.f); o[3] = make_float4(p.w, 0.f, 0.f, 0.f);
}

// reflection count not a multiple of the slot group: the tail of the last group holds no impulses
__global__ void mark_unused_slots_kernel(float *tkey, uint32_t nrays, uint32_t nrefl)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrays) return;
    for (uint32_t b = nrefl; b % RVB_SLOT_GROUP != 0; ++b) tkey[slot_of(r, b, nrays)] = -1.0f;
}

__global__ void export_images_kernel(const float4 *volA, const float4 *volB, const float4 *pt, uint32_t n, rvb_impulse *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pt[i];
    float4 *o = reinterpret_cast<float4 *>(out + i);
    o[0] = volA[i]; o[1] = volB[i]; o[2] = make_float4(p.x, p.y, p.z, 0.f); o[3] = make_float4(p.w, 0.f, 0.f, 0.f);
}

__global__ void import_images_kernel(const rvb_impulse *in, uint32_t n, float4 *volA, float4 *volB, float4 *pt)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *s = reinterpret_cast<const float4 *>(in + i);
    float4 p = s[2];
    volA[i] = s[0]; volB[i] = s[1]; pt[i] = make_float4(p.x, p.y, p.z, s[3].x);
}

__global__ void export_hits_kernel(const uint32_t *hit, uint32_t nrays, uint32_t nrefl, uint32_t *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrays * nrefl) return;
    uint32_t r = (uint32_t)(i / nrefl), b = (uint32_t)(i % nrefl);
    out[i] = hit[slot_of(r, b, nrays)];
}

// Traces `njobs` (microphone, source) pairs over the same scene and directions in one launch pair.
static int trace_run_impl(rvb_scene *scene, size_t njobs, const float *mics /* [njobs][3] */, const float *sources /* [njobs][3] */,
                          const float4 *d_dirs, size_t nrays, size_t nrefl, uint64_t ray_offset, unsigned flags, rvb_trace **out)
{
    cudaStream_t st = lib_stream();
    const size_t nslots = slot_count(nrays, nrefl);
    const bool images = !(flags & RVB_TRACE_NO_IMAGES);
    const size_t ncand = images ? nrays * RVB_NUM_IMAGE_SOURCE : 0;
    std::vector<rvb_trace *> trs(njobs, nullptr);
    std::vector<CandBuffers> cands(njobs);
    int rc = RVB_OK;
    StageTimer t_all("trace.total");
    auto bail = [&](int code) {
        for (size_t j = 0; j < njobs; ++j) {
            dev_free(cands[j].volA); dev_free(cands[j].volB); dev_free(cands[j].pt); dev_free(cands[j].idx);
            if (trs[j]) rvb_trace_destroy(trs[j]);
            out[j] = nullptr;
        }
        return code;
    };
    {
    StageTimer t_alloc("trace.alloc");
    for (size_t j = 0; j < njobs; ++j) {
        rvb_trace *tr = new rvb_trace();
        trs[j] = tr;
        tr->scene = scene; tr->nrays = nrays; tr->nrefl = nrefl; tr->ray_offset = ray_offset; tr->flags = flags;
        for (int i = 0; i < 3; ++i) { tr->mic[i] = mics[3 * j + i]; tr->src[i] = sources[3 * j + i]; }
        tr->stats.nrays = nrays; tr->stats.nreflections = nrefl;
        if ((rc = dev_alloc(&tr->d_vol, 2 * nslots)) || (rc = dev_alloc(&tr->d_pt, nslots)) || (rc = dev_alloc(&tr->d_tkey, nslots)))
            return bail(rc);
        if ((nrefl % RVB_SLOT_GROUP) && nrays) note_launch(), mark_unused_slots_kernel<<<(unsigned)((nrays + 255) / 256), 256, 0, st>>>(tr->d_tkey, (uint32_t)nrays, (uint32_t)nrefl);
        if (flags & RVB_TRACE_DEBUG) {
            if ((rc = dev_alloc(&tr->d_hit, nslots)) || (rc = dev_alloc(&tr->d_count, nrays))) return bail(rc);
        }
        if (images && ncand) {
            CandBuffers &cb = cands[j];
            if ((rc = dev_alloc(&cb.volA, ncand)) || (rc = dev_alloc(&cb.volB, ncand)) || (rc = dev_alloc(&cb.pt, ncand)) || (rc = dev_alloc(&cb.idx, ncand)))
                return bail(rc);
            cudaMemsetAsync(cb.idx, 0, ncand * sizeof(uint32_t), st);
            // slot 0 of ray 0 is read even if the direct path is occluded (key {0} is unconditional)
            cudaMemsetAsync(cb.volA, 0, sizeof(float4), st); cudaMemsetAsync(cb.volB, 0, sizeof(float4), st); cudaMemsetAsync(cb.pt, 0, sizeof(float4), st);
        }
    }
    }
    {
        StageTimer t_k("trace.kernel");
        std::vector<Trace *> base(trs.begin(), trs.end());
        if ((rc = trace_launch(base.data(), njobs, d_dirs, cands.data()))) return bail(rc);
    }
    if (images && ncand) {
        StageTimer t_d("trace.dedupe");
        for (size_t j = 0; j < njobs; ++j)
            if ((rc = images_dedupe(trs[j], cands[j].volA, cands[j].volB, cands[j].pt, cands[j].idx))) return bail(rc);
    }
    StageTimer t_f("trace.free");
    for (size_t j = 0; j < njobs; ++j) {
        dev_free(cands[j].volA); dev_free(cands[j].volB); dev_free(cands[j].pt); dev_free(cands[j].idx);
        out[j] = trs[j];
    }
    return RVB_OK;
}

}  // namespace rvb

using namespace rvb;

extern "C" {

int rvb_trace_run_device(rvb_scene *scene, const float mic[3], const float source[3], const rvb_float3 *d_directions,
                         size_t nrays, size_t nreflections, uint64_t ray_offset, unsigned flags, rvb_trace **out)
{
    if (!out) return fail(RVB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!scene || !mic || !source || (!d_directions && nrays)) return fail(RVB_ERR_INVALID, "null argument");
    if (nrays >= (1ull << 32) || nreflections >= (1ull << 32)) return fail(RVB_ERR_UNSUPPORTED, "ray or reflection count too large");
    RVB_CUDA(cudaSetDevice(scene->device));
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    cudaEvent_t e0, e1;
    RVB_CUDA(cudaEventCreate(&e0)); RVB_CUDA(cudaEventCreate(&e1));
    RVB_CUDA(cudaEventRecord(e0, st));
    int rc = trace_run_impl(scene, 1, mic, source, reinterpret_cast<const float4 *>(d_directions), nrays, nreflections, ray_offset, flags, out);
    if (rc == RVB_OK) {
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&(*out)->stats.total_ms, e0, e1);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

int rvb_trace_run_batch(rvb_scene *scene, size_t njobs, const float *mics, const float *sources, const rvb_float3 *d_directions,
                        size_t nrays, size_t nreflections, unsigned flags, rvb_trace **out)
{
    if (!out) return fail(RVB_ERR_INVALID, "out is null");
    for (size_t j = 0; j < njobs; ++j) out[j] = nullptr;
    if (njobs == 0) return RVB_OK;
    if (!scene || !mics || !sources || (!d_directions && nrays)) return fail(RVB_ERR_INVALID, "null argument");
    if (nrays >= (1ull << 32) || nreflections >= (1ull << 32)) return fail(RVB_ERR_UNSUPPORTED, "ray or reflection count too large");
    if (njobs > 65535) return fail(RVB_ERR_UNSUPPORTED, "more than 65535 jobs in one batch");
    RVB_CUDA(cudaSetDevice(scene->device));
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    cudaEvent_t e0, e1;
    RVB_CUDA(cudaEventCreate(&e0)); RVB_CUDA(cudaEventCreate(&e1));
    RVB_CUDA(cudaEventRecord(e0, st));
    int rc = trace_run_impl(scene, njobs, mics, sources, reinterpret_cast<const float4 *>(d_directions), nrays, nreflections, 0, flags, out);
    if (rc == RVB_OK) {
        float ms = 0.f;
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        for (size_t j = 0; j < njobs; ++j) out[j]->stats.total_ms = ms;   // of the whole batch
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

int rvb_trace_run(rvb_scene *scene, const float mic[3], const float source[3], const rvb_float3 *directions, size_t nrays,
                  size_t nreflections, uint64_t ray_offset, unsigned flags, rvb_trace **out)
{
    if (!out) return fail(RVB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!scene || !mic || !source || (!directions && nrays)) return fail(RVB_ERR_INVALID, "null argument");
    RVB_CUDA(cudaSetDevice(scene->device));
    RVB_TRY(ensure_device());
    cudaStream_t st = lib_stream();
    cudaEvent_t e0, e1;
    RVB_CUDA(cudaEventCreate(&e0)); RVB_CUDA(cudaEventCreate(&e1));
    RVB_CUDA(cudaEventRecord(e0, st));
    float4 *d_dirs = nullptr;
    RVB_TRY(dev_alloc(&d_dirs, nrays));
    if (nrays) RVB_CUDA(cudaMemcpyAsync(d_dirs, directions, nrays * sizeof(rvb_float3), cudaMemcpyHostToDevice, st));
    int rc = rvb_trace_run_device(scene, mic, source, reinterpret_cast<const rvb_float3 *>(d_dirs), nrays, nreflections, ray_offset, flags, out);
    dev_free(d_dirs);
    if (rc == RVB_OK) {
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&(*out)->stats.total_ms, e0, e1);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

int rvb_trace_destroy(rvb_trace *tr)
{
    if (!tr) return RVB_OK;
    if (tr->scene) cudaSetDevice(tr->scene->device);
    dev_free(tr->d_vol); dev_free(tr->d_pt); dev_free(tr->d_tkey);
    dev_free(tr->d_hit); dev_free(tr->d_count);
    dev_free(tr->d_img_volA); dev_free(tr->d_img_volB); dev_free(tr->d_img_pt);
    delete tr;
    return RVB_OK;
}

int rvb_trace_diffuse_count(const rvb_trace *tr, size_t *count)
{
    if (!tr || !count) return fail(RVB_ERR_INVALID, "null argument");
    *count = tr->nrays * tr->nrefl;
    return RVB_OK;
}

int rvb_trace_copy_diffuse(const rvb_trace *tr, rvb_impulse *out)
{
    if (!tr || !out) return fail(RVB_ERR_INVALID, "null argument");
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    cudaStream_t st = lib_stream();
    if (tr->nrays == 0 || tr->nrefl == 0) return RVB_OK;
    // stage through a bounded AoS buffer (<= 256 MiB)
    size_t rays_per_chunk = std::max<size_t>(1, (256ull << 20) / (tr->nrefl * sizeof(rvb_impulse)));
    rays_per_chunk = std::min(rays_per_chunk, tr->nrays);
    rvb_impulse *d_stage = nullptr;
    RVB_TRY(dev_alloc(&d_stage, rays_per_chunk * tr->nrefl));
    for (size_t r0 = 0; r0 < tr->nrays; r0 += rays_per_chunk) {
        size_t nr = std::min(rays_per_chunk, tr->nrays - r0);
        size_t n = nr * tr->nrefl;
        note_launch(), export_diffuse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tr->d_vol, tr->d_pt, (uint32_t)tr->nrays,
                                                                           (uint32_t)tr->nrefl, (uint32_t)r0, (uint32_t)nr, d_stage);
        RVB_CUDA(cudaMemcpyAsync(out + r0 * tr->nrefl, d_stage, n * sizeof(rvb_impulse), cudaMemcpyDeviceToHost, st));
        RVB_CUDA(cudaStreamSynchronize(st));
    }
    dev_free(d_stage);
    return RVB_OK;
}

int rvb_trace_image_count(const rvb_trace *tr, int remove_direct, size_t *count)
{
    if (!tr || !count) return fail(RVB_ERR_INVALID, "null argument");
    *count = tr->n_images - ((remove_direct && tr->first_is_direct) ? 1 : 0);   // temp.erase({0}), rayverb.cpp:695-696
    return RVB_OK;
}

int rvb_trace_copy_images(const rvb_trace *tr, int remove_direct, rvb_impulse *out, uint32_t *keys, int32_t *key_lengths)
{
    if (!tr) return fail(RVB_ERR_INVALID, "null argument");
    const size_t skip = (remove_direct && tr->first_is_direct) ? 1 : 0;
    const size_t n = tr->n_images - skip;
    if (n == 0) return RVB_OK;
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    cudaStream_t st = lib_stream();
    if (out) {
        rvb_impulse *d_stage = nullptr;
        RVB_TRY(dev_alloc(&d_stage, n));
        note_launch(), export_images_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tr->d_img_volA + skip, tr->d_img_volB + skip, tr->d_img_pt + skip, (uint32_t)n, d_stage);
        RVB_CUDA(cudaMemcpyAsync(out, d_stage, n * sizeof(rvb_impulse), cudaMemcpyDeviceToHost, st));
        RVB_CUDA(cudaStreamSynchronize(st));
        dev_free(d_stage);
    }
    if (keys && !tr->img_keys.empty()) std::copy(tr->img_keys.begin() + skip * RVB_NUM_IMAGE_SOURCE, tr->img_keys.end(), keys);
    if (key_lengths && !tr->img_len.empty()) std::copy(tr->img_len.begin() + skip, tr->img_len.end(), key_lengths);
    return RVB_OK;
}

int rvb_trace_set_images(rvb_trace *tr, const rvb_impulse *images, size_t n)
{
    if (!tr || (!images && n)) return fail(RVB_ERR_INVALID, "null argument");
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    cudaStream_t st = lib_stream();
    dev_free(tr->d_img_volA); dev_free(tr->d_img_volB); dev_free(tr->d_img_pt);
    tr->n_images = n; tr->img_keys.clear(); tr->img_len.clear(); tr->first_is_direct = false;
    if (n == 0) return RVB_OK;
    rvb_impulse *d_stage = nullptr;
    RVB_TRY(dev_alloc(&d_stage, n));
    RVB_TRY(dev_alloc(&tr->d_img_volA, n)); RVB_TRY(dev_alloc(&tr->d_img_volB, n)); RVB_TRY(dev_alloc(&tr->d_img_pt, n));
    RVB_CUDA(cudaMemcpyAsync(d_stage, images, n * sizeof(rvb_impulse), cudaMemcpyHostToDevice, st));
    note_launch(), import_images_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_stage, (uint32_t)n, tr->d_img_volA, tr->d_img_volB, tr->d_img_pt);
    RVB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_stage);
    return RVB_OK;
}

// Cross-shard merge on the host: the lists are tiny (unique image sources), and the rule is the
// reference's own map insert (rayverb.cpp:669-673): first occurrence of a key wins.
int rvb_images_merge(const rvb_impulse *images, const uint32_t *keys, const int32_t *key_lengths, size_t n, rvb_impulse *out,
                     uint32_t *out_keys, int32_t *out_key_lengths, size_t *nout)
{
    if (!nout || (n && (!images || !keys || !key_lengths || !out))) return fail(RVB_ERR_INVALID, "null argument");
    // std::map<vector<u64>> order = lexicographic with a proper prefix first.  Keys hold unsigned
    // values, so comparing the zero-padded 10-tuples and then the lengths gives the same order.
    struct PKey {
        uint32_t k[RVB_NUM_IMAGE_SOURCE + 1];   // zero-padded key, then its length
        bool operator==(const PKey &o) const { return memcmp(k, o.k, sizeof k) == 0; }
        bool operator<(const PKey &o) const
        {
            for (int j = 0; j <= RVB_NUM_IMAGE_SOURCE; ++j) if (k[j] != o.k[j]) return k[j] < o.k[j];
            return false;
        }
    };
    struct PHash {
        size_t operator()(const PKey &p) const
        {
            uint64_t h = 1469598103934665603ull;
            for (int j = 0; j <= RVB_NUM_IMAGE_SOURCE; ++j) { h ^= p.k[j]; h *= 1099511628211ull; }
            return (size_t)h;
        }
    };
    // first occurrence of every key (O(n) with a hash set), then only the unique ones are sorted
    std::unordered_map<PKey, uint32_t, PHash> first;
    first.reserve(2 * n + 1);
    std::vector<std::pair<PKey, uint32_t>> uniq;
    for (size_t i = 0; i < n; ++i) {
        PKey p;
        for (int j = 0; j < RVB_NUM_IMAGE_SOURCE; ++j) p.k[j] = j < key_lengths[i] ? keys[i * RVB_NUM_IMAGE_SOURCE + j] : 0u;
        p.k[RVB_NUM_IMAGE_SOURCE] = (uint32_t)key_lengths[i];
        if (first.emplace(p, (uint32_t)i).second) uniq.emplace_back(p, (uint32_t)i);
    }
    std::sort(uniq.begin(), uniq.end(), [](const std::pair<PKey, uint32_t> &a, const std::pair<PKey, uint32_t> &b) { return a.first < b.first; });
    size_t m = 0;
    for (const auto &u : uniq) {
        out[m] = images[u.second];
        if (out_keys) memcpy(out_keys + m * RVB_NUM_IMAGE_SOURCE, u.first.k, RVB_NUM_IMAGE_SOURCE * sizeof(uint32_t));
        if (out_key_lengths) out_key_lengths[m] = (int32_t)u.first.k[RVB_NUM_IMAGE_SOURCE];
        ++m;
    }
    *nout = m;
    return RVB_OK;
}

int rvb_trace_copy_debug(const rvb_trace *tr, uint32_t *hit_triangle, uint32_t *reflection_count)
{
    if (!tr) return fail(RVB_ERR_INVALID, "null argument");
    if (!tr->d_hit || !tr->d_count) return fail(RVB_ERR_INVALID, "trace was not run with RVB_TRACE_DEBUG");
    RVB_CUDA(cudaSetDevice(tr->scene->device));
    cudaStream_t st = lib_stream();
    const size_t n = tr->nrays * tr->nrefl;
    if (hit_triangle && n) {
        uint32_t *d_stage = nullptr;
        RVB_TRY(dev_alloc(&d_stage, n));
        note_launch(), export_hits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tr->d_hit, (uint32_t)tr->nrays, (uint32_t)tr->nrefl, d_stage);
        RVB_CUDA(cudaMemcpyAsync(hit_triangle, d_stage, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        RVB_CUDA(cudaStreamSynchronize(st));
        dev_free(d_stage);
    }
    if (reflection_count && tr->nrays) {
        RVB_CUDA(cudaMemcpyAsync(reflection_count, tr->d_count, tr->nrays * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        RVB_CUDA(cudaStreamSynchronize(st));
    }
    return RVB_OK;
}

int rvb_trace_get_stats(const rvb_trace *tr, rvb_trace_stats *stats)
{
    if (!tr || !stats) return fail(RVB_ERR_INVALID, "null argument");
    *stats = tr->stats;
    return RVB_OK;
}

}  // extern "C"
