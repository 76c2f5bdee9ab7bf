// Image-source de-duplication on the device.
//
// Replaces the host loop of Raytracer::raytrace (rayverb/rayverb.cpp:654-676) and the ordered
// walk of getRawImages (:692-706): the reference inserts, for every ray and every prefix length
// k, the key idx[0..k) into a std::map<vector<ulong>, Impulse> unless it is already there, so
// the FIRST (lowest-index) ray wins and the result comes out in lexicographic key order with a
// proper prefix before its extensions.
//
// Here: stable-compact the candidate slots (ray-major), then LSD radix sort on the key columns
// (column value = entry + 1, 0 = "past the end", so prefixes sort first); stability keeps the
// lowest ray first within equal keys; then keep the head of every run of equal keys.
// The compactions (rvb_scan.cuh) and the radix sort (rvb_sort.cuh) are the library's own.
#include <cstring>
#include <vector>

#include "rvb_internal.cuh"
#include "rvb_scan.cuh"
#include "rvb_sort.cuh"

namespace rvb {

struct CandValid {
    const uint32_t *idx;
    uint64_t ray_offset;
    __device__ bool operator()(uint32_t flat) const
    {
        uint32_t k0 = flat % RVB_NUM_IMAGE_SOURCE;
        if (k0 == 0) return ray_offset + flat / RVB_NUM_IMAGE_SOURCE == 0;   // only the very first ray carries key {0}
        return idx[flat] != 0;                                               // rayverb.cpp:667
    }
};

// column j of the key of candidate `flat`: entry+1 inside the key, 0 past its end
__device__ __forceinline__ uint32_t key_col(const uint32_t *idx, uint32_t flat, uint32_t j)
{
    uint32_t k0 = flat % RVB_NUM_IMAGE_SOURCE;
    if (j > k0) return 0u;
    if (j == 0) return 1u;                                                   // idx[..][0] is always 0 (direct slot)
    return idx[flat - k0 + j] + 1u;
}

__global__ void col_keys_kernel(const uint32_t *idx, const uint32_t *sel, const uint32_t *perm, uint32_t n, uint32_t j,
                                uint32_t *keys)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) keys[p] = key_col(idx, sel[perm[p]], j);
}

// all key columns 1..last_col packed into one 64-bit sort key, column 1 most significant
__global__ void packed_keys_kernel(const uint32_t *idx, const uint32_t *sel, uint32_t n, uint32_t last_col, uint32_t bits,
                                   unsigned long long *keys, uint32_t *perm)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t flat = sel[p];
    unsigned long long k = 0;
    for (uint32_t j = 1; j <= last_col; ++j) k = (k << bits) | key_col(idx, flat, j);
    keys[p] = k;
    perm[p] = p;
}

__global__ void head_flags_packed_kernel(const unsigned long long *keys, uint32_t n, uint8_t *head)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) head[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1 : 0;
}

struct HeadSet {
    const uint8_t *head;
    __device__ bool operator()(uint32_t p) const { return head[p] != 0; }
};

__global__ void iota_kernel(uint32_t *a, uint32_t n)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) a[p] = p;
}

__global__ void head_flags_kernel(const uint32_t *idx, const uint32_t *sel, const uint32_t *perm, uint32_t n, uint8_t *head)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    bool h = (p == 0);
    if (!h) {
        uint32_t a = sel[perm[p]], b = sel[perm[p - 1]];
        for (uint32_t j = 0; j < RVB_NUM_IMAGE_SOURCE && !h; ++j) h = key_col(idx, a, j) != key_col(idx, b, j);
    }
    head[p] = h ? 1 : 0;
}

__global__ void gather_unique_kernel(const uint32_t *idx, const uint32_t *sel, const uint32_t *perm, const uint32_t *upos,
                                     uint32_t m, const float4 *cvA, const float4 *cvB, const float4 *cpt, float4 *ovA,
                                     float4 *ovB, float4 *opt, uint32_t *okeys, int32_t *olen)
{
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= m) return;
    uint32_t flat = sel[perm[upos[u]]];
    ovA[u] = cvA[flat]; ovB[u] = cvB[flat]; opt[u] = cpt[flat];
    uint32_t k0 = flat % RVB_NUM_IMAGE_SOURCE;
    olen[u] = (int32_t)k0 + 1;
    for (uint32_t j = 0; j < RVB_NUM_IMAGE_SOURCE; ++j) okeys[u * RVB_NUM_IMAGE_SOURCE + j] = (j <= k0) ? idx[flat - k0 + j] : 0u;
}

// ---- few candidates: the whole de-duplication in ONE CTA ------------------------------------------
// A large scene yields few valid image-source candidates (C3's 100k-triangle hall: a few hundred per
// million rays, a few dozen distinct sequences), and the radix-sort path above then costs ~150 tiny
// launches (0.43 ms, launch-bound).  Up to RVB_DEDUPE_SMALL candidates are instead ranked by direct
// comparison in shared memory: candidate p (compacted, ray-major order) is a HEAD if no earlier
// candidate has the same key -- what the stable sort's "first of a run" selects -- and a head's output
// position is the number of heads with a lexicographically smaller key (past-the-end columns are 0, so a
// proper prefix comes before its extensions): the same list in the same order.
#define RVB_DEDUPE_SMALL 4096u
#define RVB_DEDUPE_THREADS 1024u
#define RVB_DEDUPE_COLS (RVB_NUM_IMAGE_SOURCE - 1)   // column 0 is constant

__device__ __forceinline__ int key_cmp_small(const uint32_t *skeys, uint32_t n, uint32_t a, uint32_t b)
{
#pragma unroll
    for (uint32_t j = 0; j < RVB_DEDUPE_COLS; ++j) {
        const uint32_t ka = skeys[j * n + a], kb = skeys[j * n + b];
        if (ka != kb) return ka < kb ? -1 : 1;
    }
    return 0;
}

__global__ void __launch_bounds__(RVB_DEDUPE_THREADS) dedupe_small_kernel(const uint32_t *idx, const uint32_t *sel, uint32_t n, const float4 *cvA,
                                                                          const float4 *cvB, const float4 *cpt, float4 *ovA, float4 *ovB, float4 *opt,
                                                                          uint32_t *out /* [0] m, [1 + u] key length, [1 + n + 10 u + j] key */)
{
    extern __shared__ uint32_t skeys[];                 // [RVB_DEDUPE_COLS][n], then head flags [n]
    uint32_t *s_head = skeys + (size_t)RVB_DEDUPE_COLS * n;
    const uint32_t tid = threadIdx.x;
    for (uint32_t p = tid; p < n; p += RVB_DEDUPE_THREADS) {
        const uint32_t flat = sel[p];
        for (uint32_t j = 0; j < RVB_DEDUPE_COLS; ++j) skeys[j * n + p] = key_col(idx, flat, j + 1);
    }
    __syncthreads();
    for (uint32_t p = tid; p < n; p += RVB_DEDUPE_THREADS) {
        bool head = true;
        for (uint32_t q = 0; q < p && head; ++q) head = key_cmp_small(skeys, n, q, p) != 0;
        s_head[p] = head ? 1u : 0u;
    }
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t p = tid; p < n; p += RVB_DEDUPE_THREADS) {
        if (!s_head[p]) continue;
        ++mine;
        uint32_t u = 0;
        for (uint32_t q = 0; q < n; ++q)
            if (s_head[q] && key_cmp_small(skeys, n, q, p) < 0) ++u;
        const uint32_t flat = sel[p], k0 = flat % RVB_NUM_IMAGE_SOURCE;
        ovA[u] = cvA[flat]; ovB[u] = cvB[flat]; opt[u] = cpt[flat];
        out[1 + u] = k0 + 1u;
        for (uint32_t j = 0; j < RVB_NUM_IMAGE_SOURCE; ++j) out[1 + n + u * RVB_NUM_IMAGE_SOURCE + j] = (j <= k0) ? idx[flat - k0 + j] : 0u;
    }
    // m = number of heads
    __shared__ uint32_t s_m;
    if (tid == 0) s_m = 0;
    __syncthreads();
    if (mine) atomicAdd(&s_m, mine);
    __syncthreads();
    if (tid == 0) out[0] = s_m;
}

static int dedupe_small(Trace *tr, const float4 *cvA, const float4 *cvB, const float4 *cpt, const uint32_t *cidx, const uint32_t *d_sel, uint32_t n,
                        cudaStream_t st)
{
    // outputs sized for the worst case (every candidate distinct): n <= RVB_DEDUPE_SMALL records
    RVB_TRY(dev_alloc(&tr->d_img_volA, n));
    RVB_TRY(dev_alloc(&tr->d_img_volB, n));
    RVB_TRY(dev_alloc(&tr->d_img_pt, n));
    uint32_t *d_out = nullptr;
    const auto cleanup = on_scope_exit([&] { dev_free(d_out); });
    const size_t words = 1 + (size_t)n * (1 + RVB_NUM_IMAGE_SOURCE);
    RVB_TRY(dev_alloc(&d_out, words));
    const size_t smem = ((size_t)RVB_DEDUPE_COLS + 1) * n * sizeof(uint32_t);
    RVB_CUDA(cudaFuncSetAttribute(dedupe_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)RVB_DEDUPE_COLS + 1) * RVB_DEDUPE_SMALL * sizeof(uint32_t))));
    note_launch(), dedupe_small_kernel<<<1, RVB_DEDUPE_THREADS, smem, st>>>(cidx, d_sel, n, cvA, cvB, cpt, tr->d_img_volA, tr->d_img_volB, tr->d_img_pt, d_out);
    RVB_CUDA(cudaGetLastError());
    std::vector<uint32_t> h(words);
    RVB_TRY(read_back(h.data(), d_out, words * sizeof(uint32_t)));
    const uint32_t m = h[0];
    tr->img_len.resize(m);
    tr->img_keys.resize((size_t)m * RVB_NUM_IMAGE_SOURCE);
    for (uint32_t u = 0; u < m; ++u) tr->img_len[u] = (int32_t)h[1 + u];
    if (m) memcpy(tr->img_keys.data(), h.data() + 1 + n, (size_t)m * RVB_NUM_IMAGE_SOURCE * sizeof(uint32_t));
    tr->n_images = m;
    tr->first_is_direct = (m > 0 && tr->img_len[0] == 1);   // key {0} sorts first when present
    return RVB_OK;
}

int images_dedupe(Trace *tr, const float4 *cvA, const float4 *cvB, const float4 *cpt, const uint32_t *cidx)
{
    cudaStream_t st = lib_stream();
    tr->n_images = 0; tr->img_keys.clear(); tr->img_len.clear(); tr->first_is_direct = false;
    const size_t nflat64 = tr->nrays * RVB_NUM_IMAGE_SOURCE;
    if (nflat64 == 0) return RVB_OK;
    if (nflat64 > 0xffffffffull) return fail(RVB_ERR_UNSUPPORTED, "too many rays for one image-source pass");
    const uint32_t nflat = (uint32_t)nflat64;

    uint32_t *d_sel = nullptr, *d_nsel = nullptr;
    uint32_t *d_perm[2] = {nullptr, nullptr}, *d_keys[2] = {nullptr, nullptr};
    unsigned long long *d_k64[2] = {nullptr, nullptr};
    uint8_t *d_head = nullptr;
    uint32_t *d_okeys = nullptr;
    int32_t *d_olen = nullptr;
    const auto cleanup = on_scope_exit([&] {
        dev_free(d_okeys); dev_free(d_olen); dev_free(d_head);
        for (int i = 0; i < 2; ++i) { dev_free(d_perm[i]); dev_free(d_keys[i]); dev_free(d_k64[i]); }
        dev_free(d_sel); dev_free(d_nsel);
    });
    RVB_TRY(dev_alloc(&d_sel, nflat));
    RVB_TRY(dev_alloc(&d_nsel, 1));
    CandValid pred{cidx, tr->ray_offset};
    RVB_TRY(compact_indices(pred, nflat, d_sel, d_nsel, st));
    uint32_t n = 0;
    RVB_TRY(read_back(&n, d_nsel, sizeof n));
    if (n == 0) return RVB_OK;
    if (n <= RVB_DEDUPE_SMALL && !(tr->flags & RVB_TRACE_SORT_DEDUPE)) {
        return dedupe_small(tr, cvA, cvB, cpt, cidx, d_sel, n, st);
    }

    for (int i = 0; i < 2; ++i) { RVB_TRY(dev_alloc(&d_perm[i], n)); RVB_TRY(dev_alloc(&d_keys[i], n)); }
    const unsigned blk = 256, grd = (n + blk - 1) / blk;
    int end_bit = 1;
    while ((1ull << end_bit) < (unsigned long long)tr->scene->T + 2ull) ++end_bit;
    int cur = 0;
    const uint32_t last_col = (uint32_t)std::min<size_t>(RVB_NUM_IMAGE_SOURCE - 1, tr->nrefl);
    RVB_TRY(dev_alloc(&d_head, n));
    if ((uint32_t)end_bit * last_col <= 64u) {
        // small scenes: every key fits one 64-bit word -> ONE radix sort instead of one per column
        // (stable, so the lowest ray still comes first within equal keys)
        for (int i = 0; i < 2; ++i) RVB_TRY(dev_alloc(&d_k64[i], n));
        note_launch(), packed_keys_kernel<<<grd, blk, 0, st>>>(cidx, d_sel, n, last_col, (uint32_t)end_bit, d_k64[1], d_perm[1]);
        const int rc_sort = rsort::sort_arrays<unsigned long long>(d_k64, d_perm, n, 0, end_bit * (int)last_col, &cur, st);
        RVB_TRY(rc_sort);
        note_launch(), head_flags_packed_kernel<<<grd, blk, 0, st>>>(d_k64[cur], n, d_head);
    } else {
        cur = 1;
        note_launch(), iota_kernel<<<grd, blk, 0, st>>>(d_perm[cur], n);
        for (uint32_t j = last_col; j >= 1; --j) {   // LSD over key columns; column 0 is constant
            // input in buffer pair 1 (sort_arrays), so move the current permutation there first if needed
            if (cur != 1) { std::swap(d_perm[0], d_perm[1]); cur = 1; }
            note_launch(), col_keys_kernel<<<grd, blk, 0, st>>>(cidx, d_sel, d_perm[1], n, j, d_keys[1]);
            const int rc_sort = rsort::sort_arrays<uint32_t>(d_keys, d_perm, n, 0, end_bit, &cur, st);
            RVB_TRY(rc_sort);
        }
        note_launch(), head_flags_kernel<<<grd, blk, 0, st>>>(cidx, d_sel, d_perm[cur], n, d_head);
    }
    uint32_t *d_upos = d_keys[0], *d_m = d_nsel;
    RVB_TRY(compact_indices(HeadSet{d_head}, n, d_upos, d_m, st));
    uint32_t m = 0;
    RVB_TRY(read_back(&m, d_m, sizeof m));

    RVB_TRY(dev_alloc(&tr->d_img_volA, m));
    RVB_TRY(dev_alloc(&tr->d_img_volB, m));
    RVB_TRY(dev_alloc(&tr->d_img_pt, m));
    RVB_TRY(dev_alloc(&d_okeys, (size_t)m * RVB_NUM_IMAGE_SOURCE));
    RVB_TRY(dev_alloc(&d_olen, m));
    note_launch(), gather_unique_kernel<<<(m + blk - 1) / blk, blk, 0, st>>>(cidx, d_sel, d_perm[cur], d_upos, m, cvA, cvB, cpt, tr->d_img_volA,
                                                              tr->d_img_volB, tr->d_img_pt, d_okeys, d_olen);
    RVB_CUDA(cudaGetLastError());
    tr->img_keys.resize((size_t)m * RVB_NUM_IMAGE_SOURCE);
    tr->img_len.resize(m);
    RVB_TRY(read_back(tr->img_keys.data(), d_okeys, tr->img_keys.size() * sizeof(uint32_t)));
    RVB_TRY(read_back(tr->img_len.data(), d_olen, m * sizeof(int32_t)));
    tr->n_images = m;
    tr->first_is_direct = (m > 0 && tr->img_len[0] == 1);   // key {0} sorts first when present
    return RVB_OK;
}

}  // namespace rvb
