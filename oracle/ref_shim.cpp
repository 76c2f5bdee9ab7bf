// C entry points over the UNMODIFIED reference rayverb/filters.cpp, built into
// oracle/_ref/libref_filters.so by oracle/Makefile (only when /root/reference
// is present).  Test infrastructure: used to pin oracle/rayverb_oracle.c's
// filter restatement against the reference's own code.
#include "filters.h"
#include <cstdlib>
#include <cstdio>
#include <cstdint>

static void no_fftw() { std::fprintf(stderr, "ref_shim: FFTW path is not available\n"); std::abort(); }
extern "C" {
float *fftwf_alloc_real(size_t) { no_fftw(); return nullptr; }
fftwf_complex *fftwf_alloc_complex(size_t) { no_fftw(); return nullptr; }
fftwf_plan fftwf_plan_dft_r2c_1d(int, float *, fftwf_complex *, unsigned) { no_fftw(); return nullptr; }
fftwf_plan fftwf_plan_dft_c2r_1d(int, fftwf_complex *, float *, unsigned) { no_fftw(); return nullptr; }
void fftwf_execute(const fftwf_plan) { no_fftw(); }
void fftwf_destroy_plan(fftwf_plan) {}
void fftwf_free(void *) {}

// data: [nch][8][len] contiguous; filtered in place through RayverbFiltering::filter
int ref_filter(int ft, float *data, uint64_t nch, uint64_t len, float sr, float lo_cutoff)
{
    if (ft == RayverbFiltering::FILTER_TYPE_WINDOWED_SINC) return -1;
    std::vector<std::vector<std::vector<float>>> d(nch, std::vector<std::vector<float>>(8, std::vector<float>(len)));
    for (uint64_t c = 0; c < nch; ++c)
        for (int b = 0; b < 8; ++b)
            std::copy(data + (c * 8 + b) * len, data + (c * 8 + b + 1) * len, d[c][b].begin());
    RayverbFiltering::filter((RayverbFiltering::FilterType)ft, d, sr, lo_cutoff);
    for (uint64_t c = 0; c < nch; ++c)
        for (int b = 0; b < 8; ++b)
            std::copy(d[c][b].begin(), d[c][b].end(), data + (c * 8 + b) * len);
    return 0;
}
}
