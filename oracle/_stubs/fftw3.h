/* Stand-in for <fftw3.h> (FFTW3 is not installed in this image).
 * Written for this repo; declares only what rayverb/filters.h names so that
 * the reference's filters.cpp can be compiled UNMODIFIED into oracle/_ref for
 * the biquad / Linkwitz-Riley parity check.  The definitions in ref_shim.cpp
 * abort(): the windowed-sinc path (the only FFTW user) is never exercised. */
#pragma once
#include <cstring>
#include <memory>
#include <stdexcept>
#include <cstddef>
extern "C" {
typedef float fftwf_complex[2];
typedef struct fftwf_plan_s *fftwf_plan;
#define FFTW_ESTIMATE (1U << 6)
float *fftwf_alloc_real(size_t n);
fftwf_complex *fftwf_alloc_complex(size_t n);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftwf_free(void *p);
}
