// Multiband crossover filtering: the API of rayverb/filters.h:196-210.  The work runs on the GPU
// (rvb_filter: chunk-parallel biquad scans, csrc/render.cu).
#pragma once
#include <vector>

namespace RayverbFiltering
{
    /// Enum denoting available filter types (same order as the reference).
    enum FilterType
    {   FILTER_TYPE_WINDOWED_SINC
    ,   FILTER_TYPE_BIQUAD_ONEPASS
    ,   FILTER_TYPE_BIQUAD_TWOPASS
    ,   FILTER_TYPE_LINKWITZ_RILEY
    };

    /// data[channel][band][sample] is band-pass filtered in place with the band edges
    /// {lo_cutoff, 175, 350, 700, 1400, 2800, 5600, 11200, 20000} (filters.cpp:268-306).
    /// FILTER_TYPE_WINDOWED_SINC grows every band to 29 + (length of the first band) - 1 samples,
    /// scaled by that length like the reference's un-normalised FFT round trip.
    void filter
    (   FilterType ft
    ,   std::vector <std::vector <std::vector <float>>> & data
    ,   float sr
    ,   float lo_cutoff
    );
}
