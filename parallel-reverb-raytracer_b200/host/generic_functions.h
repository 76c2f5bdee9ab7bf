// Recursive helpers over nested vectors (the API of rayverb/generic_functions.h).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

inline float max_amp (float t) { return std::fabs (t); }
template <typename T>
inline float max_amp (const std::vector <T> & t)
{
    float m = 0;
    for (const auto & i : t) m = std::max (m, max_amp (i));
    return m;
}

inline void mul (float & ret, float f) { ret *= f; }
template <typename T>
inline void mul (std::vector <T> & ret, float f) { for (auto && i : ret) mul (i, f); }

inline void div (float & ret, float f) { ret /= f; }
template <typename T>
inline void div (std::vector <T> & ret, float f) { for (auto && i : ret) div (i, f); }

/// Scale so that the largest magnitude becomes 1 (multiplies by float (1.0 / max), like the reference).
template <typename T>
inline void normalize (std::vector <T> & ret) { mul (ret, (float) (1.0 / max_amp (ret))); }

template <typename T, typename U>
inline T elementwise (const T & a, const T & b, const U & u)
{
    T ret;
    std::transform (std::begin (a.s), std::end (a.s), std::begin (b.s), std::begin (ret.s), u);
    return ret;
}
