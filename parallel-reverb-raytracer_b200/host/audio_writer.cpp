#include "audio_writer.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <stdexcept>

using namespace std;

namespace
{
    void put_be (vector <uint8_t> & b, uint64_t v, int bytes) { for (int i = bytes - 1; i >= 0; --i) b.push_back ((uint8_t) (v >> (8 * i))); }
    void put_le (vector <uint8_t> & b, uint64_t v, int bytes) { for (int i = 0; i != bytes; ++i) b.push_back ((uint8_t) (v >> (8 * i))); }
    void put_tag (vector <uint8_t> & b, const char * t) { b.insert (b.end(), t, t + 4); }

    /// IEEE 754 80-bit extended, big endian (AIFF sample rate field)
    void put_extended (vector <uint8_t> & b, double v)
    {
        int e = 0;
        const double m = frexp (v, &e);               // v = m * 2^e, 0.5 <= m < 1
        const uint64_t mant = (uint64_t) ldexp (m, 64);
        put_be (b, (uint16_t) (e - 1 + 16383), 2);
        put_be (b, mant, 8);
    }

    /// float in [-1, 1] -> signed PCM (scale by 2^(bits-1) - 1, round to nearest, clip)
    int32_t quantise (float x, unsigned bits)
    {
        const double full = (double) ((1u << (bits - 1)) - 1);
        const double v = nearbyint ((double) x * full);
        return (int32_t) max (-full - 1, min (full, std::isfinite (v) ? v : 0.0));
    }
}

void write_audio_file (const string & fname, const vector <vector <float>> & outdata, double sr, unsigned bit_depth, AudioFileType ftype)
{
    if (bit_depth != 16 && bit_depth != 24)
        throw runtime_error ("Invalid bitdepth - valid bitdepths are: 16 24");
    const size_t nch = outdata.size();
    size_t frames = 0;
    for (const auto & c : outdata) frames = max (frames, c.size());
    const int bytes = (int) bit_depth / 8;
    vector <uint8_t> pcm;
    pcm.reserve (frames * nch * bytes);
    for (size_t j = 0; j != frames; ++j)
        for (size_t i = 0; i != nch; ++i)
        {
            const int32_t q = quantise (j < outdata [i].size() ? outdata [i] [j] : 0.0f, bit_depth);
            if (ftype == AUDIO_AIFF) put_be (pcm, (uint32_t) q & (bit_depth == 16 ? 0xffffu : 0xffffffu), bytes);
            else put_le (pcm, (uint32_t) q & (bit_depth == 16 ? 0xffffu : 0xffffffu), bytes);
        }
    vector <uint8_t> out;
    if (ftype == AUDIO_AIFF)
    {
        const uint32_t ssnd = 8 + (uint32_t) pcm.size(), pad = ssnd & 1;
        put_tag (out, "FORM"); put_be (out, 4 + (8 + 18) + (8 + ssnd + pad), 4); put_tag (out, "AIFF");
        put_tag (out, "COMM"); put_be (out, 18, 4); put_be (out, nch, 2); put_be (out, frames, 4); put_be (out, bit_depth, 2);
        put_extended (out, sr);
        put_tag (out, "SSND"); put_be (out, ssnd, 4); put_be (out, 0, 4); put_be (out, 0, 4);
        out.insert (out.end(), pcm.begin(), pcm.end());
        if (pad) out.push_back (0);
    }
    else
    {
        const uint32_t data = (uint32_t) pcm.size(), pad = data & 1;
        put_tag (out, "RIFF"); put_le (out, 4 + (8 + 16) + (8 + data + pad), 4); put_tag (out, "WAVE");
        put_tag (out, "fmt "); put_le (out, 16, 4); put_le (out, 1, 2); put_le (out, nch, 2); put_le (out, (uint32_t) sr, 4);
        put_le (out, (uint32_t) sr * nch * bytes, 4); put_le (out, nch * bytes, 2); put_le (out, bit_depth, 2);
        put_tag (out, "data"); put_le (out, data, 4);
        out.insert (out.end(), pcm.begin(), pcm.end());
        if (pad) out.push_back (0);
    }
    ofstream f (fname, ios::binary);
    if (! f || ! f.write (reinterpret_cast <const char *> (out.data()), (streamsize) out.size()))
        throw runtime_error ("output file " + fname + " cannot be written");
}
