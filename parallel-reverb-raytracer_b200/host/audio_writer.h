// PCM writer for the impulse response (the reference uses libsndfile's SndfileHandle,
// cmd/main.cpp:26-48, which is not available in this image): AIFF and WAV, 16 or 24 bit,
// channels interleaved, every channel padded to the longest.
#pragma once
#include <string>
#include <vector>

enum AudioFileType { AUDIO_AIFF, AUDIO_WAV };

void write_audio_file
(   const std::string & fname
,   const std::vector <std::vector <float>> & outdata
,   double sr
,   unsigned bit_depth   // 16 or 24
,   AudioFileType ftype
);
