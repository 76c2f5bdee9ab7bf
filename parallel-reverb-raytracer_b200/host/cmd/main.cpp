// parallel_raytrace <config.json> <model.obj> <materials.json> <output.aif|aiff|wav>
//
// Drop-in for the reference's CLI (cmd/main.cpp:104-339): same arguments, same JSON schema
// (required: rays, reflections, sample_rate, bit_depth, source_position, mic_position,
// attenuation_model; optional: filter, hipass, normalize, volumme_scale [sic], trim_predelay,
// remove_direct, trim_tail, output_mode, verbose), same defaults, same exit codes.  The trace and
// everything after it run on the GPU through Raytracer::renderImpulseResponse.
//
// Extensions, all via the environment so the argument list stays the reference's:
//   RAYVERB_SEED=<n>             seed for getRandomDirections (default: wall clock, like the reference)
//   RAYVERB_DUMP_DIRECTIONS=<f>  write the ray directions used (raw float32 x4 per ray)
//   RAYVERB_DUMP_IR=<f>          write the float IR before PCM quantisation (raw float32, channel-major, header: nch, len)
//   RAYVERB_HRTF_TABLE=<f>       2*360*180*8 float32 HRTF table replacing the synthetic built-in
#include "rayverb.h"
#include "helpers.h"
#include "config.h"
#include "audio_writer.h"

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>

#include <sys/stat.h>

using namespace std;

static bool file_is_readable (const string & i)
{
    struct stat buffer;
    return stat (i.c_str(), &buffer) == 0;
}

static bool file_is_writable (const string & i)
{
    ofstream os (i);
    const bool ret = os.is_open() && (os << "test string" << endl);
    os.close();
    remove (i.c_str());
    return ret;
}

int main (int argc, const char * argv [])
{
    if (argc - 1 != 4)
    {
        cerr << "Command-line parameters are <config file (.json)> <model file> <material file (.json)> <output file (.aif)>" << endl;
        return 1;
    }
    const string config_filename (argv [1]), model_filename (argv [2]), material_filename (argv [3]), output_filename (argv [4]);

    for (const auto & i : {config_filename, model_filename, material_filename})
        if (! file_is_readable (i))
        {
            cerr << "input file " << i << " does not exist" << endl;
            return 1;
        }
    if (! file_is_writable (output_filename))
    {
        cerr << "output file " << output_filename << " cannot be written" << endl;
        return 1;
    }

    //  required params (defaults as cmd/main.cpp:140-146)
    cl_float3 source = {{0, 0, 0, 0}};
    cl_float3 mic = {{0, 0, 1, 0}};
    int numRays = 1024 * 8;
    int numImpulses = 64;
    double sampleRate = 44100.0;
    int bitDepth = 16;

    //  optional params (cmd/main.cpp:148-157)
    auto filter = RayverbFiltering::FILTER_TYPE_BIQUAD_ONEPASS;
    double hipass = 45.0;
    bool normalize = true;
    double volumme_scale = 1.0;
    bool trim_predelay = false;
    bool remove_direct = false;
    bool trim_tail = true;
    OutputMode output_mode = ALL;
    bool show_diagnostics = false;
    AttenuationModel attenuationModel;

    rvbjson::Document document;
    attemptJsonParse (config_filename, document);
    if (document.HasParseError())
    {
        cerr << "Encountered error while parsing config file:" << endl;
        cerr << document.GetParseError() << endl;
        return 1;
    }
    if (! document.IsObject())
    {
        cerr << "Rayverb config must be stored in a JSON object" << endl;
        return 1;
    }

    ConfigValidator cv;
    cv.addRequiredValidator ("rays", numRays);
    cv.addRequiredValidator ("reflections", numImpulses);
    cv.addRequiredValidator ("sample_rate", sampleRate);
    cv.addRequiredValidator ("bit_depth", bitDepth);
    cv.addRequiredValidator ("source_position", source);
    cv.addRequiredValidator ("mic_position", mic);
    cv.addRequiredValidator ("attenuation_model", attenuationModel);
    cv.addOptionalValidator ("filter", filter);
    cv.addOptionalValidator ("hipass", hipass);
    cv.addOptionalValidator ("normalize", normalize);
    cv.addOptionalValidator ("volumme_scale", volumme_scale);   // the reference reads this spelling (cmd/main.cpp:191)
    cv.addOptionalValidator ("trim_predelay", trim_predelay);
    cv.addOptionalValidator ("remove_direct", remove_direct);
    cv.addOptionalValidator ("trim_tail", trim_tail);
    cv.addOptionalValidator ("output_mode", output_mode);
    cv.addOptionalValidator ("verbose", show_diagnostics);
    try
    {
        cv.run (document);
    }
    catch (const runtime_error & error)
    {
        cerr << "encountered error reading config file:" << endl;
        cerr << error.what() << endl;
        return 1;
    }

    if (bitDepth != 16 && bitDepth != 24)
    {
        cerr << "Invalid bitdepth - valid bitdepths are: 16 24 " << endl;
        return 1;
    }
    const map <string, AudioFileType> ftypeTable {{"aif", AUDIO_AIFF}, {"aiff", AUDIO_AIFF}, {"wav", AUDIO_WAV}};
    const auto extension = output_filename.substr (output_filename.find_last_of (".") + 1);
    const auto ftypeIt = ftypeTable.find (extension);
    if (ftypeIt == ftypeTable.end())
    {
        cerr << "Invalid output file extension - valid extensions are: aif aiff wav " << endl;
        return 1;
    }

    const char * seed_env = getenv ("RAYVERB_SEED");
    const auto directions = seed_env ? getRandomDirections (numRays, strtoul (seed_env, nullptr, 10))
                                     : getRandomDirections (numRays);
    if (const char * dump = getenv ("RAYVERB_DUMP_DIRECTIONS"))
    {
        ofstream f (dump, ios::binary);
        f.write (reinterpret_cast <const char *> (directions.data()), (streamsize) (directions.size() * sizeof (cl_float3)));
    }

    vector <vector <float>> processed;
    try
    {
        Raytracer raytracer (numImpulses, model_filename, material_filename, show_diagnostics);
        raytracer.raytrace (mic, source, directions, show_diagnostics);
        // cmd/main.cpp:271-278 of the reference (compiled in with -DDIAGNOSTIC there)
        bool dump_rays = getenv ("RAYVERB_DIAGNOSTIC") && string (getenv ("RAYVERB_DIAGNOSTIC")) == "1";
#ifdef DIAGNOSTIC
        dump_rays = true;
#endif
        if (dump_rays) print_diagnostic (numRays, numImpulses, raytracer.getRawDiffuse().impulses, "impulse.dump");
        processed = raytracer.renderImpulseResponse
        (   attenuationModel
        ,   output_mode
        ,   remove_direct
        ,   trim_predelay
        ,   filter
        ,   (float) sampleRate
        ,   normalize
        ,   (float) hipass
        ,   trim_tail
        ,   (float) volumme_scale
        );
    }
    catch (const runtime_error & error)
    {
        cerr << "encountered runtime error:" << endl;
        cerr << error.what() << endl;
        return 1;
    }
    catch (...)
    {
        cerr << "encountered unknown runtime error" << endl;
        return 1;
    }

    if (processed.empty())
    {
        cerr << "No raytrace results returned." << endl;
        return 1;
    }
    if (const char * dump = getenv ("RAYVERB_DUMP_IR"))
    {
        ofstream f (dump, ios::binary);
        const uint64_t hdr [2] = {processed.size(), processed [0].size()};
        f.write (reinterpret_cast <const char *> (hdr), sizeof hdr);
        for (const auto & c : processed) f.write (reinterpret_cast <const char *> (c.data()), (streamsize) (c.size() * sizeof (float)));
    }
    try
    {
        write_audio_file (output_filename, processed, sampleRate, (unsigned) bitDepth, ftypeIt->second);
    }
    catch (const runtime_error & error)
    {
        cerr << error.what() << endl;
        return 1;
    }
    return 0;
}
