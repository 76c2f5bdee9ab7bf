// Host classes of the rayverb API over the C ABI (include/rayverb_b200.h).
// Mirrors the behaviour of the reference's rayverb/rayverb.cpp; citations are to that file.
#include "rayverb.h"
#include "helpers.h"

#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>

using namespace std;

namespace
{
    /// C status -> exception (the reference throws cl::Error / runtime_error; both end in exit (1)
    /// in cmd/main.cpp:299-316).
    void check (int rc)
    {
        if (rc != RVB_OK)
            throw runtime_error (string ("rayverb_b200: ") + rvb_last_error());
    }

    pair <cl_float3, cl_float3> getBounds (const vector <cl_float3> & vertices)   // rayverb.cpp:195-227
    {
        cl_float3 lo = vertices.front(), hi = vertices.front();
        for (const auto & v : vertices)
            for (int i = 0; i != 4; ++i)
            {
                lo.s [i] = min (lo.s [i], v.s [i]);
                hi.s [i] = max (hi.s [i], v.s [i]);
            }
        return make_pair (lo, hi);
    }

    bool inside (const pair <cl_float3, cl_float3> & bounds, const cl_float3 & point)   // rayverb.cpp:230-239
    {
        for (int i = 0; i != 4; ++i)
            if (! (bounds.first.s [i] <= point.s [i] && point.s [i] <= bounds.second.s [i]))
                return false;
        return true;
    }

    void printVec (const char * label, const cl_float3 & v)
    {
        cerr << label << "[" << v.s [0] << ", " << v.s [1] << ", " << v.s [2] << "]" << endl;
    }
}

// ---- free functions ------------------------------------------------------------------------------

vector <vector <float>> flattenImpulses (const vector <AttenuatedImpulse> & impulse, float samplerate)
{
    float * data = nullptr;
    size_t length = 0;
    check (rvb_flatten (impulse.data(), impulse.size(), samplerate, &data, &length));
    vector <vector <float>> flattened (RVB_NBANDS);
    for (int b = 0; b != RVB_NBANDS; ++b)
        flattened [b].assign (data + b * length, data + (b + 1) * length);
    rvb_free (data);
    return flattened;
}

vector <vector <vector <float>>> flattenImpulses (const vector <vector <AttenuatedImpulse>> & attenuated, float samplerate)
{
    vector <vector <vector <float>>> flattened (attenuated.size());
    transform (attenuated.begin(), attenuated.end(), flattened.begin(),
               [samplerate] (const vector <AttenuatedImpulse> & i) { return flattenImpulses (i, samplerate); });
    return flattened;
}

namespace
{
    /// [band][sample] vectors of one channel -> contiguous [8][len]
    vector <float> packChannel (const vector <vector <float>> & channel, size_t & len)
    {
        if (channel.size() != RVB_NBANDS)
            throw runtime_error ("each channel must hold 8 bands");
        len = channel.front().size();
        vector <float> packed (RVB_NBANDS * len);
        for (int b = 0; b != RVB_NBANDS; ++b)
        {
            if (channel [b].size() != len)
                throw runtime_error ("bands of one channel must have the same length");
            copy (channel [b].begin(), channel [b].end(), packed.begin() + b * len);
        }
        return packed;
    }
    void unpackChannel (const vector <float> & packed, size_t len, vector <vector <float>> & channel)
    {
        for (int b = 0; b != RVB_NBANDS; ++b)
            channel [b].assign (packed.begin() + b * len, packed.begin() + (b + 1) * len);
    }
}

void RayverbFiltering::filter (FilterType ft, vector <vector <vector <float>>> & data, float sr, float lo_cutoff)
{
    if (ft == FILTER_TYPE_WINDOWED_SINC)
    {
        //  The convolver is sized once by the first channel's first band (filters.cpp:280-283):
        //  every band of every channel comes back with 29 + that length - 1 samples.
        if (data.empty()) return;
        const size_t len = data.front().front().size();
        vector <float> packed (data.size() * RVB_NBANDS * len, 0.0f);
        for (size_t c = 0; c != data.size(); ++c)
            for (int b = 0; b != RVB_NBANDS; ++b)
            {
                if (data [c] [b].size() > len + RVB_SINC_KERNEL_LENGTH - 1)
                    throw runtime_error ("windowed-sinc: a channel longer than the first overflows the convolution buffer");
                if (data [c] [b].size() > len)   // fits the FFT buffer of the reference, not this packed layout
                    throw runtime_error ("windowed-sinc: channels longer than the first are not supported");
                copy (data [c] [b].begin(), data [c] [b].end(), packed.begin() + (c * RVB_NBANDS + b) * len);
            }
        const size_t grown = len + RVB_SINC_KERNEL_LENGTH - 1;
        vector <float> out (data.size() * RVB_NBANDS * grown);
        check (rvb_filter_sinc (packed.data(), data.size(), len, sr, lo_cutoff, out.data()));
        for (size_t c = 0; c != data.size(); ++c)
            for (int b = 0; b != RVB_NBANDS; ++b)
                data [c] [b].assign (out.begin() + (c * RVB_NBANDS + b) * grown, out.begin() + (c * RVB_NBANDS + b + 1) * grown);
        return;
    }
    for (auto && channel : data)
    {
        size_t len = 0;
        vector <float> packed = packChannel (channel, len);
        check (rvb_filter ((int) ft, packed.data(), 1, len, sr, lo_cutoff));
        unpackChannel (packed, len, channel);
    }
}

vector <vector <float>> process
(   RayverbFiltering::FilterType filtertype
,   vector <vector <vector <float>>> & data
,   float sr
,   bool do_normalize
,   float lo_cutoff
,   bool do_trim_tail
,   float volumme_scale
)
{
    const size_t nch = data.size();
    if (filtertype == RayverbFiltering::FILTER_TYPE_WINDOWED_SINC && nch != 0)
    {
        //  The rows grow to 29 + (first length) - 1: the device computes the processed channels from
        //  the unfiltered data, then `data` itself is filtered (and grown) as the reference leaves it.
        vector <vector <float>> packed (nch);
        vector <size_t> lengths (nch);
        vector <float *> in (nch), out (nch);
        for (size_t c = 0; c != nch; ++c) packed [c] = packChannel (data [c], lengths [c]);
        const size_t grown = lengths [0] + RVB_SINC_KERNEL_LENGTH - 1;
        vector <vector <float>> ret (nch, vector <float> (grown, 0.0f));
        for (size_t c = 0; c != nch; ++c) { in [c] = packed [c].data(); out [c] = ret [c].data(); }
        int64_t trimmed = -1;
        check (rvb_process ((int) filtertype, in.data(), lengths.data(), nch, sr, do_normalize, lo_cutoff, do_trim_tail,
                            volumme_scale, out.data(), &trimmed));
        if (trimmed >= 0)
            for (auto && ch : ret) ch.resize ((size_t) trimmed);
        RayverbFiltering::filter (filtertype, data, sr, lo_cutoff);
        return ret;
    }
    vector <vector <float>> packed (nch);
    vector <size_t> lengths (nch);
    vector <float *> in (nch), out (nch);
    vector <vector <float>> ret (nch);
    for (size_t c = 0; c != nch; ++c)
    {
        packed [c] = packChannel (data [c], lengths [c]);
        ret [c].assign (lengths [c], 0.0f);
        in [c] = packed [c].data();
        out [c] = ret [c].data();
    }
    int64_t trimmed = -1;
    check (rvb_process ((int) filtertype, in.data(), lengths.data(), nch, sr, do_normalize, lo_cutoff, do_trim_tail,
                        volumme_scale, out.data(), &trimmed));
    for (size_t c = 0; c != nch; ++c)
    {
        unpackChannel (packed [c], lengths [c], data [c]);   // data is filtered in place (rayverb.cpp:135)
        if (trimmed >= 0) ret [c].resize ((size_t) trimmed);  // trimTail resizes every channel (rayverb.cpp:120-121)
    }
    return ret;
}

void attemptJsonParse (const string & fname, rvbjson::Document & doc)
{
    ifstream in (fname);
    string file ((istreambuf_iterator <char> (in)), istreambuf_iterator <char>());
    doc.Parse (file);
}

// ---- helpers (helpers.cpp:63-81) -------------------------------------------------------------------

cl_float3 spherePoint (float z, float theta)
{
    const float ztemp = sqrtf (1 - z * z);
    return cl_float3 {{ztemp * cosf (theta), ztemp * sinf (theta), z, 0}};
}

vector <cl_float3> getRandomDirections (unsigned long num, unsigned long seed)
{
    vector <cl_float3> ret (num);
    uniform_real_distribution <float> zDist (-1, 1);
    uniform_real_distribution <float> thetaDist (-M_PI, M_PI);
    default_random_engine engine (seed);
    for (auto && i : ret)
    {
        const float z = zDist (engine);          // z first, then theta: a fixed order (the reference
        const float theta = thetaDist (engine);  // leaves the evaluation order unspecified)
        i = spherePoint (z, theta);
    }
    return ret;
}

vector <cl_float3> getRandomDirections (unsigned long num)
{
    return getRandomDirections (num, (unsigned long) chrono::system_clock::now().time_since_epoch().count());
}

// ---- diagnostic ray dump (helpers.cpp:19-59) ------------------------------------------------------------

// rapidjson's Writer::Double: shortest decimal that reads back as the same double, always with a
// fraction or an exponent
static string json_double (double v)
{
    char buf [64];
    const auto r = to_chars (buf, buf + sizeof buf, v);
    string out (buf, r.ptr);
    if (out.find_first_of (".eE") == string::npos && isfinite (v)) out += ".0";
    return out;
}

void print_diagnostic (unsigned long nrays, unsigned long nreflections, const vector <Impulse> & impulses, const string & fname)
{
    if (impulses.size() < (size_t) nrays * nreflections) throw runtime_error ("print_diagnostic: fewer impulses than rays x reflections");
    ofstream out (fname);
    for (unsigned long i = 0; i != nrays; ++i)
    {
        string line = "[";
        for (unsigned long j = 0; j != nreflections; ++j)
        {
            const Impulse & reflection = impulses [i * nreflections + j];
            float average = 0;
            for (int k = 0; k != 8; ++k) average += reflection.volume.s [k];
            average /= 8;
            line += j ? ",{" : "{";
            line += "\"position\":[" + json_double (reflection.position.s [0]) + "," + json_double (reflection.position.s [1]) + ","
                    + json_double (reflection.position.s [2]) + "],\"volume\":" + json_double (average) + "}";
        }
        out << line << "]" << endl;
    }
}

// ---- Raytracer -------------------------------------------------------------------------------------

struct Raytracer::SceneData
{
    SceneData (const string & objpath, const string & materialFileName, bool verbose)
    {
        loadScene (objpath, materialFileName, verbose, triangles, vertices, surfaces);
    }
    vector <Triangle> triangles;
    vector <cl_float3> vertices;
    vector <Surface> surfaces;
};

void Raytracer::init (vector <Triangle> & triangles, vector <cl_float3> & vertices, vector <Surface> & surfaces, bool)
{
    if (triangles.empty() || vertices.empty() || surfaces.empty())
        throw runtime_error ("Raytracer needs at least one triangle, vertex and surface");
    ntriangles = triangles.size();
    bounds = getBounds (vertices);
    check (rvb_scene_create (triangles.data(), triangles.size(), vertices.data(), vertices.size(), surfaces.data(),
                             surfaces.size(), RVB_ACCEL_AUTO, &scene));
}

Raytracer::Raytracer
(   unsigned long nreflections
,   vector <Triangle> & triangles
,   vector <cl_float3> & vertices
,   vector <Surface> & surfaces
,   bool verbose
)
:   nreflections (nreflections), ntriangles (0), storedMicpos {{0, 0, 0, 0}}, scene (nullptr), trace (nullptr)
{
    init (triangles, vertices, surfaces, verbose);
}

Raytracer::Raytracer (unsigned long nreflections, const string & objpath, const string & materialFileName, bool verbose)
:   Raytracer (nreflections, SceneData (objpath, materialFileName, verbose), verbose)
{
}

Raytracer::Raytracer (unsigned long nreflections, SceneData sceneData, bool verbose)
:   nreflections (nreflections), ntriangles (0), storedMicpos {{0, 0, 0, 0}}, scene (nullptr), trace (nullptr)
{
    init (sceneData.triangles, sceneData.vertices, sceneData.surfaces, verbose);
}

Raytracer::~Raytracer()
{
    rvb_trace_destroy (trace);
    rvb_scene_destroy (scene);
}

void Raytracer::raytrace (const cl_float3 & micpos, const cl_float3 & source, const vector <cl_float3> & directions, bool verbose)
{
    storedMicpos = micpos;

    //  warn when mic or source fall outside the model bounds (rayverb.cpp:547-583)
    const bool micinside = inside (bounds, micpos);
    const bool srcinside = inside (bounds, source);
    if (verbose && (! (micinside && srcinside)))
    {
        cerr << "model bounds: [" << bounds.first.s [0] << ", " << bounds.first.s [1] << ", " << bounds.first.s [2]
             << "], [" << bounds.second.s [0] << ", " << bounds.second.s [1] << ", " << bounds.second.s [2] << "]" << endl;
        if (! micinside)
        {
            cerr << "WARNING: microphone position may be outside model" << endl;
            printVec ("mic position: ", micpos);
        }
        if (! srcinside)
        {
            cerr << "WARNING: source position may be outside model" << endl;
            printVec ("src position: ", source);
        }
    }

    rvb_trace_destroy (trace);
    trace = nullptr;
    //  One launch over all rays; the reference's groups of 4096 (rayverb.cpp:587-684) only bounded
    //  its per-group buffers.
    check (rvb_trace_run (scene, micpos.s, source.s, directions.data(), directions.size(), nreflections, 0, 0, &trace));
}

void Raytracer::requireTrace() const
{
    if (! trace)
        throw runtime_error ("raytrace() has not been run");
}

RaytracerResults Raytracer::getRawDiffuse()
{
    requireTrace();
    size_t n = 0;
    check (rvb_trace_diffuse_count (trace, &n));
    vector <Impulse> ret (n);
    check (rvb_trace_copy_diffuse (trace, ret.data()));
    return RaytracerResults (ret, storedMicpos);
}

RaytracerResults Raytracer::getRawImages (bool removeDirect)
{
    requireTrace();
    size_t n = 0;
    check (rvb_trace_image_count (trace, removeDirect, &n));
    vector <Impulse> ret (n);
    check (rvb_trace_copy_images (trace, removeDirect, ret.data(), nullptr, nullptr));
    return RaytracerResults (ret, storedMicpos);
}

RaytracerResults Raytracer::getAllRaw (bool removeDirect)
{
    auto diffuse = getRawDiffuse().impulses;
    const auto image = getRawImages (removeDirect).impulses;
    diffuse.insert (diffuse.end(), image.begin(), image.end());
    return RaytracerResults (diffuse, storedMicpos);
}

vector <vector <float>> Raytracer::renderImpulseResponse
(   const AttenuationModel & model
,   OutputMode output_mode
,   bool remove_direct
,   bool trim_predelay
,   RayverbFiltering::FilterType filtertype
,   float sr
,   bool do_normalize
,   float lo_cutoff
,   bool do_trim_tail
,   float volumme_scale
,   const HrtfTable * hrtf_table
)
{
    requireTrace();
    rvb_render_config cfg;
    memset (&cfg, 0, sizeof cfg);
    if (model.mode == AttenuationModel::SPEAKER)
    {
        cfg.model = RVB_MODEL_SPEAKER;
        cfg.speakers = model.speakers.data();
        cfg.nspeakers = model.speakers.size();
    }
    else
    {
        const HrtfTable & table = hrtf_table ? *hrtf_table : HrtfAttenuator::builtinTable();
        cfg.model = RVB_MODEL_HRTF;
        cfg.hrtf_table = &table [0] [0] [0];
        for (int i = 0; i != 3; ++i) { cfg.facing [i] = model.hrtf.facing.s [i]; cfg.up [i] = model.hrtf.up.s [i]; }
    }
    cfg.sample_rate = sr;
    cfg.filter_type = (int) filtertype;
    cfg.lo_cutoff = lo_cutoff;
    cfg.do_normalize = do_normalize;
    cfg.volume_scale = volumme_scale;
    cfg.do_trim_tail = do_trim_tail;
    cfg.trim_predelay = trim_predelay;
    cfg.remove_direct = remove_direct;
    cfg.output_mode = (int) output_mode;
    rvb_ir * ir = nullptr;
    check (rvb_render_ir (trace, &cfg, &ir));
    size_t nch = 0;
    rvb_ir_channels (ir, &nch);
    vector <vector <float>> ret (nch);
    for (size_t c = 0; c != nch; ++c)
    {
        size_t len = 0;
        rvb_ir_length (ir, c, &len);
        ret [c].resize (len);
        rvb_ir_copy (ir, c, ret [c].data());
    }
    rvb_ir_destroy (ir);
    return ret;
}

// ---- attenuators -----------------------------------------------------------------------------------

SpeakerAttenuator::SpeakerAttenuator() {}

vector <vector <AttenuatedImpulse>> SpeakerAttenuator::attenuate (const RaytracerResults & results, const vector <Speaker> & speakers)
{
    vector <vector <AttenuatedImpulse>> attenuated (speakers.size());
    transform (speakers.begin(), speakers.end(), attenuated.begin(),
               [this, &results] (const Speaker & i) { return attenuate (results.mic, i, results.impulses); });
    return attenuated;
}

vector <AttenuatedImpulse> SpeakerAttenuator::attenuate (const cl_float3 & mic_pos, const Speaker & speaker, const vector <Impulse> & impulses)
{
    vector <AttenuatedImpulse> ret (impulses.size());
    check (rvb_attenuate_speaker (impulses.data(), impulses.size(), mic_pos.s, &speaker, ret.data()));
    return ret;
}

HrtfAttenuator::HrtfAttenuator() {}

vector <vector <AttenuatedImpulse>> HrtfAttenuator::attenuate (const RaytracerResults & results, const HrtfConfig & config)
{
    return attenuate (results, config.facing, config.up);
}

vector <vector <AttenuatedImpulse>> HrtfAttenuator::attenuate (const RaytracerResults & results, const cl_float3 & facing, const cl_float3 & up)
{
    vector <vector <AttenuatedImpulse>> attenuated (2);
    for (unsigned long channel = 0; channel != 2; ++channel)
        attenuated [channel] = attenuate (results.mic, channel, facing, up, results.impulses);
    return attenuated;
}

vector <AttenuatedImpulse> HrtfAttenuator::attenuate
(   const cl_float3 & mic_pos
,   unsigned long channel
,   const cl_float3 & facing
,   const cl_float3 & up
,   const vector <Impulse> & impulses
)
{
    //  [azimuth][elevation] of one ear is already contiguous: a * 180 + e (rayverb.cpp:774-780)
    const HrtfTable & table = getHrtfData();
    vector <AttenuatedImpulse> ret (impulses.size());
    check (rvb_attenuate_hrtf (impulses.data(), impulses.size(), mic_pos.s, &table [channel] [0] [0], facing.s, up.s, channel, ret.data()));
    return ret;
}

namespace
{
    unique_ptr <HrtfTable> g_table;

    void fillSynthetic (HrtfTable & t)
    {
        for (int ear = 0; ear != 2; ++ear)
            for (int a = 0; a != 360; ++a)
                for (int e = 0; e != 180; ++e)
                    for (int b = 0; b != 8; ++b)
                    {
                        const double az = a * M_PI / 180.0, el = e * M_PI / 180.0, side = ear == 0 ? -1.0 : 1.0;
                        t [ear] [a] [e].s [b] = (float) (0.55 + 0.35 * side * sin (az) * sin (el) * (0.3 + 0.1 * b)
                                                         + 0.08 * cos (2 * az + 0.3 * b) * sin (el));
                    }
    }

    void readTable (const string & path, HrtfTable & t)
    {
        ifstream in (path, ios::binary);
        if (! in)
            throw runtime_error ("cannot open HRTF table " + path);
        in.read (reinterpret_cast <char *> (&t [0] [0] [0]), sizeof (HrtfTable));
        if ((size_t) in.gcount() != sizeof (HrtfTable))
            throw runtime_error ("HRTF table " + path + " must hold 2*360*180*8 float32 values");
    }
}

const HrtfTable & HrtfAttenuator::builtinTable()
{
    if (! g_table)
    {
        g_table.reset (new HrtfTable);
        const char * env = getenv ("RAYVERB_HRTF_TABLE");
        if (env && *env) readTable (env, *g_table);
        else fillSynthetic (*g_table);
    }
    return *g_table;
}

void HrtfAttenuator::loadTable (const string & path)
{
    unique_ptr <HrtfTable> t (new HrtfTable);
    readTable (path, *t);
    g_table = move (t);
}

const HrtfTable & HrtfAttenuator::getHrtfData() const
{
    return builtinTable();
}
