// rayverb host API on the B200 path: the classes and free functions of the reference's
// rayverb/rayverb.h (Raytracer / SpeakerAttenuator / HrtfAttenuator / flattenImpulses / process /
// findPredelay / fixPredelay) with the same names, argument meaning and error behaviour
// (std::runtime_error where the reference throws std::runtime_error or cl::Error).  Each object
// holds an opaque handle and forwards to the C ABI (include/rayverb_b200.h); there is no OpenCL
// context to create and no kernel to JIT (ContextProvider / KernelLoader of the reference,
// rayverb.h:100-118, have no counterpart).
#pragma once

#include "filters.h"
#include "clstructs.h"
#include "generic_functions.h"
#include "json.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

/// Sum impulses occurring at the same (sampled) time: ret[band][sample] (rayverb.cpp:48-77).
std::vector <std::vector <float>> flattenImpulses
(   const std::vector <AttenuatedImpulse> & impulse
,   float samplerate
);

/// Maps flattenImpulses over channels.
std::vector <std::vector <std::vector <float>>> flattenImpulses
(   const std::vector <std::vector <AttenuatedImpulse>> & impulse
,   float samplerate
);

/// Filter and mix down each channel; optionally normalize, scale and trim the tail
/// (rayverb.cpp:125-149).  `data` is left filtered, as in the reference.
std::vector <std::vector <float>> process
(   RayverbFiltering::FilterType filtertype
,   std::vector <std::vector <std::vector <float>>> & data
,   float sr
,   bool do_normalize
,   float lo_cutoff
,   bool do_trim_tail
,   float volumme_scale
);

/// Earliest non-zero time in a (nested) collection of AttenuatedImpulses (rayverb.h:47-76).
inline float findPredelay (const AttenuatedImpulse & i) { return i.time; }
template <typename T>
inline float findPredelay (const std::vector <T> & ret)
{
    float acc = findPredelay (ret.front());
    for (auto i = ret.begin() + 1; i != ret.end(); ++i)
    {
        const float pd = findPredelay (*i);
        if (acc == 0) acc = pd;
        else if (pd != 0) acc = std::min (acc, pd);
    }
    return acc;
}

/// Subtract `seconds` from every time, clamping at zero (rayverb.h:78-89).
inline void fixPredelay (AttenuatedImpulse & ret, float seconds)
{
    ret.time = ret.time > seconds ? ret.time - seconds : 0;
}
template <typename T>
inline void fixPredelay (std::vector <T> & ret, float seconds)
{
    for (auto && i : ret) fixPredelay (i, seconds);
}
/// Find and remove the predelay.
template <typename T>
inline void fixPredelay (std::vector <T> & ret)
{
    fixPredelay (ret, findPredelay (ret));
}

/// Impulses together with the mic position they were traced for.
struct RaytracerResults
{
    RaytracerResults() {}
    RaytracerResults (const std::vector <Impulse> impulses, const cl_float3 & c)
    :   impulses (impulses)
    ,   mic (c)
    {}

    std::vector <Impulse> impulses;
    cl_float3 mic;
};

/// HRTF parameters.
struct HrtfConfig
{
    cl_float3 facing;
    cl_float3 up;
};

/// Different components of the output impulse (config.h:19-23).
enum OutputMode
{   ALL
,   IMAGE_ONLY
,   DIFFUSE_ONLY
};

/// Which attenuation to apply (config.h:29-39).
struct AttenuationModel
{
    enum Mode
    {   SPEAKER
    ,   HRTF
    };
    Mode mode;
    HrtfConfig hrtf;
    std::vector <Speaker> speakers;
};

typedef std::array <std::array <std::array <cl_float8, 180>, 360>, 2> HrtfTable;

/// The raytracer: scene staged once on the GPU, results kept on the GPU until asked for.
class Raytracer
{
public:
    /// Use your own geometry (rayverb.h:142-148).
    Raytracer
    (   unsigned long nreflections
    ,   std::vector <Triangle> & triangles
    ,   std::vector <cl_float3> & vertices
    ,   std::vector <Surface> & surfaces
    ,   bool verbose
    );

    /// Load a 3d model and materials from files (rayverb.h:151-156).
    Raytracer
    (   unsigned long nreflections
    ,   const std::string & objpath
    ,   const std::string & materialFileName
    ,   bool verbose
    );

    virtual ~Raytracer();
    Raytracer (const Raytracer &) = delete;
    Raytracer & operator= (const Raytracer &) = delete;

    /// Run the raytrace with a specific mic, source, and set of directions (rayverb.h:159-164).
    void raytrace
    (   const cl_float3 & micpos
    ,   const cl_float3 & source
    ,   const std::vector <cl_float3> & directions
    ,   bool verbose
    );

    /// Raw diffuse results: nrays * nreflections impulses, ray-major.
    RaytracerResults getRawDiffuse();
    /// De-duplicated image-source results in key order.
    RaytracerResults getRawImages (bool removeDirect);
    /// Diffuse followed by image-source results.
    RaytracerResults getAllRaw (bool removeDirect);

    /// Everything cmd/main.cpp does between raytrace() and write_sndfile(), on the GPU without
    /// copying the impulses to the host: attenuate -> fixPredelay -> flattenImpulses -> process.
    std::vector <std::vector <float>> renderImpulseResponse
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
    ,   const HrtfTable * hrtf_table = nullptr
    );

    unsigned long getNumTriangles() const { return ntriangles; }

private:
    struct SceneData;
    Raytracer (unsigned long nreflections, SceneData sceneData, bool verbose);
    void init
    (   std::vector <Triangle> & triangles
    ,   std::vector <cl_float3> & vertices
    ,   std::vector <Surface> & surfaces
    ,   bool verbose
    );
    void requireTrace() const;

    const unsigned long nreflections;
    unsigned long ntriangles;
    std::pair <cl_float3, cl_float3> bounds;
    cl_float3 storedMicpos;
    rvb_scene * scene;
    rvb_trace * trace;
};

/// Common base kept for source compatibility (rayverb.h:230-234).
struct Attenuator
{
    virtual ~Attenuator() {}
};

/// HRTF attenuation of raytrace results: [channel][impulse].
class HrtfAttenuator: public Attenuator
{
public:
    HrtfAttenuator();

    std::vector <std::vector <AttenuatedImpulse>> attenuate
    (   const RaytracerResults & results
    ,   const HrtfConfig & config
    );
    std::vector <std::vector <AttenuatedImpulse>> attenuate
    (   const RaytracerResults & results
    ,   const cl_float3 & facing
    ,   const cl_float3 & up
    );

    /// [channel][azimuth][elevation]; override to inject a table (the reference's tests do).
    virtual const HrtfTable & getHrtfData() const;

    /// Replace the built-in table with 2*360*180*8 float32 values read from `path`.
    static void loadTable (const std::string & path);
    /// The built-in table.  The reference's table (rayverb/hrtf.cpp, generated from the IRCAM
    /// Listen set) is not redistributable here; this is a smooth synthetic stand-in unless
    /// loadTable() or the RAYVERB_HRTF_TABLE environment variable supplies real data.
    static const HrtfTable & builtinTable();
private:
    std::vector <AttenuatedImpulse> attenuate
    (   const cl_float3 & mic_pos
    ,   unsigned long channel
    ,   const cl_float3 & facing
    ,   const cl_float3 & up
    ,   const std::vector <Impulse> & impulses
    );
};

/// Speaker attenuation of raytrace results: [channel][impulse].
class SpeakerAttenuator: public Attenuator
{
public:
    SpeakerAttenuator();

    std::vector <std::vector <AttenuatedImpulse>> attenuate
    (   const RaytracerResults & results
    ,   const std::vector <Speaker> & speakers
    );
private:
    std::vector <AttenuatedImpulse> attenuate
    (   const cl_float3 & mic_pos
    ,   const Speaker & speaker
    ,   const std::vector <Impulse> & impulses
    );
};

/// Try to open and parse a json file; like the reference, a missing or malformed file leaves a
/// document whose HasParseError() / IsObject() the caller checks.
void attemptJsonParse
(   const std::string & fname
,   rvbjson::Document & doc
);

/// Scene loading used by Raytracer (objpath, materialFileName): Wavefront OBJ + JSON materials
/// with the observable behaviour of the reference's Assimp-based SceneData (rayverb.cpp:296-461).
void loadScene
(   const std::string & objpath
,   const std::string & materialFileName
,   bool verbose
,   std::vector <Triangle> & triangles
,   std::vector <cl_float3> & vertices
,   std::vector <Surface> & surfaces
);
