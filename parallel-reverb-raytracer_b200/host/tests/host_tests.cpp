// The reference's gtest suite (tests/raytrace_tests.h:30-48, tests/attenuation_tests.h:67-101,
// tests/hrtf_tests.cpp:42-85) re-expressed against the B200 host classes, without gtest (not in
// this image).  Like the reference's fixtures, the test classes INHERIT from the class under test.
//   host_tests <large_square.obj> <materials.json>
// Needs a GPU; exit code 0 = all passed.
#include "rayverb.h"
#include "helpers.h"

#include <cstdio>
#include <cstring>
#include <random>

using namespace std;

static int failures = 0, checks = 0;

static bool float_eq (float a, float b)   // gtest ASSERT_FLOAT_EQ: within 4 ULPs
{
    if (a == b) return true;
    int32_t ia, ib;
    memcpy (&ia, &a, 4); memcpy (&ib, &b, 4);
    if (ia < 0) ia = (int32_t) 0x80000000 - ia;
    if (ib < 0) ib = (int32_t) 0x80000000 - ib;
    return llabs ((long long) ia - ib) <= 4;
}
#define EXPECT_FLOAT_EQ(a, b) do { ++checks; if (! float_eq ((a), (b))) { ++failures; printf ("FAIL %s:%d  %s = %.9g, %s = %.9g\n", __FILE__, __LINE__, #a, (double) (a), #b, (double) (b)); } } while (0)
#define EXPECT_TRUE(x) do { ++checks; if (! (x)) { ++failures; printf ("FAIL %s:%d  %s\n", __FILE__, __LINE__, #x); } } while (0)

// ---- RaytracerTest.ImpulseDirections ---------------------------------------------------------------
struct RaytracerTest: public Raytracer
{
    static const int NUM_REFLECTIONS = 128;
    RaytracerTest (const string & obj, const string & mat): Raytracer (NUM_REFLECTIONS, obj, mat, false)
    {
        directions = {{{0, 0, -1, 0}}, {{0, 0, 1, 0}}, {{0, -1, 0, 0}}, {{0, 1, 0, 0}}, {{-1, 0, 0, 0}}, {{1, 0, 0, 0}}};
        directions.resize (64 * 1000, cl_float3 {{0, 0, -1, 0}});
    }
    void test_eq (const cl_float3 & a, float x, float y, float z)
    {
        EXPECT_FLOAT_EQ (a.s [0], x); EXPECT_FLOAT_EQ (a.s [1], y); EXPECT_FLOAT_EQ (a.s [2], z);
    }
    void run()
    {
        const cl_float3 mic_pos = {{0, 2, 0, 0}}, src_pos = {{0, 2, 2, 0}};
        raytrace (mic_pos, src_pos, directions, false);
        auto diffuse = getRawDiffuse().impulses;
        EXPECT_TRUE (diffuse.size() == directions.size() * NUM_REFLECTIONS);
        test_eq (diffuse [0 * NUM_REFLECTIONS + 0].position, 0, 2, -27);
        test_eq (diffuse [1 * NUM_REFLECTIONS + 0].position, 0, 2, 27);
        test_eq (diffuse [2 * NUM_REFLECTIONS + 0].position, 0, 0, 2);
        test_eq (diffuse [3 * NUM_REFLECTIONS + 0].position, 0, 27, 2);
        test_eq (diffuse [4 * NUM_REFLECTIONS + 0].position, -25, 2, 2);
        test_eq (diffuse [5 * NUM_REFLECTIONS + 0].position, 25, 2, 2);
        test_eq (diffuse [0 * NUM_REFLECTIONS + 1].position, 0, 0, 0);
        test_eq (diffuse [1 * NUM_REFLECTIONS + 1].position, 0, 0, 0);
        test_eq (diffuse [2 * NUM_REFLECTIONS + 1].position, 0, 27, 2);
        test_eq (diffuse [3 * NUM_REFLECTIONS + 1].position, 0, 0, 2);
        test_eq (diffuse [4 * NUM_REFLECTIONS + 1].position, -25, 2, -2);
        test_eq (diffuse [5 * NUM_REFLECTIONS + 1].position, 25, 2, -2);
        //  getAllRaw = diffuse followed by images (rayverb.cpp:708-714)
        auto images = getRawImages (false).impulses;
        EXPECT_TRUE (getAllRaw (false).impulses.size() == diffuse.size() + images.size());
        EXPECT_TRUE (getRawImages (true).impulses.size() + 1 == images.size());
    }
    vector <cl_float3> directions;
};

static vector <Impulse> attenuationInput (size_t n)
{
    default_random_engine generator;
    uniform_real_distribution <float> dist (0, 100);
    auto make = [&] (float x, float y, float z)
    {
        Impulse i;
        memset (&i, 0, sizeof i);
        for (int b = 0; b != 8; ++b) i.volume.s [b] = 1;
        i.position = cl_float3 {{x, y, z, 0}};
        i.time = dist (generator);
        return i;
    };
    vector <Impulse> in = {make (-10, 0, 0), make (10, 0, 0), make (0, -10, 0), make (0, 10, 0), make (0, 0, -10), make (0, 0, 10)};
    while (in.size() < n) in.push_back (make (0, 0, 0));
    return in;
}

// ---- AttenuationTest --------------------------------------------------------------------------------
struct AttenuationTest: public SpeakerAttenuator
{
    AttenuationTest(): in (attenuationInput (1024 * 64)) {}
    void run (float coefficient)
    {
        const Speaker speaker = {{{0, 0, 1, 0}}, coefficient, {0, 0, 0}};
        out = attenuate (RaytracerResults (in, cl_float3 {{0, 0, 0, 0}}), {speaker}).front();
        for (const auto & j : out)
            for (int i = 1; i != 8; ++i) EXPECT_FLOAT_EQ (j.volume.s [0], j.volume.s [i]);
    }
    void all()
    {
        run (0);
        for (const auto & j : out) EXPECT_FLOAT_EQ (j.volume.s [0], 1.0f);
        for (size_t i = 0; i != in.size(); ++i) EXPECT_TRUE (in [i].time == out [i].time);   // Timing: ASSERT_EQ
        run (0.5f);
        const float e1 [6] = {.5f, .5f, .5f, .5f, 0, 1};
        for (int i = 0; i != 6; ++i) EXPECT_FLOAT_EQ (out [i].volume.s [0], e1 [i]);
        run (1);
        const float e2 [6] = {0, 0, 0, 0, -1, 1};
        for (int i = 0; i != 6; ++i) EXPECT_FLOAT_EQ (out [i].volume.s [0], e2 [i]);
    }
    vector <Impulse> in;
    vector <AttenuatedImpulse> out;
};

// ---- HrtfTest ------------------------------------------------------------------------------------------
struct HrtfTest: public HrtfAttenuator
{
    HrtfTest(): in (attenuationInput (1000)), table (new HrtfTable)
    {
        //  the reference's synthetic test table (hrtf_analysis/generate_test_hrtf_data.py through
        //  analyse_hrtf.py:28-109): value (a', e, 0...), a' = a up to 345 then ramping to 0
        for (int ear = 0; ear != 2; ++ear)
            for (int a = 0; a != 360; ++a)
                for (int e = 0; e != 180; ++e)
                {
                    cl_float8 v = {{0, 0, 0, 0, 0, 0, 0, 0}};
                    v.s [0] = a <= 345 ? (float) a : (float) (345.0 + (0.0 - 345.0) * ((a - 345) / 15.0));
                    v.s [1] = (float) e;
                    (*table) [ear] [a] [e] = v;
                }
    }
    virtual const HrtfTable & getHrtfData() const { return *table; }
    void run (float fx, float fy, float fz, const int order [4])
    {
        const HrtfConfig config = {cl_float3 {{fx, fy, fz, 0}}, cl_float3 {{0, 1, 0, 0}}};
        out = attenuate (RaytracerResults (in, cl_float3 {{0, 0, 0, 0}}), config).front();
        const int cells [4] [2] = {{180, 90}, {0, 90}, {90, 90}, {270, 90}};
        for (int k = 0; k != 4; ++k)
            for (int i = 0; i != 8; ++i)
                EXPECT_FLOAT_EQ ((*table) [0] [cells [k] [0]] [cells [k] [1]].s [i], out [order [k]].volume.s [i]);
    }
    void all()
    {
        const int o0 [4] = {5, 4, 0, 1}, o1 [4] = {1, 0, 5, 4}, o2 [4] = {4, 5, 1, 0}, o3 [4] = {0, 1, 4, 5};
        run (0, 0, 1, o0); run (1, 0, 0, o1); run (0, 0, -1, o2); run (-1, 0, 0, o3);
    }
    vector <Impulse> in;
    vector <AttenuatedImpulse> out;
    unique_ptr <HrtfTable> table;
};

static void postprocess_test()
{
    //  flattenImpulses / fixPredelay / process glue (rayverb.cpp:28-149, rayverb.h:47-97)
    vector <AttenuatedImpulse> a (4);
    memset (a.data(), 0, sizeof (AttenuatedImpulse) * a.size());
    const float times [4] = {0.0005f, 0.0015f, 0.010f, 0.0f};
    for (int i = 0; i != 4; ++i) { a [i].time = times [i]; for (int b = 0; b != 8; ++b) a [i].volume.s [b] = (float) (i + 1); }
    vector <vector <AttenuatedImpulse>> channels = {a, a};
    EXPECT_FLOAT_EQ (findPredelay (channels), 0.0005f);
    auto flat = flattenImpulses (channels, 1000.0f);
    EXPECT_TRUE (flat.size() == 2 && flat [0].size() == 8 && flat [0] [0].size() == 11);
    EXPECT_FLOAT_EQ (flat [1] [3] [0], 4.0f);
    EXPECT_FLOAT_EQ (flat [1] [3] [1], 1.0f);
    EXPECT_FLOAT_EQ (flat [1] [3] [2], 2.0f);
    EXPECT_FLOAT_EQ (flat [1] [3] [10], 3.0f);
    fixPredelay (channels);
    EXPECT_FLOAT_EQ (channels [0] [0].time, 0.0f);
    EXPECT_FLOAT_EQ (channels [0] [1].time, 0.0015f - 0.0005f);
    auto out = process (RayverbFiltering::FILTER_TYPE_LINKWITZ_RILEY, flat, 44100.0f, true, 45.0f, false, 1.0f);
    EXPECT_TRUE (out.size() == 2 && out [0].size() == 11);
    EXPECT_FLOAT_EQ (max_amp (out), 1.0f);
    //  windowed sinc: the data grows by 28 samples (filters.cpp:126-131)
    auto flat2 = flattenImpulses (channels, 1000.0f);
    const size_t before = flat2 [0] [0].size();
    auto sinc = process (RayverbFiltering::FILTER_TYPE_WINDOWED_SINC, flat2, 44100.0f, true, 45.0f, false, 1.0f);
    EXPECT_TRUE (sinc.size() == 2 && sinc [0].size() == before + 28 && flat2 [0] [0].size() == before + 28);
    EXPECT_FLOAT_EQ (max_amp (sinc), 1.0f);
}

int main (int argc, char ** argv)
{
    if (argc != 3) { fprintf (stderr, "usage: host_tests <large_square.obj> <materials.json>\n"); return 2; }
    try
    {
        RaytracerTest (argv [1], argv [2]).run();
        AttenuationTest().all();
        HrtfTest().all();
        postprocess_test();
    }
    catch (const exception & e)
    {
        printf ("FAIL exception: %s\n", e.what());
        return 1;
    }
    printf ("%s: %d checks, %d failures\n", failures ? "FAILED" : "PASSED", checks, failures);
    return failures ? 1 : 0;
}
