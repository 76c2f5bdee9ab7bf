// dump_scene <model.obj> <materials.json> <out.bin>: runs the host scene loader and writes
// T, V, S (u64) followed by Triangle[T], cl_float3[V], Surface[S].  CPU only; used by the tests to
// compare the C++ loader with the Python one.
// dump_scene --diagnostic <out.dump>: print_diagnostic (helpers.cpp:19-59) on 2 rays x 3 fixed impulses.
#include "rayverb.h"
#include "helpers.h"

#include <cstring>

#include <fstream>

int main (int argc, char ** argv)
{
    if (argc == 3 && !strcmp (argv [1], "--diagnostic"))
    {
        std::vector <Impulse> imp (6);
        for (int i = 0; i < 6; ++i)
        {
            for (int b = 0; b < 8; ++b) imp [i].volume.s [b] = 0.125f * (float) (i + 1) * (b < 4 ? 1.0f : -0.5f);
            imp [i].position = cl_float3 {{(float) i, 27.0f, -0.1f * (float) i, 0}};
            imp [i].time = 0.01f * (float) i;
        }
        print_diagnostic (2, 3, imp, argv [2]);
        return 0;
    }
    if (argc != 4) return 2;
    try
    {
        std::vector <Triangle> t; std::vector <cl_float3> v; std::vector <Surface> s;
        loadScene (argv [1], argv [2], false, t, v, s);
        std::ofstream f (argv [3], std::ios::binary);
        const uint64_t hdr [3] = {t.size(), v.size(), s.size()};
        f.write ((const char *) hdr, sizeof hdr);
        f.write ((const char *) t.data(), t.size() * sizeof (Triangle));
        f.write ((const char *) v.data(), v.size() * sizeof (cl_float3));
        f.write ((const char *) s.data(), s.size() * sizeof (Surface));
    }
    catch (const std::exception & e)
    {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
