// Wavefront OBJ + JSON materials loader.
//
// The reference loads models through Assimp (aiProcess_Triangulate | GenSmoothNormals | FlipUVs,
// rayverb.cpp:447-461), which is not available in this image.  This loader reproduces what the
// raytracer observes of that import (rayverb.cpp:331-445):
//   * surface 0 is the hard-coded default surface, surfaces 1.. are the materials of the JSON
//     file sorted by name (std::map order);
//   * every run of faces under one `usemtl` is a mesh whose material name selects the surface
//     (unknown names fall back to surface 0);
//   * vertices are NOT shared between faces (Assimp's OBJ importer emits one vertex per face
//     corner), so V = sum of face corners;
//   * polygons with more than three corners are triangulated (fan for convex quads, ear clipping
//     otherwise).  Triangle order inside a concave n-gon may differ from Assimp's; results that
//     are compared across implementations must therefore be computed from the same Triangle[]
//     (the array constructors, rayverb.h:142-148).
#include "rayverb.h"
#include "config.h"

#include <fstream>
#include <sstream>

using namespace std;

namespace
{
    struct P3 { double x, y, z; };

    double cross2 (const double * o, const double * a, const double * b)
    {
        return (a [0] - o [0]) * (b [1] - o [1]) - (a [1] - o [1]) * (b [0] - o [0]);
    }

    /// Triangulate a planar simple polygon; returns index triples into `poly`.
    vector <array <int, 3>> triangulate (const vector <P3> & poly)
    {
        const int n = (int) poly.size();
        vector <array <int, 3>> out;
        if (n < 3) return out;
        if (n == 3) { out.push_back ({{0, 1, 2}}); return out; }
        //  Newell normal -> drop the dominant axis
        double nx = 0, ny = 0, nz = 0;
        for (int i = 0; i != n; ++i)
        {
            const P3 & a = poly [i], & b = poly [(i + 1) % n];
            nx += (a.y - b.y) * (a.z + b.z);
            ny += (a.z - b.z) * (a.x + b.x);
            nz += (a.x - b.x) * (a.y + b.y);
        }
        const double ax = fabs (nx), ay = fabs (ny), az = fabs (nz);
        int drop = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
        const double nd = drop == 0 ? nx : (drop == 1 ? ny : nz);
        vector <array <double, 2>> q (n);
        for (int i = 0; i != n; ++i)
        {
            const double c [3] = {poly [i].x, poly [i].y, poly [i].z};
            double u = c [(drop + 1) % 3], v = c [(drop + 2) % 3];
            if (nd < 0) swap (u, v);
            q [i] = {{u, v}};
        }
        vector <int> idx (n);
        iota (idx.begin(), idx.end(), 0);
        int guard = 0;
        while (idx.size() > 3 && guard++ < 10 * n)
        {
            const int m = (int) idx.size();
            bool clipped = false;
            for (int kk = 0; kk != m; ++kk)
            {
                const int k = (kk + 1) % m;   // start at corner 1: a convex polygon comes out as a fan around vertex 0
                const int i0 = idx [(k + m - 1) % m], i1 = idx [k], i2 = idx [(k + 1) % m];
                if (cross2 (q [i0].data(), q [i1].data(), q [i2].data()) <= 1e-12) continue;   // reflex or flat corner
                bool contains = false;
                for (int j : idx)
                {
                    if (j == i0 || j == i1 || j == i2) continue;
                    if (cross2 (q [i0].data(), q [i1].data(), q [j].data()) >= 0 &&
                        cross2 (q [i1].data(), q [i2].data(), q [j].data()) >= 0 &&
                        cross2 (q [i2].data(), q [i0].data(), q [j].data()) >= 0) { contains = true; break; }
                }
                if (contains) continue;
                out.push_back ({{i0, i1, i2}});
                idx.erase (idx.begin() + k);
                clipped = true;
                break;
            }
            if (! clipped) break;   // degenerate polygon: fan over the rest
        }
        for (size_t k = 1; k + 1 < idx.size(); ++k) out.push_back ({{idx [0], idx [k], idx [k + 1]}});
        return out;
    }

    map <string, Surface> extractSurfaces (const string & materialFileName)   // rayverb.cpp:304-327
    {
        rvbjson::Document document;
        attemptJsonParse (materialFileName, document);
        if (! document.IsObject())
            throw runtime_error ("Materials must be stored in a JSON object");
        map <string, Surface> ret;
        for (const auto & m : document.members)
        {
            Surface surface;
            ValueJsonValidator <Surface> getter (surface);
            getter.run (m.second);
            ret [m.first] = surface;
        }
        return ret;
    }
}

void loadScene
(   const string & objpath
,   const string & materialFileName
,   bool verbose
,   vector <Triangle> & triangles
,   vector <cl_float3> & vertices
,   vector <Surface> & surfaces
)
{
    ifstream in (objpath);
    if (! in)
        throw runtime_error ("Failed to load object file.");

    const Surface defaultSurface =    // rayverb.cpp:336-339
    {   {{0.92f, 0.92f, 0.93f, 0.93f, 0.94f, 0.95f, 0.95f, 0.95f}}
    ,   {{0.50f, 0.90f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f}}
    };
    surfaces.clear(); vertices.clear(); triangles.clear();
    surfaces.push_back (defaultSurface);

    map <string, unsigned long> materialIndices;
    for (const auto & i : extractSurfaces (materialFileName))
    {
        surfaces.push_back (i.second);
        materialIndices [i.first] = surfaces.size() - 1;
    }

    vector <P3> positions;
    unsigned long mat_index = 0;
    string line;
    bool any_face = false;
    while (getline (in, line))
    {
        istringstream ls (line);
        string tag;
        if (! (ls >> tag)) continue;
        if (tag == "v")
        {
            P3 p {0, 0, 0};
            ls >> p.x >> p.y >> p.z;
            positions.push_back (p);
        }
        else if (tag == "usemtl")
        {
            string name;
            ls >> name;
            auto it = materialIndices.find (name);
            mat_index = it != materialIndices.end() ? it->second : 0;
            if (verbose)
                cerr << "Found mesh with material: " << name << " -> surface " << mat_index << endl;
        }
        else if (tag == "f")
        {
            vector <P3> corners;
            string tok;
            while (ls >> tok)
            {
                const long i = strtol (tok.c_str(), nullptr, 10);   // "v", "v/vt", "v/vt/vn", "v//vn"
                const long resolved = i > 0 ? i - 1 : (long) positions.size() + i;
                if (resolved < 0 || resolved >= (long) positions.size())
                    throw runtime_error ("Failed to load object file.");
                corners.push_back (positions [resolved]);
            }
            if (corners.size() < 3) continue;
            const unsigned long base = vertices.size();
            for (const auto & c : corners)
                vertices.push_back (cl_float3 {{(float) c.x, (float) c.y, (float) c.z, 0}});
            for (const auto & t : triangulate (corners))
                triangles.push_back (Triangle {mat_index, base + t [0], base + t [1], base + t [2]});
            any_face = true;
        }
    }
    if (! any_face)
        throw runtime_error ("Failed to load object file.");
    if (verbose)
        cerr << "Loaded 3D model with " << triangles.size() << " triangles" << endl;
}
