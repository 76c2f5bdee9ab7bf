// Ray direction generators (the API of rayverb/helpers.h:19-30).
#pragma once
#include "clstructs.h"

#include <string>
#include <vector>

/// Point on the unit sphere: -1 <= z <= 1, -pi <= theta <= pi (helpers.cpp:63-67).
cl_float3 spherePoint (float z, float theta);

/// `num` random unit vectors; seeded from the wall clock like the reference (helpers.cpp:69-81).
std::vector <cl_float3> getRandomDirections (unsigned long num);
/// Same generator with an explicit seed (the reference has no seed API).
std::vector <cl_float3> getRandomDirections (unsigned long num, unsigned long seed);

/// The reference's DIAGNOSTIC ray dump (helpers.cpp:19-59, consumed by viewer/viewer.pde): one JSON
/// array per ray and line, one {"position": [x, y, z], "volume": mean of the 8 bands} per reflection.
/// The reference compiles it in with -DDIAGNOSTIC; here it is always available and the CLI writes
/// "impulse.dump" when built with -DDIAGNOSTIC or run with RAYVERB_DIAGNOSTIC=1.
void print_diagnostic (unsigned long nrays, unsigned long nreflections, const std::vector <Impulse> & impulses,
                       const std::string & fname);
