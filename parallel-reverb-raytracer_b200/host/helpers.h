// Ray direction generators (the API of rayverb/helpers.h:19-30).
#pragma once
#include "clstructs.h"

#include <vector>

/// Point on the unit sphere: -1 <= z <= 1, -pi <= theta <= pi (helpers.cpp:63-67).
cl_float3 spherePoint (float z, float theta);

/// `num` random unit vectors; seeded from the wall clock like the reference (helpers.cpp:69-81).
std::vector <cl_float3> getRandomDirections (unsigned long num);
/// Same generator with an explicit seed (the reference has no seed API).
std::vector <cl_float3> getRandomDirections (unsigned long num, unsigned long seed);
