// synthetic_scene.h -- procedural benchmark meshes written as OBJ + MTL (BASELINE.json configs 3 and 5;
// SURVEY.md 8(d)).  Files, because Model only loads files (include/lens_trace/model.h).
#pragma once
#include <stdint.h>

#include <string>

namespace lt {

// Room (floor, back, left, right, ceiling quads, non-emissive), one emissive quad under the ceiling
// (2 triangles <= the 64 light slots) and a displaced gridN x gridN height field (2*gridN^2
// triangles) inside x,z in [-2.5,2.5], y in [0,5].  Vertex heights come from a PCG32 stream seeded
// with `seed`.  Every face carries normals and a material.  Returns the triangle count, 0 on failure.
uint64_t writeSyntheticScene(const std::string& objPath, uint32_t gridN, uint64_t seed);

}  // namespace lt
