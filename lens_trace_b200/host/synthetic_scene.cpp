// synthetic_scene.cpp -- see include/lens_trace/b200/synthetic_scene.h.
#include "lens_trace/b200/synthetic_scene.h"

#include <math.h>
#include <stdio.h>

#include <vector>

namespace lt {
namespace {

struct Pcg32 {
  uint64_t state, inc;
  explicit Pcg32(uint64_t seed) : state(0), inc((seed << 1) | 1u) {
    next();
    state += 0x853c49e6748fea9bULL ^ seed;
    next();
  }
  uint32_t next() {
    uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
  float unit() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
};

void quad(FILE* f, const float p[4][3], const float n[3], int* vbase, int* nbase, const char* mtl) {
  fprintf(f, "usemtl %s\n", mtl);
  for (int i = 0; i < 4; i++) fprintf(f, "v %.6f %.6f %.6f\n", p[i][0], p[i][1], p[i][2]);
  fprintf(f, "vn %.4f %.4f %.4f\n", n[0], n[1], n[2]);
  int v = *vbase, k = *nbase;
  fprintf(f, "f %d//%d %d//%d %d//%d %d//%d\n", v + 1, k + 1, v + 2, k + 1, v + 3, k + 1, v + 4, k + 1);
  *vbase += 4;
  *nbase += 1;
}

}  // namespace

uint64_t writeSyntheticScene(const std::string& objPath, uint32_t gridN, uint64_t seed) {
  if (gridN < 1) return 0;
  std::string mtlPath = objPath;
  size_t dot = mtlPath.find_last_of('.');
  if (dot != std::string::npos) mtlPath = mtlPath.substr(0, dot);
  mtlPath += ".mtl";
  std::string mtlName = mtlPath;
  size_t slash = mtlName.find_last_of('/');
  if (slash != std::string::npos) mtlName = mtlName.substr(slash + 1);

  FILE* m = fopen(mtlPath.c_str(), "wb");
  if (!m) return 0;
  fprintf(m, "newmtl Light\nKd 0.800000 0.800000 0.800000\nKe 1.000000 1.000000 1.000000\nNi 1.450000\nd 1.000000\n\n");
  fprintf(m, "newmtl White\nKd 0.800000 0.800000 0.800000\nKe 0.000000 0.000000 0.000000\nNi 1.450000\nd 1.000000\n\n");
  fprintf(m, "newmtl Red\nKd 1.000000 0.000000 0.000000\nKe 0.000000 0.000000 0.000000\nNi 1.450000\nd 1.000000\n\n");
  fprintf(m, "newmtl Green\nKd 0.000000 1.000000 0.000000\nKe 0.000000 0.000000 0.000000\nNi 1.450000\nd 1.000000\n\n");
  fprintf(m, "newmtl Terrain\nKd 0.600000 0.700000 0.500000\nKe 0.000000 0.000000 0.000000\nNi 1.450000\nd 1.000000\n");
  fclose(m);

  FILE* f = fopen(objPath.c_str(), "wb");
  if (!f) return 0;
  static char buf[1 << 20];
  setvbuf(f, buf, _IOFBF, sizeof buf);
  fprintf(f, "# lens_trace_b200 synthetic scene: grid %u, seed %llu\nmtllib %s\no Room\n", gridN,
          (unsigned long long)seed, mtlName.c_str());
  int vbase = 0, nbase = 0;
  const float floorQ[4][3] = {{2.5f, 0, 2.5f}, {2.5f, 0, -2.5f}, {-2.5f, 0, -2.5f}, {-2.5f, 0, 2.5f}};
  const float ceilQ[4][3] = {{2.5f, 5, 2.5f}, {-2.5f, 5, 2.5f}, {-2.5f, 5, -2.5f}, {2.5f, 5, -2.5f}};
  const float backQ[4][3] = {{2.5f, 0, 2.5f}, {-2.5f, 0, 2.5f}, {-2.5f, 5, 2.5f}, {2.5f, 5, 2.5f}};
  const float rightQ[4][3] = {{2.5f, 5, 2.5f}, {2.5f, 5, -2.5f}, {2.5f, 0, -2.5f}, {2.5f, 0, 2.5f}};
  const float leftQ[4][3] = {{-2.5f, 0, 2.5f}, {-2.5f, 0, -2.5f}, {-2.5f, 5, -2.5f}, {-2.5f, 5, 2.5f}};
  const float lightQ[4][3] = {{1.0f, 4.98f, 1.0f}, {-1.0f, 4.98f, 1.0f}, {-1.0f, 4.98f, -1.0f}, {1.0f, 4.98f, -1.0f}};
  const float up[3] = {0, 1, 0}, down[3] = {0, -1, 0}, front[3] = {0, 0, -1}, nx[3] = {-1, 0, 0}, px[3] = {1, 0, 0};
  quad(f, floorQ, up, &vbase, &nbase, "White");
  quad(f, ceilQ, down, &vbase, &nbase, "White");
  quad(f, backQ, front, &vbase, &nbase, "White");
  quad(f, rightQ, nx, &vbase, &nbase, "Red");
  quad(f, leftQ, px, &vbase, &nbase, "Green");
  quad(f, lightQ, down, &vbase, &nbase, "Light");

  // height field: (gridN+1)^2 vertices, y = smooth hills + per-vertex PCG32 jitter
  const uint32_t n = gridN, nv = n + 1;
  std::vector<float> h((size_t)nv * nv);
  Pcg32 rng(seed);
  const float cell = 4.8f / (float)n;
  for (uint32_t j = 0; j < nv; j++)
    for (uint32_t i = 0; i < nv; i++) {
      float x = -2.4f + cell * i, z = -2.4f + cell * j;
      float hills = 0.9f + 0.45f * sinf(1.7f * x + 0.3f) * cosf(1.3f * z - 0.2f) + 0.25f * sinf(4.1f * x - 2.0f * z);
      float jitter = (rng.unit() - 0.5f) * cell * 0.8f;
      h[(size_t)j * nv + i] = hills + jitter;
    }
  fprintf(f, "o Terrain\nusemtl Terrain\n");
  for (uint32_t j = 0; j < nv; j++)
    for (uint32_t i = 0; i < nv; i++)
      fprintf(f, "v %.6f %.6f %.6f\n", -2.4f + cell * i, h[(size_t)j * nv + i], -2.4f + cell * j);
  for (uint32_t j = 0; j < nv; j++)
    for (uint32_t i = 0; i < nv; i++) {
      uint32_t i0 = i > 0 ? i - 1 : i, i1 = i < n ? i + 1 : i, j0 = j > 0 ? j - 1 : j, j1 = j < n ? j + 1 : j;
      float dx = (h[(size_t)j * nv + i1] - h[(size_t)j * nv + i0]) / (cell * (float)(i1 - i0));
      float dz = (h[(size_t)j1 * nv + i] - h[(size_t)j0 * nv + i]) / (cell * (float)(j1 - j0));
      float len = sqrtf(dx * dx + 1.0f + dz * dz);
      fprintf(f, "vn %.4f %.4f %.4f\n", -dx / len, 1.0f / len, -dz / len);
    }
  for (uint32_t j = 0; j < n; j++)
    for (uint32_t i = 0; i < n; i++) {
      int a = vbase + (int)(j * nv + i) + 1, b = a + 1, c = a + (int)nv, d = c + 1;
      int na = nbase + (int)(j * nv + i) + 1, nb = na + 1, nc = na + (int)nv, nd = nc + 1;
      fprintf(f, "f %d//%d %d//%d %d//%d\n", a, na, c, nc, b, nb);
      fprintf(f, "f %d//%d %d//%d %d//%d\n", b, nb, c, nc, d, nd);
    }
  fclose(f);
  return 12ull + 2ull * n * n;
}

}  // namespace lt
