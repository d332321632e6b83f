// Diagnostic plug-in: per-pixel count of the box-pair steps and triangle tests the camera ray's traversal makes
// (reference order, no culling) -- the length of the ray's dependent load chain.  Output: (steps, triangle tests, t).
// Uses the library's internal device helpers (slab, tri_test) that the plug-in header brings along.
#include "lens_trace_b200_device.cuh"

struct Camera { float position[3]; float yaw, pitch, roll; unsigned int frameCount; };

extern "C" __global__ void linearKernel(void* nodes, void* prims, void* mats, void* lights, Camera* cam, float* out,
                                        int width, int height, int depth) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x, idy = blockIdx.y * blockDim.y + threadIdx.y;
  if (idx >= width || idy >= height) return;
  float fx = __fadd_rn(__fdiv_rn((float)idx, (float)width), -0.5f);
  float fy = __fadd_rn(__fdiv_rn((float)idy, (float)height), -0.5f);
  float c = cosf(cam->yaw), s = sinf(cam->yaw);
  float dx0 = __fsub_rn(0.0f, fx);
  Ray r;
  r.ox = __fadd_rn(fx, cam->position[0]);
  r.oy = __fadd_rn(fy, cam->position[1]);
  r.oz = __fadd_rn(cam->position[2], 0.0f);
  r.dx = __fmaf_rn(dx0, c, __fmul_rn(s, 5.0f));
  r.dy = __fsub_rn(0.0f, fy);
  r.dz = __fmaf_rn(c, 5.0f, -__fmul_rn(dx0, s));
  int steps = 0, tris = 0;
  Hit h;
  h.t = 10000000.0f; h.u = 0.0f; h.v = 0.0f; h.prim = 0; h.hit = 0;
  float ix = __frcp_rn(r.dx), iy = __frcp_rn(r.dy), iz = __frcp_rn(r.dz);
  bool nx = ix < 0.0f, ny = iy < 0.0f, nz = iz < 0.0f;
  unsigned negMask = (nx ? 1u : 0u) | (ny ? 2u : 0u) | (nz ? 4u : 0u);
  int id = (idy * width + idx) * depth;
  if (slab(nx ? lt_scene.rootMax[0] : lt_scene.rootMin[0], nx ? lt_scene.rootMin[0] : lt_scene.rootMax[0],
              ny ? lt_scene.rootMax[1] : lt_scene.rootMin[1], ny ? lt_scene.rootMin[1] : lt_scene.rootMax[1],
              nz ? lt_scene.rootMax[2] : lt_scene.rootMin[2], nz ? lt_scene.rootMin[2] : lt_scene.rootMax[2], r, ix, iy, iz)) {
    int stack[64];
    int sp = 0;
    int cur = lt_scene.rootRef;
    while (cur != LT_DONE) {
      if (cur >= 0) {
        steps++;
        const float4* np = reinterpret_cast<const float4*>(lt_scene.wnodes + cur);
        float4 bx = __ldg(np), by = __ldg(np + 1), bz = __ldg(np + 2);
        int4 m = __ldg(reinterpret_cast<const int4*>(np) + 3);
        bool hl = slab(nx ? bx.y : bx.x, nx ? bx.x : bx.y, ny ? by.y : by.x, ny ? by.x : by.y, nz ? bz.y : bz.x,
                          nz ? bz.x : bz.y, r, ix, iy, iz);
        bool hr = slab(nx ? bx.w : bx.z, nx ? bx.z : bx.w, ny ? by.w : by.z, ny ? by.z : by.w, nz ? bz.w : bz.z,
                          nz ? bz.z : bz.w, r, ix, iy, iz);
        bool axisNeg = (negMask >> m.z) & 1u;
        int nearRef = axisNeg ? m.y : m.x, farRef = axisNeg ? m.x : m.y;
        bool hn = axisNeg ? hr : hl, hf = axisNeg ? hl : hr;
        if (hn) {
          cur = nearRef;
          if (hf) stack[sp++] = farRef;
        } else if (hf) {
          cur = farRef;
        } else {
          cur = sp > 0 ? stack[--sp] : LT_DONE;
        }
      } else {
        tris++;
        int prim = ~cur;
        if (tri_test(lt_scene.tris, prim, r, 1.00000001168609742e-07f, h)) {
          h.prim = prim;
          h.hit = 1;
        }
        cur = sp > 0 ? stack[--sp] : LT_DONE;
      }
    }
  }
  out[id + 0] = (float)steps;
  out[id + 1] = (float)tris;
  out[id + 2] = h.hit == 1 ? h.t : 0.0f;
}
