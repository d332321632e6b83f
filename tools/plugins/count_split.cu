// Diagnostic plug-in: how evenly would a camera ray's traversal split over the four depth-2 subtrees of the root?
// Output per pixel: (total box-pair steps, steps in the largest of the four subtrees, steps in the larger of the two
// depth-1 subtrees).
#include "lens_trace_b200_device.cuh"

struct Camera { float position[3]; float yaw, pitch, roll; unsigned int frameCount; };

__device__ int countSubtree(int root, const Ray& r, float ix, float iy, float iz, bool nx, bool ny, bool nz) {
  if (root < 0) return 0;
  int steps = 0, sp = 0, cur = root;
  int stack[64];
  while (cur != LT_DONE) {
    if (cur >= 0) {
      steps++;
      const float4* np = reinterpret_cast<const float4*>(lt_scene.wnodes + cur);
      float4 bx = __ldg(np), by = __ldg(np + 1), bz = __ldg(np + 2);
      int4 m = __ldg(reinterpret_cast<const int4*>(np) + 3);
      bool hl = slab(nx ? bx.y : bx.x, nx ? bx.x : bx.y, ny ? by.y : by.x, ny ? by.x : by.y, nz ? bz.y : bz.x, nz ? bz.x : bz.y, r, ix, iy, iz);
      bool hr = slab(nx ? bx.w : bx.z, nx ? bx.z : bx.w, ny ? by.w : by.z, ny ? by.z : by.w, nz ? bz.w : bz.z, nz ? bz.z : bz.w, r, ix, iy, iz);
      if (hl) { cur = m.x; if (hr) stack[sp++] = m.y; }
      else if (hr) cur = m.y;
      else cur = sp > 0 ? stack[--sp] : LT_DONE;
    } else {
      cur = sp > 0 ? stack[--sp] : LT_DONE;
    }
  }
  return steps;
}

extern "C" __global__ void linearKernel(void* nodes, void* prims, void* mats, void* lights, Camera* cam, float* out,
                                        int width, int height, int depth) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x, idy = blockIdx.y * blockDim.y + threadIdx.y;
  if (idx >= width || idy >= height) return;
  float fx, fy;
  RefCamera rc;
  rc.position[0] = cam->position[0]; rc.position[1] = cam->position[1]; rc.position[2] = cam->position[2];
  rc.yaw = cam->yaw; rc.pitch = 0; rc.roll = 0; rc.frameCount = 0;
  Ray r = camera_ray(rc, idx, idy, width, height, fx, fy);
  float ix = __frcp_rn(r.dx), iy = __frcp_rn(r.dy), iz = __frcp_rn(r.dz);
  bool nx = ix < 0.0f, ny = iy < 0.0f, nz = iz < 0.0f;
  int id = (idy * width + idx) * depth;
  int total = 0, best4 = 0, best2 = 0;
  if (slab(nx ? lt_scene.rootMax[0] : lt_scene.rootMin[0], nx ? lt_scene.rootMin[0] : lt_scene.rootMax[0],
           ny ? lt_scene.rootMax[1] : lt_scene.rootMin[1], ny ? lt_scene.rootMin[1] : lt_scene.rootMax[1],
           nz ? lt_scene.rootMax[2] : lt_scene.rootMin[2], nz ? lt_scene.rootMin[2] : lt_scene.rootMax[2], r, ix, iy, iz) &&
      lt_scene.rootRef >= 0) {
    const float4* np = reinterpret_cast<const float4*>(lt_scene.wnodes + lt_scene.rootRef);
    float4 bx = __ldg(np), by = __ldg(np + 1), bz = __ldg(np + 2);
    int4 m = __ldg(reinterpret_cast<const int4*>(np) + 3);
    bool h[2];
    h[0] = slab(nx ? bx.y : bx.x, nx ? bx.x : bx.y, ny ? by.y : by.x, ny ? by.x : by.y, nz ? bz.y : bz.x, nz ? bz.x : bz.y, r, ix, iy, iz);
    h[1] = slab(nx ? bx.w : bx.z, nx ? bx.z : bx.w, ny ? by.w : by.z, ny ? by.z : by.w, nz ? bz.w : bz.z, nz ? bz.z : bz.w, r, ix, iy, iz);
    int child[2] = {m.x, m.y};
    total = 1;
    for (int c = 0; c < 2; c++) {
      if (!h[c] || child[c] < 0) continue;
      const float4* cp = reinterpret_cast<const float4*>(lt_scene.wnodes + child[c]);
      float4 cx = __ldg(cp), cy = __ldg(cp + 1), cz = __ldg(cp + 2);
      int4 cm = __ldg(reinterpret_cast<const int4*>(cp) + 3);
      bool gl = slab(nx ? cx.y : cx.x, nx ? cx.x : cx.y, ny ? cy.y : cy.x, ny ? cy.x : cy.y, nz ? cz.y : cz.x, nz ? cz.x : cz.y, r, ix, iy, iz);
      bool gr = slab(nx ? cx.w : cx.z, nx ? cx.z : cx.w, ny ? cy.w : cy.z, ny ? cy.z : cy.w, nz ? cz.w : cz.z, nz ? cz.z : cz.w, r, ix, iy, iz);
      int a = gl ? countSubtree(cm.x, r, ix, iy, iz, nx, ny, nz) : 0;
      int b = gr ? countSubtree(cm.y, r, ix, iy, iz, nx, ny, nz) : 0;
      total += 1 + a + b;
      best4 = max(best4, max(a, b));
      best2 = max(best2, 1 + a + b);
    }
  }
  out[id + 0] = (float)total;
  out[id + 1] = (float)best4;
  out[id + 2] = (float)best2;
}
