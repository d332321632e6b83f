// Test plug-in that uses the device API handed to plug-ins (lens_trace_b200_device.cuh): builds the reference's
// camera ray and calls lt_trace() on the re-flattened scene, then casts an any-hit shadow ray towards a fixed
// point.  Output per pixel: (bits of primitiveIndex+1 or 0, t, shadow hitType).
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
  LtRay r;
  r.ox = __fadd_rn(fx, cam->position[0]);
  r.oy = __fadd_rn(fy, cam->position[1]);
  r.oz = __fadd_rn(cam->position[2], 0.0f);
  r.dx = __fmaf_rn(dx0, c, __fmul_rn(s, 5.0f));
  r.dy = __fsub_rn(0.0f, fy);
  r.dz = __fmaf_rn(c, 5.0f, -__fmul_rn(dx0, s));
  LtHit h = lt_trace(r, 10000000.0f);
  float shadow = 0.0f;
  if (h.hitType == 1) {
    LtRay sr;
    sr.ox = r.ox + h.t * r.dx; sr.oy = r.oy + h.t * r.dy; sr.oz = r.oz + h.t * r.dz;
    sr.dx = 0.0f - sr.ox; sr.dy = 4.9f - sr.oy; sr.dz = 0.0f - sr.oz;
    LtHit sh = lt_trace(sr, 0.999f, h.primitiveIndex, true);
    shadow = (float)sh.hitType;
  }
  int id = (idy * width + idx) * depth;
  out[id + 0] = h.hitType == 1 ? __int_as_float(h.primitiveIndex + 1) : 0.0f;
  out[id + 1] = h.hitType == 1 ? h.t : 0.0f;
  out[id + 2] = shadow;
}
