// The example's global-illumination kernel (examples/global_illumination/resources/kernels/global_illumination.cl:
// shade :242-376, linearKernel :378-412) re-written as a CUDA plug-in on the device API of
// lens_trace_b200_device.cuh: lt_trace / lt_occluded / lt_random / lt_sample_light / lt_sample_hemisphere.
// One sample per launch, seeded by camera->frameCount, MAX_DEPTH bounces, clamped to [0, 1] (linearKernel).
// tests/test_gpu_host_api.py compares it bit for bit with the built-in LT_KERNEL_GI pipeline.
#include "lens_trace_b200_device.cuh"

#ifndef MAX_DEPTH
#define MAX_DEPTH 4
#endif

__device__ void giSample(const RefCamera* cam, int idx, int idy, int width, int height, float color[3]) {
  float fx, fy;
  LtRay ray = lt_camera_ray(*cam, idx, idy, width, height, fx, fy);
  const unsigned seed = cam->frameCount;
  float direct[3] = {0.0f, 0.0f, 0.0f}, indirect[3] = {0.0f, 0.0f, 0.0f};
  LtHit h = lt_trace(ray, 3.402823466e+38F, -1, false, LT_EPSILON_GI);
  if (lt_is_light(h.primitiveIndex)) {  // :257-266 (the payload's primitiveIndex is 0 after a miss)
    direct[0] = direct[1] = direct[2] = 1.0f;
  } else if (h.hitType == 1) {
    float pos[3], nrm[3], diffuse[3];
    lt_interpolate(h.primitiveIndex, h.u, h.v, pos, nrm);
    const RefMaterial& m = lt_material_of(h.primitiveIndex);
    diffuse[0] = m.diffuse[0]; diffuse[1] = m.diffuse[1]; diffuse[2] = m.diffuse[2];
    int hitPrim = h.primitiveIndex;
    // direct light: one shadow ray (:276-309)
    LtRay sr;
    float tMax = lt_sample_light(pos, lt_random(fx, fy, (float)seed), lt_random(fx, fy, (float)(seed + 1u)),
                                 lt_random(fx, fy, (float)(seed + 2u)), sr);
    bool lit = !lt_occluded(sr, tMax, hitPrim, LT_EPSILON_GI);
    float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(sr.dx, nrm[0]), __fmul_rn(sr.dy, nrm[1])), __fmul_rn(sr.dz, nrm[2])), 0.0f);
    if (lit)
      for (int k = 0; k < 3; k++) direct[k] = __fmul_rn(diffuse[k], d);
    // bounces (:310-372)
    int depth = 0;
    float dir[4];
    lt_sample_hemisphere(lt_random(fx, fy, (float)(seed + 3u)), lt_random(fx, fy, (float)(seed + 4u)), nrm, dir);
    LtRay ext;
    ext.ox = pos[0]; ext.oy = pos[1]; ext.oz = pos[2];
    ext.dx = dir[0]; ext.dy = dir[1]; ext.dz = dir[2];
    float extW = dir[3];
    while (depth < MAX_DEPTH) {
      LtHit eh = lt_trace(ext, 3.402823466e+38F, hitPrim, false, LT_EPSILON_GI);
      if (lt_is_light(eh.primitiveIndex)) {
        // :314-331 -- the ray is not advanced, so every remaining depth finds the light again
        for (; depth < MAX_DEPTH; depth++) {
          float w = (float)__ddiv_rn(1.0, (double)(depth + 1));
          float dd = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nrm[0], ext.dx), __fmul_rn(nrm[1], ext.dy)), __fmul_rn(nrm[2], ext.dz)),
                               __fmul_rn(1.0f, extW));
          float c = __fmul_rn(__fmul_rn(w, 1.0f), dd);
          for (int k = 0; k < 3; k++) indirect[k] = __fadd_rn(indirect[k], c);
        }
        break;
      }
      if (eh.hitType != 1) break;
      lt_interpolate(eh.primitiveIndex, eh.u, eh.v, pos, nrm);
      const RefMaterial& em = lt_material_of(eh.primitiveIndex);
      diffuse[0] = em.diffuse[0]; diffuse[1] = em.diffuse[1]; diffuse[2] = em.diffuse[2];
      hitPrim = eh.primitiveIndex;
      const unsigned sb = seed + (unsigned)depth + 5u;
      tMax = lt_sample_light(pos, lt_random(fx, fy, (float)sb), lt_random(fx, fy, (float)(sb + 1u)),
                             lt_random(fx, fy, (float)(sb + 2u)), sr);
      lit = !lt_occluded(sr, tMax, hitPrim, LT_EPSILON_GI);
      if (!lit) break;
      d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(sr.dx, nrm[0]), __fmul_rn(sr.dy, nrm[1])), __fmul_rn(sr.dz, nrm[2])), 0.0f);
      float w = (float)__ddiv_rn(1.0, (double)(depth + 1));
      for (int k = 0; k < 3; k++) indirect[k] = __fadd_rn(indirect[k], __fmul_rn(__fmul_rn(w, diffuse[k]), d));
      const unsigned se = seed + (unsigned)depth + 8u;
      depth++;
      if (depth >= MAX_DEPTH) break;  // (the reference still draws a direction here but never traces it)
      lt_sample_hemisphere(lt_random(fx, fy, (float)se), lt_random(fx, fy, (float)(se + 1u)), nrm, dir);
      ext.ox = pos[0]; ext.oy = pos[1]; ext.oz = pos[2];
      ext.dx = dir[0]; ext.dy = dir[1]; ext.dz = dir[2];
      extW = dir[3];
    }
  }
  for (int k = 0; k < 3; k++) color[k] = fminf(fmaxf(__fadd_rn(direct[k], indirect[k]), 0.0f), 1.0f);
}

extern "C" __global__ void linearKernel(void* nodes, void* prims, void* mats, void* lights, RefCamera* cam, float* out,
                                        int width, int height, int depth) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x, idy = blockIdx.y * blockDim.y + threadIdx.y;
  if (idx >= width || idy >= height) return;
  float c[3];
  giSample(cam, idx, idy, width, height, c);
  int id = (idy * width + idx) * depth;
  out[id + 0] = c[0];
  out[id + 1] = c[1];
  out[id + 2] = c[2];
}
