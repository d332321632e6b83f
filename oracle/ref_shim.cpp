// ref_shim.cpp -- flat C entry points over the UNMODIFIED reference classes, compiled by
// oracle/build_ref.sh together with the reference's own sources (where they lie under
// /root/reference) into oracle/_ref/libltref.so.  TEST INFRASTRUCTURE ONLY (see lt_oracle.h): used by
// tests/ to A/B the new kernels against the reference's CUDA backend on identical buffers, and by
// bench.py to report the reference kernel's own rate on the same GPU.
//
// Everything here is this project's code; it includes the reference headers and calls the
// reference's public API (tests/cuda_renderer_test.cc:12-49 shows the call sequence).
#include <cuda.h>
#include <nvrtc.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "lens_trace/cuda/renderer_cuda.h"
#include "lens_trace/model.h"

extern "C" {

void* ltref_model_create(const char* path) { return new Model(path); }
void ltref_model_destroy(void* m) { delete (Model*)m; }
uint64_t ltref_model_primitive_count(void* m) { return ((Model*)m)->getPrimitiveInfoListP()->size(); }
void* ltref_model_material_buffer(void* m) { return ((Model*)m)->getMaterialBuffer(); }
uint64_t ltref_model_material_bytes(void* m) { return ((Model*)m)->getMaterialBufferSize(); }

void* ltref_as_create(void* model) {
  AccelerationStructureExplicitProperties props = {};
  props.sType = STRUCTURE_TYPE_ACCELERATION_STRUCTURE_PROPERTIES;
  props.pNext = NULL;
  props.accelerationStructureExplicitType = ACCELERATION_STRUCTURE_TYPE_BVH;
  props.pModel = model;
  return new AccelerationStructureExplicit(props);
}
void ltref_as_destroy(void* a) { delete (AccelerationStructureExplicit*)a; }
void* ltref_as_node_buffer(void* a) { return ((AccelerationStructureExplicit*)a)->getNodeBuffer(); }
uint64_t ltref_as_node_bytes(void* a) { return ((AccelerationStructureExplicit*)a)->getNodeBufferSize(); }
void* ltref_as_primitive_buffer(void* a) { return ((AccelerationStructureExplicit*)a)->getOrderedPrimitiveBuffer(); }
uint64_t ltref_as_primitive_bytes(void* a) { return ((AccelerationStructureExplicit*)a)->getOrderedPrimitiveBufferSize(); }
void* ltref_as_light_buffer(void* a) { return ((AccelerationStructureExplicit*)a)->getLightContainerBuffer(); }
uint64_t ltref_as_light_bytes(void* a) { return ((AccelerationStructureExplicit*)a)->getLightContainerBufferSize(); }

void* ltref_camera_create(float x, float y, float z, float yaw) { return new Camera(x, y, z, yaw); }
void ltref_camera_destroy(void* c) { delete (Camera*)c; }
void* ltref_camera_buffer(void* c) { return ((Camera*)c)->getCameraBuffer(); }

void* ltref_renderer_cuda_create() { return new RendererCUDA(); }

// RendererCUDA::render exactly as an application calls it (src/cuda/renderer_cuda.cpp:41-140)
void ltref_render_cuda(void* renderer, const char* kernel_path, int kernel_mode, int custom_block, uint64_t bx,
                       uint64_t by, uint64_t w, uint64_t h, uint64_t d, float* out, uint64_t out_bytes, void* as,
                       void* model, void* camera) {
  RenderPropertiesCUDA p = {};
  p.sType = STRUCTURE_TYPE_RENDER_PROPERTIES_CUDA;
  p.pNext = NULL;
  p.kernelFilePath = kernel_path;
  p.kernelMode = kernel_mode ? KERNEL_MODE_TILE : KERNEL_MODE_LINEAR;
  p.threadOrganizationMode = custom_block ? THREAD_ORGANIZATION_MODE_CUSTOM : THREAD_ORGANIZATION_MODE_MAX_FIT;
  if (custom_block) {
    p.threadOrganization.sType = STRUCTURE_TYPE_THREAD_ORGANIZATION_CUDA;
    p.threadOrganization.pNext = NULL;
    p.threadOrganization.blockSize[0] = bx;
    p.threadOrganization.blockSize[1] = by;
  }
  p.imageDimensions[0] = w;
  p.imageDimensions[1] = h;
  p.imageDimensions[2] = d;
  p.pOutputBuffer = out;
  p.outputBufferSize = out_bytes;
  p.pAccelerationStructureExplicit = as;
  p.pModel = model;
  p.pCamera = camera;
  ((RendererCUDA*)renderer)->render(&p);
}

// Kernel-only time of the reference kernel: the same NVRTC compile (zero options), module load and
// launch shape as RendererCUDA::render, with the buffers resident and CUDA events around `iters`
// launches.  Returns average milliseconds per launch, < 0 on error.  Needs a current context
// (create a RendererCUDA first).
double ltref_time_kernel(const char* kernel_path, const char* entry, const void* nodes, uint64_t node_bytes,
                         const void* prims, uint64_t prim_bytes, const void* mats, uint64_t mat_bytes,
                         const void* lights, uint64_t light_bytes, const void* camera, uint64_t w, uint64_t h,
                         uint64_t d, unsigned bx, unsigned by, int warmup, int iters, float* out) {
  FILE* f = fopen(kernel_path, "rb");
  if (!f) return -1;
  std::string src;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) src.append(buf, n);
  fclose(f);
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), kernel_path, 0, NULL, NULL) != NVRTC_SUCCESS) return -2;
  if (nvrtcCompileProgram(prog, 0, NULL) != NVRTC_SUCCESS) return -3;
  size_t ptxSize;
  nvrtcGetPTXSize(prog, &ptxSize);
  std::string ptx(ptxSize, '\0');
  nvrtcGetPTX(prog, &ptx[0]);
  nvrtcDestroyProgram(&prog);
  CUmodule module;
  CUfunction fn;
  if (cuModuleLoadDataEx(&module, ptx.c_str(), 0, 0, 0) != CUDA_SUCCESS) return -4;
  if (cuModuleGetFunction(&fn, module, entry) != CUDA_SUCCESS) return -5;
  CUdeviceptr dN, dP, dM, dL, dC, dO;
  cuMemAlloc(&dN, node_bytes); cuMemcpyHtoD(dN, nodes, node_bytes);
  cuMemAlloc(&dP, prim_bytes); cuMemcpyHtoD(dP, prims, prim_bytes);
  cuMemAlloc(&dM, mat_bytes); cuMemcpyHtoD(dM, mats, mat_bytes);
  cuMemAlloc(&dL, light_bytes); cuMemcpyHtoD(dL, lights, light_bytes);
  cuMemAlloc(&dC, 28); cuMemcpyHtoD(dC, camera, 28);
  cuMemAlloc(&dO, sizeof(float) * w * h * d);
  void* args[] = {&dN, &dP, &dM, &dL, &dC, &dO, &w, &h, &d};
  unsigned gx = (unsigned)((w + bx - 1) / bx), gy = (unsigned)((h + by - 1) / by);
  CUevent e0, e1;
  cuEventCreate(&e0, 0);
  cuEventCreate(&e1, 0);
  for (int i = 0; i < warmup; i++) cuLaunchKernel(fn, gx, gy, 1, bx, by, 1, 0, NULL, args, 0);
  cuEventRecord(e0, NULL);
  for (int i = 0; i < iters; i++) cuLaunchKernel(fn, gx, gy, 1, bx, by, 1, 0, NULL, args, 0);
  cuEventRecord(e1, NULL);
  CUresult r = cuCtxSynchronize();
  float ms = 0;
  cuEventElapsedTime(&ms, e0, e1);
  if (out) cuMemcpyDtoH(out, dO, sizeof(float) * w * h * d);
  cuMemFree(dN); cuMemFree(dP); cuMemFree(dM); cuMemFree(dL); cuMemFree(dC); cuMemFree(dO);
  cuEventDestroy(e0);
  cuEventDestroy(e1);
  cuModuleUnload(module);
  if (r != CUDA_SUCCESS) return -6;
  return (double)ms / (iters > 0 ? iters : 1);
}

}  // extern "C"
