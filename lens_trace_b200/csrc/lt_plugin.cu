// lt_plugin.cu -- user-written CUDA kernels with the reference's plug-in ABI.
//
// The reference compiles the file named by RenderPropertiesCUDA::kernelFilePath at run time with NVRTC
// and launches its `linearKernel` / `tileKernel` entry point with
//   (LinearBVHNode*, Primitive*, Material*, LightContainer*, Camera*, float* out, int W, int H, int depth)
// (src/cuda/renderer_cuda.cpp:20-39,67-72,113-133; resources/kernels/cuda/basic.cu:331-340).  The seven
// shipped kernels map to built-in pipelines; any OTHER .cu file goes through here: compiled once per
// context for sm_100a, cached, and launched on the buffers of the uploaded scene in the reference's
// layouts.  NVRTC and the driver API are loaded lazily with dlopen so that liblt_b200.so itself has no
// link-time dependency on them (it must load on machines without a GPU).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "lt_internal.h"

namespace {

struct Api {
  bool tried = false, ok = false;
  std::string why;
  nvrtcResult (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*compileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*getProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*getCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*destroyProgram)(nvrtcProgram*);
  CUresult (*moduleLoadData)(CUmodule*, const void*);
  CUresult (*moduleUnload)(CUmodule);
  CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**);
  CUresult (*moduleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*);
  CUresult (*memcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream);
  CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int);
};

Api g_api;

template <class F>
bool sym(void* lib, const char* name, F& out, std::string& why) {
  out = (F)dlsym(lib, name);
  if (!out) why = std::string("missing symbol ") + name;
  return out != nullptr;
}

bool load_api() {
  Api& a = g_api;
  if (a.tried) return a.ok;
  a.tried = true;
  void* rtc = nullptr;
  const char* rtcNames[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                            "/usr/local/cuda/lib64/libnvrtc.so"};
  for (const char* n : rtcNames)
    if ((rtc = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
  void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!rtc || !drv) {
    a.why = !rtc ? "libnvrtc not found" : "libcuda.so.1 not found";
    return false;
  }
  a.ok = sym(rtc, "nvrtcCreateProgram", a.createProgram, a.why) && sym(rtc, "nvrtcCompileProgram", a.compileProgram, a.why) &&
         sym(rtc, "nvrtcGetProgramLogSize", a.getProgramLogSize, a.why) && sym(rtc, "nvrtcGetProgramLog", a.getProgramLog, a.why) &&
         sym(rtc, "nvrtcGetCUBINSize", a.getCUBINSize, a.why) && sym(rtc, "nvrtcGetCUBIN", a.getCUBIN, a.why) &&
         sym(rtc, "nvrtcDestroyProgram", a.destroyProgram, a.why) && sym(drv, "cuModuleLoadData", a.moduleLoadData, a.why) &&
         sym(drv, "cuModuleUnload", a.moduleUnload, a.why) && sym(drv, "cuModuleGetFunction", a.moduleGetFunction, a.why) &&
         sym(drv, "cuLaunchKernel", a.launchKernel, a.why) && sym(drv, "cuModuleGetGlobal_v2", a.moduleGetGlobal, a.why) &&
         sym(drv, "cuMemcpyHtoDAsync_v2", a.memcpyHtoDAsync, a.why) &&
         sym(drv, "cuFuncSetAttribute", a.funcSetAttribute, a.why);
  return a.ok;
}

}  // namespace

struct LtPlugin {
  std::string path;
  CUmodule module = nullptr;
  CUfunction linear = nullptr, tile = nullptr;
  CUdeviceptr sceneSymbol = 0;  // address of the plug-in's `lt_scene` constant, 0 if it does not use the device API
};

// texts of include/lens_trace_b200_device.cuh, lt_device.cuh and lt_device_types.h, handed to NVRTC as headers
#include "lt_device_api_embed.inc"

// Compiles `path`; on failure returns nullptr and the compiler log / reason in `err`.
LtPlugin* lt_plugin_compile(const char* path, std::string* err) {
  if (!load_api()) {
    *err = "plug-in kernels need NVRTC and the CUDA driver: " + g_api.why;
    return nullptr;
  }
  FILE* f = fopen(path, "rb");
  if (!f) {
    *err = std::string("cannot open kernel file ") + path;
    return nullptr;
  }
  std::string src;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) src.append(buf, n);
  fclose(f);
  nvrtcProgram prog;
  const char* headerNames[] = {"lens_trace_b200_device.cuh", "lt_device.cuh", "lt_device_types.h"};
  const char* headerTexts[] = {kEmbedDeviceApi, kEmbedDeviceCode, kEmbedDeviceTypes};
  if (g_api.createProgram(&prog, src.c_str(), path, 3, headerTexts, headerNames) != NVRTC_SUCCESS) {
    *err = "nvrtcCreateProgram failed";
    return nullptr;
  }
  // NVRTC's defaults otherwise, as in the reference (renderer_cuda.cpp:28: no options): a user's a * b + c contracts
  // exactly as it would there.  The library's own device code is written with explicit rounding intrinsics and does
  // not depend on the contraction mode.
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17"};
  nvrtcResult rc = g_api.compileProgram(prog, 2, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t ls = 0;
    g_api.getProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) g_api.getProgramLog(prog, &log[0]);
    *err = std::string("NVRTC could not compile ") + path + ":\n" + log;
    g_api.destroyProgram(&prog);
    return nullptr;
  }
  size_t cs = 0;
  g_api.getCUBINSize(prog, &cs);
  std::vector<char> cubin(cs);
  g_api.getCUBIN(prog, cubin.data());
  g_api.destroyProgram(&prog);
  LtPlugin* p = new LtPlugin();
  p->path = path;
  if (g_api.moduleLoadData(&p->module, cubin.data()) != CUDA_SUCCESS) {
    *err = "cuModuleLoadData failed for the compiled plug-in";
    delete p;
    return nullptr;
  }
  g_api.moduleGetFunction(&p->linear, p->module, "linearKernel");
  g_api.moduleGetFunction(&p->tile, p->module, "tileKernel");
  {
    size_t bytes = 0;
    CUdeviceptr addr = 0;
    if (g_api.moduleGetGlobal(&addr, &bytes, p->module, "lt_scene") == CUDA_SUCCESS && bytes == sizeof(LtSceneDev))
      p->sceneSymbol = addr;
  }
  if (!p->linear && !p->tile) {
    *err = std::string(path) + " exports neither linearKernel nor tileKernel (extern \"C\" __global__)";
    g_api.moduleUnload(p->module);
    delete p;
    return nullptr;
  }
  return p;
}

void lt_plugin_free(LtPlugin* p) {
  if (!p) return;
  if (p->module && g_api.ok) g_api.moduleUnload(p->module);
  delete p;
}

// Launch shape of the reference: grid = ceil(W/bx) x ceil(H/by), block = bx x by (32x1 for MAX_FIT,
// src/cuda/renderer_cuda.cpp:74-88).  Returns 0, or -1 with `err` set.
int lt_plugin_launch(LtPlugin* p, int kernelMode, const LtSceneDev* sceneDev, const void* dNodes, const void* dPrims,
                     const void* dMats, const void* dLights, const void* dCamera, float* dOut, int width, int height,
                     int depth, int bx, int by, cudaStream_t stream, std::string* err) {
  CUfunction fn = kernelMode ? p->tile : p->linear;
  if (p->sceneSymbol && sceneDev &&
      g_api.memcpyHtoDAsync(p->sceneSymbol, sceneDev, sizeof(LtSceneDev), (CUstream)stream) != CUDA_SUCCESS) {
    *err = "cannot set lt_scene of plug-in " + p->path;
    return -1;
  }
  if (!fn) {
    *err = p->path + (kernelMode ? " has no tileKernel" : " has no linearKernel");
    return -1;
  }
  if (bx <= 0 || by <= 0) {
    bx = 32;
    by = 1;
  }
  if ((long long)bx * by > 1024) {
    *err = "plug-in block size exceeds 1024 threads";
    return -1;
  }
  // a plug-in that uses the device API traverses with per-thread stacks / leaf FIFOs in dynamic shared memory
  unsigned smem = 0;
  if (p->sceneSymbol && sceneDev) {
    const int levels = sceneDev->tnodes ? 0 : (sceneDev->stackDepth < 1 ? 1 : sceneDev->stackDepth);
    smem = (unsigned)(levels + 16) * (unsigned)(bx * by) * 4u;
    if (smem > 200u * 1024u) {
      *err = "plug-in block too large for the traversal stacks of this scene (use a smaller block)";
      return -1;
    }
    if (smem > 48u * 1024u) g_api.funcSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
  }
  void* args[] = {(void*)&dNodes, (void*)&dPrims, (void*)&dMats, (void*)&dLights, (void*)&dCamera,
                  (void*)&dOut,   (void*)&width,  (void*)&height, (void*)&depth};
  CUresult rc = g_api.launchKernel(fn, (unsigned)((width + bx - 1) / bx), (unsigned)((height + by - 1) / by), 1,
                                   (unsigned)bx, (unsigned)by, 1, smem, (CUstream)stream, args, nullptr);
  if (rc != CUDA_SUCCESS) {
    *err = "cuLaunchKernel failed for plug-in " + p->path + " (CUresult " + std::to_string((int)rc) + ")";
    return -1;
  }
  return 0;
}
