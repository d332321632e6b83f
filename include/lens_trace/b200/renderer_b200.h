// renderer_b200.h -- the one renderer behind both RendererCUDA and RendererOpenCL.  Owns an lt_ctx
// (include/lens_trace_b200.h), uploads each AccelerationStructureExplicit once (cached by buffer
// identity) and maps kernelFilePath to a built-in pipeline.  Errors are printed and execution
// continues, as in the reference (src/cuda/renderer_cuda.cpp:44-46,134-136).
#pragma once
#include <map>
#include <string>

#include "lens_trace/acceleration_structure_explicit.h"
#include "lens_trace/camera.h"
#include "lens_trace/renderer.h"
#include "lens_trace/structures.h"

struct lt_ctx;
struct lt_scene;

class RendererB200 {
private:
  struct SceneKey {
    const void* nodes;
    const void* prims;
    const void* materials;
    uint64_t nodeBytes, primBytes, materialBytes;
    bool operator<(const SceneKey& o) const;
  };
  lt_ctx* ctx;
  std::map<SceneKey, lt_scene*> sceneCache;
  std::map<std::string, int> kernelCache;
  std::map<std::string, int> pluginCache;  // user .cu kernels compiled for this context

public:
  RendererB200();
  ~RendererB200();
  RendererB200(const RendererB200&) = delete;
  RendererB200& operator=(const RendererB200&) = delete;

  bool valid() const { return ctx != nullptr; }
  void forgetScenes();  // call when an AccelerationStructureExplicit was rebuilt in place

  void renderCommon(const std::string& kernelFilePath, KernelMode kernelMode, const uint64_t blockSize[2],
                    const uint64_t imageDimensions[3], void* pOutputBuffer, uint64_t outputBufferSize,
                    void* pAccelerationStructureExplicit, void* pModel, void* pCamera, void* pNext);
};
