// renderer_b200.cpp -- RendererCUDA / RendererOpenCL on top of the C-ABI (include/lens_trace_b200.h).
// Replaces src/cuda/renderer_cuda.cpp:41-140 and src/opencl/renderer_opencl.cpp:56-153.
#include <stdio.h>
#include <string.h>

#include "lens_trace/cuda/renderer_cuda.h"
#include "lens_trace/model.h"
#include "lens_trace/opencl/renderer_opencl.h"
#include "lens_trace/resource.h"
#include "lens_trace_b200.h"

bool RendererB200::SceneKey::operator<(const SceneKey& o) const { return memcmp(this, &o, sizeof(SceneKey)) < 0; }

RendererB200::RendererB200() : ctx(nullptr) {
  if (lt_ctx_create(0, &ctx) != LT_OK) {
    printf("ERROR: %s\n", lt_last_error(nullptr));
    ctx = nullptr;
  }
}

RendererB200::~RendererB200() {
  forgetScenes();
  if (ctx) lt_ctx_destroy(ctx);
}

void RendererB200::forgetScenes() {
  for (std::map<SceneKey, lt_scene*>::iterator it = sceneCache.begin(); it != sceneCache.end(); ++it)
    lt_scene_release(ctx, it->second);
  sceneCache.clear();
}

void RendererB200::renderCommon(const std::string& kernelFilePath, KernelMode kernelMode, const uint64_t blockSize[2],
                                const uint64_t imageDimensions[3], void* pOutputBuffer, uint64_t outputBufferSize,
                                void* pAccelerationStructureExplicit, void* pModel, void* pCamera, void* pNext) {
  if (!ctx) {
    printf("ERROR: no B200 context; nothing rendered (there is no CPU fallback)\n");
    return;
  }
  AccelerationStructureExplicit* as = (AccelerationStructureExplicit*)pAccelerationStructureExplicit;
  Model* model = (Model*)pModel;
  Camera* camera = (Camera*)pCamera;
  if (!as || !model || !camera || !pOutputBuffer) {
    printf("ERROR: render properties carry a NULL object\n");
    return;
  }

  int kernel;
  std::map<std::string, int>::iterator kc = kernelCache.find(kernelFilePath);
  if (kc != kernelCache.end()) {
    kernel = kc->second;
  } else {
    std::string resolved = Resource::findResource(kernelFilePath);
    kernel = lt_kernel_from_path(resolved == "INVALID RESOURCE" ? kernelFilePath.c_str() : resolved.c_str());
    kernelCache[kernelFilePath] = kernel;
  }
  // not one of the shipped kernels: a user-written .cu with the reference's kernel ABI is compiled
  // (once) with NVRTC for sm_100a and launched like the reference launches it; anything else is an error
  int plugin = -1;
  if (kernel < 0) {
    std::string resolved = Resource::findResource(kernelFilePath);
    bool isCu = resolved.size() > 3 && resolved.compare(resolved.size() - 3, 3, ".cu") == 0;
    std::map<std::string, int>::iterator pc = pluginCache.find(kernelFilePath);
    if (pc != pluginCache.end()) {
      plugin = pc->second;
    } else if (isCu) {
      if (lt_plugin_load(ctx, resolved.c_str(), &plugin) != LT_OK) {
        printf("%s\n", lt_last_error(ctx));
        plugin = -1;
      }
      pluginCache[kernelFilePath] = plugin;
    }
    if (plugin < 0) {
      printf("Kernel Error: '%s' is not a kernel this renderer provides (%s)\n", kernelFilePath.c_str(),
             isCu ? "plug-in compilation failed" : "only the shipped kernels and CUDA plug-ins are supported");
      return;
    }
  }

  SceneKey key;
  memset(&key, 0, sizeof key);
  key.nodes = as->getNodeBuffer();
  key.prims = as->getOrderedPrimitiveBuffer();
  key.materials = model->getMaterialBuffer();
  key.nodeBytes = as->getNodeBufferSize();
  key.primBytes = as->getOrderedPrimitiveBufferSize();
  key.materialBytes = model->getMaterialBufferSize();
  lt_scene* scene = nullptr;
  std::map<SceneKey, lt_scene*>::iterator sc = sceneCache.find(key);
  if (sc != sceneCache.end()) {
    scene = sc->second;
  } else {
    int rc = lt_scene_upload(ctx, key.nodes, key.nodeBytes, key.prims, key.primBytes, key.materials, key.materialBytes,
                             as->getLightContainerBuffer(), as->getLightContainerBufferSize(), &scene);
    if (rc != LT_OK) {
      printf("ERROR: scene upload failed: %s\n", lt_last_error(ctx));
      return;
    }
    sceneCache[key] = scene;
  }

  RenderExtensionB200* ext = nullptr;
  for (void* p = pNext; p;) {
    RenderExtensionB200* e = (RenderExtensionB200*)p;
    if (e->sType == STRUCTURE_TYPE_RENDER_EXTENSION_B200) {
      ext = e;
      break;
    }
    p = e->pNext;
  }

  lt_render_params params;
  memset(&params, 0, sizeof params);
  params.struct_size = sizeof params;
  params.kernel = kernel;
  params.kernel_mode = kernelMode == KERNEL_MODE_TILE ? 1 : 0;
  params.width = (int)imageDimensions[0];
  params.height = (int)imageDimensions[1];
  params.depth = (int)imageDimensions[2];
  params.frames = 1;
  params.frame_stride = 1;
  params.block_x = blockSize ? (int)blockSize[0] : 0;
  params.block_y = blockSize ? (int)blockSize[1] : 0;
  if (ext) {
    params.frames = ext->frames ? (int)ext->frames : 1;
    params.accum_mode = ext->accumulate ? LT_ACCUM_RUNNING_MEAN : LT_ACCUM_NONE;
    params.max_ray_depth = (int)ext->maxRayDepth;
    if (ext->collectStats) params.flags |= LT_FLAG_STATS;
  }
  uint64_t need = sizeof(float) * imageDimensions[0] * imageDimensions[1] * imageDimensions[2];
  if (outputBufferSize < need) {
    printf("ERROR: output buffer is %llu bytes, image needs %llu\n", (unsigned long long)outputBufferSize,
           (unsigned long long)need);
    return;
  }
  int rc = plugin >= 0
               ? lt_render_plugin(ctx, scene, camera->getCameraBuffer(), plugin, params.kernel_mode, params.width,
                                  params.height, params.depth, params.block_x, params.block_y, (float*)pOutputBuffer)
               : lt_render(ctx, scene, camera->getCameraBuffer(), &params, (float*)pOutputBuffer);
  if (rc != LT_OK) {
    printf("Kernel Error: %d (%s)\n", rc, lt_last_error(ctx));
    return;
  }
  if (ext) {
    lt_stats st;
    lt_last_stats(ctx, &st);
    ext->rays = st.rays;
    ext->nodeTests = st.node_tests;
    ext->triTests = st.tri_tests;
    ext->kernelMilliseconds = st.kernel_ms;
  }
}

RendererCUDA::RendererCUDA() : impl(new RendererB200()) {}
RendererCUDA::~RendererCUDA() { delete impl; }

void RendererCUDA::render(void* pRenderProperties) {
  RenderPropertiesCUDA* p = (RenderPropertiesCUDA*)pRenderProperties;
  if (p->sType != STRUCTURE_TYPE_RENDER_PROPERTIES_CUDA) printf("ERROR: RenderPropertiesCUDA sType\n");
  uint64_t block[2] = {32, 1};
  if (p->threadOrganizationMode == THREAD_ORGANIZATION_MODE_CUSTOM) {
    if (p->threadOrganization.sType != STRUCTURE_TYPE_THREAD_ORGANIZATION_CUDA)
      printf("ERROR: ThreadOrganizationCUDA sType\n");
    block[0] = p->threadOrganization.blockSize[0];
    block[1] = p->threadOrganization.blockSize[1];
  }
  impl->renderCommon(p->kernelFilePath, p->kernelMode, block, p->imageDimensions, p->pOutputBuffer, p->outputBufferSize,
                    p->pAccelerationStructureExplicit, p->pModel, p->pCamera, p->pNext);
}

RendererOpenCL::RendererOpenCL() : impl(new RendererB200()) {}
RendererOpenCL::~RendererOpenCL() { delete impl; }

void RendererOpenCL::render(void* pRenderProperties) {
  RenderPropertiesOpenCL* p = (RenderPropertiesOpenCL*)pRenderProperties;
  if (p->sType != STRUCTURE_TYPE_RENDER_PROPERTIES_OPENCL) printf("ERROR: RenderPropertiesOpenCL sType\n");
  uint64_t block[2] = {0, 0};
  if (p->threadOrganizationMode == THREAD_ORGANIZATION_MODE_CUSTOM) {
    if (p->threadOrganization.sType != STRUCTURE_TYPE_THREAD_ORGANIZATION_OPENCL)
      printf("ERROR: ThreadOrganizationOpenCL sType\n");
    block[0] = p->threadOrganization.threadGroupSize[0];
    block[1] = p->threadOrganization.threadGroupSize[1];
  }
  // The reference splits the image into work blocks and drops the remainder
  // (src/opencl/renderer_opencl.cpp:90); the whole image is rendered here.
  impl->renderCommon(p->kernelFilePath, p->kernelMode, block, p->imageDimensions, p->pOutputBuffer, p->outputBufferSize,
                    p->pAccelerationStructureExplicit, p->pModel, p->pCamera, p->pNext);
}
