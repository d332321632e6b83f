// scene_parser.h -- JSON .scene file -> Camera / Model / AccelerationStructureExplicit / RenderProperties*
// (reference: include/lens_trace/scene_parser.h, src/scene_parser.cpp; same keys, same defaults).
// The JSON reader is this project's own (lens_trace_b200/host/scene_parser.cpp).  B200 additions,
// all optional, under "renderer": "frames", "accumulate", "max_ray_depth".
#pragma once
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "lens_trace/acceleration_structure_explicit.h"
#include "lens_trace/camera.h"
#include "lens_trace/cuda/renderer_cuda.h"
#include "lens_trace/model.h"
#include "lens_trace/opencl/renderer_opencl.h"
#include "lens_trace/renderer.h"
#include "lens_trace/structures.h"

struct RendererParsed {
  RenderPlatform renderPlatform = RENDER_PLATFORM_OPENCL;
  std::string kernelFilePath = "resources/kernels/opencl/basic.cl";
  std::string kernelName = "basic";
  KernelMode kernelMode = KERNEL_MODE_LINEAR;
  ThreadOrganizationMode threadOrganizationMode = THREAD_ORGANIZATION_MODE_MAX_FIT;
  uint64_t workBlockSize[2] = {32, 32};
  uint64_t threadGroupSize[2] = {32, 32};
  uint64_t blockSize[2] = {32, 32};
  uint64_t imageDimensions[3] = {2048, 2048, 3};
  uint32_t frames = 1;        // B200
  uint32_t accumulate = 0;    // B200
  uint32_t maxRayDepth = 0;   // B200 (0 -> 16)
};

struct CameraParsed {
  float position[3] = {0, 0, 0};
  float pitch = 0;
  float yaw = 0;
  float roll = 0;
};

struct ModelParsed {
  std::string filePath;
};

struct WorldParsed {
  std::vector<ModelParsed> models;
};

struct OutputParsed {
  std::string filePath = "output.jpg";
};

class SceneParser {
private:
  RendererParsed rendererParsed;
  CameraParsed cameraParsed;
  WorldParsed worldParsed;
  OutputParsed outputParsed;
  bool parsedOk;

public:
  SceneParser(std::string filename);
  ~SceneParser();

  bool ok() const { return parsedOk; }

  uint64_t getOutputBufferSize();
  RenderPlatform getRenderPlatform();

  void* createOutputBuffer();

  Camera* createCamera();
  Model* createModel();
  AccelerationStructureExplicit* createAccelerationStructure(Model* model);

  RenderPropertiesOpenCL getRenderPropertiesOpenCL(void* outputBuffer,
                                                   AccelerationStructureExplicit* accelerationStructureExplicit,
                                                   Model* model, Camera* camera);
  RenderPropertiesCUDA getRenderPropertiesCUDA(void* outputBuffer,
                                               AccelerationStructureExplicit* accelerationStructureExplicit,
                                               Model* model, Camera* camera);
  BufferToImageProperties getBufferToImageProperties(void* outputBuffer);
  RenderExtensionB200 getRenderExtensionB200();
};
