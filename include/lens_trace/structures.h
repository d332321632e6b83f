// structures.h -- public enums and property structs of the renderer surface.
// Field order is API: callers use designated initialisers in declaration order
// (reference: include/lens_trace/structures.h:5-95, tests/cuda_renderer_test.cc:30-43).
// B200 additions are opt-in and chained through the (so far unused) pNext pointer.
#pragma once
#include <stdint.h>
#include <string>

enum StructureType {
  STRUCTURE_TYPE_RENDER_PROPERTIES_OPENCL,
  STRUCTURE_TYPE_THREAD_ORGANIZATION_OPENCL,
  STRUCTURE_TYPE_RENDER_PROPERTIES_CUDA,
  STRUCTURE_TYPE_THREAD_ORGANIZATION_CUDA,
  STRUCTURE_TYPE_BUFFER_TO_IMAGE_PROPERTIES,
  STRUCTURE_TYPE_ACCELERATION_STRUCTURE_PROPERTIES,
  // --- B200 extension (not in the reference) ---
  STRUCTURE_TYPE_RENDER_EXTENSION_B200 = 1000
};

enum RenderPlatform { RENDER_PLATFORM_OPENCL, RENDER_PLATFORM_CUDA, RENDER_PLATFORM_OPTIX };

enum KernelMode { KERNEL_MODE_LINEAR, KERNEL_MODE_TILE };

enum ThreadOrganizationMode { THREAD_ORGANIZATION_MODE_MAX_FIT, THREAD_ORGANIZATION_MODE_CUSTOM };

enum ImageType {
  IMAGE_TYPE_JPEG,
};

enum AccelerationStructureExplicitType {
  ACCELERATION_STRUCTURE_TYPE_BVH,
};

struct ThreadOrganizationOpenCL {
  StructureType sType;
  void* pNext;
  uint64_t workBlockSize[2];
  uint64_t threadGroupSize[2];
};

struct ThreadOrganizationCUDA {
  StructureType sType;
  void* pNext;
  uint64_t blockSize[2];
};

struct RenderPropertiesOpenCL {
  StructureType sType;
  void* pNext;
  std::string kernelFilePath;
  KernelMode kernelMode;
  ThreadOrganizationMode threadOrganizationMode;
  ThreadOrganizationOpenCL threadOrganization;
  uint64_t imageDimensions[3];
  void* pOutputBuffer;
  uint64_t outputBufferSize;
  void* pAccelerationStructureExplicit;
  void* pModel;
  void* pCamera;
};

struct RenderPropertiesCUDA {
  StructureType sType;
  void* pNext;
  std::string kernelFilePath;
  KernelMode kernelMode;
  ThreadOrganizationMode threadOrganizationMode;
  ThreadOrganizationCUDA threadOrganization;
  uint64_t imageDimensions[3];
  void* pOutputBuffer;
  uint64_t outputBufferSize;
  void* pAccelerationStructureExplicit;
  void* pModel;
  void* pCamera;
};

struct BufferToImageProperties {
  StructureType sType;
  void* pNext;
  void* pBuffer;
  uint64_t bufferSize;
  uint64_t imageDimensions[3];
  ImageType imageType;
  const char* filename;
};

struct AccelerationStructureExplicitProperties {
  StructureType sType;
  void* pNext;
  AccelerationStructureExplicitType accelerationStructureExplicitType;
  void* pModel;
};

// Optional, chained on RenderProperties*::pNext.  Lets one render() call take several frames and
// keep the running mean of examples/accumulator/resources/shaders/accumulator.frag on the device.
struct RenderExtensionB200 {
  StructureType sType;     // STRUCTURE_TYPE_RENDER_EXTENSION_B200
  void* pNext;
  uint32_t frames;         // frames per render() call (0 -> 1); frame k uses frameCount + k
  uint32_t accumulate;     // 0: output = last sample; 1: running mean (accumulator.frag:10-19)
  uint32_t maxRayDepth;    // GI bounce cap (0 -> 16, global_illumination.cl:308)
  uint32_t collectStats;   // fill rays/nodeTests/triTests below (slower; not for timing)
  uint64_t rays;           // out
  uint64_t nodeTests;      // out
  uint64_t triTests;       // out
  float kernelMilliseconds;  // out: device time of the render kernels
};
