// Shared driver of the two progressive examples (accumulator, global_illumination): the frame loop of
// the reference's examples/*/src/main.cpp:269-340 without the GL window.  Two modes:
//   per-frame : one Renderer::render per displayed frame, frameCount advanced by the application,
//               running mean applied by the application on the host (what the GL shader did);
//   batched   : one render() call for all frames with the RenderExtensionB200 on pNext; the running
//               mean stays on the device and only the final picture is copied back.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "lens_trace/acceleration_structure_explicit.h"
#include "lens_trace/camera.h"
#include "lens_trace/image_writer.h"
#include "lens_trace/model.h"
#include "lens_trace/opencl/renderer_opencl.h"
#include "lens_trace/structures.h"

inline int runProgressiveExample(const char* kernelPath, int argc, char** argv) {
  uint64_t width = 800, height = 800;
  uint32_t frames = 64, maxDepth = 0;
  bool perFrame = false;
  const char* outName = "output.ppm";
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--size") && i + 2 < argc) { width = strtoull(argv[++i], NULL, 10); height = strtoull(argv[++i], NULL, 10); }
    else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = (uint32_t)atoi(argv[++i]);
    else if (!strcmp(argv[i], "--depth") && i + 1 < argc) maxDepth = (uint32_t)atoi(argv[++i]);
    else if (!strcmp(argv[i], "--per-frame")) perFrame = true;
    else if (!strcmp(argv[i], "--out") && i + 1 < argc) outName = argv[++i];
  }
  std::vector<float> sample(width * height * 3), mean(width * height * 3, 0.0f);

  Camera camera(0, 2.5, -50, 0);
  Model model("resources/models/cornell_box.obj");
  AccelerationStructureExplicitProperties asProps = {};
  asProps.sType = STRUCTURE_TYPE_ACCELERATION_STRUCTURE_PROPERTIES;
  asProps.accelerationStructureExplicitType = ACCELERATION_STRUCTURE_TYPE_BVH;
  asProps.pModel = &model;
  AccelerationStructureExplicit accel(asProps);

  RendererOpenCL renderer;
  RenderPropertiesOpenCL props = {};
  props.sType = STRUCTURE_TYPE_RENDER_PROPERTIES_OPENCL;
  props.kernelFilePath = kernelPath;
  props.kernelMode = KERNEL_MODE_LINEAR;
  props.threadOrganizationMode = THREAD_ORGANIZATION_MODE_MAX_FIT;
  props.imageDimensions[0] = width;
  props.imageDimensions[1] = height;
  props.imageDimensions[2] = 3;
  props.pOutputBuffer = sample.data();
  props.outputBufferSize = sample.size() * sizeof(float);
  props.pAccelerationStructureExplicit = &accel;
  props.pModel = &model;
  props.pCamera = &camera;

  auto t0 = std::chrono::steady_clock::now();
  camera.resetFrameCount();  // a camera move restarts the accumulation with frameCount = 0
  if (perFrame) {
    RenderExtensionB200 one = {};  // only to forward --depth; one frame, no device accumulation
    one.sType = STRUCTURE_TYPE_RENDER_EXTENSION_B200;
    one.frames = 1;
    one.maxRayDepth = maxDepth;
    props.pNext = &one;
    for (uint32_t f = 0; f < frames; f++) {
      renderer.render(&props);
      uint32_t fc = camera.getFrameCount();
      for (size_t i = 0; i < mean.size(); i++)  // accumulator.frag:10-19
        mean[i] = fc > 0 ? (sample[i] + mean[i] * (float)fc) / (float)(fc + 1) : sample[i];
      camera.incrementFrameCount();
    }
  } else {
    RenderExtensionB200 ext = {};
    ext.sType = STRUCTURE_TYPE_RENDER_EXTENSION_B200;
    ext.frames = frames;
    ext.accumulate = 1;
    ext.maxRayDepth = maxDepth;
    props.pNext = &ext;
    props.pOutputBuffer = mean.data();
    renderer.render(&props);
    printf("device time %.3f ms for %u frames\n", ext.kernelMilliseconds, frames);
  }
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  printf("%u frames of %llux%llu in %.1f ms (%s)\n", frames, (unsigned long long)width, (unsigned long long)height, ms,
         perFrame ? "one render() per frame" : "one batched render()");

  BufferToImageProperties toImage = {};
  toImage.sType = STRUCTURE_TYPE_BUFFER_TO_IMAGE_PROPERTIES;
  toImage.pBuffer = mean.data();
  toImage.bufferSize = mean.size() * sizeof(float);
  toImage.imageDimensions[0] = width;
  toImage.imageDimensions[1] = height;
  toImage.imageDimensions[2] = 3;
  toImage.imageType = IMAGE_TYPE_JPEG;
  toImage.filename = outName;
  ImageWriter::writeBufferToImage(toImage);
  return 0;
}
