// custom_kernel example: render the Cornell box once through RendererOpenCL with the
// custom (barycentric) kernel and write the picture.  Headless equivalent of the reference's
// examples/custom_kernel/src/main.cpp (same API calls, same 800x800 default).
//   usage: custom_kernel [width height [output.ppm|.pfm]]      (run from the repository root)
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "lens_trace/acceleration_structure_explicit.h"
#include "lens_trace/camera.h"
#include "lens_trace/image_writer.h"
#include "lens_trace/model.h"
#include "lens_trace/opencl/renderer_opencl.h"
#include "lens_trace/structures.h"

int main(int argc, char** argv) {
  const uint64_t width = argc > 2 ? strtoull(argv[1], NULL, 10) : 800;
  const uint64_t height = argc > 2 ? strtoull(argv[2], NULL, 10) : 800;
  const char* outName = argc > 3 ? argv[3] : "output.ppm";
  std::vector<float> image(width * height * 3);

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
  props.kernelFilePath = "examples/custom_kernel/resources/kernels/custom_opencl.cl";
  props.kernelMode = KERNEL_MODE_LINEAR;
  props.threadOrganizationMode = THREAD_ORGANIZATION_MODE_MAX_FIT;
  props.imageDimensions[0] = width;
  props.imageDimensions[1] = height;
  props.imageDimensions[2] = 3;
  props.pOutputBuffer = image.data();
  props.outputBufferSize = image.size() * sizeof(float);
  props.pAccelerationStructureExplicit = &accel;
  props.pModel = &model;
  props.pCamera = &camera;
  renderer.render(&props);

  BufferToImageProperties toImage = {};
  toImage.sType = STRUCTURE_TYPE_BUFFER_TO_IMAGE_PROPERTIES;
  toImage.pBuffer = image.data();
  toImage.bufferSize = image.size() * sizeof(float);
  toImage.imageDimensions[0] = width;
  toImage.imageDimensions[1] = height;
  toImage.imageDimensions[2] = 3;
  toImage.imageType = IMAGE_TYPE_JPEG;
  toImage.filename = outName;
  ImageWriter::writeBufferToImage(toImage);
  printf("custom_kernel: wrote %s (%llux%llu)\n", outName, (unsigned long long)width, (unsigned long long)height);
  return 0;
}
