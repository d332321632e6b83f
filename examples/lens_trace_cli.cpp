// LensTrace <file.scene>: the reference's command-line program (src/main.cpp) on this repo's surface.
#include <stdio.h>
#include <stdlib.h>

#include "lens_trace/image_writer.h"
#include "lens_trace/scene_parser.h"

int main(int argc, const char** argv) {
  if (argc < 2) {
    printf("usage: LensTrace <file.scene>\n");
    return 2;
  }
  SceneParser parser(argv[1]);
  if (!parser.ok()) return 1;
  void* out = parser.createOutputBuffer();
  Camera* camera = parser.createCamera();
  Model* model = parser.createModel();
  if (!model) return 1;
  AccelerationStructureExplicit* accel = parser.createAccelerationStructure(model);
  RenderExtensionB200 ext = parser.getRenderExtensionB200();
  if (parser.getRenderPlatform() == RENDER_PLATFORM_OPENCL) {
    RendererOpenCL renderer;
    RenderPropertiesOpenCL p = parser.getRenderPropertiesOpenCL(out, accel, model, camera);
    p.pNext = &ext;
    renderer.render(&p);
  } else if (parser.getRenderPlatform() == RENDER_PLATFORM_CUDA) {
    RendererCUDA renderer;
    RenderPropertiesCUDA p = parser.getRenderPropertiesCUDA(out, accel, model, camera);
    p.pNext = &ext;
    renderer.render(&p);
  }
  ImageWriter::writeBufferToImage(parser.getBufferToImageProperties(out));
  delete accel;
  delete model;
  delete camera;
  free(out);
  return 0;
}
