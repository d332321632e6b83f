// renderer_opencl.h -- RendererOpenCL with the reference's interface
// (include/lens_trace/opencl/renderer_opencl.h:16-43).  There is no OpenCL dispatch: the kernel
// file named by RenderPropertiesOpenCL::kernelFilePath selects one of the built-in sm_100a
// pipelines; an unknown .cl is a reported error.
#pragma once
#include <stdio.h>

#include "lens_trace/b200/renderer_b200.h"

class RendererOpenCL final : public Renderer {
private:
  RendererB200* impl;  // behind a pointer: the class layout stays fixed for compiled applications

public:
  RendererOpenCL();
  ~RendererOpenCL();

  RendererOpenCL(const RendererOpenCL&) = delete;
  RendererOpenCL& operator=(const RendererOpenCL&) = delete;

  void render(void* pRenderProperties);
};
