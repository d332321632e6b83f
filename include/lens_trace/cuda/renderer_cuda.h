// renderer_cuda.h -- RendererCUDA with the reference's interface
// (include/lens_trace/cuda/renderer_cuda.h:16-29); implemented on the B200 kernels, no NVRTC.
#pragma once
#include <stdio.h>

#include "lens_trace/b200/renderer_b200.h"

class RendererCUDA final : public Renderer {
private:
  RendererB200* impl;  // behind a pointer: the class layout stays fixed for compiled applications

public:
  RendererCUDA();
  ~RendererCUDA();

  RendererCUDA(const RendererCUDA&) = delete;
  RendererCUDA& operator=(const RendererCUDA&) = delete;

  void render(void* pRenderProperties);
};
