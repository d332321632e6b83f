// renderer_cuda.h -- RendererCUDA with the reference's interface
// (include/lens_trace/cuda/renderer_cuda.h:16-29); implemented on the B200 kernels, no NVRTC.
#pragma once
#include <stdio.h>

#include "lens_trace/b200/renderer_b200.h"

class RendererCUDA final : public Renderer {
private:
  RendererB200 impl;

public:
  RendererCUDA();
  ~RendererCUDA();

  void render(void* pRenderProperties);
};
