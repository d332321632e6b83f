// renderer.h -- the drop-in boundary: one virtual render(void*) (reference: include/lens_trace/renderer.h:5-9).
#pragma once
#include <stddef.h>
#include <stdint.h>

class Renderer {
protected:
public:
  virtual void render(void* pRenderProperties) = 0;
};
