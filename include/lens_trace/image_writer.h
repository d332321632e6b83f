// image_writer.h -- float RGB buffer -> image file (reference: include/lens_trace/image_writer.h,
// src/image_writer.cpp:8-23).  JPEG encoding is out of scope here: the 8-bit conversion is the
// reference's (value * 255, no clamp) and the file is written as binary PPM; a ".pfm" filename
// writes the raw floats instead (lossless, for diffing).
#pragma once
#include "lens_trace/structures.h"

class ImageWriter {
public:
  static void writeBufferToImage(BufferToImageProperties bufferToImageProperties);
};
