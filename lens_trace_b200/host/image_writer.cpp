// image_writer.cpp -- see include/lens_trace/image_writer.h.
#include "lens_trace/image_writer.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

void ImageWriter::writeBufferToImage(BufferToImageProperties props) {
  if (props.sType != STRUCTURE_TYPE_BUFFER_TO_IMAGE_PROPERTIES || !props.pBuffer || !props.filename) return;
  const float* image = (const float*)props.pBuffer;
  const uint64_t w = props.imageDimensions[0], h = props.imageDimensions[1], d = props.imageDimensions[2];
  std::string name(props.filename);
  bool pfm = name.size() > 4 && name.compare(name.size() - 4, 4, ".pfm") == 0;
  bool ppm = name.size() > 4 && name.compare(name.size() - 4, 4, ".ppm") == 0;
  if (!pfm && !ppm) {
    name += ".ppm";
    printf("ImageWriter: JPEG encoding is not built in; writing %s\n", name.c_str());
  }
  FILE* f = fopen(name.c_str(), "wb");
  if (!f) {
    printf("ImageWriter: cannot open %s\n", name.c_str());
    return;
  }
  if (pfm) {
    fprintf(f, "PF\n%llu %llu\n-1.0\n", (unsigned long long)w, (unsigned long long)h);
    for (uint64_t y = 0; y < h; y++)
      for (uint64_t x = 0; x < w; x++) fwrite(image + (y * w + x) * d, sizeof(float), 3, f);
  } else {
    fprintf(f, "P6\n%llu %llu\n255\n", (unsigned long long)w, (unsigned long long)h);
    for (uint64_t i = 0; i < w * h; i++) {
      unsigned char px[3];
      for (int c = 0; c < 3; c++) px[c] = (unsigned char)(char)(image[i * d + c] * 255);  // src/image_writer.cpp:18
      fwrite(px, 1, 3, f);
    }
  }
  fclose(f);
}
