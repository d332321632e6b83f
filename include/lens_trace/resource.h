// resource.h -- path lookup used for kernel and model files (reference: src/resource.cpp:3-16).
#pragma once
#include <fstream>
#include <string>

class Resource {
public:
  // the path itself if it opens, else /usr/local/share/lens_trace/<path>, else "INVALID RESOURCE"
  static std::string findResource(std::string resourcePath);
};
