// resource.cpp -- see include/lens_trace/resource.h (reference behaviour: src/resource.cpp:3-16).
#include "lens_trace/resource.h"

static bool opens(const std::string& p) {
  std::ifstream f(p.c_str());
  return f.good();
}

std::string Resource::findResource(std::string resourcePath) {
  if (opens(resourcePath)) return resourcePath;
  const std::string installed = std::string("/usr/local/share/lens_trace/") + resourcePath;
  if (opens(installed)) return installed;
  return "INVALID RESOURCE";
}
