// acceleration_structure_explicit.h -- median-split BVH over a Model's triangles, flattened to the
// 32-byte DFS node array the kernels traverse (reference:
// include/lens_trace/acceleration_structure_explicit.h, src/acceleration_structure_explicit.cpp).
// Differences from the reference, all deliberate: centroid bounds are initialised (the reference
// reads them uninitialised, src/acceleration_structure_explicit.cpp:81-91, which makes its tree
// nondeterministic), primitives with coincident centroids are still split so every leaf holds one
// primitive (the kernels never test more than the first primitive of a leaf, basic.cu:168-172),
// nodes are emitted straight into the linear array, the light list is capped at its 64 slots, and
// the primitive buffer is freed.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "lens_trace/model.h"
#include "lens_trace/structures.h"

struct LinearBVHNode {
  float boundsMin[3];
  float boundsMax[3];

  union {
    int primitivesOffset;   // leaf
    int secondChildOffset;  // inner: first child is the next node
  };

  uint16_t primitiveCount;  // 0 = inner
  uint8_t axis;
  uint8_t pad[1];
};

struct Primitive {
  float positionA[3];
  float positionB[3];
  float positionC[3];
  float normalA[3];
  float normalB[3];
  float normalC[3];
  int materialIndex;
};

struct LightContainer {
  uint32_t count;
  uint32_t primitives[64];
};

class AccelerationStructureExplicit {
private:
  std::vector<LinearBVHNode> linearNodes;
  std::vector<Primitive> orderedPrimitives;
  LightContainer lightContainer;

public:
  AccelerationStructureExplicit(AccelerationStructureExplicitProperties accelerationStructureExplicitProperties);
  ~AccelerationStructureExplicit();

  uint64_t getNodeBufferSize();
  void* getNodeBuffer();

  uint64_t getOrderedPrimitiveBufferSize();
  void* getOrderedPrimitiveBuffer();

  uint64_t getLightContainerBufferSize();
  void* getLightContainerBuffer();
};
