// lt_device_types.h -- launch descriptor and counters read by the device code (lt_device.cuh).  Device-safe: no host
// headers, so that NVRTC can compile it for plug-in kernels (it is handed to NVRTC from memory, like lt_device.cuh).
#pragma once
#include "lens_trace_b200_device.cuh"  // buffer layouts

struct LtLaunch {
  int kernel;        // lt_kernel
  int kernelMode;    // 0 linear, 1 tile
  int width, height, depth;
  int maxRayDepth;
  int frames;
  unsigned frameStride;
  int accumMode;
  float accumWeight;
  int flags;
  int refillThreshold;  // k_path: leave the traversal loop when fewer lanes than this still have a ray
  int batchAnyHit;      // k_path: leaves recorded before a shadow ray tests them (early-out granularity)
  int batchClosest;     // k_path: leaves recorded before a closest-hit ray tests them
  int iterNodeSteps;    // trav_iter: box-pair tests per iteration of a persistent loop
  int iterTriTests;     // trav_iter: triangle tests per iteration
  // Row window of a tile split (multi-GPU): this launch renders `height` rows of an image of `fullHeight` rows;
  // local row j is image row ((j / rowBlock) * rowStride + rowPhase) * rowBlock + j % rowBlock (blocks of rowBlock
  // rows dealt round-robin to rowStride devices).  rowStride <= 1: the whole image (fullHeight == height).
  int fullHeight, rowBlock, rowStride, rowPhase;
  RefCamera cam;
};

struct LtCounters {
  unsigned long long rays, nodeTests, triTests;
};

