// lt_internal.h -- types shared by the CUDA kernels and the C-ABI layer of liblt_b200.so.
// Host-visible mirrors of the reference buffer layouts plus the re-flattened device layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lens_trace_b200_device.cuh"  // buffer layouts shared with plug-in kernels

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
  RefCamera cam;
};

struct LtCounters {
  unsigned long long rays, nodeTests, triTests;
};

struct lt_ctx;
int lt_internal_ctx_device(const lt_ctx* ctx);  // lt_capi.cu

// launchers implemented in lt_kernels.cu (all asynchronous on `stream`; return launches made)
int lt_launch_reflatten(const RefNode* dNodes, int nodeCount, const RefPrim* dPrims, int primCount,
                        const int* dInnerRank, LtWideNode* dWide, LtTri* dTris, cudaStream_t stream);
int lt_launch_inner_flags(const RefNode* dNodes, int nodeCount, int* dFlags, cudaStream_t stream);
size_t lt_scan_temp_bytes(int n);
int lt_launch_exclusive_scan(void* dTemp, size_t tempBytes, const int* dIn, int* dOut, int n, cudaStream_t stream);
// dWork: one zero-initialisable int of device memory (work counter of the persistent kernels)
int lt_launch_render(const LtSceneDev& sc, const LtLaunch& L, float* dOut, LtCounters* dCounters, int* dWork,
                     int smCount, cudaStream_t stream);
int lt_launch_primary_hits(const LtSceneDev& sc, const RefCamera& cam, int kernel, int flags, int width, int height,
                           int* dIds, int* dHit, float* dTuv, cudaStream_t stream);
int lt_launch_debug_random(const float* fx, const float* fy, const float* seed, int n, float* out, cudaStream_t stream);
int lt_launch_debug_hemisphere(const float* u1, const float* u2, const float* up, int n, float* out,
                               cudaStream_t stream);

// wavefront pipeline (lt_wavefront.cu)
size_t lt_wf_workspace_bytes_padded(long long nPaths);
size_t lt_wf_primary_hits_bytes(long long pixels);  // per-launch primary hit records, kept behind the batch workspaces
// traceEvents: optional pool of 2*maxTraceLaunches events; when given, every traversal launch is bracketed by a
// pair of events and *traceLaunches receives the number of pairs recorded
// aux: extra streams + events (fork, one batch-order event per stream) for overlapping consecutive batches (the
// L1-bound trace kernel of one batch runs beside the DRAM-bound shade kernel of another and fills the tail of the
// persistent kernels); `workspace` then holds aux->streams
// batch workspaces of lt_wf_workspace_bytes_padded(batchFrames * pixels) each.  aux == nullptr: one stream.
#define LT_WF_MAX_STREAMS 4
struct LtWfAux {
  int streams;                              // batches in flight (2..LT_WF_MAX_STREAMS), one stream each
  cudaStream_t extra[LT_WF_MAX_STREAMS];    // [0] unused: side 0 runs on the caller's stream
  cudaEvent_t fork, order[LT_WF_MAX_STREAMS];
};
// pairKinds: optional, one LT_TIMED_* per recorded pair; when given, the shade / primary-shade / accumulate launches
// are bracketed as well (per-kernel times of a serialised step)
enum { LT_TIMED_TRAVERSAL = 0, LT_TIMED_SHADE = 1, LT_TIMED_PRIMARY_SHADE = 2, LT_TIMED_ACCUMULATE = 3, LT_TIMED_KINDS = 4 };
int lt_launch_render_wavefront(const LtSceneDev& sc, const LtLaunch& L, float* dOut, LtCounters* dCounters,
                               void* workspace, int batchFrames, int smCount, cudaStream_t stream,
                               cudaEvent_t* traceEvents, int maxTraceLaunches, int* traceLaunches,
                               unsigned char* pairKinds, const LtWfAux* aux);

// user-written CUDA kernels with the reference's plug-in ABI (lt_plugin.cu)
#include <string>
struct LtPlugin;
LtPlugin* lt_plugin_compile(const char* path, std::string* err);
void lt_plugin_free(LtPlugin* p);
int lt_plugin_launch(LtPlugin* p, int kernelMode, const LtSceneDev* sceneDev, const void* dNodes, const void* dPrims,
                     const void* dMats, const void* dLights, const void* dCamera, float* dOut, int width, int height,
                     int depth, int bx, int by, cudaStream_t stream, std::string* err);

// device-side LBVH construction in the reference's flattened layout (lt_bvh.cu)
size_t lt_bvh_scratch_bytes(int n);
int lt_launch_bvh_build(const RefPrim* dPrims, int n, const RefMaterial* dMats, int matCount, RefNode* dNodes,
                        RefPrim* dOrderedPrims, RefLights* dLights, void* scratch, size_t scratchBytes, int* outMaxDepth,
                        cudaStream_t stream);
