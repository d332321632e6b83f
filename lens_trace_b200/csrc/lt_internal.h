// lt_internal.h -- types shared by the CUDA kernels and the C-ABI layer of liblt_b200.so.
// Host-visible mirrors of the reference buffer layouts plus the re-flattened device layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// ---- reference layouts (as uploaded by the caller) ----
struct RefNode {  // include/lens_trace/acceleration_structure_explicit.h:20-32 (32 B)
  float boundsMin[3];
  float boundsMax[3];
  int32_t offset;  // primitivesOffset | secondChildOffset
  uint16_t primitiveCount;
  uint8_t axis;
  uint8_t pad;
};
static_assert(sizeof(RefNode) == 32, "LinearBVHNode is 32 bytes");

struct RefPrim {  // include/lens_trace/acceleration_structure_explicit.h:34-42 (76 B)
  float a[3], b[3], c[3];
  float na[3], nb[3], nc[3];
  int32_t materialIndex;
};
static_assert(sizeof(RefPrim) == 76, "Primitive is 76 bytes");

struct RefMaterial {  // include/lens_trace/model.h:26-31 (32 B)
  float diffuse[3];
  float ior;
  float dissolve;
  float emission[3];
};
static_assert(sizeof(RefMaterial) == 32, "Material is 32 bytes");

struct RefLights {  // include/lens_trace/acceleration_structure_explicit.h:44-47 (260 B)
  uint32_t count;
  uint32_t primitives[64];
};

struct RefCamera {  // src/camera.cpp:14-19 (28 B)
  float position[3];
  float yaw, pitch, roll;
  uint32_t frameCount;
};
static_assert(sizeof(RefCamera) == 28, "camera buffer is 28 bytes");

// ---- re-flattened device layouts ----
// One 64-byte record per INNER reference node, holding the boxes of both children so that one
// dependent round trip tests two reference nodes.  Child reference: >= 0 -> index of the child's
// own LtWideNode; < 0 -> ~primitivesOffset of a leaf child.  Four 128-bit loads.
struct __align__(16) LtWideNode {
  float4 bx;   // Lmin.x Lmax.x Rmin.x Rmax.x   (L = reference node i+1, R = secondChildOffset)
  float4 by;   // Lmin.y Lmax.y Rmin.y Rmax.y
  float4 bz;   // Lmin.z Lmax.z Rmin.z Rmax.z
  int4 meta;   // Lref, Rref, axis, (Lcount | Rcount << 16)  (leaf primitiveCount, stats only)
};
static_assert(sizeof(LtWideNode) == 64, "wide node is 64 bytes");

// 48-byte triangle record for the intersection test: A, e1 = B - A, e2 = C - A (the same IEEE
// subtractions basic.cu:100-101 performs per test, done once at upload).  Three 128-bit loads.
struct __align__(16) LtTri {
  float4 q0;  // A.x A.y A.z e1.x
  float4 q1;  // e1.y e1.z e2.x e2.y
  float4 q2;  // e2.z bits(materialIndex) 0 0
};
static_assert(sizeof(LtTri) == 48, "triangle record is 48 bytes");

#define LT_DONE ((int)0x80000000)  // traversal sentinel (== ~0x7fffffff, never a primitive)

struct LtSceneDev {
  const LtWideNode* wnodes;
  const LtTri* tris;
  const RefPrim* prims;       // original 76-byte records, read only when shading a hit
  const RefMaterial* mats;
  const RefLights* lights;
  float rootMin[3];
  float rootMax[3];
  int rootRef;                // child reference of the root (leaf root -> ~primitivesOffset)
  int rootCount;              // primitiveCount of a leaf root (stats only)
  int stackDepth;             // entries a traversal stack needs (tree depth), <= 64
  int nodeCount, primCount, matCount;
};

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

// launchers implemented in lt_kernels.cu (all asynchronous on `stream`; return launches made)
int lt_launch_reflatten(const RefNode* dNodes, int nodeCount, const RefPrim* dPrims, int primCount,
                        const int* dInnerRank, LtWideNode* dWide, LtTri* dTris, cudaStream_t stream);
int lt_launch_inner_flags(const RefNode* dNodes, int nodeCount, int* dFlags, cudaStream_t stream);
size_t lt_scan_temp_bytes(int n);
int lt_launch_exclusive_scan(void* dTemp, size_t tempBytes, const int* dIn, int* dOut, int n, cudaStream_t stream);
int lt_launch_render(const LtSceneDev& sc, const LtLaunch& L, float* dOut, LtCounters* dCounters,
                     cudaStream_t stream);
int lt_launch_primary_hits(const LtSceneDev& sc, const RefCamera& cam, int kernel, int flags, int width, int height,
                           int* dIds, int* dHit, float* dTuv, cudaStream_t stream);
int lt_launch_debug_random(const float* fx, const float* fy, const float* seed, int n, float* out, cudaStream_t stream);
int lt_launch_debug_hemisphere(const float* u1, const float* u2, const float* up, int n, float* out,
                               cudaStream_t stream);

// wavefront pipeline (lt_wavefront.cu)
size_t lt_wf_workspace_bytes_padded(long long nPaths);
// traceEvents: optional pool of 2*maxTraceLaunches events; when given, every traversal launch is bracketed by a
// pair of events and *traceLaunches receives the number of pairs recorded
int lt_launch_render_wavefront(const LtSceneDev& sc, const LtLaunch& L, float* dOut, LtCounters* dCounters,
                               void* workspace, int batchFrames, int smCount, cudaStream_t stream,
                               cudaEvent_t* traceEvents, int maxTraceLaunches, int* traceLaunches);

// user-written CUDA kernels with the reference's plug-in ABI (lt_plugin.cu)
#include <string>
struct LtPlugin;
LtPlugin* lt_plugin_compile(const char* path, std::string* err);
void lt_plugin_free(LtPlugin* p);
int lt_plugin_launch(LtPlugin* p, int kernelMode, const void* dNodes, const void* dPrims, const void* dMats,
                     const void* dLights, const void* dCamera, float* dOut, int width, int height, int depth, int bx,
                     int by, cudaStream_t stream, std::string* err);

// device-side LBVH construction in the reference's flattened layout (lt_bvh.cu)
size_t lt_bvh_scratch_bytes(int n);
int lt_launch_bvh_build(const RefPrim* dPrims, int n, const RefMaterial* dMats, int matCount, RefNode* dNodes,
                        RefPrim* dOrderedPrims, RefLights* dLights, void* scratch, size_t scratchBytes, int* outMaxDepth,
                        cudaStream_t stream);
