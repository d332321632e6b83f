// lens_trace_b200_device.cuh -- device-side view of an uploaded scene, for plug-in kernels.
//
// A user-written .cu kernel with the reference's kernel ABI (linearKernel / tileKernel, see
// include/lens_trace_b200.h: lt_plugin_load) may `#include "lens_trace_b200_device.cuh"` -- the library hands
// this header to NVRTC itself -- and call lt_trace() instead of walking the 32-byte node array on its own.
// lt_trace() has the reference's intersect()/intersectIgnorePrimitiveIndex() semantics (basic.cu:156-243: near
// child first, strict t < best, no t > 0 test, first primitive of a leaf only) and the same FP32 operation
// order as the built-in pipelines, on the re-flattened 64-byte child-pair nodes and 48-byte triangles.
// The library fills `lt_scene` (a __constant__ in the plug-in's module) before every launch.
//
// The same file defines the buffer layouts for the library itself (lens_trace_b200/csrc/lt_internal.h).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef int int32_t;
typedef unsigned int uint32_t;
typedef unsigned short uint16_t;
typedef unsigned char uint8_t;
#endif

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

// Threaded ("skip pointer") form of the same tree, built for scenes small enough that eight copies stay
// cache resident.  The reference visits children near-first by the sign of the ray direction on the split
// axis and never culls by t, so for each of the 8 sign octants the visit order of the whole tree is one
// fixed depth-first sequence.  Copy o (records [o*nodeCount, (o+1)*nodeCount)) holds the reference nodes in
// that sequence: the next record is the first node to visit after a box hit, `skip` the first node after
// the subtree -- a traversal is `i = hit && inner ? i + 1 : skip` with no stack.  The bounds are stored
// already selected by the octant's dirIsNeg (basic.cu:88-91), so the slab test needs no selects either.
// One 32-byte sector, two 128-bit loads.
struct __align__(16) LtThreadNode {
  float4 a;  // lo.x lo.y lo.z hi.x      lo = dirIsNeg ? boundsMax : boundsMin, hi the other one
  float4 b;  // hi.y hi.z bits(link) bits(skip)   link < 0: leaf, ~primitivesOffset; >= 0: inner node
             //                                   skip: absolute record index, LT_DONE at the end of the tree
};
static_assert(sizeof(LtThreadNode) == 32, "threaded node is 32 bytes");

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
  const LtThreadNode* tnodes;  // 8 * nodeCount threaded records, or NULL (scene too large for the copies)
};


// ------------------------------------------------------------------------------------------------
// plug-in API (compiled only for plug-ins: NVRTC, or nvcc with -DLT_PLUGIN_API)
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC_RTC__) || defined(LT_PLUGIN_API)

__constant__ LtSceneDev lt_scene;  // set by the library before each launch of a plug-in that references it

struct LtRay {
  float ox, oy, oz;  // origin
  float dx, dy, dz;  // direction (not necessarily normalised: t is in units of its length)
};

struct LtHit {
  float t, u, v;
  int primitiveIndex;  // 0 when nothing was hit, like the reference's payload
  int hitType;         // 1 = hit
};

__device__ __forceinline__ float lt_dot3z(float ax, float ay, float az, float bx, float by, float bz) {
  return __fadd_rn(__fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by))), 0.0f);
}

// intersectBounds (basic.cu:136-154) with dirIsNeg-selected bounds
__device__ __forceinline__ bool lt_slab(float lox, float hix, float loy, float hiy, float loz, float hiz, const LtRay& r,
                                        float ix, float iy, float iz) {
  float tx0 = __fmul_rn(__fsub_rn(lox, r.ox), ix), tx1 = __fmul_rn(__fsub_rn(hix, r.ox), ix);
  float ty0 = __fmul_rn(__fsub_rn(loy, r.oy), iy), ty1 = __fmul_rn(__fsub_rn(hiy, r.oy), iy);
  float tz0 = __fmul_rn(__fsub_rn(loz, r.oz), iz), tz1 = __fmul_rn(__fsub_rn(hiz, r.oz), iz);
  bool miss1 = (tx0 > ty1) || (ty0 > tx1);
  float a = (ty0 > tx0) ? ty0 : tx0;
  float b = (ty1 < tx1) ? ty1 : tx1;
  bool miss2 = (a > tz1) || (tz0 > b);
  float b2 = (tz1 < b) ? tz1 : b;
  return !miss1 && !miss2 && (b2 > 0.0f);
}

// intersectTriangle (basic.cu:93-134); epsilon 1e-7f as in basic.cu
__device__ __forceinline__ bool lt_triangle(int prim, const LtRay& r, float eps, LtHit& h) {
  const float4* tp = reinterpret_cast<const float4*>(lt_scene.tris + prim);
  float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
  float e1x = q0.w, e1y = q1.x, e1z = q1.y, e2x = q1.z, e2y = q1.w, e2z = q2.x;
  float pvx = __fmaf_rn(r.dy, e2z, -__fmul_rn(r.dz, e2y));
  float pvy = __fmaf_rn(r.dz, e2x, -__fmul_rn(r.dx, e2z));
  float pvz = __fmaf_rn(r.dx, e2y, -__fmul_rn(r.dy, e2x));
  float det = lt_dot3z(e1x, e1y, e1z, pvx, pvy, pvz);
  if (fabsf(det) < eps) return false;
  float inv = __frcp_rn(det);
  float tx = __fsub_rn(r.ox, q0.x), ty = __fsub_rn(r.oy, q0.y), tz = __fsub_rn(r.oz, q0.z);
  float u = __fmul_rn(lt_dot3z(tx, ty, tz, pvx, pvy, pvz), inv);
  if (u < 0.0f || u > 1.0f) return false;
  float qx = __fmaf_rn(ty, e1z, -__fmul_rn(tz, e1y));
  float qy = __fmaf_rn(tz, e1x, -__fmul_rn(tx, e1z));
  float qz = __fmaf_rn(tx, e1y, -__fmul_rn(ty, e1x));
  float v = __fmul_rn(lt_dot3z(r.dx, r.dy, r.dz, qx, qy, qz), inv);
  if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
  float t = __fmul_rn(lt_dot3z(e2x, e2y, e2z, qx, qy, qz), inv);
  if (t < h.t) {
    h.t = t; h.u = u; h.v = v;
    return true;
  }
  return false;
}

// The reference's intersect (ignorePrimitiveIndex < 0) / intersectIgnorePrimitiveIndex on the re-flattened scene.
// tMax is the payload's initial t (the reference passes FLT_MAX = 1e7, basic.cu:1,308); anyHit stops at the first
// accepted triangle (valid when only hitType is read, as the reference's shadow rays do).
__device__ inline LtHit lt_trace(const LtRay& r, float tMax, int ignorePrimitiveIndex = -1, bool anyHit = false,
                                 float epsilon = 1.00000001168609742e-07f) {
  LtHit h;
  h.t = tMax; h.u = 0.0f; h.v = 0.0f; h.primitiveIndex = 0; h.hitType = 0;
  float ix = __frcp_rn(r.dx), iy = __frcp_rn(r.dy), iz = __frcp_rn(r.dz);
  bool nx = ix < 0.0f, ny = iy < 0.0f, nz = iz < 0.0f;
  unsigned negMask = (nx ? 1u : 0u) | (ny ? 2u : 0u) | (nz ? 4u : 0u);
  if (!lt_slab(nx ? lt_scene.rootMax[0] : lt_scene.rootMin[0], nx ? lt_scene.rootMin[0] : lt_scene.rootMax[0],
               ny ? lt_scene.rootMax[1] : lt_scene.rootMin[1], ny ? lt_scene.rootMin[1] : lt_scene.rootMax[1],
               nz ? lt_scene.rootMax[2] : lt_scene.rootMin[2], nz ? lt_scene.rootMin[2] : lt_scene.rootMax[2], r, ix, iy, iz))
    return h;
  int stack[64];
  int sp = 0;
  int cur = lt_scene.rootRef;
  while (cur != LT_DONE) {
    if (cur >= 0) {
      const float4* np = reinterpret_cast<const float4*>(lt_scene.wnodes + cur);
      float4 bx = __ldg(np), by = __ldg(np + 1), bz = __ldg(np + 2);
      int4 m = __ldg(reinterpret_cast<const int4*>(np) + 3);
      bool hl = lt_slab(nx ? bx.y : bx.x, nx ? bx.x : bx.y, ny ? by.y : by.x, ny ? by.x : by.y, nz ? bz.y : bz.x,
                        nz ? bz.x : bz.y, r, ix, iy, iz);
      bool hr = lt_slab(nx ? bx.w : bx.z, nx ? bx.z : bx.w, ny ? by.w : by.z, ny ? by.z : by.w, nz ? bz.w : bz.z,
                        nz ? bz.z : bz.w, r, ix, iy, iz);
      bool axisNeg = (negMask >> m.z) & 1u;
      int nearRef = axisNeg ? m.y : m.x, farRef = axisNeg ? m.x : m.y;
      bool hn = axisNeg ? hr : hl, hf = axisNeg ? hl : hr;
      if (hn) {
        cur = nearRef;
        if (hf) stack[sp++] = farRef;
      } else if (hf) {
        cur = farRef;
      } else {
        cur = sp > 0 ? stack[--sp] : LT_DONE;
      }
    } else {
      int prim = ~cur;
      if (prim != ignorePrimitiveIndex && lt_triangle(prim, r, epsilon, h)) {
        h.primitiveIndex = prim;
        h.hitType = 1;
        if (anyHit) return h;
      }
      cur = sp > 0 ? stack[--sp] : LT_DONE;
    }
  }
  return h;
}

#endif  // plug-in API
