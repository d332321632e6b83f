// lt_bvh.cu -- device-side BVH construction: an LBVH (Morton order + Karras 2012
// hierarchy) emitted in the REFERENCE's flattened layout -- 32-byte LinearBVHNode in DFS order with the
// first child adjacent, ordered 76-byte Primitive array, LightContainer -- so every kernel, the CPU
// oracle and even the reference's own renderer can traverse it unchanged.
//
// This is an opt-in alternative to the host builder (lens_trace_b200/host/acceleration_structure_explicit.cpp,
// which restates the reference's median split).  The tree differs from the median-split tree, so hit ids on
// exact ties can differ; every parity test therefore compares results ON THE SAME BUFFERS.
//
//   1. centroid of each triangle's bounding box (as src/model.cpp:49-53), scene bounds (atomics on
//      order-preserving integer images of the floats)
//   2. 63-bit Morton code (21 bits per axis); key = (code, primitive index) made unique by sorting pairs with
//      a stable radix sort (cub) -- ties keep input order
//   3. Karras: internal node i covers a key range found by binary search on common-prefix lengths; its
//      children are leaves or internal nodes; parents are recorded
//   4. bottom-up: boxes and subtree shapes, second arrival proceeds (atomic flag per internal node)
//   5. DFS index of a node = 2*firstLeaf + (number of ancestors that hold it in their LEFT subtree) -- found
//      by walking the parent chain; the same walk yields the depth (stack bound)
//   6. emit LinearBVHNode[], ordered Primitive[], LightContainer; then the usual re-flatten
#include "lt_device.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

__device__ __forceinline__ unsigned orderedBits(float f) {  // monotonic float -> uint
  unsigned u = (unsigned)__float_as_int(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fromOrderedBits(unsigned u) {
  return __int_as_float((int)((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u));
}

struct BuildScratch {
  unsigned* bounds;             // [6] ordered-int min xyz, max xyz of the centroids
  unsigned long long* keys[2];  // Morton keys (double buffer for the sort)
  int* order[2];                // primitive indices (double buffer)
  int2* children;               // internal node -> (left, right); >= 0 internal, < 0 -> ~leaf position
  int* parent;                  // [0, n-1): parent of internal i; [n-1, 2n-1): parent of leaf position j at n-1+j
  int2* range;                  // internal node -> (first, last) leaf position
  float* boxes;                 // 6 floats per internal node
  int* flags;                   // arrival counters
  int* axis;                    // split axis of internal nodes (from the Morton bit that separates the children)
  int* dfsIndex;                // DFS index of internal nodes [0,n-1) and leaves [n-1, 2n-1)
  int* maxDepth;
};

__global__ void k_centroid_bounds(const RefPrim* __restrict__ prims, int n, unsigned* bounds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RefPrim& p = prims[i];
  for (int k = 0; k < 3; k++) {
    float lo = fminf(fminf(p.a[k], p.b[k]), p.c[k]), hi = fmaxf(fmaxf(p.a[k], p.b[k]), p.c[k]);
    float c = FADD(FMUL(0.5f, lo), FMUL(0.5f, hi));
    atomicMin(&bounds[k], orderedBits(c));
    atomicMax(&bounds[3 + k], orderedBits(c));
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned v) {  // 21 bits -> every third bit
  unsigned long long x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void k_morton(const RefPrim* __restrict__ prims, int n, const unsigned* __restrict__ bounds,
                         unsigned long long* keys, int* order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RefPrim& p = prims[i];
  unsigned q[3];
  for (int k = 0; k < 3; k++) {
    float lo = fminf(fminf(p.a[k], p.b[k]), p.c[k]), hi = fmaxf(fmaxf(p.a[k], p.b[k]), p.c[k]);
    float c = FADD(FMUL(0.5f, lo), FMUL(0.5f, hi));
    float mn = fromOrderedBits(bounds[k]), mx = fromOrderedBits(bounds[3 + k]);
    float ext = mx - mn;
    float u = ext > 0.0f ? (c - mn) / ext : 0.0f;
    q[k] = (unsigned)fminf(fmaxf(u * 2097152.0f, 0.0f), 2097151.0f);
  }
  keys[i] = spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
  order[i] = i;
}

// common-prefix length of the keys at sorted positions i and j; equal codes are separated by position
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

__global__ void k_karras(const unsigned long long* __restrict__ keys, int n, int2* children, int* parent,
                         int2* range, int* axis) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + min(d, 0);
  int first = min(i, j), last = max(i, j);
  int left = (first == gamma) ? ~gamma : gamma;
  int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  children[i] = make_int2(left, right);
  range[i] = make_int2(first, last);
  // the children differ first in Morton bit b; bits cycle x,y,z from the top, and the left child has the 0
  // bit, i.e. the lower coordinate on that axis -- exactly what near/far ordering by axis sign expects
  unsigned long long diff = keys[gamma] ^ keys[gamma + 1];
  int b = diff ? 63 - __clzll(diff) : 2;
  axis[i] = 2 - (b % 3);
  if (left >= 0) parent[left] = i; else parent[n - 1 + gamma] = i;
  if (right >= 0) parent[right] = i; else parent[n - 1 + gamma + 1] = i;
  if (i == 0) parent[0] = -1;
}

// bottom-up boxes: one thread per leaf climbs; the second child to arrive at a node computes it
__global__ void k_fit_boxes(const RefPrim* __restrict__ prims, const int* __restrict__ order, int n,
                            const int2* __restrict__ children, const int* __restrict__ parent, float* boxes,
                            int* flags) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int node = parent[n - 1 + j];
  while (node >= 0) {
    if (atomicAdd(&flags[node], 1) == 0) return;  // first arrival: the sibling subtree is not ready yet
    __threadfence();
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int2 ch = children[node];
    int refs[2] = {ch.x, ch.y};
    for (int c = 0; c < 2; c++) {
      if (refs[c] < 0) {
        const RefPrim& p = prims[order[~refs[c]]];
        for (int k = 0; k < 3; k++) {
          mn[k] = fminf(mn[k], fminf(fminf(p.a[k], p.b[k]), p.c[k]));
          mx[k] = fmaxf(mx[k], fmaxf(fmaxf(p.a[k], p.b[k]), p.c[k]));
        }
      } else {
        const volatile float* b = boxes + 6 * refs[c];
        for (int k = 0; k < 3; k++) {
          mn[k] = fminf(mn[k], b[k]);
          mx[k] = fmaxf(mx[k], b[3 + k]);
        }
      }
    }
    for (int k = 0; k < 3; k++) {
      boxes[6 * node + k] = mn[k];
      boxes[6 * node + 3 + k] = mx[k];
    }
    __threadfence();
    node = parent[node];
  }
}

// DFS (pre-order, first child adjacent) index = 2*firstLeaf + #ancestors holding the node in their left subtree
__global__ void k_dfs_index(int n, const int2* __restrict__ children, const int* __restrict__ parent,
                            const int2* __restrict__ range, int* dfsIndex, int* maxDepth) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;  // [0, n-1) internal, [n-1, 2n-1) leaves
  if (v >= 2 * n - 1) return;
  bool leaf = v >= n - 1;
  int first = leaf ? v - (n - 1) : range[v].x;
  int lefts = 0, depth = 0;
  int childRef = leaf ? ~(v - (n - 1)) : v;
  int node = parent[v];
  while (node >= 0) {
    if (children[node].x == childRef) lefts++;
    depth++;
    childRef = node;
    node = parent[node];
  }
  dfsIndex[v] = 2 * first + lefts;
  if (leaf) atomicMax(maxDepth, depth);  // inner nodes above a leaf = stack bound
}

__global__ void k_emit_nodes(const RefPrim* __restrict__ prims, const int* __restrict__ order, int n,
                             const int2* __restrict__ children, const float* __restrict__ boxes,
                             const int* __restrict__ axisArr, const int* __restrict__ dfsIndex, RefNode* nodes,
                             RefPrim* orderedPrims) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= 2 * n - 1) return;
  RefNode out;
  if (v >= n - 1) {  // leaf at sorted position j
    int j = v - (n - 1);
    RefPrim p = prims[order[j]];
    orderedPrims[j] = p;
    for (int k = 0; k < 3; k++) {
      out.boundsMin[k] = fminf(fminf(p.a[k], p.b[k]), p.c[k]);
      out.boundsMax[k] = fmaxf(fmaxf(p.a[k], p.b[k]), p.c[k]);
    }
    out.offset = j;
    out.primitiveCount = 1;
    out.axis = 0;
  } else {
    for (int k = 0; k < 3; k++) {
      out.boundsMin[k] = boxes[6 * v + k];
      out.boundsMax[k] = boxes[6 * v + 3 + k];
    }
    int2 ch = children[v];
    int rightV = ch.y < 0 ? n - 1 + ~ch.y : ch.y;
    out.offset = dfsIndex[rightV];
    out.primitiveCount = 0;
    out.axis = (uint8_t)axisArr[v];
  }
  out.pad = 0;
  nodes[dfsIndex[v]] = out;
}

// lights in leaf order, first 64 (single thread: the list is tiny and must be ordered)
__global__ void k_collect_lights(const RefPrim* __restrict__ orderedPrims, int n, const RefMaterial* __restrict__ mats,
                                 int matCount, RefLights* lights) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  RefLights out;
  out.count = 0;
  for (int k = 0; k < 64; k++) out.primitives[k] = 0;
  for (int i = 0; i < n && out.count < 64; i++) {
    int m = orderedPrims[i].materialIndex;
    if (m >= 0 && m < matCount) {
      const RefMaterial& mat = mats[m];
      if (mat.emission[0] > 0 || mat.emission[1] > 0 || mat.emission[2] > 0) out.primitives[out.count++] = (unsigned)i;
    }
  }
  *lights = out;
}

// parallel variant for large scenes: flag emissive primitives, the host keeps the first 64 in order
__global__ void k_flag_emissive(const RefPrim* __restrict__ orderedPrims, int n, const RefMaterial* __restrict__ mats,
                                int matCount, int* counter, int* firstEmissive /* up to 4096 indices */) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int m = orderedPrims[i].materialIndex;
  if (m < 0 || m >= matCount) return;
  const RefMaterial& mat = mats[m];
  if (mat.emission[0] > 0 || mat.emission[1] > 0 || mat.emission[2] > 0) {
    int slot = atomicAdd(counter, 1);
    if (slot < 4096) firstEmissive[slot] = i;
  }
}

}  // namespace

size_t lt_bvh_scratch_bytes(int n) {
  size_t sortBytes = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DoubleBuffer<int> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, k, v, n, 0, 63);
  size_t per = sizeof(unsigned long long) * 2 + sizeof(int) * 2 + sizeof(int2) * 2 + sizeof(int) * 2 + 6 * sizeof(float) +
               sizeof(int) * 2 + sizeof(int) * 2;
  return sortBytes + per * (size_t)n + 4096 * sizeof(int) + 256 * 32;
}

// Builds nodes / ordered primitives / lights (all device buffers in the reference layouts).  Returns the
// number of kernels launched, or -1 on a CUDA error.  *outMaxDepth receives the stack bound.
int lt_launch_bvh_build(const RefPrim* dPrims, int n, const RefMaterial* dMats, int matCount, RefNode* dNodes,
                        RefPrim* dOrderedPrims, RefLights* dLights, void* scratch, size_t scratchBytes, int* outMaxDepth,
                        cudaStream_t stream) {
  if (n < 2) return -1;  // a single triangle has no hierarchy: callers use lt_scene_upload for that
  char* p = (char*)scratch;
  auto take = [&](size_t bytes) {
    char* r = p;
    p += (bytes + 255) & ~(size_t)255;
    return r;
  };
  BuildScratch S;
  S.bounds = (unsigned*)take(6 * sizeof(unsigned));
  S.maxDepth = (int*)take(sizeof(int));
  int* emissiveCount = (int*)take(sizeof(int));
  int* emissive = (int*)take(4096 * sizeof(int));
  S.keys[0] = (unsigned long long*)take(sizeof(unsigned long long) * n);
  S.keys[1] = (unsigned long long*)take(sizeof(unsigned long long) * n);
  S.order[0] = (int*)take(sizeof(int) * n);
  S.order[1] = (int*)take(sizeof(int) * n);
  S.children = (int2*)take(sizeof(int2) * n);
  S.range = (int2*)take(sizeof(int2) * n);
  S.parent = (int*)take(sizeof(int) * 2 * n);
  S.boxes = (float*)take(sizeof(float) * 6 * n);
  S.flags = (int*)take(sizeof(int) * n);
  S.axis = (int*)take(sizeof(int) * n);
  S.dfsIndex = (int*)take(sizeof(int) * 2 * n);
  size_t sortBytes = 0;
  cub::DoubleBuffer<unsigned long long> kb(S.keys[0], S.keys[1]);
  cub::DoubleBuffer<int> vb(S.order[0], S.order[1]);
  cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, kb, vb, n, 0, 63);
  void* sortTemp = take(sortBytes);
  if ((size_t)(p - (char*)scratch) > scratchBytes) return -1;

  const int T = 256;
  const unsigned initBounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  cudaMemcpyAsync(S.bounds, initBounds, sizeof initBounds, cudaMemcpyHostToDevice, stream);
  cudaMemsetAsync(S.maxDepth, 0, sizeof(int), stream);
  cudaMemsetAsync(emissiveCount, 0, sizeof(int), stream);
  cudaMemsetAsync(S.flags, 0, sizeof(int) * n, stream);
  k_centroid_bounds<<<(n + T - 1) / T, T, 0, stream>>>(dPrims, n, S.bounds);
  k_morton<<<(n + T - 1) / T, T, 0, stream>>>(dPrims, n, S.bounds, S.keys[0], S.order[0]);
  cub::DeviceRadixSort::SortPairs(sortTemp, sortBytes, kb, vb, n, 0, 63, stream);
  const unsigned long long* keys = kb.Current();
  const int* order = vb.Current();
  k_karras<<<(n - 1 + T - 1) / T, T, 0, stream>>>(keys, n, S.children, S.parent, S.range, S.axis);
  k_fit_boxes<<<(n + T - 1) / T, T, 0, stream>>>(dPrims, order, n, S.children, S.parent, S.boxes, S.flags);
  k_dfs_index<<<(2 * n - 1 + T - 1) / T, T, 0, stream>>>(n, S.children, S.parent, S.range, S.dfsIndex, S.maxDepth);
  k_emit_nodes<<<(2 * n - 1 + T - 1) / T, T, 0, stream>>>(dPrims, order, n, S.children, S.boxes, S.axis, S.dfsIndex,
                                                          dNodes, dOrderedPrims);
  k_collect_lights<<<1, 1, 0, stream>>>(dOrderedPrims, n < 65536 ? n : 0, dMats, matCount, dLights);
  int launches = 8;
  if (n >= 65536) {  // large scenes: parallel flagging, ordered selection on the host
    k_flag_emissive<<<(n + T - 1) / T, T, 0, stream>>>(dOrderedPrims, n, dMats, matCount, emissiveCount, emissive);
    launches++;
    int count = 0;
    cudaMemcpyAsync(&count, emissiveCount, sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    if (count > 4096) count = 4096;
    std::vector<int> idx(count);
    if (count) cudaMemcpy(idx.data(), emissive, sizeof(int) * count, cudaMemcpyDeviceToHost);
    std::sort(idx.begin(), idx.end());
    RefLights L;
    memset(&L, 0, sizeof L);
    for (int k = 0; k < count && L.count < 64; k++) L.primitives[L.count++] = (unsigned)idx[k];
    cudaMemcpyAsync(dLights, &L, sizeof L, cudaMemcpyHostToDevice, stream);
  }
  if (outMaxDepth) {
    cudaMemcpyAsync(outMaxDepth, S.maxDepth, sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}
