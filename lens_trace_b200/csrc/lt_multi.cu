// lt_multi.cu -- multi-GPU contexts of the C-ABI (lt_ctx_create_multi): one process, one single-GPU context per
// device, the scene replicated, one lt_render call divided among the devices (the reference renders
// on device 0 only, src/cuda/renderer_cuda.cpp:12).
//
//   sample split  device g renders frames g, g+G, ... of the call into its own FP32 accumulator as a weighted sum
//                 (LT_ACCUM_WEIGHTED_SUM, weight 1/N); the accumulators are combined by ONE all-reduce per call:
//                 ncclAllReduce over NVLink (single-process communicators, ncclCommInitAll; libnccl is loaded with
//                 dlopen on first use), or -- LT_MULTI_EXCHANGE=p2p, and whenever NCCL is unavailable -- this
//                 file's own kernel, in which every device sums all peers' accumulators through peer-mapped memory
//                 in a fixed order (deterministic) behind an event barrier.  The result differs from the sequential
//                 running mean of accumulator.frag:10-19 by FP32 summation order only.
//   tile split    device g renders every G-th block of 8 image rows (cost varies smoothly over the image, so the
//                 blocks balance); nothing is exchanged, every device copies its rows into the caller's buffer over
//                 its own PCIe link.
//
// The kernels of every device are launched from that device's own host thread: a 64-spp wavefront step is ~240
// launches, and eight devices' worth of them issued from one thread would cost more than the rendering.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "lt_internal.h"

#define LT_TILE_ROWS 8

namespace {

struct NcclApi {
  bool tried = false, ok = false;
  std::string why;
  ncclResult_t (*commInitAll)(ncclComm_t*, int, const int*);
  ncclResult_t (*commDestroy)(ncclComm_t);
  ncclResult_t (*allReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*groupStart)();
  ncclResult_t (*groupEnd)();
  const char* (*getErrorString)(ncclResult_t);
};
NcclApi g_nccl;

bool load_nccl() {
  NcclApi& a = g_nccl;
  if (a.tried) return a.ok;
  a.tried = true;
  void* lib = nullptr;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names)
    if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!lib) {
    a.why = "libnccl.so.2 not found";
    return false;
  }
  a.commInitAll = (decltype(a.commInitAll))dlsym(lib, "ncclCommInitAll");
  a.commDestroy = (decltype(a.commDestroy))dlsym(lib, "ncclCommDestroy");
  a.allReduce = (decltype(a.allReduce))dlsym(lib, "ncclAllReduce");
  a.groupStart = (decltype(a.groupStart))dlsym(lib, "ncclGroupStart");
  a.groupEnd = (decltype(a.groupEnd))dlsym(lib, "ncclGroupEnd");
  a.getErrorString = (decltype(a.getErrorString))dlsym(lib, "ncclGetErrorString");
  a.ok = a.commInitAll && a.commDestroy && a.allReduce && a.groupStart && a.groupEnd && a.getErrorString;
  if (!a.ok) a.why = "libnccl lacks a needed symbol";
  return a.ok;
}

#define LT_MAX_DEVICES 16
struct PeerPointers {
  const float* p[LT_MAX_DEVICES];
};

// dst[i] = sum over the devices, in device order, of their accumulators (peer-mapped reads over NVLink)
__global__ void k_sum_peers(float* __restrict__ dst, PeerPointers src, int devices, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 s = reinterpret_cast<const float4*>(src.p[0])[i];
    for (int d = 1; d < devices; d++) {
      float4 v = reinterpret_cast<const float4*>(src.p[d])[i];
      s.x = __fadd_rn(s.x, v.x);
      s.y = __fadd_rn(s.y, v.y);
      s.z = __fadd_rn(s.z, v.z);
      s.w = __fadd_rn(s.w, v.w);
    }
    reinterpret_cast<float4*>(dst)[i] = s;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = src.p[0][i];
    for (int d = 1; d < devices; d++) s = __fadd_rn(s, src.p[d][i]);
    dst[i] = s;
  }
}

__global__ void k_scale(float* __restrict__ p, size_t n, float f) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = __fmul_rn(p[i], f);
}

}  // namespace

struct LtGroup {
  std::vector<lt_ctx*> ctx;         // one single-GPU context per device
  std::vector<ncclComm_t> comms;    // sample split, NCCL exchange
  bool ncclTried = false, ncclReady = false;
  bool peersTried = false, peersReady = false;
  std::vector<float*> sumBuf;       // p2p exchange: second accumulator per device (the kernel's destination)
  size_t sumFloats = 0;
  std::vector<cudaEvent_t> rendered, summed;
  int lastSplit = -1;               // split of the accumulators the devices hold (a change resets them)
  long long lastFloats = -1;
};

extern "C" int lt_ctx_device_count(const lt_ctx* ctx) {
  if (!ctx) return 0;
  return ctx->group ? (int)ctx->group->ctx.size() : 1;
}

extern "C" int lt_ctx_create_multi(const int* device_ordinals, int device_count, lt_ctx** out_ctx) {
  if (!out_ctx) return lt_internal_fail(nullptr, LT_ERR_INVALID, "lt_ctx_create_multi: out_ctx is NULL");
  *out_ctx = nullptr;
  if (!device_ordinals || device_count < 1 || device_count > LT_MAX_DEVICES)
    return lt_internal_fail(nullptr, LT_ERR_INVALID, "lt_ctx_create_multi: needs 1..16 device ordinals");
  for (int a = 0; a < device_count; a++)
    for (int b = a + 1; b < device_count; b++)
      if (device_ordinals[a] == device_ordinals[b])
        return lt_internal_fail(nullptr, LT_ERR_INVALID, "lt_ctx_create_multi: a device ordinal is listed twice");
  if (device_count == 1) return lt_ctx_create(device_ordinals[0], out_ctx);
  LtGroup* g = new LtGroup();
  for (int k = 0; k < device_count; k++) {
    lt_ctx* c = nullptr;
    int rc = lt_ctx_create(device_ordinals[k], &c);
    if (rc != LT_OK) {
      for (lt_ctx* d : g->ctx) lt_ctx_destroy(d);
      delete g;
      return rc;  // message already set by lt_ctx_create
    }
    g->ctx.push_back(c);
  }
  g->rendered.resize(device_count, nullptr);
  g->summed.resize(device_count, nullptr);
  for (int k = 0; k < device_count; k++) {
    cudaSetDevice(g->ctx[k]->device);
    cudaEventCreateWithFlags(&g->rendered[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->summed[k], cudaEventDisableTiming);
  }
  lt_ctx* ctx = new lt_ctx();
  ctx->device = device_ordinals[0];
  ctx->group = g;
  ctx->stats.sm_count = g->ctx[0]->stats.sm_count;
  *out_ctx = ctx;
  return LT_OK;
}

void lt_multi_destroy(lt_ctx* ctx) {
  LtGroup* g = ctx->group;
  if (!g) return;
  for (size_t k = 0; k < g->ctx.size(); k++) {
    cudaSetDevice(g->ctx[k]->device);
    cudaDeviceSynchronize();
    if (k < g->comms.size() && g->comms[k] && g_nccl.ok) g_nccl.commDestroy(g->comms[k]);
    if (k < g->sumBuf.size() && g->sumBuf[k]) cudaFree(g->sumBuf[k]);
    if (g->rendered[k]) cudaEventDestroy(g->rendered[k]);
    if (g->summed[k]) cudaEventDestroy(g->summed[k]);
  }
  for (lt_ctx* c : g->ctx) lt_ctx_destroy(c);
  delete g;
  ctx->group = nullptr;
}

int lt_multi_scene_upload(lt_ctx* ctx, const void* nodes, uint64_t node_bytes, const void* primitives,
                          uint64_t primitive_bytes, const void* materials, uint64_t material_bytes,
                          const void* light_container, uint64_t light_bytes, lt_scene** out_scene) {
  LtGroup* g = ctx->group;
  *out_scene = nullptr;
  lt_scene* s = new lt_scene();
  s->parts.resize(g->ctx.size(), nullptr);
  std::vector<int> rcs(g->ctx.size(), LT_OK);
  std::vector<std::thread> workers;
  for (size_t k = 0; k < g->ctx.size(); k++)
    workers.emplace_back([&, k]() {
      rcs[k] = lt_scene_upload(g->ctx[k], nodes, node_bytes, primitives, primitive_bytes, materials, material_bytes,
                               light_container, light_bytes, &s->parts[k]);
    });
  for (std::thread& t : workers) t.join();
  float ms = 0.0f;
  for (size_t k = 0; k < g->ctx.size(); k++)
    if (g->ctx[k]->stats.upload_ms > ms) ms = g->ctx[k]->stats.upload_ms;
  ctx->stats.upload_ms = ms;
  for (size_t k = 0; k < g->ctx.size(); k++)
    if (rcs[k] != LT_OK) {
      std::string why = lt_last_error(g->ctx[k]);
      for (size_t j = 0; j < g->ctx.size(); j++)
        if (s->parts[j]) lt_scene_release(g->ctx[j], s->parts[j]);
      delete s;
      return lt_internal_fail(ctx, rcs[k], why);
    }
  *out_scene = s;
  return LT_OK;
}

void lt_multi_scene_release(lt_ctx* ctx, lt_scene* scene) {
  LtGroup* g = ctx ? ctx->group : nullptr;
  for (size_t k = 0; k < scene->parts.size(); k++)
    if (scene->parts[k]) lt_scene_release(g && k < g->ctx.size() ? g->ctx[k] : nullptr, scene->parts[k]);
  delete scene;
}

int lt_multi_accum_reset(lt_ctx* ctx) {
  LtGroup* g = ctx->group;
  for (lt_ctx* c : g->ctx) {
    int rc = lt_accum_reset(c);
    if (rc != LT_OK) return lt_internal_fail(ctx, rc, lt_last_error(c));
  }
  return LT_OK;
}

static bool ensure_nccl(LtGroup* g, std::string* why) {
  if (g->ncclTried) return g->ncclReady;
  g->ncclTried = true;
  if (!load_nccl()) {
    *why = g_nccl.why;
    return false;
  }
  std::vector<int> devs;
  for (lt_ctx* c : g->ctx) devs.push_back(c->device);
  g->comms.assign(devs.size(), nullptr);
  ncclResult_t r = g_nccl.commInitAll(g->comms.data(), (int)devs.size(), devs.data());
  if (r != ncclSuccess) {
    *why = std::string("ncclCommInitAll: ") + g_nccl.getErrorString(r);
    g->comms.clear();
    return false;
  }
  g->ncclReady = true;
  return true;
}

static bool ensure_peers(LtGroup* g, size_t floats, std::string* why) {
  if (!g->peersTried) {
    g->peersTried = true;
    g->peersReady = true;
    for (size_t a = 0; a < g->ctx.size() && g->peersReady; a++) {
      cudaSetDevice(g->ctx[a]->device);
      for (size_t b = 0; b < g->ctx.size(); b++) {
        if (a == b) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, g->ctx[a]->device, g->ctx[b]->device);
        if (!can) {
          *why = "devices cannot access each other's memory (no peer access)";
          g->peersReady = false;
          break;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[b]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          *why = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
          g->peersReady = false;
          break;
        }
        cudaGetLastError();
      }
    }
  }
  if (!g->peersReady) {
    if (why->empty()) *why = "peer access is unavailable";
    return false;
  }
  if (g->sumFloats < floats) {
    g->sumBuf.resize(g->ctx.size(), nullptr);
    for (size_t k = 0; k < g->ctx.size(); k++) {
      cudaSetDevice(g->ctx[k]->device);
      if (g->sumBuf[k]) cudaFree(g->sumBuf[k]);
      g->sumBuf[k] = nullptr;
      if (cudaMalloc(&g->sumBuf[k], floats * sizeof(float)) != cudaSuccess) {
        *why = "cannot allocate the exchange buffer";
        g->sumFloats = 0;
        return false;
      }
    }
    g->sumFloats = floats;
  }
  return true;
}

// rows of the image device g of G renders in a tile split: every G-th block of LT_TILE_ROWS rows
static int tile_rows_of(int height, int G, int g) {
  int rows = 0;
  for (int b = g; b * LT_TILE_ROWS < height; b += G) {
    int left = height - b * LT_TILE_ROWS;
    rows += left < LT_TILE_ROWS ? left : LT_TILE_ROWS;
  }
  return rows;
}

// Host-only (no device needed): the image rows device `device` of `devices` renders in a tile split, in the order of
// its local rows -- the mapping the kernels apply (lt_device.cuh: lt_image_row) to the row count lt_multi_render hands
// them.  Exposed so that the partition can be checked on a machine without a GPU.
extern "C" int lt_debug_tile_rows(int height, int devices, int device, int32_t* out_rows, int capacity) {
  if (height < 1 || devices < 1 || device < 0 || device >= devices) return LT_ERR_INVALID;
  const int rows = tile_rows_of(height, devices, device);
  if (out_rows) {
    if (capacity < rows) return LT_ERR_INVALID;
    for (int j = 0; j < rows; j++) {
      const int b = j / LT_TILE_ROWS;
      out_rows[j] = (b * devices + device) * LT_TILE_ROWS + (j - b * LT_TILE_ROWS);
    }
  }
  return rows;
}

int lt_multi_render(lt_ctx* ctx, lt_scene* scene, const void* camera28, const lt_render_params* params, float* host_out) {
  LtGroup* g = ctx->group;
  if (!scene || !camera28 || !params) return lt_internal_fail(ctx, LT_ERR_INVALID, "lt_render: NULL argument");
  if (scene->parts.size() != g->ctx.size())
    return lt_internal_fail(ctx, LT_ERR_INVALID, "lt_render: the scene was not uploaded through this multi-GPU context");
  if (params->struct_size != sizeof(lt_render_params) && params->struct_size != offsetof(lt_render_params, split_mode))
    return lt_internal_fail(ctx, LT_ERR_INVALID, "lt_render: params.struct_size mismatch");
  lt_render_params P;
  memset(&P, 0, sizeof P);
  memcpy(&P, params, params->struct_size < sizeof P ? params->struct_size : sizeof P);
  P.struct_size = sizeof P;
  if (P.width <= 0 || P.height <= 0 || P.depth < 3 || P.frames < 1)
    return lt_internal_fail(ctx, LT_ERR_INVALID, "lt_render: width/height must be > 0, depth >= 3, frames >= 1");
  if (P.flags & LT_FLAG_STATS)
    return lt_internal_fail(ctx, LT_ERR_UNSUPPORTED, "lt_render: LT_FLAG_STATS counts on a single-GPU context");
  const int G = (int)g->ctx.size();
  int split = P.split_mode;
  if (split == LT_SPLIT_AUTO) split = (P.frames >= 2 && P.accum_mode != LT_ACCUM_NONE) ? LT_SPLIT_SAMPLES : LT_SPLIT_TILES;
  if (split == LT_SPLIT_SAMPLES && P.accum_mode == LT_ACCUM_NONE)
    return lt_internal_fail(ctx, LT_ERR_UNSUPPORTED,
                            "lt_render: a sample split needs an accumulating call (LT_ACCUM_NONE keeps only the last frame)");
  if (split != LT_SPLIT_SAMPLES && split != LT_SPLIT_TILES) return lt_internal_fail(ctx, LT_ERR_INVALID, "lt_render: bad split_mode");
  RefCamera cam;
  memcpy(&cam, camera28, sizeof cam);
  const size_t fullFloats = (size_t)P.width * P.height * P.depth;
  // the devices' accumulators are only comparable while split and size stay the same
  if (g->lastSplit != split || g->lastFloats != (long long)fullFloats) {
    int rc = lt_multi_accum_reset(ctx);
    if (rc != LT_OK) return rc;
    g->lastSplit = split;
    g->lastFloats = (long long)fullFloats;
  }
  std::vector<int> rcs(G, LT_OK);
  std::vector<std::thread> workers;
  const uint32_t stride = P.frame_stride ? P.frame_stride : 1u;

  if (split == LT_SPLIT_TILES) {
    std::vector<int> rows(G);
    for (int k = 0; k < G; k++) rows[k] = tile_rows_of(P.height, G, k);
    const size_t rowBytes = (size_t)P.width * P.depth * sizeof(float);
    for (int k = 0; k < G; k++)
      workers.emplace_back([&, k]() {
        lt_ctx* c = g->ctx[k];
        if (rows[k] == 0) return;
        if (cudaSetDevice(c->device) != cudaSuccess) { rcs[k] = LT_ERR_CUDA; return; }
        rcs[k] = lt_internal_ensure_out(c, (size_t)P.width * rows[k] * P.depth);
        if (rcs[k] != LT_OK) return;
        lt_render_params Pk = P;
        Pk.height = rows[k];
        Pk.split_mode = 0;
        rcs[k] = lt_internal_render_rows(c, scene->parts[k], &cam, &Pk, c->dOut, P.height, LT_TILE_ROWS, G, k, 0);
        if (rcs[k] != LT_OK || !host_out) return;
        // blocks k, k+G, ... of the image: full blocks in one strided copy, a shorter last block on its own
        const int fullBlocks = rows[k] / LT_TILE_ROWS, tail = rows[k] % LT_TILE_ROWS;
        char* dst = (char*)host_out + (size_t)k * LT_TILE_ROWS * rowBytes;
        cudaError_t e = cudaSuccess;
        if (fullBlocks > 0)
          e = cudaMemcpy2DAsync(dst, (size_t)G * LT_TILE_ROWS * rowBytes, c->dOut, LT_TILE_ROWS * rowBytes,
                                LT_TILE_ROWS * rowBytes, fullBlocks, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && tail > 0)
          e = cudaMemcpyAsync(dst + (size_t)fullBlocks * G * LT_TILE_ROWS * rowBytes,
                              (const char*)c->dOut + (size_t)fullBlocks * LT_TILE_ROWS * rowBytes, tail * rowBytes,
                              cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rcs[k] = lt_internal_fail(c, LT_ERR_CUDA, std::string("tile copy: ") + cudaGetErrorString(e));
      });
    for (std::thread& t : workers) t.join();
    if (!host_out)
      for (int k = 0; k < G; k++) {
        cudaSetDevice(g->ctx[k]->device);
        cudaStreamSynchronize(g->ctx[k]->stream);
      }
  } else {
    // ---- sample split ----
    if (P.accum_mode == LT_ACCUM_RUNNING_MEAN && stride != 1u)
      return lt_internal_fail(ctx, LT_ERR_UNSUPPORTED, "lt_render: a multi-GPU running mean needs frame_stride 1");
    const bool mean = P.accum_mode == LT_ACCUM_RUNNING_MEAN;
    // running mean: new = old * fc/(fc+F) + sum(samples)/(fc+F)  (fc = frames already in the accumulator; 0 discards it,
    // like frameCount == 0 in accumulator.frag:10-19)
    const double total = (double)cam.frameCount + (double)P.frames;
    const float weight = mean ? (float)(1.0 / total) : P.accum_weight;
    const float keep = mean ? (float)((double)cam.frameCount / total) : 1.0f;
    std::string why;
    const char* ex = getenv("LT_MULTI_EXCHANGE");
    bool useNccl = !(ex && strcmp(ex, "p2p") == 0) && ensure_nccl(g, &why);
    if (!useNccl && !ensure_peers(g, fullFloats, &why))
      return lt_internal_fail(ctx, LT_ERR_UNSUPPORTED, "lt_render: no way to combine the devices' accumulators: " + why);
    for (int k = 0; k < G; k++)
      workers.emplace_back([&, k]() {
        lt_ctx* c = g->ctx[k];
        if (cudaSetDevice(c->device) != cudaSuccess) { rcs[k] = LT_ERR_CUDA; return; }
        rcs[k] = lt_internal_ensure_out(c, fullFloats);
        if (rcs[k] != LT_OK) return;
        // device 0 carries the accumulator that was there before this call, the others start from zero
        if (k == 0) {
          if (keep == 0.0f) cudaMemsetAsync(c->dOut, 0, fullFloats * sizeof(float), c->stream);
          else if (keep != 1.0f) k_scale<<<c->stats.sm_count * 8, 256, 0, c->stream>>>(c->dOut, fullFloats, keep);
        } else {
          cudaMemsetAsync(c->dOut, 0, fullFloats * sizeof(float), c->stream);
        }
        const int nk = (P.frames - k + G - 1) / G;  // frames k, k+G, ... < P.frames
        if (nk > 0 && k < P.frames) {
          lt_render_params Pk = P;
          Pk.frames = nk;
          Pk.frame_stride = stride * (uint32_t)G;
          Pk.accum_mode = LT_ACCUM_WEIGHTED_SUM;
          Pk.accum_weight = weight;
          Pk.split_mode = 0;
          RefCamera ck = cam;
          ck.frameCount = cam.frameCount + (uint32_t)k * stride;
          rcs[k] = lt_internal_render_rows(c, scene->parts[k], &ck, &Pk, c->dOut, P.height, P.height, 1, 0, 0);
        }
        cudaEventRecord(g->rendered[k], c->stream);
      });
    for (std::thread& t : workers) t.join();
    for (int k = 0; k < G; k++)
      if (rcs[k] != LT_OK) return lt_internal_fail(ctx, rcs[k], lt_last_error(g->ctx[k]));
    // ---- the exchange step: one all-reduce(sum) of W*H*depth floats ----
    if (useNccl) {
      g_nccl.groupStart();
      for (int k = 0; k < G; k++) {
        lt_ctx* c = g->ctx[k];
        ncclResult_t r = g_nccl.allReduce(c->dOut, c->dOut, fullFloats, ncclFloat, ncclSum, g->comms[k], c->stream);
        if (r != ncclSuccess) {
          g_nccl.groupEnd();
          return lt_internal_fail(ctx, LT_ERR_CUDA, std::string("ncclAllReduce: ") + g_nccl.getErrorString(r));
        }
      }
      ncclResult_t r = g_nccl.groupEnd();
      if (r != ncclSuccess) return lt_internal_fail(ctx, LT_ERR_CUDA, std::string("ncclGroupEnd: ") + g_nccl.getErrorString(r));
    } else {
      PeerPointers src;
      for (int k = 0; k < G; k++) src.p[k] = g->ctx[k]->dOut;
      for (int k = 0; k < G; k++) {
        lt_ctx* c = g->ctx[k];
        cudaSetDevice(c->device);
        for (int j = 0; j < G; j++)
          if (j != k) cudaStreamWaitEvent(c->stream, g->rendered[j], 0);  // every device has finished rendering
        k_sum_peers<<<c->stats.sm_count * 4, 256, 0, c->stream>>>(g->sumBuf[k], src, G, fullFloats);
        cudaEventRecord(g->summed[k], c->stream);
      }
      for (int k = 0; k < G; k++) {  // nobody overwrites its accumulator while a peer still reads it
        lt_ctx* c = g->ctx[k];
        cudaSetDevice(c->device);
        for (int j = 0; j < G; j++)
          if (j != k) cudaStreamWaitEvent(c->stream, g->summed[j], 0);
        cudaMemcpyAsync(c->dOut, g->sumBuf[k], fullFloats * sizeof(float), cudaMemcpyDeviceToDevice, c->stream);
      }
    }
    lt_ctx* c0 = g->ctx[0];
    cudaSetDevice(c0->device);
    if (host_out) {
      int rc = lt_internal_download(c0, host_out, c0->dOut, fullFloats * sizeof(float));
      if (rc != LT_OK) rcs[0] = rc;
    }
    for (int k = 0; k < G; k++) {
      cudaSetDevice(g->ctx[k]->device);
      cudaError_t e = cudaStreamSynchronize(g->ctx[k]->stream);
      if (e != cudaSuccess) rcs[k] = lt_internal_fail(g->ctx[k], LT_ERR_CUDA, std::string("multi-GPU render: ") + cudaGetErrorString(e));
    }
  }
  for (int k = 0; k < G; k++)
    if (rcs[k] != LT_OK) return lt_internal_fail(ctx, rcs[k], lt_last_error(g->ctx[k]));
  // statistics of the call: the slowest device's kernels, all devices' launches
  ctx->stats.kernel_ms = 0.0f;
  ctx->stats.kernel_launches = 0;
  for (int k = 0; k < G; k++) {
    lt_ctx* c = g->ctx[k];
    float ms = 0.0f;
    cudaSetDevice(c->device);
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess && ms > ctx->stats.kernel_ms) ctx->stats.kernel_ms = ms;
    cudaGetLastError();
    ctx->stats.kernel_launches += c->stats.kernel_launches;
  }
  cudaSetDevice(g->ctx[0]->device);
  return LT_OK;
}
