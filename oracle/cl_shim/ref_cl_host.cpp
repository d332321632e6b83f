/*
 * ref_cl_host.cpp -- runs the reference's own OpenCL kernel files on the host CPU, one work-item
 * at a time, through opencl_c_on_cpp.h.  TEST INFRASTRUCTURE ONLY (see oracle/lt_oracle.h).
 *
 * Built by oracle/build_ref_cl.sh into oracle/_ref/libltref_cl.so; the six `*.inc` files included
 * below are produced by that script from /root/reference (vector-literal rewrite only) and live
 * in oracle/_ref/cl/ -- they are never committed.
 *
 * The launch loop restates the reference's launcher, src/opencl/renderer_opencl.cpp:84-146:
 *   workBlockCount = (W / workBlockSize[0]) * (H / workBlockSize[1])        (integer division)
 *   for x in 0..workBlockCount: NDRange(global = workBlockSize, local = implementation's choice)
 *     with args (nodes, primitives, materials, lightContainer, camera, output, x, W, H, depth)
 * The local size (which only tileKernel can observe) is a parameter here.
 */
#include "opencl_c_on_cpp.h"

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

thread_local clshim_work_item clshim_wi;

/* global_illumination.cl hard-codes `int maxRayDepth = 16;` (resources/...:308, examples/...:307).
 * The workloads of BASELINE.json use other caps, so build_ref_cl.sh also emits a second copy of the
 * two GI files in which that one initialiser reads this variable instead. */
static thread_local int clshim_max_ray_depth = 16;

namespace k_basic {
#include "cl/basic.inc"
}
namespace k_basic_lighting {
#include "cl/basic_lighting.inc"
}
namespace k_global_illumination25 {
#include "cl/global_illumination_resources.inc"
}
namespace k_accumulator {
#include "cl/accumulator.inc"
}
namespace k_custom {
#include "cl/custom_opencl.inc"
}
namespace k_global_illumination {
#include "cl/global_illumination_example.inc"
}
namespace k_global_illumination25_depth {
#include "cl/global_illumination_resources_depth.inc"
}
namespace k_global_illumination_depth {
#include "cl/global_illumination_example_depth.inc"
}

namespace {

/* kernel ids = oracle/lt_oracle.h's LTO_KERNEL_* (1..6; 0 is basic.cu, which is CUDA and runs for real) */
struct launch {
  int kernel, mode;
  void *nodes, *prims, *mats, *lights, *cam;
  float* out;
  uint W, H, depth;
  size_t wbs[2], ls[2];
  int maxRayDepth;
};

#define CALL(NS)                                                                                  \
  do {                                                                                            \
    auto n = (NS::LinearBVHNode*)L.nodes;                                                         \
    auto p = (NS::Primitive*)L.prims;                                                             \
    auto m = (NS::Material*)L.mats;                                                               \
    auto l = (NS::LightContainer*)L.lights;                                                       \
    auto c = (NS::Camera*)L.cam;                                                                  \
    if (L.mode == 0) NS::linearKernel(n, p, m, l, c, L.out, block, L.W, L.H, L.depth);            \
    else NS::tileKernel(n, p, m, l, c, L.out, block, L.W, L.H, L.depth);                          \
  } while (0)

void run_item(const launch& L, uint block) {
  bool custom = L.maxRayDepth != 16;
  switch (L.kernel) {
    case 1: CALL(k_basic); break;
    case 2: CALL(k_custom); break;
    case 3: CALL(k_basic_lighting); break;
    case 4: CALL(k_accumulator); break;
    case 5: if (custom) CALL(k_global_illumination25_depth); else CALL(k_global_illumination25); break;
    case 6: if (custom) CALL(k_global_illumination_depth); else CALL(k_global_illumination); break;
  }
}

}  // namespace

extern "C" {

/* sizes of the kernel files' own struct declarations, so the caller can check its buffers match */
void ltrefcl_struct_sizes(int out[5]) {
  out[0] = (int)sizeof(k_global_illumination::LinearBVHNode);
  out[1] = (int)sizeof(k_global_illumination::Primitive);
  out[2] = (int)sizeof(k_global_illumination::Material);
  out[3] = (int)sizeof(k_global_illumination::LightContainer);
  out[4] = (int)sizeof(k_global_illumination::Camera);
}

/* One RendererOpenCL::render() worth of launches of kernel file `kernel` (LTO_KERNEL_* id, 1..6).
 * mode 0 = linearKernel, 1 = tileKernel.  workBlock = the NDRange global size, local = work-group
 * size (must divide workBlock).  maxRayDepth 16 = the files as shipped.  Returns 0 or -1. */
int ltrefcl_render(int kernel, int mode, void* nodes, void* prims, void* mats, void* lights, void* cam,
                   float* out, int W, int H, int depth, int workBlockX, int workBlockY, int localX, int localY,
                   int maxRayDepth, int threads) {
  if (kernel < 1 || kernel > 6 || (mode != 0 && mode != 1) || W <= 0 || H <= 0 || depth < 3) return -1;
  if (workBlockX <= 0 || workBlockY <= 0 || localX <= 0 || localY <= 0) return -1;
  if (workBlockX % localX || workBlockY % localY) return -1;
  if (maxRayDepth != 16 && kernel != 5 && kernel != 6) return -1;
  launch L{kernel, mode, nodes, prims, mats, lights, cam, out, (uint)W, (uint)H, (uint)depth,
           {(size_t)workBlockX, (size_t)workBlockY}, {(size_t)localX, (size_t)localY}, maxRayDepth};
  uint64_t blocks = (uint64_t)(W / workBlockX) * (uint64_t)(H / workBlockY); /* renderer_opencl.cpp:90 */
  uint64_t rows = blocks * (uint64_t)workBlockY;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  std::atomic<uint64_t> next{0};
  auto worker = [&]() {
    clshim_max_ray_depth = L.maxRayDepth;
    for (;;) {
      uint64_t r = next.fetch_add(1);
      if (r >= rows) break;
      uint block = (uint)(r / L.wbs[1]);
      size_t gy = (size_t)(r % L.wbs[1]);
      for (size_t gx = 0; gx < L.wbs[0]; gx++) {
        clshim_work_item& wi = clshim_wi;
        wi.global_id[0] = gx; wi.global_id[1] = gy; wi.global_id[2] = 0;
        wi.global_size[0] = L.wbs[0]; wi.global_size[1] = L.wbs[1]; wi.global_size[2] = 1;
        wi.local_size[0] = L.ls[0]; wi.local_size[1] = L.ls[1]; wi.local_size[2] = 1;
        wi.local_id[0] = gx % L.ls[0]; wi.local_id[1] = gy % L.ls[1]; wi.local_id[2] = 0;
        wi.group_id[0] = gx / L.ls[0]; wi.group_id[1] = gy / L.ls[1]; wi.group_id[2] = 0;
        wi.num_groups[0] = L.wbs[0] / L.ls[0]; wi.num_groups[1] = L.wbs[1] / L.ls[1]; wi.num_groups[2] = 1;
        run_item(L, block);
      }
    }
  };
  if (threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  return 0;
}

}  // extern "C"
