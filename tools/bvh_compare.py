"""Host median-split builder vs device LBVH builder: build time, tree depth, traversal work and render time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_trace_b200 import capi, host, layouts as L  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 707
path = "/tmp/bvhcmp_%d.obj" % grid
host.write_synthetic_scene(path, grid, 0x5EED)
t0 = time.perf_counter()
model = host.Model(path)
t_parse = time.perf_counter() - t0
t0 = time.perf_counter()
accel = host.AccelerationStructure(model)
t_host = time.perf_counter() - t0
median = accel.buffers()
print("triangles %d: OBJ parse %.2f s, host median-split build %.2f s" % (len(median.prims), t_parse, t_host), flush=True)

ctx = capi.Context(0)
t0 = time.perf_counter()
sc_l = ctx.build_lbvh(median.prims, median.materials)  # any order works; builder re-sorts
t_gpu_wall = time.perf_counter() - t0
t_gpu_dev = ctx.stats().upload_ms
lb, depth_l = ctx.download(sc_l)
sc_m = ctx.upload(median)
_, depth_m = ctx.download(sc_m)
print("GPU LBVH build: %.1f ms device (H2D of primitives + build + re-flatten), %.1f ms wall; stack depth %d (median split %d)" % (
    t_gpu_dev, t_gpu_wall * 1e3, depth_l, depth_m), flush=True)
cam = L.make_camera(0, 2.5, -50, 0.0, 1)
w, h = 1920, 1080
for name, sc in (("median-split", sc_m), ("LBVH", sc_l)):
    for kernel, frames, label in ((L.KERNEL_BASIC_CU, 1, "primary"), (L.KERNEL_GI, 8, "GI 8 spp")):
        p = capi.make_params(kernel, w, h, max_ray_depth=4, frames=frames, accum_mode=L.ACCUM_RUNNING_MEAN)
        ts = []
        for _ in range(3):
            ctx.accum_reset()
            ctx.render(sc, cam, p, want_output=False)
            ts.append(ctx.stats().kernel_ms)
        ps = capi.make_params(kernel, w, h, max_ray_depth=4, frames=1, accum_mode=L.ACCUM_RUNNING_MEAN, flags=L.FLAG_STATS)
        ctx.render(sc, cam, ps, want_output=False)
        st = ctx.stats()
        print("  %-12s %-9s %8.3f ms   box tests/ray %.1f  triangle tests/ray %.2f" % (
            name, label, min(ts), st.node_tests / st.rays, st.tri_tests / st.rays), flush=True)
a = ctx.render(sc_m, cam, capi.make_params(L.KERNEL_BASIC_CU, w, h))
b = ctx.render(sc_l, cam, capi.make_params(L.KERNEL_BASIC_CU, w, h))
print("pixels whose colour differs between the two trees: %d of %d" % ((np.abs(a - b).max(-1) > 0).sum(), w * h))
for f in (path, path[:-4] + ".mtl"):
    os.remove(f)
