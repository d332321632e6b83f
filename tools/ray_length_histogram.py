#!/usr/bin/env python3
"""Distribution of the camera rays' traversal lengths (box-pair steps = dependent node fetches per ray) on a scene,
through the diagnostic plug-in tools/plugins/count_tests.cu:  python tools/ray_length_histogram.py [synth:707] [W H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from lens_trace_b200 import capi, layouts as L  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "synth:707"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
sb = bench.load_scene(model)
ctx = capi.Context(0)
sc = ctx.upload(sb)
pid = ctx.plugin_load(os.path.join(ROOT, "tools", "plugins", "count_tests.cu"))
out = ctx.render_plugin(sc, L.make_camera(0, 2.5, -50), pid, w, h, block=(8, 4))
steps, tris = out[..., 0], out[..., 1]
print("%s %dx%d: steps/ray mean %.1f max %d; percentiles 50/90/99/99.9/99.99: %s" % (
    model, w, h, steps.mean(), steps.max(), np.percentile(steps, [50, 90, 99, 99.9, 99.99]).tolist()))
print("triangle tests/ray mean %.2f max %d" % (tris.mean(), tris.max()))
# per 8x4 tile (one warp of coherent rays): the warp runs as long as its longest ray
t = steps[:h // 4 * 4, :w // 8 * 8].reshape(h // 4, 4, w // 8, 8)
tmax, tmean = t.max(axis=(1, 3)), t.mean(axis=(1, 3))
print("per 8x4 tile: sum of max %.0f vs sum of mean %.0f (lane utilisation %.2f); longest tile %d steps" % (
    tmax.sum(), tmean.sum(), tmean.sum() / tmax.sum(), tmax.max()))
rows = steps.mean(axis=1)
print("rows with mean > 100 steps:", np.nonzero(rows > 100)[0].min(initial=-1), "..", np.nonzero(rows > 100)[0].max(initial=-1))
ctx.close()
