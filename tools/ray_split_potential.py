#!/usr/bin/env python3
"""How much shorter would the longest dependent chain of a camera ray get if its traversal were split over the
subtrees below the root (tools/plugins/count_split.cu)?  python tools/ray_split_potential.py [synth:707]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from lens_trace_b200 import capi, layouts as L  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "synth:707"
w, h = 1920, 1080
sb = bench.load_scene(model)
ctx = capi.Context(0)
sc = ctx.upload(sb)
pid = ctx.plugin_load(os.path.join(ROOT, "tools", "plugins", "count_split.cu"))
out = ctx.render_plugin(sc, L.make_camera(0, 2.5, -50), pid, w, h, block=(8, 4))
total, best4, best2 = out[..., 0], out[..., 1], out[..., 2]
heavy = total > 200
print("rays with > 200 steps: %d; their mean total %.0f, mean largest-of-4 %.0f (%.2f), mean larger-of-2 %.0f (%.2f)" % (
    heavy.sum(), total[heavy].mean(), best4[heavy].mean(), (best4[heavy] / total[heavy]).mean(), best2[heavy].mean(),
    (best2[heavy] / total[heavy]).mean()))
t = lambda a: a[:h // 4 * 4, :w // 8 * 8].reshape(h // 4, 4, w // 8, 8).max(axis=(1, 3))
print("sum over 8x4 tiles of the tile's longest chain: whole rays %.0f, 2-way split %.0f, 4-way split %.0f; longest: %d / %d / %d" % (
    t(total).sum(), t(best2).sum(), t(best4).sum(), t(total).max(), t(best2).max(), t(best4).max()))
ctx.close()
