"""Times the two schedules of the stochastic kernels (persistent megakernel vs wavefront) on a few workloads."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lens_trace_b200 import capi, layouts as L  # noqa: E402

cases = [("cornell_box", L.KERNEL_ACCUMULATOR, 1920, 1080, 1, 0), ("cornell_box", L.KERNEL_ACCUMULATOR, 1920, 1080, 64, 0),
         ("cornell_box", L.KERNEL_GI, 1920, 1080, 1, 4), ("cornell_box", L.KERNEL_GI, 1920, 1080, 8, 4),
         ("cornell_box", L.KERNEL_GI25, 1920, 1080, 1, 4), ("cornell_box", L.KERNEL_GI, 512, 512, 4, 4),
         ("synth:707", L.KERNEL_ACCUMULATOR, 1920, 1080, 1, 0), ("synth:707", L.KERNEL_ACCUMULATOR, 1920, 1080, 16, 0),
         ("synth:707", L.KERNEL_GI, 1920, 1080, 1, 4), ("synth:707", L.KERNEL_GI, 1920, 1080, 8, 4)]
ctx = capi.Context(0)
scenes = {}
for model, kernel, w, h, frames, depth in cases:
    if model not in scenes:
        scenes[model] = ctx.upload(bench.load_scene(model))
    sc = scenes[model]
    cam = L.make_camera(0, 2.5, -50)
    res = []
    for flag in (L.FLAG_MEGAKERNEL, L.FLAG_WAVEFRONT):
        p = capi.make_params(kernel, w, h, max_ray_depth=depth, frames=frames, accum_mode=L.ACCUM_RUNNING_MEAN, flags=flag)
        ts = []
        for i in range(4):
            ctx.render(sc, cam, p, want_output=False)
            ts.append(ctx.stats().kernel_ms)
        res.append(min(ts[1:]))
    print("%-12s kernel %d %dx%d frames %3d: megakernel %8.3f ms  wavefront %8.3f ms  ratio %.2f" % (
        model, kernel, w, h, frames, res[0], res[1], res[0] / res[1]), flush=True)
