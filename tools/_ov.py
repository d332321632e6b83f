import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import bench
from lens_trace_b200 import capi, layouts as L
ctx = capi.Context(0); cam = L.make_camera(0, 2.5, -50)
for model, frames in (("cornell_box", 64), ("synth:707", 16)):
    sc = ctx.upload(bench.load_scene(model))
    p = capi.make_params(L.KERNEL_GI, 1920, 1080, max_ray_depth=4, frames=frames, accum_mode=L.ACCUM_RUNNING_MEAN, flags=L.FLAG_WAVEFRONT)
    ts = []
    for i in range(4):
        ctx.render(sc, cam, p, want_output=False); ts.append(ctx.stats().kernel_ms)
    print(os.environ.get("LT_WF_OVERLAP"), os.environ.get("LT_WAVEFRONT_MAX_PATHS"), model, "%%.2f ms" %% min(ts[1:]), flush=True)
''' % ROOT
for ov in ("2",):
    for mp in (str(1 << 25), str(1 << 26)):
        env = dict(os.environ, LT_WF_OVERLAP=ov, LT_WAVEFRONT_MAX_PATHS=mp)
        subprocess.run([sys.executable, "-c", code], env=env)
