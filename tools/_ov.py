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
    for fl in (L.FLAG_WAVEFRONT, L.FLAG_WAVEFRONT | L.FLAG_SERIAL):
        p = capi.make_params(L.KERNEL_GI, 1920, 1080, max_ray_depth=4, frames=frames, accum_mode=L.ACCUM_RUNNING_MEAN, flags=fl)
        ts = []
        for i in range(4):
            ctx.render(sc, cam, p, want_output=False); ts.append(ctx.stats().kernel_ms)
        print(os.environ.get("LT_TAG"), model, "serial" if fl & L.FLAG_SERIAL else "overlap", "%%.2f ms" %% min(ts[1:]), flush=True)
''' % ROOT
for tag, env in (("shade8", {"LT_WF_SHADE_BLOCKS_PER_SM": "8"}), ("shade6", {"LT_WF_SHADE_BLOCKS_PER_SM": "6"}),
                 ("shade4", {"LT_WF_SHADE_BLOCKS_PER_SM": "4"})):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, LT_TAG=tag, **env))
