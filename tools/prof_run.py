"""Small fixed workload for ncu captures: one k_path launch (Cornell GI, 1080p, N frames) and one k_flat
launch (primary rays).  Usage: python tools/prof_run.py [frames] [workload-model]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_trace_b200 import capi, host, layouts as L  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = int(os.environ.get("LT_PROF_FLAGS", "0"))
model = sys.argv[2] if len(sys.argv) > 2 else "cornell_box"
if model.startswith("synth:"):
    path = "/tmp/prof_synth.obj"
    host.write_synthetic_scene(path, int(model.split(":")[1]), 0x5EED)
    sb = host.load_scene_buffers(path)
else:
    sb = host.load_scene_buffers(os.path.join(ROOT, "resources", "models", model + ".obj"))
ctx = capi.Context(0)
sc = ctx.upload(sb)
cam = L.make_camera(0, 2.5, -50)
w, h = 1920, 1080
for rep in range(2):
    ctx.render(sc, cam, capi.make_params(L.KERNEL_GI, w, h, max_ray_depth=4, frames=frames,
                                         accum_mode=L.ACCUM_RUNNING_MEAN, flags=flags), want_output=False)
    print("k_path %d frames: %.3f ms" % (frames, ctx.stats().kernel_ms))
    ctx.render(sc, cam, capi.make_params(L.KERNEL_BASIC_CU, w, h), want_output=False)
    print("k_flat: %.3f ms" % ctx.stats().kernel_ms)
sc.release()
ctx.close()
