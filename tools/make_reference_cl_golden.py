"""Generates tests/golden/ref_cl_*.npz: images computed by the reference's OWN OpenCL kernel files
(resources/kernels/opencl/*.cl, examples/*/resources/kernels/*.cl) executed on the CPU through
oracle/cl_shim (oracle/build_ref_cl.sh -> oracle/_ref/libltref_cl.so).  Needs /root/reference (the
library is built from the kernel text there), so it runs in the build container only; the fixtures
it writes are small and committed, and tests/test_oracle_cl.py checks the hand restatement
(lt_oracle.c, LTO_FP_PLAIN) against them everywhere.

    python tools/make_reference_cl_golden.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import lt_ref_cl as R  # noqa: E402
import util  # noqa: E402
from lens_trace_b200 import layouts as L  # noqa: E402

# (kernel id, tag, width, height)
KERNELS = [(L.KERNEL_BASIC_CL, "basic", 64, 48), (L.KERNEL_CUSTOM_BARY, "custom_opencl", 64, 48),
           (L.KERNEL_LIGHTING25, "basic_lighting", 32, 24), (L.KERNEL_ACCUMULATOR, "accumulator", 64, 48),
           (L.KERNEL_GI25, "global_illumination_resources", 32, 24), (L.KERNEL_GI, "global_illumination_example", 64, 48)]


def main():
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref_cl.sh")])
    out_dir = os.path.join(ROOT, "tests", "golden")
    for scene in ("cornell_box", "cornell_box_lens", "green_wall"):
        sb = util.scene(scene)
        for kernel, tag, w, h in KERNELS:
            cases = []
            for mode in (0, 1):
                for frame_count, yaw in ((0, 0.0), (1, 0.0), (9, 0.03)):
                    depths = (16, 4) if kernel in (L.KERNEL_GI25, L.KERNEL_GI) else (16,)
                    for depth in depths:
                        cam = util.default_camera(yaw, frame_count)
                        img = R.render(kernel, sb, cam, w, h, kernel_mode=mode, max_ray_depth=depth)
                        cases.append((mode, frame_count, yaw, depth, img))
            path = os.path.join(out_dir, "ref_cl_%s_%s.npz" % (tag, scene))
            np.savez_compressed(
                path, kernel=kernel, width=w, height=h,
                mode=np.array([c[0] for c in cases], np.int32), frame_count=np.array([c[1] for c in cases], np.uint32),
                yaw=np.array([c[2] for c in cases], np.float32), max_ray_depth=np.array([c[3] for c in cases], np.int32),
                color=np.stack([c[4] for c in cases]),
                nodes=sb.nodes, prims=sb.prims, materials=sb.materials, lights=sb.lights)
            print(path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
