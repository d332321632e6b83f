#!/usr/bin/env python3
"""Generates tests/golden/ref_cuda_*.npz ON THE GPU BOX from the reference itself (oracle/_ref):
buffers built by the reference's own Model + AccelerationStructureExplicit, and the outputs of the
reference's RendererCUDA (NVRTC build of its basic.cu, and the id-dump variants of that kernel) on a B200.
Run: gpurun -- python tools/make_reference_golden.py ; then copy gpurun_out/golden/*.npz to tests/golden/.
The CPU test tests/test_oracle.py::test_oracle_against_reference_generated_golden checks the oracle
against these files in every round, without a GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import test_gpu_reference_ab as ab  # noqa: E402  (RefWorld: the ctypes view of oracle/_ref/libltref.so)
import util  # noqa: E402


def main():
    lib = ab._load_ref()
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    w, h = 96, 72
    for name in ("green_wall", "cornell_box", "cornell_box_lens"):
        for yaw in (0.0, 0.05):
            world = ab.RefWorld(lib, name)
            color = world.render("basic.cu", w, h, yaw)
            data = dict(nodes=world.sb.nodes, prims=world.sb.prims, materials=world.sb.materials,
                        lights=world.sb.lights, camera=util.default_camera(yaw).view(np.uint8), width=w, height=h,
                        color=color)
            if name != "cornell_box_lens":  # id dumps report the post-lens payload on the lens scene
                a = world.render("id_dump_a.cu", w, h, yaw)
                b = world.render("id_dump_b.cu", w, h, yaw)
                hit = (b[..., 0].view(np.int32) == 2).astype(np.int32)
                ids = np.where(hit == 1, a[..., 0].view(np.int32) - 1, 0).astype(np.int32)
                tuv = np.stack([a[..., 1], a[..., 2], b[..., 1]], axis=-1).astype(np.float32)
                data.update(ids=ids, hit=hit, tuv=tuv)
            path = os.path.join(out_dir, "ref_cuda_%s_yaw%03d.npz" % (name, int(round(yaw * 100))))
            np.savez_compressed(path, **data)
            print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
