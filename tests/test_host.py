"""CPU tests of the host surface: OBJ loader, BVH builder, scene parser, kernel-path registry, and the
C-ABI libraries' exported symbols (no compute without a GPU)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import util
from lens_trace_b200 import capi, host, layouts as L


def test_libraries_export_every_declared_symbol():
    lib = capi.load()
    for name in capi.SYMBOLS:
        assert hasattr(lib, name), name
    hl = host.load()
    for name in host.SYMBOLS:
        assert hasattr(hl, name), name
    assert lib.lt_api_version() == 1
    # every function the header declares is bound above
    header = open(os.path.join(util.ROOT, "include", "lens_trace_b200.h")).read()
    import re
    declared = set(re.findall(r"\b(lt_[a-z_]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.LtError) as e:
        capi.Context(0)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_kernel_registry():
    k = capi.kernel_from_path
    assert k("resources/kernels/cuda/basic.cu") == L.KERNEL_BASIC_CU
    assert k("resources/kernels/opencl/basic.cl") == L.KERNEL_BASIC_CL
    assert k("resources/kernels/custom_opencl.cl") == L.KERNEL_CUSTOM_BARY
    assert k("resources/kernels/opencl/basic_lighting.cl") == L.KERNEL_LIGHTING25
    assert k("resources/kernels/accumulator.cl") == L.KERNEL_ACCUMULATOR
    assert k("/nonexistent/resources/kernels/opencl/global_illumination.cl") == L.KERNEL_GI25
    assert k("/nonexistent/resources/kernels/global_illumination.cl") == L.KERNEL_GI
    assert k(os.path.join(util.ROOT, "resources/kernels/opencl/global_illumination.cl")) == L.KERNEL_GI25
    assert k(os.path.join(util.ROOT, "examples/global_illumination/resources/kernels/global_illumination.cl")) == L.KERNEL_GI
    assert k("my_own_kernel.cl") < 0  # unknown .cl is an error, not a fallback


def test_cornell_box_loads_like_the_reference():
    sb = util.scene("cornell_box")
    assert len(sb.prims) == 42 and len(sb.nodes) == 83 and len(sb.materials) == 5
    assert sb.lights["count"][0] == 2
    lens = util.scene("cornell_box_lens")
    assert len(lens.prims) == 166 and len(lens.nodes) == 331 and len(lens.materials) == 6
    assert (lens.materials["dissolve"] < 1).sum() == 1  # Material.005, d 0.25
    # light primitives are the emissive ones
    for sbx in (sb, lens):
        n = sbx.lights["count"][0]
        for p in sbx.lights["prims"][0][:n]:
            assert sbx.materials["emission"][sbx.prims["mat"][p]].max() > 0


def test_quad_split_rule(tmp_path):
    # shorter diagonal; a tie takes [0,1,3],[1,2,3] (tiny_obj_loader.h:1447-1487)
    obj = tmp_path / "q.obj"
    (tmp_path / "q.mtl").write_text("newmtl M\nKd 1 1 1\n")
    obj.write_text("mtllib q.mtl\nusemtl M\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 4 0 1\nv 5 1 1\nv 0 1 1\n"
                   "vn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4//1 -3//1 -2//1 -1//1\n")
    m = host.Model(str(obj))
    assert m.primitive_count == 4
    a = host.AccelerationStructure(m)
    sb = a.buffers()
    tris = {tuple(map(tuple, np.stack([p["a"], p["b"], p["c"]]).tolist())) for p in sb.prims}
    assert ((0, 0, 0), (1, 0, 0), (0, 1, 0)) in tris and ((1, 0, 0), (1, 1, 0), (0, 1, 0)) in tris  # tie
    # second quad: diagonal 0-2 is (5,1) long, 1-3 is (4,1): 1-3 shorter -> [0,1,3],[1,2,3]
    assert ((0, 0, 1), (4, 0, 1), (0, 1, 1)) in tris and ((4, 0, 1), (5, 1, 1), (0, 1, 1)) in tris
    a.close()
    m.close()


def _check_tree(sb):
    nodes, n = sb.nodes, len(sb.nodes)
    seen = np.zeros(len(sb.prims), int)

    def walk(i):
        nd = nodes[i]
        if nd["count"] > 0:
            assert nd["count"] == 1
            p = sb.prims[nd["offset"]]
            pts = np.stack([p["a"], p["b"], p["c"]])
            np.testing.assert_array_equal(nd["min"], pts.min(0))
            np.testing.assert_array_equal(nd["max"], pts.max(0))
            seen[nd["offset"]] += 1
            return nd["min"], nd["max"], i + 1
        lmin, lmax, nxt = walk(i + 1)
        assert nd["offset"] == nxt  # DFS order: second child follows the first subtree
        rmin, rmax, nxt = walk(nd["offset"])
        np.testing.assert_array_equal(nd["min"], np.minimum(lmin, rmin))
        np.testing.assert_array_equal(nd["max"], np.maximum(lmax, rmax))
        assert nd["axis"] in (0, 1, 2)
        return nd["min"], nd["max"], nxt

    _, _, end = walk(0)
    assert end == n and (seen == 1).all()


def test_bvh_invariants():
    for name in ("green_wall", "cornell_box", "cornell_box_lens"):
        _check_tree(util.scene(name))


def test_builder_is_deterministic():
    a = host.load_scene_buffers(os.path.join(util.MODELS, "cornell_box_lens.obj"))
    b = host.load_scene_buffers(os.path.join(util.MODELS, "cornell_box_lens.obj"))
    assert a.nodes.tobytes() == b.nodes.tobytes() and a.prims.tobytes() == b.prims.tobytes()
    assert a.lights.tobytes() == b.lights.tobytes()


def test_synthetic_scene(tmp_path):
    p = str(tmp_path / "s.obj")
    n = host.write_synthetic_scene(p, 24, 0x5EED)
    assert n == 12 + 2 * 24 * 24
    sb = host.load_scene_buffers(p)
    assert len(sb.prims) == n and len(sb.nodes) == 2 * n - 1
    assert sb.lights["count"][0] == 2
    _check_tree(sb)
    q = str(tmp_path / "t.obj")
    host.write_synthetic_scene(q, 24, 0x5EED)
    assert open(p).read().split("\n", 1)[1] == open(q).read().split("\n", 1)[1].replace("t.mtl", "s.mtl")
    assert sb.prims["min" if False else "a"].min() >= -2.5001 and sb.prims["a"][:, 1].max() <= 5.0001


def test_camera_buffer_layout():
    c = host.Camera(1, 2, 3, 0.5)
    b = c.buffer()
    assert tuple(b["pos"][0]) == (1, 2, 3) and b["yaw"][0] == np.float32(0.5) and b["frameCount"][0] == 0
    c.increment_frame_count()
    c.increment_frame_count()
    assert c.buffer()["frameCount"][0] == 2
    c.set_frame_count(0)
    assert c.buffer()["frameCount"][0] == 0
    c.close()


def test_scene_files_parse(tmp_path):
    # run_scene_file needs a GPU to render; here only the parser: a bad file is reported, not fatal
    bad = tmp_path / "bad.scene"
    bad.write_text("{ not json")
    out = np.zeros(4, np.float32)
    dims = (C.c_uint64 * 3)()
    rc = host.load().lth_run_scene_file(str(bad).encode(), out.ctypes.data, out.nbytes, C.byref(dims))
    assert rc == -1
    for f in os.listdir(os.path.join(util.ROOT, "resources", "scenes")):
        json.load(open(os.path.join(util.ROOT, "resources", "scenes", f)))


def test_reference_sources_compile_against_these_headers(tmp_path):
    """Drop-in at the source level: the reference's own example and CLI compile, unedited, against
    include/lens_trace (needs /root/reference; skipped on the GPU box)."""
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference sources not present")
    inc = os.path.join(util.ROOT, "include")
    for src, defs in ((ref + "/examples/custom_kernel/src/main.cpp", []),
                      (ref + "/src/main.cpp", ["-DCUDA_ENABLED", "-DOPENCL_ENABLED"])):
        r = subprocess.run(["/usr/bin/g++", "-fsyntax-only", "-I", inc] + defs + [src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
