"""CPU tests that pin the hand restatement (oracle/lt_oracle.c) for everything the reference has only as
OpenCL C -- hash RNG, light sampling, shadow rays, GI bounce loop, 25-sample blend, barycentric shade, the
linearKernel clamp -- against the reference's OWN kernel text:

* tests/golden/ref_cl_*.npz were computed by the reference's .cl files, compiled for the CPU through
  oracle/cl_shim and run one work-item at a time (tools/make_reference_cl_golden.py); the restatement in
  LTO_FP_PLAIN mode must reproduce them bit for bit.  Runs everywhere.
* where oracle/_ref/libltref_cl.so exists (built in the container that has /root/reference; it travels
  to the GPU box), the same comparison runs live on more cases: work-block and work-group shapes,
  other cameras, a deeper tree, degenerate scenes.

LTO_FP_PLAIN and LTO_FP_DEVICE (what the CUDA path is compared with) differ only in the helpers OpenCL C
leaves to the implementation (dot/cross fusing, camera rotation, refract, cosf/sinf): see lt_oracle.h."""
import glob
import os

import numpy as np
import pytest

import lt_oracle as O
import lt_ref_cl as R
import util
from lens_trace_b200 import host, layouts as L

STOCHASTIC = (L.KERNEL_LIGHTING25, L.KERNEL_ACCUMULATOR, L.KERNEL_GI25, L.KERNEL_GI)
ALL_CL = (L.KERNEL_BASIC_CL, L.KERNEL_CUSTOM_BARY) + STOCHASTIC

needs_ref_cl = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libltref_cl.so not built (needs /root/reference)")


def plain(kernel, sb, cam, w, h, **kw):
    with O.fp_mode(O.FP_PLAIN):
        return O.render(kernel, sb, cam, w, h, threads=4, **kw)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(util.GOLDEN, "ref_cl_*.npz"))),
                         ids=lambda p: os.path.basename(p)[7:-4])
def test_restatement_reproduces_reference_cl_golden(path):
    g = np.load(path)
    sb = L.SceneBuffers(g["nodes"], g["prims"], g["materials"], g["lights"])
    kernel, w, h = int(g["kernel"]), int(g["width"]), int(g["height"])
    assert len(g["color"]) >= 6
    for i in range(len(g["color"])):
        cam = util.default_camera(float(g["yaw"][i]), int(g["frame_count"][i]))
        got = plain(kernel, sb, cam, w, h, kernel_mode=int(g["mode"][i]), max_ray_depth=int(g["max_ray_depth"][i]))
        util.assert_bit_equal(got, g["color"][i], "%s case %d" % (os.path.basename(path), i))


def test_golden_covers_every_cl_kernel_file():
    kernels = {int(np.load(p)["kernel"]) for p in glob.glob(os.path.join(util.GOLDEN, "ref_cl_*.npz"))}
    assert kernels == set(ALL_CL)


def test_linear_kernel_clamps_tile_kernel_does_not():
    # accumulator.cl:316-318 vs :356-358, GI example :409-411 vs :449-451 (and the 25-sample files)
    sb = util.scene("cornell_box")
    cam = util.default_camera(0.0, 1)
    lin = plain(L.KERNEL_GI, sb, cam, 96, 80, kernel_mode=0)
    tile = plain(L.KERNEL_GI, sb, cam, 96, 80, kernel_mode=1)
    assert tile.max() > 1.0 and tile.min() < 0.0  # unclamped radiance leaves [0,1] on both sides
    assert lin.max() == 1.0 and lin.min() == 0.0
    np.testing.assert_array_equal(lin, np.clip(tile, 0.0, 1.0))


def test_device_and_plain_modes_differ_only_in_rounding():
    # the two arithmetic flavours are the same program: images agree except where a last-place difference
    # moves a hit across an edge or flips a random light/hemisphere choice downstream of it
    sb = util.scene("cornell_box")
    cam = util.default_camera(0.0, 3)
    for kernel in (L.KERNEL_BASIC_CL, L.KERNEL_ACCUMULATOR):
        dev = O.render(kernel, sb, cam, 96, 80)
        pl = plain(kernel, sb, cam, 96, 80)
        close = np.isclose(dev, pl, rtol=1e-4, atol=1e-5).all(axis=-1)
        assert close.mean() > 0.995, (kernel, close.mean())


@needs_ref_cl
def test_cl_struct_layouts_match_the_buffers():
    assert R.struct_sizes() == [L.NODE.itemsize, L.PRIM.itemsize, L.MATERIAL.itemsize, L.LIGHTS.itemsize, 28]


@needs_ref_cl
@pytest.mark.parametrize("kernel", ALL_CL)
@pytest.mark.parametrize("scene", ["cornell_box", "cornell_box_lens"])
def test_live_reference_cl_more_cameras(kernel, scene):
    sb = util.scene(scene)
    w, h = (40, 30) if kernel in (L.KERNEL_LIGHTING25, L.KERNEL_GI25) else (80, 60)
    for yaw, fc, mode in ((-0.02, 2, 0), (0.045, 100, 1), (0.0, 4097, 0)):
        cam = util.default_camera(yaw, fc)
        want = R.render(kernel, sb, cam, w, h, kernel_mode=mode)
        util.assert_bit_equal(plain(kernel, sb, cam, w, h, kernel_mode=mode), want, "%s k%d yaw %g" % (scene, kernel, yaw))


@needs_ref_cl
def test_live_reference_cl_camera_inside_the_box():
    # camera between the walls: rays leave in every direction, negative-t hits (no t > 0 test) matter
    sb = util.scene("cornell_box")
    for kernel in (L.KERNEL_BASIC_CL, L.KERNEL_GI):
        cam = L.make_camera(0.3, 2.0, 1.0, 0.4, 5)
        want = R.render(kernel, sb, cam, 72, 56, kernel_mode=1)
        util.assert_bit_equal(plain(kernel, sb, cam, 72, 56, kernel_mode=1), want, "inside k%d" % kernel)


@needs_ref_cl
@pytest.mark.parametrize("depth", [1, 2, 4, 7])
def test_live_reference_cl_bounce_caps(depth):
    sb = util.scene("cornell_box")
    cam = util.default_camera(0.0, 11)
    want = R.render(L.KERNEL_GI, sb, cam, 64, 48, kernel_mode=1, max_ray_depth=depth)
    util.assert_bit_equal(plain(L.KERNEL_GI, sb, cam, 64, 48, kernel_mode=1, max_ray_depth=depth), want, "depth %d" % depth)


@needs_ref_cl
def test_live_reference_cl_work_blocks_and_groups():
    # the reference's launcher splits the image in work blocks (renderer_opencl.cpp:90,128-146); tileKernel
    # re-derives the pixel from group/local ids.  Any exact tiling gives the same image.
    sb = util.scene("cornell_box")
    cam = util.default_camera(0.0, 2)
    whole = R.render(L.KERNEL_ACCUMULATOR, sb, cam, 64, 48, kernel_mode=0)
    for mode, wb, ls in ((0, (32, 24), (1, 1)), (1, (64, 48), (8, 4)), (1, (32, 16), (16, 2)), (0, (16, 48), (4, 4))):
        img = R.render(L.KERNEL_ACCUMULATOR, sb, cam, 64, 48, kernel_mode=mode, work_block=wb, local=ls)
        if mode == 1:
            np.testing.assert_array_equal(np.clip(img, 0, 1), whole)
        else:
            util.assert_bit_equal(img, whole, "work block %s" % (wb,))
    # integer division drops the remainder: 64x48 with 48x32 blocks renders only the first block
    part = R.render(L.KERNEL_ACCUMULATOR, sb, cam, 64, 48, work_block=(48, 32), fill=-7.0)
    assert (part[:32, :48] == whole[:32, :48]).all() and (part[32:] == -7.0).all() and (part[:, 48:] == -7.0).all()


@needs_ref_cl
def test_live_reference_cl_deeper_tree(tmp_path):
    p = str(tmp_path / "synth.obj")
    host.write_synthetic_scene(p, 20, 0x5EED)  # ~800 triangles, depth ~10
    sb = host.load_scene_buffers(p)
    for kernel, w, h in ((L.KERNEL_GI, 64, 48), (L.KERNEL_ACCUMULATOR, 64, 48), (L.KERNEL_GI25, 24, 16)):
        cam = util.default_camera(0.0, 6)
        want = R.render(kernel, sb, cam, w, h, kernel_mode=1, max_ray_depth=4 if kernel != L.KERNEL_ACCUMULATOR else 16)
        got = plain(kernel, sb, cam, w, h, kernel_mode=1, max_ray_depth=4 if kernel != L.KERNEL_ACCUMULATOR else 16)
        util.assert_bit_equal(got, want, "synthetic k%d" % kernel)


@needs_ref_cl
@pytest.mark.parametrize("maker", ["single", "single_light", "multi_leaf"])
def test_live_reference_cl_degenerate_scenes(maker):
    sb = {"single": util.single_triangle_scene, "single_light": lambda: util.single_triangle_scene(True),
          "multi_leaf": util.multi_prim_leaf_scene}[maker]()
    for kernel in ALL_CL:
        cam = util.default_camera(0.0, 1)
        want = R.render(kernel, sb, cam, 32, 24, kernel_mode=0)
        util.assert_bit_equal(plain(kernel, sb, cam, 32, 24, kernel_mode=0), want, "%s k%d" % (maker, kernel))
