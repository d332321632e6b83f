"""The reference's five renderer tests (tests/cuda_renderer_test.cc, tests/opencl_renderer_test.cc),
re-expressed against this repo's RendererCUDA / RendererOpenCL through the C++ host surface."""
import os

import numpy as np
import pytest

import lt_oracle as O
import util
from lens_trace_b200 import host, layouts as L

pytestmark = pytest.mark.gpu

CU = "resources/kernels/cuda/basic.cu"
CL = "resources/kernels/opencl/basic.cl"


@pytest.fixture(scope="module")
def world():
    os.chdir(util.ROOT)  # resource paths are CWD-relative (src/resource.cpp:3-16)
    cam = host.Camera(0, 2.5, -50, 0)
    model = host.Model("resources/models/green_wall.obj")
    accel = host.AccelerationStructure(model)
    yield cam, model, accel
    accel.close()
    model.close()
    cam.close()


@pytest.mark.parametrize("platform,kernel", [(host.PLATFORM_CUDA, CU), (host.PLATFORM_OPENCL, CL)])
def test_create_engine_and_valid_buffer(world, platform, kernel):
    cam, model, accel = world
    r = host.Renderer(platform)  # CreateEngineTEST.ValidEngine
    assert r.h
    out = r.render(kernel, 100, 100, accel, model, cam)  # RenderBufferTEST.ValidBuffer
    assert out.shape == (100, 100, 3)
    r.close()


@pytest.mark.parametrize("platform,kernel", [(host.PLATFORM_CUDA, CU), (host.PLATFORM_OPENCL, CL)])
def test_custom_block_size(world, platform, kernel):
    cam, model, accel = world
    r = host.Renderer(platform)
    a = r.render(kernel, 100, 100, accel, model, cam).reshape(-1)
    b = r.render(kernel, 100, 100, accel, model, cam, block=(8, 8)).reshape(-1)
    c = r.render(kernel, 100, 100, accel, model, cam, block=(4, 4)).reshape(-1)
    for x in range(0, 100 * 100, 32):  # tests/cuda_renderer_test.cc:106-109
        assert a[x] == b[x] == c[x]
    r.close()


@pytest.mark.parametrize("platform,kernel", [(host.PLATFORM_CUDA, CU), (host.PLATFORM_OPENCL, CL)])
def test_kernel_mode(world, platform, kernel):
    cam, model, accel = world
    r = host.Renderer(platform)
    a = r.render(kernel, 100, 100, accel, model, cam, kernel_mode=0).reshape(-1)
    b = r.render(kernel, 100, 100, accel, model, cam, kernel_mode=1).reshape(-1)
    for x in range(0, 100 * 100 * 3, 32):  # tests/cuda_renderer_test.cc:173-175
        assert a[x] == b[x]
    r.close()


@pytest.mark.parametrize("platform,kernel", [(host.PLATFORM_CUDA, CU), (host.PLATFORM_OPENCL, CL)])
def test_correct_color(world, platform, kernel):
    cam, model, accel = world
    r = host.Renderer(platform)
    flat = r.render(kernel, 100, 100, accel, model, cam).reshape(-1)
    for x in range(0, 100 * 100, 8 * 3):  # tests/cuda_renderer_test.cc:217-221
        assert flat[x] == 0.0 and flat[x + 1] == 1.0 and flat[x + 2] == 0.0
    r.close()


def test_examples_flow_progressive_accumulation():
    """examples/global_illumination/src/main.cpp:269-340 without the GL window: RendererOpenCL, the
    example's kernel path, frameCount protocol, running mean kept on the device."""
    os.chdir(util.ROOT)
    cam = host.Camera(0, 2.5, -50, 0)
    model = host.Model("resources/models/cornell_box.obj")
    accel = host.AccelerationStructure(model)
    r = host.Renderer(host.PLATFORM_OPENCL)
    kernel = "examples/global_illumination/resources/kernels/global_illumination.cl"
    w, h, frames = 96, 72, 5
    ext = host.make_extension(frames=frames, accumulate=True, max_ray_depth=4, collect_stats=True)
    got = r.render(kernel, w, h, accel, model, cam, ext=ext)
    sb = accel.buffers()
    acc = np.zeros((h, w, 3), np.float32)
    for f in range(frames):
        O.accumulate(acc, O.render(L.KERNEL_GI, sb, util.default_camera(0.0, f), w, h, max_ray_depth=4, threads=0), f)
    bad = (util.bits(got) != util.bits(acc)).any(axis=-1)
    assert bad.sum() <= 3
    np.testing.assert_allclose(got, acc, rtol=1e-4, atol=1e-6)
    assert ext.rays >= frames * w * h and ext.kernelMilliseconds > 0
    # an unknown kernel file is a reported error and leaves the buffer untouched
    out = np.full((h, w, 3), -1, np.float32)
    r.render("resources/kernels/opencl/not_a_kernel.cl", w, h, accel, model, cam, out=out)
    assert (out == -1).all()
    r.close()
    accel.close()
    model.close()
    cam.close()


def test_scene_file_end_to_end():
    import ctypes as C
    os.chdir(util.ROOT)
    out = np.zeros((2048, 2048, 3), np.float32)
    dims = (C.c_uint64 * 3)()
    rc = host.load().lth_run_scene_file(b"resources/scenes/cornell_basic_cuda.scene", out.ctypes.data, out.nbytes, C.byref(dims))
    assert rc == 0 and tuple(dims) == (2048, 2048, 3)
    want = O.render(L.KERNEL_BASIC_CU, util.scene("cornell_box"), util.default_camera(), 2048, 2048, rows=(1000, 1016),
                    threads=0)
    util.assert_bit_equal(out[1000:1016], want[1000:1016])


def test_reference_example_binaries_run_unchanged(tmp_path):
    """The reference's examples/custom_kernel/src/main.cpp and src/main.cpp, compiled without edits against
    this repo's headers and liblenstrace.so (oracle/build_ref.sh), run on the B200 path."""
    import subprocess
    exe = os.path.join(util.ROOT, "oracle", "_ref", "bin", "ref_custom_kernel")
    cli = os.path.join(util.ROOT, "oracle", "_ref", "bin", "ref_LensTrace")
    if not (os.path.exists(exe) and os.path.exists(cli)):
        pytest.skip("oracle/_ref/bin not built")
    # the example resolves resources/kernels/custom_opencl.cl and resources/models/cornell_box.obj from the CWD
    work = tmp_path / "custom_kernel"
    (work / "resources" / "kernels").mkdir(parents=True)
    os.symlink(os.path.join(util.ROOT, "resources", "models"), work / "resources" / "models")
    os.symlink(os.path.join(util.ROOT, "examples", "custom_kernel", "resources", "kernels", "custom_opencl.cl"),
               work / "resources" / "kernels" / "custom_opencl.cl")
    r = subprocess.run([exe], cwd=work, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ppm = (work / "output.jpg.ppm").read_bytes()
    assert ppm.startswith(b"P6\n800 800\n255\n")
    pixels = np.frombuffer(ppm[len(b"P6\n800 800\n255\n"):], dtype=np.uint8).reshape(800, 800, 3)
    want = O.render(L.KERNEL_CUSTOM_BARY, util.scene("cornell_box"), util.default_camera(), 800, 800, threads=0)
    np.testing.assert_array_equal(pixels, (want * 255).astype(np.int8).view(np.uint8))
    # the command-line program on a scene file of this repo (2048x2048, basic.cu pipeline)
    r = subprocess.run([cli, "resources/scenes/cornell_basic_cuda.scene"], cwd=util.ROOT, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = os.path.join(util.ROOT, "output.jpg.ppm")
    assert os.path.getsize(out) > 2048 * 2048 * 3
    os.remove(out)


def _echo_expected(sb, cam_z, frame_count, w, h):
    idx = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0)
    idy = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1)
    p = (np.arange(w * h) % 42).reshape(h, w)
    want = np.zeros((h, w, 3), np.float32)
    want[..., 0] = np.float32(cam_z) + idx + np.float32(frame_count)
    want[..., 1] = sb.materials["diffuse"][sb.prims["mat"][p], 1] + sb.prims["b"][p, 0]
    want[..., 2] = np.float32(sb.lights["count"][0]) + sb.nodes["max"][0, 1] + idy
    return want


def test_plugin_kernel_through_the_c_abi():
    """A user-written .cu with the reference's kernel ABI: compiled by NVRTC for sm_100a, launched on the
    uploaded buffers in the reference layouts, both entry points, default and custom block shapes."""
    from lens_trace_b200 import capi
    ctx = capi.Context(0)
    sb = util.scene("cornell_box")
    sc = ctx.upload(sb)
    pid = ctx.plugin_load(os.path.join(util.ROOT, "tests", "plugins", "echo_buffers.cu"))
    cam = util.default_camera(0.0, 5)
    w, h = 101, 37
    want = _echo_expected(sb, -50.0, 5, w, h)
    for mode, block in ((0, (0, 0)), (1, (0, 0)), (0, (8, 8)), (1, (4, 4)), (0, (16, 2))):
        got = ctx.render_plugin(sc, cam, pid, w, h, kernel_mode=mode, block=block)
        util.assert_bit_equal(got, want, "plug-in mode %d block %s" % (mode, block))
    with pytest.raises(capi.LtError) as e:
        ctx.plugin_load(os.path.join(util.ROOT, "tests", "plugins", "broken.cu"))
    assert "undefined_symbol" in str(e.value)
    with pytest.raises(capi.LtError):
        ctx.plugin_load("/nonexistent/kernel.cu")
    sc.release()
    ctx.close()


def test_plugin_kernel_through_renderer_cuda(world):
    """kernelFilePath naming a .cu that is not a shipped kernel goes to the plug-in path (custom_kernel surface)."""
    os.chdir(util.ROOT)
    cam = host.Camera(0, 2.5, -50, 0)
    model = host.Model("resources/models/cornell_box.obj")
    accel = host.AccelerationStructure(model)
    r = host.Renderer(host.PLATFORM_CUDA)
    got = r.render("tests/plugins/echo_buffers.cu", 64, 40, accel, model, cam, block=(8, 8))
    util.assert_bit_equal(got, _echo_expected(accel.buffers(), -50.0, 0, 64, 40))
    out = np.full((40, 64, 3), -1, np.float32)
    r.render("tests/plugins/broken.cu", 64, 40, accel, model, cam, out=out)  # reported, buffer untouched
    assert (out == -1).all()
    r.close()
    accel.close()
    model.close()
    cam.close()


def test_reference_kernel_source_as_a_plugin():
    """The reference's own basic.cu, run UNMODIFIED through the plug-in path (NVRTC for sm_100a), gives the
    same picture as the built-in pipeline that replaces it."""
    from lens_trace_b200 import capi
    src = os.path.join(util.ROOT, "oracle", "_ref", "resources", "kernels", "cuda", "basic.cu")
    if not os.path.exists(src):
        pytest.skip("oracle/_ref not built")
    ctx = capi.Context(0)
    for name in ("cornell_box", "cornell_box_lens"):
        sb = util.scene(name)
        sc = ctx.upload(sb)
        pid = ctx.plugin_load(src)
        for yaw in (0.0, 0.04):
            cam = util.default_camera(yaw)
            a = ctx.render_plugin(sc, cam, pid, 200, 150, block=(8, 8))
            b = ctx.render(sc, cam, capi.make_params(L.KERNEL_BASIC_CU, 200, 150))
            util.assert_bit_equal(a, b, "%s yaw %g: reference source as plug-in vs built-in pipeline" % (name, yaw))
        sc.release()
    ctx.close()


def test_headless_examples_build_and_run(tmp_path):
    """examples/: the three example programs and the CLI, built against liblenstrace.so, run from the repo root."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(util.ROOT, "examples"), "-s"])
    out = str(tmp_path / "gi.pfm")
    r = subprocess.run([os.path.join(util.ROOT, "examples", "bin", "global_illumination"), "--size", "160", "120",
                        "--frames", "6", "--depth", "4", "--out", out], cwd=util.ROOT, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    header = b"PF\n160 120\n-1.0\n"
    assert raw.startswith(header)
    batched = np.frombuffer(raw[len(header):], dtype=np.float32).reshape(120, 160, 3)
    # the per-frame protocol of the reference's example (one render() per frame, mean on the host) gives the same picture
    out2 = str(tmp_path / "gi2.pfm")
    r = subprocess.run([os.path.join(util.ROOT, "examples", "bin", "global_illumination"), "--size", "160", "120",
                        "--frames", "6", "--depth", "4", "--per-frame", "--out", out2], cwd=util.ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw2 = open(out2, "rb").read()
    per_frame = np.frombuffer(raw2[len(header):], dtype=np.float32).reshape(120, 160, 3)
    acc = np.zeros((120, 160, 3), np.float32)
    sb = util.scene("cornell_box")
    for f in range(6):
        O.accumulate(acc, O.render(L.KERNEL_GI, sb, util.default_camera(0.0, f), 160, 120, max_ray_depth=4, threads=0), f)
    assert ((util.bits(batched) != util.bits(acc)).any(axis=-1)).sum() <= 3
    np.testing.assert_allclose(batched, acc, rtol=1e-4, atol=1e-6)
    util.assert_bit_equal(per_frame, batched, "one render() per frame + host mean vs one batched render()")
    for exe, args in (("accumulator", ["--size", "96", "64", "--frames", "3", "--out", str(tmp_path / "a.ppm")]),
                      ("custom_kernel", ["96", "64", str(tmp_path / "c.ppm")])):
        r = subprocess.run([os.path.join(util.ROOT, "examples", "bin", exe)] + args, cwd=util.ROOT, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr


def test_acceleration_structure_gpu_lbvh_option():
    """AccelerationStructureExplicit with ACCELERATION_STRUCTURE_TYPE_LBVH_B200: built on the GPU, held on the host
    in the reference layout, rendered through RendererCUDA like any other."""
    os.chdir(util.ROOT)
    cam = host.Camera(0, 2.5, -50, 0)
    model = host.Model("resources/models/cornell_box.obj")
    accel = host.AccelerationStructure(model, host.AccelerationStructure.GPU_LBVH)
    sb = accel.buffers()
    assert len(sb.nodes) == 83 and sb.lights["count"][0] == 2
    r = host.Renderer(host.PLATFORM_CUDA)
    got = r.render("resources/kernels/cuda/basic.cu", 128, 96, accel, model, cam)
    util.assert_bit_equal(got, O.render(L.KERNEL_BASIC_CU, sb, util.default_camera(), 128, 96))
    r.close(); accel.close(); model.close(); cam.close()


def test_plugin_device_api_lt_trace():
    """A plug-in that includes lens_trace_b200_device.cuh and calls lt_trace(): same hit ids and t as the built-in
    kernels (bit for bit), any-hit shadow query works."""
    from lens_trace_b200 import capi
    ctx = capi.Context(0)
    for name in ("cornell_box", "cornell_box_lens"):
        sb = util.scene(name)
        sc = ctx.upload(sb)
        pid = ctx.plugin_load(os.path.join(util.ROOT, "tests", "plugins", "fast_trace.cu"))
        for yaw in (0.0, 0.04):
            cam = util.default_camera(yaw)
            got = ctx.render_plugin(sc, cam, pid, 240, 160, block=(8, 8))
            ids, hit, tuv = ctx.primary_hits(sc, cam, L.KERNEL_BASIC_CU, 240, 160)
            m = hit == 1
            np.testing.assert_array_equal(got[..., 0].view(np.int32)[m] - 1, ids[m])
            assert (got[..., 0][~m] == 0).all()
            util.assert_bit_equal(got[..., 1][m], tuv[..., 0][m], "t from lt_trace vs built-in kernel")
            assert set(np.unique(got[..., 2]).tolist()) <= {0.0, 1.0} and got[..., 2][m].max() == 1.0
        sc.release()
    ctx.close()


def test_registry_goes_by_content_through_the_renderer(tmp_path):
    """a user who edits basic.cu (the reference's custom-kernel mechanism) must get THEIR kernel, not
    the built-in pipeline of that name; an edited .cl is a reported error that leaves the buffer untouched."""
    ref_cu = os.path.join(util.ROOT, "oracle", "_ref", "resources", "kernels", "cuda", "basic.cu")
    if not os.path.exists(ref_cu):
        pytest.skip("oracle/_ref not built")
    os.chdir(util.ROOT)
    cam = host.Camera(0, 2.5, -50, 0)
    model = host.Model("resources/models/cornell_box.obj")
    accel = host.AccelerationStructure(model)
    r = host.Renderer(host.PLATFORM_CUDA)
    w, h = 160, 120
    builtin = r.render("resources/kernels/cuda/basic.cu", w, h, accel, model, cam)
    text = open(ref_cu).read()
    # the unmodified reference text under another directory: recognised by content -> built-in pipeline
    (tmp_path / "a").mkdir()
    same = tmp_path / "a" / "basic.cu"
    same.write_text(text)
    from lens_trace_b200 import capi
    assert capi.kernel_from_path(str(same)) == L.KERNEL_BASIC_CU
    util.assert_bit_equal(r.render(str(same), w, h, accel, model, cam), builtin)
    # same NAME, edited colour store: must run as a plug-in and show the edit
    (tmp_path / "b").mkdir()
    edited = tmp_path / "b" / "basic.cu"
    assert text.count("output[id + 0] = outputColor.x;") >= 1  # linearKernel and tileKernel
    edited.write_text(text.replace("output[id + 0] = outputColor.x;", "output[id + 0] = outputColor.x * 0.5f + 0.25f;"))
    assert capi.kernel_from_path(str(edited)) < 0
    got = r.render(str(edited), w, h, accel, model, cam, block=(8, 8))
    want = builtin.copy()
    want[..., 0] = want[..., 0] * np.float32(0.5) + np.float32(0.25)
    util.assert_bit_equal(got, want, "edited basic.cu runs as the user's kernel")
    # edited .cl under the shipped name: error, buffer untouched
    cl = tmp_path / "custom_opencl.cl"
    cl.write_text("// my own shading\n__kernel void linearKernel() {}\n")
    out = np.full((h, w, 3), -1, np.float32)
    host.Renderer(host.PLATFORM_OPENCL).render(str(cl), w, h, accel, model, cam, out=out)
    assert (out == -1).all()
    r.close(); accel.close(); model.close(); cam.close()


def test_scene_cache_follows_object_identity_and_content():
    """the device-scene cache was keyed by host addresses.  Now: ids that are never reused + a content
    checksum; destroyed objects are dropped; at most 8 device scenes are kept."""
    os.chdir(util.ROOT)
    cam = host.Camera(0, 2.5, -50, 0)
    r = host.Renderer(host.PLATFORM_CUDA)
    K = "resources/kernels/cuda/basic.cu"
    w, h = 96, 64
    imgs = {}
    for name in ("green_wall", "cornell_box"):
        imgs[name] = O.render(L.KERNEL_BASIC_CU, util.scene(name), util.default_camera(), w, h)
    # alternate models; every structure is destroyed before the next is created (same addresses are likely reused)
    for rep in range(6):
        name = ("green_wall", "cornell_box")[rep % 2]
        model = host.Model("resources/models/%s.obj" % name)
        accel = host.AccelerationStructure(model)
        util.assert_bit_equal(r.render(K, w, h, accel, model, cam), imgs[name], "rep %d %s" % (rep, name))
        util.assert_bit_equal(r.render(K, w, h, accel, model, cam), imgs[name])
        accel.close()
        model.close()
    assert r.cached_scenes() <= 1  # dead pairs are dropped at the next render()
    # an in-place edit of the host buffers is seen (checksum), and the bound on live scenes holds
    model = host.Model("resources/models/green_wall.obj")
    accel = host.AccelerationStructure(model)
    before = r.render(K, w, h, accel, model, cam)
    mats = host._view(host.load().lth_model_material_buffer(model.h), host.load().lth_model_material_bytes(model.h), L.MATERIAL)
    mats["diffuse"][:] = (0.25, 0.5, 0.75)
    after = r.render(K, w, h, accel, model, cam)
    assert (before != after).any() and set(np.unique(after).tolist()) <= {0.0, 0.25, 0.5, 0.75}
    keep = []
    for k in range(10):
        m = host.Model("resources/models/green_wall.obj")
        a = host.AccelerationStructure(m)
        r.render(K, w, h, a, m, cam)
        keep.append((m, a))
    assert r.cached_scenes() == 8
    for m, a in keep:
        a.close()
        m.close()
    r.render(K, w, h, accel, model, cam)
    assert r.cached_scenes() == 1
    r.close(); accel.close(); model.close(); cam.close()


def test_gi_kernel_as_a_plugin_on_the_device_api():
    """the example's GI kernel re-written as a .cu plug-in on lt_trace / lt_occluded / lt_random /
    lt_sample_light / lt_sample_hemisphere equals the built-in LT_KERNEL_GI pipeline bit for bit (4 bounces, several
    frameCounts, Cornell box and the lens scene, a mesh large enough for the stack traversal), and its kernel time stays
    within 1.5x of the built-in single-frame launch."""
    from lens_trace_b200 import capi
    ctx = capi.Context(0)
    pid = ctx.plugin_load(os.path.join(util.ROOT, "tests", "plugins", "gi_plugin.cu"))
    for name in ("cornell_box", "cornell_box_lens"):
        sb = util.scene(name)
        sc = ctx.upload(sb)
        for frame in (0, 1, 9):
            cam = util.default_camera(0.02, frame)
            for block in ((8, 8), (32, 1), (16, 4)):
                got = ctx.render_plugin(sc, cam, pid, 320, 200, block=block)
                want = ctx.render(sc, cam, capi.make_params(L.KERNEL_GI, 320, 200, max_ray_depth=4))
                util.assert_bit_equal(got, want, "%s frame %d block %s: GI plug-in vs built-in" % (name, frame, block))
        sc.release()
    # timing at 1080p on the Cornell box, one frame per launch
    sb = util.scene("cornell_box")
    sc = ctx.upload(sb)
    cam = util.default_camera(0.0, 3)
    t_plugin, t_builtin = [], []
    for _ in range(5):
        ctx.render_plugin(sc, cam, pid, 1920, 1080, block=(8, 8))
        t_plugin.append(ctx.stats().kernel_ms)
        ctx.render(sc, cam, capi.make_params(L.KERNEL_GI, 1920, 1080, max_ray_depth=4), want_output=False)
        t_builtin.append(ctx.stats().kernel_ms)
    print("GI plug-in %.3f ms, built-in %.3f ms" % (min(t_plugin), min(t_builtin)))
    assert min(t_plugin) <= 1.5 * min(t_builtin), (min(t_plugin), min(t_builtin))
    sc.release()
    ctx.close()


def test_gi_plugin_on_a_large_mesh(tmp_path):
    """... and on a tree too large for the threaded copies (stack traversal in the plug-in's dynamic shared memory)."""
    from lens_trace_b200 import capi
    p = str(tmp_path / "synth.obj")
    host.write_synthetic_scene(p, 200, 0x5EED)  # 80 012 triangles, 160 023 nodes
    sb = host.load_scene_buffers(p)
    ctx = capi.Context(0)
    sc = ctx.upload(sb)
    pid = ctx.plugin_load(os.path.join(util.ROOT, "tests", "plugins", "gi_plugin.cu"))
    cam = util.default_camera(0.0, 2)
    for block in ((8, 8), (32, 2)):
        got = ctx.render_plugin(sc, cam, pid, 400, 240, block=block)
        want = ctx.render(sc, cam, capi.make_params(L.KERNEL_GI, 400, 240, max_ray_depth=4))
        util.assert_bit_equal(got, want, "synthetic mesh, block %s: GI plug-in vs built-in" % (block,))
    sc.release()
    ctx.close()
