"""Multi-GPU contexts of the C-ABI (lt_ctx_create_multi) and of the C++ surface (RenderExtensionB200::deviceCount):
the sample split with one all-reduce per call (NCCL, and this repo's own peer-memory kernel) and the tile split,
against one GPU rendering the same call.  Needs two or more devices (`gpurun --gpus 2`); skipped otherwise.

Tolerance: the sample split sums per-device partial sums, the single GPU keeps a sequential running mean -- the
two differ by FP32 rounding order only; bound 1e-5 relative (measured ~2e-7).  The tile split is bit-exact."""
import os
import subprocess

import numpy as np
import pytest

import util
from lens_trace_b200 import capi, host, layouts as L

pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


needs_two = pytest.mark.skipif("device_count() < 2", reason="needs two or more GPUs")


def rel_diff(a, b):
    return float((np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-3)).max())


@needs_two
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_sample_split_equals_one_gpu(exchange, monkeypatch):
    monkeypatch.setenv("LT_MULTI_EXCHANGE", exchange)
    n = min(device_count(), 8)
    sb = util.scene("cornell_box")
    one = capi.Context(0)
    s1 = one.upload(sb)
    many = capi.Context(list(range(n)))
    assert many.device_count() == n
    sm = many.upload(sb)
    cam = util.default_camera(0.0, 0)
    for kernel, depth, w, h, frames in ((L.KERNEL_GI, 4, 640, 360, 16), (L.KERNEL_ACCUMULATOR, 0, 333, 111, 5),
                                        (L.KERNEL_GI, 3, 1920, 1080, 8)):
        p = dict(max_ray_depth=depth, frames=frames, accum_mode=L.ACCUM_RUNNING_MEAN)
        one.accum_reset()
        want = one.render(s1, cam, capi.make_params(kernel, w, h, **p))
        many.accum_reset()
        got = many.render(sm, cam, capi.make_params(kernel, w, h, split_mode=L.SPLIT_SAMPLES, **p))
        assert rel_diff(got, want) <= 1e-5, (exchange, kernel, rel_diff(got, want))
        # progressive: the same frames in two calls (the second continues the accumulator at frameCount = half)
        half = frames // 2
        many.accum_reset()
        many.render(sm, cam, capi.make_params(kernel, w, h, split_mode=L.SPLIT_SAMPLES, max_ray_depth=depth, frames=half,
                                             accum_mode=L.ACCUM_RUNNING_MEAN))
        got2 = many.render(sm, util.default_camera(0.0, half),
                           capi.make_params(kernel, w, h, split_mode=L.SPLIT_SAMPLES, max_ray_depth=depth,
                                            frames=frames - half, accum_mode=L.ACCUM_RUNNING_MEAN))
        assert rel_diff(got2, want) <= 1e-5
    # a sample split of a non-accumulating call is refused; AUTO picks tiles for it
    with pytest.raises(capi.LtError):
        many.render(sm, cam, capi.make_params(L.KERNEL_GI, 64, 64, split_mode=L.SPLIT_SAMPLES))
    auto = many.render(sm, cam, capi.make_params(L.KERNEL_GI, 64, 48, max_ray_depth=2))
    util.assert_bit_equal(auto, one.render(s1, cam, capi.make_params(L.KERNEL_GI, 64, 48, max_ray_depth=2)))
    sm.release(); many.close(); s1.release(); one.close()


@needs_two
def test_tile_split_is_bit_exact():
    n = min(device_count(), 8)
    one = capi.Context(0)
    many = capi.Context(list(range(n)))
    for name in ("cornell_box", "cornell_box_lens"):
        sb = util.scene(name)
        s1, sm = one.upload(sb), many.upload(sb)
        for w, h in ((1920, 1080), (101, 75), (64, 5), (33, 8 * n + 3)):
            for kernel, depth, frames in ((L.KERNEL_BASIC_CU, 0, 1), (L.KERNEL_CUSTOM_BARY, 0, 3), (L.KERNEL_GI, 3, 4),
                                          (L.KERNEL_GI25, 2, 1)):
                if kernel == L.KERNEL_GI25 and w * h > 20000:
                    continue
                cam = util.default_camera(0.03, 1)
                acc = L.ACCUM_RUNNING_MEAN if frames > 1 else L.ACCUM_NONE
                p = dict(max_ray_depth=depth, frames=frames, accum_mode=acc)
                one.accum_reset()
                want = one.render(s1, cam, capi.make_params(kernel, w, h, **p))
                many.accum_reset()
                got = many.render(sm, cam, capi.make_params(kernel, w, h, split_mode=L.SPLIT_TILES, **p))
                util.assert_bit_equal(got, want, "%s kernel %d %dx%d tiles over %d GPUs" % (name, kernel, w, h, n))
        s1.release(); sm.release()
    many.close(); one.close()


@needs_two
def test_multi_gpu_through_the_renderer_classes(root):
    """RendererOpenCL::render with RenderExtensionB200{deviceCount, splitMode}: same picture as one device."""
    os.chdir(root)
    n = min(device_count(), 8)
    cam = host.Camera(0, 2.5, -50, 0)
    model = host.Model("resources/models/cornell_box.obj")
    accel = host.AccelerationStructure(model)
    r = host.Renderer(host.PLATFORM_OPENCL)
    K = "examples/global_illumination/resources/kernels/global_illumination.cl"
    w, h = 320, 200
    want = r.render(K, w, h, accel, model, cam, ext=host.make_extension(frames=8, accumulate=True, max_ray_depth=4))
    for split in (1, 2, 0):
        cam.set_frame_count(0)
        got = r.render(K, w, h, accel, model, cam,
                       ext=host.make_extension(frames=8, accumulate=True, max_ray_depth=4, devices=n, split=split))
        if split == 2:
            util.assert_bit_equal(got, want, "tile split through RendererOpenCL")
        else:
            assert rel_diff(got, want) <= 1e-5
    r.close(); accel.close(); model.close(); cam.close()


@needs_two
def test_multi_gpu_example_program(root, tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(root, "examples"), "-s"])
    n = min(device_count(), 8)
    for split in ("samples", "tiles"):
        r = subprocess.run([os.path.join(root, "examples", "bin", "multi_gpu"), "--devices", str(n), "--size", "640", "360",
                            "--frames", "16", "--split", split], cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "GPUs:" in r.stdout


def test_one_device_group_is_a_plain_context():
    c = capi.Context([0])
    assert c.device_count() == 1
    sb = util.scene("green_wall")
    sc = c.upload(sb)
    img = c.render(sc, util.default_camera(), capi.make_params(L.KERNEL_BASIC_CU, 32, 32))
    assert img[16, 16].tolist() == [0.0, 1.0, 0.0]
    sc.release(); c.close()
    with pytest.raises(capi.LtError):
        capi.Context([0, 0])
