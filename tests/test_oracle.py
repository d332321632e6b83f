"""CPU tests of the oracle: the reference's own known-answer test and invariants, pinned RNG values,
and the reference-generated golden vectors under tests/golden/ (made on the GPU box by the reference's
CUDA backend, tools/make_reference_golden.py)."""
import glob
import os

import numpy as np
import pytest

import lt_oracle as O
import util
from lens_trace_b200 import layouts as L


def test_correct_color_known_answer():
    # tests/cuda_renderer_test.cc:182-225 -- green wall, camera (0,2.5,-50), 100x100x3:
    # floats [x..x+2] == (0,1,0) for x = 0,24,48.. < 10000
    img = O.render(L.KERNEL_BASIC_CU, util.scene("green_wall"), util.default_camera(), 100, 100)
    flat = img.reshape(-1)
    for x in range(0, 100 * 100, 8 * 3):
        assert flat[x + 0] == 0.0 and flat[x + 1] == 1.0 and flat[x + 2] == 0.0
    # same answer from the OpenCL text of the kernel
    img_cl = O.render(L.KERNEL_BASIC_CL, util.scene("green_wall"), util.default_camera(), 100, 100)
    util.assert_bit_equal(img, img_cl, "basic.cu vs basic.cl")


def test_kernel_mode_invariance():
    # tests/cuda_renderer_test.cc:117-180 -- linearKernel == tileKernel
    sb = util.scene("cornell_box")
    a = O.render(L.KERNEL_BASIC_CU, sb, util.default_camera(), 100, 100, kernel_mode=0)
    b = O.render(L.KERNEL_BASIC_CU, sb, util.default_camera(), 100, 100, kernel_mode=1)
    util.assert_bit_equal(a, b)


def test_multi_prim_leaf_quirk():
    # basic.cu:168-172: a leaf with primitiveCount 2 only ever tests primitives[primitivesOffset]
    sb = util.multi_prim_leaf_scene()
    ids, hit, tuv, st = O.primary_hits(0, sb, util.default_camera(), 64, 64)
    assert set(np.unique(ids[hit == 1]).tolist()) == {0}
    assert 0 < (hit == 1).mean() < 0.9  # half of the wall is unreachable
    assert st.triTests == 2 * st.rays  # tested twice per ray, both times triangle 0


def test_cornell_colors_and_stats():
    sb = util.scene("cornell_box")
    img, st = O.render(L.KERNEL_BASIC_CU, sb, util.default_camera(), 128, 128, with_stats=True)
    colors = {tuple(round(float(v), 4) for v in c) for c in np.unique(img.reshape(-1, 3), axis=0)}
    assert colors == {(0, 0, 0), (0, 1, 0), (1, 0, 0), (0.8, 0.8, 0.8), (1, 1, 1)}
    assert st.rays == 128 * 128
    assert st.nodeTests >= st.rays and st.triTests > 0


def test_lens_scene_refracts():
    sb = util.scene("cornell_box_lens")
    img, st = O.render(L.KERNEL_BASIC_CU, sb, util.default_camera(), 96, 96, with_stats=True)
    assert st.rays > 96 * 96  # lens pixels trace two more segments (basic.cu:271,297)
    assert np.isfinite(img).all()


def test_custom_kernel_barycentrics():
    sb = util.scene("cornell_box")
    img = O.render(L.KERNEL_CUSTOM_BARY, sb, util.default_camera(), 64, 64)
    ids, hit, tuv, _ = O.primary_hits(1, sb, util.default_camera(), 64, 64)
    m = hit == 1
    np.testing.assert_array_equal(img[..., 0][m], tuv[..., 1][m])
    np.testing.assert_array_equal(img[..., 1][m], tuv[..., 2][m])
    np.testing.assert_allclose(img[m].sum(axis=1), 1.0, atol=2e-7)
    assert (img[~m] == 0).all()


def test_random_matches_its_definition():
    # basic_lighting.cl:64-67 evaluated independently with numpy float64/float32
    rng = np.random.default_rng(7)
    for _ in range(200):
        u, v = np.float32(rng.uniform(-0.5, 0.5)), np.float32(rng.uniform(-0.5, 0.5))
        seed = np.float32(rng.integers(0, 5000))
        d = np.float32(np.float32(u * np.float32(12.9898)) + np.float32(v * np.float32(78.233)))
        x = np.float64(d) + np.float64(1113.1) * np.float64(seed)
        a = np.float32(np.sin(np.fmod(x, np.pi)) * 43758.5453)
        want = np.float32(a - np.floor(a))
        got = np.float32(O.random(float(u), float(v), float(seed)))
        assert got == want and 0.0 <= got <= 1.0


def test_cuda_trig_restatement_is_accurate():
    lib = O.load()
    for x in np.linspace(-20, 20, 401, dtype=np.float32):
        assert abs(lib.lto_cuda_cosf(float(x)) - np.cos(np.float64(x))) < 3e-7
        assert abs(lib.lto_cuda_sinf(float(x)) - np.sin(np.float64(x))) < 3e-7
    assert lib.lto_cuda_cosf(0.0) == 1.0 and lib.lto_cuda_sinf(0.0) == 0.0


def test_accumulate_is_the_shader_formula():
    # accumulator.frag:10-19
    rng = np.random.default_rng(3)
    acc = np.zeros(32, np.float32)
    samples = rng.random((5, 32), dtype=np.float32)
    for k in range(5):
        want = samples[k].copy() if k == 0 else (samples[k] + acc * np.float32(k)) / np.float32(k + 1)
        O.accumulate(acc, samples[k], k)
        util.assert_bit_equal(acc, want.astype(np.float32))
    np.testing.assert_allclose(acc, samples.mean(axis=0), rtol=1e-6)


def test_gi_sample_is_deterministic_and_bounded():
    sb = util.scene("cornell_box")
    cam = util.default_camera(frame_count=3)
    a, st = O.render(L.KERNEL_GI, sb, cam, 48, 48, max_ray_depth=4, with_stats=True)
    b = O.render(L.KERNEL_GI, sb, cam, 48, 48, max_ray_depth=4, threads=4)
    util.assert_bit_equal(a, b, "thread count must not matter")
    assert np.isfinite(a).all()
    per_pixel = st.rays / (48 * 48)
    assert 1.0 <= per_pixel <= 2 + 2 * 4
    c = O.render(L.KERNEL_GI, sb, util.default_camera(frame_count=4), 48, 48, max_ray_depth=4)
    assert (a != c).any()  # the seed is frameCount


def test_blend25_clamps_only_in_linear_mode():
    sb = util.scene("cornell_box")
    lin = O.render(L.KERNEL_GI25, sb, util.default_camera(frame_count=1), 24, 24, kernel_mode=0, max_ray_depth=2)
    til = O.render(L.KERNEL_GI25, sb, util.default_camera(frame_count=1), 24, 24, kernel_mode=1, max_ray_depth=2)
    assert lin.min() >= 0 and lin.max() <= 1
    util.assert_bit_equal(lin, np.clip(til, 0, 1))


def test_rows_argument_renders_a_band():
    sb = util.scene("cornell_box")
    full = O.render(L.KERNEL_ACCUMULATOR, sb, util.default_camera(frame_count=2), 40, 40)
    band = O.render(L.KERNEL_ACCUMULATOR, sb, util.default_camera(frame_count=2), 40, 40, rows=(10, 20))
    util.assert_bit_equal(full[10:20], band[10:20])
    assert (band[:10] == 0).all() and (band[20:] == 0).all()


def _golden_files():
    return sorted(glob.glob(os.path.join(util.GOLDEN, "ref_cuda_*.npz")))


@pytest.mark.parametrize("path", _golden_files() or [None])
def test_oracle_against_reference_generated_golden(path):
    """tests/golden/ref_cuda_*.npz hold buffers built by the reference's own builder and the outputs of
    the reference's CUDA backend (NVRTC build of basic.cu) on a B200; the oracle must reproduce them
    bit for bit."""
    if path is None:
        pytest.skip("no reference-generated golden files committed yet")
    z = np.load(path)
    sb = L.SceneBuffers(z["nodes"], z["prims"], z["materials"], z["lights"])
    cam = z["camera"].view(L.CAMERA)
    w, h = int(z["width"]), int(z["height"])
    img = O.render(L.KERNEL_BASIC_CU, sb, cam, w, h)
    util.assert_bit_equal(img, z["color"], os.path.basename(path) + " colour")
    if "ids" in z:
        ids, hit, tuv, _ = O.primary_hits(0, sb, cam, w, h)
        m = z["hit"] == 1
        np.testing.assert_array_equal(hit, z["hit"])
        np.testing.assert_array_equal(ids[m], z["ids"][m])
        util.assert_bit_equal(tuv[m], z["tuv"][m], os.path.basename(path) + " t,u,v")
