"""World-size-2 test of the sample split on CPU (gloo): every rank computes its share with the CPU oracle
exactly as a GPU rank would (frames r, r+G, .., weight 1/N), the all-reduce combines them, and the result
equals the sequential running mean of the same frames."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames_per_rank, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import lt_oracle as O
    import util
    from lens_trace_b200 import layouts as L, multi_gpu

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sb = util.scene("cornell_box")
    w, h = 40, 30
    share = multi_gpu.sample_split(rank, world, frames_per_rank)
    acc = np.zeros((h, w, 3), np.float32)
    weight = np.float32(share["accum_weight"])
    for fc in share["frame_counts"]:  # LT_ACCUM_WEIGHTED_SUM: acc += weight * sample, FP32
        sample = O.render(L.KERNEL_ACCUMULATOR, sb, util.default_camera(0.0, fc), w, h)
        acc = (acc + weight * sample).astype(np.float32)
    t = torch.from_numpy(acc)
    multi_gpu.combine(t, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "combined.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_split_partition():
    from lens_trace_b200 import layouts as L, multi_gpu
    seen = []
    for r in range(4):
        s = multi_gpu.sample_split(r, 4, 16)
        assert s["frame_stride"] == 4 and s["frames"] == 16 and s["accum_mode"] == L.ACCUM_WEIGHTED_SUM
        assert abs(s["accum_weight"] - 1 / 64) < 1e-12
        seen += s["frame_counts"]
    assert sorted(seen) == list(range(64))  # every frame exactly once
    one = multi_gpu.sample_split(0, 1, 64)
    assert one["accum_mode"] == L.ACCUM_RUNNING_MEAN and one["frame_counts"] == list(range(64))
    with pytest.raises(ValueError):
        multi_gpu.sample_split(2, 2, 1)


def test_two_rank_split_equals_sequential_mean(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lt_oracle as O
    import util
    from lens_trace_b200 import layouts as L

    world, frames_per_rank = 2, 3
    mp.spawn(_worker, args=(world, _free_port(), frames_per_rank, str(tmp_path)), nprocs=world, join=True)
    combined = np.load(tmp_path / "combined.npy")
    sb = util.scene("cornell_box")
    acc = np.zeros((30, 40, 3), np.float32)
    for f in range(world * frames_per_rank):
        O.accumulate(acc, O.render(L.KERNEL_ACCUMULATOR, sb, util.default_camera(0.0, f), 40, 30), f)
    np.testing.assert_allclose(combined, acc, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("height", [1, 7, 8, 9, 75, 1080, 2160])
@pytest.mark.parametrize("devices", [1, 2, 3, 4, 8, 16])
def test_tile_split_partitions_the_image_rows(height, devices):
    """LT_SPLIT_TILES (lt_multi.cu): blocks of 8 rows dealt round-robin -- every image row belongs to exactly one
    device, local rows ascend, and the shares differ by at most one block."""
    from lens_trace_b200 import capi
    seen = []
    sizes = []
    for g in range(devices):
        rows = capi.tile_rows(height, devices, g)
        assert (np.diff(rows) > 0).all()
        assert ((rows // 8) % devices == g).all()
        seen += rows.tolist()
        sizes.append(len(rows))
    assert sorted(seen) == list(range(height))
    assert max(sizes) - min(sizes) <= 8
    with pytest.raises(capi.LtError):
        capi.tile_rows(height, devices, devices)
