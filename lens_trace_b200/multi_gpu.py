"""Sample split across the GPUs of one box: rank r of G renders frames r, r+G, ... of an
N = G*frames_per_rank sample frame into its own FP32 accumulator with weight 1/N
(LT_ACCUM_WEIGHTED_SUM, frame_stride = G); one all-reduce(sum) per frame -- the only exchange step of the
path -- turns the partial sums into the mean.  The scene is replicated on every GPU."""
from . import layouts as L


def sample_split(rank, world, frames_per_rank, first_frame_count=0):
    """Parameters of one rank's share: which frameCount values it renders and how they are weighted."""
    if not (0 <= rank < world) or frames_per_rank < 1:
        raise ValueError("bad rank/world/frames")
    total = frames_per_rank * world
    return {
        "camera_frame_count": first_frame_count + rank,
        "frame_stride": world,
        "frames": frames_per_rank,
        "accum_mode": L.ACCUM_WEIGHTED_SUM if world > 1 else L.ACCUM_RUNNING_MEAN,
        "accum_weight": 1.0 / total,
        "frame_counts": [first_frame_count + rank + k * world for k in range(frames_per_rank)],
    }


def combine(acc, world):
    """The exchange step: in-place sum over ranks (NCCL on GPUs, gloo in the CPU tests)."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(acc)
    return acc
