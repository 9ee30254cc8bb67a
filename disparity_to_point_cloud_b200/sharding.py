"""Frame-level sharding across GPUs: frames are independent, so frame i goes to rank
i mod G with no data-path collective.  The only collectives are the bench's barrier and its max-over-ranks of the
timed interval.  Works on any torch.distributed backend (nccl on the box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List


def frames_of_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Global frame indices rank `rank` owns under the i mod G rule (order preserved per rank)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def frames_per_rank(n_frames: int, world: int, scaling: str) -> List[int]:
    """How many frames each rank processes per step: weak = n_frames each, strong = n_frames split i mod G."""
    if scaling == "weak":
        return [n_frames] * world
    if scaling == "strong":
        return [len(range(r, n_frames, world)) for r in range(world)]
    raise ValueError("scaling must be 'weak' or 'strong'")


def merge_by_frame_index(per_rank_results, n_frames: int, world: int):
    """Inverse of frames_of_rank: per_rank_results[r][k] belongs to frame r + k*world."""
    out = [None] * n_frames
    for r, results in enumerate(per_rank_results):
        for k, item in enumerate(results):
            out[r + k * world] = item
    if any(o is None for o in out):
        raise ValueError("missing frames")
    return out


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
