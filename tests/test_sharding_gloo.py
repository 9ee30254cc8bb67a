"""N > 1 host logic on CPU: two gloo ranks shard a frame list, merge results back in frame order, and reduce a
timing with max-over-ranks the way bench.py does."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from disparity_to_point_cloud_b200 import sharding


def test_partition_is_exact():
    for n, g in [(1024, 1), (1024, 2), (1024, 8), (1001, 8), (3, 8), (0, 4)]:
        owned = [sharding.frames_of_rank(n, r, g) for r in range(g)]
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(n))
        assert sharding.frames_per_rank(n, g, "strong") == [len(o) for o in owned]
        assert sharding.frames_per_rank(n, g, "weak") == [n] * g
        merged = sharding.merge_by_frame_index([[f"f{i}" for i in o] for o in owned], n, g)
        assert merged == [f"f{i}" for i in range(n)]
    with pytest.raises(ValueError):
        sharding.frames_of_rank(10, 3, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.frames_of_rank(n_frames, rank, world)
        # stand-in for the per-frame GPU work: the "cloud" of frame i is its index squared
        results = [i * i for i in mine]
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        merged = sharding.merge_by_frame_index(gathered, n_frames, world)
        assert merged == [i * i for i in range(n_frames)]
        dist.barrier()
        t = sharding.max_over_ranks(10.0 + rank)          # the slowest rank defines the step time
        assert t == 10.0 + world - 1
        total = sharding.sum_over_ranks(len(mine))
        assert total == n_frames
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 37), nprocs=2, join=True)
