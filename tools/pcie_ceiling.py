#!/usr/bin/env python
"""Raw pinned-copy ceiling of the end-to-end path, per direction and both at once, on every rank at the same time.

    python tools/pcie_ceiling.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Buffer sizes are those of the BASELINE configs (bench.py): what the pipeline moves per frame, with no kernels.
Prints one JSON line per (config, direction) on rank 0.  This is the evidence behind bench.py's e2e.ceiling_gbs.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    import torch
    from disparity_to_point_cloud_b200 import pcie
    rank, world, local = bench.dist_setup()
    torch.cuda.set_device(local)
    shapes = {"config4 3840x2160 f32": (3840 * 2160 * 4, bench.n_points(3840, 2160) * 16),
              "config2 752x480 mono8": (752 * 480, bench.n_points(752, 480) * 16),
              "config5 4 x 1280x720 mono8 -> 665x665 cloud": (4 * 1280 * 720, 5475600)}
    n_bufs = [int(x) for x in sys.argv[1:]] or [3]
    for name, (bi, bo) in shapes.items():
        for direction in ("h2d", "d2h", "both"):
            for n_buf in n_bufs:
                gbs, units = pcie.measure(torch, world, bi, bo, 0.6, bench.barrier_sync, bench.max_over_ranks,
                                          bench.sum_over_ranks, n_buf=n_buf, direction=direction)
                if rank == 0:
                    print(json.dumps({"shape": name, "direction": direction, "n_gpus": world, "n_buf": n_buf,
                                      "GB/s_total": round(gbs, 2), "GB/s_per_gpu": round(gbs / world, 2),
                                      "frames/s_total": round(units, 1)}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
