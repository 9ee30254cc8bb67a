#!/usr/bin/env python
"""Kernel-only timing sweep over the tuning knobs (rows_per_unit, ctas_per_sm, arithmetic) -- development tool.
Prints one line per combination: config, us/launch, GB/s (algorithmic), fraction of measured HBM peak."""
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(stream, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    shapes = [(3840, 2160, 16), (1280, 720, 64)]
    rows = [int(x) for x in os.environ.get("SWEEP_ROWS", "8,16,32").split(",")]
    ctas = [int(x) for x in os.environ.get("SWEEP_CTAS", "4,6,7,8").split(",")]
    ariths = os.environ.get("SWEEP_ARITH", "exact,fast,generic").split(",")
    ctx = d2pc.Context()
    stream = torch.cuda.ExternalStream(ctx.compute_stream())
    for w, h, f in shapes:
        n = (w - 80) * (h - 80)
        base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
        d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
        d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
        for ar, r, c in itertools.product(ariths, rows, ctas):
            ctx.set_arith_mode(d2pc.ARITH_FAST if ar == "fast" else d2pc.ARITH_EXACT)
            ctx.set_tuning("force_generic", 1 if ar == "generic" else 0)
            ctx.set_tuning("rows_per_unit", r)
            ctx.set_tuning("ctas_per_sm", c)
            s = timeit(stream, lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4,
                                                                d_out.data_ptr(), n * 16))
            gbs = 20 * n * f / s / 1e9
            print(f"{w}x{h}x{f} arith={ar:7s} rows={r:2d} ctas/sm={c} : {s*1e6:8.1f} us  {gbs:7.1f} GB/s  "
                  f"frac={gbs/PEAK:.3f}", flush=True)
        del d_in, d_out
    ctx.close()


if __name__ == "__main__":
    main()
