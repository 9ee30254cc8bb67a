#!/usr/bin/env python
"""CROP kernel: rows per work unit x L2 prefetch distance, per frame size (python tools/rows_sweep.py [rows,..] [dist,..] [ctas,..])."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import torch  # noqa: E402

import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
import timing  # noqa: E402

ctx = d2pc.Context()
t = timing.timer(ctx)
rows_list = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 2, 4, 6, 8, 12, 16, 32]
dist_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
ctas_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
for (w, h, f) in [(1280, 720, 64), (752, 480, 256), (640, 480, 256), (3840, 2160, 16)]:
    n = (w - 80) * (h - 80)
    d_in = timing.float_batch(w, h, f)
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    for rows in rows_list:
        for dist in dist_list:
            for ctas in ctas_list:
                ctx.set_tuning("rows_per_unit", rows)
                ctx.set_tuning("prefetch_dist", dist)
                ctx.set_tuning("ctas_per_sm", ctas)
                s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16), 30)
                print(w, h, f, "rows_per_unit", rows or "auto", "prefetch", dist or "auto", "ctas_per_sm", ctas or "auto",
                      "%.1f us  frac %.3f" % (s * 1e6, 20 * n * f / s / 1e9 / timing.PEAK), flush=True)
    ctx.set_tuning("rows_per_unit", 0)
    ctx.set_tuning("prefetch_dist", 0)
    ctx.set_tuning("ctas_per_sm", 0)
