"""Times CROP_FINITE kernels (classify-first vs park) on resident batches -- development tool."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
stream = torch.cuda.ExternalStream(ctx.compute_stream())


def t(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(it):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3


for (w, h, f, kind) in [(1280, 720, 64, "s3"), (1280, 720, 64, "s2"), (3840, 2160, 16, "s3")]:
    n = (w - 80) * (h - 80)
    if kind == "s3":
        base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
    else:
        base = torch.from_numpy(synth.s2_scene(h, w, 3).astype(np.float32) * np.float32(0.125)).cuda()
    d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
    ctx.set_filter_mode(1)
    for park, mb in ((0, 0), (3, 4), (2, 4), (1, 4)):
        ctx.set_tuning("force_park", park)
        ctx.set_tuning("ctas_per_sm", mb)
        s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16,
                                               d_cnt.data_ptr()))
        kept = int(d_cnt.sum().item())
        by = 4 * n * f + 16 * kept
        print(w, h, f, kind, ["two-pass", "park", "classify", "band"][park], "minB", mb,
              "%.1f us  %.1f GB/s frac %.3f kept %.3f" % (s * 1e6, by / s / 1e9, by / s / 1e9 / 6534.8, kept / (n * f)))
    ctx.set_tuning("force_park", 0)
    ctx.set_tuning("ctas_per_sm", 0)
    ctx.set_filter_mode(0)
    del d_in, d_out
