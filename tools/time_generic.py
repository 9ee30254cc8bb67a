"""Times the generic-Q exact CROP path (force_generic) -- development tool."""
import sys

import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
stream = torch.cuda.ExternalStream(ctx.compute_stream())


def t(fn, it=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(it):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3


for (w, h, f) in [(3840, 2160, 16), (1280, 720, 64)]:
    n = (w - 80) * (h - 80)
    base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
    d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    for gen in (0, 1):
        ctx.set_tuning("force_generic", gen)
        s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16))
        print(w, h, f, "generic" if gen else "rectified", "%.1f us  frac %.3f" % (s * 1e6, 20 * n * f / s / 1e9 / 6534.8), flush=True)
    ctx.set_tuning("force_generic", 0)
    del d_in, d_out
