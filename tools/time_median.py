"""Times the mono8 callback (median 11 + reproject) and the median alone for both median kernels."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
stream = torch.cuda.ExternalStream(ctx.compute_stream())


def t(fn, it=10):
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


for (w, h, f, kind) in [(752, 480, 256, "s2"), (752, 480, 256, "s1"), (3840, 2160, 8, "s2"), (752, 480, 1, "s2")]:
    gen = synth.s2_scene if kind == "s2" else synth.s1_uniform
    d = torch.from_numpy(np.stack([gen(h, w, i) for i in range(min(f, 8))])).cuda().repeat(max(1, f // 8), 1, 1)[:f].contiguous()
    n = (w - 80) * (h - 80)
    o = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    for variant in (0, 3, 2):
        ctx.set_tuning("median_variant", variant)
        for strip in ([0, 32, 128] if variant == 0 and f > 1 else ([0, 8] if variant == 0 else [0])):
            ctx.set_tuning("median_strip", strip)
            s_all = t(lambda: ctx.reproject_mono8_device(d.data_ptr(), f, w, h, w, w * h, o.data_ptr(), n * 16))
            m = torch.empty_like(d)
            s_med = t(lambda: [ctx.median_u8_device(d[i].data_ptr(), w, h, w, m[i].data_ptr(), w, 11) for i in range(min(f, 4))]) / min(f, 4)
            print(w, h, f, kind, "variant", variant, "strip", strip,
                  "callback: %.1f us/frame  %.1f Gpix/s | full-frame median alone: %.1f us/frame" %
                  (s_all / f * 1e6, f * w * h / s_all / 1e9, s_med * 1e6), flush=True)
    ctx.set_tuning("median_strip", 0)
