"""Times the matching-score preprocessing (device entry, both callbacks) and the fusion merge at 1280x720."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context(offset_x=-7, offset_y=15)
stream = torch.cuda.ExternalStream(ctx.compute_stream())


def t(fn, it=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(it):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


for (w, h) in [(1280, 720), (752, 480)]:
    s = torch.from_numpy(synth.s2_scene(h, w, 1)).cuda()
    n = ctx.fuse_geometry(w, h)[4][0]
    o = torch.empty((n, n), dtype=torch.uint8, device="cuda")
    for which in (1, 2):
        us = t(lambda: ctx.preprocess_score_device(s.data_ptr(), w, h, w, which, o.data_ptr()))
        print(f"{w}x{h} MatchingScoreCb{which} (n = {n}): {us:.1f} us", flush=True)
