"""Synchronous per-call latency of the host entry points (one frame in, one cloud out) -- what one camera stream sees."""
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
import oracle  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402


def lat(fn, n=200):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return statistics.median(ts) * 1e6, ts[int(0.99 * n)] * 1e6


with d2pc.Context() as ctx:
    q = ctx.get_q()
    for (w, h) in [(640, 480), (752, 480), (1280, 720)]:
        img = synth.s2_scene(h, w, 1)
        d = synth.s3_float(h, w, 1)
        m8 = lat(lambda: ctx.process_mono8(img, copy=False))
        f32 = lat(lambda: ctx.process_f32(d, copy=False))
        t0 = time.perf_counter()
        for _ in range(3):
            oracle.disparity_cb_mono8(img, q)
        cpu = (time.perf_counter() - t0) / 3 * 1e6
        print(f"{w}x{h}: process_mono8 median {m8[0]:.0f} us (p99 {m8[1]:.0f}), process_f32 median {f32[0]:.0f} us "
              f"(p99 {f32[1]:.0f}); CPU oracle callback (1 thread) {cpu:.0f} us", flush=True)
