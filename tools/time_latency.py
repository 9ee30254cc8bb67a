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
        # the message a subscriber publishes: library-owned cloud + the copy into the message (what pcl::toROSMsg
        # does in the reference, cpp:84-85) versus the cloud DMA'd straight into a page-locked message buffer
        n = oracle.n_points(w, h) * 16
        msg = np.empty(n, np.uint8)

        def with_copy():
            msg[:] = ctx.process_mono8(img, copy=False)

        reg = d2pc.RegisteredArray(np.empty(n, np.uint8))
        cp = lat(with_copy)
        into = lat(lambda: ctx.process_into(img, reg.array))
        reg.free()
        print(f"{w}x{h}: mono8 -> message buffer: library cloud + memcpy {cp[0]:.0f} us (p99 {cp[1]:.0f}), "
              f"d2pc_process_mono8_into a registered buffer {into[0]:.0f} us (p99 {into[1]:.0f})", flush=True)
        t0 = time.perf_counter()
        for _ in range(3):
            oracle.disparity_cb_mono8(img, q)
        cpu = (time.perf_counter() - t0) / 3 * 1e6
        print(f"{w}x{h}: process_mono8 median {m8[0]:.0f} us (p99 {m8[1]:.0f}), process_f32 median {f32[0]:.0f} us "
              f"(p99 {f32[1]:.0f}); CPU oracle callback (1 thread) {cpu:.0f} us", flush=True)
