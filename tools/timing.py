#!/usr/bin/env python
"""Development timings, one tool (none of it is product code, none of it imports the oracle):

    python tools/timing.py crop [prefetch distances...]   headline CROP kernel vs L2 prefetch distance, 4K x16 / 720p x64
    python tools/timing.py compact                        CROP_FINITE: band kernel vs the park fallback
    python tools/timing.py generic                        rectified vs generic-Q exact arithmetic
    python tools/timing.py median                         mono8 callback + median alone, per median variant / strip
    python tools/timing.py score                          MatchingScoreCb1/2 (device entry)
    python tools/timing.py latency                        synchronous per-call latency of the host entry points
    python tools/timing.py direct [calls]                 A/B of direct_out (kernel writes the cloud into host memory) per call
    python tools/timing.py callbacks                      per-callback latency of the fusion node's synchronous entries
    python tools/timing.py stream                         end-to-end stream throughput vs pipeline depth
    python tools/timing.py fusion [slots...]              config 5: frame sets / s vs slots, spans of one set, submit cost
    python tools/timing.py numer                          cost of an integral principal point, both kernel variants

Kernel timings use CUDA events on the context's compute stream after warm-up.
"""
import ctypes
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

PEAK = 6534.8  # GB/s, MEASURED_PEAKS.json on this pool


def timer(ctx):
    stream = torch.cuda.ExternalStream(ctx.compute_stream())

    def t(fn, it=20, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(it):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it * 1e-3
    return t


def float_batch(w, h, f, kind="s3"):
    base = synth.s3_float(h, w, 3) if kind == "s3" else synth.s2_scene(h, w, 3).astype(np.float32) * np.float32(0.125)
    base = torch.from_numpy(base).cuda()
    return torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()


def crop(argv):
    ctx = d2pc.Context()
    t = timer(ctx)
    dists = [int(x) for x in argv] or [0, 256, 512, 1024, 2048]
    for (w, h, f) in [(3840, 2160, 16), (1280, 720, 64)]:
        n = (w - 80) * (h - 80)
        d_in = float_batch(w, h, f)
        d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
        for dist in dists:
            ctx.set_tuning("prefetch_dist", dist)
            s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16), 30)
            print(w, h, f, "prefetch", dist, "%.1f us  frac %.3f" % (s * 1e6, 20 * n * f / s / 1e9 / PEAK), flush=True)


def compact(argv):
    ctx = d2pc.Context()
    t = timer(ctx)
    for (w, h, f, kind) in [(1280, 720, 64, "s3"), (1280, 720, 64, "s2"), (3840, 2160, 16, "s3")]:
        n = (w - 80) * (h - 80)
        d_in = float_batch(w, h, f, kind)
        d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
        ctx.set_filter_mode(1)
        for variant, knobs in ((0, {}), (0, {"rows_per_unit": 4}), (0, {"ctas_per_sm": 3}), (1, {})):
            ctx.set_tuning("compact_variant", variant)
            for k, v in knobs.items():
                ctx.set_tuning(k, v)
            s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16,
                                                   d_cnt.data_ptr()))
            for k in knobs:
                ctx.set_tuning(k, 0)
            kept = int(d_cnt.sum().item())
            by = 4 * n * f + 16 * kept
            print(w, h, f, kind, "band" if variant == 0 else "park", knobs,
                  "%.1f us  %.1f GB/s frac %.3f kept %.3f" % (s * 1e6, by / s / 1e9, by / s / 1e9 / PEAK, kept / (n * f)), flush=True)
        ctx.set_tuning("compact_variant", 0)
        ctx.set_filter_mode(0)


def numer(argv):
    """Cost of the degenerate numerator column / row (u == cx, v == cy): the same kernels with the principal point
    moved out of the image, where no pixel has a zero numerator."""
    ctx = d2pc.Context()
    t = timer(ctx)
    for (w, h, f) in [(1280, 720, 64), (3840, 2160, 16), (752, 480, 256)]:
        n = (w - 80) * (h - 80)
        d_in = float_batch(w, h, f)
        d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
        for mode in (0, 1):
            ctx.set_filter_mode(mode)
            for name, (cx, cy), zn in (("cx,cy inside", (376.0, 240.0), 0), ("cx,cy inside, ordinary kernel", (376.0, 240.0), -1),
                                       ("cx,cy outside", (-5000.5, -5000.5), 0),
                                       ("cx,cy outside, zero-numerator kernel", (-5000.5, -5000.5), 1)):
                q = np.array([[1, 0, 0, -cx], [0, 1, 0, -cy], [0, 0, 0, 714.24], [0, 0, 1 / 0.09, 0]], dtype=np.float64)
                ctx.set_q(q)
                ctx.set_tuning("zero_numer", zn)
                s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16,
                                                       d_cnt.data_ptr() if mode else 0))
                kept = int(d_cnt.sum().item()) if mode else n * f
                by = 4 * n * f + 16 * kept
                print(w, h, f, "CROP_FINITE" if mode else "CROP", name, "%.1f us  frac %.3f" % (s * 1e6, by / s / 1e9 / PEAK), flush=True)
        ctx.set_filter_mode(0)
        ctx.set_tuning("zero_numer", 0)


def generic(argv):
    ctx = d2pc.Context()
    t = timer(ctx)
    for (w, h, f) in [(3840, 2160, 16), (1280, 720, 64)]:
        n = (w - 80) * (h - 80)
        d_in = float_batch(w, h, f)
        d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
        for gen, ctas in [(0, 0), (1, 0)] + [(1, int(c)) for c in argv]:
            ctx.set_tuning("force_generic", gen)
            ctx.set_tuning("ctas_per_sm", ctas)
            s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16))
            print(w, h, f, "generic" if gen else "rectified", "ctas_per_sm", ctas or "default",
                  "%.1f us  frac %.3f" % (s * 1e6, 20 * n * f / s / 1e9 / PEAK), flush=True)
        ctx.set_tuning("force_generic", 0)
        ctx.set_tuning("ctas_per_sm", 0)


def median(argv):
    ctx = d2pc.Context()
    t = timer(ctx)
    variants = [int(x) for x in argv] or [0, 2]
    for (w, h, f, kind) in [(752, 480, 256, "s2"), (752, 480, 256, "s1"), (3840, 2160, 8, "s2"), (752, 480, 1, "s2")]:
        gen = synth.s2_scene if kind == "s2" else synth.s1_uniform
        d = torch.from_numpy(np.stack([gen(h, w, i) for i in range(min(f, 8))])).cuda().repeat(max(1, f // 8), 1, 1)[:f].contiguous()
        n = (w - 80) * (h - 80)
        o = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
        m = torch.empty_like(d)
        for variant in variants:
            ctx.set_tuning("median_variant", variant)
            ctx.set_tuning("fuse_median", -1)
            s_two = t(lambda: ctx.reproject_mono8_device(d.data_ptr(), f, w, h, w, w * h, o.data_ptr(), n * 16), 10, 3)
            ctx.set_tuning("fuse_median", 0)
            print(w, h, f, kind, "variant", variant, "callback as two launches: %.1f us/frame  %.1f Gpix/s"
                  % (s_two / f * 1e6, f * w * h / s_two / 1e9), flush=True)
            s_all = t(lambda: ctx.reproject_mono8_device(d.data_ptr(), f, w, h, w, w * h, o.data_ptr(), n * 16), 10, 3)
            k = min(f, 4)
            s_med = t(lambda: [ctx.median_u8_device(d[i].data_ptr(), w, h, w, m[i].data_ptr(), w, 11) for i in range(k)], 10, 3) / k
            print(w, h, f, kind, "variant", variant, "callback: %.1f us/frame  %.1f Gpix/s | full-frame median alone: %.1f us/frame"
                  % (s_all / f * 1e6, f * w * h / s_all / 1e9, s_med * 1e6), flush=True)
        ctx.set_tuning("median_variant", 0)


def score(argv):
    ctx = d2pc.Context(offset_x=-7, offset_y=15)
    t = timer(ctx)
    for (w, h) in [(1280, 720), (752, 480)]:
        s = torch.from_numpy(synth.s2_scene(h, w, 1)).cuda()
        n = ctx.fuse_geometry(w, h)[4][0]
        o = torch.empty((n, n), dtype=torch.uint8, device="cuda")
        for which in (1, 2):
            us = t(lambda: ctx.preprocess_score_device(s.data_ptr(), w, h, w, which, o.data_ptr()), 50) * 1e6
            print(f"{w}x{h} MatchingScoreCb{which} (n = {n}): {us:.1f} us", flush=True)


def latency(argv):
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

    # argv[0]: direct_out tuning (-1 = D2H copy after the kernel, 0 = the default: the kernel of a synchronous call
    # writes the cloud straight into page-locked host memory)
    direct = int(argv[0]) if argv else 0
    with d2pc.Context() as ctx:
        ctx.set_tuning("direct_out", direct)
        print(f"direct_out = {direct}", flush=True)
        for (w, h) in [(640, 480), (752, 480), (1280, 720)]:
            img = synth.s2_scene(h, w, 1)
            d = synth.s3_float(h, w, 1)
            m8 = lat(lambda: ctx.process_mono8(img, copy=False))
            f32 = lat(lambda: ctx.process_f32(d, copy=False))
            # the message a subscriber publishes: library-owned cloud + the copy into the message (what pcl::toROSMsg
            # does in the reference, cpp:84-85) versus the cloud DMA'd straight into a page-locked message buffer
            n = (w - 80) * (h - 80) * 16
            msg = np.empty(n, np.uint8)

            def with_copy():
                msg[:] = ctx.process_mono8(img, copy=False)

            reg = d2pc.RegisteredArray(np.empty(n, np.uint8))
            cp = lat(with_copy)
            into = lat(lambda: ctx.process_into(img, reg.array))
            reg.free()
            own = np.empty(n, np.uint8)
            into_pageable = lat(lambda: ctx.process_into(img, own))
            print(f"{w}x{h}: process_mono8 median {m8[0]:.0f} us (p99 {m8[1]:.0f}), process_f32 median {f32[0]:.0f} us "
                  f"(p99 {f32[1]:.0f}); mono8 -> message buffer: library cloud + memcpy {cp[0]:.0f} us (p99 {cp[1]:.0f}), "
                  f"d2pc_process_mono8_into a registered buffer {into[0]:.0f} us (p99 {into[1]:.0f}), into a pageable "
                  f"buffer {into_pageable[0]:.0f} us (p99 {into_pageable[1]:.0f})", flush=True)


def direct(argv):
    """A/B of direct_out for the synchronous entries: the two modes alternate call by call in one process (box noise
    hits both alike); wall-clock median / p10 / p90 per mode, then the device-side spans of a timed call."""
    n = int(argv[0]) if argv else 1500
    with d2pc.Context() as ctx:
        for (w, h) in [(640, 480), (752, 480), (1280, 720)]:
            img = synth.s2_scene(h, w, 1)
            d = synth.s3_float(h, w, 1)
            for name, fn in (("mono8", lambda: ctx.process_mono8(img, copy=False)), ("f32", lambda: ctx.process_f32(d, copy=False))):
                modes = [-1, 0]  # direct_out
                ts = {m: [] for m in modes}
                for i in range(len(modes) * n + 60):
                    mode = modes[i % len(modes)]
                    ctx.set_tuning("direct_out", mode)
                    t0 = time.perf_counter()
                    fn()
                    dt = time.perf_counter() - t0
                    if i >= 60:
                        ts[mode].append(dt * 1e6)
                out = []
                for mode in modes:
                    v = sorted(ts[mode])
                    out.append(f"direct_out {mode:2d}: median {v[len(v)//2]:6.1f} us (p10 {v[len(v)//10]:6.1f}, p90 {v[9*len(v)//10]:6.1f})")
                print(f"{w}x{h} {name}: " + "; ".join(out), flush=True)
            ctx.set_timing(True)
            for mode in (-1, 0):
                ctx.set_tuning("direct_out", mode)
                sp = []
                for _ in range(50):
                    ctx.process_mono8(img, copy=False)
                    t = ctx.slot_timing(0)
                    sp.append((t.total_us, t.h2d_us, t.kernels_us, t.d2h_us))
                sp.sort()
                m = sp[len(sp) // 2]
                print(f"{w}x{h} mono8 spans, direct_out {mode:2d}: total {m[0]:.1f} us = H2D {m[1]:.1f} + kernels {m[2]:.1f} + D2H {m[3]:.1f}", flush=True)
            ctx.set_timing(False)
            ctx.set_tuning("direct_out", 0)
    # the synchronous config-5 call: four 1280x720 frames -> fused 665x665 map -> DisparityCb
    with d2pc.Context(offset_x=-7, offset_y=15) as ctx:
        four = [synth.s2_scene(720, 1280, 5 + i) for i in range(4)]
        ts = {-1: [], 0: []}
        for i in range(2 * 300 + 20):
            mode = -1 if i % 2 else 0
            ctx.set_tuning("direct_out", mode)
            cl = d2pc.Cloud()
            t0 = time.perf_counter()
            rc = d2pc.lib().d2pc_fuse_then_process(ctx._h, *(a.ctypes.data for a in four), 1280, 720, 1280, ctypes.byref(cl))
            dt = time.perf_counter() - t0
            assert rc == 0 and cl.width == 585 * 585
            if i >= 20:
                ts[mode].append(dt * 1e6)
        for mode in (-1, 0):
            v = sorted(ts[mode])
            print(f"fuse_then_process 4 x 1280x720, direct_out {mode:2d}: median {v[len(v)//2]:6.1f} us (p10 {v[len(v)//10]:6.1f}, p90 {v[9*len(v)//10]:6.1f})", flush=True)


def callbacks(argv):
    """Per-callback latency of the fusion node's synchronous entries (what include/d2pc_b200/nodes.hpp calls once per
    message): MatchingScoreCb1/2, publishFusedDepthMap, the colouriser of a debug topic, then DisparityCb on the fused
    map.  Pageable inputs (a ROS message's data vector), C ABI called directly, wall clock."""
    L = d2pc.lib()

    def lat(fn, n=300):
        for _ in range(10):
            fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            rc = fn()
            ts.append(time.perf_counter() - t0)
            assert rc == 0
        ts.sort()
        return ts[len(ts) // 2] * 1e6

    with d2pc.Context(offset_x=-7, offset_y=15) as ctx:
        for (w, h) in [(752, 480), (1280, 720)]:
            d1, d2, s1, s2 = (synth.s2_scene(h, w, 40 + i) for i in range(4))
            im1, im2, fu, co, col = d2pc.Image(), d2pc.Image(), d2pc.Image(), d2pc.Image(), d2pc.Image()
            t1 = lat(lambda: L.d2pc_preprocess_score(ctx._h, s1.ctypes.data, w, h, w, 1, ctypes.byref(im1)))
            t2 = lat(lambda: L.d2pc_preprocess_score(ctx._h, s2.ctypes.data, w, h, w, 2, ctypes.byref(im2)))
            p1, p2 = im1.array().copy(), im2.array().copy()
            tf = lat(lambda: L.d2pc_fuse_preprocessed(ctx._h, d1.ctypes.data, d2.ctypes.data, p1.ctypes.data, p2.ctypes.data,
                                                      w, h, w, ctypes.byref(fu), ctypes.byref(co)))
            fused = fu.array().copy()
            fh, fw = fused.shape
            tc = lat(lambda: L.d2pc_colorize_depth(ctx._h, fused.ctypes.data, fw, fh, fw, ctypes.byref(col)))
            cl = d2pc.Cloud()
            td = lat(lambda: L.d2pc_process_mono8(ctx._h, fused.ctypes.data, fw, fh, fw, ctypes.byref(cl)))
            print(f"{w}x{h}: MatchingScoreCb1 {t1:.0f} us, MatchingScoreCb2 {t2:.0f} us, DisparityCb2 + publishFusedDepthMap "
                  f"{tf:.0f} us, colorizeDepth({fw}x{fh}) {tc:.0f} us, DisparityCb on the fused map {td:.0f} us", flush=True)


def stream(argv):
    for (w, h, dt, nfr) in [(752, 480, np.uint8, 2000), (3840, 2160, np.float32, 128), (1280, 720, np.float32, 1000)]:
        pin = d2pc.PinnedArray((8, h, w), dt)
        for i in range(8):
            pin.array[i] = synth.s2_scene(h, w, 100 + i) if dt == np.uint8 else synth.s3_float(h, w, 100 + i)
        for slots in (2, 3, 4, 6):
            ctx = d2pc.Context(n_slots=slots)
            if argv:
                ctx.set_tuning("direct_out", int(argv[0]))  # 1: streamed frames also write their clouds directly
            ctx.process_stream(pin.array, collect=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.process_stream(pin.array, collect=False, n_frames=nfr)
            torch.cuda.synchronize()
            s = time.perf_counter() - t0
            n = (w - 80) * (h - 80)
            print(f"{w}x{h} {np.dtype(dt).name} slots={slots}: {nfr/s:9.1f} frames/s  {nfr*w*h/s/1e6:9.1f} Mpix/s  "
                  f"D2H {nfr*n*16/s/1e9:5.1f} GB/s  H2D {nfr*w*h*np.dtype(dt).itemsize/s/1e9:5.1f} GB/s", flush=True)
            ctx.close()
        pin.free()


def fusion(argv):
    """Config 5 stream: frame sets per second vs slots, the spans of one set, and the CPU time of a submission."""
    w, h, nsets = 1280, 720, 1500
    pin = d2pc.PinnedArray((4, 4, h, w), np.uint8)
    for i in range(4):
        for j in range(4):
            pin.array[i, j] = synth.s2_scene(h, w, 10 * i + j)
    for slots in [int(x) for x in argv] or (2, 3, 4, 6):
        ctx = d2pc.Context(n_slots=slots, offset_x=-7, offset_y=15)
        for pre in (False, True):   # without / with MatchingScoreCb1/2 (two 21 us kernels less per set)
            ctx.process_fusion_stream(pin.array, collect=False, preprocess_scores=pre)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.process_fusion_stream(pin.array, collect=False, n_sets=nsets, preprocess_scores=pre)
            torch.cuda.synchronize()
            s = time.perf_counter() - t0
            if not pre:
                print(f"   scores already preprocessed: {nsets/s:8.1f} sets/s")
        # one set alone with the spans, then the host cost of a submission with the slot already idle
        ctx.set_timing(True)
        ctx.submit_fusion(0, *pin.array[0])
        ctx.wait(0)
        t = ctx.slot_timing(0)
        ctx.set_timing(False)
        cpu = []
        for i in range(50):
            t1 = time.perf_counter()
            ctx.submit_fusion(0, *pin.array[i % 4])
            cpu.append(time.perf_counter() - t1)
            ctx.wait(0)
        # the same pipeline driven from here with timing on: spans of a set while its neighbours are in flight
        ctx.set_timing(True)
        spans = []
        sub = ret = 0
        t2 = time.perf_counter()
        while ret < 400:
            while sub < 400 and sub - ret < slots:
                ctx.submit_fusion(sub % slots, *pin.array[sub % 4])
                sub += 1
            cl = d2pc.Cloud()
            assert d2pc.lib().d2pc_wait(ctx._h, ret % slots, ctypes.byref(cl)) == 0   # no copy of the cloud
            ctx._keep.pop(ret % slots, None)
            tt = ctx.slot_timing(ret % slots)
            spans.append((tt.h2d_us, tt.kernels_us, tt.d2h_us, tt.total_us))
            ret += 1
        s2 = time.perf_counter() - t2
        ctx.set_timing(False)
        sp = np.median(np.array(spans[50:]), axis=0)
        print(f"   in the pipeline (python loop, {400/s2:.0f} sets/s): h2d {sp[0]:.0f} kernels {sp[1]:.0f} d2h {sp[2]:.0f} total {sp[3]:.0f} us")
        print(f"fusion {w}x{h} slots={slots}: {nsets/s:8.1f} sets/s ({1e6*s/nsets:6.1f} us/set); one set alone: h2d {t.h2d_us:.0f} "
              f"kernels {t.kernels_us:.0f} d2h {t.d2h_us:.0f} total {t.total_us:.0f} us; submit call (python) "
              f"{1e6*statistics.median(cpu):.0f} us", flush=True)
        ctx.close()
    pin.free()


if __name__ == "__main__":
    cmds = {"crop": crop, "compact": compact, "generic": generic, "numer": numer, "median": median, "score": score, "latency": latency, "direct": direct, "callbacks": callbacks,
            "stream": stream, "fusion": fusion}
    if len(sys.argv) < 2 or sys.argv[1] not in cmds:
        raise SystemExit(__doc__)
    cmds[sys.argv[1]](sys.argv[2:])
