#!/usr/bin/env python
"""Targets for ncu / compute-sanitizer: each runs a few launches of one path and exits.

    python tools/ncu_target.py crop | compact | median | median1 | score | config5 | all [key=value tuning pairs]

crop      16 x 3840x2160 float frames through the headline CROP kernel
compact   64 x 1280x720 float frames, CROP_FINITE (band kernel)
median    256 x 752x480 mono8 callback (median 11 + reproject);  median1: ONE frame (latency case)
score     matching-score preprocessing of one 1280x720 frame, both callbacks
config5   one whole config-5 frame set through the pipelined host entry
all       every kernel once at small sizes (the compute-sanitizer target)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402


def tune(ctx, pairs):
    for kv in pairs:
        k, v = kv.split("=")
        ctx.set_tuning(k, int(v))


def float_run(w, h, f, compact, pairs):
    ctx = d2pc.Context()
    tune(ctx, pairs)
    n = (w - 80) * (h - 80)
    base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
    d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
    ctx.set_filter_mode(1 if compact else 0)
    for _ in range(3):
        ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16, d_cnt.data_ptr())
    ctx.sync()


def median_run(f, pairs):
    ctx = d2pc.Context()
    tune(ctx, pairs)
    w, h = 752, 480
    n = (w - 80) * (h - 80)
    d = torch.from_numpy(np.stack([synth.s2_scene(h, w, i) for i in range(8)])).cuda().repeat((f + 7) // 8, 1, 1)[:f].contiguous()
    o = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.reproject_mono8_device(d.data_ptr(), f, w, h, w, w * h, o.data_ptr(), n * 16)
    ctx.sync()


def score_run(pairs):
    w, h = 1280, 720
    ctx = d2pc.Context(offset_x=-7, offset_y=15)
    s = torch.from_numpy(synth.s2_scene(h, w, 5)).cuda()
    n = ctx.fuse_geometry(w, h)[4][0]
    out = torch.empty((n, n), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        for which in (1, 2):
            ctx.preprocess_score_device(s.data_ptr(), w, h, w, which, out.data_ptr())
    ctx.sync()


def config5_run(pairs):
    w, h = 1280, 720
    ctx = d2pc.Context(offset_x=-7, offset_y=15)
    sets = np.stack([np.stack([synth.s2_scene(h, w, 200 + 4 * i + k) for k in range(4)]) for i in range(2)])
    ctx.process_fusion_stream(sets, collect=False, n_sets=6)
    ctx.sync()


def all_run(pairs):
    with d2pc.Context(offset_x=-7, offset_y=15) as ctx:
        d = synth.s4_stress(131, 333, 1)
        d.reshape(-1)[::17] = 0.0
        d.reshape(-1)[5::97] = np.float32(1e-41)
        img = synth.s2_scene(150, 260, 2)
        ctx.process_f32(d)
        ctx.process_f32(d[:, 3:300])              # scalar-load path
        ctx.process_mono8(img)
        for variant in (0, 1):
            ctx.set_tuning("compact_variant", variant)
            ctx.set_filter_mode(1)
            ctx.process_f32(d)
            ctx.process_mono8(img)
            ctx.set_filter_mode(0)
        ctx.set_tuning("compact_variant", 0)
        ctx.set_arith_mode(1)
        ctx.process_f32(d)
        ctx.set_arith_mode(0)
        ctx.set_tuning("force_generic", 1)
        ctx.process_f32(d)
        ctx.set_tuning("force_generic", 0)
        ctx.set_tuning("median_variant", 2)
        ctx.process_mono8(img)
        ctx.set_tuning("median_variant", 0)
        four = [synth.s2_scene(200, 300, 3 + i) for i in range(4)]
        ctx.fuse(*four)
        p1 = ctx.preprocess_score(four[2], 1)
        p2 = ctx.preprocess_score(four[3], 2)
        ctx.fuse_preprocessed(four[0], four[1], p1, p2)
        ctx.fuse_then_process(*four)
        ctx.process_fusion_stream(np.stack([np.stack(four)] * 2))
        ctx.process_stream(np.stack([img, img, img, img]))
        ctx.colorize_depth(img)


if __name__ == "__main__":
    what, pairs = (sys.argv[1] if len(sys.argv) > 1 else ""), sys.argv[2:]
    if what == "crop":
        float_run(3840, 2160, 16, False, pairs)
    elif what == "compact":
        float_run(1280, 720, 64, True, pairs)
    elif what == "median":
        median_run(256, pairs)
    elif what == "median1":
        median_run(1, pairs)
    elif what == "score":
        score_run(pairs)
    elif what == "config5":
        config5_run(pairs)
    elif what == "all":
        all_run(pairs)
    else:
        raise SystemExit(__doc__)
    print("ok")
