"""compute-sanitizer target: every kernel once, small sizes."""
import sys

import numpy as np

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

with d2pc.Context(offset_x=-7, offset_y=15) as ctx:
    d = synth.s4_stress(131, 333, 1)
    d.reshape(-1)[::17] = 0.0
    d.reshape(-1)[5::97] = np.float32(1e-41)
    img = synth.s2_scene(150, 260, 2)
    ctx.process_f32(d)
    ctx.process_f32(d[:, 3:300])              # scalar-load path
    ctx.process_mono8(img)
    for variant in (0, 1, 2, 3, 4, 5):
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
    ctx.set_tuning("median_variant", 1)
    ctx.process_mono8(img)
    ctx.set_tuning("median_variant", 0)
    four = [synth.s2_scene(200, 300, 3 + i) for i in range(4)]
    ctx.fuse(*four)
    p1 = ctx.preprocess_score(four[2], 1)
    p2 = ctx.preprocess_score(four[3], 2)
    ctx.fuse_preprocessed(four[0], four[1], p1, p2)
    ctx.fuse_then_process(*four)
    ctx.process_stream(np.stack([img, img, img, img]))
print("sanitize target done")
