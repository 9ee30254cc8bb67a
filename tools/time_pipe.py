"""Times the CROP_FINITE pipeline kernel (variant 5) against the band kernel (variant 3) -- development tool."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
import oracle  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
stream = torch.cuda.ExternalStream(ctx.compute_stream())
PEAK = 6534.8


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


configs = [(3, 0, 0, 0, 0), (5, 1, 8, 4, 0), (5, 2, 8, 4, 0), (5, 4, 8, 4, 0), (5, 2, 8, 5, 0), (5, 2, 12, 4, 0), (5, 4, 12, 4, 0), (5, 4, 12, 5, 0), (5, 2, 12, 6, 0)]
if len(sys.argv) > 1:
    configs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]

# correctness first, on a small frame with zeros / NaN / inf / denormals sprinkled in
q = ctx.get_q()
d = synth.s3_float(480, 640, 0)
d[100, 50:90] = 0
d[240, :] = np.where(np.arange(640) % 3 == 0, 0, d[240, :])
d[300, 100] = np.nan
d[301, 101] = np.inf
d[302, 102] = 1e-40
d[303, 103] = 1e-38
want = oracle.filter_finite(oracle.disparity_cb_f32(d, q)).tobytes()
ctx.set_filter_mode(1)
for (var, pw, cw_, st, rows) in configs:
    if var == 3:
        ctx.set_tuning('prefetch_dist', pw)

    ctx.set_tuning("compact_variant", var)
    ctx.set_tuning("pipe_producers", pw)
    ctx.set_tuning("pipe_consumers", cw_)
    ctx.set_tuning("pipe_stages", st)
    ctx.set_tuning("rows_per_unit", rows)
    got = ctx.process_f32(d).tobytes()
    print("parity", (var, pw, cw_, st, rows), "OK" if got == want else "MISMATCH (%d vs %d bytes)" % (len(got), len(want)), flush=True)

for (w, h, f, kind) in [(1280, 720, 64, "s3"), (1280, 720, 64, "s2"), (3840, 2160, 16, "s3")]:
    n = (w - 80) * (h - 80)
    if kind == "s3":
        base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
    else:
        base = torch.from_numpy(synth.s2_scene(h, w, 3).astype(np.float32) * np.float32(0.125)).cuda()
    d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
    for (var, pw, cw_, st, rows) in configs:
        if var == 3:
            ctx.set_tuning('prefetch_dist', pw)
        ctx.set_tuning("compact_variant", var)
        ctx.set_tuning("pipe_producers", pw)
        ctx.set_tuning("pipe_consumers", cw_)
        ctx.set_tuning("pipe_stages", st)
        ctx.set_tuning("rows_per_unit", rows)
        s = t(lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16,
                                               d_cnt.data_ptr()))
        kept = int(d_cnt.sum().item())
        by = 4 * n * f + 16 * kept
        print(w, h, f, kind, "variant %d pw %d cw %d stages %d rows %d" % (var, pw, cw_, st, rows),
              "%.1f us  %.1f GB/s frac %.3f kept %.3f" % (s * 1e6, by / s / 1e9, by / s / 1e9 / PEAK, kept / (n * f)), flush=True)
    del d_in, d_out
ctx.set_tuning("compact_variant", 0)
ctx.set_tuning("rows_per_unit", 0)
ctx.set_filter_mode(0)
