"""ncu target: a few CROP_FINITE launches (classify-first and park kernels) on 1280x720x64."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
w, h, f = 1280, 720, 64
n = (w - 80) * (h - 80)
base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
ctx.set_filter_mode(1)
# optional: key=value tuning pairs on the command line (e.g. compact_variant=5 pipe_producers=4)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_tuning(k, int(v))
for park in (0, 0, 0):
    ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4, d_out.data_ptr(), n * 16, d_cnt.data_ptr())
ctx.sync()
print("kept", int(d_cnt.sum().item()))
