"""ncu target: the mono8 callback on ONE 752x480 frame (single-stream latency case)."""
import sys

import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
w, h, f = 752, 480, 1
n = (w - 80) * (h - 80)
d = torch.from_numpy(synth.s2_scene(h, w, 0)).cuda()
o = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.reproject_mono8_device(d.data_ptr(), f, w, h, w, w * h, o.data_ptr(), n * 16)
ctx.sync()
print("ok")
