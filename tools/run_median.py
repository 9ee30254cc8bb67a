"""ncu target: the mono8 callback (median 11 + reproject) on 752x480 x 256."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

ctx = d2pc.Context()
w, h, f = 752, 480, 256
n = (w - 80) * (h - 80)
d = torch.from_numpy(np.stack([synth.s2_scene(h, w, i) for i in range(8)])).cuda().repeat(f // 8, 1, 1).contiguous()
o = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_tuning(k, int(v))
for _ in range(3):
    ctx.reproject_mono8_device(d.data_ptr(), f, w, h, w, w * h, o.data_ptr(), n * 16)
ctx.sync()
print("ok")
