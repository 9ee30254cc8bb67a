"""ncu target: one config-5 frame set (score preprocessing x2, fuse, median 3, DisparityCb on the fused map)."""
import sys

import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

w, h = 1280, 720
ctx = d2pc.Context(offset_x=-7, offset_y=15)
four = [synth.s2_scene(h, w, 200 + i) for i in range(4)]
for _ in range(3):
    ctx.fuse_then_process(*four)
ctx.sync()
print("ok")
