"""ncu target: matching-score preprocessing of one 1280x720 frame, both callbacks."""
import sys

import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

w, h = 1280, 720
ctx = d2pc.Context(offset_x=-7, offset_y=15)
s = torch.from_numpy(synth.s2_scene(h, w, 5)).cuda()
st, r1, r2, rc, dims = ctx.fuse_geometry(w, h)
out = torch.empty((dims[0], dims[0]), dtype=torch.uint8, device="cuda")
for _ in range(2):
    for which in (1, 2):
        ctx.preprocess_score_device(s.data_ptr(), w, h, w, which, out.data_ptr())
ctx.sync()
print("ok")
