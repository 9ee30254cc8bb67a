import sys
sys.path.insert(0, ".")
import numpy as np, torch
import disparity_to_point_cloud_b200 as d2pc
from disparity_to_point_cloud_b200 import synth
ctx = d2pc.Context()
stream = torch.cuda.ExternalStream(ctx.compute_stream())
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(it): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/it*1e3
for (w,h) in [(752,480),(665,665),(1280,720)]:
    d = torch.from_numpy(synth.s2_scene(h, w, 0)).cuda()
    n = (w-80)*(h-80)
    o = torch.empty((1, n*16), dtype=torch.uint8, device="cuda")
    for strip in (0, 2, 3, 4, 6, 8, 12, 16):
        ctx.set_tuning("median_strip", strip)
        print(w, h, "strip", strip, "%.1f us" % t(lambda: ctx.reproject_mono8_device(d.data_ptr(), 1, w, h, w, w*h, o.data_ptr(), n*16)), flush=True)
