"""End-to-end stream throughput vs pipeline depth (n_slots) -- development tool."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import disparity_to_point_cloud_b200 as d2pc  # noqa: E402
from disparity_to_point_cloud_b200 import synth  # noqa: E402

for (w, h, dt, nfr) in [(752, 480, np.uint8, 2000), (3840, 2160, np.float32, 128), (1280, 720, np.float32, 1000)]:
    pin = d2pc.PinnedArray((8, h, w), dt)
    for i in range(8):
        pin.array[i] = synth.s2_scene(h, w, 100 + i) if dt == np.uint8 else synth.s3_float(h, w, 100 + i)
    for slots in (2, 3, 4, 6, 8):
        ctx = d2pc.Context(n_slots=slots)
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
