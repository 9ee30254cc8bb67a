"""How should a cloud reach a PAGEABLE destination?  (a) D2H into pinned memory + memcpy, what d2pc_wait does for a
pageable caller buffer; (b) cudaMemcpy straight into the pageable buffer (driver-staged).  Not part of the library."""
import time

import numpy as np
import torch

for (w, h) in [(640, 480), (752, 480), (1280, 720)]:
    n = (w - 80) * (h - 80) * 16
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    pin = torch.empty(n, dtype=torch.uint8).pin_memory()
    page = torch.empty(n, dtype=torch.uint8)
    page_np = page.numpy()
    pin_np = pin.numpy()

    def a():
        pin.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        page_np[:] = pin_np

    def b():
        page.copy_(d)  # cudaMemcpy D2H into pageable memory
        torch.cuda.synchronize()

    for name, fn in (("pinned + memcpy", a), ("direct into pageable", b)):
        for _ in range(10):
            fn()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        print(f"{w}x{h} cloud {n/1e6:.1f} MB  {name:22s} median {ts[100]*1e6:7.1f} us", flush=True)
