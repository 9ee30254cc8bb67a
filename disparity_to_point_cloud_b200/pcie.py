"""Raw host <-> device copy ceiling of the end-to-end path (bench.py e2e.ceiling_gbs, tools/pcie_ceiling.py).

The end-to-end figure moves every frame into the GPU and every cloud out of it through pinned host memory; what
bounds it is the PCIe link of each GPU and, with several GPUs, whatever their links share on the host side.  This
measures that bound directly: plain cudaMemcpyAsync between pinned host buffers and device buffers of the same
sizes as the pipeline's, H2D and D2H on separate streams, no kernels, on every rank at the same time.
"""
from __future__ import annotations

import time


def _library_pinned(torch, nbytes, keep):
    """A pinned host tensor from the library's own allocator (d2pc_host_alloc: cudaHostAlloc, or the huge-page
    backed mapping when D2PC_PINNED_THP=1), i.e. the memory the product's slots use."""
    import numpy as np

    import disparity_to_point_cloud_b200 as d2pc
    pin = d2pc.PinnedArray((nbytes,), np.uint8)
    keep.append(pin)
    return torch.from_numpy(pin.array)


def measure(torch, world, h2d_bytes, d2h_bytes, seconds, barrier_sync, max_over_ranks, sum_over_ranks, n_buf=3,
            direction="both", library_alloc=True):
    """-> (GB/s summed over ranks and directions, units/s summed over ranks).  direction: both | h2d | d2h."""
    do_in = direction in ("both", "h2d") and h2d_bytes > 0
    do_out = direction in ("both", "d2h") and d2h_bytes > 0
    keep = []
    if library_alloc:
        host = lambda nb: _library_pinned(torch, nb, keep)  # noqa: E731
    else:
        host = lambda nb: torch.empty(nb, dtype=torch.uint8).pin_memory()  # noqa: E731
    h_in = [host(max(h2d_bytes, 1)) for _ in range(n_buf)] if do_in else []
    d_in = [torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device="cuda") for _ in range(n_buf)] if do_in else []
    h_out = [host(max(d2h_bytes, 1)) for _ in range(n_buf)] if do_out else []
    d_out = [torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device="cuda") for _ in range(n_buf)] if do_out else []
    for t in h_in + h_out:
        t.zero_()  # touch the pages on this rank's NUMA node
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def run(n):
        for i in range(n):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in[i % n_buf].copy_(h_in[i % n_buf], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out[i % n_buf].copy_(d_out[i % n_buf], non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    run(4)
    barrier_sync(world)
    t0 = time.perf_counter()
    run(8)
    pilot = max(time.perf_counter() - t0, 1e-6) / 8
    n = int(-max_over_ranks(-max(8, int(seconds / pilot)), world))  # the smallest count any rank proposes
    barrier_sync(world)
    t0 = time.perf_counter()
    run(n)
    barrier_sync(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    units = sum_over_ranks(n, world)
    per_unit = (h2d_bytes if do_in else 0) + (d2h_bytes if do_out else 0)
    del h_in, h_out
    for pin in keep:
        pin.free()
    return units * per_unit / dt / 1e9, units / dt
