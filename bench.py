#!/usr/bin/env python
"""bench.py -- disparity -> PointCloud2 throughput (BASELINE.json's metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1 is launched by the driver under torch.distributed.run, one rank per GPU)

Workload (BASELINE.json configs[3], the one the 1/2/4/8-GPU metric is quoted on): 3840x2160 float32 disparity
frames (S3 distribution) through the reference's reprojection path (src/disparity_to_point_cloud.cpp:63-85):
reproject with Q -> 40-px crop -> pack {x,y,z,1.0f}.  A step is one pass over a batch of 1024 frames PER GPU
(frames are independent, so ranks shard them with no collective: weak scaling).

  value      Mpixel/s of input disparity, kernel-only: inputs resident in HBM in a ring of frame slots that is
             larger than L2 (so every launch streams from HBM), CUDA events on the launching stream.
  e2e        same metric through the C ABI's host entry (d2pc_process_stream): pinned host frames, H2D, kernel,
             D2H of every PointCloud2 payload, three streams / three slots overlapped.  Wall clock bracketed by
             device synchronisation (the work spans three streams), max over ranks.
  roofline   algorithmic bytes (20 B per kept point) / measured kernel time vs measured HBM peak.
  cpu_baseline  the CPU oracle port of the same path on a bounded sample, timed on this host (N=1, rank 0).

--impl reference times the CPU oracle port (the reference itself needs ROS/OpenCV/PCL and cannot be built in
this image) with all host threads on the same workload, a bounded sample of frames per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, BORDER = 3840, 2160, 40
FRAMES_PER_STEP = 1024
RING = 16                 # device-resident frame slots: 16 x (33.2 MB in + 125.1 MB out) = 2.5 GB >> 126 MB L2
N_PTS = (W - 2 * BORDER) * (H - 2 * BORDER)
ALG_BYTES_PER_FRAME = 20 * N_PTS   # 4 B disparity read + 16 B point written per kept pixel
METRIC = "Mpixels/s disparity->PointCloud2 (3840x2160 float32 disparity: reproject + 40px crop + XYZ1 pack)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and so its first-touch pinned buffers) to the NUMA node of its GPU: the
    end-to-end path is PCIe / host-DRAM bound, and a rank on the far socket halves it.  Best effort."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier_sync(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    from disparity_to_point_cloud_b200 import sharding
    return sharding.max_over_ranks(x, device="cuda") if world > 1 else x


def cpu_port_run(n_frames, threads, repeat=1):
    """Times the oracle port of cpp:63-85 on n_frames 4K S3 frames; returns (Mpix/s, seconds)."""
    import oracle
    from disparity_to_point_cloud_b200 import synth
    q = oracle.q_from_intrinsics()
    frames = np.empty((n_frames, H, W), dtype=np.float32)
    base = synth.s3_float(H, W, 0)
    for i in range(n_frames):
        frames[i] = np.roll(base, 97 * i, axis=1)
    cloud = np.empty((n_frames, N_PTS * 16), dtype=np.uint8)
    cloud[:] = 0  # touch the pages: the reference's publisher would reuse a warm allocator too
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        oracle.run_frames(frames, q, False, threads, cloud)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_frames * W * H / best / 1e6, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_frames = max(threads, 8) if threads <= 32 else threads
    times = []
    for i in range(args.warmup + args.steps):
        mpix, dt = cpu_port_run(n_frames, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n_frames * W * H / (ms / 1e3) / 1e6
    sample = f"{n_frames} frames of 3840x2160 f32 (S3) per step, frame-parallel on {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3840x2160 float32 disparity frames, reproject+crop+pack (BASELINE configs[3])",
                   "note": "reference node needs ROS1+OpenCV+PCL (unbuildable here): this is the CPU oracle port of "
                           "src/disparity_to_point_cloud.cpp:63-85, pinned bit-exact against cv2"},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch

    import disparity_to_point_cloud_b200 as d2pc
    from disparity_to_point_cloud_b200 import synth

    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    from disparity_to_point_cloud_b200 import sharding
    # frames are independent: weak = every rank runs args.frames frames per step, strong = args.frames are
    # sharded i mod G (BASELINE configs[3] wording); no data-path collective either way
    per_rank = sharding.frames_per_rank(args.frames, world, args.scaling)
    frames_per_step = per_rank[rank]
    total_frames_per_step = sum(per_rank)
    ring = min(RING, max(frames_per_step, 1))
    launches_per_step = (frames_per_step + ring - 1) // ring

    ctx = d2pc.Context(device=local, n_slots=args.slots)
    stream = torch.cuda.ExternalStream(ctx.compute_stream(), device=torch.device("cuda", local))

    # ---- resident inputs: RING distinct S3 frames (seeded per rank so ranks do not share data)
    base = torch.from_numpy(synth.s3_float(H, W, 1000 * rank)).cuda()
    d_in = torch.empty((ring, H, W), dtype=torch.float32, device="cuda")
    for i in range(ring):
        d_in[i] = torch.roll(base, shifts=131 * i + 7, dims=1)
    out_stride = N_PTS * 16
    d_out = torch.empty((ring, out_stride), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def kernel_step():
        left = frames_per_step
        while left > 0:
            n = min(ring, left)
            ctx.reproject_f32_device(d_in.data_ptr(), n, W, H, W * 4, W * H * 4, d_out.data_ptr(), out_stride)
            left -= n

    for _ in range(args.warmup):
        kernel_step()
    barrier_sync(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        kernel_step()
    ev1.record(stream)
    barrier_sync(world)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), world)
    ms_step = ms_total / args.steps
    value = total_frames_per_step * W * H / (ms_step / 1e3) / 1e6

    # one frame checked on the device path so a broken kernel can not post a number
    probe = d_out[1, :64].cpu().numpy().view(np.float32)
    assert probe[3] == 1.0 and probe[7] == 1.0, "kernel output is not XYZ1 points"

    peak, peak_src = measured_peak()
    per_launch_s = (ms_total / 1e3) / launches
    achieved = ALG_BYTES_PER_FRAME * ring / per_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "reproject_crop_kernel<float,vec16,exact-rectified>",
                "bytes_per_launch": ALG_BYTES_PER_FRAME * ring, "launch_us": per_launch_s * 1e6}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_ring16")
        except Exception:
            pass

    # ---- end to end: pinned host frames -> H2D -> kernel -> D2H, through d2pc_process_stream
    e2e_frames = args.e2e_frames if args.e2e_frames > 0 else frames_per_step
    e2e_total = sharding.sum_over_ranks(e2e_frames, device="cuda") if world > 1 else e2e_frames
    host_ring = 8
    pin = d2pc.PinnedArray((host_ring, H, W), np.float32)
    hb = synth.s3_float(H, W, 1000 * rank + 1)
    for i in range(host_ring):
        pin.array[i] = np.roll(hb, 61 * i, axis=1)
    checks = {}

    def sink(idx, cloud):
        if idx == 1:
            checks["first"] = cloud.bytes_view()[:32].copy()
            checks["width"] = cloud.width

    def e2e_step(n, check=False):
        # the timed call hands every cloud to a NULL sink inside the C ABI (no Python per frame);
        # the warm-up call looks at one cloud so a broken pipeline can not post a number
        ctx.process_stream(pin.array, collect=False, sink=sink if check else None, n_frames=n)

    e2e_step(host_ring, check=True)  # warm-up: allocates the slot buffers
    barrier_sync(world)
    t0 = time.perf_counter()
    e2e_step(e2e_frames)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    e2e_value = e2e_total * W * H / e2e_s / 1e6
    assert checks.get("width") == N_PTS
    e2e = {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": frames_per_step * W * H * 4,
           "d2h_bytes_per_step": frames_per_step * N_PTS * 16, "frames_timed": e2e_frames,
           "frames_per_s": e2e_total / e2e_s, "timer": "wall clock between device syncs, max over ranks"}
    pin.free()

    cpu_baseline = None
    extras = {}
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        n = 8 if threads <= 16 else min(threads, 64)
        mpix, dt = cpu_port_run(n, threads)
        mpix1, dt1 = cpu_port_run(2, 1)
        cpu_baseline = {"value": mpix, "unit": "Mpixel/s", "cores": threads, "kind": "port",
                        "sample": f"{n} frames 3840x2160 f32 (S3), frame-parallel on {threads} threads, {dt:.2f} s; "
                                  f"single thread (as the reference node runs): {mpix1:.1f} Mpixel/s"}
        if not args.no_extras:
            extras = run_extras(ctx, stream, torch, d2pc, synth)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3840x2160 float32 disparity (S3), %d frames per GPU per step, reproject+crop+pack "
                                   "(BASELINE configs[3])" % frames_per_step,
                       "frames_per_step_per_gpu": frames_per_step, "resident_ring_frames": ring,
                       "l2_policy": "inputs+outputs of one launch are 2.5 GB (>> 126 MB L2), no flush needed",
                       "arith": "EXACT (bit-identical to cv::reprojectImageTo3D)", "filter": "CROP (reference)",
                       "frames_per_s": total_frames_per_step / (ms_step / 1e3),
                       "sharding": "frame i -> rank i mod G, no collective on the data path"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _time_launches(torch, stream, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3


def run_extras(ctx, stream, torch, d2pc, synth):
    """Other BASELINE configs, kernel-only, short.  Reported for context; not the headline."""
    peak, _ = measured_peak()
    out = {}
    # config 3: 1280x720 x 64 resident (236 MB in + 786 MB out > L2)
    w, h, f = 1280, 720, 64
    n = (w - 80) * (h - 80)
    base = torch.from_numpy(synth.s3_float(h, w, 3)).cuda()
    d_in = torch.stack([torch.roll(base, 17 * i, dims=1) for i in range(f)]).contiguous()
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    s = _time_launches(torch, stream, lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4,
                                                                       d_out.data_ptr(), n * 16), 20)
    out["config3_1280x720x64_kernel"] = {"Mpixel/s": f * w * h / s / 1e6, "frames/s": f / s,
                                         "GB/s": 20 * n * f / s / 1e9, "frac_of_hbm_peak": 20 * n * f / s / 1e9 / peak}
    # same batch, CROP_FINITE (decoupled look-back compaction); S3 has ~0.4% zero disparities
    d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
    s = _time_launches(torch, stream, lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4,
                                                                       d_out.data_ptr(), n * 16, d_cnt.data_ptr()), 20)
    ctx.set_filter_mode(d2pc.FILTER_CROP)
    kept = int(d_cnt.sum().item())
    by = 4 * n * f + 16 * kept
    out["config3_crop_finite_kernel"] = {"Mpixel/s": f * w * h / s / 1e6, "GB/s": by / s / 1e9,
                                         "frac_of_hbm_peak": by / s / 1e9 / peak, "kept_fraction": kept / (n * f)}
    # the literal any-Q exact path on the same batch (what a non-rectified Q would take)
    ctx.set_tuning("force_generic", 1)
    s = _time_launches(torch, stream, lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4,
                                                                       d_out.data_ptr(), n * 16), 20)
    ctx.set_tuning("force_generic", 0)
    out["config3_generic_q_exact_kernel"] = {"Mpixel/s": f * w * h / s / 1e6, "frac_of_hbm_peak": 20 * n * f / s / 1e9 / peak}
    # FAST arithmetic on the same batch
    ctx.set_arith_mode(d2pc.ARITH_FAST)
    s = _time_launches(torch, stream, lambda: ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, w * h * 4,
                                                                       d_out.data_ptr(), n * 16), 20)
    ctx.set_arith_mode(d2pc.ARITH_EXACT)
    out["config3_fast_arith_kernel"] = {"Mpixel/s": f * w * h / s / 1e6, "frac_of_hbm_peak": 20 * n * f / s / 1e9 / peak}
    del d_in, d_out
    # mono8 entry (median 11 + reproject), 752x480 x 256 resident
    w, h, f = 752, 480, 256
    n = (w - 80) * (h - 80)
    d_img = torch.from_numpy(np.stack([synth.s2_scene(h, w, i) for i in range(8)])).cuda().repeat(f // 8, 1, 1)
    d_out = torch.empty((f, n * 16), dtype=torch.uint8, device="cuda")
    s = _time_launches(torch, stream, lambda: ctx.reproject_mono8_device(d_img.data_ptr(), f, w, h, w, w * h,
                                                                         d_out.data_ptr(), n * 16), 10)
    by = ((w - 70) * (h - 70) + 16 * n) * f
    out["config2_752x480_mono8x256_kernel"] = {"Mpixel/s": f * w * h / s / 1e6, "frames/s": f / s,
                                               "frac_of_hbm_peak": by / s / 1e9 / peak}
    # config 2 end to end: 752x480 mono8 stream, 1000 frames, pinned, 3 slots
    pin = d2pc.PinnedArray((8, h, w), np.uint8)
    for i in range(8):
        pin.array[i] = synth.s2_scene(h, w, 100 + i)
    ctx.process_stream(pin.array, collect=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.process_stream(pin.array, collect=False, n_frames=1000)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config2_752x480_mono8_stream_e2e"] = {"frames/s": 1000 / dt, "Mpixel/s": 1000 * w * h / dt / 1e6}
    pin.free()
    # config 5: four 1280x720 maps -> score preprocessing -> fuse -> median 3 -> trim -> DisparityCb (665x665)
    w, h = 1280, 720
    fctx = d2pc.Context(device=torch.cuda.current_device(), offset_x=-7, offset_y=15)
    fstream = torch.cuda.ExternalStream(fctx.compute_stream())
    four = [torch.from_numpy(synth.s2_scene(h, w, 200 + i)).cuda() for i in range(4)]
    st, r1, r2, rc, dims = fctx.fuse_geometry(w, h)
    n_sq, fw, fh = dims
    d_fused = torch.empty((fh, fw), dtype=torch.uint8, device="cuda")
    d_comb = torch.empty((n_sq, n_sq), dtype=torch.uint8, device="cuda")
    d_pre = [torch.empty((n_sq, n_sq), dtype=torch.uint8, device="cuda") for _ in range(2)]
    npts = (fw - 80) * (fh - 80)
    d_cloud = torch.empty(npts * 16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    s = _time_launches(torch, fstream, lambda: fctx.fuse_device(four[0].data_ptr(), four[1].data_ptr(),
                                                                four[2].data_ptr(), four[3].data_ptr(), w, h, w,
                                                                d_fused.data_ptr(), d_comb.data_ptr()), 50)
    out["config5_fuse_merge_median3_kernels"] = {"us": s * 1e6, "Mpixel/s (merged)": n_sq * n_sq / s / 1e6,
                                                 "GB/s (6 n^2)": 6 * n_sq * n_sq / s / 1e9}
    s = _time_launches(torch, fstream, lambda: [fctx.preprocess_score_device(four[2 + i].data_ptr(), w, h, w, i + 1,
                                                                             d_pre[i].data_ptr()) for i in range(2)], 50)
    out["config5_score_preprocess_x2_kernels"] = {"us": s * 1e6}
    s = _time_launches(torch, fstream, lambda: fctx.reproject_mono8_device(d_fused.data_ptr(), 1, fw, fh, fw, fw * fh,
                                                                           d_cloud.data_ptr(), npts * 16), 50)
    out["config5_fused_665x665_callback_kernels"] = {"us": s * 1e6, "points": npts}
    hfour = [synth.s2_scene(h, w, 200 + i) for i in range(4)]
    fctx.fuse_then_process(*hfour)
    t0 = time.perf_counter()
    for _ in range(50):
        fctx.fuse_then_process(*hfour)
    dt = (time.perf_counter() - t0) / 50
    out["config5_fuse_then_process_host_e2e"] = {"ms": dt * 1e3, "frame_sets/s": 1 / dt,
                                                 "note": "4 pageable 1280x720 frames in, 342225-point cloud out, synchronous"}
    fctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=0, help="frames timed end to end (0 = one full step)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--slots", type=int, default=3, help="pipeline slots per GPU of the end-to-end path")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU per step; strong: --frames in total, sharded i mod G")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
