#!/usr/bin/env python3
"""Throughput of the RGB-D -> point-cloud hot path on B200: 720p frames/s and achieved HBM GB/s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the reference's numpy path on the host cores

A step is one pass of the fused deprojection + validity/distance mask + ordered compaction kernel
(rv_deproject_mask, BASELINE.json configs[4]) over a batch of synthetic 1280x720 RGB-D frames that is already
resident in HBM; the batch is walked in chunks whose outputs go to a two-deep ring (a whole batch of dense float32
clouds would not fit in 180 GB).  `value` = frames/s over all ranks; `e2e` = the same work through the public
host-buffer API (pinned host arrays in, host clouds out, copies inside the timed region); `roofline` = algorithmic
bytes of the kernel / its CUDA-event duration against MEASURED_PEAKS.json; `cpu_baseline` = the oracle port of the
reference's numpy statements timed on this box's cores (rank 0, N=1 only).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 720, 1280
P = H * W
# RealSense D415 factory colour intrinsics at 1280x720 (tests/golden/calibration/factory_color_intrinsics_1280_720.json)
FX, FY, CX, CY = 912.350341796875, 911.7763061523438, 628.7836303710938, 348.9772644042969
R_MAX = 1.0  # distance_masking_on_ply.py:15
METRIC = "rgbd_to_pointcloud_720p_frames_per_s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=8192, help="frames per GPU per step (BASELINE configs[4] batch)")
    ap.add_argument("--chunk", type=int, default=2048, help="frames per kernel launch (resident output ring slot)")
    ap.add_argument("--e2e-frames", type=int, default=512, help="frames per end-to-end step (pinned host buffers)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    ap.add_argument("--mode", default="compact_ordered", choices=["compact_ordered", "compact_unordered", "dense_zero"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "tma"], help="K1 kernel (A/B runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic frames (SURVEY.md 8d distributions)
def synth_chunk(n, gen, device):
    """n frames on `device`: piecewise-smooth planes 600-2200 mm with 2 mm noise, 5 % far background, ~42 % holes in
    horizontal runs, 0.03 % saturated pixels; colour = gradient blended with uniform noise.  Returns (u16 depth, u8 bgr)."""
    import torch
    f32 = torch.float32
    u = torch.arange(W, device=device, dtype=f32).view(1, 1, W) - W / 2
    v = torch.arange(H, device=device, dtype=f32).view(1, H, 1) - H / 2
    r = lambda *s: torch.rand(*s, generator=gen, device=device, dtype=f32)
    z0 = 600 + 1600 * r(n, 1, 1)
    a, b = 1.2 * r(n, 1, 1) - 0.6, 1.2 * r(n, 1, 1) - 0.6
    z = z0 + a * u + b * v
    z = z + (u > (0.6 * r(n, 1, 1) - 0.3) * W) * (900 * r(n, 1, 1) - 300)
    z = z + 2.0 * torch.randn(n, H, W, generator=gen, device=device, dtype=f32)
    far = r(n, H, W) < 0.05
    z = torch.where(far, 3000 + 7000 * r(n, H, W), z).clamp_(150, 65000)
    fld = torch.randn(n, H, W // 16 + 1, generator=gen, device=device, dtype=f32).repeat_interleave(16, dim=2)[:, :, :W]
    fld = fld + 0.35 * torch.randn(n, H, W, generator=gen, device=device, dtype=f32)
    z = torch.where(fld < -0.2139, torch.zeros_like(z), z)  # Phi^-1(0.42) * sqrt(1 + 0.35^2)
    z = torch.where(r(n, H, W) < 0.0003, torch.full_like(z, 65535.0), z)
    depth = z.to(torch.int32).to(torch.uint16)
    g = ((torch.arange(W, device=device).view(1, 1, W, 1) * 255 // (W - 1))
         + (torch.arange(H, device=device).view(1, H, 1, 1) * 255 // (H - 1)) * torch.tensor([1, 2, 3], device=device)) % 256
    noise = torch.randint(0, 256, (n, H, W, 3), generator=gen, device=device, dtype=torch.int32)
    bgr = ((g + noise) // 2).to(torch.uint8)
    return depth, bgr


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml, else nvidia-smi)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            try:
                return int(ids[self.index])
            except Exception:
                return self.index
        return self.index

    def _run(self):
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    n = self._nvml
                    self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                    try:
                        mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                else:
                    import subprocess
                    out = subprocess.run(["nvidia-smi", "-i", str(self._physical_index()),
                                          "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(float(out[0]))
                    self.max_mhz = float(out[1])
                    for k, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                        if "Active" in val and "Not" not in val:
                            self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(2.0)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------- CPU reference path
def _cpu_frame(args):
    """The reference's per-frame statements, restated by the oracle: depth_to_meters (better_three_capture.py:118-125)
    -> create_masked_pointcloud (create_masked_ply.py:74-100) -> ||p|| < 1.0 m (distance_masking_on_ply.py:12-19)."""
    depth, bgr, mask = args
    from oracle import oracle_np as O
    dm = O.depth_to_meters(depth, "mul_f32")
    pts, cols = O.create_masked_pointcloud(bgr, dm, mask, FX, FY, CX, CY)
    keep = O.distance_mask(pts, R_MAX)
    return int(pts[keep].shape[0]) + 0 * int(cols[keep].shape[0])


def _cpu_worker_main(conn):
    """Persistent worker: receives its frames once, then runs `reps` passes per "go" message."""
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    import numpy as np
    mask = np.full((H, W), 255, np.uint8)
    frames = conn.recv()
    _cpu_frame((frames[0][0], frames[0][1], mask))  # warm-up: imports, page faults
    conn.send("ready")
    while True:
        reps = conn.recv()
        if reps is None:
            return
        n = 0
        for _ in range(reps):
            for d, c in frames:
                _cpu_frame((d, c, mask))
                n += 1
        conn.send(n)


class CpuPool:
    """One spawned process per core, each holding a disjoint slice of the sample (frames are sent once, outside the
    timed region; a timed pass only exchanges a repetition count and a frame count)."""

    def __init__(self, depth, bgr, cores):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        n = depth.shape[0]
        self.cores = max(1, min(cores, n))
        self.procs, self.conns = [], []
        for w in range(self.cores):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker_main, args=(child,), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(parent)
        for w, c in enumerate(self.conns):
            c.send([(depth[i], bgr[i]) for i in range(w, n, self.cores)])
        for c in self.conns:
            assert c.recv() == "ready"

    def rate(self, reps=1) -> float:
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(reps)
        done = sum(c.recv() for c in self.conns)
        return done / (time.perf_counter() - t0)

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(5)


def host_frames(n_frames, seed=4321):
    """Synthetic frames for the CPU leg: same generator as the GPU workload when a GPU is present, numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            gen = torch.Generator(device="cuda").manual_seed(seed)
            out_d, out_c = [], []
            for f0 in range(0, n_frames, 32):
                d, c = synth_chunk(min(32, n_frames - f0), gen, torch.device("cuda", torch.cuda.current_device()))
                out_d.append(d.cpu())
                out_c.append(c.cpu())
            return torch.cat(out_d).numpy(), torch.cat(out_c).numpy()
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import synth_batch
    return synth_batch(n_frames, H, W, seed0=seed)


def cpu_reference_rates(n_frames, cores, steps, warmup, reps=1):
    depth, bgr = host_frames(n_frames)
    pool = CpuPool(depth, bgr, cores)
    try:
        rates = [pool.rate(reps) for _ in range(warmup + steps)][warmup:]
    finally:
        pool.close()
    return rates, pool.cores


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    reps = 4
    n_frames = a.cpu_frames or cores * 4
    rates, cores = cpu_reference_rates(n_frames, cores, a.steps, a.warmup, reps)
    value = statistics.median(rates)
    sample = (f"{n_frames} synthetic 720p frames x {reps} passes per step over {cores} worker processes "
              f"(disjoint frame slices, one process per core)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * n_frames * reps / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, 1),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(a, world):
    return {"workload": "BASELINE configs[4]: synthetic 1280x720 RGB-D (u16 depth + BGR8) -> validity + ||p||<1.0 m mask -> "
                        "ordered compacted float32 SoA xyz+rgb cloud",
            "frames_per_gpu_per_step": a.frames, "global_batch": a.frames * world, "chunk_frames": a.chunk,
            "resolution": [W, H], "mode": a.mode, "kernel": a.kernel, "r_max_m": R_MAX, "unit_rule": "mul_f32",
            "l2": f"inputs ({a.chunk * P * 5 / 1e9:.2f} GB per launch, {a.frames * P * 5 / 1e9:.1f} GB per step) exceed the 126 MB L2; no flush needed",
            "parallelism": f"frames sharded over {world} GPU(s), no collective on the hot path"}


# ------------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import repas_vision_b200 as rv
    from repas_vision_b200 import _ops, shard
    numa = shard.bind_to_gpu_numa_node(local) if world > 1 else {}  # pinned staging buffers local to each GPU's PCIe root

    cam = rv.Camera(FX, FY, CX, CY, W, H)
    frames, chunk = a.frames, min(a.chunk, a.frames)
    n_chunks = (frames + chunk - 1) // chunk
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    depth = torch.empty((frames, H, W), dtype=torch.uint16, device=dev)
    bgr = torch.empty((frames, H, W, 3), dtype=torch.uint8, device=dev)
    gstep = 64
    for f0 in range(0, frames, gstep):
        n = min(gstep, frames - f0)
        d, c = synth_chunk(n, gen, dev)
        depth[f0:f0 + n], bgr[f0:f0 + n] = d, c
    del d, c
    ring = [torch.empty((6, chunk * P), dtype=torch.float32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()

    kw = dict(max_distance=R_MAX, mode=a.mode, dtype="f32", kernel=a.kernel)
    launches_ctx = rv._lib.context(local)

    def one_step(events=None):
        kept = []
        for i in range(n_chunks):
            f0, f1 = i * chunk, min(frames, (i + 1) * chunk)
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            r = rv.deproject_batch(depth[f0:f1], bgr[f0:f1], cam, out=ring[i & 1], **kw)
            if events is not None:
                e1.record()
                events.append((e0, e1, f1 - f0))
            kept.append(r.counts)
        return kept

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        kept = one_step()
    barrier()
    valid_per_step = int(torch.cat(kept).sum().item())

    sampler = ClockSampler(local).start()
    events = []
    l0 = launches_ctx.launches
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        one_step(events)
    t1.record()
    barrier()
    gpu_launches = launches_ctx.launches - l0
    clocks = sampler.stop()
    ms_local = t0.elapsed_time(t1)
    ms = shard.max_over_ranks(ms_local, dev)
    total_frames = shard.sum_over_ranks(frames * a.steps, dev)
    value = total_frames / (ms * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean CUDA-event duration of a launch
    full = [(e0.elapsed_time(e1), n) for e0, e1, n in events if n == chunk]
    k_ms = statistics.mean(t for t, _ in full)
    valid_frac = valid_per_step / (frames * P)
    written_frac = 1.0 if a.mode.startswith("dense") else valid_frac  # dense modes write every pixel
    alg_bytes = chunk * (P * 5 + written_frac * P * 24)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "B200_PROFILING.md fallback"
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))  # ncu capture of one launch; scaled to this run's frames per launch
            traffic = float(tj["dram_bytes_per_launch"]) * chunk / float(tj["frames_per_launch"])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_deproject", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k_ms, "frames_per_launch": chunk,
                "valid_fraction": valid_frac}

    # end to end through the host-buffer API: pinned numpy in, host clouds out, copies inside the timed region
    e2e = None
    if not a.no_e2e:
        ef = min(a.e2e_frames, frames)
        hd = torch.empty((ef, H, W), dtype=torch.uint16).pin_memory()
        hc = torch.empty((ef, H, W, 3), dtype=torch.uint8).pin_memory()
        hd.copy_(depth[:ef])
        hc.copy_(bgr[:ef])
        torch.cuda.synchronize()
        from repas_vision_b200.pipeline import HostPipeline
        pipe = HostPipeline(cam, H, W, max_distance=R_MAX, chunk_frames=32, device=dev)
        for _ in range(max(1, a.warmup)):
            res = pipe.run(hd.numpy(), hc.numpy())
        barrier()
        ts = time.perf_counter()
        d2h = 0
        for _ in range(a.steps):
            res = pipe.run(hd.numpy(), hc.numpy())
            d2h = res.d2h_bytes
        torch.cuda.synchronize()
        e_ms_local = (time.perf_counter() - ts) * 1e3
        e_ms = shard.max_over_ranks(e_ms_local, dev)
        e_frames = shard.sum_over_ranks(ef * a.steps, dev)
        e2e = {"value": e_frames / (e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": ef * P * 5,
               "d2h_bytes_per_step": int(d2h), "frames_per_step_per_gpu": ef, "numa": numa,
               "api": "repas_vision_b200.pipeline.HostPipeline.run(depth_u16[B,H,W], bgr[B,H,W,3]) -> host float32 xyz+rgb + counts"}
        del hd, hc, pipe, res

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = usable_cores()
        n_frames = a.cpu_frames or max(cores * 4, 8)
        del depth, bgr, ring
        torch.cuda.empty_cache()
        rates, cores = cpu_reference_rates(n_frames, cores, 3, 1, 2)
        cpu = {"value": statistics.median(rates), "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n_frames} synthetic 720p frames (same generator) x 2 passes, median of 3, through oracle_np "
                         f"depth_to_meters -> create_masked_pointcloud -> ||p||<1.0 m, one worker process per core"}

    # the other rows of the hot path (a6 registration, a11 + a12 fusion) with the CPU oracle timed beside them on a
    # bounded sample; tools/bench_kernels.py has the full per-kernel table
    rows = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            rows = other_rows(rv, _ops, dev, gen)
        except Exception as e:  # the headline line must still print; the failure is reported in it
            rows = [{"row": "other rows failed", "error": repr(e)}]

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(a, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(gpu_launches), "clocks": clocks, "valid_points_per_step_rank0": valid_per_step, "rows": rows,
        }))
    if world > 1:
        dist.destroy_process_group()


def other_rows(rv, _ops, dev, gen):
    """Registration (BASELINE configs[2]) and four-pose fusion (configs[3]): GPU time by CUDA events next to the C oracle on
    one host core (registration: 2 frames; voxel grid: the same merged cloud).  Reported, not part of `value`."""
    import numpy as np
    import torch
    from oracle import oracle_c, oracle_np as O
    oracle_c.build()

    def gpu_ms(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    out = []
    B = 64
    d720, bgr = synth_chunk(B, gen, dev)
    depth = d720[:, 120:600, 320:960].contiguous()
    dcam = rv.Camera(504.3227233886719, 504.2591247558594, 320.16888427734375, 345.57403564453125 - 48.0, 640, 480)
    ccam = rv.Camera(748.8987426757812, 748.3513793945312, 639.8699951171875, 361.9516906738281, 1280, 720)
    ang = np.deg2rad(6.0)
    R = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    t = np.array([0.032, -0.002, 0.004])
    ms = gpu_ms(lambda: rv.register_depth_to_color(depth, dcam, ccam, R, t))
    host = depth[:2].cpu().numpy()
    t0 = time.perf_counter()
    for f in host:
        oracle_c.register_depth_to_color(f, dcam.as_dict(), ccam.as_dict(), R.T.reshape(9), t)
    cpu_s = (time.perf_counter() - t0) / len(host)
    out.append({"row": "a6 registration 640x480 -> 1280x720", "gpu_frames_per_s": B / (ms * 1e-3), "batch": B,
                "cpu_frames_per_s": 1.0 / cpu_s, "cpu": "oracle.c register_depth_to_color, 1 core, 2 frames"})

    cam = rv.Camera(FX, FY, CX, CY, W, H)
    batch = rv.deproject_batch(d720[:4].contiguous(), bgr[:4].contiguous(), cam, max_distance=2.5, dtype="f32")
    clouds = [batch.frame(i) for i in range(4)]
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    ms = gpu_ms(lambda: rv.fuse_views(clouds, poses, 0.005))
    n = sum(len(c) for c in clouds)
    merged, total, _ = _ops.transform_merge([(c._data, c._n) for c in clouds], [rv.world_from_camera(T) for T in poses], True)
    hp = merged[:3, :total].t().contiguous().cpu().numpy().astype(np.float64)
    hc = merged[3:, :total].t().contiguous().cpu().numpy().astype(np.float64)
    t0 = time.perf_counter()
    parts = [O.transform(c.points, rv.world_from_camera(T)) for c, T in zip(clouds, poses)]
    t_tr = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle_c.voxel_down_sample(hp, hc, 0.005)
    t_vx = time.perf_counter() - t0
    del parts
    out.append({"row": "a11 + a12 four-pose fusion: transform, merge, 5 mm voxel grid", "points": n, "gpu_ms": ms,
                "cpu_ms": (t_tr + t_vx) * 1e3, "cpu": "oracle_np.transform + oracle.c voxel_down_sample, 1 core"})

    # 8f-4 on the fused cloud (create_masked_ply.py:168-174, mpa_icp_export.py:166-208): GPU on the whole cloud, the CPU side
    # (KD-tree formulations of the oracle, one core) on a contiguous slab of it, both as points per second
    down = rv.fuse_views(clouds, poses, 0.005)
    ms_sor = gpu_ms(lambda: down.remove_statistical_outlier(20, 2.0), 3)
    kept, _ = down.remove_statistical_outlier(20, 2.0)
    ms_nrm = gpu_ms(lambda: kept.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30)), 3)
    D = rv.registration.vector6d_to_matrix4d([0.01, -0.008, 0.012, 0.003, -0.002, 0.004])
    src = kept.select_by_index(np.arange(0, len(kept), 2)).transform(D)
    crit = rv.ICPConvergenceCriteria(max_iteration=30)
    icp = lambda: rv.registration_icp(src, kept, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), crit)  # noqa: E731
    reg = icp()
    ms_icp = gpu_ms(icp, 3)
    P = kept.points
    slab = P[np.argsort(P[:, 0], kind="stable")[:min(len(P), 100000)]]
    t0 = time.perf_counter()
    avg = O.knn_mean_distance_kdtree(slab, 20)
    O.statistical_outlier_indices(avg, 2.0)
    t_sor = time.perf_counter() - t0
    t0 = time.perf_counter()
    Ns = O.estimate_normals_kdtree(slab, 0.02, 30)
    t_nrm = time.perf_counter() - t0
    moved = O.transform(slab[::2], D)
    t0 = time.perf_counter()
    near, _, _ = O.nearest_correspondences_kdtree(moved, slab, 0.02)
    O.point_to_plane_update(moved, slab, Ns, near)
    t_icp = time.perf_counter() - t0
    cpu = "oracle_np KD-tree formulation (scipy cKDTree), 1 core, %d-point slab" % len(slab)
    out.append({"row": "8f-4 remove_statistical_outlier(20, 2.0)", "points": len(down), "gpu_ms": ms_sor,
                "gpu_points_per_s": len(down) / (ms_sor * 1e-3), "cpu_points_per_s": len(slab) / t_sor, "cpu": cpu})
    out.append({"row": "8f-4 estimate_normals(Hybrid(0.02, 30))", "points": len(kept), "gpu_ms": ms_nrm,
                "gpu_points_per_s": len(kept) / (ms_nrm * 1e-3), "cpu_points_per_s": len(slab) / t_nrm, "cpu": cpu})
    out.append({"row": "8f-4 registration_icp point-to-plane, per iteration (match + estimate + transform)",
                "source_points": len(src), "target_points": len(kept), "iterations": reg.iterations, "fitness": reg.fitness,
                "gpu_ms": ms_icp, "gpu_matches_per_s": len(src) * (reg.iterations + 1) / (ms_icp * 1e-3),
                "cpu_matches_per_s": len(moved) / t_icp, "cpu": cpu + " (one iteration, tree build included)"})
    return out


def main():
    a = parse_args()
    # stdout carries the one JSON line and nothing else: libraries that write to file descriptor 1 on their own (NCCL prints
    # its version line there when NCCL_DEBUG is set) are pointed at stderr, and the line goes out through the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_b200(a)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    main()
