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
    ap.add_argument("--ring-slots", type=int, default=2, help="output ring depth (slots of `chunk` frames at dense capacity)")
    ap.add_argument("--e2e-frames", type=int, default=512, help="frames per end-to-end step (pinned host buffers)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    ap.add_argument("--mode", default="compact_ordered", choices=["compact_ordered", "compact_unordered", "dense_zero"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "tma"], help="K1 kernel (A/B runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rows", action="store_true", help="skip the rows of the other BASELINE configs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU; strong: the --frames batch of BASELINE configs[4] cut over the GPUs")
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * n_frames * reps / value, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, max(world, a.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def frames_per_gpu(a, world):
    """weak: every GPU takes a.frames; strong: the a.frames-frame batch of BASELINE configs[4] is cut into contiguous blocks."""
    if a.scaling == "strong":
        return max(1, a.frames // max(world, 1))
    return a.frames


def workload_config(a, world):
    fpg = frames_per_gpu(a, world)
    chunk = min(a.chunk, fpg)
    return {"workload": "BASELINE configs[4]: synthetic 1280x720 RGB-D (u16 depth + BGR8) -> validity + ||p||<1.0 m mask -> "
                        "ordered compacted float32 SoA xyz+rgb cloud",
            "frames_per_gpu_per_step": fpg, "global_batch": fpg * world, "chunk_frames": chunk, "output_ring_slots": a.ring_slots,
            "resolution": [W, H], "mode": a.mode, "kernel": a.kernel, "r_max_m": R_MAX, "unit_rule": "mul_f32",
            "e2e_workload": f"the same frames through HostPipeline.run: {min(a.e2e_frames, fpg)} frames per GPU per step from pinned "
                            "host memory, u16 depth + NV12 colour in (3.5 B/px, the camera's format), float32 xyz + r,g,b bytes "
                            "out (16 B/point), 32-frame chunks on three streams",
            "l2": f"inputs ({chunk * P * 5 / 1e9:.2f} GB per launch, {fpg * P * 5 / 1e9:.1f} GB per step) exceed the 126 MB L2; "
                  "no flush needed (the small rows flush it with a 512 MB write between repetitions)",
            "parallelism": f"frames sharded over {world} GPU(s), no collective on the hot path"}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
    return FALLBACK_HBM_GBS, "B200_PROFILING.md fallback"


def stored_traffic(key):
    """dram bytes per frame of a kernel from a stored `ncu --set full` capture (profiles/k1_traffic.json); None if absent."""
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        tj = json.load(open(tp))
        tj = tj.get(key, tj) if key else tj
        return float(tj["dram_bytes_per_launch"]) / float(tj["frames_per_launch"]), tj.get("source", "profiles/k1_traffic.json")
    except Exception:
        return None, None


def gather_floats(x, world, dev):
    """One float per rank -> list on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [float(x)]
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


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
    frames = frames_per_gpu(a, world)
    chunk = min(a.chunk, frames)
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
    ring = [torch.empty((6, chunk * P), dtype=torch.float32, device=dev) for _ in range(max(1, a.ring_slots))]
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
            r = rv.deproject_batch(depth[f0:f1], bgr[f0:f1], cam, out=ring[i % len(ring)], **kw)
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
    per_rank_value = gather_floats(frames * a.steps / (ms_local * 1e-3), world, dev)

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean CUDA-event duration of a launch
    full = [(e0.elapsed_time(e1), n) for e0, e1, n in events if n == chunk]
    k_ms = statistics.mean(t for t, _ in full)
    valid_frac = valid_per_step / (frames * P)
    written_frac = 1.0 if a.mode.startswith("dense") else valid_frac  # dense modes write every pixel
    alg_bytes = chunk * (P * 5 + written_frac * P * 24)
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    per_frame, tsrc = stored_traffic("headline")
    roofline = {"bound": "hbm", "kernel": "k_deproject_tma", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None if per_frame is None else per_frame * chunk,
                "traffic_source": None if per_frame is None else f"stored ncu --set full capture ({tsrc}), dram read + write per "
                                                                 "frame scaled to this run's frames per launch; not measured in this run",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k_ms,
                "frames_per_launch": chunk, "valid_fraction": valid_frac,
                "algorithmic_bytes_per_frame": "W*H*(2+3) + kept*24 (SURVEY 8d)"}

    # end to end through the host-buffer API: pinned host frames in (u16 depth + NV12 colour, what the camera delivers), host
    # clouds out, copies inside the timed region.  Next to it the same copies without the kernel (the host-side / PCIe ceiling of
    # this box at this N) and, on one GPU, the BGR-in / float-colour-out transport of round 1.
    e2e = None
    if not a.no_e2e:
        ef = min(a.e2e_frames, frames)
        hd = torch.empty((ef, H, W), dtype=torch.uint16).pin_memory()
        hn = torch.empty((ef, H * 3 // 2, W), dtype=torch.uint8).pin_memory()
        hd.copy_(depth[:ef])
        hn.copy_(torch.randint(0, 256, (ef, H * 3 // 2, W), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8))
        torch.cuda.synchronize()
        from repas_vision_b200.pipeline import HostPipeline

        def timed_e2e(fn, steps):
            barrier()
            ts = time.perf_counter()
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
            loc = (time.perf_counter() - ts) * 1e3
            return loc, shard.max_over_ranks(loc, dev)

        pipe = HostPipeline(cam, H, W, max_distance=R_MAX, chunk_frames=32, color_format="nv12", colors="u8", device=dev)
        state = {}

        def e2e_step():
            res = pipe.run(hd.numpy(), hn.numpy())
            state["points"] = int(res.counts.sum())  # the step's result is read on the host
            state["d2h"], state["h2d"] = res.d2h_bytes, res.h2d_bytes
            res.release()

        for _ in range(max(2, a.warmup)):
            e2e_step()
        e_loc, e_ms = timed_e2e(e2e_step, a.steps)
        e_frames = shard.sum_over_ranks(ef * a.steps, dev)
        like = pipe.run(hd.numpy(), hn.numpy())

        def probe_step():
            pipe.copy_probe(hd.numpy(), hn.numpy(), like=like).release()

        probe_step()
        p_loc, p_ms = timed_e2e(probe_step, max(3, a.steps // 2))
        p_frames = shard.sum_over_ranks(ef * max(3, a.steps // 2), dev)
        like.release()
        e2e = {"value": e_frames / (e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(state["h2d"]),
               "d2h_bytes_per_step": int(state["d2h"]), "frames_per_step_per_gpu": ef, "points_per_step_rank0": state["points"],
               "per_rank": gather_floats(ef * a.steps / (e_loc * 1e-3), world, dev),
               "copy_probe": {"value": p_frames / (p_ms * 1e-3), "unit": "frames/s",
                              "per_rank": gather_floats(ef * max(3, a.steps // 2) / (p_loc * 1e-3), world, dev),
                              "what": "the same host->device and device->host copies on the same streams with NO kernel: the "
                                      "ceiling the host side and PCIe of this box set at this number of GPUs"},
               "h2d_GBps_per_gpu": state["h2d"] * a.steps / (e_loc * 1e-3) / 1e9,
               "d2h_GBps_per_gpu": state["d2h"] * a.steps / (e_loc * 1e-3) / 1e9,
               "numa": numa, "cpu_affinity_cores": usable_cores(),
               "api": "repas_vision_b200.pipeline.HostPipeline(color_format='nv12', colors='u8').run(depth_u16[B,H,W], "
                      "nv12[B,H*3/2,W]) -> host float32 xyz + uint8 rgb + counts"}
        del pipe
        if world == 1:
            hc = torch.empty((ef, H, W, 3), dtype=torch.uint8).pin_memory()
            hc.copy_(bgr[:ef])
            pipe5 = HostPipeline(cam, H, W, max_distance=R_MAX, chunk_frames=32, colors="float", device=dev)

            def bgr_step():
                res = pipe5.run(hd.numpy(), hc.numpy())
                state["d2h5"], state["h2d5"] = res.d2h_bytes, res.h2d_bytes
                res.release()

            for _ in range(2):
                bgr_step()
            b_loc, _ = timed_e2e(bgr_step, max(3, a.steps // 2))
            e2e["bgr_in_float_colours_out"] = {"value": ef * max(3, a.steps // 2) / (b_loc * 1e-3), "unit": "frames/s",
                                               "h2d_bytes_per_step": int(state["h2d5"]), "d2h_bytes_per_step": int(state["d2h5"]),
                                               "what": "round 1's transport (5 B/px in, 24 B/point out) on the same frames"}
            del pipe5, hc
        del hd, hn

    cpu = None
    rows = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = usable_cores()
        n_frames = a.cpu_frames or max(cores * 4, 8)
        del depth, bgr, ring
        torch.cuda.empty_cache()
        rates, cores = cpu_reference_rates(n_frames, cores, 3, 1, 2)
        cpu = {"value": statistics.median(rates), "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n_frames} synthetic 720p frames (same generator) x 2 passes, median of 3, through oracle_np "
                         f"depth_to_meters -> create_masked_pointcloud -> ||p||<1.0 m, one worker process per core"}
    elif world > 1:
        del depth, bgr, ring
        torch.cuda.empty_cache()

    # the other BASELINE configs and rows of the hot path, each with its own roofline object (rank 0, one GPU); under torchrun
    # the fusion configuration's gather over NCCL instead (timed separately, excluded from frames/s)
    if world == 1 and not a.no_rows:
        try:
            rows = other_rows(rv, _ops, dev, gen, cpu_side=not a.no_cpu_baseline)
        except Exception as e:  # the headline line must still print; the failure is reported in it
            rows = [{"row": "other rows failed", "error": repr(e)}]
    elif world > 1 and not a.no_rows:
        try:
            rows = [fusion_gather_row(rv, _ops, shard, dev, gen, rank, world)]
        except Exception as e:
            rows = [{"row": "fusion gather failed", "error": repr(e)}]

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f32 storage, f64 geometry (north_star: float32 SoA xyz+rgb; x=(u-cx)*z/fx evaluated in float64 like the reference)",
            "data": "synthetic", "config": workload_config(a, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(gpu_launches), "clocks": clocks, "valid_points_per_step_rank0": valid_per_step,
            "per_rank_value": per_rank_value, "rows": rows,
        }))
    if world > 1:
        dist.destroy_process_group()


def fusion_gather_row(rv, _ops, shard, dev, gen, rank, world):
    """BASELINE configs[3] under N ranks: every rank fuses its own four-pose problem (K1 -> K3 -> K4), then the per-rank voxel
    counts are all-gathered and the fused clouds go to rank 0 over NCCL (counts first, then grouped send/recv).  The gather is
    timed on its own with CUDA events and is not part of frames/s."""
    import numpy as np
    import torch
    import torch.distributed as dist
    cam = rv.Camera(FX, FY, CX, CY, W, H)
    d4, c4 = synth_chunk(4, gen, dev)
    batch = rv.deproject_batch(d4, c4, cam, max_distance=2.5, dtype="f32")
    clouds = [batch.frame(i) for i in range(4)]
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    fused = rv.fuse_views(clouds, poses, 0.005)
    shard.gather_clouds(fused._data, len(fused))  # warm-up: NCCL channel set-up
    dist.barrier()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        counts = shard.gather_counts(torch.tensor([len(fused)], dtype=torch.int64, device=dev))
        merged, sizes = shard.gather_clouds(fused._data, len(fused))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = shard.max_over_ranks(statistics.median(ts), dev)
    ok = True
    if rank == 0:
        ok = merged.shape[1] == int(counts.sum().item()) == sum(sizes) and bool(torch.equal(merged[:, :sizes[0]], fused._data[:, :sizes[0]]))
    nbytes = int(sum(sizes)) * 24
    return {"row": "configs[3] gather: per-rank fused 5 mm clouds -> rank 0 over NCCL (counts all-gather, then grouped send/recv)",
            "ranks": world, "voxels_per_rank": [int(x) for x in sizes], "gather_ms": ms, "bytes_to_rank0": nbytes,
            "GBps_into_rank0": nbytes / (ms * 1e-3) / 1e9, "merged_equals_local_on_rank0": bool(ok),
            "note": "after the timed kernels; not part of frames/s (SURVEY 8e)"}


def other_rows(rv, _ops, dev, gen, cpu_side=True):
    """The other BASELINE configs and K1 workloads, one row each: GPU time by CUDA events (median, L2 flushed between
    repetitions for the small ones), algorithmic bytes, achieved GB/s against the measured HBM peak with the bound
    that actually limits the kernel named, and the CPU oracle on a bounded sample beside it.  Reported, not part of `value`."""
    import numpy as np
    import torch
    from oracle import oracle_c, oracle_np as O
    oracle_c.build()
    peak, _ = hbm_peak()
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=dev)

    def gpu_ms(fn, reps=7, do_flush=True):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if do_flush:
                flush.add_(1)  # 512 MB write: evicts the 126 MB L2 between repetitions
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    def roof(alg_bytes, ms, bound="hbm", note=None):
        ach = alg_bytes / (ms * 1e-3) / 1e9
        r = {"bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes": alg_bytes,
             "traffic": None}
        if note:
            r["note"] = note
        return r

    def cpu_time(fn, n):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n

    out = []
    cam = rv.Camera(FX, FY, CX, CY, W, H)

    # ---- K1 on 256 resident 720p frames: the headline workload with the camera's formats, then the two dense bounds
    B = 256
    d, c = synth_chunk(B, gen, dev)
    out6 = torch.empty((6, B * P), dtype=torch.float32, device=dev)
    nv12 = torch.randint(0, 256, (B, H * 3 // 2, W), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
    ms = gpu_ms(lambda: rv.deproject_batch(d, nv12, cam, max_distance=R_MAX, out=out6[:4], color_format="nv12",
                                           color_scale="packed8"), do_flush=False)
    kept = float(rv.deproject_batch(d, nv12, cam, max_distance=R_MAX, out=out6[:4], color_format="nv12",
                                    color_scale="packed8").counts.sum().item()) / (B * P)
    out.append({"row": "K1 720p, NV12 colour in, r,g,b bytes out (the transport of the e2e arm), 1 m mask, B=256", "frames_per_s": B / (ms * 1e-3),
                "ms": ms, "kept": kept, "roofline": roof(B * P * (3.5 + 16 * kept), ms, note="W*H*(2+1.5) + kept*16 bytes per frame")})
    del nv12
    di = d.to(torch.int32)
    full = torch.where(di == 0, torch.full_like(di, 900), di).to(torch.uint16)
    del di
    ms = gpu_ms(lambda: rv.deproject_batch(full, c, cam, out=out6), do_flush=False)
    out.append({"row": "K1 720p, every pixel valid (dense upper bound, SURVEY 8d adversarial set), ordered compact, B=256",
                "frames_per_s": B / (ms * 1e-3), "ms": ms, "kept": 1.0, "roofline": roof(B * P * 29, ms)})
    del full
    ms = gpu_ms(lambda: rv.deproject_batch(d, c, cam, out=out6), do_flush=False)
    kept = float(rv.deproject_batch(d, c, cam, out=out6).counts.sum().item()) / (B * P)
    out.append({"row": "K1 720p, validity mask only (the real-data 58 % validity of BASELINE.md), ordered compact, B=256",
                "frames_per_s": B / (ms * 1e-3), "ms": ms, "kept": kept, "roofline": roof(B * P * (5 + 24 * kept), ms)})

    # ---- BASELINE configs[1]: RealSense 640x480, /1000 rule, 1 m canopy mask, 1024-frame batch
    Bv = 1024
    dv = d[:, :480, :640].contiguous().repeat(4, 1, 1)
    cv = c[:, :480, :640].contiguous().repeat(4, 1, 1, 1)
    camv = rv.Camera(608.2335815429688, 607.8508911132812, 312.52239990234375, 232.65150451660156, 640, 480)
    outv = torch.empty((6, Bv * 480 * 640), dtype=torch.float32, device=dev)
    ms = gpu_ms(lambda: rv.deproject_batch(dv, cv, camv, max_distance=1.0, unit_rule="div_f32", out=outv), do_flush=False)
    kept = float(rv.deproject_batch(dv, cv, camv, max_distance=1.0, unit_rule="div_f32", out=outv).counts.sum().item()) / (Bv * 480 * 640)
    row = {"row": "configs[1]: RealSense 640x480 aligned deprojection + 1 m canopy mask, 1024-frame batch (f32(d)/1000 rule)",
           "frames_per_s": Bv / (ms * 1e-3), "ms": ms, "kept": kept, "roofline": roof(Bv * 480 * 640 * (5 + 24 * kept), ms)}
    if cpu_side:
        hd4, hc4 = dv[:4].cpu().numpy(), cv[:4].cpu().numpy()
        m480 = np.full((480, 640), 255, np.uint8)

        def cpu_vga():
            for i in range(4):
                dm = O.depth_to_meters(hd4[i], "div_f32")
                pts, cols = O.create_masked_pointcloud(hc4[i], dm, m480, camv.fx, camv.fy, camv.cx, camv.cy)
                k = O.distance_mask(pts, 1.0)
                pts[k], cols[k]
        row["cpu"] = {"value": 4.0 / cpu_time(cpu_vga, 2), "unit": "frames/s", "cores": 1, "what": "oracle_np, 4 frames x 2"}
    out.append(row)
    del dv, cv, outv

    # ---- BASELINE configs[2]: Femto ToF 640x480 registered into 1280x720 colour (a6), then deprojected
    depth = d[:, 120:600, 320:960].contiguous()
    dcam = rv.Camera(504.3227233886719, 504.2591247558594, 320.16888427734375, 345.57403564453125 - 48.0, 640, 480)
    ccam = rv.Camera(748.8987426757812, 748.3513793945312, 639.8699951171875, 361.9516906738281, 1280, 720)
    ang = np.deg2rad(6.0)
    R = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    t = np.array([0.032, -0.002, 0.004])
    ms = gpu_ms(lambda: rv.register_depth_to_color(depth, dcam, ccam, R, t), do_flush=False)
    reg_bytes = B * (640 * 480 * 2 + 1280 * 720 * 2)
    row = {"row": "configs[2] a6: registration 640x480 ToF depth -> 1280x720 colour grid (z-buffer, bit-exact winners), B=256",
           "frames_per_s": B / (ms * 1e-3), "ms": ms,
           "roofline": roof(reg_bytes, ms, bound="issue",
                            note="bound by instruction issue on the prescribed float32 geometry, not by HBM: k_reg_rects 82 % / k_reg_tile "
                                 "77 % issue-active in the stored ncu capture (profiles/r01_k2_register_ncu_raw_subset.csv); frac is the "
                                 "algorithmic-byte rate against the HBM peak")}
    if cpu_side:
        host = depth[:2].cpu().numpy()
        row["cpu"] = {"value": 1.0 / cpu_time(lambda: [oracle_c.register_depth_to_color(f, dcam.as_dict(), ccam.as_dict(), R.T.reshape(9), t)
                                                       for f in host], 1) * 2.0, "unit": "frames/s", "cores": 1,
                      "what": "oracle.c register_depth_to_color, 2 frames"}
    out.append(row)

    def reg_then_deproject():
        al, _ = _ops.register(depth, dcam, ccam, np.asarray(R).T.reshape(9), t)
        return rv.deproject_batch(al, c, ccam, max_distance=1.0, out=out6)

    ms = gpu_ms(reg_then_deproject, do_flush=False)
    kept = float(reg_then_deproject().counts.sum().item()) / (B * P)
    out.append({"row": "configs[2]: registration then deprojection + 1 m mask (two launches per batch), B=256", "frames_per_s": B / (ms * 1e-3),
                "ms": ms, "kept": kept, "roofline": roof(reg_bytes + B * P * (5 + 24 * kept), ms, bound="issue",
                                                         note="registration (issue-bound) is two thirds of the time")})
    del depth, out6

    # ---- BASELINE configs[3]: four poses -> tag frame -> merge -> 5 mm voxel grid
    batch = rv.deproject_batch(d[:4].contiguous(), c[:4].contiguous(), cam, max_distance=2.5, dtype="f32")
    clouds = [batch.frame(i) for i in range(4)]
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    n = sum(len(cl) for cl in clouds)
    down = rv.fuse_views(clouds, poses, 0.005)
    m = len(down)
    ms = gpu_ms(lambda: rv.fuse_views(clouds, poses, 0.005))
    row = {"row": "configs[3]: four-pose fusion: K3 transform into the tag frame + merge, K4 5 mm voxel grid (one count read back)",
           "points": n, "voxels": m, "ms": ms, "fusions_per_s": 1.0 / (ms * 1e-3), "points_per_s": n / (ms * 1e-3),
           "roofline": roof(n * 24 + m * 24, ms, note="N*24 read + M*24 written: the merged cloud is not counted (SURVEY 8d: 0 written if fused)")}
    Ts = [rv.world_from_camera(T) for T in poses]
    merged, total, bounds = _ops.transform_merge([(cl._data, cl._n) for cl in clouds], Ts, True, want_bounds=True)
    ms4 = gpu_ms(lambda: _ops.voxel_downsample(merged, total, True, 0.005, bounds=bounds))
    row["k4_alone"] = {"ms": ms4, "roofline": roof(n * 24 + m * 24, ms4)}
    if cpu_side:
        hp = merged[:3, :total].t().contiguous().cpu().numpy().astype(np.float64)
        hc = merged[3:, :total].t().contiguous().cpu().numpy().astype(np.float64)
        t_tr = cpu_time(lambda: [O.transform(cl.points, T) for cl, T in zip(clouds, Ts)], 1)
        t_vx = cpu_time(lambda: oracle_c.voxel_down_sample(hp, hc, 0.005), 1)
        row["cpu"] = {"value": 1.0 / (t_tr + t_vx), "unit": "fusions/s", "cores": 1, "ms": (t_tr + t_vx) * 1e3,
                      "what": "oracle_np.transform + oracle.c voxel_down_sample"}
        del hp, hc
    out.append(row)
    del merged

    # ---- 8f-4 on the fused cloud (create_masked_ply.py:168-174, mpa_icp_export.py:166-208): GPU on the whole cloud, the CPU side
    # (KD-tree formulations of the oracle, one core) on a contiguous slab of it, both as points per second
    ms_sor = gpu_ms(lambda: down.remove_statistical_outlier(20, 2.0), 3)
    kept_c, _ = down.remove_statistical_outlier(20, 2.0)
    ms_nrm = gpu_ms(lambda: kept_c.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30)), 3)
    D = rv.registration.vector6d_to_matrix4d([0.01, -0.008, 0.012, 0.003, -0.002, 0.004])
    src = kept_c.select_by_index(np.arange(0, len(kept_c), 2)).transform(D)
    crit = rv.ICPConvergenceCriteria(max_iteration=30)
    icp = lambda: rv.registration_icp(src, kept_c, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), crit)  # noqa: E731
    reg = icp()
    ms_icp = gpu_ms(icp, 3)
    lat = "search on a hash grid: bound by latency and the candidate distance tests, not by HBM"
    r_sor = {"row": "8f-4 remove_statistical_outlier(20, 2.0) on the fused 5 mm cloud", "points": len(down), "ms": ms_sor,
             "points_per_s": len(down) / (ms_sor * 1e-3), "roofline": roof(len(down) * 24 + len(kept_c) * 24, ms_sor, "latency", lat)}
    r_nrm = {"row": "8f-4 estimate_normals(Hybrid(0.02, 30))", "points": len(kept_c), "ms": ms_nrm,
             "points_per_s": len(kept_c) / (ms_nrm * 1e-3), "roofline": roof(len(kept_c) * 48, ms_nrm, "latency", lat)}
    r_icp = {"row": "8f-4 registration_icp point-to-plane (match + estimate + transform per iteration)", "source_points": len(src),
             "target_points": len(kept_c), "iterations": reg.iterations, "fitness": reg.fitness, "ms": ms_icp,
             "ms_per_iteration": ms_icp / (reg.iterations + 1), "matches_per_s": len(src) * (reg.iterations + 1) / (ms_icp * 1e-3),
             "roofline": roof((reg.iterations + 1) * len(src) * (24 + 4 + 48 + 48), ms_icp, "latency", lat)}
    if cpu_side:
        Pk = kept_c.points
        slab = Pk[np.argsort(Pk[:, 0], kind="stable")[:min(len(Pk), 100000)]]
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
        what = "oracle_np KD-tree formulation (scipy cKDTree), %d-point slab" % len(slab)
        r_sor["cpu"] = {"value": len(slab) / t_sor, "unit": "points/s", "cores": 1, "what": what}
        r_nrm["cpu"] = {"value": len(slab) / t_nrm, "unit": "points/s", "cores": 1, "what": what}
        r_icp["cpu"] = {"value": len(moved) / t_icp, "unit": "matches/s", "cores": 1, "what": what + " (one iteration, tree build included)"}
    out += [r_sor, r_nrm, r_icp]
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
