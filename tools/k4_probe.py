#!/usr/bin/env python3
"""K3/K4 on the four-view cloud of the benchmark, timed with CUDA events (and a target for per-kernel ncu passes).
    python tools/k4_probe.py [--voxel 0.005] [--reps 10]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import repas_vision_b200 as rv  # noqa: E402
from repas_vision_b200 import _ops  # noqa: E402
from bench import synth_chunk, H, W, FX, FY, CX, CY  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxel", type=float, default=0.005)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--flush", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(7)
    d, c = synth_chunk(4, gen, dev)
    cam = rv.Camera(FX, FY, CX, CY, W, H)
    batch = rv.deproject_batch(d, c, cam, max_distance=2.5, dtype="f32")
    clouds = [batch.frame(i) for i in range(4)]
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    Ts = [rv.world_from_camera(T) for T in poses]
    views = [(cl._data, cl._n) for cl in clouds]
    merged, total, bounds = _ops.transform_merge(views, Ts, True, want_bounds=True)
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=dev)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            if a.flush:
                flush.add_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    r = _ops.voxel_downsample(merged, total, True, a.voxel, bounds=bounds)
    m = int(r["m"].item())
    out = {"points": total, "voxels": m, "voxel": a.voxel}
    out["k4_ms"] = timed(lambda: _ops.voxel_downsample(merged, total, True, a.voxel, bounds=bounds))
    out["k4_nobounds_ms"] = timed(lambda: _ops.voxel_downsample(merged, total, True, a.voxel))
    out["fuse_ms"] = timed(lambda: _ops.fuse_voxel(views, Ts, True, a.voxel))
    out["k3_ms"] = timed(lambda: _ops.transform_merge(views, Ts, True, want_bounds=True))
    out["algorithmic_bytes"] = total * 24 + m * 24
    print(json.dumps(out))


if __name__ == "__main__":
    main()
