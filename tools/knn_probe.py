#!/usr/bin/env python3
"""remove_statistical_outlier(20, 2.0) and estimate_normals(Hybrid(0.02, 30)) on the benchmark's fused 5 mm cloud, timed with
CUDA events (and a target for ncu launch lists).
    python tools/knn_probe.py [--reps 5]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import repas_vision_b200 as rv  # noqa: E402
from bench import synth_chunk, H, W, FX, FY, CX, CY  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(7)
    d, c = synth_chunk(4, gen, dev)
    cam = rv.Camera(FX, FY, CX, CY, W, H)
    batch = rv.deproject_batch(d, c, cam, max_distance=2.5, dtype="f32")
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    down = rv.fuse_views([batch.frame(i) for i in range(4)], poses, 0.005)
    kept, _ = down.remove_statistical_outlier(20, 2.0)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    ms_sor = timed(lambda: down.remove_statistical_outlier(20, 2.0))
    ms_nrm = timed(lambda: kept.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30)))
    print(json.dumps({"points": len(down), "kept": len(kept), "sor_ms": ms_sor, "normals_hybrid_ms": ms_nrm}))


if __name__ == "__main__":
    main()
