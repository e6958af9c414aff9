#!/usr/bin/env python3
"""registration_icp (point-to-plane) on the benchmark's fused 5 mm cloud, timed with CUDA events (and a target for ncu launch
lists): half of the cloud, displaced, onto the whole cloud with its normals.
    python tools/icp_probe.py [--reps 5] [--host-loop]"""
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
    ap.add_argument("--host-loop", action="store_true")
    ap.add_argument("--steps-per-readback", type=int, default=0, help="override registration.ICP_STEPS_PER_READBACK")
    ap.add_argument("--dump-state", action="store_true", help="print the reserved words of the device loop's state (timing builds)")
    a = ap.parse_args()
    if a.steps_per_readback:
        rv.registration.ICP_STEPS_PER_READBACK = a.steps_per_readback
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
    kept.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30))
    D = rv.registration.vector6d_to_matrix4d([0.01, -0.008, 0.012, 0.003, -0.002, 0.004])
    src = kept.select_by_index(np.arange(0, len(kept), 2)).transform(D)
    crit = rv.ICPConvergenceCriteria(max_iteration=30)
    run = lambda: rv.registration_icp(src, kept, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), crit,  # noqa: E731
                                      device_loop=not a.host_loop)
    states = []
    if a.dump_state:
        from repas_vision_b200 import _ops
        make = _ops.icp_state
        _ops.icp_state = lambda dev: states.append(make(dev)) or states[-1]
    reg = run()
    torch.cuda.synchronize()
    if states:
        print("state words 40..47:", states[-1].cpu().numpy()[40:48].tolist())
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reg = run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(json.dumps({"source_points": len(src), "target_points": len(kept), "iterations": reg.iterations, "fitness": reg.fitness,
                      "ms": ms, "ms_per_iteration": ms / (reg.iterations + 1), "device_loop": not a.host_loop, "steps_per_readback": rv.registration.ICP_STEPS_PER_READBACK}))


if __name__ == "__main__":
    main()
