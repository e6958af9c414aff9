#!/usr/bin/env python3
"""K1 on one resident batch in a chosen transport, timed with CUDA events (and a target for `ncu -k regex:k_deproject`).
    python tools/k1_probe.py --color nv12 --colors packed8 --frames 2048 --reps 5"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import repas_vision_b200 as rv  # noqa: E402
from bench import synth_chunk, H, W, FX, FY, CX, CY, P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--color", default="bgr", choices=["bgr", "nv12"])
    ap.add_argument("--colors", default="unit", choices=["unit", "packed8"])
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--r-max", type=float, default=1.0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(1234)
    B = a.frames
    depth = torch.empty((B, H, W), dtype=torch.uint16, device=dev)
    color = torch.empty((B, H, W, 3) if a.color == "bgr" else (B, H * 3 // 2, W), dtype=torch.uint8, device=dev)
    for f0 in range(0, B, 64):
        n = min(64, B - f0)
        d, c = synth_chunk(n, gen, dev)
        depth[f0:f0 + n] = d
        if a.color == "bgr":
            color[f0:f0 + n] = c
        else:
            color[f0:f0 + n] = torch.randint(0, 256, (n, H * 3 // 2, W), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
    planes = 4 if a.colors == "packed8" else 6
    out = torch.empty((planes, B * P), dtype=torch.float32, device=dev)
    cam = rv.Camera(FX, FY, CX, CY, W, H)
    kw = dict(max_distance=a.r_max if a.r_max > 0 else None, out=out, color_format=a.color, color_scale=a.colors)
    for _ in range(3):
        r = rv.deproject_batch(depth, color, cam, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = rv.deproject_batch(depth, color, cam, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    kept = float(r.counts.sum().item()) / (B * P)
    bpp_in = 5.0 if a.color == "bgr" else 3.5
    bpp_out = 24.0 if a.colors == "unit" else 16.0
    alg = B * P * (bpp_in + bpp_out * kept)
    print(json.dumps({"color": a.color, "colors": a.colors, "frames": B, "ms": ms, "frames_per_s": B / (ms * 1e-3), "kept": kept,
                      "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "library": rv.library_path()}))


if __name__ == "__main__":
    main()
