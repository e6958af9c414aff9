#!/usr/bin/env python3
"""Per-kernel timings of the rest of the hot path: registration, transform + merge,
voxel grid, PLY records.  One JSON line per kernel with its algorithmic bytes and the fraction of the measured HBM peak.
bench.py stays the headline (K1); this tool produces the numbers quoted in profiles/ for K2-K4.

    python tools/bench_kernels.py [--reps 20]
"""
from __future__ import annotations

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

sys.path.insert(0, ROOT)
from bench import synth_chunk, H, W, FX, FY, CX, CY  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.add_(1)  # 512 MB write: evicts the 126 MB L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def line(name, ms, alg_bytes, units, unit_name, extra=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    d = {"kernel": name, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak(),
         unit_name + "_per_s": units / (ms * 1e-3)}
    if extra:
        d.update(extra)
    print(json.dumps(d))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(7)

    # ---- K2 registration: Femto ToF 640x480 (centre crop of the 640x576 intrinsics) -> 1280x720, 32 mm baseline, 6 deg pitch
    B = 256
    d720, _ = synth_chunk(B, gen, dev)
    depth = d720[:, 120:600, 320:960].contiguous()
    dcam = rv.Camera(504.3227233886719, 504.2591247558594, 320.16888427734375, 345.57403564453125 - 48.0, 640, 480)
    ccam = rv.Camera(748.8987426757812, 748.3513793945312, 639.8699951171875, 361.9516906738281, 1280, 720)
    ang = np.deg2rad(6.0)
    R = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    t = np.array([0.032, -0.002, 0.004])
    ms = timed(lambda: rv.register_depth_to_color(depth, dcam, ccam, R, t), a.reps, flush)
    line("K2 register 640x480 -> 1280x720 (B=256)", ms, B * (640 * 480 * 2 + 1280 * 720 * 2), B, "frames")

    # ---- BASELINE configs[2] as a pipeline: register 256 ToF frames into the colour grid, then deproject + 1 m mask
    _, bgr256 = synth_chunk(B, gen, dev)
    out6 = torch.empty((6, B * H * W), dtype=torch.float32, device=dev)

    def reg_then_deproject():
        al, _ = _ops.register(depth, dcam, ccam, np.asarray(R).T.reshape(9), t)
        return rv.deproject_batch(al, bgr256, ccam, max_distance=1.0, out=out6)

    ms = timed(reg_then_deproject, a.reps, flush)
    kept = float(reg_then_deproject().counts.sum().item()) / (B * H * W)
    line("K2 -> K1 pipeline: register 640x480 into 720p, deproject + 1 m mask (B=256)", ms,
         B * (640 * 480 * 2 + 1280 * 720 * 2) + B * H * W * (5 + 24 * kept), B, "frames", {"kept": kept})
    del out6, bgr256

    # ---- K1 on the registered frames feeds K3/K4: four views, distance-masked
    _, bgr = synth_chunk(4, gen, dev)
    cam = rv.Camera(FX, FY, CX, CY, W, H)
    batch = rv.deproject_batch(d720[:4].contiguous(), bgr, cam, max_distance=2.5, dtype="f32")
    clouds = [batch.frame(i) for i in range(4)]
    n = [len(c) for c in clouds]
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(rv.world_from_camera(T))
    views = [(c._data, c._n) for c in clouds]
    ms = timed(lambda: _ops.transform_merge(views, poses, True, want_bounds=True), a.reps, flush)
    N = sum(n)
    line("K3 transform+merge 4 views f32", ms, N * 24 * 2, N, "points", {"points": N})

    merged, total, bounds = _ops.transform_merge(views, poses, True, want_bounds=True)
    for vs in (0.005, 0.02):
        r = _ops.voxel_downsample(merged, total, True, vs, bounds=bounds)
        m = int(r["m"].item())
        ms = timed(lambda: _ops.voxel_downsample(merged, total, True, vs, bounds=bounds), a.reps, flush)
        line(f"K4 voxel_down_sample {vs * 1000:.0f} mm f32", ms, total * 24 + m * 24, total, "points", {"points": total, "voxels": m})

    # ---- a8 / a9 on an existing cloud (distance_masking_on_ply.py:12-19, view_point_cloud.py:109-116): ordered compaction of a
    # 32 M-point float32 cloud (768 MB in): ||p|| < 1 keeps about a third, the z-clip about half
    nb = 32 * 1024 * 1024
    big = (torch.rand((6, nb), generator=gen, device=dev, dtype=torch.float32) * 2.0 - 0.5)
    big[2] = torch.rand(nb, generator=gen, device=dev) * 1.2
    for name, kw in (("a8 distance mask ||p|| < 1 m", dict(r_max=1.0)), ("a9 z-clip 0.15 .. 0.8 m", dict(z_clip=(0.15, 0.8)))):
        o_, cnt_ = _ops.filter_cloud(big, nb, True, **kw)
        keptf = float(cnt_.item()) / nb
        del o_
        ms = timed(lambda: _ops.filter_cloud(big, nb, True, **kw), max(3, a.reps // 4), flush)
        line(f"{name} on a 32 M-point cloud (ordered compaction)", ms, nb * 24 * (1 + keptf), nb, "points", {"kept": keptf})
    del big

    # ---- BASELINE configs[3] as one call: transform 4 views into the tag frame, merge, 5 mm voxel grid (device-resident)
    cam_T = [np.linalg.inv(np.asarray(p).reshape(4, 4)) for p in poses]
    ms = timed(lambda: rv.fuse_views(clouds, cam_T, 0.005), a.reps, flush)
    line("fuse_views: 4 views -> tag frame -> merge -> 5 mm voxel grid (one int64 read back)", ms, N * 24 * 2 + N * 24, N, "points",
         {"points": N, "views_per_s": 4.0 / (ms * 1e-3)})

    # ---- create_masked_ply.py:163-170 on the fused cloud: 5 mm voxel grid, then remove_statistical_outlier(20, 2.0)
    down = rv.fuse_views(clouds, cam_T, 0.005)
    ms = timed(lambda: down.remove_statistical_outlier(20, 2.0), max(3, a.reps // 4), flush)
    kept, _ = down.remove_statistical_outlier(20, 2.0)
    line("remove_statistical_outlier(20, 2.0) on the 5 mm fused cloud (index list left on the device)", ms, len(down) * 24 + len(kept) * 24,
         len(down), "points", {"points": len(down), "kept": len(kept)})

    ms = timed(lambda: kept.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30)).orient_normals_towards_camera_location(),
               max(3, a.reps // 4), flush)
    line("estimate_normals(Hybrid(0.02, 30)) + orientation on that cloud", ms, len(kept) * 24 + len(kept) * 24, len(kept), "points",
         {"points": len(kept)})

    # ---- refine_with_icp (mpa_icp_export.py:166-208): point-to-plane ICP of a displaced half of that cloud onto it
    icp_T = np.eye(4)
    icp_T[:3, :3] = rv.registration.vector6d_to_matrix4d([0.01, -0.008, 0.012, 0, 0, 0])[:3, :3]
    icp_T[:3, 3] = (0.003, -0.002, 0.004)
    icp_src = kept.select_by_index(np.arange(0, len(kept), 2)).transform(icp_T)
    crit = rv.ICPConvergenceCriteria(max_iteration=30, relative_fitness=1e-6, relative_rmse=1e-6)
    run_icp = lambda: rv.registration_icp(icp_src, kept, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), crit)  # noqa: E731
    reg = run_icp()
    ms = timed(run_icp, max(3, a.reps // 4), flush)
    # the same loop on 2 k points: launch + read-back + host solve per iteration, the floor under any cloud size
    tiny_t = kept.select_by_index(np.arange(0, 4000, 2))
    tiny_t.estimate_normals(rv.KDTreeSearchParamHybrid(0.05, 30))
    tiny_s = tiny_t.select_by_index(np.arange(0, len(tiny_t), 2)).transform(icp_T)
    run_tiny = lambda: rv.registration_icp(tiny_s, tiny_t, 0.05, np.eye(4), rv.TransformationEstimationPointToPlane(), crit)  # noqa: E731
    tiny = run_tiny()
    ms_tiny = timed(run_tiny, max(3, a.reps // 4), flush)
    line("registration_icp point-to-plane, max_dist 0.02, half of that cloud onto it (host solve per iteration)", ms,
         (reg.iterations + 1) * (len(icp_src) * (24 + 4 + 48) + len(icp_src) * 48), len(icp_src) * (reg.iterations + 1), "matches",
         {"source_points": len(icp_src), "target_points": len(kept), "iterations": reg.iterations, "fitness": reg.fitness,
          "inlier_rmse": reg.inlier_rmse, "ms_per_iteration": ms / (reg.iterations + 1),
          "ms_per_iteration_on_2k_points": ms_tiny / (tiny.iterations + 1),
          "residual_rotation_translation": [float(np.abs((reg.transformation @ icp_T)[:3, :3] - np.eye(3)).max()),
                                            float(np.abs((reg.transformation @ icp_T)[:3, 3]).max())]})

    ms = timed(lambda: _ops.pack_ply_records(merged, total, True, "unit", "f32"), a.reps, flush)
    line("PLY records float xyz + uchar rgb", ms, total * 24 + total * 15, total, "points")

    # ---- K1 variants for reference: all-valid dense-capacity compact (upper bound) and VGA canopy config (BASELINE configs[1])
    B = 256
    d, c = synth_chunk(B, gen, dev)
    di = d.to(torch.int32)
    full = torch.where(di == 0, torch.full_like(di, 900), di).to(torch.uint16)
    out = torch.empty((6, B * H * W), dtype=torch.float32, device=dev)
    ms = timed(lambda: rv.deproject_batch(full, c, cam, out=out), a.reps, flush)
    line("K1 720p all pixels valid, ordered compact (B=256)", ms, B * H * W * 29, B, "frames")
    ms = timed(lambda: rv.deproject_batch(d, c, cam, out=out), a.reps, flush)
    kept = float(rv.deproject_batch(d, c, cam, out=out).counts.sum().item()) / (B * H * W)
    line("K1 720p validity mask only, ordered compact (B=256)", ms, B * H * W * (5 + 24 * kept), B, "frames", {"kept": kept})
    Bv = 1024
    dv = d[:, :480, :640].contiguous().repeat(4, 1, 1)
    cv = c[:, :480, :640].contiguous().repeat(4, 1, 1, 1)
    camv = rv.Camera(608.2335815429688, 607.8508911132812, 312.52239990234375, 232.65150451660156, 640, 480)
    outv = torch.empty((6, Bv * 480 * 640), dtype=torch.float32, device=dev)
    ms = timed(lambda: rv.deproject_batch(dv, cv, camv, max_distance=1.0, unit_rule="div_f32", out=outv), a.reps, flush)
    kept = float(rv.deproject_batch(dv, cv, camv, max_distance=1.0, unit_rule="div_f32", out=outv).counts.sum().item()) / (Bv * 480 * 640)
    line("K1 VGA canopy config (BASELINE configs[1], 1024 frames, /1000 rule, 1 m mask)", ms, Bv * 480 * 640 * (5 + 24 * kept), Bv,
         "frames", {"kept": kept})


    # ---- host-buffer pipeline with NV12 colour frames (3.5 instead of 5 bytes per pixel across PCIe)
    import time as _t
    from repas_vision_b200.pipeline import HostPipeline
    ef = 512
    hd = torch.empty((ef, H, W), dtype=torch.uint16).pin_memory()
    hd.copy_(d.repeat(2, 1, 1)[:ef])
    hn = torch.randint(0, 256, (ef, H * 3 // 2, W), dtype=torch.uint8).pin_memory()
    pipe = HostPipeline(cam, H, W, max_distance=1.0, chunk_frames=32, color_format="nv12", device=dev)
    res, runs = None, []
    for _ in range(9):  # the first runs page-lock the result blocks (two sets: a result is alive while the next is made)
        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        res = pipe.run(hd.numpy(), hn.numpy())
        torch.cuda.synchronize()
        runs.append(_t.perf_counter() - t0)
        res.release()
    dt = float(np.median(runs[3:]))
    print(json.dumps({"kernel": "e2e HostPipeline with NV12 colour frames (pinned host in, host clouds out), 512 x 720p",
                      "ms": dt * 1e3, "frames_per_s": ef / dt, "h2d_GBps": ef * H * W * 3.5 / dt / 1e9}))
    del pipe, res, hd, hn

    # ---- BASELINE configs[0]: ONE frame, numpy in -> numpy out through the scripts' call shape (host copies included)
    import time
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import synth_color, synth_depth
    rgb1 = synth_color(H, W, 3)
    dm1 = synth_depth(H, W, 3).astype(np.float32) * np.float32(0.001)
    msk1 = np.full((H, W), 255, np.uint8)

    def one_frame():
        pc = rv.create_masked_pointcloud(rgb1, dm1, msk1, FX, FY, CX, CY)
        return pc.points, pc.colors  # (N,3) float64 on the host, like np.asarray(pcd.points)

    for _ in range(3):
        one_frame()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        pts, _ = one_frame()
        ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    print(json.dumps({"kernel": "configs[0]: create_masked_pointcloud, one 720p frame, numpy float64 in/out (host wall clock)",
                      "ms": ms, "points": int(pts.shape[0]), "frames_per_s": 1e3 / ms}))


if __name__ == "__main__":
    main()
