#!/usr/bin/env python3
"""Generate tests/golden/reference_golden.json (+ frames/) by RUNNING THE REFERENCE'S OWN
FUNCTIONS in the build container, where /root/reference is mounted.

The reference is a folder of scripts; its pure numpy/cv2 functions import fine once the
camera / viewer modules that only `main()` touches are stubbed (open3d, pyrealsense2,
pyorbbecsdk, pupil_apriltags are not installable here).  Nothing from the reference is
copied: this script imports it, feeds it the captured frames that ship in its checkout
and seeded synthetic inputs, and records outputs (digests + samples).  The GPU box has
no /root/reference, so tests read only the committed JSON and the frame copies.

Run:  python tests/golden/make_golden.py        (from the repo root)
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import tempfile
import types

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = os.path.join(HERE, "frames")
CAL = os.path.join(HERE, "calibration")


# ----------------------------------------------------------------------------- stubs
class _FakeVec(list):
    pass


class _FakePointCloud:
    def __init__(self):
        self.points = None
        self.colors = None


def _install_stubs():
    o3d = types.ModuleType("open3d")
    o3d.geometry = types.SimpleNamespace(PointCloud=_FakePointCloud, TriangleMesh=object, LineSet=object)
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.array(a, dtype=np.float64, copy=True),
                                        Vector2iVector=lambda a: np.array(a))
    o3d.io = types.SimpleNamespace()
    o3d.visualization = types.SimpleNamespace()
    sys.modules["open3d"] = o3d
    sys.modules["pyrealsense2"] = types.ModuleType("pyrealsense2")
    pa = types.ModuleType("pupil_apriltags")
    pa.Detector = object
    sys.modules["pupil_apriltags"] = pa
    ob = types.ModuleType("pyorbbecsdk")
    for name in ("Pipeline", "Config", "OBSensorType", "OBFormat", "OBStreamType", "OBFrameAggregateOutputMode",
                 "AlignFilter", "PointCloudFilter", "OBError", "save_point_cloud_to_ply", "VideoStreamProfile"):
        setattr(ob, name, type(name, (), {}))
    ob.__getattr__ = lambda name: type(name, (), {})
    sys.modules["pyorbbecsdk"] = ob


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod  # dataclasses look their module up while the class body runs
    spec.loader.exec_module(mod)
    return mod


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(a: np.ndarray) -> dict:
    a = np.ascontiguousarray(a)
    return {"shape": list(a.shape), "dtype": str(a.dtype), "sha256": sha(a)}


# --------------------------------------------------------------------------- fixtures
CANOPY = ["2025-11-14T143013", "2025-11-14T143028", "2025-11-14T143037", "2025-11-14T143042", "2025-12-05T152733"]


def copy_fixtures():
    os.makedirs(FRAMES, exist_ok=True)
    os.makedirs(CAL, exist_ok=True)
    src = os.path.join(REF, "realsense_d415i/canopy_detection/new-captures")
    for ts in CANOPY:
        for stem in ("canopy_capture", "depth_snapshot"):
            shutil.copyfile(os.path.join(src, f"{stem}_{ts}_HD.png"), os.path.join(FRAMES, f"{stem}_{ts}_HD.png"))
    for rel in ("femto_bolt_code/scripts/calibration_parameters/factory_color_intrinsics_2025-09-08T143506.json",
                "femto_bolt_code/scripts/calibration_parameters/factory_depth_intrinsics_2025-09-08T143506.json",
                "femto_bolt_code/scripts/calibration_parameters/factory_extrinsics_d2c_2025-09-08T143506.json",
                "femto_bolt_code/scripts/calibration_parameters/checkerboard_color_intrinsics_2025-08-26T183535.json",
                "realsense_d415i/april_tag_detection_caliberation/factory_color_intrinsics_1280_720.json",
                "realsense_d415i/april_tag_detection_caliberation/factory_color_intrinsics_640_480.json",
                "realsense_d415i/april_tag_detection_caliberation/factory_d2c_extrinsics.json",
                "6dof/20250917_164430.txt"):
        shutil.copyfile(os.path.join(REF, rel), os.path.join(CAL, os.path.basename(rel)))
    for ts in CANOPY[:4]:
        shutil.copyfile(os.path.join(src, f"canopy_y_{ts}.txt"), os.path.join(FRAMES, f"canopy_y_{ts}.txt"))
    os.chmod(FRAMES, 0o755)
    for d in (FRAMES, CAL):
        for f in os.listdir(d):
            os.chmod(os.path.join(d, f), 0o644)


def blob_mask(h, w, seed):
    """Seeded segmentation-like mask: a few filled ellipses, ~15 % coverage."""
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.uint8)
    for _ in range(4):
        c = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        ax = (int(rng.integers(40, 220)), int(rng.integers(40, 160)))
        cv2.ellipse(m, c, ax, float(rng.uniform(0, 180)), 0, 360, 255, -1)
    return m


class _Intr:
    def __init__(self, d):
        self.fx, self.fy, self.ppx, self.ppy = d["fx"], d["fy"], d["ppx"], d["ppy"]
        self.width, self.height = d["width"], d["height"]


class _DepthFrame:
    def __init__(self, arr):
        self._a = arr

    def get_data(self):
        return self._a

    def get_width(self):
        return self._a.shape[1]

    def get_height(self):
        return self._a.shape[0]


def main():
    _install_stubs()
    copy_fixtures()
    cmp_ = _load(os.path.join(REF, "femto_bolt_code/scripts/create_masked_ply.py"), "ref_create_masked_ply")
    canopy = _load(os.path.join(REF, "realsense_d415i/canopy_detection/canopy_return.py"), "ref_canopy_return")
    fview = _load(os.path.join(REF, "femto_bolt_code/scripts/final_view.py"), "ref_final_view")
    btc = _load(os.path.join(REF, "femto_bolt_code/scripts/better_three_capture.py"), "ref_better_three_capture")
    G: dict = {"generator": "tests/golden/make_golden.py", "numpy": np.__version__, "cv2": cv2.__version__}

    rs_intr = json.load(open(os.path.join(CAL, "factory_color_intrinsics_1280_720.json")))
    fx, fy, cx, cy = rs_intr["fx"], rs_intr["fy"], rs_intr["ppx"], rs_intr["ppy"]

    # ---- 1. intrinsics I/O (create_masked_ply.py:27-52)
    G["intrinsics"] = {}
    from pathlib import Path
    for name in ("factory_color_intrinsics_2025-09-08T143506.json", "factory_depth_intrinsics_2025-09-08T143506.json",
                 "checkerboard_color_intrinsics_2025-08-26T183535.json"):
        vals = cmp_.load_color_intrinsics(Path(os.path.join(CAL, name)))
        G["intrinsics"][name] = {"load": list(vals),
                                 "scaled_640x360": list(cmp_.scale_intrinsics(*vals[:4], vals[4], vals[5], 640, 360)),
                                 "scaled_noop": list(cmp_.scale_intrinsics(*vals[:4], 0, 0, 640, 360))}

    # ---- 2. the canopy known-answer pipeline (canopy_return.py:319-409) on the 4 golden pairs
    G["canopy"] = []
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "src/camera_sensor/camera_z_data"))
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        for ts in CANOPY[:4]:
            color = cv2.imread(os.path.join(FRAMES, f"canopy_capture_{ts}_HD.png"), cv2.IMREAD_COLOR)
            depth = cv2.imread(os.path.join(FRAMES, f"depth_snapshot_{ts}_HD.png"), cv2.IMREAD_UNCHANGED)
            calls = []
            orig = canopy.deproject_pixel_to_point

            def spy(intr, pixel, depth_value, _o=orig, _c=calls):
                r = _o(intr, pixel, depth_value)
                _c.append((tuple(int(p) for p in pixel), float(depth_value), tuple(float(v) for v in r)))
                return r

            canopy.deproject_pixel_to_point = spy
            canopy.process_canopy_detection(color, _DepthFrame(depth), _Intr(rs_intr), ts)
            canopy.deproject_pixel_to_point = orig
            written = open("src/camera_sensor/camera_z_data/camera_z.txt").read()
            stored = open(os.path.join(FRAMES, f"canopy_y_{ts}.txt")).read().strip()
            (pix, dval, xyz), = calls
            G["canopy"].append({"ts": ts, "pixel": list(pix), "depth_m": dval, "xyz": list(xyz),
                                "written_now": written, "stored_in_reference": stored})
            assert written == stored, (ts, written, stored)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp)

    # ---- 3. get_depth_at_pixel / deproject_pixel_to_point on seeded pixels (incl. borders, holes)
    depth0 = cv2.imread(os.path.join(FRAMES, f"depth_snapshot_{CANOPY[0]}_HD.png"), cv2.IMREAD_UNCHANGED)
    rng = np.random.default_rng(7)
    pix = [(0, 0), (1279, 719), (1279, 0), (0, 719), (2, 1), (1278, 717)] + [
        (int(rng.integers(-5, 1285)), int(rng.integers(-5, 725))) for _ in range(58)]
    G["median_depth"] = []
    for (x, y) in pix:
        for win in (5, 11):
            d = canopy.get_depth_at_pixel(_DepthFrame(depth0), x, y, window_size=win)
            rec = {"x": x, "y": y, "window": win, "depth_m": None if d is None else float(d)}
            if d is not None:
                rec["xyz"] = [float(v) for v in canopy.deproject_pixel_to_point(_Intr(rs_intr), (x, y), d)]
            G["median_depth"].append(rec)
    G["median_depth_frame"] = f"depth_snapshot_{CANOPY[0]}_HD.png"

    # ---- 4. depth_to_meters (better_three_capture.py:118-125)
    raw, depth_m, scale = btc.depth_to_meters(_DepthFrame(depth0))
    G["depth_to_meters"] = {"frame": G["median_depth_frame"], "scale": scale, "depth_m": digest(depth_m)}
    allvals = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    _, all_m, _ = btc.depth_to_meters(_DepthFrame(allvals))
    G["depth_to_meters"]["all_u16"] = digest(all_m)

    # ---- 5. create_masked_pointcloud (create_masked_ply.py:56-107) on the five real frames
    G["masked_cloud"] = []
    for i, ts in enumerate(CANOPY):
        color = cv2.imread(os.path.join(FRAMES, f"canopy_capture_{ts}_HD.png"), cv2.IMREAD_COLOR)
        depth = cv2.imread(os.path.join(FRAMES, f"depth_snapshot_{ts}_HD.png"), cv2.IMREAD_UNCHANGED)
        _, dm, _ = btc.depth_to_meters(_DepthFrame(depth))
        h, w = depth.shape
        for variant in ("all", "blob", "blob_inverted"):
            mask = np.full((h, w), 255, np.uint8) if variant == "all" else blob_mask(h, w, 100 + i)
            pcd = cmp_.create_masked_pointcloud(color, dm, mask, fx, fy, cx, cy, invert_mask=(variant == "blob_inverted"))
            P, Cc = np.asarray(pcd.points), np.asarray(pcd.colors)
            rec = {"ts": ts, "variant": variant, "mask_seed": 100 + i, "n": int(P.shape[0]), "points": digest(P),
                   "colors": digest(Cc), "first": P[:3].tolist(), "last": P[-3:].tolist(),
                   "first_colors": Cc[:3].tolist()}
            if variant == "all":
                # distance_masking_on_ply.py:12-19 executed literally on the reference's points
                distances = np.linalg.norm(P, axis=1)
                keep = distances < 1.0
                rec["dist_lt_1m"] = {"kept": int(keep.sum()), "mask": digest(np.packbits(keep)),
                                     "points": digest(P[keep])}
                # view_point_cloud.py:109-116 with the example range 0.15 .. 8.0 m
                Z = P[:, 2]
                kz = np.ones(Z.shape[0], dtype=bool)
                kz &= (Z >= float(0.15))
                kz &= (Z <= float(8.0))
                rec["zclip_0p15_8"] = {"kept": int(kz.sum()), "mask": digest(np.packbits(kz))}
                # april_tag_bg_removal_pl.py:450-455 with a fixed box
                min_b, max_b = np.array([-0.3, -0.25, 0.5]), np.array([0.35, 0.2, 1.2])
                ka = ((P[:, 0] >= min_b[0]) & (P[:, 0] <= max_b[0]) & (P[:, 1] >= min_b[1]) & (P[:, 1] <= max_b[1])
                      & (P[:, 2] >= min_b[2]) & (P[:, 2] <= max_b[2]))
                rec["aabb"] = {"min": min_b.tolist(), "max": max_b.tolist(), "kept": int(ka.sum()),
                               "mask": digest(np.packbits(ka))}
            G["masked_cloud"].append(rec)

    # float depth with NaN / inf / negative entries (the isfinite & >0 predicate)
    rngf = np.random.default_rng(11)
    dm = (rngf.uniform(0.2, 4.0, (48, 64))).astype(np.float32)
    dm[rngf.random((48, 64)) < 0.1] = np.nan
    dm[rngf.random((48, 64)) < 0.05] = np.inf
    dm[rngf.random((48, 64)) < 0.05] = -1.0
    dm[rngf.random((48, 64)) < 0.1] = 0.0
    col = rngf.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    msk = (rngf.random((48, 64)) < 0.7).astype(np.uint8) * 255
    pcd = cmp_.create_masked_pointcloud(col, dm, msk, 60.0, 61.0, 31.5, 23.25)
    G["masked_cloud_float_small"] = {"seed": 11, "n": int(len(pcd.points)), "points": digest(np.asarray(pcd.points)),
                                     "colors": digest(np.asarray(pcd.colors))}

    # ---- 6. median_depth / color_pixel_to_3d (final_view.py:132-146)
    Zm = depth0.astype(np.float32) * 0.001
    G["final_view_median"] = []
    for (x, y) in pix[:24]:
        if 0 <= x < 1280 and 0 <= y < 720:
            for win in (5, 9):
                zc = fview.median_depth(Zm, x, y, win)
                p3 = fview.color_pixel_to_3d(float(x), float(y), zc, fx, fy, cx, cy)
                G["final_view_median"].append({"u": x, "v": y, "win": win, "z": float(zc), "p": p3.tolist()})

    # ---- 7. solve_pnp_with_best_obj_order (final_view.py:171-225) on exactly projected tag corners
    K = np.array([[748.8987426757812, 0, 639.8699951171875], [0, 748.3513793945312, 361.9516906738281], [0, 0, 1.0]])
    dist = np.zeros((5, 1))
    tag = 0.0303
    h2 = tag / 2
    # detector corner convention of the reference's winning order "BL,BR,TR,TL" 
    obj = np.array([[-h2, h2, 0], [h2, h2, 0], [h2, -h2, 0], [-h2, -h2, 0]], dtype=np.float64)
    G["solvepnp"] = []
    rngp = np.random.default_rng(3)
    for k in range(8):
        rvec = rngp.uniform(-0.5, 0.5, 3)
        rvec[0] += np.pi * (k % 2) * 0.1
        tvec = np.array([rngp.uniform(-0.2, 0.2), rngp.uniform(-0.15, 0.15), rngp.uniform(0.4, 1.2)])
        img, _ = cv2.projectPoints(obj, rvec, tvec, K, dist)
        img = img.reshape(4, 2).astype(np.float64)
        o, rv, tv, err, label = fview.solve_pnp_with_best_obj_order(img, K, dist, tag)
        R, _ = cv2.Rodrigues(rv)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = tv.reshape(3)
        G["solvepnp"].append({"corners_px": img.tolist(), "true_rvec": rvec.tolist(), "true_tvec": tvec.tolist(),
                              "rvec": rv.reshape(3).tolist(), "tvec": tv.reshape(3).tolist(), "err_px": float(err),
                              "label": label, "T_cam_tag": T.tolist()})
    G["solvepnp_K"] = K.tolist()
    G["solvepnp_tag_size"] = tag

    # ---- 8. the reference's own statement of a rigid transform of points: transform_point_tag_local_to_camera
    # (april_tag_bg_removal_pl.py:177-179, `R @ p + t`), what geom.transform(T) does to every point of a cloud
    # (final_view_with_cad.py:333).  Poses: the 4x4 the reference ships (6dof/20250917_164430.txt) and the solvePnP results above.
    bgr_mod = _load(os.path.join(REF, "femto_bolt_code/scripts/april_tag_bg_removal_pl.py"), "ref_april_tag_bg_removal_pl")
    rngt = np.random.default_rng(11)
    pts = np.concatenate([rngt.uniform(-0.5, 0.5, (96, 3)), rngt.uniform(-3.0, 3.0, (32, 3))])
    poses = [np.loadtxt(os.path.join(CAL, "20250917_164430.txt"))] + [np.array(g["T_cam_tag"]) for g in G["solvepnp"][:3]]
    G["rigid_transform"] = {"points": pts.tolist(), "cases": []}
    for T in poses:
        R, t = T[:3, :3].copy(), T[:3, 3].copy()
        out = np.stack([bgr_mod.transform_point_tag_local_to_camera(p, R, t) for p in pts])
        G["rigid_transform"]["cases"].append({"T": T.tolist(), "out": out.tolist()})

    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(G, f, indent=1)
    print("wrote", os.path.join(HERE, "reference_golden.json"))
    for c in G["canopy"]:
        print(c)


if __name__ == "__main__":
    main()
