"""Calibration / pose file loaders with the reference scripts' call shapes.

Every function here mirrors one the reference defines (paths relative to the reference
checkout); names, argument order, return tuples and the exceptions raised are the same so a
script can switch its import and keep working:

* load_color_intrinsics / scale_intrinsics ... femto_bolt_code/scripts/create_masked_ply.py:27-52
  (identical copies in final_view.py:32-57, april_tag_bg_removal_pl.py:55-69)
* load_intrinsics_json ..................... femto_bolt_code/scripts/april_tag_detector_solvepnp.py:51-66
* load_intrinsics (RealSense ppx/ppy/coeffs) . realsense_d415i/vis_tool/vis_tool_april_tag_pose_validaiton.py:38-47
* read_depth_to_color_extrinsics ........... realsense_d415i/vis_tool/vis_tool_april_tag_pose_validaiton.py:93-98
* load_transform_matrix .................... femto_bolt_code/scripts/6dof_icp_export.py:55-70

`load_camera` / `load_extrinsics` are the one addition: they fold the three on-disk
dialects into the `Camera` record the CUDA entry points take.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_RS_MODELS = {
    "distortion.none": "none", "none": "none",
    "distortion.brown_conrady": "brown_conrady", "brown_conrady": "brown_conrady",
    "distortion.inverse_brown_conrady": "inverse_brown_conrady", "inverse_brown_conrady": "inverse_brown_conrady",
    "distortion.modified_brown_conrady": "modified_brown_conrady", "modified_brown_conrady": "modified_brown_conrady",
}


@dataclass
class Camera:
    """Pinhole intrinsics + optional Brown-Conrady coefficients (k1,k2,p1,p2,k3)."""
    fx: float
    fy: float
    cx: float
    cy: float
    width: int = 0
    height: int = 0
    dist: tuple = field(default_factory=lambda: (0.0, 0.0, 0.0, 0.0, 0.0))
    model: str = "none"

    @property
    def K(self) -> np.ndarray:
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]], dtype=np.float64)

    @property
    def distorted(self) -> bool:
        return self.model != "none" and any(float(c) != 0.0 for c in self.dist)

    def scaled(self, dst_w: int, dst_h: int) -> "Camera":
        fx, fy, cx, cy = scale_intrinsics(self.fx, self.fy, self.cx, self.cy, self.width, self.height, dst_w, dst_h)
        return Camera(fx, fy, cx, cy, int(dst_w), int(dst_h), tuple(self.dist), self.model)

    def as_dict(self) -> dict:
        return dict(fx=self.fx, fy=self.fy, cx=self.cx, cy=self.cy, width=self.width, height=self.height,
                    dist=list(self.dist), model=self.model)


def load_color_intrinsics(json_path):
    """(fx, fy, cx, cy, width, height) from a flat JSON or one nested under "color_intrinsics".

    FileNotFoundError when the file is missing, KeyError when one of fx/fy/cx/cy is absent;
    width/height default to 0 (create_masked_ply.py:27-43)."""
    json_path = Path(json_path)
    if not json_path.exists():
        raise FileNotFoundError(f"Intrinsics JSON not found: {json_path}")
    data = json.loads(json_path.read_text())
    intr = data["color_intrinsics"] if (isinstance(data, dict) and "color_intrinsics" in data) else data
    for k in ("fx", "fy", "cx", "cy"):
        if k not in intr:
            raise KeyError(f"Missing '{k}' in intrinsics JSON: {json_path}")
    return (float(intr["fx"]), float(intr["fy"]), float(intr["cx"]), float(intr["cy"]),
            int(intr.get("width", 0)), int(intr.get("height", 0)))


def scale_intrinsics(fx, fy, cx, cy, src_w, src_h, dst_w, dst_h):
    """Per-axis rescale; a non-positive source size means "unknown" and returns the input unchanged."""
    if src_w <= 0 or src_h <= 0:
        return fx, fy, cx, cy
    sx = float(dst_w) / float(src_w)
    sy = float(dst_h) / float(src_h)
    return fx * sx, fy * sy, cx * sx, cy * sy


def load_intrinsics_json(path):
    """(K float32 3x3, dist float32[5], width, height); dist_coeffs padded / cut to 5 entries."""
    with open(path, "r") as f:
        J = json.load(f)
    K = np.array([[float(J["fx"]), 0.0, float(J["cx"])], [0.0, float(J["fy"]), float(J["cy"])], [0.0, 0.0, 1.0]],
                 dtype=np.float32)
    dist = np.array(J.get("dist_coeffs", [0, 0, 0, 0, 0]), dtype=np.float32).reshape(-1)
    if dist.size < 5:
        dist = np.pad(dist, (0, 5 - dist.size)).astype(np.float32)
    elif dist.size > 5:
        dist = dist[:5].astype(np.float32)
    return K, dist, int(J["width"]), int(J["height"])


def load_intrinsics(json_path):
    """RealSense dump {fx,fy,ppx,ppy,coeffs,width,height} -> (K float64, dist float64, (w, h))."""
    with open(json_path, "r") as f:
        d = json.load(f)
    K = np.array([[d["fx"], 0.0, d["ppx"]], [0.0, d["fy"], d["ppy"]], [0.0, 0.0, 1.0]], dtype=np.float64)
    dist = np.array(d.get("coeffs", [0, 0, 0, 0, 0]), dtype=np.float64)
    return K, dist, (int(d["width"]), int(d["height"]))


def read_depth_to_color_extrinsics(json_path):
    """{R_dc, t_dc} -> (R float64 3x3 row-major as stored, t float64[3])."""
    with open(json_path, "r") as f:
        data = json.load(f)
    return (np.array(data["R_dc"], dtype=np.float64).reshape(3, 3), np.array(data["t_dc"], dtype=np.float64).reshape(3))


def load_extrinsics(json_path):
    """Either dialect: Femto {R, t} (fetch_intrinsics.py:81-96) or RealSense {R_dc, t_dc}."""
    with open(json_path, "r") as f:
        data = json.load(f)
    if "R_dc" in data:
        R, t = data["R_dc"], data["t_dc"]
    elif "R" in data:
        R, t = data["R"], data["t"]
    else:
        raise KeyError(f"Missing 'R'/'R_dc' in extrinsics JSON: {json_path}")
    return np.array(R, dtype=np.float64).reshape(3, 3), np.array(t, dtype=np.float64).reshape(3)


def load_camera(json_path) -> Camera:
    """Any of the three intrinsics dialects -> Camera."""
    json_path = Path(json_path)
    if not json_path.exists():
        raise FileNotFoundError(f"Intrinsics JSON not found: {json_path}")
    data = json.loads(json_path.read_text())
    intr = data["color_intrinsics"] if (isinstance(data, dict) and "color_intrinsics" in data) else data
    if "ppx" in intr:
        cx, cy = float(intr["ppx"]), float(intr["ppy"])
        dist = list(intr.get("coeffs", [0, 0, 0, 0, 0]))
        model = _RS_MODELS.get(str(intr.get("distortion_model", intr.get("model", "none"))).lower(), "none")
    else:
        for k in ("fx", "fy", "cx", "cy"):
            if k not in intr:
                raise KeyError(f"Missing '{k}' in intrinsics JSON: {json_path}")
        cx, cy = float(intr["cx"]), float(intr["cy"])
        dist = list(intr.get("dist_coeffs", [0, 0, 0, 0, 0]))
        # cv2.calibrateCamera coefficients are the forward Brown-Conrady model
        model = "brown_conrady" if any(float(c) != 0.0 for c in dist) else "none"
    dist = [float(c) for c in dist[:5]] + [0.0] * max(0, 5 - len(dist))
    if not any(dist):
        model = "none"
    return Camera(float(intr["fx"]), float(intr["fy"]), cx, cy, int(intr.get("width", 0)), int(intr.get("height", 0)),
                  tuple(dist), model)


def load_transform_matrix(txt_path) -> np.ndarray:
    """4x4 pose text file (np.loadtxt layout); ValueError unless the shape is (4, 4)."""
    txt_path = Path(txt_path)
    if not txt_path.exists():
        raise FileNotFoundError(f"Transform file not found: {txt_path}")
    T = np.loadtxt(str(txt_path))
    if T.shape != (4, 4):
        raise ValueError(f"Expected 4x4 matrix, got shape {T.shape}")
    if not np.allclose(T[3, :], np.array([0, 0, 0, 1]), atol=1e-6):
        print(f"[WARN] Last row is {T[3, :]}, expected [0, 0, 0, 1]")
    return T
