"""Host-side pose math: AprilTag corners -> solvePnP -> 4x4, and the fusion convention.

The pose stays on the host (microseconds of cv2 per view); the GPU consumes only the 4x4.

* solve_pnp_with_best_obj_order ... femto_bolt_code/scripts/final_view.py:171-225 (ten more copies,
  e.g. april_tag_bg_removal_pl.py:116-142): eight candidate corner orders x
  cv2.solvePnP(SOLVEPNP_IPPE_SQUARE); score = mean reprojection error in pixels
  + 1000 when the tag lands behind the camera; ties keep the earlier candidate.
* to_4x4 .......................... femto_bolt_code/scripts/final_view_with_cad.py:138-142,
  final_view.py:353-355 (T[:3,:3] = R; T[:3,3] = t)
* world_from_camera ............... SURVEY.md Appendix D.4: world := tag frame, a view's cloud moves by
  inv(T_cam_tag) = [R^T | -R^T t] computed in float64.
"""
from __future__ import annotations

import numpy as np

_ORDERS = ("TL,TR,BR,BL", "TR,BR,BL,TL", "BR,BL,TL,TR", "BL,TL,TR,BR",
           "TR,TL,BL,BR", "TL,BL,BR,TR", "BL,BR,TR,TL", "BR,TR,TL,BL")


def _corner_table(tag_size_m: float) -> dict:
    h = float(tag_size_m) / 2.0
    return {"TL": (-h, -h, 0.0), "TR": (+h, -h, 0.0), "BR": (+h, +h, 0.0), "BL": (-h, +h, 0.0)}


def compute_reproj_error(obj_pts, img_pts, rvec, tvec, K, dist) -> float:
    """Mean L2 reprojection error in pixels (final_view.py:164-169)."""
    import cv2
    proj, _ = cv2.projectPoints(obj_pts, rvec, tvec, K, dist)
    return float(np.mean(np.linalg.norm(proj.reshape(-1, 2) - np.asarray(img_pts).reshape(-1, 2), axis=1)))


def solve_pnp_with_best_obj_order(img_corners_px, K, dist, tag_size_m):
    """Returns (obj_pts 4x3, rvec 3x1, tvec 3x1, err_px, order_label); RuntimeError when every order fails."""
    import cv2
    table = _corner_table(tag_size_m)
    img = np.asarray(img_corners_px)
    best, best_score, best_err, best_label = None, np.inf, None, None
    for label in _ORDERS:
        obj_pts = np.array([table[name] for name in label.split(",")], dtype=np.float64)
        ok, rvec, tvec = cv2.solvePnP(obj_pts, img, K, dist, flags=cv2.SOLVEPNP_IPPE_SQUARE)
        if not ok:
            continue
        err = compute_reproj_error(obj_pts, img, rvec, tvec, K, dist)
        score = err + (1000.0 if float(tvec[2, 0] if tvec.ndim == 2 else tvec[2]) <= 0 else 0.0)
        if score < best_score:
            best, best_score, best_err, best_label = (obj_pts, rvec, tvec), score, err, label
    if best is None:
        raise RuntimeError("solvePnP failed for all candidate corner orders.")
    return best[0], best[1], best[2], best_err, best_label


def to_4x4(R, t) -> np.ndarray:
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = np.asarray(R, dtype=np.float64).reshape(3, 3)
    T[:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
    return T


def pose_from_tag_corners(img_corners_px, K, dist, tag_size_m) -> np.ndarray:
    """Corners -> T_cam_tag (tag -> camera, OpenCV basis), the composition final_view.py:341-355 performs."""
    import cv2
    _, rvec, tvec, _, _ = solve_pnp_with_best_obj_order(img_corners_px, K, dist, tag_size_m)
    R, _ = cv2.Rodrigues(rvec)
    return to_4x4(R, tvec)


def invert_rigid(T) -> np.ndarray:
    T = np.asarray(T, dtype=np.float64)
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4, dtype=np.float64)
    out[:3, :3] = R.T
    out[:3, 3] = -(R.T @ t)
    return out


def world_from_camera(T_cam_tag) -> np.ndarray:
    """The transform applied to a view's cloud in the four-pose fusion (tag frame = world)."""
    return invert_rigid(T_cam_tag)


def transform_point_tag_local_to_camera(p_tag, R, t):
    """R @ p + t (april_tag_bg_removal_pl.py:177-179)."""
    return np.asarray(R, dtype=np.float64) @ np.asarray(p_tag, dtype=np.float64) + np.asarray(t, dtype=np.float64).reshape(3)
