"""ctypes binding of oracle/liboracle.so (oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_MODELS = {"none": 0, "brown_conrady": 1, "inverse_brown_conrady": 2, "modified_brown_conrady": 3}


class OrcCam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("ppx", C.c_float), ("ppy", C.c_float),
                ("coeffs", C.c_float * 5), ("model", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_register_z16.restype = C.c_int
        _lib.orc_voxel_down_sample.restype = C.c_int64
        _lib.orc_deproject_masked.restype = C.c_int64
    return _lib


def make_cam(cam: dict) -> OrcCam:
    c = OrcCam()
    c.fx, c.fy, c.ppx, c.ppy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    d = list(cam.get("dist", [0, 0, 0, 0, 0]))
    for i in range(5):
        c.coeffs[i] = d[i]
    c.model = _MODELS[cam.get("model", "none")]
    c.width, c.height = cam["width"], cam["height"]
    return c


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def register_depth_to_color(depth_u16, dcam, ccam, R_colmajor, t, depth_units=0.001):
    """One frame.  Returns (aligned u16 [Hc,Wc], winner i32 [Hc,Wc])."""
    depth = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    Hd, Wd = depth.shape
    Hc, Wc = ccam["height"], ccam["width"]
    out = np.empty((Hc, Wc), np.uint16)
    win = np.empty((Hc, Wc), np.int32)
    R = np.ascontiguousarray(R_colmajor, dtype=np.float32).reshape(9)
    tt = np.ascontiguousarray(t, dtype=np.float32).reshape(3)
    dc, cc = make_cam(dcam), make_cam(ccam)
    lib().orc_register_z16(_p(depth, C.c_uint16), Hd, Wd, C.byref(dc), C.byref(cc), _p(R, C.c_float), _p(tt, C.c_float),
                           C.c_float(depth_units), _p(out, C.c_uint16), _p(win, C.c_int32))
    return out, win


def voxel_down_sample(points, colors, voxel_size):
    """Returns (keys [M,3] i32, centroids [M,3], colours [M,3] or None, counts [M]) in first-insertion order."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = pts.shape[0]
    cols = None if colors is None else np.ascontiguousarray(colors, dtype=np.float64)
    cap = max(n, 1)
    keys = np.empty((cap, 3), np.int32)
    cent = np.empty((cap, 3), np.float64)
    col = np.empty((cap, 3), np.float64)
    cnt = np.empty(cap, np.int64)
    m = lib().orc_voxel_down_sample(_p(pts, C.c_double), None if cols is None else _p(cols, C.c_double), C.c_int64(n),
                                    C.c_double(voxel_size), _p(keys, C.c_int32), _p(cent, C.c_double),
                                    _p(col, C.c_double), _p(cnt, C.c_int64), C.c_int64(cap))
    if m < 0:
        raise ValueError({-1: "voxel_size <= 0", -2: "voxel_size is too small", -3: "capacity"}.get(m, str(m)))
    return keys[:m].copy(), cent[:m].copy(), (None if cols is None else col[:m].copy()), cnt[:m].copy()


def deproject_masked(depth_u16, bgr, mask, fx, fy, cx, cy, scale=0.001, r_max=None):
    depth = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    H, W = depth.shape
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    pts = np.empty((H * W, 3), np.float64)
    cols = np.empty((H * W, 3), np.float64)
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    m = lib().orc_deproject_masked(_p(depth, C.c_uint16), _p(bgr, C.c_uint8), None if mk is None else _p(mk, C.c_uint8),
                                   H, W, C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                   C.c_float(scale), int(r_max is not None), C.c_double(r_max or 0.0),
                                   _p(pts, C.c_double), _p(cols, C.c_double))
    return pts[:m], cols[:m]
