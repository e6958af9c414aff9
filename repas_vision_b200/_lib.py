"""ctypes binding of include/repas_vision.h (librepasvision.so).

There is no CPU fallback: if the shared library is missing, or no sm_100-class GPU is
visible when a compute entry point is used, the caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# RV_LIBRARY_PATH: a differently tuned build of the same sources (tools/k1_sweep.sh builds such variants for A/B runs)
SO_PATH = os.environ.get("RV_LIBRARY_PATH") or os.path.join(HERE, "librepasvision.so")

# enums (mirror include/repas_vision.h)
RV_OK, RV_EINVAL, RV_ECAPACITY, RV_ECUDA, RV_EALIGN, RV_EWORKSPACE = range(6)
RV_F32, RV_F64 = 0, 1
DIST_MODELS = {"none": 0, "brown_conrady": 1, "inverse_brown_conrady": 2, "modified_brown_conrady": 3}
UNIT_RULES = {"mul_f32": 0, "div_f32": 1, "div_f64": 2}
DEPTH_KINDS = {"u16": 0, "f32": 1}
MODES = {"compact_ordered": 0, "compact_unordered": 1, "dense_zero": 2, "dense_nan": 3, "compact_packed": 4}
COLOR_SCALES = {"unit": 0, "255": 1, "packed8": 2}
COLOR_FORMATS = {"bgr": 0, "nv12": 1}
GEOMETRIES = {"reference": 0, "sdk_f32": 1}
KERNELS = {"auto": 0, "generic": 1, "tma": 2}

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class RvCam(C.Structure):
    _fields_ = [("fx", c_f64), ("fy", c_f64), ("cx", c_f64), ("cy", c_f64), ("dist", c_f64 * 5), ("model", c_i32),
                ("width", c_i32), ("height", c_i32), ("reserved", c_i32)]


class RvDeprojectParams(C.Structure):
    _fields_ = [("cam", RvCam), ("depth_kind", c_i32), ("unit_rule", c_i32), ("unit_scale", c_f64),
                ("use_seg_mask", c_i32), ("invert_mask", c_i32), ("use_depth_trunc", c_i32), ("use_zclip", c_i32),
                ("depth_trunc", c_f64), ("z_min", c_f64), ("z_max", c_f64), ("use_radius", c_i32), ("use_aabb", c_i32),
                ("r_max", c_f64), ("aabb_min", c_f64 * 3), ("aabb_max", c_f64 * 3), ("mode", c_i32),
                ("out_dtype", c_i32), ("color_scale", c_i32), ("kernel_select", c_i32), ("color_format", c_i32),
                ("geometry", c_i32)]


# name -> (restype, argtypes); every symbol include/repas_vision.h declares
SIGNATURES = {
    "rv_abi_version": (C.c_int, []),
    "rv_sizeof_cam": (C.c_int, []),
    "rv_sizeof_deproject_params": (C.c_int, []),
    "rv_build_info": (C.c_char_p, []),
    "rv_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "rv_destroy": (C.c_int, [c_vp]),
    "rv_last_error": (C.c_char_p, [c_vp]),
    "rv_status_string": (C.c_char_p, [C.c_int]),
    "rv_sm_count": (C.c_int, [c_vp]),
    "rv_launch_count": (c_i64, [c_vp]),
    "rv_depth_to_meters": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_f64, c_vp, c_vp]),
    "rv_register_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rv_register_depth_to_color": (C.c_int, [c_vp, c_vp, C.c_int, C.POINTER(RvCam), C.POINTER(RvCam),
                                             C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, c_vp, c_vp, c_vp,
                                             C.c_size_t, c_vp]),
    "rv_build_ray_table": (C.c_int, [c_vp, C.POINTER(RvCam), c_vp, c_vp]),
    "rv_deproject_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rv_deproject_mask": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(RvDeprojectParams), c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp,
                                    C.c_size_t, c_vp]),
    "rv_filter_workspace_bytes": (C.c_size_t, [c_i64]),
    "rv_filter_cloud": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.c_int, C.POINTER(RvDeprojectParams), c_vp,
                                  c_i64, c_vp, c_vp, c_vp, C.c_size_t, c_vp]),
    "rv_bounds_init": (C.c_int, [c_vp, c_vp, c_vp]),
    "rv_cloud_stats": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp]),
    "rv_transform_merge": (C.c_int, [c_vp, C.c_int, C.POINTER(c_vp), C.POINTER(c_i64), C.POINTER(c_i64),
                                     C.POINTER(c_f64), C.c_int, C.c_int, c_vp, c_i64, C.c_int, c_vp, c_vp]),
    "rv_voxel_workspace_bytes": (C.c_size_t, [c_i64]),
    "rv_voxel_downsample": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.c_int, c_f64, c_vp, c_vp, c_i64, C.c_int,
                                      c_i64, c_vp, c_vp, c_vp, c_vp, C.c_size_t, c_vp]),
    "rv_fuse_voxel": (C.c_int, [c_vp, C.c_int, C.POINTER(c_vp), C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_f64), C.c_int,
                                C.c_int, c_f64, c_vp, c_i64, C.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, C.c_size_t, c_vp]),
    "rv_pack_ply_records": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "rv_unpack_ply_records": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.POINTER(c_i32), C.c_int, C.POINTER(c_i32), c_vp, c_i64,
                                        C.c_int, c_vp]),
    "rv_knn_workspace_bytes": (C.c_size_t, [c_i64]),
    "rv_knn_mean_distance": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.c_int, c_vp, c_vp, C.c_size_t, c_vp]),
    "rv_estimate_normals": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_f64, C.c_int, C.POINTER(c_f64), c_vp, c_i64, c_vp,
                                      C.c_size_t, c_vp]),
    "rv_orient_normals": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, C.POINTER(c_f64), c_vp]),
    "rv_statistical_outlier_mask": (C.c_int, [c_vp, c_vp, c_i64, c_f64, c_vp, c_vp, c_vp]),
    "rv_select_by_mask": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp,
                                    C.c_size_t, c_vp]),
    "rv_nn_index_build": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_f64, c_vp, C.c_size_t, c_vp]),
    "rv_nn_search": (C.c_int, [c_vp, c_vp, C.c_size_t, c_i64, c_vp, c_i64, c_i64, C.c_int, c_f64, c_vp, c_vp]),
    "rv_icp_sums_bytes": (C.c_size_t, []),
    "rv_icp_sums": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, c_vp, c_vp,
                              c_vp]),
    "rv_icp_state_bytes": (C.c_size_t, []),
    "rv_icp_begin": (C.c_int, [c_vp, c_vp, C.POINTER(c_f64), C.c_int, c_f64, c_f64, c_i64, c_vp]),
    "rv_icp_iterate": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_i64, c_i64, C.c_int, c_vp, C.c_size_t, c_i64, c_vp, c_i64,
                                 C.c_int, c_vp, c_i64, c_f64, c_vp, c_vp, c_vp]),
    "rv_median_depth_window": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_i64, C.c_int, c_vp, c_vp]),
    "rv_nv12_to_bgr": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
}

_lock = threading.Lock()
_lib = None
_ctxs: dict[int, "Context"] = {}


def load():
    """dlopen the in-tree shared library and declare every prototype."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: the sm_100a CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.rv_sizeof_cam() != C.sizeof(RvCam) or lib.rv_sizeof_deproject_params() != C.sizeof(RvDeprojectParams):
            raise RuntimeError("ctypes struct layout does not match the compiled library")
        _lib = lib
        return lib


class RvError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"repas_vision status {status}: {text}")
        self.status = status


class Context:
    """One rv_ctx per CUDA device."""

    def __init__(self, device: int):
        self.lib = load()
        self.device = device
        h = c_vp()
        st = self.lib.rv_create(device, C.byref(h))
        if st != RV_OK:
            raise RuntimeError(
                f"rv_create(device={device}) failed with status {st} ({self.lib.rv_status_string(st).decode()}): "
                "a Blackwell-class (sm_100) CUDA device is required; there is no CPU fallback")
        self.handle = h

    def check(self, status: int):
        if status != RV_OK:
            raise RvError(status, self.lib.rv_last_error(self.handle).decode() or
                          self.lib.rv_status_string(status).decode())

    @property
    def sm_count(self) -> int:
        return self.lib.rv_sm_count(self.handle)

    @property
    def launches(self) -> int:
        return int(self.lib.rv_launch_count(self.handle))


def context(device: int) -> Context:
    with _lock:
        ctx = _ctxs.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _ctxs[device] = ctx
    return ctx


def make_cam(fx, fy, cx, cy, width, height, dist=None, model="none") -> RvCam:
    cam = RvCam()
    cam.fx, cam.fy, cam.cx, cam.cy = float(fx), float(fy), float(cx), float(cy)
    d = [0.0] * 5 if dist is None else [float(v) for v in list(dist)[:5]] + [0.0] * max(0, 5 - len(list(dist)))
    for i in range(5):
        cam.dist[i] = d[i]
    cam.model = DIST_MODELS[model] if isinstance(model, str) else int(model)
    cam.width, cam.height = int(width), int(height)
    return cam
