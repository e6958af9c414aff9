"""Torch-tensor front of the C ABI: pointer extraction, workspace allocation, stream hand-off.

PyTorch is plumbing here (device memory from its caching allocator, the current stream);
every computation is a kernel of librepasvision.so reached through include/repas_vision.h.
There is no fallback: without the library or an sm_100 device these raise RuntimeError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .calibration import Camera

_TORCH_DT = {"f32": torch.float32, "f64": torch.float64}
_RV_DT = {torch.float32: _lib.RV_F32, torch.float64: _lib.RV_F64}


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("repas_vision_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"repas_vision_b200 computes on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ctx_for(t_or_dev) -> _lib.Context:
    dev = t_or_dev.device if isinstance(t_or_dev, torch.Tensor) else t_or_dev
    return _lib.context(dev.index)


def stream_ptr(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def pstride(t: torch.Tensor) -> int:
    """Plane stride (elements) of an SoA cloud tensor [planes, n]; column slices of a batch are fine, the points of a
    plane must be contiguous."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("cloud planes must be rows of a 2-D tensor with unit column stride")
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def cam_struct(cam: Camera, width=None, height=None) -> _lib.RvCam:
    return _lib.make_cam(cam.fx, cam.fy, cam.cx, cam.cy, cam.width if width is None else width,
                         cam.height if height is None else height, cam.dist, cam.model)


def to_device(a, dev: torch.device, dtype=None) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous tensor on dev.  No copy for a conforming CUDA tensor."""
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a))
    elif isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.as_tensor(np.asarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


# ---------------------------------------------------------------------------- K1
def deproject(depth, bgr, mask, cam: Camera, *, depth_kind, unit_rule="mul_f32", unit_scale=None, invert_mask=False,
              depth_trunc=None, z_clip=None, r_max=None, aabb=None, mode="compact_ordered", out_dtype="f32",
              color_scale="unit", want_valid=False, want_src_index=False, frame_capacity=None, out=None,
              rays=None, kernel="auto", color_format="bgr", geometry="reference"):
    """depth [B,H,W] (uint16 or float32), bgr [B,H,W,3] uint8 (or NV12 frames [B,H*3/2,W] with color_format="nv12") or
    None, mask [B,H,W] uint8 or None, all on one CUDA device.  Returns dict(data [6 or 3, B*cap], counts [B] int64, cap,
    valid, src_index); with color_scale="packed8" data has four float32 planes, the last one holding the bytes r,g,b,0
    of every point."""
    dev = depth.device
    ctx = ctx_for(dev)
    B, H, W = depth.shape
    P = H * W
    p = _lib.RvDeprojectParams()
    p.cam = cam_struct(cam, W, H)
    p.depth_kind = _lib.DEPTH_KINDS[depth_kind]
    p.unit_rule = _lib.UNIT_RULES[unit_rule]
    p.unit_scale = float(unit_scale) if unit_scale is not None else (0.001 if unit_rule == "mul_f32" else 1000.0)
    p.use_seg_mask = int(mask is not None)
    p.invert_mask = int(bool(invert_mask))
    p.use_depth_trunc = int(depth_trunc is not None)
    p.depth_trunc = float(depth_trunc or 0.0)
    p.use_zclip = int(z_clip is not None)
    if z_clip is not None:
        p.z_min = -np.inf if z_clip[0] is None else float(z_clip[0])
        p.z_max = np.inf if z_clip[1] is None else float(z_clip[1])
    p.use_radius = int(r_max is not None)
    p.r_max = float(r_max or 0.0)
    p.use_aabb = int(aabb is not None)
    if aabb is not None:
        for i in range(3):
            p.aabb_min[i] = float(aabb[0][i])
            p.aabb_max[i] = float(aabb[1][i])
    p.mode = _lib.MODES[mode]
    p.out_dtype = _lib.RV_F32 if out_dtype == "f32" else _lib.RV_F64
    p.color_scale = _lib.COLOR_SCALES[color_scale]
    p.kernel_select = _lib.KERNELS[kernel]
    p.color_format = _lib.COLOR_FORMATS[color_format]
    p.geometry = _lib.GEOMETRIES[geometry]
    if bgr is not None:
        want = (B, H, W, 3) if color_format == "bgr" else (B, H * 3 // 2, W)
        if tuple(bgr.shape) != want or bgr.dtype != torch.uint8:
            raise ValueError(f"colour frames must be uint8 {want} for color_format={color_format!r}, got {tuple(bgr.shape)}")
    if color_scale == "packed8" and out_dtype != "f32":
        raise ValueError("color_scale='packed8' stores one 32-bit colour word per point: float32 output only")
    cap = int(frame_capacity) if frame_capacity is not None else P
    planes = 3 if bgr is None else (4 if color_scale == "packed8" else 6)
    packed = mode == "compact_packed"
    plane_stride = B * cap
    if out is None:
        out = torch.empty((planes, plane_stride), dtype=_TORCH_DT[out_dtype], device=dev)
    else:
        if out.dtype != _TORCH_DT[out_dtype] or out.device != dev or out.dim() != 2 or out.shape[0] < planes \
                or out.shape[1] < plane_stride or not out.is_contiguous():
            raise ValueError("out must be a contiguous [planes, >= B*capacity] tensor of the output dtype on the input device")
        plane_stride = out.shape[1]
    counts = torch.empty(B + 1 if packed else B, dtype=torch.int64, device=dev)
    valid = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if want_valid else None
    src = torch.empty(plane_stride, dtype=torch.int32, device=dev) if want_src_index else None
    if cam.distorted and cam.model != "modified_brown_conrady" and rays is None:
        rays = ray_table(cam, H, W, dev)
    ws_bytes = ctx.lib.rv_deproject_workspace_bytes(B, H, W)
    ws = workspace(ws_bytes, dev)
    ctx.check(ctx.lib.rv_deproject_mask(ctx.handle, ptr(depth), ptr(bgr), ptr(mask), ptr(rays), B, H, W, C.byref(p),
                                        ptr(out), plane_stride, cap, ptr(valid), ptr(src), ptr(counts), ptr(ws),
                                        ws.numel(), stream_ptr(dev)))
    return dict(data=out, counts=counts, cap=cap, valid=valid, src_index=src, planes=planes, plane_stride=plane_stride)


_ray_cache: dict = {}


def ray_table(cam: Camera, H: int, W: int, dev: torch.device) -> torch.Tensor:
    key = (cam.fx, cam.fy, cam.cx, cam.cy, tuple(cam.dist), cam.model, H, W, dev.index)
    t = _ray_cache.get(key)
    if t is None:
        ctx = ctx_for(dev)
        t = torch.empty((H, W, 2), dtype=torch.float64, device=dev)
        cs = cam_struct(cam, W, H)
        ctx.check(ctx.lib.rv_build_ray_table(ctx.handle, C.byref(cs), ptr(t), stream_ptr(dev)))
        if len(_ray_cache) > 16:
            _ray_cache.clear()
        _ray_cache[key] = t
    return t


def depth_to_meters(depth_u16: torch.Tensor, rule="mul_f32", scale=None) -> torch.Tensor:
    dev = depth_u16.device
    ctx = ctx_for(dev)
    out = torch.empty(depth_u16.shape, dtype=torch.float32, device=dev)
    s = float(scale) if scale is not None else (0.001 if rule == "mul_f32" else 1000.0)
    ctx.check(ctx.lib.rv_depth_to_meters(ctx.handle, ptr(depth_u16), depth_u16.numel(), _lib.UNIT_RULES[rule], s, ptr(out),
                                         stream_ptr(dev)))
    return out


# ---------------------------------------------------------------------------- K2
def register(depth: torch.Tensor, dcam: Camera, ccam: Camera, R_colmajor, t, depth_units=0.001, want_winner=False,
             chunk_frames=None):
    """depth [B,Hd,Wd] uint16 -> (aligned [B,Hc,Wc] uint16, winner [B,Hc,Wc] int32 or None)."""
    dev = depth.device
    ctx = ctx_for(dev)
    B, Hd, Wd = depth.shape
    Hc, Wc = ccam.height, ccam.width
    out = torch.empty((B, Hc, Wc), dtype=torch.uint16, device=dev)
    win = torch.empty((B, Hc, Wc), dtype=torch.int32, device=dev) if want_winner else None
    nb = ctx.lib.rv_register_workspace_bytes(B if chunk_frames is None else min(B, chunk_frames), Hd, Wd, Hc, Wc)
    ws = workspace(nb, dev)
    R = (C.c_float * 9)(*[float(v) for v in np.asarray(R_colmajor, dtype=np.float32).reshape(9)])
    tt = (C.c_float * 3)(*[float(v) for v in np.asarray(t, dtype=np.float32).reshape(3)])
    dc, cc = cam_struct(dcam, Wd, Hd), cam_struct(ccam)
    ctx.check(ctx.lib.rv_register_depth_to_color(ctx.handle, ptr(depth), B, C.byref(dc), C.byref(cc), R, tt,
                                                 C.c_float(depth_units), ptr(out), ptr(win), ptr(ws), ws.numel(),
                                                 stream_ptr(dev)))
    return out, win


# --------------------------------------------------------------------- cloud ops
def filter_cloud(data: torch.Tensor, n: int, has_color: bool, *, z_clip=None, r_max=None, aabb=None, want_index=False):
    """Ordered compaction by the cloud predicates: (planes [P, n], count int64[1][, source indices int64 [n]])."""
    dev = data.device
    ctx = ctx_for(dev)
    p = _lib.RvDeprojectParams()
    p.use_zclip = int(z_clip is not None)
    if z_clip is not None:
        p.z_min = -np.inf if z_clip[0] is None else float(z_clip[0])
        p.z_max = np.inf if z_clip[1] is None else float(z_clip[1])
    p.use_radius = int(r_max is not None)
    p.r_max = float(r_max or 0.0)
    p.use_aabb = int(aabb is not None)
    if aabb is not None:
        for i in range(3):
            p.aabb_min[i] = float(aabb[0][i])
            p.aabb_max[i] = float(aabb[1][i])
    out = torch.empty((data.shape[0], max(n, 1)), dtype=data.dtype, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = workspace(ctx.lib.rv_filter_workspace_bytes(n), dev)
    index = torch.empty(max(n, 1), dtype=torch.int64, device=dev) if want_index else None
    ctx.check(ctx.lib.rv_filter_cloud(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], int(has_color), C.byref(p),
                                      ptr(out), out.shape[1], ptr(count), ptr(index), ptr(ws), ws.numel(), stream_ptr(dev)))
    return (out, count, index) if want_index else (out, count)


def cloud_stats(data: torch.Tensor, n: int) -> np.ndarray:
    """float64 [9] on the host: min xyz, max xyz, sum xyz of the first n points (one kernel, one 72-byte read-back)."""
    dev = data.device
    ctx = ctx_for(dev)
    stats = torch.empty(9, dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.rv_cloud_stats(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], ptr(stats), stream_ptr(dev)))
    return stats.cpu().numpy()


def planes_to_host(data: torch.Tensor, first: int, rows: int, n: int) -> np.ndarray:
    """(n, rows) float64 numpy copy of planes [first, first+rows) of an SoA cloud: one device->host copy per plane (each is
    contiguous), transpose and widening on the host -- no device kernel."""
    host = torch.empty((rows, max(n, 1)), dtype=data.dtype)
    for r in range(rows):
        host[r, :n].copy_(data[first + r, :n])
    return np.ascontiguousarray(host[:, :n].numpy().T, dtype=np.float64)


def transform_merge(views, Ts, has_color: bool, out_dtype=None, want_bounds=False):
    """views: list of (data [planes, stride], n).  Ts: list of 4x4.  Returns (merged [planes, total], total, bounds)."""
    dev = views[0][0].device
    ctx = ctx_for(dev)
    in_dt = views[0][0].dtype
    if any(v[0].dtype != in_dt for v in views):
        raise ValueError("all views must share one dtype")
    out_dt = in_dt if out_dtype is None else _TORCH_DT[out_dtype]
    nv = len(views)
    total = int(sum(v[1] for v in views))
    planes = 6 if has_color else 3
    out = torch.empty((planes, max(total, 1)), dtype=out_dt, device=dev)
    ptrs = (C.c_void_p * nv)(*[v[0].data_ptr() for v in views])
    strides = (C.c_int64 * nv)(*[pstride(v[0]) for v in views])
    ns = (C.c_int64 * nv)(*[int(v[1]) for v in views])
    Tflat = np.ascontiguousarray(np.stack([np.asarray(T, dtype=np.float64).reshape(4, 4) for T in Ts]))  # alive during the call
    Tc = Tflat.ctypes.data_as(C.POINTER(C.c_double))
    bounds = None
    if want_bounds:
        bounds = torch.empty(6, dtype=torch.float64, device=dev)
        ctx.check(ctx.lib.rv_bounds_init(ctx.handle, ptr(bounds), stream_ptr(dev)))
    ctx.check(ctx.lib.rv_transform_merge(ctx.handle, nv, ptrs, strides, ns, Tc, _RV_DT[in_dt], int(has_color), ptr(out),
                                         out.shape[1], _RV_DT[out_dt], ptr(bounds), stream_ptr(dev)))
    return out, total, bounds


def transform_xyz_into(src: torch.Tensor, n: int, T, out: torch.Tensor) -> None:
    """K3 on the coordinate planes only, into a caller-owned buffer of the same dtype (the ICP working copy: no allocation, no
    colour traffic)."""
    dev = src.device
    ctx = ctx_for(dev)
    ptrs = (C.c_void_p * 1)(src.data_ptr())
    strides = (C.c_int64 * 1)(pstride(src))
    ns = (C.c_int64 * 1)(int(n))
    Tc = (C.c_double * 16)(*np.asarray(T, dtype=np.float64).reshape(-1))
    ctx.check(ctx.lib.rv_transform_merge(ctx.handle, 1, ptrs, strides, ns, Tc, _RV_DT[src.dtype], 0, ptr(out), pstride(out),
                                         _RV_DT[out.dtype], None, stream_ptr(dev)))


def voxel_downsample(data: torch.Tensor, n: int, has_color: bool, voxel_size: float, *, bounds=None, out_dtype=None,
                     want_keys=False, want_counts=False, out_capacity=None):
    """Returns dict(data [planes, cap], m (int64 tensor[1]), keys [3,cap] int32 | None, counts [cap] | None)."""
    dev = data.device
    ctx = ctx_for(dev)
    out_dt = data.dtype if out_dtype is None else _TORCH_DT[out_dtype]
    cap = int(n if out_capacity is None else out_capacity)
    planes = 6 if has_color else 3
    out = torch.empty((planes, max(cap, 1)), dtype=out_dt, device=dev)
    keys = torch.empty((3, max(cap, 1)), dtype=torch.int32, device=dev) if want_keys else None
    cnts = torch.empty(max(cap, 1), dtype=torch.int32, device=dev) if want_counts else None
    m = torch.empty(1, dtype=torch.int64, device=dev)  # always written by the call (0 for an empty cloud, -1 on overflow)
    ws = workspace(ctx.lib.rv_voxel_workspace_bytes(n), dev)
    if not (float(voxel_size) > 0.0):
        raise ValueError("voxel_size <= 0")
    ctx.check(ctx.lib.rv_voxel_downsample(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], int(has_color),
                                          float(voxel_size), ptr(bounds), ptr(out), out.shape[1], _RV_DT[out_dt],
                                          min(cap, out.shape[1]) if cap > 0 else 0, ptr(keys), ptr(cnts), ptr(m), ptr(ws),
                                          ws.numel(), stream_ptr(dev)))
    return dict(data=out, m=m, keys=keys, counts=cnts, cap=out.shape[1])


def fuse_voxel(views, Ts, has_color: bool, voxel_size: float, *, out_dtype=None, want_keys=False, want_counts=False):
    """K3 + K4 in one call (rv_fuse_voxel): views = list of (data [planes, stride], n), Ts = list of 4x4 moving each view into
    the common frame.  The merged cloud is never written.  Returns the dict of voxel_downsample."""
    dev = views[0][0].device
    ctx = ctx_for(dev)
    in_dt = views[0][0].dtype
    if any(v[0].dtype != in_dt for v in views):
        raise ValueError("all views must share one dtype")
    out_dt = in_dt if out_dtype is None else _TORCH_DT[out_dtype]
    nv = len(views)
    total = int(sum(v[1] for v in views))
    planes = 6 if has_color else 3
    out = torch.empty((planes, max(total, 1)), dtype=out_dt, device=dev)
    keys = torch.empty((3, max(total, 1)), dtype=torch.int32, device=dev) if want_keys else None
    cnts = torch.empty(max(total, 1), dtype=torch.int32, device=dev) if want_counts else None
    m = torch.empty(1, dtype=torch.int64, device=dev)  # always written by the call (0 for an empty cloud, -1 on overflow)
    ws = workspace(ctx.lib.rv_voxel_workspace_bytes(total), dev)
    ptrs = (C.c_void_p * nv)(*[v[0].data_ptr() for v in views])
    strides = (C.c_int64 * nv)(*[pstride(v[0]) for v in views])
    ns = (C.c_int64 * nv)(*[int(v[1]) for v in views])
    Tflat = np.ascontiguousarray(np.stack([np.asarray(T, dtype=np.float64).reshape(4, 4) for T in Ts]))  # alive during the call
    Tc = Tflat.ctypes.data_as(C.POINTER(C.c_double))
    ctx.check(ctx.lib.rv_fuse_voxel(ctx.handle, nv, ptrs, strides, ns, Tc, _RV_DT[in_dt], int(has_color), float(voxel_size),
                                    ptr(out), out.shape[1], _RV_DT[out_dt], total, ptr(keys), ptr(cnts), ptr(m), ptr(ws),
                                    ws.numel(), stream_ptr(dev)))
    return dict(data=out, m=m, keys=keys, counts=cnts, cap=out.shape[1])


def pack_ply_records(data: torch.Tensor, n: int, has_color: bool, color_scale="unit", coord_dtype="f32") -> torch.Tensor:
    dev = data.device
    ctx = ctx_for(dev)
    rec = 3 * (4 if coord_dtype == "f32" else 8) + 3
    out = torch.empty(max(n, 1) * rec, dtype=torch.uint8, device=dev)
    ctx.check(ctx.lib.rv_pack_ply_records(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], int(has_color),
                                          _lib.COLOR_SCALES[color_scale], _lib.RV_F32 if coord_dtype == "f32" else _lib.RV_F64,
                                          ptr(out), stream_ptr(dev)))
    return out[:n * rec]


def unpack_ply_records(records: torch.Tensor, n: int, record_bytes: int, xyz_offset, coord_dtype: str, rgb_offset,
                       out_dtype: str = "f64") -> torch.Tensor:
    """uint8 vertex records as they lie in a binary PLY (on the GPU) -> SoA planes [3 or 6, n]."""
    dev = records.device
    ctx = ctx_for(dev)
    planes = 6 if rgb_offset is not None else 3
    out = torch.empty((planes, max(n, 1)), dtype=torch.float64 if out_dtype == "f64" else torch.float32, device=dev)
    xo = (C.c_int32 * 3)(*[int(v) for v in xyz_offset])
    ro = (C.c_int32 * 3)(*[int(v) for v in rgb_offset]) if rgb_offset is not None else None
    ctx.check(ctx.lib.rv_unpack_ply_records(ctx.handle, ptr(records), n, int(record_bytes), xo,
                                            _lib.RV_F32 if coord_dtype == "f32" else _lib.RV_F64, ro, ptr(out), pstride(out),
                                            _RV_DT[out.dtype], stream_ptr(dev)))
    return out


def knn_mean_distance(data: torch.Tensor, n: int, k: int) -> torch.Tensor:
    """Mean distance of every point to its k nearest neighbours, itself included (float64 [n])."""
    dev = data.device
    ctx = ctx_for(dev)
    out = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    nb = ctx.lib.rv_knn_workspace_bytes(n)
    ws = workspace(nb, dev)
    ctx.check(ctx.lib.rv_knn_mean_distance(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], int(k), ptr(out),
                                           ptr(ws), ws.numel(), stream_ptr(dev)))
    return out[:n]


def estimate_normals(data: torch.Tensor, n: int, radius: float, max_nn: int, camera_location=None) -> torch.Tensor:
    """float64 [3, n] unit normals (hybrid radius / k-nearest neighbourhoods), turned towards camera_location when given."""
    dev = data.device
    ctx = ctx_for(dev)
    out = torch.empty((3, max(n, 1)), dtype=torch.float64, device=dev)
    ws = workspace(ctx.lib.rv_knn_workspace_bytes(n), dev)
    cam = None if camera_location is None else (C.c_double * 3)(*[float(v) for v in np.asarray(camera_location).reshape(3)])
    ctx.check(ctx.lib.rv_estimate_normals(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], float(radius), int(max_nn),
                                          cam, ptr(out), pstride(out), ptr(ws), ws.numel(), stream_ptr(dev)))
    return out


def orient_normals(data: torch.Tensor, n: int, normals: torch.Tensor, camera_location) -> None:
    dev = data.device
    ctx = ctx_for(dev)
    cam = (C.c_double * 3)(*[float(v) for v in np.asarray(camera_location, dtype=np.float64).reshape(3)])
    ctx.check(ctx.lib.rv_orient_normals(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], ptr(normals), pstride(normals),
                                        cam, stream_ptr(dev)))


def nn_index_build(data: torch.Tensor, n: int, max_distance: float) -> torch.Tensor:
    """Hash-grid index over the first n points of `data` (the reference's KDTreeFlann(target)); returns the workspace that
    holds it."""
    dev = data.device
    ctx = ctx_for(dev)
    ws = workspace(ctx.lib.rv_knn_workspace_bytes(n), dev)
    ctx.check(ctx.lib.rv_nn_index_build(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], float(max_distance), ptr(ws),
                                        ws.numel(), stream_ptr(dev)))
    return ws


def nn_search(index_ws: torch.Tensor, n_indexed: int, query: torch.Tensor, n_query: int, max_distance: float,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """int32 [n_query]: nearest indexed point closer than max_distance, -1 without one."""
    dev = query.device
    ctx = ctx_for(dev)
    if out is None:
        out = torch.empty(max(n_query, 1), dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.rv_nn_search(ctx.handle, ptr(index_ws), index_ws.numel(), n_indexed, ptr(query), pstride(query), n_query,
                                   _RV_DT[query.dtype], float(max_distance), ptr(out), stream_ptr(dev)))
    return out


def icp_scratch(dev: torch.device) -> torch.Tensor:
    return torch.empty(int(ctx_for(dev).lib.rv_icp_sums_bytes()) // 8, dtype=torch.float64, device=dev)


def icp_sums(source: torch.Tensor, n_source: int, target: torch.Tensor, n_target: int, target_normals, nearest: torch.Tensor,
             point_to_plane: bool, scratch: torch.Tensor | None = None) -> torch.Tensor:
    """float64 [32] sums of one ICP estimation step over the correspondences (see include/repas_vision.h); a view of
    `scratch` (rv_icp_sums_bytes() bytes)."""
    dev = source.device
    ctx = ctx_for(dev)
    if scratch is None:
        scratch = icp_scratch(dev)
    ctx.check(ctx.lib.rv_icp_sums(ctx.handle, int(bool(point_to_plane)), ptr(source), pstride(source), n_source,
                                  _RV_DT[source.dtype], ptr(target), pstride(target), n_target, _RV_DT[target.dtype],
                                  ptr(target_normals), 0 if target_normals is None else pstride(target_normals), ptr(nearest),
                                  ptr(scratch), stream_ptr(dev)))
    return scratch[:32]


def icp_state(dev: torch.device) -> torch.Tensor:
    return torch.zeros(int(ctx_for(dev).lib.rv_icp_state_bytes()) // 8, dtype=torch.float64, device=dev)


def icp_begin(state: torch.Tensor, T_init, max_iteration: int, relative_fitness: float, relative_rmse: float, n_source: int) -> None:
    dev = state.device
    ctx = ctx_for(dev)
    Tc = (C.c_double * 16)(*np.asarray(T_init, dtype=np.float64).reshape(-1))
    ctx.check(ctx.lib.rv_icp_begin(ctx.handle, ptr(state), Tc, int(max_iteration), float(relative_fitness), float(relative_rmse),
                                   int(n_source), stream_ptr(dev)))


def icp_iterate(state: torch.Tensor, first: bool, steps: int, work: torch.Tensor, n_source: int, index_ws: torch.Tensor,
                target: torch.Tensor, n_target: int, target_normals: torch.Tensor, max_distance: float, nearest: torch.Tensor,
                scratch: torch.Tensor) -> None:
    """Queues `steps` point-to-plane iterations (and the initial evaluation when `first`) on the current stream."""
    dev = work.device
    ctx = ctx_for(dev)
    ctx.check(ctx.lib.rv_icp_iterate(ctx.handle, ptr(state), int(bool(first)), int(steps), ptr(work), pstride(work), n_source,
                                     _RV_DT[work.dtype], ptr(index_ws), index_ws.numel(), n_target, ptr(target), pstride(target),
                                     _RV_DT[target.dtype], ptr(target_normals), pstride(target_normals), float(max_distance),
                                     ptr(nearest), ptr(scratch), stream_ptr(dev)))


def statistical_outlier_mask(mean: torch.Tensor, std_ratio: float):
    """(keep uint8 [n], stats float64 [4] = cloud mean, std dev, threshold, points counted)."""
    dev = mean.device
    ctx = ctx_for(dev)
    n = int(mean.numel())
    keep = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    stats = torch.zeros(520, dtype=torch.float64, device=dev)  # 4 results + flag + 2 x 256 block sums
    ctx.check(ctx.lib.rv_statistical_outlier_mask(ctx.handle, ptr(mean), n, float(std_ratio), ptr(keep), ptr(stats),
                                                  stream_ptr(dev)))
    return keep[:n], stats[:4]


def select_by_mask(data: torch.Tensor, n: int, has_color: bool, keep: torch.Tensor, want_index: bool = True):
    """Ordered compaction by a uint8 mask: (planes [P, n], count int64[1], source indices int64 [n] | None)."""
    dev = data.device
    ctx = ctx_for(dev)
    out = torch.empty((data.shape[0], max(n, 1)), dtype=data.dtype, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    index = torch.empty(max(n, 1), dtype=torch.int64, device=dev) if want_index else None
    nb = ctx.lib.rv_filter_workspace_bytes(n)
    ws = workspace(nb, dev)
    ctx.check(ctx.lib.rv_select_by_mask(ctx.handle, ptr(data), pstride(data), n, _RV_DT[data.dtype], int(has_color), ptr(keep),
                                        ptr(out), pstride(out), ptr(count), ptr(index), ptr(ws), ws.numel(), stream_ptr(dev)))
    return out, count, index


def median_depth_window(depth_u16: torch.Tensor, uv: torch.Tensor, window: int) -> torch.Tensor:
    dev = depth_u16.device
    ctx = ctx_for(dev)
    H, W = depth_u16.shape
    n = uv.shape[0]
    out = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.rv_median_depth_window(ctx.handle, ptr(depth_u16), H, W, ptr(uv), n, int(window), ptr(out),
                                             stream_ptr(dev)))
    return out[:n]


def nv12_to_bgr(nv12: torch.Tensor, H: int, W: int, out: torch.Tensor | None = None) -> torch.Tensor:
    dev = nv12.device
    ctx = ctx_for(dev)
    B = nv12.shape[0]
    if out is None:
        out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    ctx.check(ctx.lib.rv_nv12_to_bgr(ctx.handle, ptr(nv12), B, H, W, ptr(out), stream_ptr(dev)))
    return out
