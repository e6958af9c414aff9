"""User-facing RGB-D -> point-cloud functions with the reference scripts' call shapes.

numpy in -> numpy-backed results, torch CUDA in -> results stay on that device (zero copy).
All arithmetic runs in the sm_100a kernels of librepasvision.so; this module only validates
shapes, moves buffers and mirrors the reference's argument meaning and error behaviour:

* depth_to_meters ............. femto_bolt_code/scripts/better_three_capture.py:118-125
* create_masked_pointcloud .... femto_bolt_code/scripts/create_masked_ply.py:56-107
* create_from_rgbd_image ...... Open3D 0.19 RGBDImage + PointCloud.create_from_rgbd_image
* PointCloud.transform / + / voxel_down_sample / select_by_index ... Open3D legacy PointCloud (Appendix B.1), call sites
  final_view_with_cad.py:333, mpa_icp_export.py:174, view_point_cloud.py:116-120
* PointCloud.select_within_distance ... realsense_d415i/capture_scripts/distance_masking_on_ply.py:9-23
* PointCloud.clip_z ........... femto_bolt_code/scripts/view_point_cloud.py:109-116
* PointCloud.crop_aabb ........ femto_bolt_code/scripts/april_tag_bg_removal_pl.py:450-468
* register_depth_to_color ..... AlignFilter / rs.align call sites (better_three_capture.py:169,188;
  capture_aligned_all.py:75,197), semantics SURVEY Appendix B.3
* get_depth_at_pixel / median_depth ... canopy_return.py:279-317, final_view.py:132-141
* fuse_views .................. SURVEY Appendix D.4 (four_pose_captures composition)
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _ops
from .calibration import Camera

_NP_DT = {"f32": np.float32, "f64": np.float64}


def _as_camera(cam=None, fx=None, fy=None, cx=None, cy=None, width=0, height=0) -> Camera:
    if isinstance(cam, Camera):
        return cam
    if isinstance(cam, dict):
        return Camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam.get("width", 0), cam.get("height", 0),
                      tuple(cam.get("dist", (0, 0, 0, 0, 0))), cam.get("model", "none"))
    return Camera(float(fx), float(fy), float(cx), float(cy), int(width), int(height))


def _is_torch_cuda(a) -> bool:
    return isinstance(a, torch.Tensor) and a.is_cuda


def _lex_order(keys: torch.Tensor) -> torch.Tensor:
    """Permutation that sorts the columns of an integer [3, m] key array lexicographically (three stable sorts)."""
    order = torch.argsort(keys[2], stable=True)
    order = order[torch.argsort(keys[1][order], stable=True)]
    return order[torch.argsort(keys[0][order], stable=True)]


class KDTreeSearchParamHybrid:
    """o3d.geometry.KDTreeSearchParamHybrid(radius, max_nn): up to max_nn nearest neighbours closer than radius."""

    def __init__(self, radius: float, max_nn: int):
        self.radius, self.max_nn = float(radius), int(max_nn)


class DeviceIndexList:
    """The `ind` of remove_statistical_outlier: Open3D hands back an IntVector, the reference script drops it
    (create_masked_ply.py:169).  It stays on the GPU until somebody looks at it; then it behaves like the int64 numpy array
    of kept indices (len, indexing, iteration, np.asarray, select_by_index)."""

    def __init__(self, index: torch.Tensor):
        self.device_tensor = index
        self._host = None

    def _array(self) -> np.ndarray:
        if self._host is None:
            self._host = self.device_tensor.cpu().numpy()
        return self._host

    def __len__(self):
        return int(self.device_tensor.numel())

    def __getitem__(self, i):
        return self._array()[i]

    def __iter__(self):
        return iter(self._array())

    def __array__(self, dtype=None, copy=None):
        a = self._array()
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return f"DeviceIndexList with {len(self)} indices"


class PointCloud:
    """Coloured cloud held on the GPU as structure-of-arrays planes x,y,z[,r,g,b].

    Mirrors the slice of Open3D's legacy PointCloud the reference uses: `.points` / `.colors`
    ((N,3) float64 numpy, like np.asarray(pcd.points)), `transform`, `+`, `voxel_down_sample`,
    `select_by_index`, `is_empty`, `has_colors`.  `.xyz` / `.rgb` expose the device planes ([3,N] torch views).
    """

    def __init__(self, data: torch.Tensor | None = None, n: int = 0, has_color: bool = True, device=None):
        if data is None:
            dev = _ops.require_cuda(device)
            data = torch.empty((6 if has_color else 3, 1), dtype=torch.float64, device=dev)
            n = 0
        self._data = data
        self._normals = None  # float64 [3, n] once estimate_normals ran; rotated by transform, averaged by voxel_down_sample,
        # carried through select_by_index, the filters, remove_statistical_outlier and + (Open3D SelectByIndex / operator+=)
        self._n = int(n)
        self._has_color = bool(has_color) and data.shape[0] >= 6

    # ---- construction
    @classmethod
    def from_arrays(cls, points, colors=None, device=None, dtype="f64") -> "PointCloud":
        dev = points.device if _is_torch_cuda(points) else _ops.require_cuda(device)
        tdt = torch.float64 if dtype == "f64" else torch.float32
        p = _ops.to_device(points, dev, tdt)
        if p.dim() != 2 or p.shape[1] != 3:
            raise ValueError("points must have shape (N, 3)")
        n = p.shape[0]
        planes = 6 if colors is not None else 3
        data = torch.empty((planes, max(n, 1)), dtype=tdt, device=dev)
        data[:3, :n] = p.t()
        if colors is not None:
            c = _ops.to_device(colors, dev, tdt)
            if c.shape != p.shape:
                raise ValueError("colors must match points in shape")
            data[3:6, :n] = c.t()
        return cls(data, n, colors is not None)

    # ---- views
    def __len__(self):
        return self._n

    @property
    def device(self):
        return self._data.device

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def xyz(self) -> torch.Tensor:
        return self._data[:3, :self._n]

    @property
    def rgb(self):
        return self._data[3:6, :self._n] if self._has_color else None

    @property
    def points(self) -> np.ndarray:
        """(N,3) float64 like np.asarray(pcd.points): one device->host copy per plane, transposed on the host."""
        return _ops.planes_to_host(self._data, 0, 3, self._n)

    @points.setter
    def points(self, value) -> None:
        """`pcd.points = o3d.utility.Vector3dVector(pts)` (create_masked_ply.py:103): a new coordinate vector.  Colours and
        normals stay only when they still have one row per point (Open3D keeps them as separate vectors; has_colors() /
        has_normals() are false there as well when the lengths differ)."""
        dt = self._data.dtype
        p = _ops.to_device(np.asarray(value) if not isinstance(value, torch.Tensor) else value, self.device, dt)
        if p.dim() != 2 or p.shape[1] != 3:
            raise ValueError("points must have shape (N, 3)")
        n = int(p.shape[0])
        keep_color = self._has_color and n == self._n and n > 0
        data = torch.empty((6 if keep_color else 3, max(n, 1)), dtype=dt, device=self.device)
        data[:3, :n] = p.t()
        if keep_color:
            data[3:6, :n] = self._data[3:6, :n]
        if self._normals is not None and n != self._n:
            self._normals = None
        self._data, self._n, self._has_color = data, n, keep_color

    @property
    def colors(self) -> np.ndarray:
        if not self._has_color:
            return np.zeros((0, 3))
        return _ops.planes_to_host(self._data, 3, 3, self._n)

    @colors.setter
    def colors(self, value) -> None:
        """`pcd.colors = o3d.utility.Vector3dVector(rgb01)` (create_masked_ply.py:104): one unit-RGB row per point; an empty
        vector removes the colours."""
        c = _ops.to_device(np.asarray(value) if not isinstance(value, torch.Tensor) else value, self.device, self._data.dtype)
        if c.numel() == 0:
            self._data, self._has_color = self._data[:3].contiguous(), False
            return
        if c.dim() != 2 or c.shape[1] != 3 or int(c.shape[0]) != self._n:
            raise ValueError("colors must have shape (N, 3) with one row per point")
        if self._data.shape[0] < 6:
            data = torch.empty((6, self._data.shape[1]), dtype=self._data.dtype, device=self.device)
            data[:3] = self._data[:3]
            self._data = data
        self._data[3:6, :self._n] = c.t()
        self._has_color = True

    def has_colors(self) -> bool:
        return self._has_color and self._n > 0

    def has_points(self) -> bool:
        return self._n > 0

    def is_empty(self) -> bool:
        return self._n == 0

    def get_min_bound(self) -> np.ndarray:
        return _ops.cloud_stats(self._data, self._n)[0:3] if self._n else np.zeros(3)

    def get_max_bound(self) -> np.ndarray:
        return _ops.cloud_stats(self._data, self._n)[3:6] if self._n else np.zeros(3)

    def __repr__(self):
        return f"PointCloud with {self._n} points."

    # ---- Open3D-shaped operations
    def transform(self, T) -> "PointCloud":
        """In place, returns self (Open3D semantics): p' = (T [p,1])[:3] / (T [p,1])[3] in float64."""
        T = np.asarray(T, dtype=np.float64)
        if T.shape != (4, 4):
            raise ValueError(f"Expected 4x4 matrix, got shape {T.shape}")
        if self._n:
            out, total, _ = _ops.transform_merge([(self._data, self._n)], [T], self._has_color)
            self._data = out
            if self._normals is not None:  # Open3D TransformNormals: n' = (T [n,0])[:3], the upper-left 3x3 only
                R = np.eye(4)
                R[:3, :3] = T[:3, :3]
                self._normals = _ops.transform_merge([(self._normals, self._n)], [R], False)[0]
        return self

    def transformed(self, T) -> "PointCloud":
        """A moved copy; this cloud is left as it is (normals are rotated along, as `transform` does)."""
        if not self._n:
            return self
        pc = PointCloud(self._data, self._n, self._has_color)
        pc._normals = self._normals
        return pc.transform(T)

    def __add__(self, other: "PointCloud") -> "PointCloud":
        return merge([self, other])

    def voxel_down_sample(self, voxel_size: float, return_keys: bool = False):
        """Open3D PointCloud.voxel_down_sample: mean position / colour per occupied voxel of the grid anchored
        at min_bound - voxel_size/2.  Output order is unspecified (as in Open3D).  RuntimeError for
        voxel_size <= 0 or a grid whose index would overflow, as Open3D raises."""
        if not (float(voxel_size) > 0.0):
            raise RuntimeError("[Open3D-compatible] voxel_size <= 0.")
        if self._n == 0:
            return (self, np.zeros((0, 3), np.int32), np.zeros(0, np.int32)) if return_keys else self
        carry = self._normals is not None
        r = _ops.voxel_downsample(self._data, self._n, self._has_color, float(voxel_size), want_keys=return_keys or carry,
                                  want_counts=return_keys)
        m = int(r["m"].item())
        if m < 0:
            raise RuntimeError("[Open3D-compatible] voxel_size is too small.")
        pc = PointCloud(r["data"], m, self._has_color)
        if carry and m:
            # Open3D averages the normals per voxel as well (not renormalised).  K4 averages whatever sits in the colour
            # planes, so a second pass runs on xyz + normals; the two outputs list the voxels in different (hash) orders
            # and are matched through their voxel keys.
            tmp = torch.empty((6, self._n), dtype=torch.float64, device=self.device)
            tmp[:3] = self._data[:3, :self._n]
            tmp[3:] = self._normals[:, :self._n]
            r2 = _ops.voxel_downsample(tmp, self._n, True, float(voxel_size), want_keys=True)
            if int(r2["m"].item()) != m:
                raise RuntimeError("voxel grids of the colour and the normal pass differ")
            oa, ob = _lex_order(r["keys"][:, :m]), _lex_order(r2["keys"][:, :m])
            nrm = torch.empty((3, m), dtype=torch.float64, device=self.device)
            nrm[:, oa] = r2["data"][3:6, :m][:, ob]
            pc._normals = nrm
        if return_keys:
            return pc, r["keys"][:, :m].t().cpu().numpy(), r["counts"][:m].cpu().numpy()
        return pc

    # ---- normals (create_masked_ply.py:172-174)
    def has_normals(self) -> bool:
        return self._normals is not None and self._n > 0

    @property
    def normals(self) -> np.ndarray:
        return self._normals[:, :self._n].t().cpu().numpy() if self._normals is not None else np.zeros((0, 3))

    @normals.setter
    def normals(self, value) -> None:
        """`pcd.normals = o3d.utility.Vector3dVector(n)`: one row per point (float64 on the device), empty = none."""
        v = _ops.to_device(np.asarray(value) if not isinstance(value, torch.Tensor) else value, self.device, torch.float64)
        if v.numel() == 0:
            self._normals = None
            return
        if v.dim() != 2 or v.shape[1] != 3 or int(v.shape[0]) != self._n:
            raise ValueError("normals must have shape (N, 3) with one row per point")
        self._normals = v.t().contiguous()

    def estimate_normals(self, search_param=None, fast_normal_computation: bool = True) -> "PointCloud":
        """Open3D PointCloud.estimate_normals with a KDTreeSearchParamHybrid(radius, max_nn) neighbourhood (the only form
        the reference uses; Open3D's default is KNN 30, given here as radius = inf).  In place, returns self."""
        sp = search_param if search_param is not None else KDTreeSearchParamHybrid(radius=float("inf"), max_nn=30)
        if self._n:
            radius = sp.radius if np.isfinite(sp.radius) else 1e300
            self._normals = _ops.estimate_normals(self._data, self._n, radius, sp.max_nn)
        return self

    def orient_normals_towards_camera_location(self, camera_location=(0.0, 0.0, 0.0)) -> "PointCloud":
        """Flip every normal that points away from camera_location (in place)."""
        if self._normals is None:
            raise RuntimeError("[Open3D-compatible] No normals in the PointCloud. Call estimate_normals() first.")
        if self._n:
            _ops.orient_normals(self._data, self._n, self._normals, camera_location)
        return self

    def remove_statistical_outlier(self, nb_neighbors: int = 20, std_ratio: float = 2.0, print_progress: bool = False):
        """Open3D PointCloud.remove_statistical_outlier (create_masked_ply.py:169): drops the points whose mean distance
        to their nb_neighbors nearest neighbours (themselves included) is not below cloud mean + std_ratio * std dev.
        Returns (cloud, ind) like Open3D: the kept points in order and their indices in this cloud (a DeviceIndexList
        where Open3D hands back an IntVector: it indexes, iterates, converts with np.asarray and feeds select_by_index
        alike, but crosses PCIe only when looked at)."""
        if int(nb_neighbors) < 1 or not (float(std_ratio) > 0.0):
            raise RuntimeError("[Open3D-compatible] Illegal input parameters, the number of neighbors and standard "
                               "deviation ratio must be positive.")
        if self._n == 0:
            return (PointCloud(None, 0, self._has_color, device=self.device),
                    DeviceIndexList(torch.zeros(0, dtype=torch.int64, device=self.device)))
        mean = _ops.knn_mean_distance(self._data, self._n, int(nb_neighbors))
        keep, _ = _ops.statistical_outlier_mask(mean, float(std_ratio))
        out, count, index = _ops.select_by_mask(self._data, self._n, self._has_color, keep)
        m = int(count.item())
        pc = PointCloud(out, m, self._has_color)
        if self._normals is not None:
            pc._normals = self._normals[:, :self._n].index_select(1, index[:m]).contiguous()
        return pc, DeviceIndexList(index[:m])

    def compute_nearest_neighbor_distance(self) -> np.ndarray:
        """Open3D PointCloud.compute_nearest_neighbor_distance (ply_to_stl.py:45,56): distance from every point to its
        nearest other point, i.e. the second entry of SearchKNN(point, 2); zeros for fewer than two points.  Runs as the
        k = 2 case of the outlier-removal search: mean(0, d) * 2 = d exactly."""
        if self._n < 2:
            return np.zeros(self._n)
        return (_ops.knn_mean_distance(self._data, self._n, 2) * 2.0).cpu().numpy()

    def translate(self, translation, relative: bool = True) -> "PointCloud":
        """Open3D PointCloud.translate: p += t (relative) or p += t - centre (absolute).  In place, returns self."""
        t = np.asarray(translation, dtype=np.float64).reshape(3)
        if not relative:
            t = t - self.get_center()
        T = np.eye(4)
        T[:3, 3] = t
        normals, self._normals = self._normals, None  # a shift leaves the normals as they are
        self.transform(T)
        self._normals = normals
        return self

    def scale(self, scale: float, center) -> "PointCloud":
        """Open3D PointCloud.scale (manual_pose_verify.py:293): p = (p - center) * scale + center, in that order (two K3
        passes so the rounding is the reference's).  In place, returns self."""
        c = np.asarray(center, dtype=np.float64).reshape(3)
        normals, self._normals = self._normals, None  # Open3D scales the points only
        if np.any(c != 0.0):
            self.translate(-c)
        T = np.diag([float(scale)] * 3 + [1.0])
        T[:3, 3] = c
        self.transform(T)
        self._normals = normals
        return self

    def rotate(self, R, center=None) -> "PointCloud":
        """Open3D PointCloud.rotate (mpa_icp.py:411, mpa_final_view_with_export.py:412 on the sampled CAD cloud):
        p = R (p - center) + center, normals n = R n; `center` defaults to the cloud's centre.  Two K3 passes, in that
        order, so every operation rounds as written.  In place, returns self."""
        R = np.asarray(R, dtype=np.float64)
        if R.shape != (3, 3):
            raise ValueError(f"Expected 3x3 matrix, got shape {R.shape}")
        c = self.get_center() if center is None else np.asarray(center, dtype=np.float64).reshape(3)
        if np.any(c != 0.0):
            self.translate(-c)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = c
        return self.transform(T)  # (rotates the normals by R as Open3D does)

    @staticmethod
    def get_rotation_matrix_from_xyz(rotation) -> np.ndarray:
        """Open3D Geometry3D.get_rotation_matrix_from_xyz: Rx(a) Ry(b) Rz(c) for rotation = (a, b, c) in radians."""
        a, b, c = (float(v) for v in np.asarray(rotation, dtype=np.float64).reshape(3))
        ca, sa, cb, sb, cc, sc = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(c), math.sin(c)
        Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
        Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
        Rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
        return Rx @ Ry @ Rz

    def get_axis_aligned_bounding_box(self) -> "AxisAlignedBoundingBox":
        """Open3D PointCloud.get_axis_aligned_bounding_box (one reduction kernel, one read-back)."""
        st = _ops.cloud_stats(self._data, self._n) if self._n else np.zeros(9)
        return AxisAlignedBoundingBox(st[0:3].copy(), st[3:6].copy())

    def crop(self, bounding_box: "AxisAlignedBoundingBox") -> "PointCloud":
        """Open3D PointCloud.crop with an axis-aligned box: the points inside it, bounds included."""
        return self.crop_aabb(bounding_box.min_bound, bounding_box.max_bound)

    def get_center(self) -> np.ndarray:
        """Mean of the points (zeros for an empty cloud)."""
        return _ops.cloud_stats(self._data, self._n)[6:9] / self._n if self._n else np.zeros(3)

    def paint_uniform_color(self, color) -> "PointCloud":
        """Open3D PointCloud.paint_uniform_color: every point gets `color` (unit RGB).  In place, returns self."""
        c = torch.as_tensor(np.asarray(color, dtype=np.float64).reshape(3), device=self.device).to(self._data.dtype)
        if not (self._has_color and self._data.shape[0] >= 6):
            data = torch.empty((6, self._data.shape[1]), dtype=self._data.dtype, device=self.device)
            data[:3] = self._data[:3]
            self._data, self._has_color = data, True
        self._data[3:6] = c[:, None]
        return self

    def select_by_index(self, indices, invert: bool = False) -> "PointCloud":
        if isinstance(indices, DeviceIndexList):
            idx = indices.device_tensor.to(self.device)
        else:
            idx = torch.as_tensor(np.asarray(indices), dtype=torch.int64, device=self.device)
        if invert:
            keep = torch.ones(self._n, dtype=torch.bool, device=self.device)
            keep[idx] = False
            idx = torch.nonzero(keep).reshape(-1)
        pc = PointCloud(self._data[:, :self._n].index_select(1, idx).contiguous(), idx.numel(), self._has_color)
        if self._normals is not None:
            pc._normals = self._normals[:, :self._n].index_select(1, idx).contiguous()
        return pc

    # ---- the reference's cloud predicates, fused into one ordered compaction kernel
    def _filtered(self, **kw) -> "PointCloud":
        if self._n == 0:
            return self
        if self._normals is None:
            out, count = _ops.filter_cloud(self._data, self._n, self._has_color, **kw)
            return PointCloud(out, int(count.item()), self._has_color)
        out, count, index = _ops.filter_cloud(self._data, self._n, self._has_color, want_index=True, **kw)
        m = int(count.item())
        pc = PointCloud(out, m, self._has_color)
        pc._normals = self._normals[:, :self._n].index_select(1, index[:m]).contiguous()
        return pc

    def select_within_distance(self, max_distance: float = 1.0) -> "PointCloud":
        """keep ||p||_2 < max_distance (strict), float64 on the stored coordinates."""
        return self._filtered(r_max=float(max_distance))

    def clip_z(self, z_min=None, z_max=None) -> "PointCloud":
        """keep z_min <= Z <= z_max, inclusive; either bound may be None."""
        if z_min is None and z_max is None:
            return self
        return self._filtered(z_clip=(z_min, z_max))

    def crop_aabb(self, min_bound, max_bound) -> "PointCloud":
        """keep min_bound <= p <= max_bound per axis, inclusive."""
        return self._filtered(aabb=(np.asarray(min_bound, dtype=np.float64), np.asarray(max_bound, dtype=np.float64)))


class AxisAlignedBoundingBox:
    """o3d.geometry.AxisAlignedBoundingBox as the scripts use it: the two corners and what follows from them."""

    def __init__(self, min_bound, max_bound):
        self.min_bound = np.asarray(min_bound, dtype=np.float64).reshape(3)
        self.max_bound = np.asarray(max_bound, dtype=np.float64).reshape(3)

    def get_min_bound(self) -> np.ndarray:
        return self.min_bound

    def get_max_bound(self) -> np.ndarray:
        return self.max_bound

    def get_center(self) -> np.ndarray:
        return (self.min_bound + self.max_bound) * 0.5

    def get_extent(self) -> np.ndarray:
        return self.max_bound - self.min_bound

    def get_max_extent(self) -> float:
        return float(self.get_extent().max())

    def volume(self) -> float:
        return float(np.prod(self.get_extent()))

    def __repr__(self):
        return f"AxisAlignedBoundingBox: min: {tuple(self.min_bound)}, max: {tuple(self.max_bound)}"


def merge(clouds) -> PointCloud:
    """Concatenation in list order (Open3D `+`)."""
    clouds = [c for c in clouds]
    if not clouds:
        raise ValueError("merge needs at least one cloud")
    has_color = all(c._has_color for c in clouds)
    eye = np.eye(4)
    views = [(c._data, c._n) for c in clouds]
    out, total, _ = _ops.transform_merge(views, [eye] * len(views), has_color)
    pc = PointCloud(out, total, has_color)
    if total and all(c._normals is not None or c._n == 0 for c in clouds):  # Open3D operator+= keeps normals when both sides have them
        pc._normals = torch.cat([c._normals[:, :c._n] for c in clouds if c._n], dim=1).contiguous()
    return pc


# --------------------------------------------------------------------------- a1/a2
def depth_to_meters(depth_raw, rule: str = "mul_f32", scale=None):
    """(raw, depth_m, scale): uint16 depth -> float32 metres, `raw.astype(float32) * 0.001` by default.

    The reference takes an SDK frame and views its buffer as uint16[H,W]; here the array is passed directly."""
    s = float(scale) if scale is not None else (0.001 if rule == "mul_f32" else 1000.0)
    if _is_torch_cuda(depth_raw):
        if depth_raw.dtype != torch.uint16:
            raise RuntimeError(f"depth must be uint16, got {depth_raw.dtype}")
        return depth_raw, _ops.depth_to_meters(depth_raw.contiguous(), rule, s), s
    raw = np.asarray(depth_raw)
    if raw.dtype != np.uint16:
        raise RuntimeError(f"depth must be uint16, got {raw.dtype}")
    dev = _ops.require_cuda()
    out = _ops.depth_to_meters(_ops.to_device(raw, dev), rule, s)
    return raw, out.cpu().numpy(), s


# --------------------------------------------------------------------------- a3
def _depth_kind(d):
    """The kernel reads uint16 raw depth or float32 metres.  The reference converts whatever `np.load` returned with
    `depth_m[valid].astype(np.float64)` (create_masked_ply.py:91): a float64 image whose values are exactly float32 numbers
    (what depth_to_meters and the capture scripts write) gives the same bits through float32 and is accepted; anything that
    would be narrowed or reinterpreted (float64 with more precision, float16, other integers) is refused instead of silently
    producing a different cloud."""
    if isinstance(d, torch.Tensor):
        if d.dtype == torch.uint16:
            return "u16", d
        if d.dtype == torch.float32:
            return "f32", d
        if d.dtype == torch.float64:
            n = d.to(torch.float32)
            if bool(((n.to(torch.float64) == d) | (d != d)).all()):
                return "f32", n
            raise RuntimeError("depth is float64 with values that are not float32 numbers: pass float32 metres or uint16 raw depth "
                               "(the kernel would narrow them and the cloud would differ from the reference's)")
        raise RuntimeError(f"depth dtype {d.dtype} is not supported: pass uint16 raw depth or float32 metres")
    if d.dtype == np.uint16:
        return "u16", d
    if d.dtype == np.float32:
        return "f32", d
    if d.dtype == np.float64:
        n = d.astype(np.float32)
        if np.array_equal(n.astype(np.float64), d, equal_nan=True):
            return "f32", n
        raise RuntimeError("depth is float64 with values that are not float32 numbers: pass float32 metres or uint16 raw depth "
                           "(the kernel would narrow them and the cloud would differ from the reference's)")
    raise RuntimeError(f"depth dtype {d.dtype} is not supported: pass uint16 raw depth or float32 metres")


def _prep_frame_inputs(rgb, depth, mask, dev):
    """Validate one frame's arrays the way create_masked_ply.py:134-139 does and upload them as [1,...] tensors."""
    d = depth if isinstance(depth, torch.Tensor) else np.asarray(depth)
    if d.ndim != 2:
        raise RuntimeError(f"depth must be HxW, got shape {tuple(d.shape)}")
    H, W = d.shape
    if rgb is not None:
        c = rgb if isinstance(rgb, torch.Tensor) else np.asarray(rgb)
        if tuple(c.shape) != (H, W, 3):
            raise RuntimeError(f"Color/depth size mismatch: color {tuple(c.shape[:2])}, depth {(H, W)}")
    if mask is not None:
        m = mask if isinstance(mask, torch.Tensor) else np.asarray(mask)
        if tuple(m.shape) != (H, W):
            raise RuntimeError(f"Mask/depth size mismatch: mask {tuple(m.shape)}, depth {(H, W)}")
    kind, d = _depth_kind(d)
    dt = _ops.to_device(d, dev)
    ct = None if rgb is None else _ops.to_device(rgb, dev, torch.uint8)
    mt = None if mask is None else _ops.to_device(mask, dev, torch.uint8)
    return (dt[None], None if ct is None else ct[None], None if mt is None else mt[None], kind, H, W)


def create_masked_pointcloud(rgb, depth_m, mask, fx, fy, cx, cy, invert_mask: bool = False, *, dtype: str = "f64",
                             unit_rule: str = "mul_f32", depth_scale=None, max_distance=None, z_clip=None, aabb=None,
                             camera: Camera | None = None) -> PointCloud:
    """Coloured cloud of the masked pixels with finite positive depth, in row-major pixel order.

    rgb: HxWx3 uint8 BGR; depth_m: HxW float metres aligned to colour (a uint16 raw-depth image is also
    accepted and converted with `unit_rule` / `depth_scale`); mask: HxW uint8 (>0 = object, or ==0 when
    invert_mask).  x=(u-cx)*z/fx, y=(v-cy)*z/fy in float64; colours RGB/255.  dtype="f64" reproduces the
    reference's float64 arrays bit for bit; "f32" stores float32 planes (within 1 ulp of the float64 values).
    The keyword-only predicates fuse the reference's later filters into the same kernel."""
    dev = depth_m.device if _is_torch_cuda(depth_m) else _ops.require_cuda()
    d, c, m, kind, H, W = _prep_frame_inputs(rgb, depth_m, mask, dev)
    cam = camera if camera is not None else Camera(float(fx), float(fy), float(cx), float(cy), W, H)
    r = _ops.deproject(d, c, m, cam, depth_kind=kind, unit_rule=unit_rule, unit_scale=depth_scale,
                       invert_mask=invert_mask, r_max=max_distance, z_clip=z_clip, aabb=aabb, mode="compact_ordered",
                       out_dtype=dtype)
    return PointCloud(r["data"], int(r["counts"][0].item()), c is not None)


def create_from_rgbd_image(color, depth, intrinsic, extrinsic=None, *, depth_scale: float = 1000.0,
                           depth_trunc: float = 3.0, project_valid_depth_only: bool = True, dtype: str = "f64"):
    """Open3D-shaped: RGBDImage.create_from_color_and_depth(depth_scale, depth_trunc, convert_rgb_to_intensity=False)
    followed by PointCloud.create_from_rgbd_image(intrinsic, extrinsic).  color is BGR uint8 (cv2 order).
    With project_valid_depth_only=False the cloud is dense (H*W points, NaN where invalid)."""
    cam = _as_camera(intrinsic)
    dev = depth.device if _is_torch_cuda(depth) else _ops.require_cuda()
    d, c, _, kind, H, W = _prep_frame_inputs(color, depth, None, dev)
    if kind != "u16":
        raise RuntimeError("create_from_rgbd_image expects a uint16 depth image")
    mode = "compact_ordered" if project_valid_depth_only else "dense_nan"
    r = _ops.deproject(d, c, None, cam, depth_kind="u16", unit_rule="div_f32", unit_scale=float(depth_scale),
                       depth_trunc=float(depth_trunc), mode=mode, out_dtype=dtype)
    n = int(r["counts"][0].item()) if project_valid_depth_only else H * W
    pc = PointCloud(r["data"], n, c is not None)
    if extrinsic is not None:
        pc.transform(np.linalg.inv(np.asarray(extrinsic, dtype=np.float64)))
    return pc


class CloudBatch:
    """Result of a batched deprojection: planes [6 or 3, B*cap] with frame b at columns [b*cap, b*cap+counts[b])."""

    def __init__(self, r: dict, B: int, H: int, W: int, has_color: bool, dense: bool, packed_color: bool = False):
        self.data, self.counts, self.cap = r["data"], r["counts"], r["cap"]
        self.valid, self.src_index = r["valid"], r["src_index"]
        self.B, self.H, self.W, self.has_color, self.dense = B, H, W, has_color, dense
        self.packed_color = packed_color and has_color  # plane 3 holds the bytes r,g,b,0 (color_scale="packed8")
        self._counts_host = None

    def counts_host(self) -> np.ndarray:
        if self._counts_host is None:
            self._counts_host = self.counts.cpu().numpy()
        return self._counts_host

    def _n(self, b: int) -> int:
        return min(self.H * self.W if self.dense else int(self.counts_host()[b]), self.cap)

    def frame(self, b: int) -> PointCloud:
        if self.packed_color:
            raise RuntimeError("this batch carries packed 8-bit colours (color_scale='packed8'): use .rgb8(b) / .xyz(b), or "
                               "deproject with color_scale='unit' for PointCloud frames")
        n = self._n(b)
        return PointCloud(self.data[:, b * self.cap:b * self.cap + max(n, 1)], n, self.has_color)

    def xyz(self, b: int) -> torch.Tensor:
        """[3, n] device view of frame b's coordinates."""
        return self.data[:3, b * self.cap:b * self.cap + self._n(b)]

    def rgb8(self, b: int) -> torch.Tensor:
        """[n, 3] uint8 device view (r, g, b) of frame b's colours; packed-colour batches only."""
        if not self.packed_color:
            raise RuntimeError("rgb8 needs a batch made with color_scale='packed8'")
        n = self._n(b)
        return self.data[3, b * self.cap:b * self.cap + n].view(torch.uint8).view(n, 4)[:, :3]

    def __len__(self):
        return self.B


def deproject_batch(depth, bgr, camera, mask=None, *, unit_rule: str = "mul_f32", depth_scale=None, invert_mask=False,
                    depth_trunc=None, max_distance=None, z_clip=None, aabb=None, mode: str = "compact_ordered",
                    dtype: str = "f32", color_scale: str = "unit", want_valid=False, want_src_index=False,
                    frame_capacity=None, out=None, kernel: str = "auto", color_format: str = "bgr",
                    geometry: str = "reference") -> CloudBatch:
    """Batched form of create_masked_pointcloud: depth [B,H,W] (uint16 raw or float32 metres), bgr [B,H,W,3] uint8,
    mask [B,H,W] uint8 or None.  One kernel launch for the whole batch.

    color_format="nv12": `bgr` holds the camera's NV12 frames [B, H*3/2, W] (better_three_capture.py:101-106,159); the
    kernel converts the kept pixels itself, identical to feeding cv2.cvtColor(..., COLOR_YUV2BGR_NV12) images.
    color_scale="packed8": colours stay bytes (one r,g,b,0 word per point in a fourth plane) instead of three float planes.
    geometry="sdk_f32": x = z * ((u - ppx) / fx) in float32 like rs.pointcloud / PointCloudFilter (capture_aligned_all.py:209-216,
    better_three_capture.py:235-237) instead of the reference's float64 numpy form; with mode="dense_zero" and
    color_scale="255" this is the SDKs' cloud, value for value."""
    cam = _as_camera(camera)
    dev = depth.device if _is_torch_cuda(depth) else _ops.require_cuda()
    kind, d = _depth_kind(depth if isinstance(depth, torch.Tensor) else np.asarray(depth))
    d = _ops.to_device(d, dev)
    if d.dim() != 3:
        raise RuntimeError(f"depth must be [B,H,W], got {tuple(d.shape)}")
    B, H, W = d.shape
    c = None if bgr is None else _ops.to_device(bgr, dev, torch.uint8)
    if color_format not in ("bgr", "nv12"):
        raise ValueError("color_format must be 'bgr' or 'nv12'")
    want = (B, H, W, 3) if color_format == "bgr" else (B, H * 3 // 2, W)
    if c is not None and (tuple(c.shape) != want or (color_format == "nv12" and (H % 2 or W % 2))):
        raise RuntimeError(f"Color/depth size mismatch: color {tuple(c.shape)}, depth {(B, H, W)}")
    m = None if mask is None else _ops.to_device(mask, dev, torch.uint8)
    if m is not None and tuple(m.shape) != (B, H, W):
        raise RuntimeError(f"Mask/depth size mismatch: mask {tuple(m.shape)}, depth {(B, H, W)}")
    r = _ops.deproject(d, c, m, cam, depth_kind=kind, unit_rule=unit_rule, unit_scale=depth_scale, invert_mask=invert_mask,
                       depth_trunc=depth_trunc, r_max=max_distance, z_clip=z_clip, aabb=aabb, mode=mode, out_dtype=dtype,
                       color_scale=color_scale, want_valid=want_valid, want_src_index=want_src_index,
                       frame_capacity=frame_capacity, out=out, kernel=kernel, color_format=color_format, geometry=geometry)
    return CloudBatch(r, B, H, W, c is not None, mode.startswith("dense"), color_scale == "packed8")


# --------------------------------------------------------------------------- a6
def register_depth_to_color(depth, depth_camera, color_camera, R=None, t=None, *, depth_units: float = 0.001,
                            rotation_layout: str = "row_major", return_winner: bool = False):
    """Depth image(s) in the depth camera's geometry -> uint16 depth on the colour pixel grid, nearest surface wins
    (what AlignFilter(COLOR_STREAM).process / rs.align(rs.stream.color).process return for the depth stream).

    depth: [Hd,Wd] or [B,Hd,Wd] uint16.  R, t: depth -> colour rigid transform; R as a 3x3 in `rotation_layout`
    ("row_major" mathematical matrix, or "col_major" = librealsense's flat storage).  Identity when omitted."""
    dcam, ccam = _as_camera(depth_camera), _as_camera(color_camera)
    single = (depth.dim() if isinstance(depth, torch.Tensor) else np.asarray(depth).ndim) == 2
    dev = depth.device if _is_torch_cuda(depth) else _ops.require_cuda()
    d = _ops.to_device(depth, dev)
    if d.dtype != torch.uint16:
        raise RuntimeError(f"depth must be uint16, got {d.dtype}")
    if single:
        d = d[None]
    if ccam.width <= 0 or ccam.height <= 0:
        raise RuntimeError("color_camera needs width/height")
    Rm = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
    Rcol = Rm.T.reshape(9) if rotation_layout == "row_major" else Rm.reshape(9)
    tt = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).reshape(3)
    out, win = _ops.register(d, dcam, ccam, Rcol, tt, depth_units, return_winner)
    numpy_out = not _is_torch_cuda(depth)
    if single:
        out = out[0]
        win = None if win is None else win[0]
    if numpy_out:
        out = out.cpu().numpy()
        win = None if win is None else win.cpu().numpy()
    return (out, win) if return_winner else out


# --------------------------------------------------------------------------- a2 (windowed median)
def median_depth_windows(depth_raw, pixels, window: int = 5) -> np.ndarray:
    """Median of the non-zero raw depths in clipped window x window neighbourhoods (raw units, float64; NaN = none)."""
    dev = depth_raw.device if _is_torch_cuda(depth_raw) else _ops.require_cuda()
    d = _ops.to_device(depth_raw, dev)
    if d.dtype != torch.uint16 or d.dim() != 2:
        raise RuntimeError("depth must be a uint16 HxW image")
    uv = _ops.to_device(np.asarray(pixels, dtype=np.int32).reshape(-1, 2), dev)
    return _ops.median_depth_window(d, uv, window).cpu().numpy()


def get_depth_at_pixel(depth_raw, x: int, y: int, window_size: int = 5):
    """canopy_return.py:279-317: median of valid depths around (x, y) in metres (raw / 1000.0), None if no valid depth."""
    v = float(median_depth_windows(depth_raw, [(int(x), int(y))], window_size)[0])
    if np.isnan(v):
        return None
    return v / 1000.0


def deproject_pixel_to_point(intrinsics, pixel, depth_value):
    """canopy_return.py:183-206 single pixel, float64 on the host (three flops; not worth a launch)."""
    cam = intrinsics
    fx, fy = float(cam.fx), float(cam.fy)
    ppx = float(getattr(cam, "ppx", getattr(cam, "cx", 0.0)))
    ppy = float(getattr(cam, "ppy", getattr(cam, "cy", 0.0)))
    x, y = pixel
    return ((x - ppx) * depth_value / fx, (y - ppy) * depth_value / fy, depth_value)


# --------------------------------------------------------------------------- config 4
def fuse_views(clouds, T_cam_tag_list, voxel_size: float = 0.005, *, return_keys: bool = False):
    """Four-pose fusion: move view i by inv(T_cam_tag_i) into the shared tag frame, concatenate in view order,
    voxel_down_sample(voxel_size).  Returns the fused PointCloud (and voxel keys / counts when asked)."""
    from .pose import world_from_camera
    if len(clouds) != len(T_cam_tag_list) or not clouds:
        raise ValueError("need one pose per cloud")
    has_color = all(c._has_color for c in clouds)
    Ts = [world_from_camera(T) for T in T_cam_tag_list]
    if not (float(voxel_size) > 0.0):
        raise RuntimeError("[Open3D-compatible] voxel_size <= 0.")
    total = sum(c._n for c in clouds)
    if total == 0:
        empty = PointCloud(None, 0, has_color, device=clouds[0].device)
        return (empty, np.zeros((0, 3), np.int32), np.zeros(0, np.int32)) if return_keys else empty
    if len(clouds) > 8:  # the fused call takes eight views; more go through the merged cloud
        out, total, bounds = _ops.transform_merge([(c._data, c._n) for c in clouds], Ts, has_color, want_bounds=True)
        r = _ops.voxel_downsample(out, total, has_color, float(voxel_size), bounds=bounds, want_keys=return_keys,
                                  want_counts=return_keys)
    else:
        r = _ops.fuse_voxel([(c._data, c._n) for c in clouds], Ts, has_color, float(voxel_size), want_keys=return_keys,
                            want_counts=return_keys)
    m = int(r["m"].item())
    if m < 0:
        raise RuntimeError("[Open3D-compatible] voxel_size is too small.")
    pc = PointCloud(r["data"], m, has_color)
    if return_keys:
        return pc, r["keys"][:, :m].t().cpu().numpy(), r["counts"][:m].cpu().numpy()
    return pc


def nv12_to_bgr(nv12, height: int, width: int):
    """cv2.cvtColor(nv12.reshape(H*3//2, W), COLOR_YUV2BGR_NV12) for [B, H*3/2, W] (or one [H*3/2, W]) uint8 frames."""
    dev = nv12.device if _is_torch_cuda(nv12) else _ops.require_cuda()
    a = _ops.to_device(nv12, dev, torch.uint8)
    single = a.dim() == 2
    if single:
        a = a[None]
    out = _ops.nv12_to_bgr(a, int(height), int(width))
    if single:
        out = out[0]
    return out if _is_torch_cuda(nv12) else out.cpu().numpy()
