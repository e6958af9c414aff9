"""The slice of the `open3d` namespace that the reference's point-cloud path touches, backed by this package.

    from repas_vision_b200 import o3d_compat as o3d      # instead of: import open3d as o3d

    pcd = o3d.geometry.PointCloud()                                   # create_masked_ply.py:102-104
    pcd.points = o3d.utility.Vector3dVector(points)
    pcd.colors = o3d.utility.Vector3dVector(colors)
    pcd = pcd.voxel_down_sample(0.005)                                # mpa_icp_export.py:174
    pcd.estimate_normals(search_param=o3d.geometry.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    reg = o3d.pipelines.registration.registration_icp(src, pcd, 0.02, np.eye(4),
            o3d.pipelines.registration.TransformationEstimationPointToPlane(),
            o3d.pipelines.registration.ICPConvergenceCriteria(max_iteration=50))      # mpa_icp_export.py:187-197
    o3d.io.write_point_cloud(str(path), pcd, write_ascii=False, compressed=True)      # mpa_icp_export.py

Everything here runs on the GPU through the C ABI (`np.asarray(pcd.points)` is a device-to-host copy, not a view: writes into
it do not reach the cloud -- assign `pcd.points = ...` instead; the reference never writes through the view).  What is not
the point-cloud path -- triangle meshes, line sets, the visualiser, FPFH / RANSAC global registration, raycasting -- is not
here, and asking for it says so.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from . import cloud as _cloud
from . import ply as _ply
from . import registration as _registration


def _rows(width: int, dtype):
    def make(values=()):
        a = np.asarray(values, dtype=dtype)
        if a.size == 0:
            return a.reshape(0, width)
        if a.ndim != 2 or a.shape[1] != width:
            raise RuntimeError(f"expected an array of shape (N, {width}), got {a.shape}")  # Open3D raises RuntimeError too
        return np.ascontiguousarray(a)
    return make


def _vector(dtype):
    return lambda values=(): np.ascontiguousarray(np.asarray(values, dtype=dtype).reshape(-1))


class _Scope(SimpleNamespace):
    """A namespace that names what it leaves out instead of failing with a bare AttributeError."""

    def __getattr__(self, name):
        raise AttributeError(f"open3d.{self._path}.{name} is outside the point-cloud path repas_vision_b200 replaces "
                             f"(available here: {', '.join(k for k in vars(self) if not k.startswith('_'))})")


utility = _Scope(_path="utility", Vector3dVector=_rows(3, np.float64), Vector3iVector=_rows(3, np.int32),
                 Vector2iVector=_rows(2, np.int32), DoubleVector=_vector(np.float64), IntVector=_vector(np.int32))

geometry = _Scope(_path="geometry", PointCloud=_cloud.PointCloud, KDTreeSearchParamHybrid=_cloud.KDTreeSearchParamHybrid,
                  AxisAlignedBoundingBox=_cloud.AxisAlignedBoundingBox,
                  get_rotation_matrix_from_xyz=_cloud.PointCloud.get_rotation_matrix_from_xyz)

io = _Scope(_path="io", read_point_cloud=_ply.read_point_cloud, write_point_cloud=_ply.write_point_cloud)

pipelines = _Scope(_path="pipelines", registration=_Scope(
    _path="pipelines.registration", registration_icp=_registration.registration_icp,
    evaluate_registration=_registration.evaluate_registration, RegistrationResult=_registration.RegistrationResult,
    ICPConvergenceCriteria=_registration.ICPConvergenceCriteria,
    TransformationEstimationPointToPlane=_registration.TransformationEstimationPointToPlane,
    TransformationEstimationPointToPoint=_registration.TransformationEstimationPointToPoint))


def __getattr__(name):
    raise AttributeError(f"open3d.{name} is outside the point-cloud path repas_vision_b200 replaces "
                         "(available here: geometry, utility, io, pipelines.registration)")
