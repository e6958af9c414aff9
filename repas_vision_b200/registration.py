"""ICP registration in the shape of o3d.pipelines.registration.

Reference call sites: `refine_with_icp` femto_bolt_code/scripts/mpa_icp_export.py:166-208, 6dof_icp_export.py:109-157,
mpa_icp.py:150-170, `align_postop_to_preop` icp_cad_model.py:62-96 and :266-275 -- all
`registration_icp(source, target, max_dist, init, TransformationEstimationPointToPlane(), ICPConvergenceCriteria(...))`.

Semantics restated from Open3D 0.19 (pipelines/registration/Registration.cpp, TransformationEstimation.cpp,
utility/Eigen.cpp; the wheel is not part of the reference checkout, see DESIGN.md "parity unpinned" note):

  registration_icp:  pcd = source transformed by init; correspondences; then up to max_iteration times
      update = estimation.compute_transformation(pcd, target, correspondences)
      transformation = update @ transformation;  pcd.transform(update);  new correspondences
      stop when |fitness - previous| < relative_fitness and |inlier_rmse - previous| < relative_rmse
  correspondences:  per source point the nearest target point with squared distance < max_dist^2;
      fitness = matched / len(source), inlier_rmse = sqrt(sum of squared distances / matched)
  point-to-plane:   r = (s - t) . n_t, J = [s x n_t, n_t];  x = solve(J^T J, -J^T r);
      update = [Rz(x2) Ry(x1) Rx(x0) | x3..5]
  point-to-point:   Eigen::umeyama(source, target, with_scaling)

Point-to-plane (what every call site of the reference uses) runs the whole loop on the device: per iteration four kernels
(move the working copy, nearest-point search on the target's hash grid, the sums of the estimation step, and one that takes
the loop's decisions: fitness / RMSE, the stopping rule, the 6x6 solve, the pose composition); the host queues a few
iterations at a time and reads 448 bytes of state back per batch.  Point-to-point (Umeyama's SVD) and
`device_loop=False` keep the solve and the stopping rule here on the host: three kernels and 32 doubles per iteration.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _ops
from .cloud import PointCloud


class ICPConvergenceCriteria:
    def __init__(self, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6, max_iteration: int = 30):
        self.relative_fitness = float(relative_fitness)
        self.relative_rmse = float(relative_rmse)
        self.max_iteration = int(max_iteration)

    def __repr__(self):
        return (f"ICPConvergenceCriteria class with relative_fitness={self.relative_fitness:e}, "
                f"relative_rmse={self.relative_rmse:e}, and max_iteration={self.max_iteration}")


class TransformationEstimationPointToPoint:
    def __init__(self, with_scaling: bool = False):
        self.with_scaling = bool(with_scaling)

    point_to_plane = False

    def update_from_sums(self, s: np.ndarray) -> np.ndarray:
        """Eigen::umeyama from the correspondence sums: s[0] count, s[2] sum |src|^2, s[3:6] sum src, s[6:9] sum tgt,
        s[9:18] sum tgt_a src_b."""
        n = s[0]
        if n == 0:
            return np.eye(4)
        src_mean, dst_mean = s[3:6] / n, s[6:9] / n
        sigma = s[9:18].reshape(3, 3) / n - np.outer(dst_mean, src_mean)
        U, d, Vt = np.linalg.svd(sigma)
        S = np.ones(3)
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            S[2] = -1.0
        R = U @ np.diag(S) @ Vt
        T = np.eye(4)
        if self.with_scaling:
            src_var = s[2] / n - float(src_mean @ src_mean)
            c = float(d @ S) / src_var
            T[:3, 3] = dst_mean - c * (R @ src_mean)
            T[:3, :3] = c * R
        else:
            T[:3, 3] = dst_mean - R @ src_mean
            T[:3, :3] = R
        return T


class TransformationEstimationPointToPlane:
    point_to_plane = True

    def update_from_sums(self, s: np.ndarray) -> np.ndarray:
        """utility::SolveJacobianSystemAndObtainExtrinsicMatrix: x = solve(JTJ, -JTr), then TransformVector6dToMatrix4d."""
        if s[0] == 0:
            return np.eye(4)
        JTJ = np.zeros((6, 6))
        JTJ[np.triu_indices(6)] = s[9:30]
        JTJ = JTJ + np.triu(JTJ, 1).T
        JTr = s[3:9]
        try:
            x = np.linalg.solve(JTJ, -JTr)
        except np.linalg.LinAlgError:  # exactly singular (Eigen's LDLT would still hand back a solution of the consistent part)
            x = np.linalg.lstsq(JTJ, -JTr, rcond=None)[0]
        if not np.all(np.isfinite(x)):
            return np.eye(4)
        return vector6d_to_matrix4d(x)


def vector6d_to_matrix4d(x) -> np.ndarray:
    """utility::TransformVector6dToMatrix4d: rotation Rz(x[2]) @ Ry(x[1]) @ Rx(x[0]), translation x[3:6]."""
    a, b, c = float(x[0]), float(x[1]), float(x[2])
    ca, sa, cb, sb, cc, sc = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(c), math.sin(c)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]], dtype=np.float64)
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]], dtype=np.float64)
    Rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]], dtype=np.float64)
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = np.asarray(x[3:6], dtype=np.float64)
    return T


def _is_identity(T: np.ndarray, prec: float = 1e-12) -> bool:
    """Eigen's Matrix4d::isIdentity(): diagonal approximately 1, everything else much smaller than 1."""
    for i in range(4):
        for j in range(4):
            v = float(T[i, j])
            if i == j:
                if not abs(v - 1.0) <= min(abs(v), 1.0) * prec:
                    return False
            elif not abs(v) <= prec:
                return False
    return True


class RegistrationResult:
    """transformation (4x4 float64), fitness, inlier_rmse, correspondence_set ((M,2) int32: source index, target index,
    ascending source index -- Open3D's order depends on its thread schedule)."""

    def __init__(self, transformation=None):
        self.transformation = np.eye(4) if transformation is None else np.array(transformation, dtype=np.float64)
        self.fitness = 0.0
        self.inlier_rmse = 0.0
        self._nearest = None  # int32 [n_source] device tensor, -1 = unmatched
        self._n = 0
        self.iterations = 0  # estimation steps registration_icp ran (not an Open3D field)

    @property
    def correspondence_set(self) -> np.ndarray:
        if self._nearest is None or self._n == 0:
            return np.zeros((0, 2), np.int32)
        near = self._nearest[:self._n]
        src = torch.nonzero(near >= 0).reshape(-1)
        return torch.stack([src.to(torch.int32), near[src]], dim=1).cpu().numpy()

    def __repr__(self):
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, and "
                f"correspondence_set size of {int(round(self.fitness * self._n))}\nAccess transformation to get result.")


class _Matcher:
    """One registration: the target's hash-grid index (the reference's KDTreeFlann), the working copy of the source
    coordinates (two buffers, K3 goes from one into the other) and the per-step buffers."""

    def __init__(self, source: PointCloud, target: PointCloud, max_distance: float, point_to_plane: bool):
        if source.device != target.device:
            raise ValueError("source and target must live on the same device")
        self.target, self.max_distance, self.plane = target, float(max_distance), bool(point_to_plane)
        self.n = len(source)
        self.index = _ops.nn_index_build(target._data, len(target), self.max_distance) if len(target) else None
        self.cur = source._data  # read-only until the first transform, which writes into a buffer of its own
        self.spare = [torch.empty((3, max(self.n, 1)), dtype=source._data.dtype, device=source.device) for _ in range(2)]
        self.nearest = torch.empty(max(self.n, 1), dtype=torch.int32, device=source.device)
        self.scratch = _ops.icp_scratch(source.device)

    def transform(self, T: np.ndarray) -> None:
        """pcd.Transform(T) on the working copy."""
        if self.n:
            out = self.spare[0] if self.cur is not self.spare[0] else self.spare[1]
            _ops.transform_xyz_into(self.cur, self.n, T, out)
            self.cur = out

    def evaluate(self, transformation: np.ndarray):
        """GetRegistrationResultAndCorrespondences; also returns the sums the estimation step needs."""
        res = RegistrationResult(transformation)
        n, nt = self.n, len(self.target)
        res._n = n
        if n == 0 or nt == 0:
            return res, np.zeros(32)
        _ops.nn_search(self.index, nt, self.cur, n, self.max_distance, out=self.nearest)
        normals = self.target._normals if self.plane else None
        s = _ops.icp_sums(self.cur, n, self.target._data, nt, normals, self.nearest, self.plane, self.scratch).cpu().numpy()
        res._nearest = self.nearest
        if s[0] > 0:
            res.fitness = float(s[0]) / float(n)
            res.inlier_rmse = math.sqrt(float(s[1]) / float(s[0]))
        return res, s


def evaluate_registration(source: PointCloud, target: PointCloud, max_correspondence_distance: float,
                          transformation=None) -> RegistrationResult:
    """o3d.pipelines.registration.evaluate_registration."""
    T = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
    if not (float(max_correspondence_distance) > 0.0):
        return RegistrationResult(T)
    m = _Matcher(source, target, max_correspondence_distance, False)
    if not _is_identity(T):
        m.transform(T)
    res, _ = m.evaluate(T)
    res._nearest = None if res._nearest is None else res._nearest.clone()
    return res


ICP_STEPS_PER_READBACK = 8  # iterations queued between two looks at the device-side state (2 / 4 / 8 / 16: 1.17 / 1.15 / 1.11 / 1.19 ms on the 349 k-point case)


def _icp_on_device(m: _Matcher, source: PointCloud, transformation: np.ndarray, crit: ICPConvergenceCriteria) -> RegistrationResult:
    """The point-to-plane loop with its decisions taken on the device (rv_icp_iterate)."""
    n, nt = m.n, len(m.target)
    dev = source.device
    work = m.spare[0]
    _ops.transform_xyz_into(source._data, n, transformation, work)  # pcd = source transformed by init (a copy, always)
    state = _ops.icp_state(dev)
    _ops.icp_begin(state, transformation, crit.max_iteration, crit.relative_fitness, crit.relative_rmse, n)
    queued, first = 0, True
    while True:
        steps = min(ICP_STEPS_PER_READBACK, crit.max_iteration - queued)
        _ops.icp_iterate(state, first, steps, work, n, m.index, m.target._data, nt, m.target._normals, m.max_distance, m.nearest,
                         m.scratch)
        queued += steps
        first = False
        host = state.cpu().numpy()
        done = int(host[48:49].view(np.int32)[0])
        if done or queued >= crit.max_iteration:
            break
    res = RegistrationResult(host[:16].reshape(4, 4).copy())
    res._n, res._nearest = n, m.nearest
    res.fitness, res.inlier_rmse, res.iterations = float(host[32]), float(host[33]), int(host[36])
    return res


def registration_icp(source: PointCloud, target: PointCloud, max_correspondence_distance: float, init=None,
                     estimation_method=None, criteria: ICPConvergenceCriteria | None = None, *,
                     device_loop: bool = True) -> RegistrationResult:
    """o3d.pipelines.registration.registration_icp (legacy pipeline), same argument order and defaults.  device_loop=False
    keeps the solve and the stopping rule on the host for point-to-plane as well (one read-back per iteration)."""
    est = estimation_method if estimation_method is not None else TransformationEstimationPointToPoint()
    crit = criteria if criteria is not None else ICPConvergenceCriteria()
    if not (float(max_correspondence_distance) > 0.0):
        raise RuntimeError("[Open3D-compatible] Invalid max_correspondence_distance.")
    if est.point_to_plane and not target.has_normals():
        raise RuntimeError("[Open3D-compatible] TransformationEstimationPointToPlane and TransformationEstimationColoredICP "
                           "require pre-computed normal vectors for target PointCloud.")
    transformation = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    if transformation.shape != (4, 4):
        raise ValueError(f"Expected 4x4 matrix, got shape {transformation.shape}")
    m = _Matcher(source, target, max_correspondence_distance, est.point_to_plane)
    if (device_loop and est.point_to_plane and type(est) is TransformationEstimationPointToPlane and m.n > 0 and len(target) > 0):
        return _icp_on_device(m, source, transformation, crit)
    if not _is_identity(transformation):
        m.transform(transformation)
    result, sums = m.evaluate(transformation)
    for it in range(1, crit.max_iteration + 1):
        update = est.update_from_sums(sums)
        transformation = update @ transformation
        m.transform(update)
        backup = result
        result, sums = m.evaluate(transformation)
        result.iterations = it
        if (abs(backup.fitness - result.fitness) < crit.relative_fitness
                and abs(backup.inlier_rmse - result.inlier_rmse) < crit.relative_rmse):
            break
    return result
