"""Cross-checks against the third-party code the reference calls (open3d 0.19, pyrealsense2, OpenCV).

OpenCV is in the image, so the distortion ray table and the RealSense deprojection formula are pinned here on the CPU.
Open3D and pyrealsense2 are not installable in the build container (no network); every Open3D-defined row of the hot path
(voxel_down_sample, transform, remove_statistical_outlier, estimate_normals, registration_icp, PLY) has a test below that runs
as soon as `import open3d` works -- on a pod that has it, or from `baseline/_ref/` -- and is skipped otherwise.  When they
run, DESIGN.md's "parity unpinned" list shrinks to what they do not cover."""
import os
import sys

import numpy as np
import pytest

from conftest import CAL, CANOPY_TS, load_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)


# ------------------------------------------------------------------ pinned here: OpenCV
def test_brown_conrady_ray_table_equals_cv2_undistort():
    """The checkerboard calibration of the reference (april_tag_detector_solvepnp.py:51-66 loads it for cv2) carries
    cv2.calibrateCamera coefficients, which `load_camera` maps to the forward Brown-Conrady model.  The 10-step fixed-point
    inverse of the ray table (librealsense's rs2_deproject_pixel_to_point) must land on OpenCV's converged undistortion of the
    same pixels, and projecting the rays back with cv2.projectPoints must return the pixel centres."""
    import cv2
    from oracle import oracle_np as O
    from repas_vision_b200.calibration import load_camera
    cam = load_camera(os.path.join(CAL, "checkerboard_color_intrinsics_2025-08-26T183535.json"))
    assert cam.model == "brown_conrady" and any(cam.dist)
    W, H = cam.width, cam.height
    rays = O.ray_table(cam.fx, cam.fy, cam.cx, cam.cy, cam.dist, cam.model, W, H)
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
    dist = np.array(cam.dist)
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    pix = np.stack([u, v], -1).reshape(-1, 1, 2)
    conv = cv2.undistortPointsIter(pix, K, dist, None, None, (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 100, 1e-15))
    assert np.abs(rays - conv.reshape(H, W, 2)).max() <= 1e-14          # measured 5.6e-16
    five = cv2.undistortPoints(pix, K, dist).reshape(H, W, 2)            # OpenCV's default: five iterations
    assert np.abs(rays - five).max() <= 1e-9                             # measured 2.2e-10: ours is the more converged one
    pts = np.concatenate([rays.reshape(-1, 2), np.ones((H * W, 1))], 1)
    back, _ = cv2.projectPoints(pts, np.zeros(3), np.zeros(3), K, dist)
    assert np.abs(back.reshape(H, W, 2) - np.stack([u, v], -1)).max() <= 1e-10  # pixels; measured 4.5e-13


def test_dense_cloud_against_float32_sdk_formula(rs720):
    """rs.pointcloud / PointCloudFilter compute x = depth * ((u - ppx) / fx) in float32; this repo computes
    (u - cx) * z / fx in float64 like the reference's numpy path and rounds once to float32.  On the reference's captured
    frames the two differ by at most ONE float32 unit in the last place (measured: exactly 1.0 on all five frames), which is
    below 1e-6 m for everything within the 8 m the depth cameras range over and 3.8e-6 m on saturated 65.535 m pixels --
    inside the 1e-5 m bar everywhere."""
    from oracle import oracle_np as O
    cam = dict(fx=rs720["fx"], fy=rs720["fy"], cx=rs720["cx"], cy=rs720["cy"], model="none")
    worst_ulp, worst_abs, worst_abs_8m = 0.0, 0.0, 0.0
    for ts in CANOPY_TS:
        color, depth = load_frame(ts)
        h, w = depth.shape
        ours = O.deproject_mask(depth, color, None, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], out_dtype="f32")
        keep = ours["valid"]
        v, u = np.nonzero(keep)
        z = depth[keep].astype(np.float32) * np.float32(0.001)
        X, Y, Z = O._rs_deproject_f32(u.astype(np.float32), v.astype(np.float32), z, cam)
        sdk = np.stack([X, Y, Z], axis=1)
        assert sdk.dtype == np.float32 and np.array_equal(sdk[:, 2], ours["points"][:, 2])  # z is the same float32 product
        d = np.abs(sdk.astype(np.float64) - ours["points"].astype(np.float64))
        ulp = np.spacing(np.abs(ours["points"])).astype(np.float64)
        worst_ulp = max(worst_ulp, float((d / np.maximum(ulp, 1e-30)).max()))
        worst_abs = max(worst_abs, float(d.max()))
        worst_abs_8m = max(worst_abs_8m, float(d[z <= 8.0].max(initial=0.0)))
    assert worst_ulp <= 1.0, worst_ulp        # float32 units in the last place
    assert worst_abs_8m <= 1e-6, worst_abs_8m  # metres, within the cameras' range
    assert worst_abs <= 4e-6, worst_abs        # saturated pixels (65.535 m); the bar is 1e-5


# ------------------------------------------------------------------ run when the wheels are there
def _o3d():
    return pytest.importorskip("open3d", reason="open3d is not installed here (SURVEY 8c); runs where it is")


def _cloud(rs720, max_distance=1.2):
    import repas_vision_b200 as rv
    color, depth = load_frame(CANOPY_TS[4])
    pc = rv.create_masked_pointcloud(color, depth, None, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], max_distance=max_distance)
    return rv, pc, pc.points, pc.colors


def _as_o3d(o3d, P, C=None):
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(P)
    if C is not None:
        pcd.colors = o3d.utility.Vector3dVector(C)
    return pcd


def _rows_sorted(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.gpu
def test_open3d_voxel_down_sample_and_transform(rs720):
    o3d = _o3d()
    rv, pc, P, C = _cloud(rs720)
    ref = _as_o3d(o3d, P, C).voxel_down_sample(0.005)
    got = pc.voxel_down_sample(0.005)
    rp, gp = _rows_sorted(np.asarray(ref.points)), _rows_sorted(got.points)
    assert rp.shape == gp.shape and np.abs(rp - gp).max() <= 1e-5 * np.abs(rp).max()
    T = np.eye(4)
    T[:3, :3] = [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]
    T[:3, 3] = [0.1, -0.2, 0.05]
    moved = np.asarray(_as_o3d(o3d, P).transform(T).points)
    assert np.array_equal(pc.transformed(T).points, moved)


@pytest.mark.gpu
def test_open3d_outliers_normals_icp(rs720):
    o3d = _o3d()
    rv, pc, _, _ = _cloud(rs720)
    down = pc.voxel_down_sample(0.005)
    P = down.points
    ref = _as_o3d(o3d, P)
    _, ind_ref = ref.remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0)
    _, ind = down.remove_statistical_outlier(20, 2.0)
    assert np.array_equal(np.asarray(ind), np.asarray(ind_ref))
    ref.estimate_normals(o3d.geometry.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    down.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30))
    dots = np.abs((np.asarray(ref.normals) * down.normals).sum(1))
    assert np.median(dots) >= 1 - 1e-9 and (dots > 1 - 1e-6).mean() > 0.99
    D = rv.registration.vector6d_to_matrix4d([0.01, -0.008, 0.012, 0.003, -0.002, 0.004])
    src = down.select_by_index(np.arange(0, len(down), 2)).transform(D)
    reg = rv.registration_icp(src, down, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), rv.ICPConvergenceCriteria(max_iteration=30))
    reg_ref = o3d.pipelines.registration.registration_icp(
        _as_o3d(o3d, src.points), ref, 0.02, np.eye(4), o3d.pipelines.registration.TransformationEstimationPointToPlane(),
        o3d.pipelines.registration.ICPConvergenceCriteria(max_iteration=30))
    assert abs(reg.fitness - reg_ref.fitness) <= 1e-9 and np.abs(reg.transformation - reg_ref.transformation).max() <= 1e-6


@pytest.mark.gpu
def test_open3d_reads_our_ply(rs720, tmp_path):
    o3d = _o3d()
    rv, pc, P, C = _cloud(rs720)
    path = str(tmp_path / "cloud.ply")
    rv.write_point_cloud(path, pc)
    back = o3d.io.read_point_cloud(path)
    assert np.array_equal(np.asarray(back.points), P)
    assert np.abs(np.asarray(back.colors) - C).max() <= 0.5 / 255 + 1e-12
    ref_path = str(tmp_path / "ref.ply")
    o3d.io.write_point_cloud(ref_path, _as_o3d(o3d, P, C), write_ascii=False, compressed=False)
    ours = rv.read_point_cloud(ref_path)
    assert np.array_equal(ours.points, P)


def test_pyrealsense2_project_deproject():
    rs = pytest.importorskip("pyrealsense2", reason="pyrealsense2 is not installed here (SURVEY 8c); runs where it is")
    from oracle import oracle_np as O
    intr = rs.intrinsics()
    intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy = 640, 480, 608.2336, 607.8509, 312.5224, 232.6515
    intr.model = rs.distortion.inverse_brown_conrady
    intr.coeffs = [0.04, -0.1, 0.001, -0.0005, 0.03]
    cam = dict(fx=intr.fx, fy=intr.fy, cx=intr.ppx, cy=intr.ppy, model="inverse_brown_conrady", dist=list(intr.coeffs))
    rng = np.random.default_rng(3)
    for _ in range(200):
        px, py, z = rng.uniform(0, 640), rng.uniform(0, 480), rng.uniform(0.2, 4.0)
        ref = rs.rs2_deproject_pixel_to_point(intr, [px, py], z)
        X, Y, Z = O._rs_deproject_f32(np.float32([px]), np.float32([py]), np.float32([z]), cam)
        assert np.allclose([X[0], Y[0], Z[0]], ref, rtol=0, atol=1e-6)
        uv = rs.rs2_project_point_to_pixel(intr, ref)
        pu, pv = O._rs_project_f32(np.float32([ref[0]]), np.float32([ref[1]]), np.float32([ref[2]]), cam)
        assert np.allclose([pu[0], pv[0]], uv, rtol=0, atol=1e-3)
