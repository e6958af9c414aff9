"""Pins the CPU oracle to the reference: every golden in tests/golden/reference_golden.json was
produced by the reference's own functions (tests/golden/make_golden.py); the four canopy_y
values are fixtures the reference itself ships.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import CAL, CANOPY_TS, FRAMES, blob_mask, load_frame, sha
from oracle import oracle_c, oracle_np as O
from synth import bumpy_surface


def test_canopy_known_answers(golden, rs720):
    """realsense_d415i/canopy_detection/new-captures/canopy_y_*.txt: median depth window +
    /1000.0 + pinhole deprojection, formatted "{Y:.4f}" (canopy_return.py:404-407)."""
    assert len(golden["canopy"]) == 4
    for rec in golden["canopy"]:
        _, depth = load_frame(rec["ts"])
        x, y = rec["pixel"]
        d = O.median_depth_window(depth, x, y, 5)
        if d is None or d <= 0:
            d = O.median_depth_window(depth, x, y, 11)
        dm = d / 1000.0
        assert dm == rec["depth_m"]
        X, Y, Z = O.deproject_pixel_to_point(rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], (x, y), dm)
        assert [X, Y, Z] == rec["xyz"]
        stored = open(os.path.join(FRAMES, f"canopy_y_{rec['ts']}.txt")).read().strip()
        assert f"{Y:.4f}" == stored == rec["stored_in_reference"]


def test_median_depth_windows(golden, rs720):
    _, depth = load_frame(CANOPY_TS[0])
    for rec in golden["median_depth"]:
        d = O.median_depth_window(depth, rec["x"], rec["y"], rec["window"])
        if rec["depth_m"] is None:
            assert d is None
        else:
            assert d / 1000.0 == rec["depth_m"]
            xyz = O.deproject_pixel_to_point(rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], (rec["x"], rec["y"]),
                                             d / 1000.0)
            assert list(xyz) == rec["xyz"]


def test_depth_to_meters(golden):
    _, depth = load_frame(CANOPY_TS[0])
    assert sha(O.depth_to_meters(depth, "mul_f32")) == golden["depth_to_meters"]["depth_m"]["sha256"]
    allv = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    assert sha(O.depth_to_meters(allv, "mul_f32")) == golden["depth_to_meters"]["all_u16"]["sha256"]
    # the three unit rules differ in the last ulp for many values
    a = O.depth_to_meters(allv, "mul_f32")
    b = O.depth_to_meters(allv, "div_f32")
    assert a.dtype == b.dtype == np.float32 and (a != b).any()
    assert np.abs(a.astype(np.float64) - b).max() < 1e-5


@pytest.mark.parametrize("idx", range(5))
def test_masked_cloud_real_frames(golden, rs720, idx):
    ts = CANOPY_TS[idx]
    color, depth = load_frame(ts)
    dm = O.depth_to_meters(depth, "mul_f32")
    recs = [r for r in golden["masked_cloud"] if r["ts"] == ts]
    assert len(recs) == 3
    for rec in recs:
        h, w = depth.shape
        mask = np.full((h, w), 255, np.uint8) if rec["variant"] == "all" else blob_mask(h, w, rec["mask_seed"])
        P, C = O.create_masked_pointcloud(color, dm, mask, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"],
                                          invert_mask=rec["variant"] == "blob_inverted")
        assert P.shape[0] == rec["n"]
        assert sha(P) == rec["points"]["sha256"]  # float64, bit-exact, same order
        assert sha(C) == rec["colors"]["sha256"]
        # the fused-kernel contract reproduces the same cloud in its float64 mode
        out = O.deproject_mask(depth, color, mask, fx=rs720["fx"], fy=rs720["fy"], cx=rs720["cx"], cy=rs720["cy"],
                               invert_mask=rec["variant"] == "blob_inverted", out_dtype="f64")
        assert sha(out["points"]) == rec["points"]["sha256"] and sha(out["colors"]) == rec["colors"]["sha256"]
        if rec["variant"] == "all":
            keep = O.distance_mask(P, 1.0)
            assert int(keep.sum()) == rec["dist_lt_1m"]["kept"]
            assert sha(np.packbits(keep)) == rec["dist_lt_1m"]["mask"]["sha256"]
            assert sha(P[keep]) == rec["dist_lt_1m"]["points"]["sha256"]
            fused = O.deproject_mask(depth, color, None, fx=rs720["fx"], fy=rs720["fy"], cx=rs720["cx"],
                                     cy=rs720["cy"], r_max=1.0, out_dtype="f64")
            assert sha(fused["points"]) == rec["dist_lt_1m"]["points"]["sha256"]
            kz = O.z_clip_mask(P, 0.15, 8.0)
            assert sha(np.packbits(kz)) == rec["zclip_0p15_8"]["mask"]["sha256"]
            ka = O.aabb_mask(P, rec["aabb"]["min"], rec["aabb"]["max"])
            assert int(ka.sum()) == rec["aabb"]["kept"] and sha(np.packbits(ka)) == rec["aabb"]["mask"]["sha256"]
            # compiled loop form agrees with the numpy form
            Pc, Cc = oracle_c.deproject_masked(depth, color, None, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"])
            assert sha(Pc) == rec["points"]["sha256"] and sha(Cc) == rec["colors"]["sha256"]


def test_masked_cloud_float_depth_nan_inf(golden):
    rngf = np.random.default_rng(11)
    dm = (rngf.uniform(0.2, 4.0, (48, 64))).astype(np.float32)
    dm[rngf.random((48, 64)) < 0.1] = np.nan
    dm[rngf.random((48, 64)) < 0.05] = np.inf
    dm[rngf.random((48, 64)) < 0.05] = -1.0
    dm[rngf.random((48, 64)) < 0.1] = 0.0
    col = rngf.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    msk = (rngf.random((48, 64)) < 0.7).astype(np.uint8) * 255
    P, C = O.create_masked_pointcloud(col, dm, msk, 60.0, 61.0, 31.5, 23.25)
    g = golden["masked_cloud_float_small"]
    assert P.shape[0] == g["n"] and sha(P) == g["points"]["sha256"] and sha(C) == g["colors"]["sha256"]
    out = O.deproject_mask(dm, col, msk, fx=60.0, fy=61.0, cx=31.5, cy=23.25, depth_kind="f32", out_dtype="f64")
    assert sha(out["points"]) == g["points"]["sha256"]


def test_final_view_median_and_pixel_to_3d(golden, rs720):
    """final_view.py:132-146 -- median of finite positive float depths; (u-cx)/fx*Z order (within 1 ulp
    of the kernel's (u-cx)*Z/fx order)."""
    _, depth = load_frame(CANOPY_TS[0])
    Zm = depth.astype(np.float32) * np.float32(0.001)
    for rec in golden["final_view_median"]:
        u, v, win = rec["u"], rec["v"], rec["win"]
        r = max(1, win // 2)
        patch = Zm[max(0, v - r):min(720, v + r + 1), max(0, u - r):min(1280, u + r + 1)]
        patch = patch[np.isfinite(patch) & (patch > 0)]
        z = float(np.median(patch)) if patch.size else 0.0
        assert z == rec["z"]
        x, y, zz = O.deproject_pixel_to_point(rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], (float(u), float(v)), z)
        assert np.allclose([x, y, zz], rec["p"], rtol=1e-15, atol=1e-15)


def test_statistical_outlier_oracle_against_independent_kdtree():
    """Open3D is absent (parity unpinned), so the brute-force restatement of RemoveStatisticalOutliers is cross-checked
    against an independent formulation: scipy's cKDTree for the k nearest neighbours, numpy reductions for the statistics."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(3)
    P = np.concatenate([rng.normal(size=(1200, 3)) * 0.1, rng.random((60, 3)) * 3.0 - 1.5])
    avg = O.knn_mean_distance(P, 20)
    d, _ = cKDTree(P).query(P, k=20)
    assert d[:, 0].max() == 0.0  # the point itself is its own first neighbour, as in KDTreeFlann::SearchKNN
    assert np.allclose(avg, d.mean(axis=1), rtol=1e-12, atol=0)
    ind, mean, std, thr = O.statistical_outlier_indices(avg, 2.0)
    assert np.isclose(mean, avg.mean(), rtol=1e-12) and np.isclose(std, avg.std(ddof=1), rtol=1e-12)
    assert np.array_equal(ind, np.nonzero(avg < avg.mean() + 2.0 * avg.std(ddof=1))[0])
    assert 0 < len(P) - len(ind) < 100  # the sparse clutter goes, the cluster stays
    assert (ind < 1200).mean() > 0.97


def test_icp_oracle_against_independent_kdtree_and_ground_truth():
    """Open3D is absent (parity unpinned): the restatement of GetRegistrationResultAndCorrespondences is cross-checked against
    scipy's cKDTree, and the whole RegistrationICP loop (point-to-plane and point-to-point) must walk a displaced copy of a
    surface back onto it, i.e. return the inverse of the displacement."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(11)
    tgt = bumpy_surface(rng, 4000)
    src = bumpy_surface(rng, 1500)
    x = np.array([0.02, -0.015, 0.03, 0.004, -0.003, 0.005])
    D = O.vector6d_to_matrix4d(x)
    assert np.allclose(D[:3, :3] @ D[:3, :3].T, np.eye(3), atol=1e-15) and np.isclose(np.linalg.det(D[:3, :3]), 1.0)
    moved = O.transform(src, D)
    near, fit, rmse = O.nearest_correspondences(moved, tgt, 0.01)
    d, j = cKDTree(tgt).query(moved, k=1)
    ok = d * d < 0.01 * 0.01
    assert np.array_equal(near >= 0, ok) and np.array_equal(near[ok], j[ok].astype(np.int32))
    assert np.isclose(fit, ok.mean(), rtol=0, atol=0) and np.isclose(rmse, np.sqrt((d[ok] ** 2).mean()), rtol=1e-12)
    # a point exactly max_distance away is NOT a correspondence (d2 < r2, strict)
    n2, f2, _ = O.nearest_correspondences(np.array([[0.0, 0.0, 0.5]]), np.array([[0.0, 0.0, 0.0]]), 0.5)
    assert n2[0] == -1 and f2 == 0.0
    # normals of the analytic surface are not needed exactly: estimate them like the scripts do
    N = O.estimate_normals(tgt, 0.02, 30, camera_location=(0.0, 0.0, 0.0))
    T, fit, rmse, near, it = O.registration_icp(moved, tgt, 0.02, target_normals=N, point_to_plane=True, max_iteration=50)
    assert it < 50 and fit > 0.95
    assert np.allclose(T @ D, np.eye(4), atol=2e-3)
    T2, fit2, rmse2, _, it2 = O.registration_icp(moved, tgt, 0.02, point_to_plane=False, max_iteration=200)
    assert fit2 > 0.95 and np.allclose(T2 @ D, np.eye(4), atol=6e-3)
    # with an initial guess the loop starts from it: the exact inverse leaves nothing to do
    T3, fit3, rmse3, _, it3 = O.registration_icp(moved, tgt, 0.02, init=np.linalg.inv(D), target_normals=N, max_iteration=30)
    assert it3 <= 3 and np.allclose(T3 @ D, np.eye(4), atol=1e-3)


def test_kdtree_formulations_agree_with_the_brute_force_restatements():
    """bench.py times the KD-tree formulations as the CPU side of the 8f-4 rows; they must say the same as the restatements
    the GPU is checked against."""
    rng = np.random.default_rng(8)
    P = bumpy_surface(rng, 3000, noise=0.0005)
    assert np.allclose(O.knn_mean_distance_kdtree(P, 20), O.knn_mean_distance(P, 20), rtol=1e-12, atol=0)
    a = O.estimate_normals(P, 0.02, 30, camera_location=(0, 0, 0))
    b = O.estimate_normals_kdtree(P, 0.02, 30, camera_location=(0, 0, 0))
    assert np.median(np.abs((a * b).sum(axis=1))) > 1 - 1e-12 and ((a * b).sum(axis=1) > 0.999).mean() > 0.99
    Q = bumpy_surface(rng, 800, noise=0.002)
    n1, f1, r1 = O.nearest_correspondences(Q, P, 0.004)
    n2, f2, r2 = O.nearest_correspondences_kdtree(Q, P, 0.004)
    assert np.array_equal(n1, n2) and f1 == f2 and np.isclose(r1, r2, rtol=1e-12)


def test_rigid_transform_matches_the_reference_point_form(golden):
    """`geom.transform(T)` is Open3D's, but the reference states the same map itself for single points:
    transform_point_tag_local_to_camera = R @ p + t (april_tag_bg_removal_pl.py:177-179).  Its outputs on 128 seeded points
    under the pose file the reference ships (6dof/20250917_164430.txt) and three solvePnP poses are stored in the goldens;
    the oracle's Open3D-ordered arithmetic agrees with them to rounding (the matrix product may fuse or reorder: <= 2e-15 m)."""
    from oracle import oracle_np as O
    g = golden["rigid_transform"]
    P = np.array(g["points"])
    assert len(g["cases"]) == 4
    for c in g["cases"]:
        got = O.transform(P, np.array(c["T"]))
        assert np.abs(got - np.array(c["out"])).max() <= 2e-15
