"""K3/K4 parity on the B200: 4x4 pose transform + merge, hash voxel grid, four-pose fusion, PLY records and the
windowed-median / NV12 helpers, through the C ABI against the CPU oracle.

Bars: voxel keys and per-voxel point counts bit-exact as a SET (Open3D's output order is unspecified, SURVEY
Appendix B.1); centroids / mean colours within 1e-5 relative; transformed coordinates bit-exact in float64."""
import os

import numpy as np
import pytest

from conftest import CAL, CANOPY_TS, load_frame, sha
from synth import synth_batch

pytestmark = pytest.mark.gpu

CENTROID_RTOL = 1e-5  # BASELINE.json north_star


@pytest.fixture(scope="module")
def rv():
    import torch
    assert torch.cuda.is_available()
    import repas_vision_b200 as rv
    return rv


@pytest.fixture(scope="module")
def O():
    from oracle import oracle_np
    return oracle_np


def _rand_pose(rng, scale=0.5):
    import cv2
    R, _ = cv2.Rodrigues(rng.uniform(-1.5, 1.5, 3))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.uniform(-scale, scale, 3)
    return T


def _sorted_voxels(keys, cent, col, cnt):
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return keys[order], cent[order], None if col is None else col[order], cnt[order]


def _check_voxels(got_pc, got_keys, got_counts, ref):
    rk, rc, rcol, rn = ref  # oracle returns keys sorted lexicographically
    gk, gc, gcol, gn = _sorted_voxels(got_keys, got_pc.points, got_pc.colors if got_pc.has_colors() else None, got_counts)
    assert gk.shape == rk.shape and np.array_equal(gk, rk), "voxel key sets differ"
    assert np.array_equal(gn, rn), "per-voxel point counts differ"
    scale = np.maximum(np.abs(rc), 1e-3)
    assert (np.abs(gc - rc) / scale).max() <= CENTROID_RTOL
    if rcol is not None:
        assert np.abs(gcol - rcol).max() <= CENTROID_RTOL


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_transform_matches_open3d_semantics(rv, O, dtype):
    rng = np.random.default_rng(5)
    P = rng.uniform(-2, 2, (100003, 3))
    C = rng.random((100003, 3))
    if dtype == "f32":
        P, C = P.astype(np.float32), C.astype(np.float32)
    T = _rand_pose(rng)
    pc = rv.PointCloud.from_arrays(P, C, dtype=dtype)
    out = pc.transform(T)
    assert out is pc  # in place, returns self like Open3D
    ref = O.transform(P, T)
    if dtype == "f64":
        assert np.array_equal(pc.points, ref)
    else:
        assert np.array_equal(pc.points.astype(np.float32), ref.astype(np.float32))
    assert np.array_equal(pc.colors, C.astype(np.float64))
    # projective last row (w != 1)
    Tp = T.copy()
    Tp[3] = [0.01, -0.02, 0.03, 1.1]
    pc2 = rv.PointCloud.from_arrays(P.astype(np.float64), None).transform(Tp)
    assert np.array_equal(pc2.points, O.transform(P.astype(np.float64), Tp))
    with pytest.raises(ValueError):
        pc.transform(np.eye(3))


def test_merge_is_concatenation(rv):
    rng = np.random.default_rng(6)
    parts = [(rng.normal(size=(n, 3)), rng.random((n, 3))) for n in (1000, 1, 0, 777)]
    clouds = [rv.PointCloud.from_arrays(p, c) for p, c in parts]
    m = rv.merge(clouds)
    assert np.array_equal(m.points, np.concatenate([p for p, _ in parts]))
    assert np.array_equal(m.colors, np.concatenate([c for _, c in parts]))
    s = clouds[0] + clouds[3]
    assert len(s) == 1777 and np.array_equal(s.points[1000:], parts[3][0])


@pytest.mark.parametrize("voxel", [0.005, 0.003, 0.05])
def test_voxel_down_sample_real_frame(rv, O, rs720, voxel):
    """mpa_icp_export.py:44,174 (5 mm), view_point_cloud.py example (3 mm) on a captured frame."""
    color, depth = load_frame(CANOPY_TS[4])
    pc = rv.create_masked_pointcloud(color, depth, None, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], max_distance=1.5)
    P, C = pc.points, pc.colors
    down, keys, counts = pc.voxel_down_sample(voxel, return_keys=True)
    _check_voxels(down, keys, counts, O.voxel_down_sample(P, C, voxel))
    assert counts.sum() == len(pc)


def test_voxel_oracle_forms_agree_and_edge_cases(rv, O):
    from oracle import oracle_c
    rng = np.random.default_rng(8)
    P = rng.uniform(-0.4, 0.4, (50000, 3))
    P[:100] = P[0]  # duplicates
    P[100:200] = np.round(P[100:200] / 0.01) * 0.01  # points exactly on voxel faces
    C = rng.random((50000, 3))
    ref = O.voxel_down_sample(P, C, 0.01)
    ck, cc, ccol, cn = _sorted_voxels(*oracle_c.voxel_down_sample(P, C, 0.01))
    assert np.array_equal(ck, ref[0]) and np.array_equal(cn, ref[3]) and np.allclose(cc, ref[1], rtol=0, atol=1e-15)
    pc = rv.PointCloud.from_arrays(P, C)
    down, keys, counts = pc.voxel_down_sample(0.01, return_keys=True)
    _check_voxels(down, keys, counts, ref)
    # no colours, float32 storage
    pc32 = rv.PointCloud.from_arrays(P.astype(np.float32), None, dtype="f32")
    d32, k32, n32 = pc32.voxel_down_sample(0.02, return_keys=True)
    _check_voxels(d32, k32, n32, O.voxel_down_sample(P.astype(np.float32), None, 0.02))
    # single point, empty cloud, bad sizes
    one = rv.PointCloud.from_arrays(P[:1], C[:1]).voxel_down_sample(0.005)
    assert len(one) == 1 and np.array_equal(one.points, P[:1])
    assert len(rv.PointCloud.from_arrays(P[:0], C[:0]).voxel_down_sample(0.005)) == 0
    with pytest.raises(RuntimeError):
        pc.voxel_down_sample(0.0)
    with pytest.raises(RuntimeError):
        pc.voxel_down_sample(1e-9)  # index range overflow -> "voxel_size is too small"


def test_four_pose_fusion(rv, O, rs720):
    """BASELINE config 4: four views, solvePnP-style T_cam_tag poses, merged in the tag frame and voxelised at 5 mm."""
    import torch
    rng = np.random.default_rng(12)
    depth, bgr = synth_batch(4, 720, 1280, seed0=40)
    cam = rv.Camera(rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], 1280, 720)
    poses = []
    for i in range(4):  # camera orbiting the tag at 0/90/180/270 degrees, 0.8 m away
        a = np.deg2rad(90.0 * i)
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    batch = rv.deproject_batch(torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda(), cam, max_distance=1.2, dtype="f64")
    clouds = [batch.frame(i) for i in range(4)]
    fused, keys, counts = rv.fuse_views(clouds, poses, 0.005, return_keys=True)
    parts, cols = [], []
    for i in range(4):
        ref = O.deproject_mask(depth[i], bgr[i], None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, r_max=1.2, out_dtype="f64")
        parts.append(O.transform(ref["points"], O.invert_rigid(poses[i])))
        cols.append(ref["colors"])
    _check_voxels(fused, keys, counts, O.voxel_down_sample(np.concatenate(parts), np.concatenate(cols), 0.005))
    assert counts.sum() == sum(len(c) for c in clouds)


def test_voxel_grid_full_size_properties(rv):
    """A 3.2 M-point cloud (several CTA parts, many joiners) through size-independent properties: the per-voxel counts sum
    to N, sum(count * centroid) reproduces the sum of the points, and the key set equals torch.unique of the float64 keys
    computed independently on the device."""
    import torch
    from repas_vision_b200 import _ops
    g = torch.Generator(device="cuda").manual_seed(99)
    n = 3_200_017
    # points on a noisy surface in scan order (runs of equal voxels) plus uniform clutter (isolated voxels)
    u = torch.arange(n, device="cuda", dtype=torch.float64)
    surf = torch.stack([(u % 2000) * 0.0007 - 0.7, torch.floor(u / 2000) * 0.0009 - 0.7,
                        0.8 + 0.001 * torch.randn(n, generator=g, device="cuda", dtype=torch.float64)])
    clutter = torch.rand((3, n), generator=g, device="cuda", dtype=torch.float64) * 2.0 - 1.0
    pick = torch.rand(n, generator=g, device="cuda") < 0.2
    xyz = torch.where(pick[None, :], clutter, surf).contiguous()
    rgb = torch.rand((3, n), generator=g, device="cuda", dtype=torch.float64)
    data = torch.cat([xyz, rgb]).contiguous()
    for voxel in (0.005, 0.03):
        r = _ops.voxel_downsample(data, n, True, voxel, want_keys=True, want_counts=True)
        m = int(r["m"].item())
        cnt = r["counts"][:m].to(torch.int64)
        assert int(cnt.sum()) == n
        cent = r["data"][:, :m]
        tot = (cent * cnt[None, :].to(torch.float64)).sum(dim=1)
        ref = data.sum(dim=1)
        assert torch.all((tot - ref).abs() <= 1e-9 * ref.abs().clamp(min=1.0))
        origin = xyz.min(dim=1).values - voxel * 0.5
        k = torch.floor((xyz - origin[:, None]) / voxel).to(torch.int64)
        packed = (k[0] << 42) | (k[1] << 21) | k[2]
        uniq, ucnt = torch.unique(packed, return_counts=True)
        got = (r["keys"][0, :m].to(torch.int64) << 42) | (r["keys"][1, :m].to(torch.int64) << 21) | r["keys"][2, :m].to(torch.int64)
        order = torch.argsort(got)
        assert m == uniq.numel() and torch.equal(got[order], uniq) and torch.equal(cnt[order], ucnt)


def test_statistical_outlier_removal_matches_oracle(rv, O, rs720):
    """create_masked_ply.py:163-170: voxel_down_sample then remove_statistical_outlier(20, 2.0).  The mean neighbour
    distances equal the brute-force oracle bit for bit, and so do the threshold and the kept index list."""
    from repas_vision_b200 import _ops
    color, depth = load_frame(CANOPY_TS[1])
    pc = rv.create_masked_pointcloud(color, depth.astype(np.float32) * np.float32(0.001), np.full(depth.shape, 255, np.uint8),
                                     rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], max_distance=1.0)
    down = pc.voxel_down_sample(0.005)
    P = down.points
    assert 2000 < len(P) < 60000
    avg = _ops.knn_mean_distance(down._data, len(down), 20).cpu().numpy()
    ref_avg = O.knn_mean_distance(P, 20)
    assert np.array_equal(avg, ref_avg)
    ref_ind, mean, std, thr = O.statistical_outlier_indices(ref_avg, 2.0)
    keep, stats = _ops.statistical_outlier_mask(torch_from(avg), 2.0)
    assert np.allclose(stats.cpu().numpy(), [mean, std, thr, float(len(P))], rtol=1e-12, atol=0)  # parallel sums; the list below is exact
    kept, ind = down.remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0)
    assert np.array_equal(ind, ref_ind) and 0 < len(ind) < len(P)
    assert int(ind[0]) == int(ref_ind[0]) and [int(v) for v in list(ind)[:5]] == ref_ind[:5].tolist()  # array-like, fetched lazily
    assert np.array_equal(down.select_by_index(ind).points, kept.points) and np.asarray(ind).dtype == np.int64
    assert np.array_equal(kept.points, P[ref_ind]) and np.array_equal(kept.colors, down.colors[ref_ind])
    # random clouds: clusters, duplicates, an isolated far point, fewer points than neighbours, float32 storage
    rng = np.random.default_rng(5)
    Q = np.concatenate([rng.normal(size=(1500, 3)) * 0.05, rng.normal(size=(800, 3)) * 0.3 + 1.0, rng.random((200, 3)) * 4.0,
                        np.repeat(rng.random((5, 3)), 30, axis=0), [[50.0, -20.0, 3.0]]])
    for k, ratio, dt in ((20, 2.0, "f64"), (1, 0.5, "f64"), (7, 1.0, "f32"), (64, 3.0, "f64")):
        pcq = rv.PointCloud.from_arrays(Q, None, dtype=dt)
        Qs = pcq.points
        got = _ops.knn_mean_distance(pcq._data, len(pcq), k).cpu().numpy()
        want = O.knn_mean_distance(Qs, k)
        assert np.array_equal(got, want), (k, dt, np.abs(got - want).max())
        _, ind = pcq.remove_statistical_outlier(k, ratio)
        assert np.array_equal(ind, O.statistical_outlier_indices(want, ratio)[0])
    # a far outlier inflates the bounding box a thousandfold: the grid is rebuilt finer and the answer stays exact
    S = np.concatenate([np.stack([rng.random(12000) * 0.6, rng.random(12000) * 0.4, rng.normal(size=12000) * 0.002], 1),
                        [[900.0, 700.0, -300.0]]])
    pcs = rv.PointCloud.from_arrays(S, None)
    assert np.array_equal(_ops.knn_mean_distance(pcs._data, len(pcs), 20).cpu().numpy(), O.knn_mean_distance(S, 20))
    # a point within rounding distance of the threshold: the statistics are redone in index order and the decision is the
    # sequential reference's
    av = rng.random(50000) * 0.01 + 0.002
    for _ in range(6):
        av[1234] = O.statistical_outlier_indices(av, 1.5)[3]
    ind_ref, mean, std, thr = O.statistical_outlier_indices(av, 1.5)
    assert abs(av[1234] - thr) <= 1e-12 * thr
    keep, stats = _ops.statistical_outlier_mask(torch_from(av), 1.5)
    assert np.array_equal(np.nonzero(keep.cpu().numpy())[0], ind_ref)
    assert stats.cpu().numpy().tolist() == [mean, std, thr, 50000.0]  # this time from the sequential pass: bit-identical
    few = rv.PointCloud.from_arrays(Q[:7], None)
    assert np.array_equal(_ops.knn_mean_distance(few._data, 7, 20).cpu().numpy(), O.knn_mean_distance(Q[:7], 20))
    assert len(rv.PointCloud.from_arrays(Q[:0], None).remove_statistical_outlier()[1]) == 0
    with pytest.raises(RuntimeError):
        pcq.remove_statistical_outlier(0, 2.0)


def torch_from(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_statistical_outlier_removal_full_size_properties(rv):
    """A regular 1200 x 1000 lattice (1.2 M points, spacing h): the 7 nearest neighbours of an interior point are itself,
    four at h and two at sqrt(2) h -- a known answer at scale -- and a few raised points are exactly the ones removed."""
    import torch
    from repas_vision_b200 import _ops
    h = 0.002
    ny, nx = 1000, 1200
    yy, xx = torch.meshgrid(torch.arange(ny, device="cuda", dtype=torch.float64), torch.arange(nx, device="cuda", dtype=torch.float64),
                            indexing="ij")
    z = torch.zeros(ny * nx, device="cuda", dtype=torch.float64)
    lifted = torch.tensor([5 * nx + 7, 500 * nx + 600, 998 * nx + 1100], device="cuda")
    z[lifted] = 0.05
    data = torch.stack([xx.reshape(-1) * h, yy.reshape(-1) * h, z]).contiguous()
    n = ny * nx
    avg = _ops.knn_mean_distance(data, n, 7)
    row, col = 300, 300
    want = (0.0 + 4 * h + 2 * (2.0 ** 0.5) * h) / 7.0
    got = float(avg[row * nx + col])
    assert abs(got - want) <= 1e-12
    keep, stats = _ops.statistical_outlier_mask(avg, 6.0)
    dropped = set(torch.nonzero(keep == 0).reshape(-1).cpu().tolist())
    # the spread of the lattice is tiny, so the threshold sits just above the interior value: out go the lifted points,
    # the lattice points that lost them as neighbours, and the four corners (three of their neighbours are further away)
    corners = {0, nx - 1, (ny - 1) * nx, ny * nx - 1}
    near = {int(i) + dy * nx + dx for i in lifted.cpu().tolist() for dy in (-1, 0, 1) for dx in (-1, 0, 1)}
    assert set(lifted.cpu().tolist()) <= dropped and corners <= dropped and dropped <= (corners | near)
    out, count, index = _ops.select_by_mask(data, n, False, keep)
    m = n - len(dropped)
    assert int(count.item()) == m
    assert torch.equal(out[:, :m], data[:, keep.bool()]) and torch.equal(index[:m], torch.nonzero(keep).reshape(-1))


def test_normals_match_oracle_and_ply_carries_them(rv, O, rs720, tmp_path):
    """create_masked_ply.py:163-177 in full: voxel grid, outlier removal, estimate_normals(Hybrid(0.02, 30)), orientation
    towards the camera, PLY with normals.  The neighbourhoods are exact; the eigenvector comes from a closed-form solver on
    the GPU and from eigh in the oracle, so normals are compared as directions (1e-7) where the two smaller eigenvalues are
    separated, and everywhere for unit length and orientation."""
    color, depth = load_frame(CANOPY_TS[2])
    pc = rv.create_masked_pointcloud(color, depth.astype(np.float32) * np.float32(0.001), np.full(depth.shape, 255, np.uint8),
                                     rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], max_distance=1.0)
    down = pc.voxel_down_sample(0.005)
    down, _ = down.remove_statistical_outlier(20, 2.0)
    P = down.points
    assert 2000 < len(P) < 60000
    down.estimate_normals(search_param=rv.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    assert down.has_normals()
    down.orient_normals_towards_camera_location(camera_location=np.array([0.0, 0.0, 0.0]))
    N = down.normals
    ref = O.estimate_normals(P, 0.02, 30, camera_location=(0.0, 0.0, 0.0))
    assert np.abs(np.linalg.norm(N, axis=1) - 1.0).max() < 1e-12
    assert ((N * (0.0 - P)).sum(axis=1) >= 0).all()
    cosang = np.abs((N * ref).sum(axis=1))
    assert np.median(cosang) > 1 - 1e-12
    assert (cosang > 1 - 1e-7).mean() > 0.995  # the rest: near-degenerate neighbourhoods (two close eigenvalues)
    same = (N * ref).sum(axis=1) > 0
    assert same[cosang > 1 - 1e-7].all()
    # a tilted plane: every interior normal is the plane normal; isolated points get (0, 0, 1) turned to the camera
    g = np.stack(np.meshgrid(np.arange(60) * 0.004, np.arange(50) * 0.004, indexing="ij"), -1).reshape(-1, 2)
    nrm = np.array([0.3, -0.2, 1.0]) / np.linalg.norm([0.3, -0.2, 1.0])
    plane = np.stack([g[:, 0], g[:, 1], 0.7 - (nrm[0] * g[:, 0] + nrm[1] * g[:, 1]) / nrm[2]], 1)
    pts = np.concatenate([plane, [[5.0, 5.0, 5.0], [-4.0, 2.0, 9.0]]])
    pp = rv.PointCloud.from_arrays(pts, None)
    pp.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30)).orient_normals_towards_camera_location((0.0, 0.0, 0.0))
    Np = pp.normals
    assert np.abs(np.abs(Np[:3000] @ nrm) - 1.0).max() < 1e-9 and (Np[:3000] @ nrm < 0).all()  # towards the camera: -n
    assert np.array_equal(Np[3000:], [[0.0, 0.0, -1.0], [0.0, 0.0, -1.0]])
    with pytest.raises(RuntimeError):
        rv.PointCloud.from_arrays(pts, None).orient_normals_towards_camera_location()
    # PLY: x y z nx ny nz red green blue, read back by the independent reader and by the device decoder
    path = tmp_path / "with_normals.ply"
    rv.write_point_cloud(str(path), down)
    header, arr = O.read_ply_minimal(str(path))
    assert [l.split()[-1] for l in header if l.startswith("property")] == ["x", "y", "z", "nx", "ny", "nz", "red", "green", "blue"]
    assert np.array_equal(np.stack([arr["nx"], arr["ny"], arr["nz"]], 1), N)
    back = rv.read_point_cloud(str(path))
    assert back.has_normals() and np.array_equal(back.normals, N) and np.array_equal(back.points, P)
    assert np.array_equal(back.colors, np.floor(down.colors * 255.0 + 0.5) / 255.0)  # std::round: halves go up


def test_pose_from_corners_feeds_fusion(rv, golden):
    """final_view.py:171-225 on exactly projected corners: same winning order and pose as the reference produced."""
    K = np.array(golden["solvepnp_K"])
    for rec in golden["solvepnp"]:
        _, rvec, tvec, err, label = rv.solve_pnp_with_best_obj_order(np.array(rec["corners_px"]), K, np.zeros((5, 1)),
                                                                     golden["solvepnp_tag_size"])
        assert label == rec["label"]
        assert np.allclose(rvec.reshape(3), rec["rvec"], atol=1e-9) and np.allclose(tvec.reshape(3), rec["tvec"], atol=1e-9)
        T = rv.pose_from_tag_corners(np.array(rec["corners_px"]), K, np.zeros((5, 1)), golden["solvepnp_tag_size"])
        assert np.allclose(T, rec["T_cam_tag"], atol=1e-9)


def test_ply_device_decode_matches_host_parser(rv, O, tmp_path):
    """read_point_cloud unpacks binary vertex records on the GPU: same arrays as the independent host reader, for float and
    double coordinates, with and without colours, and with extra vertex properties / a face element that must be skipped."""
    from repas_vision_b200 import ply as plymod
    rng = np.random.default_rng(12)
    n = 4099
    P = rng.normal(size=(n, 3)) * 0.4
    Cb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    for coord in ("double", "float"):
        path = tmp_path / f"c_{coord}.ply"
        rv.write_point_cloud(str(path), rv.PointCloud.from_arrays(P, Cb.astype(np.float64) / 255.0), coord=coord)
        assert plymod._binary_vertex_layout(str(path)) is not None
        got = rv.read_point_cloud(str(path))
        _, arr = O.read_ply_minimal(str(path))
        ref = np.stack([arr["x"], arr["y"], arr["z"]], 1).astype(np.float64)
        assert np.array_equal(got.points, ref)
        assert np.array_equal(got.colors, np.stack([arr["red"], arr["green"], arr["blue"]], 1).astype(np.float64) / 255.0)
        got32 = rv.read_point_cloud(str(path), dtype="f32")
        assert np.array_equal(got32.points, ref.astype(np.float32).astype(np.float64))
    # a hand-written file: normals and an alpha byte between the fields, colours before the coordinates, faces after
    dt = np.dtype([("red", "u1"), ("nx", "<f4"), ("x", "<f4"), ("y", "<f4"), ("green", "u1"), ("z", "<f4"), ("alpha", "u1"),
                   ("blue", "u1")])
    rec = np.zeros(n, dt)
    rec["x"], rec["y"], rec["z"] = P[:, 0], P[:, 1], P[:, 2]
    rec["red"], rec["green"], rec["blue"] = Cb[:, 0], Cb[:, 1], Cb[:, 2]
    rec["nx"], rec["alpha"] = 0.5, 7
    names = {"u1": "uchar", "<f4": "float"}
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    head += [f"property {names[dt[k].str.lstrip('|')]} {k}" for k in dt.names]
    head += ["element face 1", "property list uchar int vertex_indices", "end_header"]
    path = tmp_path / "odd.ply"
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        f.write(rec.tobytes())
        f.write(bytes([3]) + np.array([0, 1, 2], "<i4").tobytes())
    assert plymod._binary_vertex_layout(str(path))[2] == dt.itemsize
    got = rv.read_point_cloud(str(path))
    assert np.array_equal(got.points, np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float64))
    assert np.array_equal(got.colors, Cb.astype(np.float64) / 255.0)
    # no colours
    path = tmp_path / "nc.ply"
    rv.write_point_cloud(str(path), rv.PointCloud.from_arrays(P[:33], None))
    got = rv.read_point_cloud(str(path))
    assert not got.has_colors() and np.array_equal(got.points, P[:33])


def test_ply_round_trip_and_layout(rv, O, rs720, tmp_path):
    color, depth = load_frame(CANOPY_TS[0])
    pc = rv.create_masked_pointcloud(color, depth, None, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], max_distance=1.0)
    P, C = pc.points, pc.colors
    for coord, npdt in (("double", "<f8"), ("float", "<f4")):
        path = tmp_path / f"cloud_{coord}.ply"
        assert rv.write_point_cloud(str(path), pc, coord=coord)
        header, arr = O.read_ply_minimal(str(path))
        assert header[1] == "format binary_little_endian 1.0" and f"element vertex {len(pc)}" in header
        assert [h for h in header if h.startswith("property")] == [f"property {coord} x", f"property {coord} y",
                                                                   f"property {coord} z", "property uchar red",
                                                                   "property uchar green", "property uchar blue"]
        assert np.array_equal(arr["x"], P[:, 0].astype(npdt)) and np.array_equal(arr["z"], P[:, 2].astype(npdt))
        assert np.array_equal(arr["red"], np.round(C[:, 0] * 255).astype(np.uint8))
        back = rv.read_point_cloud(str(path))
        assert len(back) == len(pc)
        if coord == "double":
            assert np.array_equal(back.points, P)
        assert np.array_equal(np.round(back.colors * 255), np.round(C * 255))
    # colours survive exactly: uchar -> /255 -> *255 round trip for every byte value
    ramp = np.arange(256, dtype=np.float64)[:, None].repeat(3, 1) / 255.0
    pcr = rv.PointCloud.from_arrays(np.zeros((256, 3)), ramp)
    rv.write_point_cloud(str(tmp_path / "ramp.ply"), pcr)
    _, arr = O.read_ply_minimal(str(tmp_path / "ramp.ply"))
    assert np.array_equal(arr["green"], np.arange(256, dtype=np.uint8))
    # ascii variant parses with the independent reader too
    small = rv.PointCloud.from_arrays(P[:50], C[:50])
    rv.write_point_cloud(str(tmp_path / "a.ply"), small, write_ascii=True)
    _, arr = O.read_ply_minimal(str(tmp_path / "a.ply"))
    assert np.allclose(arr["y"], P[:50, 1], atol=1e-9)
    # no colours
    rv.write_point_cloud(str(tmp_path / "nc.ply"), rv.PointCloud.from_arrays(P[:10], None))
    header, arr = O.read_ply_minimal(str(tmp_path / "nc.ply"))
    assert arr.dtype.names == ("x", "y", "z") and np.array_equal(arr["x"], P[:10, 0])


def test_median_depth_windows_and_canopy_known_answers(rv, golden, rs720):
    """get_depth_at_pixel (canopy_return.py:279-317) on the GPU reproduces the reference's stored canopy_y goldens."""
    _, depth0 = load_frame(CANOPY_TS[0])
    recs = golden["median_depth"]
    for win in (5, 11):
        sel = [r for r in recs if r["window"] == win]
        got = rv.median_depth_windows(depth0, [(r["x"], r["y"]) for r in sel], win)
        for r, g in zip(sel, got):
            if r["depth_m"] is None:
                assert np.isnan(g)
            else:
                assert g / 1000.0 == r["depth_m"]
    class Intr:
        fx, fy, ppx, ppy = rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"]
    for rec in golden["canopy"]:
        _, depth = load_frame(rec["ts"])
        x, y = rec["pixel"]
        d = rv.get_depth_at_pixel(depth, x, y, 5)
        if d is None or d <= 0:
            d = rv.get_depth_at_pixel(depth, x, y, 11)
        assert d == rec["depth_m"]
        X, Y, Z = rv.deproject_pixel_to_point(Intr, (x, y), d)
        assert [X, Y, Z] == rec["xyz"] and f"{Y:.4f}" == rec["stored_in_reference"]


def test_nv12_to_bgr_matches_opencv(rv):
    import cv2
    rng = np.random.default_rng(2)
    # 16-byte-wide kernel (W % 16 == 0: one partial strip, five full strips, a 48-pixel tail strip) and the byte kernel
    for H, W in ((72, 128), (720, 1280), (36, 304), (50, 70)):
        nv12 = rng.integers(0, 256, (3, H * 3 // 2, W), dtype=np.uint8)
        got = rv.nv12_to_bgr(nv12, H, W)
        for b in range(3):
            ref = cv2.cvtColor(nv12[b], cv2.COLOR_YUV2BGR_NV12)
            assert np.array_equal(got[b], ref), (H, W, b)


def test_exact_division_helper_against_numpy(rv):
    """rv_div (reciprocal + two fused corrections) must equal IEEE division for the divisors this path uses: checked through
    the kernels on all 65536 raw depths and all 256 colour bytes, for several focal lengths."""
    import torch
    from oracle import oracle_np as O
    allv = np.arange(65536, dtype=np.uint16).reshape(64, 1024)
    col = (np.arange(64 * 1024 * 3) % 256).astype(np.uint8).reshape(64, 1024, 3)
    for fx, fy in ((912.350341796875, 911.7763061523438), (748.8987426757812, 748.3513793945312), (3.0, 7.0),
                   (605.2845686, 605.44233933), (1.0000000000000002, 1.9999999999999998)):
        for rule in ("mul_f32", "div_f32", "div_f64"):
            pc = rv.create_masked_pointcloud(col, allv, None, fx, fy, 511.3, 31.7, unit_rule=rule)
            ref = O.deproject_mask(allv, col, None, fx=fx, fy=fy, cx=511.3, cy=31.7, unit_rule=rule, out_dtype="f64")
            assert np.array_equal(pc.points, ref["points"]) and np.array_equal(pc.colors, ref["colors"])


@pytest.mark.parametrize("max_nn,radius", [(64, 0.05), (50, 0.012), (3, 1.0)])
def test_normals_neighbourhood_sizes(rv, O, max_nn, radius):
    """icp_cad_model.py:267 asks for max_nn=50; 64 is the largest list the kernel holds (96 KB of shared memory per CTA)."""
    from synth import bumpy_surface
    rng = np.random.default_rng(13)
    P = bumpy_surface(rng, 4000)
    pc = rv.PointCloud.from_arrays(P, None)
    pc.estimate_normals(rv.KDTreeSearchParamHybrid(radius, max_nn)).orient_normals_towards_camera_location()
    N = pc.normals
    ref = O.estimate_normals(P, radius, max_nn, camera_location=(0.0, 0.0, 0.0))
    assert np.allclose(np.linalg.norm(N, axis=1), 1.0, atol=1e-12)
    dots = (N * ref).sum(axis=1)
    if max_nn > 3:
        assert np.median(dots) > 1 - 1e-12 and (dots > 0.999).mean() > 0.99
    else:  # three neighbours: a degenerate covariance (rank <= 2), the normal of the triangle they span
        assert (np.abs(dots) > 0.999).mean() > 0.95
    with pytest.raises(Exception):
        pc.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 65))


def test_small_open3d_cloud_methods(rv, O):
    """compute_nearest_neighbor_distance (ply_to_stl.py:45,56), scale (manual_pose_verify.py:293), translate,
    paint_uniform_color, get_center: bit-exact against numpy in the reference's operation order."""
    rng = np.random.default_rng(17)
    P = np.concatenate([rng.normal(size=(3000, 3)) * 0.2, np.repeat(rng.random((3, 3)), 2, axis=0), [[9.0, 9.0, 9.0]]])
    C = rng.random(P.shape)
    pc = rv.PointCloud.from_arrays(P, C)
    d = pc.compute_nearest_neighbor_distance()
    assert d.dtype == np.float64 and np.array_equal(d, O.nearest_neighbor_distance(P))
    assert (d[3000:3006] == 0.0).all() and d[-1] > 10.0
    assert rv.PointCloud.from_arrays(P[:1], None).compute_nearest_neighbor_distance().tolist() == [0.0]
    assert np.allclose(pc.get_center(), P.mean(axis=0), rtol=1e-12)
    c = np.array([0.1, -0.2, 0.3])
    pc.estimate_normals(rv.KDTreeSearchParamHybrid(0.1, 30))
    N = pc.normals
    pc.scale(1000.0, center=(0, 0, 0))
    assert np.array_equal(pc.points, P * 1000.0) and np.array_equal(pc.normals, N) and np.array_equal(pc.colors, C)
    pc.scale(0.001, center=c)
    Q = (P * 1000.0 - c) * 0.001 + c
    assert np.array_equal(pc.points, Q)
    pc.translate((1.0, 2.0, -3.0))
    assert np.array_equal(pc.points, Q + np.array([1.0, 2.0, -3.0])) and np.array_equal(pc.normals, N)
    pc.translate((0.0, 0.0, 0.0), relative=False)
    assert np.allclose(pc.get_center(), 0.0, atol=1e-12)
    pc.paint_uniform_color([0.7, 0.7, 0.7])
    assert np.array_equal(pc.colors, np.full(P.shape, 0.7))
    bare = rv.PointCloud.from_arrays(P, None).paint_uniform_color([1.0, 0.0, 0.25])
    assert bare.has_colors() and np.array_equal(bare.colors, np.tile([1.0, 0.0, 0.25], (len(P), 1))) and np.array_equal(bare.points, P)


def test_voxel_down_sample_averages_normals(rv, O):
    """Open3D's VoxelDownSample averages colours AND normals per voxel (not renormalised)."""
    from synth import bumpy_surface
    rng = np.random.default_rng(19)
    P = bumpy_surface(rng, 20000)
    C = rng.random(P.shape)
    for colors, dtype in ((C, "f64"), (None, "f64"), (C, "f32")):
        pc = rv.PointCloud.from_arrays(P, colors, dtype=dtype)
        pc.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30)).orient_normals_towards_camera_location()
        Pg, N = pc.points, pc.normals
        down, keys, counts = pc.voxel_down_sample(0.01, return_keys=True)
        assert down.has_normals() and len(down.normals) == len(down) and 1000 < len(down) < len(P)
        rk, cent, col, cnt = O.voxel_down_sample(Pg, pc.colors if colors is not None else None, 0.01)
        _, _, nrm, _ = O.voxel_down_sample(Pg, N, 0.01)
        order = np.lexsort(keys.T[::-1])
        assert np.array_equal(keys[order], rk) and np.array_equal(counts[order], cnt)
        tol = dict(rtol=1e-12, atol=1e-14) if dtype == "f64" else dict(rtol=CENTROID_RTOL, atol=1e-7)
        assert np.allclose(down.points[order], cent, **tol)
        assert np.allclose(down.normals[order], nrm, rtol=1e-12, atol=1e-14)
        if colors is not None:
            assert np.allclose(down.colors[order], col, **tol)
        single = counts[order] == 1
        assert np.allclose(np.linalg.norm(down.normals[order][single], axis=1), 1.0, atol=1e-12)


def test_normals_follow_selection_filters_and_merge(rv):
    """Open3D's SelectByIndex and operator+= carry normals; so do the fused predicates and the outlier filter here.  A target
    cropped after estimate_normals must still feed point-to-plane ICP (mpa_icp_export.py:166-197)."""
    rng = np.random.default_rng(41)
    P = np.concatenate([rng.uniform(-0.4, 0.4, (4000, 2)), rng.normal(0.9, 0.002, (4000, 1))], axis=1)
    pc = rv.PointCloud.from_arrays(P, rng.random((4000, 3)))
    pc.estimate_normals(rv.KDTreeSearchParamHybrid(0.1, 20))
    N = pc.normals
    idx = np.array([5, 17, 3999, 0, 256])
    assert np.array_equal(pc.select_by_index(idx).normals, N[idx])
    inv = pc.select_by_index(idx, invert=True)
    keep = np.ones(4000, bool)
    keep[idx] = False
    assert np.array_equal(inv.normals, N[keep]) and np.array_equal(inv.points, P[keep])
    near = pc.select_within_distance(0.95)
    m = np.sqrt((P * P).sum(1)) < 0.95
    assert near.has_normals() and np.array_equal(near.normals, N[m]) and np.array_equal(near.points, P[m])
    box = pc.crop_aabb([-0.2, -0.1, 0.0], [0.3, 0.25, 2.0])
    mb = (P >= [-0.2, -0.1, 0.0]).all(1) & (P <= [0.3, 0.25, 2.0]).all(1)
    assert np.array_equal(box.normals, N[mb])
    zc = pc.clip_z(0.899, 0.901)
    assert np.array_equal(zc.normals, N[(P[:, 2] >= 0.899) & (P[:, 2] <= 0.901)])
    kept, ind = pc.remove_statistical_outlier(20, 2.0)
    assert np.array_equal(kept.normals, N[np.asarray(ind)])
    both = near + box
    assert both.has_normals() and np.array_equal(both.normals, np.concatenate([N[m], N[mb]]))
    plain = rv.PointCloud.from_arrays(P[:10], None)
    assert not (near + plain).has_normals()  # one side without normals: the sum has none (Open3D)
    T = _rand_pose(rng)
    moved = pc.transformed(T)
    assert moved.has_normals() and np.abs(moved.normals - N @ T[:3, :3].T).max() < 1e-12 and np.array_equal(pc.normals, N)
    # the cropped target is still a valid point-to-plane target
    src = near.select_by_index(np.arange(0, len(near), 3))
    reg = rv.registration_icp(src, near, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), rv.ICPConvergenceCriteria(max_iteration=3))
    assert reg.fitness == 1.0


def test_bounds_center_and_filter_capacity(rv):
    import ctypes as C
    import torch
    from repas_vision_b200 import _lib, _ops
    rng = np.random.default_rng(9)
    for dtype in ("f64", "f32"):
        P = rng.normal(0.0, 1.0, (50001, 3))
        pc = rv.PointCloud.from_arrays(P, None, dtype=dtype)
        Q = pc.points
        assert np.array_equal(pc.get_min_bound(), Q.min(0)) and np.array_equal(pc.get_max_bound(), Q.max(0))
        assert np.abs(pc.get_center() - Q.mean(0)).max() < 1e-12
    empty = rv.PointCloud(None, 0, False)
    assert np.array_equal(empty.get_min_bound(), np.zeros(3)) and np.array_equal(empty.get_center(), np.zeros(3))
    # the kept count is only known on the device: an output too small for all n points is refused, not overrun
    ctx = _lib.context(0)
    data = torch.zeros((3, 1000), dtype=torch.float32, device="cuda")
    out = torch.zeros((3, 10), dtype=torch.float32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    keep = torch.ones(1000, dtype=torch.uint8, device="cuda")
    ws = _ops.workspace(ctx.lib.rv_filter_workspace_bytes(1000), data.device)
    p = _lib.RvDeprojectParams()
    p.use_radius, p.r_max = 1, 1.0
    st = ctx.lib.rv_filter_cloud(ctx.handle, _ops.ptr(data), 1000, 1000, _lib.RV_F32, 0, C.byref(p), _ops.ptr(out), 10, _ops.ptr(cnt),
                                 None, _ops.ptr(ws), ws.numel(), None)
    assert st == _lib.RV_ECAPACITY
    st = ctx.lib.rv_select_by_mask(ctx.handle, _ops.ptr(data), 1000, 1000, _lib.RV_F32, 0, _ops.ptr(keep), _ops.ptr(out), 10,
                                   _ops.ptr(cnt), None, _ops.ptr(ws), ws.numel(), None)
    assert st == _lib.RV_ECAPACITY


def test_voxel_means_are_the_index_order_sums(rv, O):
    """Open3D adds the points of a voxel in index order in float64.  The grid kernel does the same whenever a voxel is made of
    at most eight runs of consecutive points: those means are bit-identical to the oracle's, not just within tolerance; the
    rest (voxels gathered from longer chains) stay within 1e-12."""
    rng = np.random.default_rng(77)
    # scan-ordered surface (runs of consecutive points per voxel, several rows per voxel) plus isolated clutter
    n = 120000
    u = np.arange(n)
    P = np.stack([(u % 600) * 0.0011 - 0.33, (u // 600) * 0.0013 - 0.13, 0.8 + 0.0005 * rng.standard_normal(n)], axis=1)
    clutter = rng.random(n) < 0.1
    P[clutter] = rng.uniform(-0.4, 0.4, (int(clutter.sum()), 3))
    C = rng.random((n, 3))
    for dtype, vox in (("f64", 0.005), ("f32", 0.004)):
        Pd = P.astype(np.float32) if dtype == "f32" else P
        Cd = C.astype(np.float32) if dtype == "f32" else C
        pc = rv.PointCloud.from_arrays(Pd, Cd, dtype=dtype)
        down, keys, counts = pc.voxel_down_sample(vox, return_keys=True)
        rk, rc, rcol, rn = O.voxel_down_sample(Pd.astype(np.float64), Cd.astype(np.float64), vox)
        gk, gc, gcol, gn = _sorted_voxels(keys, down.points, down.colors, counts)
        assert np.array_equal(gk, rk) and np.array_equal(gn, rn)
        if dtype == "f32":  # stored means are the float32 roundings of the same float64 quotients
            rc, rcol = rc.astype(np.float32).astype(np.float64), rcol.astype(np.float32).astype(np.float64)
        small = rn <= 8  # at most eight points, hence at most eight runs
        assert small.sum() > 1000
        assert np.array_equal(gc[small], rc[small]) and np.array_equal(gcol[small], rcol[small])
        exact = (gc == rc).all(axis=1)
        assert exact.mean() > 0.9
        assert np.abs(gc - rc).max() <= 1e-12 if dtype == "f64" else np.abs(gc - rc).max() <= 1e-7


def test_voxel_grid_long_chains_and_fused_equals_unfused(rv, O):
    """A coarse grid over points in random order: every voxel is a chain of thousands of single-point runs, which takes the
    atomic path (k_vox_long).  And the fused call (views transformed on the fly, merged cloud never written) gives the voxels
    of transform + merge + voxel_down_sample."""
    rng = np.random.default_rng(78)
    P = rng.uniform(0.0, 1.0, (150000, 3))
    C = rng.random((150000, 3))
    pc = rv.PointCloud.from_arrays(P, C)
    down, keys, counts = pc.voxel_down_sample(0.5, return_keys=True)
    assert len(down) <= 27 and counts.min() > 256
    _check_voxels(down, keys, counts, O.voxel_down_sample(P, C, 0.5))
    gk, gc, gcol, gn = _sorted_voxels(keys, down.points, down.colors, counts)
    ref = O.voxel_down_sample(P, C, 0.5)
    assert np.abs(gc - ref[1]).max() <= 1e-12 and np.abs(gcol - ref[2]).max() <= 1e-12
    # mixed: a few heavy voxels among many light ones
    Q = np.concatenate([rng.uniform(0.0, 0.004, (3000, 3)), rng.uniform(0.0, 1.0, (40000, 3))])
    Q = Q[rng.permutation(len(Q))]
    pq = rv.PointCloud.from_arrays(Q, None)
    dq, kq, nq = pq.voxel_down_sample(0.005, return_keys=True)
    _check_voxels(dq, kq, nq, O.voxel_down_sample(Q, None, 0.005))
    assert nq.max() >= 300
    # fused == unfused, float32 and float64 storage, 1..5 views (one of them empty)
    for dtype in ("f32", "f64"):
        views, poses = [], []
        for v in range(5):
            m = 0 if v == 2 else 20000 + 1000 * v
            pts = rng.uniform(-0.3, 0.3, (m, 3)) + [0.0, 0.0, 0.9]
            views.append(rv.PointCloud.from_arrays(pts, rng.random((m, 3)), dtype=dtype))
            poses.append(_rand_pose(rng, 0.2))
        fused, fk, fn = rv.fuse_views(views, poses, 0.01, return_keys=True)
        merged = rv.merge([c.transformed(rv.world_from_camera(T)) for c, T in zip(views, poses)])
        plain, pk, pn = merged.voxel_down_sample(0.01, return_keys=True)
        a = _sorted_voxels(fk, fused.points, fused.colors, fn)
        b = _sorted_voxels(pk, plain.points, plain.colors, pn)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])  # same values summed in the same order


def test_transform_matches_the_reference_point_form(rv, golden):
    """K3 against the reference's own `R @ p + t` (april_tag_bg_removal_pl.py:177-179) on the stored golden points and poses
    (the 4x4 the reference ships and three solvePnP results): float64 to rounding, float32 storage within one ulp."""
    g = golden["rigid_transform"]
    P = np.array(g["points"])
    for c in g["cases"]:
        T, ref = np.array(c["T"]), np.array(c["out"])
        got = rv.PointCloud.from_arrays(P, None).transform(T).points
        assert np.abs(got - ref).max() <= 2e-15
        got32 = rv.PointCloud.from_arrays(P.astype(np.float32), None, dtype="f32").transform(T).points
        ref32 = np.stack([T[:3, :3] @ p + T[:3, 3] for p in P.astype(np.float32).astype(np.float64)])
        assert np.abs(got32 - ref32).max() <= 4e-7  # float32 spacing at |coordinate| <= 4 m


def test_rotate_bounding_box_and_crop_like_open3d(rv):
    """PointCloud.rotate (R (p - c) + c, normals by R), get_axis_aligned_bounding_box, crop and get_rotation_matrix_from_xyz
    as the CAD-placement scripts call them (mpa_icp.py:95-96,411; mpa_final_view_with_export.py:412)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(12)
    P = rng.uniform(-1.0, 1.0, (5000, 3)) + (0.2, -0.1, 1.5)
    C = rng.uniform(0.0, 1.0, (5000, 3))
    ang = (0.3, -0.7, 1.1)
    R = rv.PointCloud.get_rotation_matrix_from_xyz(ang)
    assert np.allclose(R, Rotation.from_euler("XYZ", ang).as_matrix(), atol=1e-15)  # Rx Ry Rz
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-15)
    c = np.array([0.1, 0.2, 1.4])
    pc = rv.PointCloud.from_arrays(P, C)
    pc.estimate_normals(rv.KDTreeSearchParamHybrid(0.2, 20))
    N = pc.normals.copy()
    assert pc.rotate(R, center=c) is pc
    ref = ((P - c) @ R.T) + c
    assert np.allclose(pc.points, ref, rtol=0, atol=1e-15)
    # the same operations in the same order: translate, then ((R0 x + R1 y) + R2 z) + c
    d = P - c
    exact = np.stack([((R[i, 0] * d[:, 0] + R[i, 1] * d[:, 1]) + R[i, 2] * d[:, 2]) + c[i] for i in range(3)], axis=1)
    assert np.array_equal(pc.points, exact)
    assert np.allclose(pc.normals, N @ R.T, atol=1e-15) and np.array_equal(pc.colors, C)
    # default centre: the cloud's own centre stays where it is
    pc2 = rv.PointCloud.from_arrays(P, None)
    before = pc2.get_center()
    pc2.rotate(R)
    assert np.allclose(pc2.get_center(), before, atol=1e-12)
    box = pc2.get_axis_aligned_bounding_box()
    assert np.array_equal(box.min_bound, pc2.points.min(axis=0)) and np.array_equal(box.get_max_bound(), pc2.points.max(axis=0))
    assert np.allclose(box.get_center(), (box.min_bound + box.max_bound) / 2) and box.volume() > 0
    small = rv.AxisAlignedBoundingBox(box.get_center() - 0.3, box.get_center() + 0.3)
    inside = pc2.crop(small)
    Q = pc2.points
    keep = np.all((Q >= small.min_bound) & (Q <= small.max_bound), axis=1)
    assert np.array_equal(inside.points, Q[keep]) and 0 < keep.sum() < len(Q)


def test_open3d_shaped_namespace_builds_the_same_clouds(rv, tmp_path):
    """`from repas_vision_b200 import o3d_compat as o3d`: the reference's way of building and using a cloud
    (create_masked_ply.py:102-104,168-177; mpa_icp_export.py:174-197) gives what the package's own entry points give."""
    from repas_vision_b200 import o3d_compat as o3d
    from synth import bumpy_surface
    rng = np.random.default_rng(8)
    P = bumpy_surface(rng, 6000)
    C = rng.uniform(0.0, 1.0, (6000, 3))
    pcd = o3d.geometry.PointCloud()
    assert pcd.is_empty() and not pcd.has_colors() and np.asarray(pcd.points).shape == (0, 3)
    pcd.points = o3d.utility.Vector3dVector(P)
    pcd.colors = o3d.utility.Vector3dVector(C)
    ref = rv.PointCloud.from_arrays(P, C)
    assert len(pcd) == 6000 and pcd.has_colors()
    assert np.array_equal(np.asarray(pcd.points), P) and np.array_equal(np.asarray(pcd.colors), C)
    with pytest.raises(ValueError):
        pcd.colors = o3d.utility.Vector3dVector(C[:10])
    with pytest.raises(RuntimeError):
        o3d.utility.Vector3dVector(np.zeros((4, 2)))
    # the same pipeline through both spellings
    a, b = pcd.voxel_down_sample(0.01), ref.voxel_down_sample(0.01)
    ka, kb = np.lexsort(a.points.T), np.lexsort(b.points.T)
    # (voxels of more than eight runs are summed with float64 atomics: equal to rounding, not to the bit)
    assert len(a) == len(b) and np.allclose(a.points[ka], b.points[kb], rtol=0, atol=1e-13) and np.allclose(a.colors[ka], b.colors[kb], rtol=0, atol=1e-13)
    a2, ind = a.remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0)
    a2.estimate_normals(search_param=o3d.geometry.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    a2.orient_normals_towards_camera_location(np.array([0.0, 0.0, 0.0]))
    assert a2.has_normals() and np.asarray(a2.normals).shape == (len(a2), 3)
    D = rv.registration.vector6d_to_matrix4d([0.01, -0.01, 0.02, 0.002, -0.001, 0.003])
    src = o3d.geometry.PointCloud()
    src.points = o3d.utility.Vector3dVector((np.asarray(a2.points)[::2] @ D[:3, :3].T) + D[:3, 3])
    reg = o3d.pipelines.registration.registration_icp(
        src, a2, 0.02, np.eye(4), o3d.pipelines.registration.TransformationEstimationPointToPlane(),
        o3d.pipelines.registration.ICPConvergenceCriteria(max_iteration=50, relative_fitness=1e-6, relative_rmse=1e-6))
    assert reg.fitness > 0.99 and np.allclose(reg.transformation @ D, np.eye(4), atol=1e-6)
    # points assigned again with the same length keep colours and normals; another length drops them
    n0 = np.asarray(a2.normals).copy()
    a2.points = o3d.utility.Vector3dVector(np.asarray(a2.points) + 1.0)
    assert a2.has_colors() and np.array_equal(np.asarray(a2.normals), n0)
    a2.normals = o3d.utility.Vector3dVector(-n0)
    assert np.array_equal(np.asarray(a2.normals), -n0)
    a2.points = o3d.utility.Vector3dVector(P[:100])
    assert len(a2) == 100 and not a2.has_colors() and not a2.has_normals()
    # PLY through the namespace
    path = tmp_path / "cloud.ply"
    assert o3d.io.write_point_cloud(str(path), pcd, write_ascii=False, compressed=True)
    back = o3d.io.read_point_cloud(str(path))
    assert np.array_equal(np.asarray(back.points), P) and np.array_equal(np.asarray(back.colors), np.round(C * 255.0) / 255.0)
    with pytest.raises(AttributeError, match="outside the point-cloud path"):
        o3d.geometry.TriangleMesh
