"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/repas_vision.h declares,
the host logic (calibration loaders, pose helper, PLY header, sharding) behaves like the reference's, and the
product refuses to compute without a GPU instead of falling back."""
import ctypes
import json
import os
import re
import socket
import sys

import numpy as np
import pytest

from conftest import CAL, GOLDEN, ROOT

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def built():
    entry.build()
    import repas_vision_b200 as rv
    return rv


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "repas_vision.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from repas_vision_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/repas_vision.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.rv_abi_version() == 2
    assert b"sm_100a" in lib.rv_build_info()
    assert lib.rv_sizeof_cam() == ctypes.sizeof(_lib.RvCam)
    assert lib.rv_sizeof_deproject_params() == ctypes.sizeof(_lib.RvDeprojectParams)
    assert lib.rv_status_string(2) == b"output capacity exceeded"
    # size queries are pure host functions
    assert lib.rv_deproject_workspace_bytes(1, 720, 1280) == 128 + 450 * 8
    # per frame: 8 B per depth pixel, 8 B per 32-pixel row segment, counter + 256-entry list per 64x32 colour tile
    per_frame = 640 * 480 * 8 + 480 * 20 * 8 + 2048 + 20 * 23 * 256 * 4  # each part is a multiple of 256 B here
    tables = (640 + 480) * 2 * 4  # normalised corner coordinates per column / row, 256-byte multiple here
    assert lib.rv_register_workspace_bytes(64, 480, 640, 720, 1280) == tables + 16 * per_frame
    # header + part counters, keys, chain heads, list, links, run-head bits, long-voxel pool, a fusion's transformed xyz (each
    # rounded up to 256 B)
    assert lib.rv_voxel_workspace_bytes(1000) == (256 + 2048 * 8) + 256 + 24064 + 8192 + 4096 + 256 + 112 * 64 + 24064


def test_shared_object_holds_only_sm100a_code(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", built.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rv = built
    z = np.ones((4, 4), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rv.create_masked_pointcloud(np.zeros((4, 4, 3), np.uint8), z, np.ones((4, 4), np.uint8), 1, 1, 0, 0)
    with pytest.raises(RuntimeError):
        rv.register_depth_to_color(np.zeros((4, 4), np.uint16), rv.Camera(1, 1, 0, 0, 4, 4), rv.Camera(1, 1, 0, 0, 4, 4))
    from repas_vision_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.load().rv_create(0, ctypes.byref(h)) == _lib.RV_ECUDA


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "repas_vision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_intrinsics_loaders_match_reference_outputs(built, golden, tmp_path):
    rv = built
    for name, rec in golden["intrinsics"].items():
        vals = rv.load_color_intrinsics(os.path.join(CAL, name))
        assert list(vals) == rec["load"]
        assert list(rv.scale_intrinsics(*vals[:4], vals[4], vals[5], 640, 360)) == rec["scaled_640x360"]
        assert list(rv.scale_intrinsics(*vals[:4], 0, 0, 640, 360)) == rec["scaled_noop"]
    with pytest.raises(FileNotFoundError):
        rv.load_color_intrinsics(tmp_path / "missing.json")
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"fx": 1, "fy": 1, "cx": 1}))
    with pytest.raises(KeyError):
        rv.load_color_intrinsics(bad)
    nested = tmp_path / "nested.json"
    nested.write_text(json.dumps({"color_intrinsics": {"fx": 2, "fy": 3, "cx": 4, "cy": 5}}))
    assert rv.load_color_intrinsics(nested) == (2.0, 3.0, 4.0, 5.0, 0, 0)
    K, dist, w, h = rv.load_intrinsics_json(os.path.join(CAL, "checkerboard_color_intrinsics_2025-08-26T183535.json"))
    assert K.dtype == np.float32 and dist.shape == (5,) and (w, h) == (1280, 720)
    K, dist, wh = rv.load_intrinsics(os.path.join(CAL, "factory_color_intrinsics_1280_720.json"))
    assert K[0, 2] == 628.7836303710938 and wh == (1280, 720) and not dist.any()
    cam = rv.load_camera(os.path.join(CAL, "checkerboard_color_intrinsics_2025-08-26T183535.json"))
    assert cam.model == "brown_conrady" and cam.distorted and cam.dist[0] == 0.09217283086787045
    cam = rv.load_camera(os.path.join(CAL, "factory_color_intrinsics_640_480.json"))
    assert cam.model == "none" and cam.cx == 312.52239990234375 and (cam.width, cam.height) == (640, 480)
    R, t = rv.read_depth_to_color_extrinsics(os.path.join(CAL, "factory_d2c_extrinsics.json"))
    assert R.shape == (3, 3) and abs(t[0] - 0.014984656) < 1e-8
    R, t = rv.load_extrinsics(os.path.join(CAL, "factory_extrinsics_d2c_2025-09-08T143506.json"))
    assert np.array_equal(R, np.eye(3)) and not t.any()
    T = rv.load_transform_matrix(os.path.join(CAL, "20250917_164430.txt"))
    assert T.shape == (4, 4) and np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-5)
    (tmp_path / "t3.txt").write_text("1 0 0\n0 1 0\n0 0 1\n")
    with pytest.raises(ValueError):
        rv.load_transform_matrix(tmp_path / "t3.txt")


def test_pose_helper_matches_reference_goldens(built, golden):
    rv = built
    K = np.array(golden["solvepnp_K"])
    for rec in golden["solvepnp"]:
        _, rvec, tvec, err, label = rv.solve_pnp_with_best_obj_order(np.array(rec["corners_px"]), K, np.zeros((5, 1)),
                                                                     golden["solvepnp_tag_size"])
        assert label == rec["label"] and abs(err - rec["err_px"]) < 1e-9
        assert np.allclose(rvec.reshape(3), rec["rvec"], atol=1e-9) and np.allclose(tvec.reshape(3), rec["tvec"], atol=1e-9)
        T = rv.pose_from_tag_corners(np.array(rec["corners_px"]), K, np.zeros((5, 1)), golden["solvepnp_tag_size"])
        assert np.allclose(T, rec["T_cam_tag"], atol=1e-9)
        assert np.allclose(rv.world_from_camera(T) @ T, np.eye(4), atol=1e-12)


def test_ply_header_and_reader(built, tmp_path):
    from repas_vision_b200 import ply
    from oracle import oracle_np as O
    hdr = ply.ply_header(3, True, "double").decode()
    assert hdr.splitlines()[:4] == ["ply", "format binary_little_endian 1.0", "comment Created by Open3D", "element vertex 3"]
    assert hdr.endswith("end_header\n")
    rec = np.zeros(3, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    rec["x"], rec["y"], rec["z"], rec["red"] = [1, 2, 3], [4, 5, 6], [7, 8, 9], [255, 0, 128]
    p = tmp_path / "h.ply"
    p.write_bytes(ply.ply_header(3, True, "float") + rec.tobytes())
    header, arr = ply.read_ply_vertices(p)
    _, arr2 = O.read_ply_minimal(str(p))
    assert np.array_equal(arr, arr2) and arr["red"].tolist() == [255, 0, 128]
    with pytest.raises(RuntimeError):
        (tmp_path / "x.ply").write_bytes(b"not a ply")
        ply.read_ply_vertices(tmp_path / "x.ply")


def test_shard_range_partitions_exactly(built):
    from repas_vision_b200.shard import shard_range
    for total in (0, 1, 7, 8, 8192, 1023):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from repas_vision_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        total = 11
        a, b = shard.shard_range(total, rank, world)
        local = torch.arange(a, b, dtype=torch.int64) * 10 + 1  # "counts" of this rank's frames
        allc = shard.gather_counts(local, total)
        data = torch.arange(6 * (rank + 2), dtype=torch.float64).reshape(6, rank + 2) + 100 * rank
        merged, sizes = shard.gather_clouds(data, rank + 1, dst=0)
        tmax = shard.max_over_ranks(1.0 + rank)
        tsum = shard.sum_over_ranks(float(b - a))
        q.put((rank, allc.tolist(), None if merged is None else merged.tolist(), sizes, tmax, tsum))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gathers(built):
    """world_size 2 on CPU: the only collectives of the path (counts all-gather, variable-length cloud gather)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect_counts = [i * 10 + 1 for i in range(11)]
    for rank, allc, merged, sizes, tmax, tsum in res:
        assert allc == expect_counts and sizes == [1, 2] and tmax == 2.0 and tsum == 11.0
    m0 = np.array(res[0][2])
    assert m0.shape == (6, 3)
    assert m0[:, 0].tolist() == [0.0, 2.0, 4.0, 6.0, 8.0, 10.0]  # rank 0: data[:, :1] of a [6,2] arange
    assert m0[:, 1:].tolist() == (np.arange(18).reshape(6, 3)[:, :2] + 100).tolist()
    assert res[1][2] is None


def test_binary_ply_layout_parser(tmp_path):
    """Host half of the device-side PLY decode: offsets of x, y, z / red, green, blue inside a packed vertex record."""
    from repas_vision_b200 import ply
    p = tmp_path / "a.ply"
    head = ("ply\nformat binary_little_endian 1.0\ncomment x\nelement vertex 2\nproperty double x\nproperty double y\n"
            "property double z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode()
    p.write_bytes(head + bytes(2 * 27))
    assert ply._binary_vertex_layout(str(p)) == (len(head), 2, 27, [0, 8, 16], "f64", [24, 25, 26], None)
    q = tmp_path / "b.ply"
    head = ("ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty uchar blue\nproperty float z\nproperty float nx\n"
            "property float x\nproperty float y\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n").encode()
    q.write_bytes(head + bytes(17))
    assert ply._binary_vertex_layout(str(q)) == (len(head), 1, 17, [9, 13, 1], "f32", None, None)  # no red/green, no ny/nz
    for bad in ("format ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\n",
                "format binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty double y\nproperty float z\n",
                "format binary_little_endian 1.0\nelement face 1\nproperty list uchar int vertex_indices\nelement vertex 1\n"
                "property float x\nproperty float y\nproperty float z\n",
                "format binary_little_endian 1.0\nelement vertex 1\nproperty int x\nproperty int y\nproperty int z\n"):
        r = tmp_path / "c.ply"
        r.write_bytes(("ply\n" + bad + "end_header\n").encode() + bytes(64))
        assert ply._binary_vertex_layout(str(r)) is None  # the host reader takes these


def test_icp_host_steps_match_the_oracle(built):
    """registration.py keeps the 6x6 solve / Umeyama step and Eigen's isIdentity on the host: fed with the sums the GPU
    produces (computed here with numpy), they must return the oracle's update."""
    from oracle import oracle_np as O
    from repas_vision_b200 import registration as R
    from synth import bumpy_surface
    rng = np.random.default_rng(31)
    tgt = bumpy_surface(rng, 1500)
    src = O.transform(bumpy_surface(rng, 600), O.vector6d_to_matrix4d(np.array([0.02, 0.01, -0.02, 0.003, 0.002, -0.004])))
    N = O.estimate_normals(tgt, 0.03, 30, camera_location=(0, 0, 0))
    near, _, _ = O.nearest_correspondences(src, tgt, 0.02)
    sel = near >= 0
    S, T, Nt = src[sel], tgt[near[sel]], N[near[sel]]
    r = ((S - T) * Nt).sum(axis=1)
    J = np.concatenate([np.cross(S, Nt), Nt], axis=1)
    plane = np.zeros(32)
    plane[0], plane[1], plane[2] = sel.sum(), ((S - T) ** 2).sum(), (r * r).sum()
    plane[3:9], plane[9:30] = J.T @ r, (J.T @ J)[np.triu_indices(6)]
    got = R.TransformationEstimationPointToPlane().update_from_sums(plane)
    assert np.allclose(got, O.point_to_plane_update(src, tgt, N, near), rtol=0, atol=1e-12)
    point = np.zeros(32)
    point[0], point[1], point[2] = sel.sum(), plane[1], (S * S).sum()
    point[3:6], point[6:9], point[9:18] = S.sum(axis=0), T.sum(axis=0), (T.T @ S).reshape(-1)
    for scaling in (False, True):
        got = R.TransformationEstimationPointToPoint(scaling).update_from_sums(point)
        assert np.allclose(got, O.point_to_point_update(src, tgt, near, scaling), rtol=0, atol=1e-10)
    assert np.array_equal(R.TransformationEstimationPointToPlane().update_from_sums(np.zeros(32)), np.eye(4))
    assert np.array_equal(R.TransformationEstimationPointToPoint().update_from_sums(np.zeros(32)), np.eye(4))
    x = np.array([0.3, -0.2, 0.5, 1.0, 2.0, 3.0])
    assert np.allclose(R.vector6d_to_matrix4d(x), O.vector6d_to_matrix4d(x), rtol=0, atol=1e-16)
    # Eigen's isIdentity(1e-12): exact identity and rounding noise pass, a micrometre shift does not
    assert R._is_identity(np.eye(4)) and R._is_identity(np.eye(4) + 1e-14)
    shifted = np.eye(4)
    shifted[0, 3] = 1e-6
    assert not R._is_identity(shifted)
    c = R.ICPConvergenceCriteria()
    assert (c.relative_fitness, c.relative_rmse, c.max_iteration) == (1e-6, 1e-6, 30)


def test_device_index_list_behaves_like_the_index_array(built):
    """remove_statistical_outlier's `ind` (Open3D: IntVector) is fetched on first use; until then only its length is known."""
    import torch
    from repas_vision_b200.cloud import DeviceIndexList
    ind = DeviceIndexList(torch.tensor([4, 7, 9, 12], dtype=torch.int64))
    assert len(ind) == 4 and ind._host is None
    assert np.array_equal(np.asarray(ind), [4, 7, 9, 12]) and np.asarray(ind).dtype == np.int64
    assert int(ind[1]) == 7 and [int(v) for v in ind] == [4, 7, 9, 12] and np.array_equal(ind[1:3], [7, 9])
    assert np.asarray(ind, dtype=np.int32).dtype == np.int32
    assert len(DeviceIndexList(torch.zeros(0, dtype=torch.int64))) == 0


def test_depth_dtype_policy():
    """create_masked_ply.py:91 converts whatever dtype np.load returned to float64; the kernel reads uint16 or float32.  A
    float64 image of float32 numbers is the same cloud and passes, anything that would be narrowed is refused."""
    import torch
    from repas_vision_b200.cloud import _depth_kind
    u = np.arange(12, dtype=np.uint16).reshape(3, 4)
    assert _depth_kind(u)[0] == "u16" and _depth_kind(u)[1] is u
    f = (u.astype(np.float32) * np.float32(0.001))
    assert _depth_kind(f)[0] == "f32"
    f64 = f.astype(np.float64)
    f64[0, 0] = np.nan
    k, d = _depth_kind(f64)
    assert k == "f32" and d.dtype == np.float32 and np.array_equal(d.astype(np.float64), f64, equal_nan=True)
    with pytest.raises(RuntimeError):
        _depth_kind(u.astype(np.float64) * 0.001)  # float64 products are not float32 numbers
    for bad in (np.int32, np.float16, np.uint8):
        with pytest.raises(RuntimeError):
            _depth_kind(u.astype(bad))
    assert _depth_kind(torch.from_numpy(f64))[0] == "f32"
    with pytest.raises(RuntimeError):
        _depth_kind(torch.from_numpy(u.astype(np.int32)))


def test_open3d_shaped_namespace_lists_the_path_and_names_what_it_leaves_out():
    """o3d_compat imports without a GPU; the vector constructors validate shapes like Open3D's; everything outside the
    point-cloud path raises an AttributeError that says so."""
    from repas_vision_b200 import o3d_compat as o3d
    import repas_vision_b200 as rv
    assert o3d.geometry.PointCloud is rv.PointCloud and o3d.io.read_point_cloud is rv.read_point_cloud
    assert o3d.pipelines.registration.registration_icp is rv.registration_icp
    v = o3d.utility.Vector3dVector([[1, 2, 3], [4, 5, 6]])
    assert v.dtype == np.float64 and v.shape == (2, 3) and o3d.utility.Vector3dVector().shape == (0, 3)
    assert o3d.utility.IntVector([1, 2]).dtype == np.int32 and o3d.utility.DoubleVector([0.5]).tolist() == [0.5]
    with pytest.raises(RuntimeError):
        o3d.utility.Vector3dVector(np.zeros((3, 4)))
    for expr in ("o3d.visualization", "o3d.geometry.TriangleMesh", "o3d.pipelines.registration.compute_fpfh_feature",
                 "o3d.io.read_triangle_mesh"):
        with pytest.raises(AttributeError, match="outside the point-cloud path"):
            eval(expr)
    R = o3d.geometry.get_rotation_matrix_from_xyz((0.0, 0.0, np.pi / 2))
    assert np.allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)
