"""SURVEY 8f-4, ICP: nearest-point correspondence search on the target's hash grid, the sums of the estimation step and
the registration_icp loop, through the C ABI against the CPU oracle (oracle_np.registration_icp, a restatement of Open3D
0.19 RegistrationICP; Open3D itself is absent: parity unpinned).

Bars: correspondences (integer indices) bit-exact; fitness exact; sums / inlier_rmse within 1e-11 relative (parallel sum
order); the final 4x4 within 1e-8 absolute of the oracle's after the same number of iterations."""
import numpy as np
import pytest

from synth import bumpy_surface

pytestmark = pytest.mark.gpu


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


def _nearest(rv, src, tgt, max_distance, dtype="f64"):
    from repas_vision_b200 import _ops
    s = rv.PointCloud.from_arrays(src, None, dtype=dtype)
    t = rv.PointCloud.from_arrays(tgt, None, dtype=dtype)
    idx = _ops.nn_index_build(t._data, len(t), max_distance)
    return _ops.nn_search(idx, len(t), s._data, len(s), max_distance)[:len(s)].cpu().numpy()


@pytest.mark.parametrize("max_distance", [0.002, 0.01, 0.05, 10.0])
def test_nearest_search_matches_oracle(rv, O, max_distance):
    rng = np.random.default_rng(5)
    tgt = bumpy_surface(rng, 6000)
    src = np.concatenate([bumpy_surface(rng, 2500, noise=0.002),
                          rng.uniform(-1.0, 1.0, (300, 3)) + (0, 0, 0.6),   # far away, mostly outside the target's box
                          tgt[:50]])                                         # exactly on target points: distance 0
    ref, fit, rmse = O.nearest_correspondences(src, tgt, max_distance)
    got = _nearest(rv, src, tgt, max_distance)
    assert np.array_equal(got, ref)
    assert (ref[-50:] == np.arange(50)).all()
    if max_distance >= 10.0:
        assert (ref >= 0).all()
    if max_distance <= 0.002:
        assert (ref[2500:2800] == -1).mean() > 0.9


def test_nearest_search_ties_bounds_and_float32(rv, O):
    # a lattice: queries at cell centres are equidistant from 8 points -> the lowest index wins, as in the oracle's argmin
    g = np.arange(6, dtype=np.float64) * 0.25
    tgt = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    rng = np.random.default_rng(1)
    tgt = tgt[rng.permutation(len(tgt))]
    q = tgt[:40] + 0.125
    ref, _, _ = O.nearest_correspondences(q, tgt, 1.0)
    got = _nearest(rv, q, tgt, 1.0)
    assert np.array_equal(got, ref)
    # strict bound: a target exactly max_distance away is not matched, one ulp closer is
    tgt1 = np.array([[0.0, 0.0, 0.0], [5.0, 5.0, 5.0]])
    q1 = np.array([[0.5, 0.0, 0.0], [np.nextafter(0.5, 0.0), 0.0, 0.0], [np.nan, 0.0, 0.0]])
    assert _nearest(rv, q1, tgt1, 0.5).tolist() == [-1, 0, -1]
    # float32 clouds are widened exactly
    t32 = bumpy_surface(rng, 3000).astype(np.float32)
    s32 = bumpy_surface(rng, 1000, noise=0.001).astype(np.float32)
    ref32, _, _ = O.nearest_correspondences(s32.astype(np.float64), t32.astype(np.float64), 0.01)
    assert np.array_equal(_nearest(rv, s32, t32, 0.01, dtype="f32"), ref32)


def test_nearest_search_with_far_outlier_in_target(rv, O):
    """One stray target point inflates the bounding box (the grid refinement / brute-force paths of the index)."""
    rng = np.random.default_rng(9)
    tgt = np.concatenate([bumpy_surface(rng, 4000), [[40.0, -30.0, 25.0]]])
    src = np.concatenate([bumpy_surface(rng, 800, noise=0.001), [[39.0, -30.0, 25.0], [-20.0, 3.0, 1.0]]])
    for md in (0.01, 2.0, 100.0):
        ref, _, _ = O.nearest_correspondences(src, tgt, md)
        assert np.array_equal(_nearest(rv, src, tgt, md), ref)


def test_icp_sums_match_numpy(rv, O):
    from repas_vision_b200 import _ops
    import torch
    rng = np.random.default_rng(2)
    tgt = bumpy_surface(rng, 5000)
    src = bumpy_surface(rng, 3000, noise=0.001)
    N = O.estimate_normals(tgt, 0.02, 30, camera_location=(0, 0, 0))
    near, fit, rmse = O.nearest_correspondences(src, tgt, 0.004)
    assert 0.2 < fit < 1.0
    s = rv.PointCloud.from_arrays(src, None)
    t = rv.PointCloud.from_arrays(tgt, None)
    nrm = torch.as_tensor(np.ascontiguousarray(N.T), device=s.device)
    dnear = torch.as_tensor(near, device=s.device)
    sel = near >= 0
    S, T, Nt = src[sel], tgt[near[sel]], N[near[sel]]
    plane = _ops.icp_sums(s._data, len(s), t._data, len(t), nrm, dnear, True).cpu().numpy()
    r = ((S - T) * Nt).sum(axis=1)
    J = np.concatenate([np.cross(S, Nt), Nt], axis=1)
    assert plane[0] == sel.sum()
    assert np.isclose(plane[1], ((S - T) ** 2).sum(), rtol=1e-11)
    assert np.isclose(plane[2], (r * r).sum(), rtol=1e-11)
    assert np.allclose(plane[3:9], J.T @ r, rtol=1e-9, atol=1e-15)
    assert np.allclose(plane[9:30], (J.T @ J)[np.triu_indices(6)], rtol=1e-11, atol=1e-15)
    point = _ops.icp_sums(s._data, len(s), t._data, len(t), None, dnear, False).cpu().numpy()
    assert point[0] == sel.sum() and np.isclose(point[1], plane[1], rtol=0, atol=0)
    assert np.isclose(point[2], (S * S).sum(), rtol=1e-11)
    assert np.allclose(point[3:6], S.sum(axis=0), rtol=1e-11) and np.allclose(point[6:9], T.sum(axis=0), rtol=1e-11)
    assert np.allclose(point[9:18].reshape(3, 3), T.T @ S, rtol=1e-11)
    # the sums are deterministic: a second launch gives the same bits
    again = _ops.icp_sums(s._data, len(s), t._data, len(t), nrm, dnear, True).cpu().numpy()
    assert np.array_equal(again, plane)


def _scene(O, seed=21, n_t=8000, n_s=3000):
    rng = np.random.default_rng(seed)
    tgt = bumpy_surface(rng, n_t)
    src = bumpy_surface(rng, n_s)
    D = O.vector6d_to_matrix4d(np.array([0.02, -0.015, 0.03, 0.004, -0.003, 0.005]))
    return O.transform(src, D), tgt, D


def test_registration_icp_point_to_plane_matches_oracle(rv, O):
    """refine_with_icp (mpa_icp_export.py:166-208): voxel-sized clouds, estimate_normals(Hybrid(0.02, 30)) on the target,
    point-to-plane ICP with the scripts' criteria."""
    moved, tgt, D = _scene(O)
    source = rv.PointCloud.from_arrays(moved, None)
    target = rv.PointCloud.from_arrays(tgt, None)
    target.estimate_normals(search_param=rv.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    N = target.normals
    crit = rv.ICPConvergenceCriteria(max_iteration=50, relative_fitness=1e-6, relative_rmse=1e-6)
    reg = rv.registration_icp(source, target, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(), crit)
    T, fit, rmse, near, it = O.registration_icp(moved, tgt, 0.02, target_normals=N, point_to_plane=True, max_iteration=50)
    assert it < 50
    assert reg.fitness == fit
    assert np.isclose(reg.inlier_rmse, rmse, rtol=1e-9)
    assert np.allclose(reg.transformation, T, rtol=0, atol=1e-8)
    cs = reg.correspondence_set
    sel = np.nonzero(near >= 0)[0]
    assert cs.dtype == np.int32 and np.array_equal(cs[:, 0], sel) and np.array_equal(cs[:, 1], near[sel])
    assert np.allclose(reg.transformation @ D, np.eye(4), atol=2e-3)  # and it is the right answer
    assert np.array_equal(source.points, moved)  # the source cloud is left untouched, like Open3D's const reference
    # evaluate_registration at the result reports the same figures
    ev = rv.evaluate_registration(source, target, 0.02, reg.transformation)
    assert np.isclose(ev.fitness, reg.fitness, atol=2.0 / len(moved)) and np.isclose(ev.inlier_rmse, reg.inlier_rmse, rtol=1e-3)


@pytest.mark.parametrize("with_scaling", [False, True])
def test_registration_icp_point_to_point_matches_oracle(rv, O, with_scaling):
    moved, tgt, D = _scene(O, seed=22, n_t=6000, n_s=2000)
    init = np.eye(4)
    init[:3, 3] = (-0.002, 0.001, -0.002)
    source = rv.PointCloud.from_arrays(moved, None)
    target = rv.PointCloud.from_arrays(tgt, None)
    crit = rv.ICPConvergenceCriteria(max_iteration=40)
    reg = rv.registration_icp(source, target, 0.02, init, rv.TransformationEstimationPointToPoint(with_scaling), crit)
    T, fit, rmse, near, it = O.registration_icp(moved, tgt, 0.02, init=init, point_to_plane=False, with_scaling=with_scaling,
                                                max_iteration=40)
    assert reg.fitness == fit and np.isclose(reg.inlier_rmse, rmse, rtol=1e-9)
    assert np.allclose(reg.transformation, T, rtol=0, atol=1e-8)
    assert len(reg.correspondence_set) == int((near >= 0).sum())


def test_registration_icp_errors_and_empty_inputs(rv, O):
    moved, tgt, _ = _scene(O, n_t=500, n_s=200)
    source = rv.PointCloud.from_arrays(moved, None)
    target = rv.PointCloud.from_arrays(tgt, None)
    with pytest.raises(RuntimeError, match="normal"):
        rv.registration_icp(source, target, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane())
    with pytest.raises(RuntimeError, match="max_correspondence_distance"):
        rv.registration_icp(source, target, 0.0)
    with pytest.raises(ValueError):
        rv.registration_icp(source, target, 0.02, np.eye(3))
    # no point within reach: identity updates, zero fitness, the initial transformation comes back
    far = rv.PointCloud.from_arrays(moved + 50.0, None)
    reg = rv.registration_icp(far, target, 0.02)
    assert reg.fitness == 0.0 and reg.inlier_rmse == 0.0 and np.array_equal(reg.transformation, np.eye(4))
    assert reg.correspondence_set.shape == (0, 2)
    empty = rv.PointCloud(None, 0, False)
    assert rv.registration_icp(empty, target, 0.02).fitness == 0.0
    assert rv.registration_icp(source, empty, 0.02).fitness == 0.0
    assert rv.evaluate_registration(source, target, -1.0).fitness == 0.0


def test_transform_carries_normals(rv, O):
    """Open3D PointCloud::Transform rotates the normals by the upper-left 3x3 (final_view_with_cad.py:333 on a cloud
    that went through estimate_normals)."""
    rng = np.random.default_rng(4)
    P = bumpy_surface(rng, 3000)
    pc = rv.PointCloud.from_arrays(P, None)
    pc.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30))
    N = pc.normals
    T = O.vector6d_to_matrix4d(np.array([0.3, -0.2, 0.5, 0.1, 0.2, -0.3]))
    pc.transform(T)
    assert pc.has_normals()
    R = np.eye(4)
    R[:3, :3] = T[:3, :3]
    assert np.array_equal(pc.normals, O.transform(N, R))
    assert np.array_equal(pc.points, O.transform(P, T))
    assert np.allclose(np.linalg.norm(pc.normals, axis=1), 1.0, atol=1e-12)


def test_registration_icp_float32_clouds(rv, O):
    """Clouds straight out of the float32 pipeline (deproject_batch(dtype="f32")): the working copy stays float32, so the result
    follows the float64 oracle to float32 rounding only."""
    moved, tgt, D = _scene(O, seed=23, n_t=5000, n_s=2000)
    source = rv.PointCloud.from_arrays(moved.astype(np.float32), None, dtype="f32")
    target = rv.PointCloud.from_arrays(tgt.astype(np.float32), None, dtype="f32")
    target.estimate_normals(rv.KDTreeSearchParamHybrid(0.02, 30))
    reg = rv.registration_icp(source, target, 0.02, np.eye(4), rv.TransformationEstimationPointToPlane(),
                              rv.ICPConvergenceCriteria(max_iteration=50))
    T, fit, rmse, _, _ = O.registration_icp(source.points, target.points, 0.02, target_normals=target.normals, max_iteration=50)
    assert abs(reg.fitness - fit) < 5e-3 and abs(reg.inlier_rmse - rmse) < 1e-5
    assert np.allclose(reg.transformation, T, atol=2e-4) and np.allclose(reg.transformation @ D, np.eye(4), atol=3e-3)
    mixed = rv.registration_icp(rv.PointCloud.from_arrays(moved, None), target, 0.02, np.eye(4),
                                rv.TransformationEstimationPointToPlane(), rv.ICPConvergenceCriteria(max_iteration=50))
    assert np.allclose(mixed.transformation, T, atol=2e-4)


def test_device_loop_equals_host_loop(rv, O):
    """Point-to-plane ICP with the loop's decisions taken on the device (rv_icp_iterate: stopping rule, 6x6 solve, pose
    composition) against the same loop with those steps on the host: same number of estimation steps, same correspondences,
    transformation to rounding, for several caps on the iteration count (0 = evaluate only; caps that fall inside and across
    the read-back batches) and a non-identity initial alignment."""
    moved, tgt, D = _scene(O, seed=33)
    target = rv.PointCloud.from_arrays(tgt, None)
    target.estimate_normals(search_param=rv.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    est = rv.TransformationEstimationPointToPlane()
    init = O.vector6d_to_matrix4d(np.array([0.002, 0.001, -0.003, 0.0005, 0.0, -0.0004]))
    for dtype in ("f64", "f32"):
        source = rv.PointCloud.from_arrays(moved.astype(np.float32) if dtype == "f32" else moved, None, dtype=dtype)
        tg = rv.PointCloud.from_arrays(tgt.astype(np.float32) if dtype == "f32" else tgt, None, dtype=dtype)
        tg._normals = target._normals
        for cap, T0 in ((0, np.eye(4)), (1, np.eye(4)), (3, init), (4, np.eye(4)), (5, init), (8, np.eye(4)), (9, init), (50, np.eye(4)),
                        (50, init)):
            crit = rv.ICPConvergenceCriteria(max_iteration=cap, relative_fitness=1e-6, relative_rmse=1e-6)
            a = rv.registration_icp(source, tg, 0.02, T0, est, crit)
            b = rv.registration_icp(source, tg, 0.02, T0, est, crit, device_loop=False)
            assert a.iterations == b.iterations and a.iterations <= cap
            assert a.fitness == b.fitness and np.isclose(a.inlier_rmse, b.inlier_rmse, rtol=1e-9)
            assert np.allclose(a.transformation, b.transformation, rtol=0, atol=1e-9 if dtype == "f64" else 1e-6)
            if dtype == "f64":
                assert np.array_equal(a.correspondence_set, b.correspondence_set)
        assert a.iterations < 50  # converged by the stopping rule, not by the cap


def test_device_loop_survives_a_stale_match_buffer(rv, O):
    """rv_icp_iterate starts every search after the first from the match the buffer holds (the previous evaluation's, by
    contract).  A caller that breaks the contract -- a buffer of arbitrary integers, some far out of range -- still gets the
    exact nearest points: the stale entry is only ever a candidate whose distance is measured."""
    import torch
    from repas_vision_b200 import _ops
    moved, tgt, D = _scene(O, seed=41, n_t=6000, n_s=2500)
    source = rv.PointCloud.from_arrays(moved, None)
    target = rv.PointCloud.from_arrays(tgt, None)
    target.estimate_normals(search_param=rv.KDTreeSearchParamHybrid(radius=0.02, max_nn=30))
    n, nt, dev = len(source), len(target), source.device
    index = _ops.nn_index_build(target._data, nt, 0.02)
    scratch = _ops.icp_scratch(dev)
    results = []
    rng = np.random.default_rng(3)
    for garbage in (None, rng.integers(-2**31, 2**31 - 1, n), rng.integers(0, nt, n), np.full(n, nt), np.full(n, -1)):
        work = source._data.clone()
        nearest = torch.empty(n, dtype=torch.int32, device=dev)
        state = _ops.icp_state(dev)
        _ops.icp_begin(state, np.eye(4), 3, 1e-6, 1e-6, n)
        _ops.icp_iterate(state, True, 1, work, n, index, target._data, nt, target._normals, 0.02, nearest, scratch)
        if garbage is not None:
            nearest.copy_(torch.from_numpy(garbage.astype(np.int64).astype(np.int32)).to(dev))
        _ops.icp_iterate(state, False, 2, work, n, index, target._data, nt, target._normals, 0.02, nearest, scratch)
        results.append((state.cpu().numpy().copy(), nearest.cpu().numpy().copy(), work.cpu().numpy().copy()))
    for st, near, work in results[1:]:
        assert np.array_equal(st[:40], results[0][0][:40])
        assert np.array_equal(near, results[0][1]) and np.array_equal(work, results[0][2])
    # and they are the oracle's nearest points of the final working copy
    ref, _, _ = O.nearest_correspondences(results[0][2][:3, :n].T.copy(), target.points, 0.02)
    assert np.array_equal(results[0][1], ref)
