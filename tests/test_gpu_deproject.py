"""K1 parity on the B200: deprojection + masks + stream compaction through the C ABI against the CPU oracle
and against the golden vectors the reference's own functions produced (tests/golden/make_golden.py).

Bars: valid mask, point order and float64 clouds bit-exact; float32 clouds bit-exact against the oracle's
float32 contract and within 1e-5 m of the float64 reference values."""
import functools

import numpy as np
import pytest

from conftest import CANOPY_TS, blob_mask, load_frame, sha
from synth import synth_batch, synth_color, synth_depth, synth_mask

pytestmark = pytest.mark.gpu

XYZ_TOL_M = 1e-5  # BASELINE.json north_star: within 1e-5 m absolute on xyz


@pytest.fixture(scope="module")
def rv():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import repas_vision_b200 as rv
    return rv


@pytest.fixture(scope="module")
def O():
    from oracle import oracle_np
    return oracle_np


@functools.lru_cache(maxsize=None)
def _dataset(B, H, W):
    depth, bgr = synth_batch(B, H, W, seed0=1)
    return depth, bgr, np.stack([synth_mask(H, W, 50 + i, cover=0.3) for i in range(B)])


def _cloud_np(pc):
    return pc.points, pc.colors


@pytest.mark.parametrize("idx", range(5))
def test_real_frames_float64_match_reference_goldens(rv, golden, rs720, idx):
    """create_masked_pointcloud on the reference's captured frames: float64 points/colours hash-identical to what
    femto_bolt_code/scripts/create_masked_ply.py:56-107 produced, plus the fused distance mask."""
    ts = CANOPY_TS[idx]
    color, depth = load_frame(ts)
    _, dm, _ = rv.depth_to_meters(depth)
    for rec in [r for r in golden["masked_cloud"] if r["ts"] == ts]:
        h, w = depth.shape
        mask = np.full((h, w), 255, np.uint8) if rec["variant"] == "all" else blob_mask(h, w, rec["mask_seed"])
        pc = rv.create_masked_pointcloud(color, dm, mask, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"],
                                         invert_mask=rec["variant"] == "blob_inverted")
        P, C = _cloud_np(pc)
        assert len(pc) == rec["n"]
        assert sha(P) == rec["points"]["sha256"]
        assert sha(C) == rec["colors"]["sha256"]
        # raw uint16 depth in, unit conversion fused: same cloud
        pc16 = rv.create_masked_pointcloud(color, depth, mask, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"],
                                           invert_mask=rec["variant"] == "blob_inverted")
        assert sha(pc16.points) == rec["points"]["sha256"]
        if rec["variant"] == "all":
            near = rv.create_masked_pointcloud(color, dm, mask, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"],
                                               max_distance=1.0)
            assert len(near) == rec["dist_lt_1m"]["kept"]
            assert sha(near.points) == rec["dist_lt_1m"]["points"]["sha256"]
            # the stand-alone filters on the produced cloud (distance_masking_on_ply.py, view_point_cloud.py, AABB crop)
            assert sha(pc.select_within_distance(1.0).points) == rec["dist_lt_1m"]["points"]["sha256"]
            assert len(pc.clip_z(0.15, 8.0)) == rec["zclip_0p15_8"]["kept"]
            assert len(pc.crop_aabb(rec["aabb"]["min"], rec["aabb"]["max"])) == rec["aabb"]["kept"]


def test_float_depth_nan_inf_negative(rv, golden):
    rngf = np.random.default_rng(11)
    dm = (rngf.uniform(0.2, 4.0, (48, 64))).astype(np.float32)
    dm[rngf.random((48, 64)) < 0.1] = np.nan
    dm[rngf.random((48, 64)) < 0.05] = np.inf
    dm[rngf.random((48, 64)) < 0.05] = -1.0
    dm[rngf.random((48, 64)) < 0.1] = 0.0
    col = rngf.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    msk = (rngf.random((48, 64)) < 0.7).astype(np.uint8) * 255
    pc = rv.create_masked_pointcloud(col, dm, msk, 60.0, 61.0, 31.5, 23.25)
    g = golden["masked_cloud_float_small"]
    assert len(pc) == g["n"] and sha(pc.points) == g["points"]["sha256"] and sha(pc.colors) == g["colors"]["sha256"]


CASES = [
    dict(),
    dict(r_max=1.0),
    dict(z_clip=(0.15, 8.0)),
    dict(z_clip=(None, 2.0), r_max=2.5),
    dict(aabb=((-0.3, -0.25, 0.5), (0.35, 0.2, 1.2))),
    dict(depth_trunc=3.0, unit_rule="div_f32"),
    dict(unit_rule="div_f64", r_max=1.5),
    dict(mask=True),
    dict(mask=True, invert_mask=True, r_max=1.2),
    dict(color_scale="255"),
]


@pytest.mark.parametrize("shape,kernel", [((720, 1280), "tma"), ((720, 1280), "generic"), ((480, 640), "tma"),
                                          ((480, 640), "generic"), ((36, 48), "tma"), ((37, 53), "auto"), ((1, 2049), "auto")])
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_synthetic_frames_all_predicates(rv, O, rs720, shape, kernel, case, dtype):
    """Every predicate combination through BOTH K1 kernels (TMA-fed pipeline and generic), ragged shapes included
    (tile tails, odd widths; those only run the generic kernel)."""
    import torch
    H, W = shape
    kw = dict(CASES[case])
    use_mask = kw.pop("mask", False)
    B = 2 if H * W > 10000 else 5
    depth, bgr, allmask = _dataset(B, H, W)
    mask = allmask if use_mask else None
    cam = rv.Camera(rs720["fx"] * W / 1280, rs720["fy"] * H / 720, rs720["cx"] * W / 1280, rs720["cy"] * H / 720, W, H)
    okw = dict(kw)
    if okw.get("z_clip") is not None:
        okw["z_clip"] = tuple(v if v is not None else s * np.inf for v, s in zip(okw["z_clip"], (-1.0, 1.0)))
    gk = dict(unit_rule=kw.get("unit_rule", "mul_f32"), invert_mask=kw.get("invert_mask", False),
              depth_trunc=kw.get("depth_trunc"), max_distance=kw.get("r_max"), z_clip=kw.get("z_clip"), aabb=kw.get("aabb"),
              color_scale=kw.get("color_scale", "unit"))
    batch = rv.deproject_batch(torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda(), cam,
                               None if mask is None else torch.from_numpy(mask).cuda(), dtype=dtype, want_valid=True,
                               want_src_index=True, kernel=kernel, **gk)
    counts = batch.counts_host()
    for b in range(B):
        ref = O.deproject_mask(depth[b], bgr[b], None if mask is None else mask[b], fx=cam.fx, fy=cam.fy, cx=cam.cx,
                               cy=cam.cy, out_dtype=dtype, **okw)
        n = ref["points"].shape[0]
        assert counts[b] == n
        assert np.array_equal(batch.valid[b].cpu().numpy().astype(bool), ref["valid"])  # bit-exact mask
        f = batch.frame(b)
        got = f.xyz.t().cpu().numpy()
        assert np.array_equal(batch.src_index[b * batch.cap:b * batch.cap + n].cpu().numpy(), ref["src_index"])
        assert got.dtype == ref["points"].dtype and np.array_equal(got, ref["points"])  # bit-exact, same order
        assert np.array_equal(f.rgb.t().cpu().numpy(), ref["colors"])
        ref64 = O.deproject_mask(depth[b], bgr[b], None if mask is None else mask[b], fx=cam.fx, fy=cam.fy, cx=cam.cx,
                                 cy=cam.cy, out_dtype="f64", **{k: v for k, v in okw.items() if k in ("unit_rule", "invert_mask", "depth_trunc")})
        # tolerance bar against the float64 values (same pixels selected by src_index)
        sel = np.isin(ref64["src_index"], ref["src_index"])
        assert np.abs(got.astype(np.float64) - ref64["points"][sel]).max(initial=0.0) <= XYZ_TOL_M


@pytest.mark.parametrize("mode,kernel", [("compact_unordered", "auto"), ("compact_unordered", "generic"), ("compact_unordered", "tma"),
                                         ("dense_zero", "tma"), ("dense_zero", "generic"),
                                         ("dense_nan", "tma"), ("dense_nan", "generic")])
def test_other_output_modes(rv, O, rs720, mode, kernel):
    import torch
    H, W, B = 480, 640, 4
    depth, bgr = synth_batch(B, H, W, seed0=21)
    cam = rv.Camera(608.2335815429688, 607.8508911132812, 312.52239990234375, 232.65150451660156, W, H)
    batch = rv.deproject_batch(torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda(), cam, max_distance=1.0,
                               unit_rule="div_f32", mode=mode, want_src_index=True, kernel=kernel)
    counts = batch.counts_host()
    for b in range(B):
        ref = O.deproject_mask(depth[b], bgr[b], None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, r_max=1.0,
                               unit_rule="div_f32", out_dtype="f32")
        n = ref["points"].shape[0]
        assert counts[b] == n
        blk = batch.data[:, b * batch.cap:(b + 1) * batch.cap].cpu().numpy()
        src = batch.src_index[b * batch.cap:(b + 1) * batch.cap].cpu().numpy()
        if mode == "compact_unordered":
            order = np.argsort(src[:n], kind="stable")  # compare as a set keyed by source pixel
            assert np.array_equal(src[:n][order], ref["src_index"])
            assert np.array_equal(blk[:3, :n].T[order], ref["points"])
            assert np.array_equal(blk[3:, :n].T[order], ref["colors"])
        else:
            flat = ref["valid"].reshape(-1)
            assert np.array_equal(src >= 0, flat)
            assert np.array_equal(blk[:3][:, flat].T, ref["points"])
            assert np.array_equal(blk[3:][:, flat].T, ref["colors"])
            inv = blk[:3][:, ~flat]
            assert np.isnan(inv).all() if mode == "dense_nan" else (inv == 0).all()
            assert (blk[3:][:, ~flat] == 0).all()


@pytest.mark.parametrize("kernel", ["generic", "tma"])
@pytest.mark.parametrize("model", ["inverse_brown_conrady", "brown_conrady"])
def test_distorted_camera_ray_table(rv, O, model, kernel):
    """Checkerboard calibration with lens distortion (realtime_pose_estimation_april_tag.py:10-18): the float64 ray table
    reproduces rs2_deproject_pixel_to_point and the kernel multiplies by it."""
    import torch
    H, W = 480, 640
    dist = (0.04344582, 0.32076285, -0.00060687, -0.0004814, -1.40593456)
    cam = rv.Camera(605.2845686, 605.44233933, 309.95995203, 229.79166863, W, H, dist, model)
    depth, bgr = synth_batch(2, H, W, seed0=31)
    rays = O.ray_table(cam.fx, cam.fy, cam.cx, cam.cy, dist, model, W, H)
    from repas_vision_b200 import _ops
    got_rays = _ops.ray_table(cam, H, W, torch.device("cuda", 0)).cpu().numpy()
    assert np.array_equal(got_rays, rays)
    for dtype in ("f32", "f64"):
        batch = rv.deproject_batch(torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda(), cam, max_distance=1.5,
                                   dtype=dtype, kernel=kernel)
        for b in range(2):
            ref = O.deproject_mask(depth[b], bgr[b], None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, r_max=1.5,
                                   out_dtype=dtype, rays=rays)
            assert batch.counts_host()[b] == ref["points"].shape[0]
            assert np.array_equal(batch.frame(b).xyz.t().cpu().numpy(), ref["points"])
            assert np.array_equal(batch.frame(b).rgb.t().cpu().numpy(), ref["colors"])


def test_create_from_rgbd_image_open3d_shape(rv, O, rs720):
    color, depth = load_frame(CANOPY_TS[0])
    pc = rv.create_from_rgbd_image(color, depth, rs720, depth_scale=1000.0, depth_trunc=3.0)
    P, C = O.create_from_rgbd_image(color, depth, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], 1000.0, 3.0)
    assert len(pc) == P.shape[0] and np.array_equal(pc.points, P) and np.array_equal(pc.colors, C)
    E = np.eye(4)
    E[:3, 3] = [0.1, -0.2, 0.05]
    E[:3, :3] = [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]
    pc2 = rv.create_from_rgbd_image(color, depth, rs720, extrinsic=E)
    P2, _ = O.create_from_rgbd_image(color, depth, rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], 1000.0, 3.0, extrinsic=E)
    assert np.abs(pc2.points - P2).max() <= 1e-12
    dense = rv.create_from_rgbd_image(color, depth, rs720, project_valid_depth_only=False)
    assert len(dense) == 720 * 1280 and np.isnan(dense.points[:, 0]).sum() == 720 * 1280 - P.shape[0]


def test_depth_to_meters_goldens(rv, golden):
    _, depth = load_frame(CANOPY_TS[0])
    raw, dm, scale = rv.depth_to_meters(depth)
    assert scale == golden["depth_to_meters"]["scale"] and raw is depth or np.array_equal(raw, depth)
    assert dm.dtype == np.float32 and sha(dm) == golden["depth_to_meters"]["depth_m"]["sha256"]
    allv = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    assert sha(rv.depth_to_meters(allv)[1]) == golden["depth_to_meters"]["all_u16"]["sha256"]
    from oracle import oracle_np as O
    for rule in ("div_f32", "div_f64"):
        ref = O.depth_to_meters(allv, rule).astype(np.float32)
        assert np.array_equal(rv.depth_to_meters(allv, rule)[1], ref)


def test_full_size_batch_properties(rv, rs720):
    """BASELINE full-size shapes through size-independent properties: counts equal the valid-mask population, ordered
    output is strictly increasing in source index, dense and compact agree, and a second run is identical."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(1234)
    B, H, W = 64, 720, 1280
    depth = torch.randint(0, 4000, (B, H, W), generator=g, device="cuda", dtype=torch.int32)
    depth = torch.where(depth < 1500, torch.zeros_like(depth), depth).to(torch.uint16)
    bgr = torch.randint(0, 256, (B, H, W, 3), generator=g, device="cuda", dtype=torch.uint8)
    cam = rv.Camera(rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], W, H)
    a = rv.deproject_batch(depth, bgr, cam, max_distance=2.0, want_valid=True, want_src_index=True)
    counts = a.counts.cpu()
    assert torch.equal(counts, a.valid.view(B, -1).sum(dim=1, dtype=torch.int64).cpu())
    for b in (0, B // 2, B - 1):
        n = int(counts[b])
        s = a.src_index[b * a.cap:b * a.cap + n]
        assert bool((s[1:] > s[:-1]).all())
        assert torch.equal(torch.nonzero(a.valid[b].view(-1)).view(-1).to(torch.int32), s)
    d = rv.deproject_batch(depth, bgr, cam, max_distance=2.0, mode="dense_zero")
    for b in (0, B - 1):
        n = int(counts[b])
        idx = a.src_index[b * a.cap:b * a.cap + n].long()
        assert torch.equal(d.data[:, b * d.cap:(b + 1) * d.cap][:, idx], a.data[:, b * a.cap:b * a.cap + n])
    a2 = rv.deproject_batch(depth, bgr, cam, max_distance=2.0)
    assert torch.equal(a2.counts.cpu(), counts)
    for b in (1, B - 2):
        n = int(counts[b])
        assert torch.equal(a2.data[:, b * a.cap:b * a.cap + n], a.data[:, b * a.cap:b * a.cap + n])
    # byte colours: the same points, and the bytes are the float colours times 255; NV12 in: counts and coordinates do not
    # depend on the colour format, every colour byte of a kept point is written
    pk = rv.deproject_batch(depth, bgr, cam, max_distance=2.0, color_scale="packed8")
    nv12 = torch.randint(0, 256, (B, H * 3 // 2, W), generator=g, device="cuda", dtype=torch.int32).to(torch.uint8)
    nv = rv.deproject_batch(depth, nv12, cam, max_distance=2.0, color_format="nv12", color_scale="packed8")
    assert torch.equal(pk.counts.cpu(), counts) and torch.equal(nv.counts.cpu(), counts)
    for b in (0, B // 3, B - 1):
        n = int(counts[b])
        assert torch.equal(pk.data[:3, b * pk.cap:b * pk.cap + n], a.data[:3, b * a.cap:b * a.cap + n])
        assert torch.equal(nv.data[:3, b * nv.cap:b * nv.cap + n], a.data[:3, b * a.cap:b * a.cap + n])
        rgb8 = pk.rgb8(b).to(torch.float32)
        assert torch.equal(torch.round(a.data[3:, b * a.cap:b * a.cap + n].t() * 255.0), rgb8)
        idx = a.src_index[b * a.cap:b * a.cap + n].long()
        assert torch.equal(pk.rgb8(b), bgr[b].view(-1, 3)[idx].flip(1))


def test_edge_cases_and_errors(rv, rs720):
    import torch
    cam = rv.Camera(500.0, 500.0, 3.5, 2.5, 8, 6)
    z = np.zeros((6, 8), np.uint16)
    c = np.zeros((6, 8, 3), np.uint8)
    pc = rv.create_masked_pointcloud(c, z, np.full((6, 8), 255, np.uint8), 500.0, 500.0, 3.5, 2.5)
    assert len(pc) == 0 and pc.is_empty() and pc.points.shape == (0, 3)
    full = np.full((6, 8), 65535, np.uint16)
    pc = rv.create_masked_pointcloud(c, full, np.full((6, 8), 255, np.uint8), 500.0, 500.0, 3.5, 2.5, dtype="f32")
    assert len(pc) == 48 and np.allclose(pc.points[:, 2], np.float32(65535) * np.float32(0.001))
    with pytest.raises(RuntimeError):
        rv.create_masked_pointcloud(c[:5], z, None, 500.0, 500.0, 3.5, 2.5)
    with pytest.raises(RuntimeError):
        rv.create_masked_pointcloud(c, z, np.zeros((5, 8), np.uint8), 500.0, 500.0, 3.5, 2.5)
    # capacity: counts stay truthful when the per-frame capacity is too small
    b = rv.deproject_batch(torch.from_numpy(full[None]).cuda(), torch.from_numpy(c[None]).cuda(), cam, frame_capacity=10)
    assert int(b.counts[0]) == 48
    from repas_vision_b200 import _lib
    ctx = _lib.context(0)
    st = ctx.lib.rv_deproject_mask(ctx.handle, None, None, None, None, 1, 6, 8, None, None, 0, 0, None, None, None, None, 0, None)
    assert st == _lib.RV_EINVAL and b"params" in ctx.lib.rv_last_error(ctx.handle)


@pytest.mark.parametrize("kernel", ["tma", "generic"])
def test_packed_mode_and_host_pipeline(rv, O, rs720, kernel):
    """COMPACT_PACKED (frames back to back, B+1 offsets) and the host-buffer pipeline built on it: numpy in, host clouds out."""
    import torch
    from repas_vision_b200 import _ops
    from repas_vision_b200.pipeline import HostPipeline
    H, W, B = 480, 640, 7
    depth, bgr = synth_batch(B, H, W, seed0=61)
    cam = rv.Camera(608.2335815429688, 607.8508911132812, 312.52239990234375, 232.65150451660156, W, H)
    refs = [O.deproject_mask(depth[b], bgr[b], None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, r_max=1.3, out_dtype="f32")
            for b in range(B)]
    r = _ops.deproject(torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda(), None, cam, depth_kind="u16", r_max=1.3,
                       mode="compact_packed", kernel=kernel)
    off = r["counts"].cpu().numpy()
    assert off[0] == 0 and np.array_equal(np.diff(off), [x["points"].shape[0] for x in refs])
    data = r["data"].cpu().numpy()
    for b in range(B):
        assert np.array_equal(data[:3, off[b]:off[b + 1]].T, refs[b]["points"])
        assert np.array_equal(data[3:, off[b]:off[b + 1]].T, refs[b]["colors"])
    if kernel == "tma":
        pipe = HostPipeline(cam, H, W, max_distance=1.3, chunk_frames=3, colors="float")
        res = pipe.run(depth, bgr)
        assert np.array_equal(res.counts, np.diff(off)) and res.h2d_bytes == B * H * W * 5
        for b in range(B):
            xyz, rgb = res.frame(b)
            assert np.array_equal(xyz.T, refs[b]["points"]) and np.array_equal(rgb.T, refs[b]["colors"])
        assert np.array_equal(res.points(2), refs[2]["points"].astype(np.float64))
        # default colour transport: the bytes themselves, 16 bytes per point; expanded on the host to the reference's float64
        pipe8 = HostPipeline(cam, H, W, max_distance=1.3, chunk_frames=3)
        res8 = pipe8.run(depth, bgr)
        assert res8.d2h_bytes < res.d2h_bytes and np.array_equal(res8.counts, res.counts)
        for b in range(B):
            xyz, rgb8 = res8.frame(b)
            keep = refs[b]["valid"]
            assert rgb8.dtype == np.uint8 and np.array_equal(rgb8, bgr[b][keep][:, ::-1])
            assert np.array_equal(xyz.T, refs[b]["points"])
            assert np.array_equal(res8.colors(b), bgr[b][keep][:, ::-1].astype(np.float64) / 255.0)
        # NV12 colour frames (the capture script's preferred format) read by the kernel itself: same clouds as feeding the
        # BGR image cv2 makes of them
        import cv2
        rng = np.random.default_rng(7)
        nv12 = rng.integers(0, 256, (B, H * 3 // 2, W), dtype=np.uint8)
        bgr_cv = np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12) for f in nv12])
        a = HostPipeline(cam, H, W, max_distance=1.3, chunk_frames=4, color_format="nv12", colors="float").run(depth, nv12)
        b2 = pipe.run(depth, bgr_cv)
        assert a.h2d_bytes == B * H * W * 7 // 2 and np.array_equal(a.counts, b2.counts)
        for fb in range(B):
            assert np.array_equal(a.frame(fb)[0], b2.frame(fb)[0]) and np.array_equal(a.frame(fb)[1], b2.frame(fb)[1])
        a8 = HostPipeline(cam, H, W, max_distance=1.3, chunk_frames=4, color_format="nv12").run(depth, nv12)
        for fb in range(B):
            assert np.array_equal(a8.frame(fb)[0], b2.frame(fb)[0])
            assert np.array_equal(a8.frame(fb)[1].astype(np.float32) / np.float32(255.0), b2.frame(fb)[1].T)
            keep = O.deproject_mask(depth[fb], bgr_cv[fb], None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, r_max=1.3, out_dtype="f32")["valid"]
            assert np.array_equal(a8.colors(fb), bgr_cv[fb][keep][:, ::-1].astype(np.float64) / 255.0)
        res2 = pipe.run(torch.from_numpy(depth).pin_memory(), torch.from_numpy(bgr).pin_memory())  # pinned inputs, reuse
        assert np.array_equal(res2.counts, res.counts) and np.array_equal(res2.frame(B - 1)[0], res.frame(B - 1)[0])
        # results own their pinned memory: a later run must not overwrite arrays handed out earlier (only release() recycles)
        keep_xyz = res.frame(0)[0]
        snapshot = keep_xyz.copy()
        del res
        for _ in range(3):
            pipe.run(depth, bgr_cv)
        assert np.array_equal(keep_xyz, snapshot)
        res2.release()
        with pytest.raises(RuntimeError):
            res2.frame(0)
        # frames -> PLY files (the capture loop's save_point_cloud_to_ply, better_three_capture.py:242): read back by the product's
        # reader and by the independent minimal reader of the oracle
        import os
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            for res_x in (res8, pipe.run(depth, bgr)):
                path = os.path.join(tmp, "frame3.ply")
                assert res_x.write_ply(3, path)
                back = rv.read_point_cloud(path)
                assert np.array_equal(back.points, refs[3]["points"].astype(np.float64))
                assert np.array_equal(back.colors, bgr[3][refs[3]["valid"]][:, ::-1].astype(np.float64) / 255.0)
                _, arr = O.read_ply_minimal(path)
                assert np.array_equal(np.stack([arr["x"], arr["y"], arr["z"]], 1), refs[3]["points"])
                assert np.array_equal(np.stack([arr["red"], arr["green"], arr["blue"]], 1), bgr[3][refs[3]["valid"]][:, ::-1])
        # the copy-only probe replays the schedule of a finished run
        pr = pipe8.copy_probe(depth, bgr, like=res8)
        assert pr.h2d_bytes == res8.h2d_bytes and pr.d2h_bytes == res8.d2h_bytes


@pytest.mark.parametrize("shape,kernel", [((720, 1280), "tma"), ((720, 1280), "generic"), ((480, 640), "tma"), ((480, 640), "generic"),
                                          ((36, 48), "auto"), ((38, 54), "auto"), ((2, 2050), "auto")])
@pytest.mark.parametrize("case", [dict(), dict(r_max=1.0), dict(r_max=1.2, unit_rule="div_f32"),
                                  dict(z_clip=(0.3, 2.0), mask=True), dict(mode="dense_zero")])
def test_nv12_frames_and_packed_colours(rv, O, rs720, shape, kernel, case):
    """Colour as the camera delivers it (NV12, better_three_capture.py:101-106,159) read by K1 itself, and colours kept as
    bytes on the way out (RV_COLOR_PACKED8): both must give exactly the cloud of the cv2-decoded BGR image / the float
    colour planes, on both kernels, including partial tiles, rows shorter than a tile and masks."""
    import cv2
    import torch
    H, W = shape
    B = 2 if H * W > 10000 else 4
    kw = dict(case)
    use_mask = kw.pop("mask", False)
    mode = kw.pop("mode", "compact_ordered")
    depth, _, allmask = _dataset(B, H, W)
    rng = np.random.default_rng(1000 + H)
    nv12 = rng.integers(0, 256, (B, H * 3 // 2, W), dtype=np.uint8)
    bgr_cv = np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12) for f in nv12])
    cam = rv.Camera(rs720["fx"] * W / 1280, rs720["fy"] * H / 720, rs720["cx"] * W / 1280, rs720["cy"] * H / 720, W, H)
    gk = dict(unit_rule=kw.get("unit_rule", "mul_f32"), max_distance=kw.get("r_max"), z_clip=kw.get("z_clip"), mode=mode,
              kernel=kernel, want_src_index=True)
    d = torch.from_numpy(depth).cuda()
    m = torch.from_numpy(allmask).cuda() if use_mask else None
    ref = rv.deproject_batch(d, torch.from_numpy(bgr_cv).cuda(), cam, m, **gk)                       # BGR in, float colours
    got = rv.deproject_batch(d, torch.from_numpy(nv12).cuda(), cam, m, color_format="nv12", **gk)     # NV12 in, float colours
    pk = rv.deproject_batch(d, torch.from_numpy(nv12).cuda(), cam, m, color_format="nv12", color_scale="packed8", **gk)
    pk_bgr = rv.deproject_batch(d, torch.from_numpy(bgr_cv).cuda(), cam, m, color_scale="packed8", **gk)
    counts = ref.counts_host()
    assert np.array_equal(got.counts_host(), counts) and np.array_equal(pk.counts_host(), counts)
    assert pk.data.shape[0] == 4 and got.data.shape[0] == 6
    for b in range(B):
        o = O.deproject_mask(depth[b], bgr_cv[b], allmask[b] if use_mask else None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy,
                             out_dtype="f32", **{k: v for k, v in kw.items()})
        n = H * W if mode.startswith("dense") else int(counts[b])
        assert int(counts[b]) == o["points"].shape[0]
        lo = b * ref.cap
        assert torch.equal(got.data[:, lo:lo + n], ref.data[:, lo:lo + n])
        assert torch.equal(pk.data[:3, lo:lo + n], ref.data[:3, lo:lo + n])
        assert torch.equal(pk.data[3, lo:lo + n], pk_bgr.data[3, lo:lo + n])
        rgb8 = pk.rgb8(b).cpu().numpy()
        src = ref.src_index[lo:lo + n].cpu().numpy()
        flat = bgr_cv[b].reshape(-1, 3)
        if mode.startswith("dense"):
            ok = src >= 0
            assert np.array_equal(ok, o["valid"].reshape(-1))
            assert np.array_equal(rgb8[ok], flat[ok][:, ::-1]) and not rgb8[~ok].any()
            assert np.array_equal(pk.data[3, lo:lo + n].view(torch.int32).cpu().numpy()[~ok], np.zeros((~ok).sum(), np.int32))
        else:
            assert np.array_equal(src, o["src_index"])
            assert np.array_equal(rgb8, flat[src][:, ::-1])
            assert np.array_equal(ref.data[3:, lo:lo + n].t().cpu().numpy(), o["colors"])
            assert np.array_equal(ref.data[:3, lo:lo + n].t().cpu().numpy(), o["points"])
        # the fourth byte of every colour word is zero
        assert not pk.data[3, lo:lo + n].view(torch.uint8).view(-1, 4)[:, 3].any()
    with pytest.raises(ValueError):
        rv.deproject_batch(d, torch.from_numpy(nv12).cuda(), cam, color_format="nv12", color_scale="packed8", dtype="f64")
    with pytest.raises(RuntimeError):
        pk.frame(0)


@pytest.mark.parametrize("mode", ["dense_zero", "compact_ordered"])
def test_sdk_float32_geometry(rv, O, rs720, mode):
    """the SDK clouds (rs.pointcloud, PointCloudFilter) evaluate z * ((u - ppx) / fx) in float32.
    geometry="sdk_f32" reproduces that arithmetic value for value (against the oracle's restatement of
    rs2_deproject_pixel_to_point); it stays within one float32 ulp of the reference's float64-then-round form."""
    import torch
    color, depth = load_frame(CANOPY_TS[1])
    H, W = depth.shape
    cam = rv.Camera(rs720["fx"], rs720["fy"], rs720["cx"], rs720["cy"], W, H)
    d = torch.from_numpy(depth[None]).cuda()
    c = torch.from_numpy(color[None]).cuda()
    kw = dict(max_distance=3.0, mode=mode, color_scale="255", want_src_index=True)
    sdk = rv.deproject_batch(d, c, cam, geometry="sdk_f32", **kw)
    ref = rv.deproject_batch(d, c, cam, **kw)
    o = O.deproject_mask(depth, color, None, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, r_max=3.0, out_dtype="f32", color_scale="255",
                         geometry="sdk_f32")
    # the oracle's array form equals the scalar restatement of rs2_deproject_pixel_to_point
    v, u = np.nonzero(o["valid"])
    z = depth[o["valid"]].astype(np.float32) * np.float32(0.001)
    X, Y, Z = O._rs_deproject_f32(u.astype(np.float32), v.astype(np.float32), z, dict(fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy))
    assert np.array_equal(o["points"], np.stack([X, Y, Z], 1))
    n = int(sdk.counts_host()[0])
    assert n == o["points"].shape[0]
    if mode == "dense_zero":
        flat = o["valid"].reshape(-1)
        got = sdk.data[:, :H * W].cpu().numpy()
        assert np.array_equal(got[:3][:, flat].T, o["points"]) and np.array_equal(got[3:][:, flat].T, o["colors"])
        assert not got[:, ~flat].any()
    else:
        got = sdk.data[:, :n].cpu().numpy()
        assert np.array_equal(got[:3].T, o["points"]) and np.array_equal(got[3:].T, o["colors"])
        assert np.array_equal(sdk.src_index[:n].cpu().numpy(), o["src_index"])
        # against the reference form on the pixels both keep: at most one unit in the last place
        a, b = sdk.src_index[:n].cpu().numpy(), ref.src_index[:int(ref.counts_host()[0])].cpu().numpy()
        both = np.intersect1d(a, b)
        pa = got[:3].T[np.isin(a, both)]
        pb = ref.data[:3, :len(b)].cpu().numpy().T[np.isin(b, both)]
        assert len(both) > 0.99 * n
        assert (np.abs(pa.astype(np.float64) - pb) <= np.spacing(np.abs(pb)).astype(np.float64)).all()
    with pytest.raises(RuntimeError):
        rv.deproject_batch(d, c, cam, geometry="sdk_f32", kernel="tma")
