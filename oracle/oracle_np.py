"""CPU oracle (numpy) for the RGB-D -> point-cloud hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``repas_vision_b200/`` may import this
module; it is used by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the timed CPU baseline.  It restates, statement by statement but in its own
words, what the reference scripts compute (paths relative to the reference
checkout):

* depth units ............ femto_bolt_code/scripts/better_three_capture.py:118-125,
                           femto_bolt_code/scripts/custom_reader.py:34-40,
                           realsense_d415i/canopy_detection/canopy_return.py:279-317
* masked deprojection .... femto_bolt_code/scripts/create_masked_ply.py:56-107
* single pixel ........... realsense_d415i/canopy_detection/canopy_return.py:183-206
* cloud predicates ....... realsense_d415i/capture_scripts/distance_masking_on_ply.py:9-23,
                           femto_bolt_code/scripts/view_point_cloud.py:109-116,
                           femto_bolt_code/scripts/april_tag_bg_removal_pl.py:450-455
* 4x4 transform .......... Open3D 0.19 PointCloud::Transform,
                           call sites final_view_with_cad.py:333,
                           vis_tool_april_tag_pose_validaiton.py:239-245
* voxel_down_sample ...... Open3D 0.19 PointCloud::VoxelDownSample,
                           call sites mpa_icp_export.py:44,174, create_masked_ply.py:164
* create_from_rgbd_image . Open3D 0.19
* outliers, normals, ICP . Open3D 0.19 RemoveStatisticalOutliers / EstimateNormals / RegistrationICP,
                           call sites create_masked_ply.py:168-174, mpa_icp_export.py:166-208
* registration ........... librealsense 2.56 align z16->other;
                           loops live in oracle/oracle.c, a slow independent
                           numpy formulation lives here for cross-checking.

Pinning status (see DESIGN.md "Oracle"):
* deprojection + depth units: PINNED by the reference's four canopy_y goldens and
  by golden vectors generated from the reference's own functions
  (tests/golden/make_golden.py);
* voxel_down_sample, transform, outlier removal, normals, ICP, registration, Open3D/SDK clouds: PARITY UNPINNED --
  the arithmetic lives in wheels that are not in the reference checkout
  (open3d==0.19.0, pyrealsense2==2.56.5.9235, pyorbbecsdk) and the reference
  stores no expected outputs for them.  Their published algorithms are restated
  here and cross-checked by independent formulations only.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------- units
def depth_to_meters(raw: np.ndarray, rule: str = "mul_f32", scale: float | None = None):
    """uint16 raw depth -> metres.

    mul_f32: f32(u16) * f32(0.001)   (better_three_capture.py:123-124)
    div_f32: f32(u16) / f32(1000.0)  (custom_reader.py:39-40; Open3D RGBDImage)
    div_f64: f64(u16) / 1000.0       (canopy_return.py:315)
    """
    raw = np.asarray(raw)
    if rule == "mul_f32":
        s = F32(0.001 if scale is None else scale)
        return raw.astype(F32) * s
    if rule == "div_f32":
        s = F32(1000.0 if scale is None else scale)
        return raw.astype(F32) / s
    if rule == "div_f64":
        s = F64(1000.0 if scale is None else scale)
        return raw.astype(F64) / s
    raise ValueError(rule)


def median_depth_window(depth_raw: np.ndarray, x: int, y: int, window: int = 5):
    """Median of the non-zero raw depths in a clipped window (raw units, float64).

    canopy_return.py:293-312: the pixel is clamped into the image, the window is
    [y-h, y+h] x [x-h, x+h] clipped to the image, zeros are dropped, np.median.
    Returns None when the window holds no valid depth.
    """
    h_img, w_img = depth_raw.shape
    x = min(max(x, 0), w_img - 1)
    y = min(max(y, 0), h_img - 1)
    half = window // 2
    win = depth_raw[max(0, y - half):min(h_img, y + half + 1), max(0, x - half):min(w_img, x + half + 1)]
    vals = win[win > 0]
    if vals.size == 0:
        return None
    return float(np.median(vals))


def deproject_pixel_to_point(fx, fy, ppx, ppy, pixel, depth_m):
    """canopy_return.py:199-204: subtract, multiply, divide, all float64."""
    u, v = pixel
    return ((u - ppx) * depth_m / fx, (v - ppy) * depth_m / fy, depth_m)


# ------------------------------------------------------------------- deprojection
def create_masked_pointcloud(bgr, depth_m, mask, fx, fy, cx, cy, invert_mask=False):
    """float64 points / colours of create_masked_ply.py:74-100, row-major order."""
    sel = (mask == 0) if invert_mask else (mask > 0)
    keep = sel & np.isfinite(depth_m) & (depth_m > 0)
    rows, cols = np.nonzero(keep)  # row-major, identical order to a boolean gather
    z = depth_m[rows, cols].astype(F64)
    x = (cols.astype(F64) - cx) * z / fx
    y = (rows.astype(F64) - cy) * z / fy
    pts = np.stack([x, y, z], axis=1)
    rgb = bgr[rows, cols][:, ::-1].astype(F64) / 255.0
    return pts, rgb


def brown_conrady_undistort(x, y, k, iters=10):
    """librealsense rs2_deproject_pixel_to_point, BROWN_CONRADY branch, float64."""
    k1, k2, p1, p2, k3 = [F64(c) for c in k]
    xo, yo = x.copy(), y.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
        dy = 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y)
        x = (xo - dx) * icdist
        y = (yo - dy) * icdist
    return x, y


def ray_table(fx, fy, cx, cy, dist, model, width, height):
    """[H,W,2] float64 normalised rays, rs2_deproject_pixel_to_point order (Appendix B.3)."""
    u = np.arange(width, dtype=F64)[None, :].repeat(height, 0)
    v = np.arange(height, dtype=F64)[:, None].repeat(width, 1)
    x = (u - cx) / fx
    y = (v - cy) / fy
    if model == "inverse_brown_conrady":
        k1, k2, p1, p2, k3 = [F64(c) for c in dist]
        r2 = x * x + y * y
        f = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
        ux = x * f + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
        uy = y * f + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y)
        x, y = ux, uy
    elif model == "brown_conrady":
        x, y = brown_conrady_undistort(x, y, dist)
    elif model not in ("none", "modified_brown_conrady"):
        raise ValueError(model)
    return np.stack([x, y], axis=-1)


def deproject_mask(depth, bgr, mask=None, *, fx, fy, cx, cy, depth_kind="u16", unit_rule="mul_f32",
                   unit_scale=None, invert_mask=False, depth_trunc=None, z_clip=None, r_max=None, aabb=None,
                   out_dtype="f32", color_scale="unit", rays=None, geometry="reference"):
    """One frame of the fused kernel's contract (DESIGN.md "K1").

    geometry="sdk_f32": the arithmetic of the SDK clouds instead of the reference's numpy statement -- librealsense
    rs2_deproject_pixel_to_point as rs.pointcloud / Orbbec's PointCloudFilter evaluate it, all in float32:
    x = z * ((u - ppx) / fx), y = z * ((v - ppy) / fy) (pinhole cameras only).

    Returns dict(valid [H,W] bool, points [N,3], colors [N,3], src_index [N]) in
    row-major order of kept pixels.  The cloud predicates are evaluated in
    float64 on the STORED coordinates (exact up-casts when out_dtype == 'f32'),
    so fusing a predicate never changes what filtering the produced cloud
    afterwards would give.
    """
    depth = np.asarray(depth)
    H, W = depth.shape
    if depth_kind == "u16":
        z_native = depth_to_meters(depth, unit_rule, unit_scale)
        ok = depth != 0
    else:
        z_native = depth.astype(F32)
        ok = np.isfinite(z_native) & (z_native > 0)
    z32 = z_native.astype(F32)
    z64 = z_native.astype(F64)
    if mask is not None:
        ok &= (mask == 0) if invert_mask else (mask > 0)
    if depth_trunc is not None:
        ok &= ~(z32 >= F32(depth_trunc))
    u = np.arange(W, dtype=F64)[None, :]
    v = np.arange(H, dtype=F64)[:, None]
    with np.errstate(all="ignore"):
        if geometry == "sdk_f32":
            if rays is not None:
                raise ValueError("sdk_f32 geometry is restated for pinhole cameras only")
            uf, vf = np.arange(W, dtype=F32)[None, :], np.arange(H, dtype=F32)[:, None]
            x64 = (z32 * ((uf - F32(cx)) / F32(fx))).astype(F64)
            y64 = (z32 * ((vf - F32(cy)) / F32(fy))).astype(F64)
        elif rays is None:
            x64 = (u - cx) * z64 / fx
            y64 = (v - cy) * z64 / fy
        else:
            x64 = z64 * rays[..., 0]
            y64 = z64 * rays[..., 1]
        if out_dtype == "f32":
            xs, ys, zs = x64.astype(F32), y64.astype(F32), z64.astype(F32)
        else:
            xs, ys, zs = x64, y64, z64
        X, Y, Z = xs.astype(F64), ys.astype(F64), zs.astype(F64)
        if z_clip is not None:
            ok &= (Z >= z_clip[0]) & (Z <= z_clip[1])
        if r_max is not None:
            ok &= np.sqrt((X * X + Y * Y) + Z * Z) < r_max
        if aabb is not None:
            lo, hi = aabb
            ok &= (X >= lo[0]) & (X <= hi[0]) & (Y >= lo[1]) & (Y <= hi[1]) & (Z >= lo[2]) & (Z <= hi[2])
    rows, cols = np.nonzero(ok)
    pts = np.stack([xs[rows, cols], ys[rows, cols], zs[rows, cols]], axis=1)
    if bgr is not None:
        c = bgr[rows, cols][:, ::-1].astype(F64)
        if color_scale == "unit":
            c = c / 255.0
        c = c.astype(F32 if out_dtype == "f32" else F64)
    else:
        c = None
    return dict(valid=ok, points=pts, colors=c, src_index=(rows * W + cols).astype(np.int32))


def create_from_rgbd_image(bgr, depth_u16, fx, fy, cx, cy, depth_scale=1000.0, depth_trunc=3.0, extrinsic=None):
    """Open3D create_from_color_and_depth + create_from_rgbd_image (Appendix B.2), float64 out."""
    d = depth_u16.astype(F32) / F32(depth_scale)
    d = np.where(d >= F32(depth_trunc), F32(0), d)
    rows, cols = np.nonzero(d > 0)
    z = d[rows, cols].astype(F64)
    x = (cols.astype(F64) - cx) * z / fx
    y = (rows.astype(F64) - cy) * z / fy
    pts = np.stack([x, y, z], axis=1)
    if extrinsic is not None:
        pts = transform(pts, np.linalg.inv(np.asarray(extrinsic, dtype=F64)))
    rgb = bgr[rows, cols][:, ::-1].astype(F64) / 255.0
    return pts, rgb


# -------------------------------------------------------------- cloud predicates
def distance_mask(points, max_distance=1.0):
    """distance_masking_on_ply.py:12-16 -- np.linalg.norm(points, axis=1) < max_distance."""
    p = np.asarray(points).astype(F64)
    s = (p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) + p[:, 2] * p[:, 2]
    return np.sqrt(s) < max_distance


def z_clip_mask(points, z_min=None, z_max=None):
    """view_point_cloud.py:109-116, inclusive on both ends."""
    z = np.asarray(points)[:, 2].astype(F64)
    keep = np.ones(z.shape[0], dtype=bool)
    if z_min is not None:
        keep &= z >= float(z_min)
    if z_max is not None:
        keep &= z <= float(z_max)
    return keep


def aabb_mask(points, min_b, max_b):
    """april_tag_bg_removal_pl.py:450-455, inclusive."""
    p = np.asarray(points).astype(F64)
    return ((p[:, 0] >= min_b[0]) & (p[:, 0] <= max_b[0]) & (p[:, 1] >= min_b[1]) & (p[:, 1] <= max_b[1])
            & (p[:, 2] >= min_b[2]) & (p[:, 2] <= max_b[2]))


# --------------------------------------------------------------------- transform
def transform(points, T):
    """Open3D PointCloud::Transform: q = T [p,1]; p' = q[:3] / q[3], float64.

    Accumulation order fixed as ((T0*x + T1*y) + T2*z) + T3, no fused multiply-add.
    """
    p = np.asarray(points).astype(F64)
    T = np.asarray(T, dtype=F64)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    q = [((T[i, 0] * x + T[i, 1] * y) + T[i, 2] * z) + T[i, 3] for i in range(4)]
    return np.stack([q[0] / q[3], q[1] / q[3], q[2] / q[3]], axis=1)


def invert_rigid(T):
    """[R|t] -> [R^T | -R^T t], float64 on host."""
    T = np.asarray(T, dtype=F64)
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4)
    out[:3, :3] = R.T
    out[:3, 3] = -(R.T @ t)
    return out


# ------------------------------------------------------------------------- voxel
def voxel_keys(points, voxel_size, min_bound=None):
    p = np.asarray(points).astype(F64)
    if min_bound is None:
        min_bound = p.min(axis=0)
    origin = np.asarray(min_bound, dtype=F64) - voxel_size * 0.5
    return np.floor((p - origin) / voxel_size).astype(np.int32)


def voxel_down_sample(points, colors, voxel_size):
    """Open3D VoxelDownSample (Appendix B.1).  Returns (keys[M,3] int32 sorted
    lexicographically, centroids[M,3], mean colours[M,3] or None, counts[M]).
    Sums are accumulated in point-index order in float64, as the C++ loop does."""
    if voxel_size <= 0:
        raise ValueError("voxel_size <= 0")
    p = np.asarray(points).astype(F64)
    if p.shape[0] == 0:
        e = np.zeros((0, 3))
        return np.zeros((0, 3), np.int32), e, (e if colors is not None else None), np.zeros(0, np.int64)
    mn, mx = p.min(axis=0), p.max(axis=0)
    if voxel_size * np.iinfo(np.int32).max < (mx - mn).max():
        raise ValueError("voxel_size is too small")
    keys = voxel_keys(p, voxel_size, mn)
    uniq, inv = np.unique(keys, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    m = uniq.shape[0]
    cnt = np.bincount(inv, minlength=m)
    sums = np.zeros((m, 3))
    np.add.at(sums, inv, p)  # unbuffered, index order
    cent = sums / cnt[:, None].astype(F64)
    col = None
    if colors is not None:
        cs = np.zeros((m, 3))
        np.add.at(cs, inv, np.asarray(colors).astype(F64))
        col = cs / cnt[:, None].astype(F64)
    return uniq.astype(np.int32), cent, col, cnt.astype(np.int64)


# ------------------------------------------------------------------ registration
def _f32(x):
    return np.asarray(x, dtype=F32)


def _rs_deproject_f32(px, py, depth, cam):
    """rs2_deproject_pixel_to_point in float32 (Appendix B.3)."""
    fx, fy, ppx, ppy = F32(cam["fx"]), F32(cam["fy"]), F32(cam["cx"]), F32(cam["cy"])
    x = (px - ppx) / fx
    y = (py - ppy) / fy
    model = cam.get("model", "none")
    if model == "inverse_brown_conrady":
        k1, k2, p1, p2, k3 = [F32(c) for c in cam["dist"]]
        r2 = x * x + y * y
        f = F32(1) + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
        ux = x * f + F32(2) * p1 * x * y + p2 * (r2 + F32(2) * x * x)
        uy = y * f + F32(2) * p2 * x * y + p1 * (r2 + F32(2) * y * y)
        x, y = ux, uy
    elif model == "brown_conrady":
        k1, k2, p1, p2, k3 = [F32(c) for c in cam["dist"]]
        xo, yo = x.copy(), y.copy()
        for _ in range(10):
            r2 = x * x + y * y
            icdist = F32(1) / (F32(1) + ((k3 * r2 + k2) * r2 + k1) * r2)
            dx = F32(2) * p1 * x * y + p2 * (r2 + F32(2) * x * x)
            dy = F32(2) * p2 * x * y + p1 * (r2 + F32(2) * y * y)
            x = (xo - dx) * icdist
            y = (yo - dy) * icdist
    return depth * x, depth * y, depth


def _rs_project_f32(X, Y, Z, cam):
    """rs2_project_point_to_pixel in float32 (Appendix B.3)."""
    fx, fy, ppx, ppy = F32(cam["fx"]), F32(cam["fy"]), F32(cam["cx"]), F32(cam["cy"])
    with np.errstate(all="ignore"):
        x = X / Z
        y = Y / Z
        model = cam.get("model", "none")
        if model in ("modified_brown_conrady", "inverse_brown_conrady"):
            k1, k2, p1, p2, k3 = [F32(c) for c in cam["dist"]]
            r2 = x * x + y * y
            f = F32(1) + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
            # both models: x *= f; y *= f; the tangential terms then use the SCALED x, y and the old r2
            x = x * f
            y = y * f
            dx = x + F32(2) * p1 * x * y + p2 * (r2 + F32(2) * x * x)
            dy = y + F32(2) * p2 * x * y + p1 * (r2 + F32(2) * y * y)
            x, y = dx, dy
        elif model == "brown_conrady":
            k1, k2, p1, p2, k3 = [F32(c) for c in cam["dist"]]
            r2 = x * x + y * y
            f = F32(1) + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
            xf = x * f
            yf = y * f
            dx = xf + F32(2) * p1 * x * y + p2 * (r2 + F32(2) * x * x)
            dy = yf + F32(2) * p2 * x * y + p1 * (r2 + F32(2) * y * y)
            x, y = dx, dy
        return x * fx + ppx, y * fy + ppy


PIX_LIMIT = F32(1 << 30)


def register_corners(depth_u16, dcam, ccam, R_colmajor, t, depth_units=0.001):
    """Per depth pixel: (x0,y0,x1,y1,ok) of the colour rectangle it covers.

    Vectorised float32 restatement of the two-corner mapping of librealsense's
    align_images (Appendix B.3).  A corner whose projected pixel is NaN or beyond
    +-2^30 rejects the pixel (the C `(int)` cast would be undefined there).
    """
    Hd, Wd = depth_u16.shape
    R = _f32(R_colmajor)
    tt = _f32(t)
    z = depth_u16.astype(F32) * F32(depth_units)
    xs = np.arange(Wd, dtype=F32)[None, :].repeat(Hd, 0)
    ys = np.arange(Hd, dtype=F32)[:, None].repeat(Wd, 1)
    out = []
    ok = depth_u16 != 0
    for off in (F32(-0.5), F32(0.5)):
        X, Y, Z = _rs_deproject_f32(xs + off, ys + off, z, dcam)
        # rs2_transform_point_to_point, column-major rotation
        Xc = R[0] * X + R[3] * Y + R[6] * Z + tt[0]
        Yc = R[1] * X + R[4] * Y + R[7] * Z + tt[1]
        Zc = R[2] * X + R[5] * Y + R[8] * Z + tt[2]
        px, py = _rs_project_f32(Xc, Yc, Zc, ccam)
        with np.errstate(all="ignore"):
            px = px + F32(0.5)
            py = py + F32(0.5)
            fin = np.isfinite(px) & np.isfinite(py) & (np.abs(px) < PIX_LIMIT) & (np.abs(py) < PIX_LIMIT)
        ok &= fin
        ix = np.where(fin, px, 0).astype(np.int64)  # C cast: truncation toward zero
        iy = np.where(fin, py, 0).astype(np.int64)
        out += [ix, iy]
    x0, y0, x1, y1 = out
    ok &= ~((x0 < 0) | (y0 < 0) | (x1 >= ccam["width"]) | (y1 >= ccam["height"]))
    return x0, y0, x1, y1, ok


def register_depth_to_color_gather(depth_u16, dcam, ccam, R_colmajor, t, depth_units=0.001):
    """Independent (slow) formulation: for every colour pixel take the minimum over
    the depth pixels whose rectangle covers it.  Used to cross-check oracle.c on
    small images.  Returns (aligned u16 [Hc,Wc], winner int32 [Hc,Wc])."""
    Hc, Wc = ccam["height"], ccam["width"]
    x0, y0, x1, y1, ok = register_corners(depth_u16, dcam, ccam, R_colmajor, t, depth_units)
    best = np.full((Hc, Wc), np.iinfo(np.int64).max, dtype=np.int64)
    Wd = depth_u16.shape[1]
    ys, xs = np.nonzero(ok)
    for yy, xx in zip(ys, xs):
        key = (int(depth_u16[yy, xx]) << 32) | (yy * Wd + xx)
        a, b, c, d = x0[yy, xx], y0[yy, xx], x1[yy, xx], y1[yy, xx]
        if c >= a and d >= b:
            sub = best[b:d + 1, a:c + 1]
            np.minimum(sub, key, out=sub)
    empty = best == np.iinfo(np.int64).max
    aligned = np.where(empty, 0, best >> 32).astype(np.uint16)
    winner = np.where(empty, -1, best & 0xFFFFFFFF).astype(np.int32)
    return aligned, winner


# ------------------------------------------------------- statistical outlier removal
def knn_mean_distance(points, k, chunk=512):
    """avg[i] of Open3D RemoveStatisticalOutliers: mean of sqrt(d2) over the k nearest neighbours of point i, itself
    included (KDTreeFlann::SearchKNN returns them ascending); brute force in float64, d2 = (dx*dx + dy*dy) + dz*dz."""
    P = np.asarray(points, dtype=np.float64)
    n = P.shape[0]
    kk = min(int(k), n)
    out = np.empty(n, np.float64)
    for i0 in range(0, n, chunk):
        q = P[i0:i0 + chunk]
        dx = q[:, None, 0] - P[None, :, 0]
        dy = q[:, None, 1] - P[None, :, 1]
        dz = q[:, None, 2] - P[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        near = np.sort(np.partition(d2, kk - 1, axis=1)[:, :kk], axis=1)
        r = np.sqrt(near)
        acc = np.zeros(q.shape[0])
        for j in range(kk):  # std::accumulate: left to right
            acc = acc + r[:, j]
        out[i0:i0 + chunk] = acc / kk
    return out


def nearest_neighbor_distance(points, chunk=512):
    """Open3D 0.19 PointCloud::ComputeNearestNeighborDistance (ply_to_stl.py:45,56): sqrt of the second squared distance of
    SearchKNN(point, 2) -- the nearest OTHER point (0 for a duplicate); zeros for fewer than two points."""
    P = np.asarray(points, dtype=np.float64)
    n = P.shape[0]
    if n < 2:
        return np.zeros(n)
    out = np.empty(n)
    for i0 in range(0, n, chunk):
        q = P[i0:i0 + chunk]
        dx = q[:, None, 0] - P[None, :, 0]
        dy = q[:, None, 1] - P[None, :, 1]
        dz = q[:, None, 2] - P[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        out[i0:i0 + chunk] = np.sqrt(np.partition(d2, 1, axis=1)[:, 1])
    return out


def statistical_outlier_indices(avg, std_ratio):
    """Index-ordered sums of PointCloud::RemoveStatisticalOutliers (Open3D 0.19): returns (indices, mean, std, threshold)."""
    avg = np.asarray(avg, np.float64)
    valid = int((avg >= 0).sum())
    cloud_mean = 0.0
    for v in avg.tolist():
        if v > 0:
            cloud_mean = cloud_mean + v
    cloud_mean /= valid
    sq = 0.0
    for v in avg.tolist():
        if v > 0:
            sq = sq + (v - cloud_mean) * (v - cloud_mean)
    with np.errstate(all="ignore"):
        std = float(np.sqrt(np.float64(sq) / np.float64(valid - 1)))
    thr = cloud_mean + std_ratio * std
    ind = np.nonzero((avg > 0) & (avg < thr))[0]
    return ind, cloud_mean, std, thr


def remove_statistical_outlier(points, nb_neighbors=20, std_ratio=2.0):
    """create_masked_ply.py:169 -> (kept indices, avg distances)."""
    avg = knn_mean_distance(points, nb_neighbors)
    ind, _, _, _ = statistical_outlier_indices(avg, std_ratio)
    return ind, avg


def estimate_normals(points, radius, max_nn, camera_location=None, chunk=512):
    """Open3D 0.19 PointCloud::EstimateNormals with KDTreeSearchParamHybrid(radius, max_nn) (create_masked_ply.py:173):
    neighbours = the max_nn nearest (the point itself included) with squared distance < radius^2; fewer than three -> (0,0,1);
    covariance from the cumulants (utility::ComputeCovariance); normal = unit eigenvector of the smallest eigenvalue.  Open3D
    gets it from a closed-form 3x3 solver, this restatement from numpy.linalg.eigh: same line, sign arbitrary until
    orient_normals_towards_camera_location (:174) turns it towards camera_location."""
    P = np.asarray(points, dtype=np.float64)
    n = P.shape[0]
    kk = min(int(max_nn), n)
    N = np.zeros((n, 3))
    r2 = float(radius) * float(radius)
    for i0 in range(0, n, chunk):
        q = P[i0:i0 + chunk]
        dx = q[:, None, 0] - P[None, :, 0]
        dy = q[:, None, 1] - P[None, :, 1]
        dz = q[:, None, 2] - P[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        idx = np.argsort(d2, axis=1, kind="stable")[:, :kk]
        dd = np.take_along_axis(d2, idx, axis=1)
        for r in range(q.shape[0]):
            sel = idx[r][dd[r] < r2]
            if sel.size < 3:
                N[i0 + r] = (0.0, 0.0, 1.0)
                continue
            nb = P[sel]
            mu = nb.mean(axis=0)
            C = (nb[:, :, None] * nb[:, None, :]).mean(axis=0) - mu[:, None] * mu[None, :]
            w, v = np.linalg.eigh(C)
            N[i0 + r] = v[:, 0]
    if camera_location is not None:
        ref = np.asarray(camera_location, dtype=np.float64)[None, :] - P
        flip = (N * ref).sum(axis=1) < 0
        N[flip] *= -1.0
    return N


# ----------------------------------------------------------------------- ICP
def nearest_correspondences(source, target, max_distance, chunk=256):
    """Open3D 0.19 GetRegistrationResultAndCorrespondences (pipelines/registration/Registration.cpp), as
    registration_icp uses it (mpa_icp_export.py:187): per source point KDTreeFlann::SearchHybrid(point, max_distance, 1) =
    the nearest target point if its squared distance is < max_distance^2.  Returns (nearest int32 [n], -1 = unmatched;
    fitness; inlier_rmse).  Equal distances go to the lower target index (the KD-tree's choice is unspecified)."""
    S = np.asarray(source, dtype=np.float64)
    T = np.asarray(target, dtype=np.float64)
    n = S.shape[0]
    near = np.full(n, -1, np.int32)
    if n == 0 or T.shape[0] == 0 or not (max_distance > 0.0):
        return near, 0.0, 0.0
    r2 = float(max_distance) * float(max_distance)
    err2 = 0.0
    for i0 in range(0, n, chunk):
        q = S[i0:i0 + chunk]
        dx = T[None, :, 0] - q[:, None, 0]
        dy = T[None, :, 1] - q[:, None, 1]
        dz = T[None, :, 2] - q[:, None, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        j = np.argmin(d2, axis=1)  # first minimum = lowest index
        best = d2[np.arange(q.shape[0]), j]
        ok = best < r2
        near[i0:i0 + chunk][ok] = j[ok]
        err2 += float(best[ok].sum())
    m = int((near >= 0).sum())
    if m == 0:
        return near, 0.0, 0.0
    return near, m / float(n), float(np.sqrt(err2 / m))


def vector6d_to_matrix4d(x):
    """utility::TransformVector6dToMatrix4d: AngleAxis(x2, Z) * AngleAxis(x1, Y) * AngleAxis(x0, X), translation x[3:6]."""
    a, b, c = x[0], x[1], x[2]
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = x[3:6]
    return T


def point_to_plane_update(source, target, target_normals, near):
    """TransformationEstimationPointToPlane::ComputeTransformation: r = (s - t) . n_t, J = [s x n_t, n_t],
    x = solve(J^T J, -J^T r) (utility::SolveJacobianSystemAndObtainExtrinsicMatrix), identity without correspondences."""
    sel = np.nonzero(near >= 0)[0]
    if sel.size == 0:
        return np.eye(4)
    s = np.asarray(source, dtype=np.float64)[sel]
    t = np.asarray(target, dtype=np.float64)[near[sel]]
    nt = np.asarray(target_normals, dtype=np.float64)[near[sel]]
    r = ((s - t) * nt).sum(axis=1)
    J = np.concatenate([np.cross(s, nt), nt], axis=1)
    JTJ = J.T @ J
    JTr = J.T @ r
    x = np.linalg.solve(JTJ, -JTr)
    return vector6d_to_matrix4d(x)


def point_to_point_update(source, target, near, with_scaling=False):
    """TransformationEstimationPointToPoint::ComputeTransformation = Eigen::umeyama(source, target, with_scaling)."""
    sel = np.nonzero(near >= 0)[0]
    if sel.size == 0:
        return np.eye(4)
    s = np.asarray(source, dtype=np.float64)[sel]
    t = np.asarray(target, dtype=np.float64)[near[sel]]
    ms, mt = s.mean(axis=0), t.mean(axis=0)
    sd, td = s - ms, t - mt
    sigma = td.T @ sd / sel.size
    U, d, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    T = np.eye(4)
    if with_scaling:
        c = float(d @ S) / float((sd * sd).sum() / sel.size)
        T[:3, :3] = c * R
        T[:3, 3] = mt - c * (R @ ms)
    else:
        T[:3, :3] = R
        T[:3, 3] = mt - R @ ms
    return T


def registration_icp(source, target, max_distance, init=None, target_normals=None, point_to_plane=True, with_scaling=False,
                     relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
    """Open3D 0.19 RegistrationICP: transform the source by init, match, then repeat {estimate the update from the matches,
    compose it on the left, transform the working copy by the update, match again} until both fitness and inlier_rmse
    change by less than the criteria or max_iteration is reached.  Returns (transformation, fitness, inlier_rmse, nearest,
    iterations run)."""
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    pcd = np.asarray(source, dtype=np.float64)
    if not np.array_equal(T, np.eye(4)):
        pcd = transform(pcd, T)
    near, fit, rmse = nearest_correspondences(pcd, target, max_distance)
    it = 0
    for it in range(1, max_iteration + 1):
        if point_to_plane:
            upd = point_to_plane_update(pcd, target, target_normals, near)
        else:
            upd = point_to_point_update(pcd, target, near, with_scaling)
        T = upd @ T
        pcd = transform(pcd, upd)
        prev = (fit, rmse)
        near, fit, rmse = nearest_correspondences(pcd, target, max_distance)
        if abs(prev[0] - fit) < relative_fitness and abs(prev[1] - rmse) < relative_rmse:
            break
    return T, fit, rmse, near, it


# ---- the same three neighbourhood operations on a KD-tree (scipy's cKDTree in the role of Open3D's KDTreeFlann): an
# independent formulation used to cross-check the brute-force restatements above and, being O(n log n) like the
# reference, as the timed CPU side of bench.py's rows for them
def knn_mean_distance_kdtree(points, k):
    from scipy.spatial import cKDTree
    P = np.asarray(points, dtype=np.float64)
    d, _ = cKDTree(P).query(P, k=min(int(k), len(P)))
    return d.reshape(len(P), -1).mean(axis=1)


def estimate_normals_kdtree(points, radius, max_nn, camera_location=None):
    from scipy.spatial import cKDTree
    P = np.asarray(points, dtype=np.float64)
    n = len(P)
    kk = min(int(max_nn), n)
    d, idx = cKDTree(P).query(P, k=kk)
    d, idx = d.reshape(n, kk), idx.reshape(n, kk)
    ok = d * d < float(radius) * float(radius)
    cnt = ok.sum(axis=1)
    nb = P[idx] * ok[:, :, None]
    c1 = nb.sum(axis=1) / np.maximum(cnt, 1)[:, None]
    c2 = np.einsum("nki,nkj->nij", nb, nb) / np.maximum(cnt, 1)[:, None, None]
    C = c2 - c1[:, :, None] * c1[:, None, :]
    _, v = np.linalg.eigh(C)
    N = v[:, :, 0].copy()
    N[cnt < 3] = (0.0, 0.0, 1.0)
    if camera_location is not None:
        ref = np.asarray(camera_location, dtype=np.float64)[None, :] - P
        N[(N * ref).sum(axis=1) < 0] *= -1.0
    return N


def nearest_correspondences_kdtree(source, target, max_distance):
    from scipy.spatial import cKDTree
    S = np.asarray(source, dtype=np.float64)
    d, j = cKDTree(np.asarray(target, dtype=np.float64)).query(S, k=1)
    ok = d * d < float(max_distance) * float(max_distance)
    near = np.where(ok, j, -1).astype(np.int32)
    m = int(ok.sum())
    return near, (m / float(len(S)) if len(S) else 0.0), (float(np.sqrt((d[ok] ** 2).sum() / m)) if m else 0.0)


# ----------------------------------------------------------------------- PLY read
def read_ply_minimal(path):
    """Independent minimal PLY vertex reader (binary LE / ascii) used to check the
    product's writer against the format o3d.io.read_point_cloud accepts."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode("ascii").splitlines()
    assert header[0] == "ply"
    fmt = None
    n = 0
    props = []
    in_vertex = False
    for line in header[1:]:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                n = int(tok[2])
        elif tok[0] == "property" and in_vertex:
            props.append((tok[2], tok[1]))
    np_t = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1"}
    dt = np.dtype([(name, np_t[t]) for name, t in props])
    if fmt == "binary_little_endian":
        arr = np.frombuffer(data, dtype=dt, count=n, offset=end)
    else:
        rows = np.loadtxt(data[end:].decode("ascii").splitlines()[:n], ndmin=2)
        arr = np.zeros(n, dtype=dt)
        for i, (name, _) in enumerate(props):
            arr[name] = rows[:, i]
    return header, arr
