"""PLY output compatible with the reference's viewers (o3d.io.read_point_cloud) and a reader for round trips.

* write_point_cloud ... o3d.io.write_point_cloud call sites femto_bolt_code/scripts/create_masked_ply.py:177,
  april_tag_bg_removal_pl.py:535-537 (write_ascii=False, compressed=False); layout per SURVEY Appendix B.1:
  binary little-endian, `property double x/y/z` (Open3D's default) then `uchar red/green/blue` =
  round(clamp(c,0,1)*255).  coord="float" gives the 15-byte records the SDK writers emit
  (better_three_capture.py:242, capture_aligned_all.py:262).
* read_point_cloud .... the 31 o3d.io.read_point_cloud call sites (e.g. view_point_cloud.py:104): float or double
  coordinates, optional uchar colours (-> /255.0), ascii or binary little-endian; other vertex properties are skipped.

The vertex records are packed on the GPU (rv_pack_ply_records) so a cloud crosses PCIe once, already in file layout;
reading a binary file is the mirror image (rv_unpack_ply_records).
"""
from __future__ import annotations

import os

import numpy as np

from . import _ops
from .cloud import PointCloud

_NP_T = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
         "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
         "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def ply_header(n: int, has_color: bool, coord: str = "double", ascii_: bool = False, has_normals: bool = False) -> bytes:
    lines = ["ply", "format ascii 1.0" if ascii_ else "format binary_little_endian 1.0", "comment Created by Open3D",
             f"element vertex {n}", f"property {coord} x", f"property {coord} y", f"property {coord} z"]
    if has_normals:  # Open3D's order: x y z nx ny nz red green blue
        lines += [f"property {coord} nx", f"property {coord} ny", f"property {coord} nz"]
    if has_color:
        lines += ["property uchar red", "property uchar green", "property uchar blue"]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def cloud_to_ply_records(pcd: PointCloud, coord: str = "double") -> np.ndarray:
    """uint8 [N, record_bytes] vertex records in file layout, packed on the device."""
    cdt = "f64" if coord == "double" else "f32"
    rec = 3 * (8 if cdt == "f64" else 4) + 3
    if len(pcd) == 0:
        return np.zeros((0, rec if pcd._has_color else rec - 3), np.uint8)
    raw = _ops.pack_ply_records(pcd._data, len(pcd), pcd._has_color, "unit", cdt).cpu().numpy().reshape(len(pcd), rec)
    return raw if pcd._has_color else np.ascontiguousarray(raw[:, :rec - 3])


def write_point_cloud(filename, pointcloud: PointCloud, write_ascii: bool = False, compressed: bool = False,
                      print_progress: bool = False, coord: str = "double") -> bool:
    """Returns True on success like Open3D; `compressed` has no effect on PLY (as in Open3D)."""
    if coord not in ("double", "float"):
        raise ValueError("coord must be 'double' or 'float'")
    filename = os.fspath(filename)
    n = len(pointcloud)
    rec = cloud_to_ply_records(pointcloud, coord)
    normals = pointcloud.has_normals()
    if normals and n:  # splice the normals between the coordinates and the colours of every record
        cw = 8 if coord == "double" else 4
        nb = np.ascontiguousarray(pointcloud.normals.astype("<f8" if cw == 8 else "<f4")).view(np.uint8).reshape(n, 3 * cw)
        rec = np.ascontiguousarray(np.concatenate([rec[:, :3 * cw], nb, rec[:, 3 * cw:]], axis=1))
    with open(filename, "wb") as f:
        f.write(ply_header(n, pointcloud._has_color, coord, write_ascii, normals and n > 0))
        if not write_ascii:
            f.write(rec.tobytes())
        else:
            cw = 8 if coord == "double" else 4
            nv = 6 if (normals and n) else 3
            xyz = np.ascontiguousarray(rec[:, :nv * cw]).view("<f8" if cw == 8 else "<f4").reshape(n, nv)
            rgb = rec[:, nv * cw:nv * cw + 3] if pointcloud._has_color else None
            fmt = "%.10f" if cw == 8 else "%.9g"
            for i in range(n):
                row = " ".join(fmt % v for v in xyz[i])
                if rgb is not None:
                    row += " %d %d %d" % tuple(int(v) for v in rgb[i])
                f.write((row + "\n").encode("ascii"))
    return True


def read_ply_vertices(filename):
    """(header lines, structured numpy array of the vertex element)."""
    with open(os.fspath(filename), "rb") as f:
        data = f.read()
    marker = b"end_header\n"
    if not data.startswith(b"ply") or marker not in data:
        raise RuntimeError(f"not a PLY file: {filename}")
    end = data.index(marker) + len(marker)
    header = data[:end].decode("ascii", "replace").splitlines()
    fmt, n, props, in_vertex, before_vertex = None, 0, [], False, 0
    for line in header[1:]:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            if tok[1] == "vertex":
                in_vertex, n = True, int(tok[2])
            else:
                if not props:
                    before_vertex += int(tok[2])
                in_vertex = False
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise RuntimeError("list properties on the vertex element are not supported")
            props.append((tok[2], _NP_T[tok[1]]))
    if before_vertex:
        raise RuntimeError("elements before `vertex` are not supported")
    dt = np.dtype(props)
    if fmt == "binary_little_endian":
        arr = np.frombuffer(data, dtype=dt, count=n, offset=end)
    elif fmt == "ascii":
        rows = np.loadtxt(data[end:].decode("ascii").splitlines()[:n], ndmin=2) if n else np.zeros((0, len(props)))
        arr = np.zeros(n, dtype=dt)
        for i, (name, _) in enumerate(props):
            arr[name] = rows[:, i]
    else:
        raise RuntimeError(f"unsupported PLY format: {fmt}")
    return header, arr


def _binary_vertex_layout(filename):
    """For a binary little-endian PLY whose first element is `vertex` with float/double x, y, z (same type) and, if
    present, uchar red/green/blue: (data offset, n, record bytes, xyz offsets, 'f32'|'f64', rgb offsets or None,
    nx/ny/nz offsets or None).
    Anything else returns None and takes the host reader."""
    with open(filename, "rb") as f:
        head = f.read(1 << 16)
    marker = b"end_header\n"
    if not head.startswith(b"ply") or marker not in head:
        return None
    end = head.index(marker) + len(marker)
    fmt, n, props, in_vertex, seen_vertex = None, 0, [], False, False
    for line in head[:end].decode("ascii", "replace").splitlines()[1:]:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            if tok[1] == "vertex" and not seen_vertex and not props:
                in_vertex, seen_vertex, n = True, True, int(tok[2])
            elif not seen_vertex:
                return None  # an element before the vertices
            else:
                in_vertex = False
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list" or tok[1] not in _NP_T:
                return None
            props.append((tok[2], tok[1]))
    if fmt != "binary_little_endian" or not seen_vertex:
        return None
    off, offsets, types = 0, {}, {}
    for name, t in props:
        offsets[name], types[name] = off, np.dtype(_NP_T[t])
        off += types[name].itemsize
    if not all(k in offsets for k in "xyz"):
        return None
    ct = {types[k].str for k in "xyz"}
    if ct not in ({"<f4"}, {"<f8"}):
        return None
    rgb = None
    if all(k in offsets for k in ("red", "green", "blue")):
        if any(types[k].str != "|u1" for k in ("red", "green", "blue")):
            return None
        rgb = [offsets["red"], offsets["green"], offsets["blue"]]
    nrm = None
    if all(k in offsets for k in ("nx", "ny", "nz")) and {types[k].str for k in ("nx", "ny", "nz")} == ct:
        nrm = [offsets["nx"], offsets["ny"], offsets["nz"]]
    return end, n, off, [offsets[k] for k in "xyz"], "f32" if ct == {"<f4"} else "f64", rgb, nrm


def read_point_cloud(filename, device=None, dtype: str = "f64") -> PointCloud:
    """PLY -> PointCloud on the GPU.  A missing file gives an empty cloud with a warning, as Open3D does.  Binary files
    are uploaded as they lie on disk and their vertex records are unpacked on the device (rv_unpack_ply_records); ascii
    files and unusual layouts are parsed on the host."""
    filename = os.fspath(filename)
    if not os.path.exists(filename):
        print(f"[Open3D-compatible WARNING] Read PLY failed: unable to open file: {filename}")
        return PointCloud(None, 0, False, device=device)
    layout = _binary_vertex_layout(filename)
    if layout is not None and layout[1] > 0:
        import torch
        start, n, rec, xyz_off, cdt, rgb_off, nrm_off = layout
        raw = np.fromfile(filename, dtype=np.uint8, count=n * rec, offset=start)
        if raw.size != n * rec:
            raise RuntimeError(f"PLY file is truncated: {filename}")
        dev = _ops.require_cuda(device)
        records = torch.from_numpy(raw).to(dev)
        planes = _ops.unpack_ply_records(records, n, rec, xyz_off, cdt, rgb_off, dtype)
        pc = PointCloud(planes, n, rgb_off is not None)
        if nrm_off is not None:  # the same unpacker, pointed at nx ny nz
            pc._normals = _ops.unpack_ply_records(records, n, rec, nrm_off, cdt, None, "f64")
        return pc
    _, arr = read_ply_vertices(filename)
    names = arr.dtype.names or ()
    if not all(k in names for k in ("x", "y", "z")):
        raise RuntimeError("PLY vertex element has no x/y/z")
    pts = np.stack([arr["x"], arr["y"], arr["z"]], axis=1).astype(np.float64)
    cols = None
    if all(k in names for k in ("red", "green", "blue")):
        cols = np.stack([arr["red"], arr["green"], arr["blue"]], axis=1).astype(np.float64) / 255.0
    return PointCloud.from_arrays(pts, cols, device=device, dtype=dtype)
