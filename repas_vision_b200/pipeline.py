"""Host-buffer front of the hot path: numpy (ideally pinned) frames in, host clouds out, copies overlapped with compute.

This is the call a script makes when its frames live in host memory, the batched form of

    color = frame_to_bgr_image(color_frame)                  better_three_capture.py:101-106 (NV12 from the camera)
    raw, depth_m, _ = depth_to_meters(frame)                 better_three_capture.py:118-125
    pcd = create_masked_pointcloud(bgr, depth_m, mask, ...)  create_masked_ply.py:56-107
    keep = np.linalg.norm(points, axis=1) < 1.0              distance_masking_on_ply.py:12-19

Three CUDA streams walk the batch in chunks: H2D of chunk i+1, the fused kernel on chunk i (COMPACT_PACKED: the
frames of a chunk land back to back, so each plane is ONE contiguous device->host copy), D2H of chunk i-1.
The only host synchronisation is reading a chunk's B+1 offsets to size its D2H, done one chunk behind the copy
engine so PCIe stays busy.  PCIe bounds this path, so the bytes per frame are what counts:

* colour frames may arrive as the camera's NV12 ([B, H*3/2, W] uint8: 1.5 instead of 3 bytes per pixel); the kernel reads
  them directly (no BGR image is ever written), a 720p frame then crosses PCIe as 3.5 bytes per pixel;
* colours come back as the bytes they are (colors="u8": one r,g,b,0 word per point, 16 bytes per point with its float32
  xyz) and are expanded to k / 255.0 on the host only when somebody asks (`HostBatchResult.colors`); colors="float"
  returns three float32 colour planes (24 bytes per point).

The device-resident API (`deproject_batch` on CUDA tensors) is the one the HBM roofline applies to.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _ops
from .calibration import Camera


class _PinnedPool:
    """Pinned host blocks for the results.  Page-locking memory costs milliseconds per hundred megabytes, far more than the
    copy it serves, so blocks are rounded up to powers of two and reused -- but only blocks their owner handed back with
    `HostBatchResult.release()`: arrays given out by a result alias its block, so nothing is recycled behind the caller."""

    def __init__(self, max_bytes: int = 8 << 30):
        self.free: dict[int, list[torch.Tensor]] = {}
        self.max_bytes, self.held = int(max_bytes), 0

    def take(self, elements: int) -> torch.Tensor:
        cap = 1 << max(16, int(elements - 1).bit_length())
        blocks = self.free.get(cap)
        if blocks:
            self.held -= cap * 4
            return blocks.pop()
        return torch.empty(cap, dtype=torch.float32, pin_memory=True)

    def give(self, block: torch.Tensor) -> None:
        cap = block.numel()
        if self.held + cap * 4 <= self.max_bytes:
            self.free.setdefault(cap, []).append(block)
            self.held += cap * 4


class HostCloudChunk:
    """Clouds of `frames` consecutive frames on the host: float32 planes [3, 4 or 6, total] (pinned), offsets [frames+1].
    Four planes = x, y, z and the packed colour words (bytes r,g,b,0)."""

    def __init__(self, first_frame: int, frames: int, planes: torch.Tensor, offsets: np.ndarray, block=None, pool=None,
                 packed_color: bool = False):
        self.first_frame, self.frames = first_frame, frames
        self._planes, self.offsets = planes, offsets
        self._block, self._pool = block, pool
        self.packed_color = packed_color

    @property
    def total(self) -> int:
        return int(self.offsets[-1])

    def release(self) -> None:
        """Hand the pinned block back for reuse.  Every array obtained from this chunk is invalid afterwards."""
        if self._pool is not None and self._block is not None:
            self._pool.give(self._block)
        self._block = self._planes = None

    @property
    def planes(self) -> np.ndarray:
        if self._planes is None:
            raise RuntimeError("this result was released")
        return self._planes.numpy()

    def frame(self, i: int):
        """Views of frame `first_frame + i`: (xyz [3,n] float32, colours).  colours = [n,3] uint8 (r,g,b) for a packed-colour
        chunk, [3,n] float32 in [0,1] for colors="float", None without colour."""
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        pl = self.planes
        if pl.shape[0] < 4:
            return pl[:3, a:b], None
        if self.packed_color:
            return pl[:3, a:b], pl[3, a:b].view(np.uint8).reshape(b - a, 4)[:, :3]
        return pl[:3, a:b], pl[3:6, a:b]


class HostBatchResult:
    def __init__(self, chunks, n_frames, d2h_bytes, h2d_bytes):
        self.chunks, self.n_frames, self.d2h_bytes, self.h2d_bytes = chunks, n_frames, d2h_bytes, h2d_bytes

    def release(self) -> None:
        """Return the pinned result blocks to the pipeline's pool (a loop that runs batch after batch calls this when it is
        done with a result; otherwise the blocks are simply freed with the result).  Arrays obtained from `frame`, `planes`
        alias those blocks and must not be used afterwards; `points` / `colors` return copies and stay valid."""
        for c in self.chunks:
            c.release()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()
        return False

    @property
    def counts(self) -> np.ndarray:
        return np.concatenate([np.diff(c.offsets) for c in self.chunks]) if self.chunks else np.zeros(0, np.int64)

    def frame(self, b: int):
        for c in self.chunks:
            if c.first_frame <= b < c.first_frame + c.frames:
                return c.frame(b - c.first_frame)
        raise IndexError(b)

    def points(self, b: int) -> np.ndarray:
        """(n,3) float64 like np.asarray(pcd.points)."""
        return np.ascontiguousarray(self.frame(b)[0].T, dtype=np.float64)

    def write_ply(self, b: int, filename) -> bool:
        """Frame b as the binary PLY the SDK writers of the capture scripts emit (save_point_cloud_to_ply
        better_three_capture.py:242, points.export_to_ply capture_aligned_all.py:262): `float x y z` + `uchar red green blue`,
        15 bytes per vertex, readable by o3d.io.read_point_cloud / rv.read_point_cloud.  With byte colours (colors="u8") the
        record is exactly what came back from the GPU: nothing is rescaled on the way to the file."""
        import os
        from .ply import ply_header
        xyz, rgb = self.frame(b)
        n = xyz.shape[1]
        if rgb is None:
            rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
        else:
            rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
            if rgb.dtype == np.uint8:
                rec["r"], rec["g"], rec["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
            else:  # float planes in [0,1]: Open3D's round(clamp(c) * 255)
                q = np.rint(np.clip(rgb.astype(np.float64), 0.0, 1.0) * 255.0).astype(np.uint8)
                rec["r"], rec["g"], rec["b"] = q[0], q[1], q[2]
        rec["x"], rec["y"], rec["z"] = xyz[0], xyz[1], xyz[2]
        with open(os.fspath(filename), "wb") as f:
            f.write(ply_header(n, rgb is not None, "float"))
            f.write(rec.tobytes())
        return True

    def colors(self, b: int) -> np.ndarray:
        """(n,3) float64 in [0,1] like np.asarray(pcd.colors); packed bytes are expanded here as k / 255.0, which is the
        reference's own float64 statement (create_masked_ply.py:100)."""
        rgb = self.frame(b)[1]
        if rgb is None:
            return np.zeros((0, 3))
        if rgb.dtype == np.uint8:
            return rgb.astype(np.float64) / 255.0
        return np.ascontiguousarray(rgb.T, dtype=np.float64)


class HostPipeline:
    def __init__(self, camera: Camera, height: int, width: int, *, max_distance=None, z_clip=None, aabb=None,
                 unit_rule: str = "mul_f32", depth_scale=None, use_mask: bool = False, invert_mask: bool = False,
                 with_color: bool = True, color_format: str = "bgr", colors: str = "u8", chunk_frames: int = 32,
                 slots: int = 3, device=None):
        """color_format "nv12": the colour frames arrive as the camera delivers them ([B, H*3/2, W] uint8, the capture
        script's preferred format, better_three_capture.py:101-106,159) and are read by the deprojection kernel itself.
        colors "u8" (default): colours come back as bytes, 16 bytes per point; "float": three float32 planes in [0,1]."""
        if color_format not in ("bgr", "nv12"):
            raise ValueError("color_format must be 'bgr' or 'nv12'")
        if colors not in ("u8", "float"):
            raise ValueError("colors must be 'u8' or 'float'")
        if color_format == "nv12" and (int(height) % 2 or int(width) % 2):
            raise ValueError("NV12 needs even image sides")
        self.color_format = color_format
        self.dev = _ops.require_cuda(device)
        self.cam, self.H, self.W = camera, int(height), int(width)
        self.P = self.H * self.W
        self.C, self.slots = int(chunk_frames), max(2, int(slots))
        self.use_mask, self.with_color = use_mask, with_color
        self.packed_color = with_color and colors == "u8"
        self.planes = 3 if not with_color else (4 if self.packed_color else 6)
        self.kw = dict(depth_kind="u16", unit_rule=unit_rule, unit_scale=depth_scale, invert_mask=invert_mask,
                       r_max=max_distance, z_clip=z_clip, aabb=aabb, mode="compact_packed", out_dtype="f32",
                       color_format=color_format, color_scale="packed8" if self.packed_color else "unit")
        d = self.dev
        C, H, W = self.C, self.H, self.W
        cshape = (C, H, W, 3) if color_format == "bgr" else (C, H * 3 // 2, W)
        self.color_bytes = 3 * self.P if color_format == "bgr" else self.P * 3 // 2
        self.d_depth = [torch.empty((C, H, W), dtype=torch.uint16, device=d) for _ in range(self.slots)]
        self.d_color = [torch.empty(cshape, dtype=torch.uint8, device=d) for _ in range(self.slots)] if with_color else None
        self.d_mask = [torch.empty((C, H, W), dtype=torch.uint8, device=d) for _ in range(self.slots)] if use_mask else None
        self.d_out = [torch.empty((self.planes, C * self.P), dtype=torch.float32, device=d) for _ in range(self.slots)]
        self.h_off = [torch.empty(C + 1, dtype=torch.int64, pin_memory=True) for _ in range(self.slots)]
        self.pool = _PinnedPool()
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(d) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(self.slots)]
        self.ev_k = [torch.cuda.Event() for _ in range(self.slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(self.slots)]

    @staticmethod
    def _host_tensor(a, dtype) -> torch.Tensor:
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if t.dtype != dtype:
            raise RuntimeError(f"expected {dtype}, got {t.dtype}")
        return t.contiguous()

    def _check_inputs(self, depth_u16, color, mask):
        hd = self._host_tensor(depth_u16, torch.uint16)
        B = hd.shape[0]
        if tuple(hd.shape[1:]) != (self.H, self.W):
            raise RuntimeError(f"depth must be [B,{self.H},{self.W}], got {tuple(hd.shape)}")
        hc = None
        if self.with_color:
            if color is None:
                raise RuntimeError("this pipeline was built with_color=True: colour frames are required")
            hc = self._host_tensor(color, torch.uint8)
            want = (B, self.H, self.W, 3) if self.color_format == "bgr" else (B, self.H * 3 // 2, self.W)
            if tuple(hc.shape) != want:
                raise RuntimeError(f"Color/depth size mismatch: color {tuple(hc.shape)}, depth {tuple(hd.shape)}")
        hm = None
        if self.use_mask:
            if mask is None:
                raise RuntimeError("this pipeline was built use_mask=True: mask is required")
            hm = self._host_tensor(mask, torch.uint8)
            if tuple(hm.shape) != (B, self.H, self.W):
                raise RuntimeError(f"Mask/depth size mismatch: mask {tuple(hm.shape)}, depth {tuple(hd.shape)}")
        return hd, hc, hm, B

    def run(self, depth_u16, bgr=None, mask=None, *, _copy_only_like: HostBatchResult | None = None) -> HostBatchResult:
        """depth_u16 [B,H,W] uint16, colour frames ([B,H,W,3] BGR or [B,H*3/2,W] NV12) uint8, mask [B,H,W] uint8 host arrays
        (pinned memory copies asynchronously; pageable memory still works, the driver then stages it)."""
        hd, hc, hm, B = self._check_inputs(depth_u16, bgr, mask)
        probe = _copy_only_like
        n_chunks = (B + self.C - 1) // self.C
        chunks, d2h, h2d = [None] * n_chunks, 0, 0
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(cur)

        def issue(i):
            nonlocal h2d
            s = i % self.slots
            f0, f1 = i * self.C, min(B, (i + 1) * self.C)
            n = f1 - f0
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_k[s])  # the kernel that last read this input slot
                self.d_depth[s][:n].copy_(hd[f0:f1], non_blocking=True)
                h2d += n * self.P * 2
                if hc is not None:
                    self.d_color[s][:n].copy_(hc[f0:f1], non_blocking=True)
                    h2d += n * self.color_bytes
                if hm is not None:
                    self.d_mask[s][:n].copy_(hm[f0:f1], non_blocking=True)
                    h2d += n * self.P
                self.ev_in[s].record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(self.ev_in[s])
                self.s_k.wait_event(self.ev_out[s])  # the D2H that last read this output slot
                if probe is None:
                    r = _ops.deproject(self.d_depth[s][:n], None if hc is None else self.d_color[s][:n],
                                       None if hm is None else self.d_mask[s][:n], self.cam, out=self.d_out[s], **self.kw)
                    self.h_off[s][:n + 1].copy_(r["counts"], non_blocking=True)
                self.ev_k[s].record(self.s_k)

        def drain(i):
            nonlocal d2h
            s = i % self.slots
            f0, f1 = i * self.C, min(B, (i + 1) * self.C)
            n = f1 - f0
            self.ev_k[s].synchronize()
            offsets = self.h_off[s][:n + 1].numpy().copy() if probe is None else probe.chunks[i].offsets
            total = int(offsets[-1])
            block = self.pool.take(self.planes * max(total, 1))
            host = block[:self.planes * max(total, 1)].view(self.planes, max(total, 1))
            with torch.cuda.stream(self.s_out):
                if total:
                    for p in range(self.planes):
                        host[p, :total].copy_(self.d_out[s][p, :total], non_blocking=True)
                self.ev_out[s].record(self.s_out)
            d2h += total * 4 * self.planes + (n + 1) * 8
            chunks[i] = HostCloudChunk(f0, n, host[:, :total], offsets, block, self.pool, self.packed_color)

        for i in range(n_chunks + 1):
            if i < n_chunks:
                issue(i)
            if i >= 1:
                drain(i - 1)
        self.s_out.synchronize()
        cur.wait_stream(self.s_k)
        return HostBatchResult(chunks, B, d2h, h2d)

    def copy_probe(self, depth_u16, bgr=None, mask=None, *, like: HostBatchResult) -> HostBatchResult:
        """The copy schedule of `run` without the kernel: the same host->device copies into the same slots and device->host
        copies of exactly the bytes `like` (a finished run on the same inputs) brought back, on the same three streams.
        What this takes is what the host side and PCIe cost by themselves (bench.py reports it next to the end-to-end rate)."""
        return self.run(depth_u16, bgr, mask, _copy_only_like=like)
