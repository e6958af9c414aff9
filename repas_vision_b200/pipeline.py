"""Host-buffer front of the hot path: numpy (ideally pinned) frames in, host clouds out, copies overlapped with compute.

This is the call a script makes when its frames live in host memory, the batched form of

    raw, depth_m, _ = depth_to_meters(frame)                 better_three_capture.py:118-125
    pcd = create_masked_pointcloud(bgr, depth_m, mask, ...)  create_masked_ply.py:56-107
    keep = np.linalg.norm(points, axis=1) < 1.0              distance_masking_on_ply.py:12-19

Three CUDA streams walk the batch in chunks: H2D of chunk i+1, the fused kernel on chunk i (COMPACT_PACKED: the
frames of a chunk land back to back, so each plane is ONE contiguous device->host copy), D2H of chunk i-1.
The only host synchronisation is reading a chunk's B+1 offsets to size its D2H, done one chunk behind the copy
engine so PCIe stays busy.  PCIe bounds this path (4.6 MB in per 720p frame); the device-resident API
(`deproject_batch` on CUDA tensors) is the one the HBM roofline applies to.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _ops
from .calibration import Camera


class _PinnedPool:
    """Pinned host blocks for the results, recycled when a result is dropped.  Page-locking memory costs milliseconds per
    hundred megabytes, far more than the copy it serves, so blocks are rounded up to powers of two and reused."""

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
    """Clouds of `frames` consecutive frames on the host: planes [6 or 3, total] float32 (pinned), offsets [frames+1]."""

    def __init__(self, first_frame: int, frames: int, planes: torch.Tensor, offsets: np.ndarray, block=None, pool=None):
        self.first_frame, self.frames = first_frame, frames
        self._planes, self.offsets = planes, offsets
        self._block, self._pool = block, pool

    def __del__(self):
        if self._pool is not None and self._block is not None:
            self._pool.give(self._block)
            self._block = None

    @property
    def planes(self) -> np.ndarray:
        return self._planes.numpy()

    def frame(self, i: int):
        """(xyz [3,n], rgb [3,n] or None) views of frame `first_frame + i`."""
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        pl = self.planes
        return pl[:3, a:b], (pl[3:6, a:b] if pl.shape[0] >= 6 else None)


class HostBatchResult:
    def __init__(self, chunks, n_frames, d2h_bytes, h2d_bytes):
        self.chunks, self.n_frames, self.d2h_bytes, self.h2d_bytes = chunks, n_frames, d2h_bytes, h2d_bytes

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

    def colors(self, b: int) -> np.ndarray:
        rgb = self.frame(b)[1]
        return np.zeros((0, 3)) if rgb is None else np.ascontiguousarray(rgb.T, dtype=np.float64)


class HostPipeline:
    def __init__(self, camera: Camera, height: int, width: int, *, max_distance=None, z_clip=None, aabb=None,
                 unit_rule: str = "mul_f32", depth_scale=None, use_mask: bool = False, invert_mask: bool = False,
                 with_color: bool = True, color_format: str = "bgr", chunk_frames: int = 32, slots: int = 3, device=None):
        """color_format "nv12": the colour frames arrive as the camera delivers them ([B, H*3/2, W] uint8, the capture
        script's preferred format, better_three_capture.py:101-106,159) and are decoded on the GPU; the frame then crosses
        PCIe at 3.5 instead of 5 bytes per pixel."""
        if color_format not in ("bgr", "nv12"):
            raise ValueError("color_format must be 'bgr' or 'nv12'")
        if color_format == "nv12" and (int(height) % 2 or int(width) % 2):
            raise ValueError("NV12 needs even image sides")
        self.color_format = color_format
        self.dev = _ops.require_cuda(device)
        self.cam, self.H, self.W = camera, int(height), int(width)
        self.P = self.H * self.W
        self.C, self.slots = int(chunk_frames), max(2, int(slots))
        self.kw = dict(depth_kind="u16", unit_rule=unit_rule, unit_scale=depth_scale, invert_mask=invert_mask,
                       r_max=max_distance, z_clip=z_clip, aabb=aabb, mode="compact_packed", out_dtype="f32")
        self.use_mask, self.with_color = use_mask, with_color
        self.planes = 6 if with_color else 3
        d = self.dev
        C, H, W = self.C, self.H, self.W
        self.d_depth = [torch.empty((C, H, W), dtype=torch.uint16, device=d) for _ in range(self.slots)]
        self.d_bgr = [torch.empty((C, H, W, 3), dtype=torch.uint8, device=d) for _ in range(self.slots)] if with_color else None
        self.d_nv12 = ([torch.empty((C, H * 3 // 2, W), dtype=torch.uint8, device=d) for _ in range(self.slots)]
                       if with_color and color_format == "nv12" else None)
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

    def run(self, depth_u16, bgr=None, mask=None) -> HostBatchResult:
        """depth_u16 [B,H,W] uint16, bgr [B,H,W,3] uint8, mask [B,H,W] uint8 host arrays (pinned memory copies
        asynchronously; pageable memory still works, the driver then stages it)."""
        hd = self._host_tensor(depth_u16, torch.uint16)
        B = hd.shape[0]
        if tuple(hd.shape[1:]) != (self.H, self.W):
            raise RuntimeError(f"depth must be [B,{self.H},{self.W}], got {tuple(hd.shape)}")
        hc = None
        if self.with_color:
            if bgr is None:
                raise RuntimeError("this pipeline was built with_color=True: bgr is required")
            hc = self._host_tensor(bgr, torch.uint8)
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
                if hc is not None and self.d_nv12 is not None:
                    self.d_nv12[s][:n].copy_(hc[f0:f1], non_blocking=True)
                    h2d += n * self.P * 3 // 2
                elif hc is not None:
                    self.d_bgr[s][:n].copy_(hc[f0:f1], non_blocking=True)
                    h2d += n * self.P * 3
                if hm is not None:
                    self.d_mask[s][:n].copy_(hm[f0:f1], non_blocking=True)
                    h2d += n * self.P
                self.ev_in[s].record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(self.ev_in[s])
                self.s_k.wait_event(self.ev_out[s])  # the D2H that last read this output slot
                if self.d_nv12 is not None:
                    _ops.nv12_to_bgr(self.d_nv12[s][:n], self.H, self.W, out=self.d_bgr[s][:n])
                r = _ops.deproject(self.d_depth[s][:n], None if hc is None else self.d_bgr[s][:n],
                                   None if hm is None else self.d_mask[s][:n], self.cam, out=self.d_out[s], **self.kw)
                self.h_off[s][:n + 1].copy_(r["counts"], non_blocking=True)
                self.ev_k[s].record(self.s_k)

        def drain(i):
            nonlocal d2h
            s = i % self.slots
            f0, f1 = i * self.C, min(B, (i + 1) * self.C)
            n = f1 - f0
            self.ev_k[s].synchronize()
            offsets = self.h_off[s][:n + 1].numpy().copy()
            total = int(offsets[-1])
            block = self.pool.take(self.planes * max(total, 1))
            host = block[:self.planes * max(total, 1)].view(self.planes, max(total, 1))
            with torch.cuda.stream(self.s_out):
                if total:
                    for p in range(self.planes):
                        host[p, :total].copy_(self.d_out[s][p, :total], non_blocking=True)
                self.ev_out[s].record(self.s_out)
            d2h += total * 4 * self.planes + (n + 1) * 8
            chunks[i] = HostCloudChunk(f0, n, host[:, :total], offsets, block, self.pool)

        for i in range(n_chunks + 1):
            if i < n_chunks:
                issue(i)
            if i >= 1:
                drain(i - 1)
        self.s_out.synchronize()
        cur.wait_stream(self.s_k)
        return HostBatchResult(chunks, B, d2h, h2d)
