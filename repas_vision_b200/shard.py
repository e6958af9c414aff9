"""Multi-GPU sharding of independent frames: one process per GPU, no collective on the hot path.

Frames (and whole fusion problems) share nothing but a few hundred bytes of calibration, so rank g of G
takes the contiguous block [g*B/G, (g+1)*B/G) of the batch.  torch.distributed is used only
AFTER the per-frame kernels: an all-gather of per-rank point counts and, for the fusion configuration, a
variable-length gather of merged voxel clouds to rank 0 (counts first, then grouped send / recv).  Works on NCCL
(GPU; tests/test_gpu_multi.py under torchrun) and gloo (CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPU cores of the NUMA node the GPU hangs off, BEFORE it allocates pinned host buffers, so
    that the staging memory of the host-buffer pipeline is local to the GPU's PCIe root (first-touch placement).  One
    process per GPU makes this a per-rank call.  Best effort: returns {} and changes nothing when /sys does not tell."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id  # "0000:1B:00.0" (torch >= 2.5)
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bus = bus[-12:]
        except Exception:
            return {}
    try:
        node = int(open(f"/sys/bus/pci/devices/{str(bus).lower()}/numa_node").read().strip())
        if node < 0:
            return {}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        return {}


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition; the first `total % world` ranks get one extra unit."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world(group=None) -> tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def gather_counts(local_counts: torch.Tensor, total_frames: int | None = None, group=None) -> torch.Tensor:
    """Per-frame kept-point counts of every rank, concatenated in frame order on every rank.

    local_counts: int64 [frames of this rank] (block partition of `total_frames`).  One all-gather of int64."""
    rank, world = _world(group)
    local_counts = local_counts.to(torch.int64).reshape(-1)
    if world == 1:
        return local_counts.clone()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_counts.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_counts.numel()], dtype=torch.int64, device=local_counts.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    if total_frames is not None and sum(sizes) != total_frames:
        raise RuntimeError(f"ranks hold {sum(sizes)} frames, expected {total_frames}")
    width = max(sizes) if sizes else 0
    padded = torch.zeros(max(width, 1), dtype=torch.int64, device=local_counts.device)
    padded[:local_counts.numel()] = local_counts
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


def gather_clouds(data: torch.Tensor, n: int, dst: int = 0, group=None):
    """Variable-length gather of SoA clouds ([planes, >=n] each) to rank `dst`: one all-gather of the per-rank
    point counts, then ONE group of point-to-point transfers -- every other rank sends exactly its n points per plane,
    `dst` receives each straight into that rank's columns of the merged cloud (ncclGroupStart / ncclSend / ncclRecv /
    ncclGroupEnd on NCCL; nothing is padded and nothing reaches ranks that do not want it).
    Returns (merged [planes, sum n], per-rank sizes) on `dst`, (None, sizes) elsewhere."""
    rank, world = _world(group)
    planes = data.shape[0]
    if world == 1:
        return data[:, :n].contiguous(), [n]
    dev = data.device
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    sizes = [int(s.item()) for s in sizes]
    if rank == dst:
        # plane-major per source rank, so that every receive buffer is one contiguous block
        parts = [torch.empty((planes, m), dtype=data.dtype, device=dev) for m in sizes]
        parts[dst].copy_(data[:, :n])
        ops = [dist.P2POp(dist.irecv, parts[r], _global_rank(r, group), group) for r in range(world) if r != dst and sizes[r] > 0]
    else:
        mine = data[:, :n].contiguous()
        ops = [dist.P2POp(dist.isend, mine, _global_rank(dst, group), group)] if n > 0 else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != dst:
        return None, sizes
    return torch.cat(parts, dim=1), sizes


def _global_rank(group_rank: int, group=None) -> int:
    return group_rank if group is None else dist.get_global_rank(group, group_rank)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing reduction for the benchmark: the slowest rank defines the step."""
    rank, world = _world(group)
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device=None, group=None) -> float:
    rank, world = _world(group)
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
