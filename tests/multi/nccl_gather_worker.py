"""torchrun worker of tests/test_gpu_multi.py: the fusion configuration's gather on real GPUs over NCCL.

Every rank fuses its own four-pose problem (K1 -> K3 -> K4 on its GPU), the per-rank voxel counts are all-gathered and the
fused clouds travel to rank 0 (counts first, then grouped send / recv).  Rank 0 recomputes every rank's problem on its own GPU
and checks that what arrived is the same cloud as a key set (Open3D's voxel order is unspecified); it prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def problem(rv, seed, dev):
    from synth import synth_batch
    H, W = 360, 640
    depth, bgr = synth_batch(4, H, W, seed0=seed)
    cam = rv.Camera(456.0, 455.5, 319.5, 179.5, W, H)
    batch = rv.deproject_batch(torch.from_numpy(depth).to(dev), torch.from_numpy(bgr).to(dev), cam, max_distance=2.0)
    poses = []
    for i in range(4):
        an = np.deg2rad(90.0 * i)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(an), 0, np.sin(an)], [0, 1, 0], [-np.sin(an), 0, np.cos(an)]]
        T[:3, 3] = [0.02 * i, -0.01, 0.8]
        poses.append(T)
    return rv.fuse_views([batch.frame(i) for i in range(4)], poses, 0.005)


def canon(xyzrgb: np.ndarray) -> np.ndarray:
    """Rows of a [6, n] cloud sorted lexicographically: equal clouds up to voxel order compare equal."""
    a = np.ascontiguousarray(xyzrgb.T)
    return a[np.lexsort(a.T[::-1])]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import repas_vision_b200 as rv
    from repas_vision_b200 import shard
    mine = problem(rv, 100 + 7 * rank, dev)
    counts = shard.gather_counts(torch.tensor([len(mine)], dtype=torch.int64, device=dev), world)
    merged, sizes = shard.gather_clouds(mine._data, len(mine), dst=0)
    ok, detail = True, ""
    if rank == 0:
        ok = merged is not None and merged.shape == (6, sum(sizes)) and counts.cpu().tolist() == sizes
        off = 0
        for r in range(world):
            ref = problem(rv, 100 + 7 * r, dev)
            got = merged[:, off:off + sizes[r]].cpu().numpy()
            off += sizes[r]
            same = sizes[r] == len(ref) and np.array_equal(canon(got), canon(ref._data[:, :len(ref)].cpu().numpy()))
            ok = ok and same
            detail += f"rank{r}:{sizes[r]}:{'same' if same else 'DIFFERENT'} "
    else:
        ok = merged is None
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"ok": bool(flag.item()), "world": world, "sizes": sizes, "detail": detail.strip(),
                          "backend": dist.get_backend(), "so": rv.library_path()}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
