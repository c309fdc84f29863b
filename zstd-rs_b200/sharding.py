"""Host-side frame sharding for multi-GPU decode.

Frames are independent by format (ruzstd resets tables, offset history and window per frame, decoding/scratch.rs:52-68), so a
frame list splits into contiguous shards, one per rank, with no data-path collective.  Shards are balanced by compressed bytes
(the work proxy available before decoding); a frame is never split.
"""
import numpy as np


def shard_frames(src_sizes, world_size, rank):
    """Contiguous [lo, hi) frame range of `rank`, balanced by cumulative compressed bytes."""
    sizes = np.asarray(src_sizes, dtype=np.uint64)
    n = len(sizes)
    if world_size <= 1 or n == 0:
        return 0, n
    csum = np.concatenate([[0], np.cumsum(sizes)]).astype(np.float64)
    total = csum[-1]
    bounds = [int(np.searchsorted(csum, total * r / world_size, side="left")) for r in range(world_size + 1)]
    bounds[0], bounds[-1] = 0, n
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds[rank], bounds[rank + 1]


def gather_counters(local, world_size, dist=None, device="cpu"):
    """Sum-reduce a small dict of integer counters and max-reduce 'max_*' keys over ranks (the only collective used)."""
    import torch
    keys = sorted(local)
    sums = torch.tensor([float(local[k]) for k in keys if not k.startswith("max_")], dtype=torch.float64, device=device)
    maxs = torch.tensor([float(local[k]) for k in keys if k.startswith("max_")], dtype=torch.float64, device=device)
    if world_size > 1 and dist is not None:
        if len(sums):
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        if len(maxs):
            dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.startswith("max_"):
            out[k] = float(maxs[mi]); mi += 1
        else:
            out[k] = float(sums[si]); si += 1
    return out
