"""Multi-GPU host logic: the batch shards contiguously across ranks with NO data-path collective; the
only exchange is one all-gather of per-instance results at the end.  Backend-agnostic: NCCL (cuda tensors) on the B200
box, gloo (cpu tensors) in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_bounds(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: B_g = ceil(B / G); trailing ranks may get a short or empty shard."""
    per = (batch + world - 1) // world
    return min(rank * per, batch), min((rank + 1) * per, batch)


def shard_arrays(arrays: dict, rank: int, world: int) -> dict:
    """Slices every [B, ...] array of `arrays` (None entries pass through) to this rank's shard."""
    B = next(v for v in arrays.values() if v is not None).shape[0]
    lo, hi = shard_bounds(B, rank, world)
    return {k: (None if v is None else v[lo:hi]) for k, v in arrays.items()}


def gather_results(cost, iterations, status, batch: int, device=None):
    """The path's single collective: all-gather of {final cost f64, iterations, status} (24 B / instance).
    Every rank passes its shard's arrays and receives the whole batch's, in instance order.  Shards are padded to
    B_g = ceil(B/G) so the collective is a plain fixed-size all_gather."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(batch, rank, world)
    assert len(cost) == hi - lo, f"rank {rank}: shard has {len(cost)} instances, expected {hi - lo}"
    per = (batch + world - 1) // world
    mine = np.full((per, 3), np.nan)
    mine[: hi - lo, 0] = cost
    mine[: hi - lo, 1] = iterations
    mine[: hi - lo, 2] = status
    if world == 1:
        allr = mine[None]
    else:
        t = torch.from_numpy(mine)
        if device is not None:
            t = t.to(device)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        allr = torch.stack(out).cpu().numpy()
    rows = []
    for r in range(world):
        l, h = shard_bounds(batch, r, world)
        rows.append(allr[r, : h - l])
    full = np.concatenate(rows, axis=0)
    return full[:, 0].copy(), full[:, 1].astype(np.int32), full[:, 2].astype(np.int32)
