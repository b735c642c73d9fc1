"""Multi-GPU sharding of viewpoint batches.

Frames are independent units (the reference rebuilds Renderer and Pixels per frame, src/game.rs:505-519), so a batch
shards by viewpoint, one process / context per GPU, assets replicated, NO collective on the draw path.  The only exchange
is a host-side gather of the per-frame checksums (8 bytes per frame), off the timed path.

Assignment is STRIDED (`shard_indices`: rank r owns viewpoints r, r + G, r + 2G, ...): consecutive viewpoints of a walk
path see the same rooms and cost the same, so every GPU gets the same mix.  Contiguous ranges (`shard_range`, round 1) gave
each GPU a different part of the map, and the slowest range set the step time (8 GPUs: 0.907 of linear).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of viewpoints owned by `rank`: [g*N/G, (g+1)*N/G)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def shard_indices(n_total: int, rank: int, world: int) -> np.ndarray:
    """Viewpoint indices owned by `rank`: rank, rank + world, rank + 2*world, ... (< n_total)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return np.arange(rank, n_total, world, dtype=np.int64)


def unshard(parts: list[np.ndarray], n_total: int) -> np.ndarray:
    """Inverse of shard_indices: parts[r] holds the values of rank r's viewpoints in its own order."""
    world = len(parts)
    out = np.zeros(n_total, np.asarray(parts[0]).dtype)
    for r, p in enumerate(parts):
        idx = shard_indices(n_total, r, world)
        if len(p) != len(idx):
            raise ValueError("rank %d: %d values for %d viewpoints" % (r, len(p), len(idx)))
        out[idx] = p
    return out


def gather_checksums(local: np.ndarray, device=None) -> np.ndarray:
    """All-gather the per-frame u64 checksums of every rank, in rank order (ranks may own different counts).
    Works with any torch.distributed backend (gloo on CPU, nccl on GPUs); returns `local` when not initialised."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.copy()
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([local.size], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    buf = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
    if local.size:
        buf[:local.size] = torch.from_numpy(local.view(np.int64).copy()).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = [p[:c].cpu().numpy().view(np.uint64) for p, c in zip(parts, counts)]
    return np.concatenate(out) if out else np.zeros(0, np.uint64)


def checksum_of_checksums(sums: np.ndarray) -> int:
    with np.errstate(over="ignore"):
        return int(np.asarray(sums, np.uint64).sum(dtype=np.uint64))
