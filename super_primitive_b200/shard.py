"""Multi-GPU sharding of independent alignment problems.

A unit is one (source keyframe, target set) problem; units are independent, so they are dealt
round-robin to the ranks (one process per GPU), optimised with no data-path communication, and the
KB-sized results (pose 4x4, log-depth seeds, final cost) are collected once at the end with a single
all-gather over NCCL (gloo in the CPU tests).  Variable segment counts are NaN-padded to the global
maximum.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_units, rank, world):
    """Round-robin assignment: unit i belongs to rank i % world."""
    return list(range(rank, n_units, world))


def owner_of(unit, world):
    return unit % world


def gather_results(poses, k_padded, costs, n_units, group=None, device=None):
    """All ranks contribute their local results; every rank receives the global, unit-ordered
    (poses (n,4,4), k (n,Nmax) NaN-padded, costs (n,)).  Local arrays are ordered like
    ``shard_indices``.  A rank whose shard is empty (n_units < world) passes ``None`` for the three arrays (and the
    ``device`` its collectives run on): it still takes part in every collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return poses, k_padded, costs
    world = dist.get_world_size(group)
    dev = poses.device if poses is not None else torch.device(device if device is not None else "cpu")
    n_local_max = (n_units + world - 1) // world
    n_local = 0 if poses is None else poses.shape[0]
    # agree on the padded segment count
    nmax = torch.tensor([k_padded.shape[1] if n_local else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
    nmax = int(nmax.item())
    width = 16 + nmax + 1
    payload = torch.full((n_local_max, width), float('nan'), dtype=torch.float32, device=dev)
    if n_local:
        payload[:n_local, :16] = poses.reshape(n_local, 16)
        payload[:n_local, 16:16 + k_padded.shape[1]] = k_padded
        payload[:n_local, 16 + nmax] = costs
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    out_pose = torch.empty((n_units, 4, 4), dtype=torch.float32, device=dev)
    out_k = torch.empty((n_units, nmax), dtype=torch.float32, device=dev)
    out_cost = torch.empty(n_units, dtype=torch.float32, device=dev)
    for r in range(world):
        idx = shard_indices(n_units, r, world)
        if not idx:
            continue
        blk = gathered[r][:len(idx)]
        ii = torch.tensor(idx, dtype=torch.int64, device=dev)
        out_pose[ii] = blk[:, :16].reshape(-1, 4, 4)
        out_k[ii] = blk[:, 16:16 + nmax]
        out_cost[ii] = blk[:, 16 + nmax]
    return out_pose, out_k, out_cost


def gather_ragged(vectors, n_units, group=None, device=None):
    """Variable-length results (one 1-D float32 tensor per LOCAL unit, ordered like ``shard_indices``) -> list of
    ``n_units`` tensors in unit order on every rank.  Used for mapping windows, whose result (frame poses, seeds of
    every keyframe, brightness terms, loss) has a window-dependent length: two all-reduces agree on the padded
    shape, one all-gather moves the NaN-padded rows with their lengths."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(vectors)
    world = dist.get_world_size(group)
    dev = vectors[0].device if vectors else torch.device(device if device is not None else "cpu")   # empty shard: say where
    n_local_max = (n_units + world - 1) // world
    width = torch.tensor([max((int(v.numel()) for v in vectors), default=0)], dtype=torch.int64, device=dev)
    dist.all_reduce(width, op=dist.ReduceOp.MAX, group=group)
    width = int(width.item())
    payload = torch.full((n_local_max, width + 1), float('nan'), dtype=torch.float32, device=dev)
    for i, v in enumerate(vectors):
        payload[i, 0] = float(v.numel())
        payload[i, 1:1 + v.numel()] = v.to(torch.float32).reshape(-1)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    out = [None] * n_units
    for r in range(world):
        for i, u in enumerate(shard_indices(n_units, r, world)):
            n = int(gathered[r][i, 0].item())
            out[u] = gathered[r][i, 1:1 + n].clone()
    return out
