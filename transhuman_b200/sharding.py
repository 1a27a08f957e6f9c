"""Ray / view partitioning across ranks.

Rays are independent units: the query path needs no collective.  The only
exchange is the final image gather (`rgb, acc, depth` = 5 floats per ray).
These helpers are device agnostic (they run under `gloo` on CPU in the tests
and under `nccl` on the GPUs); the rendering itself is `ops.render_rays`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def tile_interleaved_ray_indices(H: int, W: int, rank: int, world: int, tile: int = 16) -> torch.Tensor:
    """Indices (row-major pixel ids, sorted) of the rays rank `rank` renders when
    one H x W view is sharded as interleaved `tile` x `tile` pixel tiles, round
    robin over ranks -- keeps the culled workload balanced (the body is centred)."""
    ty = torch.arange(H) // tile
    tx = torch.arange(W) // tile
    tiles_x = (W + tile - 1) // tile
    owner = (ty[:, None] * tiles_x + tx[None, :]) % world
    return torch.nonzero(owner.reshape(-1) == rank)[:, 0]


def views_for_rank(n_views: int, rank: int, world: int) -> list:
    """Target views rank `rank` renders when a batch of views is sharded view-per-rank."""
    return list(range(rank, n_views, world))


def gather_rays(local: torch.Tensor, idx: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-ray results.  `local` (n_local, C) are this rank's rows for
    the global ray ids `idx` (n_local,).  Returns the (n_total, C) image on every
    rank.  Shards may be ragged: they are padded to the largest shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = local.new_zeros((n_total, local.shape[1]))
        out[idx] = local
        return out
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    n_max = int(max(s.item() for s in sizes))
    pad_v = local.new_zeros((n_max, local.shape[1]))
    pad_v[: local.shape[0]] = local
    pad_i = torch.full((n_max,), -1, dtype=torch.int64, device=local.device)
    pad_i[: idx.shape[0]] = idx.to(local.device)
    all_v = [torch.empty_like(pad_v) for _ in range(world)]
    all_i = [torch.empty_like(pad_i) for _ in range(world)]
    dist.all_gather(all_v, pad_v, group=group)
    dist.all_gather(all_i, pad_i, group=group)
    out = local.new_zeros((n_total, local.shape[1]))
    for v, i, s in zip(all_v, all_i, sizes):
        n = int(s.item())
        out[i[:n]] = v[:n]
    return out


def max_over_ranks_ms(ms: float, device, group=None) -> float:
    """Device time of a step = the slowest rank's."""
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
