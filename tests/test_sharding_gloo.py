"""The N > 1 host logic on CPU: world size 2 over gloo (127.0.0.1)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transhuman_b200 import sharding


def _fake_render(idx, C=5):
    """Stand-in for ops.render_rays: a deterministic function of the ray id."""
    i = idx.to(torch.float32)
    return torch.stack([torch.sin(i * 0.01 + c) for c in range(C)], dim=1)


def _worker(rank, world, port, H, W, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = sharding.tile_interleaved_ray_indices(H, W, rank, world, tile)
        img = sharding.gather_rays(_fake_render(idx), idx, H * W)
        want = _fake_render(torch.arange(H * W))
        ok = torch.equal(img, want)
        ms = sharding.max_over_ranks_ms(10.0 + rank, torch.device("cpu"))
        views = sharding.views_for_rank(8, rank, world)
        q.put((rank, ok, ms, len(idx), views))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H,W,tile", [(64, 64, 16), (50, 70, 16), (33, 17, 8)])
def test_two_rank_shard_and_gather(H, W, tile):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + H) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, tile, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "gathered image differs from the single-rank image"
    assert all(r[2] == 11.0 for r in res)                  # max over ranks
    assert sum(r[3] for r in res) == H * W                 # shards partition the rays
    assert res[0][4] == [0, 2, 4, 6] and res[1][4] == [1, 3, 5, 7]


def test_partition_properties_single_process():
    for world in (1, 2, 3, 8):
        seen = torch.zeros(96 * 80, dtype=torch.int32)
        sizes = []
        for r in range(world):
            idx = sharding.tile_interleaved_ray_indices(96, 80, r, world)
            seen[idx] += 1
            sizes.append(len(idx))
        assert torch.all(seen == 1)
        assert max(sizes) - min(sizes) <= 16 * 16 * 2      # balanced to within two tiles
    out = sharding.gather_rays(_fake_render(torch.arange(10)), torch.arange(10), 10)
    assert torch.equal(out, _fake_render(torch.arange(10)))
