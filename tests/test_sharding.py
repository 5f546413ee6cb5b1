"""The N > 1 path on CPU: tile -> shard arithmetic against the C ABI's dust_shard_layout, and a
world_size-2 `gloo` run of the exchange step (all-gather of compact shard buffers + un-permute)."""
import os
import socket

import numpy as np
import pytest

from dust_b200 import sharding
from dust_b200.pipeline import shard_layout


@pytest.mark.parametrize("w,h,n", [(1920, 1080, 1), (1920, 1080, 2), (1920, 1080, 8), (3840, 2160, 8), (200, 120, 3), (31, 33, 4), (1, 1, 8)])
def test_layout_matches_c_abi_and_is_a_bijection(w, h, n):
    tiles_x, tiles_y, per, px = sharding.layout(w, h, n)
    assert shard_layout(w, h, n) == (per, px)
    shard, idx = sharding.pixel_to_shard(w, h, n)
    assert shard.min() >= 0 and shard.max() < n and idx.min() >= 0 and idx.max() < px
    flat = shard * px + idx
    assert len(np.unique(flat)) == w * h                                     # no two pixels share a slot
    # round-robin balance: shards differ by at most one tile
    tiles = np.bincount(((np.arange(tiles_x * tiles_y)) % n), minlength=n)
    assert tiles.max() - tiles.min() <= 1 and tiles.max() == per
    frame = np.arange(w * h * 2, dtype=np.float32).reshape(h, w, 2)
    assert np.array_equal(sharding.resolve(sharding.scatter(frame, n), w, h), frame)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, w, h, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "renders" only its own tiles: pixel value = global pixel id, others untouched
        frame = np.arange(w * h * 4, dtype=np.float32).reshape(h, w, 4)
        mine = sharding.scatter(frame, world)[rank]
        mine_t = torch.from_numpy(mine.copy())
        gathered = torch.empty((world,) + mine_t.shape, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered.view(-1), mine_t.view(-1))     # the single exchange step
        out = sharding.resolve(gathered.numpy(), w, h)
        rays = torch.tensor([float((sharding.pixel_to_shard(w, h, world)[0] == rank).sum())])
        dist.all_reduce(rays)                                                # bench.py sums per-rank ray counts the same way
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                             # ... and takes the max time over ranks
        q.put((rank, bool(np.array_equal(out, frame)), float(rays.item()), float(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allgather_resolve():
    import torch.multiprocessing as mp
    w, h, world = 200, 120, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "gathered + resolved frame differs from the single-process frame"
    assert all(r[2] == w * h for r in res) and all(r[3] == 2.0 for r in res)
