"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU sweep (sharding with halo, the single all-gather)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ball_action_spotting_b200.indexes import StackIndexesGenerator
from ball_action_spotting_b200.sweep import frames_needed, gather_predictions, prediction_bounds, shard_range


def test_shards_partition_the_prediction_range_and_carry_the_halo():
    gen = StackIndexesGenerator(15, 2)
    lo, hi = prediction_bounds(gen, 67500, 1)
    assert (lo, hi) == (15, 67500 - 15 - 1)
    for world in (1, 2, 3, 8):
        shards = [shard_range(lo, hi, r, world) for r in range(world)]
        assert shards[0][0] == lo and shards[-1][1] == hi + 1
        assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in shards]
        assert max(sizes) - min(sizes) <= 1
        for a, b in shards:
            f0, f1 = frames_needed(gen, a, b)
            assert f0 == a - 14 and f1 == b - 1 + 14 and f0 >= 0 and f1 < 67500     # 28-frame halo (predictors.py:36)
    assert shard_range(5, 4, 0, 2) == (5, 5)        # empty range


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = 15, 15 + 10          # 11 predictions -> shards of 6 and 5
        a, b = shard_range(lo, hi, rank, world)
        local = torch.stack([torch.arange(a, b, dtype=torch.float32), torch.arange(a, b, dtype=torch.float32) * 10], 1)
        counts = [shard_range(lo, hi, r, world)[1] - shard_range(lo, hi, r, world)[0] for r in range(world)]
        full = gather_predictions(local, counts)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


def test_all_gather_of_ragged_shards_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[float(i), float(i) * 10] for i in range(15, 26)]
    assert res[0] == want and res[1] == want
