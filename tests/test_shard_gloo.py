"""The N>1 path on CPU: world_size-2 (and 4) gloo processes run the shard/all-gather logic of
visma_b200/shard.py and must reproduce the single-rank pose table bit for bit (each object's solve
is independent, so the gathered table cannot depend on the sharding)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


class FakeResult:
    """Deterministic stand-in for a RegistrationResult of object b (no GPU in this test)."""

    def __init__(self, b):
        rng = np.random.default_rng(1000 + b)
        self.transformation_ = rng.normal(size=(4, 4))
        self.fitness_ = float(rng.random())
        self.inlier_rmse_ = float(rng.random())
        self.correspondence_set_ = np.zeros((int(rng.integers(0, 50000)), 2), np.int32)
        self.iterations_ = int(rng.integers(0, 31))


def _worker(rank, world, n_objects, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from visma_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_objects(n_objects, rank, world)
    rows = shard.pack_results([FakeResult(b) for b in mine], n_objects, rank, world)
    table = shard.all_gather_poses(rows, n_objects)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, table))


@pytest.mark.parametrize("world,n_objects", [(2, 32), (2, 7), (4, 5)])
def test_allgather_matches_single_rank(world, n_objects):
    import torch.multiprocessing as mp
    from visma_b200 import shard
    single = shard.unpack_table(shard.pack_results([FakeResult(b) for b in range(n_objects)], n_objects, 0, 1)[None],
                                n_objects, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + n_objects
    procs = [ctx.Process(target=_worker, args=(r, world, n_objects, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, table in got:
        assert table.shape == (n_objects, shard.ROW)
        assert np.array_equal(table, single), rank


def test_shard_partition():
    from visma_b200 import shard
    for world in (1, 2, 4, 8):
        owned = [shard.shard_objects(32, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(32))
        assert max(len(o) for o in owned) == shard.rows_per_rank(32, world)
    assert shard.shard_objects(3, 5, 8) == []
