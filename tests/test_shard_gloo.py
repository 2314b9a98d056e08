"""Multi-rank host logic on CPU: world_size 2 over gloo (the data path has no collective; ranks only
agree on shards and reduce timings)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from afec_b200 import shard


def test_shards_cover_and_balance():
    rng = np.random.default_rng(0)
    costs = rng.uniform(0.5, 30.0, size=1000).tolist()
    for world in (1, 2, 4, 8):
        parts = shard.shard_by_cost(costs, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(1000))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= 30.0
    assert shard.shard_by_cost([], 4) == [[], [], [], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [float(1 + (i * 7) % 13) for i in range(101)]
    mine = shard.my_shard(costs, rank, world)
    # every rank derives the same partition without talking; check it by gathering the index sets
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    units = sum(costs[i] for i in mine)
    secs = 1.0 + rank                     # rank 1 is slower: the job time is the max
    agg = shard.aggregate_throughput(units, secs, dist)
    dist.barrier()
    if rank == 0:
        flat = sorted(i for g in gathered for i in g)
        torch.save({"flat": flat, "agg": agg, "total": sum(costs)}, out)
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["flat"] == list(range(101))
    assert abs(r["agg"] - r["total"] / 2.0) < 1e-9
