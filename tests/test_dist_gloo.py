"""world_size-2 `gloo` run of the multi-GPU host logic on the CPU: contiguous batch sharding with no
data-path collective, and the max-over-ranks timing reduction bench.py uses."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = bench.shard_range(13, rank, world)
        owned = torch.zeros(13)
        owned[s:e] = 1
        dist.all_reduce(owned)                       # test-only check: every sample owned exactly once
        slow = bench.max_over_ranks(10.0 + rank, torch.device("cpu"))
        q.put((rank, (s, e), owned.tolist(), slow))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, 29571
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 7) and res[1][1] == (7, 13)
    assert all(r[2] == [1.0] * 13 for r in res)
    assert all(r[3] == 11.0 for r in res)            # max over ranks, identical on every rank
