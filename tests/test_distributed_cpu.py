"""world_size-2 gloo test of the N>1 host path: contiguous sharding, shard-independent per-env seeds,
the gather of episode returns and the max-over-ranks timing reduction."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
    from pybullet_robot_envs.b2env import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = 11  # ragged on purpose
    lo, hi = shard.shard_range(G, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32) * 10
    allr = shard.gather_returns(local)
    t = shard.max_over_ranks(1.0 + rank, torch.device("cpu"))
    seeds = [shard.env_seed(100, i) for i in range(lo, hi)]
    q.put((rank, lo, hi, allr.tolist(), t, seeds))
    dist.destroy_process_group()


def test_shard_gather_gloo_world2():
    world, port = 2, 29517 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, a0, t0, s0), (r1, lo1, hi1, a1, t1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 6, 6, 11)
    assert a0 == a1 == [10.0 * i for i in range(11)]
    assert t0 == t1 == 2.0
    assert s0 + s1 == list(range(100, 111))


def test_shard_range_covers_everything():
    sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
    from pybullet_robot_envs.b2env.shard import shard_range
    for G in (1, 7, 16384, 65536, 131072):
        for W in (1, 2, 4, 8):
            spans = [shard_range(G, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
