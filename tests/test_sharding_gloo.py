"""CPU, world_size 2 over gloo: the N>1 host logic (shard ranges, episode-statistics all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tetris_gymnasium_b200.sharding import allreduce_episode_stats, shard_range, summarize

    start, stop = shard_range(n_total, rank, world)
    # every rank contributes stats that depend on its shard: episodes = #envs, return = sum of global ids
    ids = torch.arange(start, stop, dtype=torch.float64)
    stats = torch.tensor([float(stop - start), float(ids.sum()), 2.0 * (stop - start), float(rank)], dtype=torch.float64)
    allreduce_episode_stats(stats)
    gathered = [None] * world
    dist.all_gather_object(gathered, (start, stop))
    if rank == 0:
        out.put((stats.tolist(), gathered, summarize(stats)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1000, 1001, 7])
def test_shards_partition_and_stats_allreduce(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, ranges, summ = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # contiguous, disjoint, complete partition
    assert ranges[0][0] == 0 and ranges[-1][1] == n_total
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
    assert stats[0] == n_total and stats[1] == n_total * (n_total - 1) / 2 and stats[2] == 2 * n_total and stats[3] == 1.0
    assert abs(summ["mean_length"] - 2.0) < 1e-12


def test_shard_range_properties():
    from tetris_gymnasium_b200.sharding import shard_range

    for n in (0, 1, 5, 16, 1 << 24):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
