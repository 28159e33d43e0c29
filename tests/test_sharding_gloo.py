"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo groups exercising the track sharding and
the stats all-gather that bench.py / production use over NCCL (tracks shard with no
data-path collective; only fixed-size per-track stats are exchanged)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_tracks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.shard_tracks(n_tracks, world, rank)
        # fake per-track stats: every field carries the global track id so ordering errors show up
        local = torch.tensor([[t + 0.001 * k for k in range(shard.STATS_DOUBLES)] for t in mine], dtype=torch.float64).reshape(
            len(mine), shard.STATS_DOUBLES)
        allst = shard.gather_track_stats(local, n_tracks, world, rank)
        ok = allst.shape == (n_tracks, shard.STATS_DOUBLES) and all(abs(float(allst[t, 0]) - t) < 1e-12 for t in range(n_tracks))
        ok = ok and all(abs(float(allst[t, 5]) - (t + 0.005)) < 1e-12 for t in range(n_tracks))
        q.put((rank, bool(ok), mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_tracks", [(2, 8), (2, 7), (3, 10)])
def test_sharding_and_stats_gather(world, n_tracks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tracks, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(t for _, _, mine in res for t in mine)
    assert owned == list(range(n_tracks))                 # every track mastered exactly once
    assert all(ok for _, ok, _ in res)


def test_shard_helpers():
    assert shard.shard_tracks(10, 4, 1) == [1, 5, 9]
    assert sum(shard.local_count(1024, 8, r) for r in range(8)) == 1024
    recs = shard.stats_to_records(np.arange(2 * shard.STATS_DOUBLES, dtype=np.float64).reshape(2, -1))
    assert set(recs[0]) >= {"lufs_in", "lufs_out", "gain_db", "peak_in", "peak_out", "mean0", "mean1", "nonfinite"}
