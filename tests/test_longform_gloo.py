"""Host logic of the time-split long-file path (BASELINE config 5) on CPU: the slice plan, and the all-reduce
callback the C side drives -- here over a world_size-2 gloo group on host memory instead of NCCL on device memory."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_b200 import longform


@pytest.mark.parametrize("n,world,margin", [(691_200_000, 8, 524288), (1_000_003, 3, 8192), (5000, 4, 4096), (4096 * 7 + 1, 2, 0)])
def test_plan_slices_partition(n, world, margin):
    plans = longform.plan_slices(n, world, margin)
    live = [p for p in plans if p is not None]
    assert live[0]["own_start"] == 0 and live[-1]["own_stop"] == n
    for a, b in zip(live, live[1:]):
        assert a["own_stop"] == b["own_start"]                        # owned ranges tile the file
    for p in live:
        assert p["own_start"] % longform.ALIGN == 0 and p["own_lo"] % 4 == 0
        assert 0 <= p["start"] <= p["own_start"] < p["own_stop"] <= p["stop"] <= n
        assert p["own_start"] - p["start"] == (min(margin, p["own_start"]) if p["own_start"] else 0)
        assert p["stop"] - p["own_stop"] == (min(margin, n - p["own_stop"]) if p["own_stop"] < n else 0)
        assert p["own_lo"] == p["own_start"] - p["start"] and p["own_hi"] - p["own_lo"] == p["own_stop"] - p["own_start"]
    assert len(plans) == world and (len(live) == world or n < world * longform.ALIGN)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}
        np_dt = {0: np.float64, 1: np.int64, 2: np.float32}

        def wrap(ptr, count, dtype):      # host memory here; device memory + __cuda_array_interface__ in production
            buf = (C.c_char * (count * np.dtype(np_dt[dtype]).itemsize)).from_address(ptr)
            return torch.from_numpy(np.frombuffer(buf, dtype=np_dt[dtype]))

        cb = longform.make_allreduce(lambda t, op: dist.all_reduce(t, op=ops[op]), wrap)
        # what the C side does at its three reduction points, on this rank's partial values
        sums = np.array([1.5 + rank, -2.0 * rank], dtype=np.float64)
        hops = np.array([10 + rank, 1 << 40, rank], dtype=np.int64)
        peak = np.array([0.25 + 0.5 * rank], dtype=np.float32)
        mins = np.array([-0.1 * (rank + 1)], dtype=np.float64)
        assert cb(None, sums.ctypes.data, sums.size, 0, 0) == 0
        assert cb(None, hops.ctypes.data, hops.size, 1, 0) == 0
        assert cb(None, peak.ctypes.data, peak.size, 2, 2) == 0
        assert cb(None, mins.ctypes.data, mins.size, 0, 1) == 0
        r = np.arange(world)
        ok = np.allclose(sums, [np.sum(1.5 + r), np.sum(-2.0 * r)]) and list(hops) == [int(np.sum(10 + r)), world << 40, int(r.sum())]
        ok = ok and float(peak[0]) == np.float32(0.25 + 0.5 * (world - 1)) and np.isclose(mins[0], -0.1 * world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_allreduce_callback_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
