"""Track sharding across the GPUs of one box: tracks are independent units, so rank r
masters tracks r, r + G, r + 2G, ... with no data-path collective; the only exchange is an all-gather
of the fixed-size per-track stats records (loudness before/after, gain, peaks) -- NCCL over NVLink on
GPUs, gloo in the CPU tests of this host logic."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import TrackStats

STATS_DOUBLES = C.sizeof(TrackStats) // 8


def shard_tracks(n_tracks: int, world: int, rank: int) -> list:
    """Global track indices owned by ``rank`` (track t -> rank t % world)."""
    return list(range(rank, n_tracks, world))


def local_count(n_tracks: int, world: int, rank: int) -> int:
    return len(range(rank, n_tracks, world))


def gather_track_stats(local, n_tracks: int, world: int, rank: int, group=None):
    """All-gather per-track stats.

    ``local``: float64 tensor ``(local_count, STATS_DOUBLES)`` on the rank's device (cuda for NCCL, cpu
    for gloo).  Returns a tensor ``(n_tracks, STATS_DOUBLES)`` in global track order on the same device.
    Ranks may own different numbers of tracks; shards are padded to the largest one for the collective.
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    cap = local_count(n_tracks, world, 0)
    buf = torch.zeros((cap, STATS_DOUBLES), dtype=torch.float64, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = torch.empty((n_tracks, STATS_DOUBLES), dtype=torch.float64, device=local.device)
    for r in range(world):
        idx = shard_tracks(n_tracks, world, r)
        out[idx] = parts[r][: len(idx)]
    return out


def stats_to_records(t) -> list:
    """(n, STATS_DOUBLES) float64 tensor/array -> list of dicts with the field names of mm_track_stats."""
    a = np.asarray(t.cpu() if hasattr(t, "cpu") else t, dtype=np.float64)
    names = []
    for name, typ in TrackStats._fields_:
        k = C.sizeof(typ) // 8
        names += [name] if k == 1 else [f"{name}{i}" for i in range(k)]
    return [dict(zip(names, row)) for row in a]
