"""One long file split in time across GPUs (BASELINE config 5).

Rank r masters a contiguous SLICE of the file: its own frames plus a margin on every cut side that is wide
enough (``mm_slice_margin``) for each recurrence on the chain -- the zero-phase IIR sweeps in both directions,
the K-weighting filters, the de-esser's envelope follower -- to forget that the slice did not start where the
file starts.  The slice goes through the same fused chain as a whole track (``mm_dev_master_slice``); what
couples the ranks are only the chain's global scalars:

* channel sums / minima / maxima   -> DC offset and -0.5 dB guard   (pipeline.py:134-149)   3 x rows float64
* BS.1770 100 ms block square sums -> gated loudness, gain          (pipeline.py:644-664)   rows x hops int64
* output peak                      -> final -0.5 dB guard           (pipeline.py:1899)      1 float32

each reduced over the ranks' OWN frames with an all-reduce in stream order: four per file (sums; negated minima and
maxima together; block sums; peak).  In production these are ``ncclAllReduce`` calls the C side enqueues itself on
the context's stream (``make_exchange``: a communicator created from C, ``csrc/nccl_shim.cu``) -- no Python, no
stream hand-over between two kernels; a callback (``make_allreduce``) serves the tests, where ranks are threads
sharing one GPU or gloo processes on the CPU.  Block sums are 64-bit fixed point, so the loudness -- and with it
every output sample's gain -- does not depend on the number of ranks.  With one rank this is ``mm_dev_master``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

ALIGN = 4096          # owned ranges start on tile-sized boundaries (float4 alignment needs only 4)

_ALLREDUCE_T = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int)


class Slice(C.Structure):
    _fields_ = [("global_n", C.c_int64), ("global_off", C.c_int64), ("own_lo", C.c_int64), ("own_hi", C.c_int64),
                ("allreduce", _ALLREDUCE_T), ("user", C.c_void_p), ("nccl_comm", C.c_void_p)]


def slice_margin(sr: int) -> int:
    return int(_lib.load().mm_slice_margin(int(sr)))


def plan_slices(n: int, world: int, margin: int, align: int = ALIGN) -> list:
    """Cut [0, n) into ``world`` contiguous owned ranges whose starts are multiples of ``align``; each slice is its
    owned range widened by ``margin`` frames on the cut sides (clamped to the file).  Returns, per rank, a dict
    ``start, stop`` (slice in the file), ``own_lo, own_hi`` (owned frames, slice-local), ``own_start, own_stop``.
    Ranks beyond the file's length in ``align`` units get an empty plan (``None``)."""
    if n <= 0 or world <= 0:
        raise ValueError("n and world must be positive")
    units = (n + align - 1) // align
    out = []
    for r in range(world):
        u0, u1 = (units * r) // world, (units * (r + 1)) // world
        a, b = min(u0 * align, n), min(u1 * align, n)
        if a >= b:
            out.append(None)
            continue
        start = max(0, a - margin) if a > 0 else 0
        stop = min(n, b + margin) if b < n else n
        out.append({"start": start, "stop": stop, "own_lo": a - start, "own_hi": b - start, "own_start": a, "own_stop": b})
    return out


_NP_DTYPES = {0: ("<f8", 8), 1: ("<i8", 8), 2: ("<f4", 4)}


class _DevView:
    """Raw device pointer -> ``__cuda_array_interface__`` so torch can wrap it without a copy."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def make_allreduce(reduce_tensor, wrap):
    """ctypes callback for ``mm_slice.allreduce``.  ``wrap(ptr, count, dtype_code)`` turns the buffer into something
    ``reduce_tensor(obj, op_code)`` reduces in place across the ranks (op 0 sum, 1 min, 2 max)."""

    def cb(user, ptr, count, dtype, op):
        try:
            reduce_tensor(wrap(ptr, count, dtype), op)
            return 0
        except Exception as e:  # pragma: no cover - surfaced as an MMError by the C side
            import sys
            print(f"[mm_b200.longform] allreduce failed: {e!r}", file=sys.stderr)
            return 1

    return _ALLREDUCE_T(cb)


def torch_allreduce(stream, device, group=None):
    """The production callback: torch.distributed all-reduce (NCCL) on the engine's stream."""
    import torch
    import torch.distributed as dist
    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}
    tdt = {0: torch.float64, 1: torch.int64, 2: torch.float32}

    def wrap(ptr, count, dtype):
        t = torch.as_tensor(_DevView(ptr, count, _NP_DTYPES[dtype][0]), device=device)
        assert t.dtype == tdt[dtype] and t.data_ptr() == int(ptr)
        return t

    def red(t, op):
        with torch.cuda.stream(stream):
            dist.all_reduce(t, op=ops[op], group=group)

    return make_allreduce(red, wrap)


class Exchange:
    """The exchange step of one rank: an NCCL communicator owned by the C library (``mm_nccl_comm_create``), or -- ``comm``
    None -- a Python all-reduce callback.  ``slice_struct`` fills the ``mm_slice`` either way."""

    def __init__(self, comm=None, callback=None, note=""):
        self.comm, self.callback, self.note = comm, callback, note

    def slice_struct(self, global_n, plan):
        return Slice(int(global_n), int(plan["start"]), int(plan["own_lo"]), int(plan["own_hi"]),
                     self.callback if (self.callback is not None and not self.comm) else _ALLREDUCE_T(), None,
                     self.comm if self.comm else None)

    def describe(self):
        return self.note

    def close(self):
        if self.comm:
            _lib.load().mm_nccl_comm_destroy(self.comm)
            self.comm = None


def make_exchange(eng, world: int, rank: int, group=None, direct: bool = True) -> Exchange:
    """Production exchange under torch.distributed (one process per GPU).  ``direct``: the C library opens its own NCCL
    communicator (rank 0's unique id travels by a torch.distributed broadcast, once) and issues the all-reduces itself;
    otherwise (or when libnccl cannot be loaded) torch.distributed's all-reduce through the callback."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    if world <= 1:
        return Exchange(note="single rank: none")
    if direct and lib.mm_nccl_version() > 0:
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.mm_nccl_unique_id(buf))
        t = torch.tensor(list(buf), dtype=torch.uint8, device=eng.tdev)
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        comm = C.c_void_p()
        _lib.check(lib.mm_nccl_comm_create(eng.ctx, raw, int(world), int(rank), C.byref(comm)))
        return Exchange(comm=comm, note=f"4 ncclAllReduce per file enqueued from C on the chain's stream (NCCL {lib.mm_nccl_version()})")
    return Exchange(callback=torch_allreduce(eng.stream, eng.tdev, group), note="torch.distributed all_reduce through the mm_slice callback")


def master_slice(eng, x_slice: np.ndarray, sr: int, plan: dict, global_n: int, style: dict, target_lufs: float, chain: str = "v2",
                 *, allreduce=None, exchange: Exchange = None, want_int16: bool = False, seed: int = 0, measure: bool = False, src=None):
    """Master one rank's slice.  ``x_slice``: (slice_frames, ch) float32 host array (or None with a device
    ``src`` Batch already holding it).  Returns dict(audio=(own, ch) float32, pcm=int16 or None, stats=record)."""
    import torch
    from .engine import style_struct, TrackStats
    from .shard import stats_to_records
    b = src if src is not None else eng.upload([np.asarray(x_slice, dtype=np.float32)], sr)
    dst = eng.like(b)
    g = b.geom
    if exchange is None:
        exchange = Exchange(callback=allreduce)
    sl = exchange.slice_struct(global_n, plan)
    st_arr = (_lib.Style * 1)(style_struct(style, target_lufs))
    flags = (_lib.FLAG_MEASURE_IN | _lib.FLAG_MEASURE_OUT) if measure else 0
    with torch.cuda.stream(eng.stream):
        pcm = torch.empty((1, b.n, b.channels), dtype=torch.int16, device=eng.tdev) if want_int16 else None
        st = torch.empty(C.sizeof(TrackStats), dtype=torch.uint8, device=eng.tdev)
        _lib.check(eng.lib.mm_dev_master_slice(
            eng.ctx, C.byref(g), _lib.CHAIN_V1 if chain == "v1" else _lib.CHAIN_V2, st_arr, b.ptr, dst.ptr,
            C.c_void_p(pcm.data_ptr()) if pcm is not None else None, None, int(seed), C.c_void_p(st.data_ptr()), int(flags),
            C.byref(sl)))
        eng.sync()
        lo, hi = int(plan["own_lo"]), int(plan["own_hi"])
        audio = dst.live()[:, lo:hi].t().contiguous().cpu().numpy()
        pcm_np = pcm[0, lo:hi].cpu().numpy() if pcm is not None else None
        rec = stats_to_records(np.frombuffer(st.cpu().numpy().tobytes(), dtype=np.float64).reshape(1, -1))[0]
    return {"audio": audio, "pcm": pcm_np, "stats": rec, "device_out": dst}


def master_long_file(x: np.ndarray, sr: int, style: dict, target_lufs: float, chain: str = "v2", *, world: int = 1, rank: int = 0,
                     eng=None, group=None, want_int16: bool = False, seed: int = 0, measure: bool = False, margin=None):
    """Time-split mastering of one file under torch.distributed (one process per GPU): every rank passes the same
    ``x`` (or at least its own slice of it: only ``x[start:stop]`` is read) and gets its OWNED part back together
    with its place in the file, ``(own_start, own_stop)``."""
    from .engine import get_engine
    eng = eng or get_engine()
    n = int(x.shape[0])
    margin = slice_margin(sr) if margin is None else int(margin)
    plan = plan_slices(n, world, margin)[rank]
    if plan is None:
        raise ValueError(f"rank {rank} of {world} owns no frames of a {n}-frame file")
    xch = make_exchange(eng, world, rank, group)
    try:
        res = master_slice(eng, x[plan["start"]:plan["stop"]], sr, plan, n, style, target_lufs, chain, exchange=xch,
                           want_int16=want_int16, seed=seed, measure=measure)
    finally:
        xch.close()
    res["own_start"], res["own_stop"] = plan["own_start"], plan["own_stop"]
    return res
