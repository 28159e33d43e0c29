#!/usr/bin/env python
"""Benchmark of the mastering hot path (BASELINE.json metric: mastered audio-seconds per wall-second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--chain v2|v1]

One step = one pass of the full mastering chain (+ TPDF dither to int16) over a batch of synthetic
tracks that is already resident in HBM.  At N = 1 the workload is BASELINE.json configs[1]:
64 synthetic 3-minute 44.1 kHz stereo tracks.  With N > 1 (torchrun, one rank per GPU) every rank
masters its own 64 tracks (sharded by track, weak scaling) and the per-track loudness/peak stats are
all-gathered over NCCL inside the timed region.  Prints ONE JSON line on rank 0.

--impl reference times the reference chain's CPU restatement (oracle/, numpy/scipy -- the reference
itself is not on the GPU box) on the host cores, one track per process, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "audio-mastering-web_b200"))

import numpy as np  # noqa: E402

SR = 44100
DUR = 180.0
TRACKS = 64
METRIC = "mastered audio-seconds per wall-second (full chain + 16-bit TPDF export)"
UNIT = "audio-s/s"

# algorithmic fp32 words moved per channel-sample by each kernel (DESIGN.md "Kernels")
STREAMS = {
    "sweep_fwd_m2_f1_i1": 2, "sweep_bwd_m2_f1_store": 2, "sweep_bwd_m2_f1_combine": 3, "sweep_bwd_m2_f1_exciter": 3,
    "sweep_fwd_m2_f2_i1": 3, "sweep_fwd_m2_f2_i2": 4, "sweep_bwd_m2_f2_store": 4, "sweep_bwd_m2_f2_combine": 4,
    "sweep_bwd_m2_f2_dynamics": 5, "sweep_fwd_m2_f4_i1": 5, "sweep_bwd_m2_f4_combine": 6,
    "sweep_fwd_m4_f1_i1": 2, "sweep_bwd_m4_f1_store": 2, "envelope_gain": 2, "deesser_smooth_apply": 4,
    "lufs_kweight_blocks": 1, "row_stats": 1, "finalize_dither_int16": 2.5, "finalize": 2, "peak_after_imager": 1,
}
CHAIN_BYTES_PER_FRAME = {"v2": 436.0, "v1": 516.0}     # style "standard", stereo


def measured_peak_gbs():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class _StdoutToStderr:
    """NCCL prints its version banner to stdout when it initialises; the contract is ONE JSON line there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def _init_nccl(local):
    import torch
    import torch.distributed as dist
    with _StdoutToStderr():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.zeros(1, device=torch.device("cuda", local))
        dist.all_reduce(t)                 # communicator creation (and its banner) happens here
        torch.cuda.synchronize()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (test infrastructure) timed as the reference's CPU path
# -------------------------------------------------------------------------------------------------------
def _cpu_one(args):
    t, dur, chain = args
    from oracle import chain as oc
    from mm_b200 import synth
    x = synth.numpy_track(t, SR, dur)
    t0 = time.time()
    out = (oc.run_v1 if chain == "v1" else oc.run_v2)(x, SR, -14.0, "standard")
    rng = np.random.default_rng(t)
    noise = (rng.random(out.shape) + rng.random(out.shape) - 1.0).astype(np.float32)
    oc.quantize_int16(out, noise)
    return time.time() - t0


def cpu_baseline(chain, dur=60.0, procs=1, tracks=None):
    """oracle port on `procs` host processes, one track each; returns audio-s/s and a description."""
    import multiprocessing as mp
    tracks = tracks or procs
    work = [(1000 + i, dur, chain) for i in range(tracks)]
    if procs == 1:
        wall = sum(_cpu_one(w) for w in work)
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            per = pool.map(_cpu_one, work)
        waves = (tracks + procs - 1) // procs
        wall = max(per) * waves          # chain + export time of the slowest worker (synthesis excluded)
    return tracks * dur / wall, f"{tracks} synthetic tracks x {dur:.0f} s, 44.1 kHz stereo, {chain} chain 'standard' + TPDF int16, oracle (numpy/scipy)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    dur = 30.0
    for _ in range(args.warmup):
        cpu_baseline(args.chain, dur=5.0, procs=procs)
    vals = []
    for _ in range(args.steps):
        val, sample = cpu_baseline(args.chain, dur=dur, procs=procs)
        vals.append(val)
    value = len(vals) / sum(1.0 / v for v in vals)      # total audio / total time
    wall = args.steps * procs * dur / value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bounded sample of configs[1]: per step {procs} tracks x {dur:.0f} s 44.1 kHz stereo, "
                               f"{args.chain} chain + TPDF int16, one track per host process", "chain": args.chain},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from mm_b200 import _lib, pipeline as P, shard, synth
    from mm_b200.engine import Engine, style_struct, TrackStats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (mm_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        _init_nccl(local)
    eng = Engine(local)
    tracks, dur, sr = args.tracks, args.sec, SR
    mixed = args.workload == "mixed"
    if mixed:      # BASELINE configs[2]: 48 kHz, genre presets cycling over the GLOBAL track index, 128 tracks per GPU
        sr = 48000
        tracks = args.tracks if args.tracks != TRACKS else 128
    n = int(round(sr * dur))
    chain = _lib.CHAIN_V1 if args.chain == "v1" else _lib.CHAIN_V2

    # synthetic batch, generated on the device (SURVEY 8d generator), resident in HBM before timing
    src = eng.empty(tracks, 2, n, sr)
    ids = shard.shard_tracks(world * tracks, world, rank)      # track t -> rank t % world
    with torch.cuda.stream(eng.stream):
        src.t.zero_()
        synth.torch_batch(ids, sr, dur, eng.tdev, out=src.t, row_stride=src.stride, lead=_lib.MM_LEAD)
    out = eng.like(src)
    names = list(P.STYLE_CONFIGS)
    style_names = [names[t % len(names)] if mixed else "standard" for t in ids]
    styles = [style_struct(P.STYLE_CONFIGS[s], P.STYLE_CONFIGS[s]["lufs"] if mixed else -14.0) for s in style_names]
    # algorithmic bytes per stereo frame of each track: 8 (W_base + 5 n_style_bands + 5 [exciter fires])
    wbase = 64.5 if args.chain == "v1" else 54.5
    def _alg_bytes(sn):
        cfg = P.STYLE_CONFIGS[sn]
        nb = sum(1 for k in ("sub", "bass", "mids", "presence", "air") if abs(cfg.get(k, 0.0)) >= 0.05)
        exc = (cfg.get("exciter_db", 0.0) > 0.05) if args.chain == "v1" else (abs(cfg.get("exciter_db", 0.0)) >= 0.05)
        return 8.0 * (wbase + 5 * nb + 5 * (1 if exc else 0))
    alg_per_frame = float(np.mean([_alg_bytes(sn) for sn in style_names]))
    arr = (_lib.Style * tracks)(*styles)
    with torch.cuda.stream(eng.stream):
        pcm = torch.empty((tracks, n, 2), dtype=torch.int16, device=eng.tdev)
        stats = torch.empty((tracks, shard.STATS_DOUBLES), dtype=torch.float64, device=eng.tdev)
    g = src.geom
    flags = _lib.FLAG_MEASURE_OUT

    def step(i):
        _lib.check(eng.lib.mm_dev_master(eng.ctx, C.byref(g), chain, arr, src.ptr, out.ptr, C.c_void_p(pcm.data_ptr()), None,
                                         1234 + i, C.c_void_p(stats.data_ptr()), flags))
        if world > 1:      # the only exchange of the sharded path: per-track stats records over NCCL
            with torch.cuda.stream(eng.stream):
                shard.gather_track_stats(stats, world * tracks, world, rank)

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(eng.stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    ktimes = eng.kernel_times()
    eng.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    audio_s = world * tracks * dur * args.steps
    value = audio_s / (ms * 1e-3)

    # sanity of the timed work: every track was mastered to its target within the gate
    eng.sync()
    recs = shard.stats_to_records(stats)
    lufs_out = np.array([r["lufs_out"] for r in recs])
    nonfinite = float(sum(r["nonfinite"] for r in recs))

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory, copies inside the timing) ----
    e2e = None
    if rank == 0 or world > 1:
        e_tracks = min(tracks, args.e2e_tracks)
        frames = e_tracks * n * 2
        # pinned host buffers in the reference's own layout: float32 (n, 2) interleaved per track
        hin = torch.empty((e_tracks, n, 2), dtype=torch.float32, pin_memory=True)
        hpcm = torch.empty((e_tracks, n, 2), dtype=torch.int16, pin_memory=True)
        hstats = (TrackStats * e_tracks)()
        with torch.cuda.stream(eng.stream):
            il = torch.empty((e_tracks, n, 2), dtype=torch.float32, device=eng.tdev)
            ge = _lib.Geom(n, src.stride, e_tracks, 2, sr, 0)
            _lib.check(eng.lib.mm_dev_interleave(eng.ctx, C.byref(ge), src.ptr, C.c_void_p(il.data_ptr())))
            eng.sync()
            hin.copy_(il)            # setup, untimed
            del il
        torch.cuda.synchronize()
        earr = (_lib.Style * e_tracks)(*styles[:e_tracks])
        e_sr = sr

        def e2e_step(i):
            _lib.check(eng.lib.mm_master_host(eng.ctx, chain, e_tracks, n, 2, sr, earr, C.c_void_p(hin.data_ptr()), None,
                                              C.c_void_p(hpcm.data_ptr()), None, 99 + i, hstats, flags))

        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for i in range(reps):
            e2e_step(1 + i)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall], dtype=torch.float64, device=eng.tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        e2e = {"value": world * e_tracks * dur * reps / wall, "unit": UNIT, "h2d_bytes_per_step": frames * 4,
               "d2h_bytes_per_step": frames * 2 + e_tracks * C.sizeof(TrackStats), "tracks_per_step": e_tracks,
               "api": "mm_master_host (C ABI, pinned host buffers in/out)"}
        # job-level variant: the upload's PCM_16 frames cross PCIe as they are (informative; the headline stays float32 in)
        hin16 = torch.empty((e_tracks, n, 2), dtype=torch.int16, pin_memory=True)
        hin16.copy_((hin * 32767.0).round().to(torch.int16))
        del hin

        def e2e16_step(i):
            _lib.check(eng.lib.mm_master_host_pcm16(eng.ctx, chain, e_tracks, n, 2, sr, earr, C.c_void_p(hin16.data_ptr()), None,
                                                    C.c_void_p(hpcm.data_ptr()), 199 + i, hstats, flags))

        e2e16_step(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(reps):
            e2e16_step(1 + i)
        torch.cuda.synchronize()
        wall16 = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall16], dtype=torch.float64, device=eng.tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall16 = float(t.item())
        e2e["pcm16_in"] = {"value": world * e_tracks * dur * reps / wall16, "unit": UNIT, "h2d_bytes_per_step": frames * 2,
                           "d2h_bytes_per_step": frames * 2 + e_tracks * C.sizeof(TrackStats),
                           "api": "mm_master_host_pcm16 (PCM_16 frames in and out, widened on the device)"}
        del hin16, hpcm

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    top = max(ktimes.items(), key=lambda kv: kv[1][0])
    kname, (kms, kcnt) = top
    rows_n = tracks * 2 * n
    alg_bytes = STREAMS.get(kname, 2) * 4.0 * rows_n
    achieved = alg_bytes / (kms / kcnt * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(kname)
        except Exception:
            traffic = None
    ksum = sum(v[0] for v in ktimes.values())
    chain_gbs = alg_per_frame * tracks * n * world * args.steps / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_share_of_step": kms / ksum,
                "chain": {"algorithmic_bytes_per_stereo_frame": alg_per_frame, "achieved": chain_gbs / world,
                          "frac": chain_gbs / world / peak},
                "kernels": {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / args.steps,
                                "gbs": STREAMS.get(k, 0) * 4.0 * rows_n / (v[0] / v[1] * 1e-3) / 1e9} for k, v in
                            sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:12]}}

    cb = None
    if world == 1 and not args.no_cpu:
        v, sample = cpu_baseline(args.chain, dur=180.0, procs=1, tracks=1)
        cb = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64/f32", "data": "synthetic",
        "config": {"workload": (f"configs[2]: {tracks} synthetic {dur:.0f} s {sr} Hz stereo tracks per GPU, genre presets cycling over the "
                                f"global track index (STYLE_CONFIGS order) at their own LUFS targets, {args.chain} chain + TPDF int16 + after-LUFS"
                                if mixed else
                                f"configs[1]: {tracks} synthetic {dur:.0f} s {sr} Hz stereo tracks per GPU, {args.chain} default chain "
                                f"(style standard, -14 LUFS) + TPDF dither to int16 + after-LUFS"),
                   "chain": args.chain, "tracks_per_gpu": tracks, "frames_per_track": n,
                   "cache": f"inputs ({tracks * n * 8 / 1e9:.2f} GB per GPU) exceed L2; no flush needed", "storage": "float32 streams; float64 chunk scan everywhere; in-chunk recurrences float64 (full-path low cut-offs) or float32 FFMA2 on balanced realizations (DESIGN.md precision policy)"},
        "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "check": {"lufs_out_mean": float(np.mean(lufs_out)), "lufs_out_min": float(np.min(lufs_out)),
                  "lufs_out_max": float(np.max(lufs_out)), "nonfinite": nonfinite},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: one 2-hour 96 kHz stereo file, time-split over the ranks (strong scaling)
# -------------------------------------------------------------------------------------------------------
def run_longform(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from mm_b200 import _lib, longform, pipeline as P, synth
    from mm_b200.engine import Engine, style_struct, TrackStats
    from mm_b200.shard import stats_to_records

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (mm_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        _init_nccl(local)
    eng = Engine(local)
    sr, dur = 96000, args.sec if args.sec != DUR else 7200.0
    n = int(round(sr * dur))
    plan = longform.plan_slices(n, world, longform.slice_margin(sr))[rank]
    ns = plan["stop"] - plan["start"]
    src = eng.empty(1, 2, ns, sr)
    with torch.cuda.stream(eng.stream):
        src.t.zero_()
        synth.torch_long_slice(0, sr, plan["start"], plan["stop"], eng.tdev, src.t, lead=_lib.MM_LEAD)
        pcm = torch.empty((1, ns, 2), dtype=torch.int16, device=eng.tdev)
        st = torch.empty(C.sizeof(TrackStats), dtype=torch.uint8, device=eng.tdev)
    out = eng.like(src)
    cb = longform.torch_allreduce(eng.stream, eng.tdev) if world > 1 else longform._ALLREDUCE_T()
    sl = longform.Slice(n, plan["start"], plan["own_lo"], plan["own_hi"], cb, None)
    style = (_lib.Style * 1)(style_struct(P.STYLE_CONFIGS["standard"], -14.0))
    chain = _lib.CHAIN_V1 if args.chain == "v1" else _lib.CHAIN_V2
    g = src.geom

    def step(i):
        _lib.check(eng.lib.mm_dev_master_slice(eng.ctx, C.byref(g), chain, style, src.ptr, out.ptr, C.c_void_p(pcm.data_ptr()), None,
                                               1234 + i, C.c_void_p(st.data_ptr()), _lib.FLAG_MEASURE_OUT, C.byref(sl)))

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(eng.stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    ktimes = eng.kernel_times()
    eng.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = dur * args.steps / (ms * 1e-3)
    rec = stats_to_records(np.frombuffer(st.cpu().numpy().tobytes(), dtype=np.float64).reshape(1, -1))[0]

    # end to end: the rank's slice from pinned host memory, its own frames' int16 back to pinned host memory
    own = plan["own_hi"] - plan["own_lo"]
    hin = torch.empty((ns, 2), dtype=torch.float32, pin_memory=True)
    hpcm = torch.empty((own, 2), dtype=torch.int16, pin_memory=True)
    with torch.cuda.stream(eng.stream):
        il = torch.empty((1, ns, 2), dtype=torch.float32, device=eng.tdev)
        _lib.check(eng.lib.mm_dev_interleave(eng.ctx, C.byref(g), src.ptr, C.c_void_p(il.data_ptr())))
        eng.sync()
        hin.copy_(il[0])

        def e2e_step(i):
            il[0].copy_(hin, non_blocking=True)
            _lib.check(eng.lib.mm_dev_deinterleave(eng.ctx, C.byref(g), C.c_void_p(il.data_ptr()), src.ptr))
            step(100 + i)
            hpcm.copy_(pcm[0, plan["own_lo"]:plan["own_hi"]], non_blocking=True)
            eng.sync()

        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 2))
        for i in range(reps):
            e2e_step(1 + i)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=eng.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    e2e = {"value": dur * reps / wall, "unit": UNIT, "h2d_bytes_per_step": ns * 2 * 4, "d2h_bytes_per_step": own * 2 * 2,
           "api": "mm_dev_master_slice on a slice copied from / to pinned host memory (per rank)"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak_gbs()
    top = max(ktimes.items(), key=lambda kv: kv[1][0])
    kname, (kms, kcnt) = top
    alg_bytes = STREAMS.get(kname, 2) * 4.0 * 2 * ns
    chain_gbs = CHAIN_BYTES_PER_FRAME[args.chain] * ns * args.steps / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": alg_bytes / (kms / kcnt * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes / (kms / kcnt * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "chain": {"algorithmic_bytes_per_stereo_frame": CHAIN_BYTES_PER_FRAME[args.chain], "achieved": chain_gbs,
                          "frac": chain_gbs / peak, "note": "rank 0: its slice including margins"}}
    cb_line = None
    if world == 1 and not args.no_cpu:
        from oracle import chain as oc
        x = synth.numpy_track(0, sr, 20.0)
        t0 = time.time()
        o = (oc.run_v1 if args.chain == "v1" else oc.run_v2)(x, sr, -14.0, "standard")
        rng = np.random.default_rng(0)
        oc.quantize_int16(o, (rng.random(o.shape) + rng.random(o.shape) - 1.0).astype(np.float32))
        cb_line = {"value": 20.0 / (time.time() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "20 s of 96 kHz stereo through the oracle (numpy/scipy), v2 chain + TPDF int16"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64/f32",
        "data": "synthetic",
        "config": {"workload": f"configs[4]: one {dur:.0f} s {sr} Hz stereo file, {args.chain} default chain + TPDF int16 + after-LUFS, "
                               f"split in time over {world} GPU(s) ({longform.slice_margin(sr)} margin frames per cut side)",
                   "chain": args.chain, "frames": n, "slice_frames_rank0": ns, "cache": "slice (GBs) exceeds L2; no flush needed"},
        "roofline": roofline, "cpu_baseline": cb_line, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "check": {"lufs_out": rec["lufs_out"], "gain_db": rec["gain_db"], "peak_out": rec["peak_out"], "nonfinite": rec["nonfinite"]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: analyzer-only path over 30 s clips (LUFS + gating, 4x true peak, correlation, spectrum bars)
# -------------------------------------------------------------------------------------------------------
def run_analyze(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from mm_b200 import _lib, shard, synth
    from mm_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (mm_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        _init_nccl(local)
    eng = Engine(local)
    sr, dur = SR, args.sec if args.sec != DUR else 30.0
    total = args.tracks if args.tracks != TRACKS else 10000
    clips = shard.local_count(total, world, rank)
    sub = min(clips, 2500)                         # clips per device batch: 26.5 GB resident
    n = int(round(sr * dur))
    src = eng.empty(sub, 2, n, sr)
    distinct = min(sub, 64)
    with torch.cuda.stream(eng.stream):
        src.t.zero_()
        synth.torch_batch(list(range(distinct)), sr, dur, eng.tdev, out=src.t, row_stride=src.stride, lead=_lib.MM_LEAD)
        for k in range(distinct, sub, distinct):   # fill the batch with copies of the 64 distinct clips
            m = min(distinct, sub - k)
            src.t[2 * k:2 * (k + m)].copy_(src.t[:2 * m])
        res = {k: torch.empty(sub * w, dtype=torch.float64, device=eng.tdev) for k, w in
               (("lufs", 1), ("tp", 1), ("corr", 1), ("peak", 1), ("bars0", 64), ("bars1", 64), ("bars2", 64))}
    g = src.geom
    nbatch = (clips + sub - 1) // sub
    ptr = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731

    def step(i):
        for _ in range(nbatch):
            _lib.check(eng.lib.mm_dev_measure_lufs(eng.ctx, C.byref(g), src.ptr, ptr(res["lufs"])))
            _lib.check(eng.lib.mm_dev_true_peak_correlation(eng.ctx, C.byref(g), src.ptr, ptr(res["tp"]), ptr(res["corr"]), ptr(res["peak"])))
            for v in range(3):
                _lib.check(eng.lib.mm_dev_spectrum_bars(eng.ctx, C.byref(g), src.ptr, v, ptr(res[f"bars{v}"])))

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for i in range(args.steps):
        step(i)
    e1.record(eng.stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    ktimes = eng.kernel_times()
    eng.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    done = nbatch * sub * world
    value = done * dur * args.steps / (ms * 1e-3)
    # end to end: one sub-batch of clips from pinned host memory, results back to the host
    hin = torch.empty((sub, n, 2), dtype=torch.float32, pin_memory=True)
    with torch.cuda.stream(eng.stream):
        il = torch.empty((sub, n, 2), dtype=torch.float32, device=eng.tdev)
        _lib.check(eng.lib.mm_dev_interleave(eng.ctx, C.byref(g), src.ptr, ptr(il)))
        eng.sync()
        hin.copy_(il)
        hres = {k: torch.empty(v.shape, dtype=torch.float64, pin_memory=True) for k, v in res.items()}

        def e2e_step():
            il.copy_(hin, non_blocking=True)
            _lib.check(eng.lib.mm_dev_deinterleave(eng.ctx, C.byref(g), ptr(il), src.ptr))
            _lib.check(eng.lib.mm_dev_measure_lufs(eng.ctx, C.byref(g), src.ptr, ptr(res["lufs"])))
            _lib.check(eng.lib.mm_dev_true_peak_correlation(eng.ctx, C.byref(g), src.ptr, ptr(res["tp"]), ptr(res["corr"]), ptr(res["peak"])))
            for v in range(3):
                _lib.check(eng.lib.mm_dev_spectrum_bars(eng.ctx, C.byref(g), src.ptr, v, ptr(res[f"bars{v}"])))
            for k in res:
                hres[k].copy_(res[k], non_blocking=True)
            eng.sync()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=eng.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    e2e = {"value": world * sub * dur / wall, "unit": "audio-s/s", "h2d_bytes_per_step": sub * n * 8,
           "d2h_bytes_per_step": int(sum(v.numel() for v in res.values()) * 8), "clips_per_step": sub,
           "api": "mm_dev_measure_lufs / true_peak_correlation / spectrum_bars on clips copied from pinned host memory"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak_gbs()
    top = max(ktimes.items(), key=lambda kv: kv[1][0])
    kname, (kms, kcnt) = top
    alg = 4.0 * 2 * n * sub                               # one read of the batch per analyzer kernel
    lufs_host = res["lufs"].cpu().numpy()
    line = {
        "metric": "analyzed audio-seconds per wall-second (integrated LUFS, 4x true peak, sample peak, correlation, 3x64 spectrum bars)",
        "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (64 distinct clips, replicated)",
        "config": {"workload": f"configs[3]: analyzer-only path over {total} synthetic {dur:.0f} s {sr} Hz stereo clips "
                               f"({clips} per GPU in device batches of {sub})", "clips": total, "frames_per_clip": n,
                   "cache": f"batch ({sub * n * 8 / 1e9:.1f} GB) exceeds L2; no flush needed"},
        "roofline": {"bound": "tensor" if False else "hbm", "kernel": kname, "achieved": alg / (kms / kcnt * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": alg / (kms / kcnt * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg,
                     "path": {"algorithmic_bytes_per_stereo_frame": 8.0, "achieved": 8.0 * done * n * args.steps / (ms * 1e-3) / 1e9 / world,
                              "note": "SURVEY 8d counts ONE fused read; this build runs two reading kernels (the meter, and the true-peak FIR "
                                      "with the correlation sums riding on it) -- the FIR is FP32-FMA bound (61 MAC/sample), not HBM bound"},
                     "kernels": {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / args.steps} for k, v in
                                 sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:8]}},
        "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "check": {"lufs_mean": float(np.mean(lufs_host)), "lufs_min": float(np.min(lufs_host)), "lufs_max": float(np.max(lufs_host))},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chain", default="v2", choices=["v1", "v2"])
    ap.add_argument("--tracks", type=int, default=TRACKS)
    ap.add_argument("--sec", type=float, default=DUR)
    ap.add_argument("--e2e-tracks", type=int, default=TRACKS)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="batch", choices=["batch", "mixed", "analyze", "longform"],
                    help="batch = BASELINE configs[1] (default, what the driver times); mixed = configs[2] (48 kHz, mixed presets); "
                         "analyze = configs[3] (analyzer-only over 30 s clips); longform = configs[4], one long file split in time")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "longform":
        run_longform(args)
    elif args.workload == "analyze":
        run_analyze(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
