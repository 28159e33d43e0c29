#!/usr/bin/env python
"""Benchmark of the mastering hot path (BASELINE.json metric: mastered audio-seconds per wall-second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--chain v2|v1]
                    [--workload batch|mixed|analyze|longform] [--extras v1,envelope,mixed,analyze,longform|none]

One step = one pass of the full mastering chain (+ TPDF dither to int16) over a batch of synthetic
tracks that is already resident in HBM.  At N = 1 the workload is BASELINE.json configs[1]:
64 synthetic 3-minute 44.1 kHz stereo tracks.  With N > 1 (torchrun, one rank per GPU) every rank
masters its own 64 tracks (sharded by track, weak scaling) and the per-track loudness/peak stats are
all-gathered over NCCL inside the timed region.  Prints ONE JSON line on rank 0.

The default line also carries, under "extra", bounded runs of the other BASELINE configurations in the same
process (same command the driver runs): the v1 chain and the envelope-compressor mode on configs[1], configs[2] (mixed presets at 48 kHz),
configs[3] (analyzer-only path) and configs[4] (one 2-hour 96 kHz file split in time over the N ranks).

Accuracy gate ("check"): track 0 of the timed batch is the synthetic track the CPU leg masters (cpu_baseline);
the timed run's own output for it is compared with the CPU leg's (samples, LUFS, true peak, int16 under a shared
dither buffer) and the run FAILS (non-zero exit) beyond the north-star tolerances.

CPU legs (cpu_baseline, --impl reference) run the UNMODIFIED reference when it is importable -- in the build container
from /root/reference, on the GPU box from oracle/_ref (the byte-compiled modules oracle/make_ref.py built from that tree;
kind "reference") -- and fall back to the oracle port (oracle/chain.py, kind "port") otherwise.  --impl reference uses all
host cores, one track per process, on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
TOL_SAMPLE, TOL_LU, TOL_DB = 1e-4, 0.01, 0.01          # BASELINE.json north_star tolerances (the gate)

# algorithmic fp32 words moved per channel-sample by each kernel (DESIGN.md "Kernels")
STREAMS = {
    "sweep_fwd_m2_f1_i1": 2, "sweep_bwd_m2_f1_store": 2, "sweep_bwd_m2_f1_combine": 3, "sweep_bwd_m2_f1_exciter": 3,
    "sweep_fwd_m2_f2_i1": 3, "sweep_fwd_m2_f2_i2": 4, "sweep_bwd_m2_f2_store": 4, "sweep_bwd_m2_f2_combine": 4,
    "sweep_bwd_m2_f2_dynamics": 5, "sweep_bwd_m2_f2_dynamics_gen": 5, "sweep_fwd_m2_f4_i1": 5, "sweep_bwd_m2_f4_combine": 6,
    "sweep_fwd_m4_f1_i1": 2, "sweep_bwd_m4_f1_store": 2, "envelope_gain": 2, "deesser_smooth_apply": 4,
    "lufs_kweight_blocks": 1, "row_stats": 1, "finalize_dither_int16": 2.5, "finalize": 2, "peak_after_imager": 1,
    "band_envelope_compress": 5,
}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12       # 148 SMs x 128 FP32 lanes x 2 flop x 1.965 GHz = 74.4 (B200_PROFILING.md)


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
def _ref_kind(compressor="soft_knee"):
    """"reference" when the unmodified reference is importable (the source tree in the build container, else oracle/_ref: the
    byte-compiled copy oracle/make_ref.py built from it, which travels to the GPU box), "port" (oracle/chain.py) otherwise and for
    the envelope-compressor mode, which exists only as a restatement (pedalboard is absent)."""
    if compressor != "soft_knee" or os.environ.get("MM_BENCH_CPU", "") == "port":
        return "port"
    try:
        from oracle import ref_harness
        return "reference" if ref_harness.available() else "port"
    except Exception:
        return "port"


def _cpu_one(args, keep=False):
    """One track through the reference's CPU path: chain + 16-bit TPDF export, timed (synthesis excluded).  kind "reference": the
    reference's own run_mastering_pipeline / MasteringChain.default_chain + apply_output_edge_fade_in (routers/mastering.py:583)
    + export_audio(..., "wav", dither_type="tpdf"); kind "port": oracle/chain.py."""
    t, dur, chain = args[:3]
    sr = args[3] if len(args) > 3 else SR
    style = args[4] if len(args) > 4 else "standard"
    compressor = args[5] if len(args) > 5 else "soft_knee"
    from oracle import chain as oc
    from mm_b200 import synth
    x = synth.numpy_track(t, sr, dur)
    target = oc.STYLE_CONFIGS[style]["lufs"]
    rng = np.random.default_rng(t)
    if _ref_kind(compressor) == "reference":
        from oracle import ref_harness
        ref = ref_harness.load()
        RP = ref.pipeline
        noise = None
        if keep:        # the gate quantises the GPU's samples under the SAME dither values: hand the reference a known buffer
            noise = (rng.random(x.shape) + rng.random(x.shape) - 1.0).astype(np.float32)
            orig = RP._dither_noise_tpdf
            RP._dither_noise_tpdf = lambda shape: noise
        try:
            t0 = time.time()
            if chain == "v1":
                out = RP.run_mastering_pipeline(x.copy(), sr, target_lufs=target, style=style)
            else:
                ch = ref.chain.MasteringChain.default_chain(target_lufs=target, style=style)
                out = RP.apply_output_edge_fade_in(ch.process(x.copy(), sr, target_lufs=target, style=style), sr, fade_ms=6.0)
            wav = RP.export_audio(out, sr, 2, "wav", dither_type="tpdf")
            dt = time.time() - t0
        finally:
            if keep:
                RP._dither_noise_tpdf = orig
        if keep:
            pcm = np.frombuffer(wav[44:], dtype="<i2").reshape(-1, out.shape[1] if out.ndim == 2 else 1)
            return dt, {"x": x, "out": out, "noise": noise, "pcm": pcm, "lufs": RP.measure_lufs(out, sr), "tp": ref.true_peak_dbfs(out, sr),
                        "kind": "reference"}
        return dt
    t0 = time.time()
    out = (oc.run_v1 if chain == "v1" else oc.run_v2)(x, sr, target, style, compressor=compressor)
    noise = (rng.random(out.shape) + rng.random(out.shape) - 1.0).astype(np.float32)
    pcm = oc.quantize_int16(out, noise)
    dt = time.time() - t0
    if keep:
        return dt, {"x": x, "out": out, "noise": noise, "pcm": pcm, "lufs": oc.measure_lufs(out, sr), "tp": oc.true_peak_dbfs(out, sr),
                    "kind": "port"}
    return dt


def cpu_baseline(chain, dur=60.0, procs=1, tracks=None, keep=False):
    """the reference's CPU path (_ref_kind: the unmodified reference when importable, else the oracle port) on `procs` host
    processes, one track each; returns audio-s/s, a description and (keep, procs == 1) its input / outputs for the accuracy gate."""
    import multiprocessing as mp
    tracks = tracks or procs
    work = [(1000 + i, dur, chain) for i in range(tracks)]
    kept = None
    if procs == 1:
        wall = 0.0
        for w in work:
            r = _cpu_one(w, keep=keep)
            if keep:
                wall += r[0]
                kept = r[1]
            else:
                wall += r
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            per = pool.map(_cpu_one, work)
        waves = (tracks + procs - 1) // procs
        wall = max(per) * waves          # chain + export time of the slowest worker (synthesis excluded)
    kind = _ref_kind()
    how = ("the unmodified reference (byte-compiled copy under oracle/_ref; pyloudnorm / soundfile stand-ins of oracle/ref_harness.py)"
           if kind == "reference" else "oracle port (numpy/scipy)")
    desc = f"{tracks} synthetic tracks x {dur:.0f} s, 44.1 kHz stereo, {chain} chain 'standard' + TPDF int16 WAV export, {how}"
    return tracks * dur / wall, desc, kept


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    dur = 30.0
    kind = _ref_kind()
    if kind == "reference":
        from oracle import ref_harness
        ref_harness.load()          # imported (and its numba kernels compiled on first use) before the pool forks
    for _ in range(args.warmup):
        cpu_baseline(args.chain, dur=5.0, procs=procs)
    vals = []
    sample = ""
    for _ in range(args.steps):
        val, sample, _ = cpu_baseline(args.chain, dur=dur, procs=procs)
        vals.append(val)
    value = len(vals) / sum(1.0 / v for v in vals)      # total audio / total time
    wall = args.steps * procs * dur / value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bounded sample of configs[1]: per step {procs} tracks x {dur:.0f} s 44.1 kHz stereo, "
                               f"{args.chain} chain + TPDF int16, one track per host process ("
                               + ("the UNMODIFIED reference: run_mastering_pipeline / MasteringChain + export_audio imported from its "
                                  "byte-compiled modules under oracle/_ref, built by oracle/make_ref.py from /root/reference"
                                  if kind == "reference" else
                                  "the CPU restatement of the reference chain, oracle/chain.py: no built reference under oracle/_ref") + ")",
                   "chain": args.chain},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------
# our arm: runtime, timing harness
# -------------------------------------------------------------------------------------------------------
class Runtime:
    def __init__(self):
        import torch
        import torch.distributed as dist
        from mm_b200.engine import Engine
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (mm_b200 has no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            with _StdoutToStderr():
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                t = torch.zeros(1, device=torch.device("cuda", self.local))
                dist.all_reduce(t)                 # communicator creation (and its banner) happens here
                torch.cuda.synchronize()
        self.eng = Engine(self.local)

    def barrier(self):
        self.eng.sync()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.eng.tdev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def free(self):
        """Between workloads: the chain's scratch and torch's cached blocks go back to the driver."""
        import gc
        self.eng.sync()
        self.eng.release_workspace()
        gc.collect()
        self.torch.cuda.empty_cache()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(rt, step, steps, warmup, clocks=True, separate_kernel_pass=False):
    """W untimed steps, then exactly K steps bracketed by barrier + synchronize, CUDA events on the context's stream (the stream
    the kernels are launched on), per-kernel events for the roofline, MAX over ranks.  separate_kernel_pass: the K timed steps run
    WITHOUT the per-kernel events (two event records per launch are ~1 % of a 9 ms step of 28 launches) and the per-kernel table
    comes from K further steps."""
    torch, eng = rt.torch, rt.eng
    for i in range(warmup):
        step(i)
    rt.barrier()
    if separate_kernel_pass:
        sampler = ClockSampler(rt.local) if (clocks and rt.rank == 0) else None
        if sampler:
            sampler.start()
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        for i in range(steps):
            step(warmup + i)
        e1.record(eng.stream)
        rt.barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0
        clk = sampler.stop() if sampler else None
        eng.timing(True)
        for i in range(steps):
            step(warmup + steps + i)
        rt.barrier()
        ktimes = eng.kernel_times()
        eng.timing(False)
        return {"ms": rt.max_over_ranks(ms), "launches": int(launches), "ktimes": ktimes, "clocks": clk}
    sampler = ClockSampler(rt.local) if (clocks and rt.rank == 0) else None
    if sampler:
        sampler.start()
    eng.timing(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for i in range(steps):
        step(warmup + i)
    e1.record(eng.stream)
    rt.barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    ktimes = eng.kernel_times()
    eng.timing(False)
    clk = sampler.stop() if sampler else None
    return {"ms": rt.max_over_ranks(ms), "launches": int(launches), "ktimes": ktimes, "clocks": clk}


def kernel_table(ktimes, steps, default_samples, top=12):
    """Per-kernel average launch time and algorithmic GB/s.  Bytes per launch = words x 4 x the channel-samples the launch
    actually visited (reported by the launcher: row lists / track runs of a mixed batch), not the whole batch."""
    out = {}
    for k, (ms, cnt, samples) in sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:top]:
        smp = samples / cnt if samples > 0 else default_samples
        out[k] = {"ms_per_launch": ms / cnt, "launches_per_step": cnt / steps, "samples_per_launch": smp,
                  "gbs": STREAMS.get(k, 0) * 4.0 * smp / (ms / cnt * 1e-3) / 1e9}
    return out


def dominant_roofline(ktimes, steps, default_samples, peak, peak_src):
    kname, (kms, kcnt, ksamples) = max(ktimes.items(), key=lambda kv: kv[1][0])
    smp = ksamples / kcnt if ksamples > 0 else default_samples
    alg_bytes = STREAMS.get(kname, 2) * 4.0 * smp
    achieved = alg_bytes / (kms / kcnt * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(kname)
        except Exception:
            traffic = None
    ksum = sum(v[0] for v in ktimes.values())
    return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
            "peak_note": "peak is the measured COPY bandwidth (one read + one write stream); a read-only kernel (row_stats) can "
                         "exceed it -- HBM3e reads alone run at up to ~7.7 TB/s nominal",
            "kernel_share_of_step": kms / ksum, "kernels": kernel_table(ktimes, steps, default_samples)}


def gate_record(max_abs, d_lufs, d_tp, int16_same_input, int16_chain_maxdiff, what):
    ok = (max_abs <= TOL_SAMPLE and d_lufs <= TOL_LU and d_tp <= TOL_DB and int16_same_input == 0)
    return {"what": what, "max_abs": max_abs, "dLUFS": d_lufs, "dTP": d_tp, "int16_mismatch": int(int16_same_input),
            "int16_chain_maxdiff_lsb": int(int16_chain_maxdiff), "tol": {"max_abs": TOL_SAMPLE, "dLUFS": TOL_LU, "dTP": TOL_DB, "int16_mismatch": 0},
            "pass": bool(ok)}


def side_gate(rt, chain, style, sr, dur=10.0, track=1001, compressor="soft_knee"):
    """Short-track gate for the bounded extras: one synthetic track through P.master_batch (same library, same precision policy)
    against the oracle."""
    from mm_b200 import pipeline as P
    from oracle import chain as oc
    _, kept = _cpu_one((track, dur, chain, sr, style, compressor), keep=True)
    target = P.STYLE_CONFIGS[style]["lufs"]
    res = P.master_batch([kept["x"]], sr, [style], [target], chain=chain, want_int16=True, noise=kept["noise"][None], measure=True, eng=rt.eng,
                         compressor=compressor)
    out = res["audio"][0]
    max_abs = float(np.max(np.abs(out.astype(np.float64) - kept["out"])))
    q = rt.eng.quantize_int16(rt.eng.upload([kept["out"]], sr), noise=kept["noise"][None])[0]
    return gate_record(max_abs, abs(res["stats"][0]["lufs_out"] - kept["lufs"]), abs(P.true_peak_dbfs(out, sr) - kept["tp"]),
                       int(np.count_nonzero(q != kept["pcm"])), int(np.max(np.abs(res["int16"][0].astype(np.int32) - kept["pcm"].astype(np.int32)))),
                       f"{dur:.0f} s synthetic track {track}, {sr} Hz, {chain}/{style}" + (" (envelope compressor)" if compressor != "soft_knee" else "") +
                       f" through master_batch vs the {kept['kind']}")


# -------------------------------------------------------------------------------------------------------
# configs[1] / configs[2]: a batch of tracks through the full chain
# -------------------------------------------------------------------------------------------------------
def bench_batch(rt, args, *, chain, mixed, steps, warmup, main, envelope=False):
    torch, eng, world, rank = rt.torch, rt.eng, rt.world, rt.rank
    from mm_b200 import _lib, pipeline as P, shard, synth
    from mm_b200.engine import style_struct, TrackStats
    tracks, dur, sr = args.tracks, args.sec, SR
    if mixed:      # BASELINE configs[2]: 48 kHz, genre presets cycling over the GLOBAL track index, 128 tracks per GPU
        sr = 48000
        tracks = args.tracks if args.tracks != TRACKS else 128
    n = int(round(sr * dur))
    chain_id = _lib.CHAIN_V1 if chain == "v1" else _lib.CHAIN_V2

    # synthetic batch, generated on the device (SURVEY 8d generator), resident in HBM before timing
    src = eng.empty(tracks, 2, n, sr)
    ids = shard.shard_tracks(world * tracks, world, rank)      # track t -> rank t % world
    if mixed and world > 1:
        # presets cycle with period 8: plain t % world would hand a whole rank one preset at world = 8 (rank 1 all "edm", 676 B per
        # frame; rank 0 all "standard", 436) and the step would wait for the heaviest rank.  Rotated round-robin (SURVEY 8e:
        # "round-robin or size-balanced by track"): the assignment shifts by one rank per period of the preset cycle, so every rank
        # holds every preset in equal numbers whenever the world size divides the period (2, 4, 8 GPUs)
        period = len(P.STYLE_CONFIGS)
        rot = period if period % world == 0 else world
        ids = [t for t in range(world * tracks) if (t + t // rot) % world == rank]
    with torch.cuda.stream(eng.stream):
        src.t.zero_()
        synth.torch_batch(ids, sr, dur, eng.tdev, out=src.t, row_stride=src.stride, lead=_lib.MM_LEAD)
    # accuracy gate: track 0 of the timed batch is the track the CPU oracle masters (cpu_baseline leg, N = 1, rank 0)
    kept = None
    cb = None
    do_gate = main and world == 1 and not args.no_cpu
    if do_gate:
        v, sample, kept = cpu_baseline(chain, dur=dur, procs=1, tracks=1, keep=True)
        cb = {"value": v, "unit": UNIT, "cores": 1, "kind": kept["kind"], "sample": sample}
        with torch.cuda.stream(eng.stream):
            xt = torch.from_numpy(np.ascontiguousarray(kept["x"].T)).to(eng.tdev)
            src.live()[0:2].copy_(xt)
            del xt
    out = eng.like(src)
    names = list(P.STYLE_CONFIGS)
    style_names = [names[t % len(names)] if mixed else "standard" for t in ids]
    styles = [style_struct(P.STYLE_CONFIGS[s], P.STYLE_CONFIGS[s]["lufs"] if mixed else -14.0) for s in style_names]
    # algorithmic bytes per stereo frame of each track: 8 (W_base + 5 n_style_bands + 5 [exciter fires])
    wbase = (64.5 if chain == "v1" else 54.5) + (4.0 if envelope else 0.0)      # envelope-compressor mode: +4 words

    def _alg_bytes(sn):
        cfg = P.STYLE_CONFIGS[sn]
        nb = sum(1 for k in ("sub", "bass", "mids", "presence", "air") if abs(cfg.get(k, 0.0)) >= 0.05)
        exc = (cfg.get("exciter_db", 0.0) > 0.05) if chain == "v1" else (abs(cfg.get("exciter_db", 0.0)) >= 0.05)
        return 8.0 * (wbase + 5 * nb + 5 * (1 if exc else 0))
    alg_per_frame = float(np.mean([_alg_bytes(sn) for sn in style_names]))
    arr = (_lib.Style * tracks)(*styles)
    with torch.cuda.stream(eng.stream):
        pcm = torch.empty((tracks, n, 2), dtype=torch.int16, device=eng.tdev)
        stats = torch.empty((tracks, shard.STATS_DOUBLES), dtype=torch.float64, device=eng.tdev)
    g = src.geom
    flags = _lib.FLAG_MEASURE_OUT | (_lib.FLAG_ENVELOPE_COMPRESSOR if envelope else 0)

    def step(i):
        _lib.check(eng.lib.mm_dev_master(eng.ctx, C.byref(g), chain_id, arr, src.ptr, out.ptr, C.c_void_p(pcm.data_ptr()), None,
                                         1234 + i, C.c_void_p(stats.data_ptr()), flags))
        if world > 1:      # the only exchange of the sharded path: per-track stats records over NCCL
            with torch.cuda.stream(eng.stream):
                shard.gather_track_stats(stats, world * tracks, world, rank)

    # the K timed steps run without per-kernel events and with the library's lanes (the batch split over concurrent streams); the
    # per-kernel table comes from K further steps with events on, which the library runs on ONE stream so a kernel is timed alone
    tm = timed_steps(rt, step, steps, warmup, clocks=main, separate_kernel_pass=True)
    ms = tm["ms"]
    value = world * tracks * dur * steps / (ms * 1e-3)

    # sanity of the timed work: every track was mastered to its target within the gate
    eng.sync()
    recs = shard.stats_to_records(stats)
    lufs_out = np.array([r["lufs_out"] for r in recs])
    check = {"lufs_out_mean": float(np.mean(lufs_out)), "lufs_out_min": float(np.min(lufs_out)),
             "lufs_out_max": float(np.max(lufs_out)), "nonfinite": float(sum(r["nonfinite"] for r in recs))}
    if kept is not None:
        # the timed run's OWN output for track 0 against the oracle's output for the same samples
        with torch.cuda.stream(eng.stream):
            got = out.live()[0:2].t().contiguous().cpu().numpy()
            got_pcm = pcm[0].cpu().numpy()
        max_abs = float(np.max(np.abs(got.astype(np.float64) - kept["out"])))
        tp = P.true_peak_dbfs(got, sr)
        q = eng.quantize_int16(eng.upload([kept["out"]], sr), noise=kept["noise"][None])[0]     # the quantiser alone: bit-exact
        own = eng.quantize_int16(eng.upload([got], sr), noise=kept["noise"][None])[0]            # chain output, shared dither buffer
        check["gate"] = gate_record(max_abs, abs(recs[0]["lufs_out"] - kept["lufs"]), abs(tp - kept["tp"]),
                                    int(np.count_nonzero(q != kept["pcm"])), int(np.max(np.abs(own.astype(np.int32) - kept["pcm"].astype(np.int32)))),
                                    f"track 0 of the timed batch ({dur:.0f} s, synthetic track 1000) vs the {kept['kind']}'s output for the same samples")
        check["gate"]["philox_int16_vs_oracle_float_lsb"] = float(np.max(np.abs(got_pcm.astype(np.float64) - kept["out"].astype(np.float64) * 32767.0)))
        del kept
    elif not args.no_cpu and rank == 0:
        check["gate"] = side_gate(rt, chain, "edm" if mixed else "standard", sr, compressor="envelope" if envelope else "soft_knee")

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory, copies inside the timing) ----
    e2e = None
    if main and args.e2e_tracks > 0:      # --e2e-tracks 0: device-resident numbers only (kernel experiments)
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
        torch.cuda.synchronize()
        earr = (_lib.Style * e_tracks)(*styles[:e_tracks])
        reps, e_warm = max(3, min(steps, 10)), 2

        def wall_of(fn):
            for i in range(e_warm):
                fn(i)
            rt.barrier()
            t0 = time.perf_counter()
            for i in range(reps):
                fn(e_warm + i)
            torch.cuda.synchronize()
            return rt.max_over_ranks(time.perf_counter() - t0)

        # what the links give: the same bytes per step as plain pinned copies, both directions at once, every rank at once
        s_in, s_out = torch.cuda.Stream(device=eng.tdev), torch.cuda.Stream(device=eng.tdev)
        with torch.cuda.stream(eng.stream):
            dpcm = torch.empty((e_tracks, n, 2), dtype=torch.int16, device=eng.tdev)
        eng.sync()

        def copy_step(i):
            with torch.cuda.stream(s_in):
                il.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s_out):
                hpcm.copy_(dpcm, non_blocking=True)
        wall_c = wall_of(copy_step)
        ceiling = world * e_tracks * dur * reps / wall_c

        def e2e_step(i):
            _lib.check(eng.lib.mm_master_host(eng.ctx, chain_id, e_tracks, n, 2, sr, earr, C.c_void_p(hin.data_ptr()), None,
                                              C.c_void_p(hpcm.data_ptr()), None, 99 + i, hstats, flags))
        wall = wall_of(e2e_step)
        e2e = {"value": world * e_tracks * dur * reps / wall, "unit": UNIT, "h2d_bytes_per_step": frames * 4,
               "d2h_bytes_per_step": frames * 2 + e_tracks * C.sizeof(TrackStats), "tracks_per_step": e_tracks, "reps": reps,
               "api": "mm_master_host (C ABI, pinned host buffers in/out)",
               "copy_ceiling": {"value": ceiling, "unit": UNIT, "h2d_gbs_per_rank": frames * 4 * reps / wall_c / 1e9,
                                "d2h_gbs_per_rank": frames * 2 * reps / wall_c / 1e9,
                                "what": "the step's bytes as plain pinned cudaMemcpyAsync, H2D and D2H concurrently, all ranks at once"},
               "frac_of_copy_ceiling": (world * e_tracks * dur * reps / wall) / ceiling}
        # job-level variant: the upload's PCM_16 frames cross PCIe as they are
        hin16 = torch.empty((e_tracks, n, 2), dtype=torch.int16, pin_memory=True)
        hin16.copy_((hin * 32767.0).round().to(torch.int16))
        del hin, il

        def copy16_step(i):
            with torch.cuda.stream(s_in):
                dpcm.copy_(hin16, non_blocking=True)
            with torch.cuda.stream(s_out):
                hpcm.copy_(dpcm, non_blocking=True)
        # (reads and writes of dpcm race; only the transfer time matters here)
        wall_c16 = wall_of(copy16_step)

        def e2e16_step(i):
            _lib.check(eng.lib.mm_master_host_pcm16(eng.ctx, chain_id, e_tracks, n, 2, sr, earr, C.c_void_p(hin16.data_ptr()), None,
                                                    C.c_void_p(hpcm.data_ptr()), 199 + i, hstats, flags))
        wall16 = wall_of(e2e16_step)
        e2e["pcm16_in"] = {"value": world * e_tracks * dur * reps / wall16, "unit": UNIT, "h2d_bytes_per_step": frames * 2,
                           "d2h_bytes_per_step": frames * 2 + e_tracks * C.sizeof(TrackStats),
                           "api": "mm_master_host_pcm16 (PCM_16 frames in and out, widened on the device: the job path of a WAV upload)",
                           "copy_ceiling": world * e_tracks * dur * reps / wall_c16,
                           "frac_of_copy_ceiling": (world * e_tracks * dur * reps / wall16) / (world * e_tracks * dur * reps / wall_c16)}
        # ragged variant (/api/v2/batch, routers/mastering.py:855-1037): the same uploads, no two of the same length, as ONE call
        rjobs = (_lib.HostJob * e_tracks)()
        r_audio_s = 0.0
        for t in range(e_tracks):
            nt = n - 4410 * t                                    # 180 s, 179.9 s, ... : every upload is a chunk of its own
            rjobs[t].n, rjobs[t].channels, rjobs[t].sr = nt, 2, sr
            rjobs[t].pcm16_in, rjobs[t].pcm16_out = hin16[t].data_ptr(), hpcm[t].data_ptr()
            rjobs[t].style, rjobs[t].dither_id = earr[t], t
            r_audio_s += nt / sr

        def ragged_step(i):
            _lib.check(eng.lib.mm_master_host_jobs(eng.ctx, chain_id, e_tracks, rjobs, 299 + i, flags))
        wall_r = wall_of(ragged_step)
        e2e["ragged_jobs"] = {"value": world * r_audio_s * reps / wall_r, "unit": UNIT, "uploads_per_step": e_tracks,
                              "frames": f"{n - 4410 * (e_tracks - 1)} .. {n} (all different)",
                              "api": "mm_master_host_jobs (one mm_host_job per upload: own length, own pinned PCM_16 buffers)"}
        del hin16, hpcm, dpcm, rjobs
        # single-job latency through the Python drop-in (pageable numpy in / out, one 180 s track), rank 0
        if rank == 0 and not args.no_cpu:
            x1 = synth.numpy_track(1000, sr, dur)
            P.run_mastering_pipeline(x1, sr) if chain == "v1" else P.master_batch([x1], sr, ["standard"], chain="v2", eng=eng)
            t0 = time.perf_counter()
            for _ in range(3):
                if chain == "v1":
                    P.run_mastering_pipeline(x1, sr)
                else:
                    P.master_batch([x1], sr, ["standard"], chain="v2", eng=eng)
            lat = (time.perf_counter() - t0) / 3
            e2e["single_job"] = {"ms": lat * 1e3, "audio_s_per_s": dur / lat,
                                 "api": ("run_mastering_pipeline" if chain == "v1" else "master_batch([x])") + " on one pageable numpy (n, 2) float32 track, result back as numpy"}

    peak, peak_src = measured_peak_gbs()
    rows_n = tracks * 2 * n
    roofline = dominant_roofline(tm["ktimes"], steps, rows_n, peak, peak_src)
    if world > 1:      # bytes of ALL ranks over the slowest rank's time (ranks may hold different presets)
        tt = torch.tensor([alg_per_frame], dtype=torch.float64, device=eng.tdev)
        rt.dist.all_reduce(tt)
        alg_per_frame = float(tt.item()) / world
    chain_gbs = alg_per_frame * tracks * n * world * steps / (ms * 1e-3) / 1e9
    roofline["chain"] = {"algorithmic_bytes_per_stereo_frame": alg_per_frame, "achieved": chain_gbs / world, "frac": chain_gbs / world / peak}
    roofline["note"] = ("per-kernel times: a separate pass of K steps with CUDA events around every launch, run by the library on ONE stream "
                        "(a kernel timed alone, over the whole batch); the K timed steps of `value` run without those events and with the "
                        "batch split over the lanes, whose kernels overlap -- so the step is shorter than the sum of the per-kernel times "
                        f"({sum(v[0] for v in tm['ktimes'].values()) / steps:.2f} ms)")
    cfg = {"workload": (f"configs[2]: {tracks} synthetic {dur:.0f} s {sr} Hz stereo tracks per GPU, genre presets cycling over the "
                        f"global track index (STYLE_CONFIGS order) at their own LUFS targets, {chain} chain + TPDF int16 + after-LUFS"
                        if mixed else
                        f"configs[1]: {tracks} synthetic {dur:.0f} s {sr} Hz stereo tracks per GPU, {chain} default chain "
                        f"(style standard, -14 LUFS) + TPDF dither to int16 + after-LUFS"),
           "chain": chain, "compressor": "envelope (pedalboard-style, parity unpinned)" if envelope else "soft_knee (pinned)",
           "tracks_per_gpu": tracks, "frames_per_track": n,
           "lanes": "MM_LANES=" + os.environ.get("MM_LANES", "auto") + ": mm_dev_master spreads the batch over concurrent streams (auto: two "
                    "halves); every timed step covers the whole batch",
           "cache": f"inputs ({tracks * n * 8 / 1e9:.2f} GB per GPU) exceed L2; no flush needed",
           "precision_policy": "MM_PASS2=" + os.environ.get("MM_PASS2", "auto") + ": float32 streams; float64 chunk scan everywhere; in-chunk "
                               "recurrences float64 (full-path low cut-offs) or float32 FFMA2 on balanced realizations (DESIGN.md)"}
    res = {"value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "weak", "config": cfg,
           "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": tm["launches"], "clocks": tm["clocks"], "check": check}
    del src, out, pcm, stats
    return res


# -------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: one 2-hour 96 kHz stereo file, time-split over the ranks (strong scaling)
# -------------------------------------------------------------------------------------------------------
def bench_longform(rt, args, *, chain, steps, warmup, main):
    torch, eng, world, rank = rt.torch, rt.eng, rt.world, rt.rank
    from mm_b200 import _lib, longform, pipeline as P, synth
    from mm_b200.engine import style_struct, TrackStats
    from mm_b200.shard import stats_to_records
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
    xch = longform.make_exchange(eng, world, rank) if world > 1 else None
    xch = xch or longform.Exchange(note="single rank: none")
    sl = xch.slice_struct(n, plan)
    style = (_lib.Style * 1)(style_struct(P.STYLE_CONFIGS["standard"], -14.0))
    chain_id = _lib.CHAIN_V1 if chain == "v1" else _lib.CHAIN_V2
    g = src.geom

    def step(i):
        _lib.check(eng.lib.mm_dev_master_slice(eng.ctx, C.byref(g), chain_id, style, src.ptr, out.ptr, C.c_void_p(pcm.data_ptr()), None,
                                               1234 + i, C.c_void_p(st.data_ptr()), _lib.FLAG_MEASURE_OUT, C.byref(sl)))

    tm = timed_steps(rt, step, steps, warmup, clocks=main, separate_kernel_pass=True)
    ms = tm["ms"]
    value = dur * steps / (ms * 1e-3)
    rec = stats_to_records(np.frombuffer(st.cpu().numpy().tobytes(), dtype=np.float64).reshape(1, -1))[0]
    e2e = None
    if main:
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
            rt.barrier()
            t0 = time.perf_counter()
            reps = max(1, min(steps, 3))
            for i in range(reps):
                e2e_step(1 + i)
            torch.cuda.synchronize()
            wall = rt.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": dur * reps / wall, "unit": UNIT, "h2d_bytes_per_step": ns * 2 * 4, "d2h_bytes_per_step": own * 2 * 2,
               "api": "mm_dev_master_slice on a slice copied from / to pinned host memory (per rank)"}
        del hin, hpcm, il
    peak, peak_src = measured_peak_gbs()
    bpf = 8.0 * (64.5 if chain == "v1" else 54.5)
    roofline = dominant_roofline(tm["ktimes"], steps, 2 * ns, peak, peak_src)
    chain_gbs = bpf * ns * steps / (ms * 1e-3) / 1e9
    roofline["chain"] = {"algorithmic_bytes_per_stereo_frame": bpf, "achieved": chain_gbs, "frac": chain_gbs / peak,
                         "note": "rank 0: its slice including margins"}
    cb_line = None
    if main and world == 1 and not args.no_cpu:
        from oracle import chain as oc
        x = synth.numpy_track(0, sr, 20.0)
        t0 = time.time()
        o = (oc.run_v1 if chain == "v1" else oc.run_v2)(x, sr, -14.0, "standard")
        rng = np.random.default_rng(0)
        oc.quantize_int16(o, (rng.random(o.shape) + rng.random(o.shape) - 1.0).astype(np.float32))
        cb_line = {"value": 20.0 / (time.time() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "20 s of 96 kHz stereo through the oracle (numpy/scipy), v2 chain + TPDF int16"}
    res = {"value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "strong",
           "config": {"workload": f"configs[4]: one {dur:.0f} s {sr} Hz stereo file, {chain} default chain + TPDF int16 + after-LUFS, "
                                  f"split in time over {world} GPU(s) ({longform.slice_margin(sr)} margin frames per cut side)",
                      "chain": chain, "frames": n, "slice_frames_rank0": ns, "cache": "slice (GBs) exceeds L2; no flush needed",
                      "exchange": xch.describe()},
           "roofline": roofline, "cpu_baseline": cb_line, "e2e": e2e, "gpu_launches": tm["launches"], "clocks": tm["clocks"],
           "check": {"lufs_out": rec["lufs_out"], "gain_db": rec["gain_db"], "peak_out": rec["peak_out"], "nonfinite": rec["nonfinite"]}}
    xch.close()
    del src, out, pcm, st
    return res


# -------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: analyzer-only path over 30 s clips (LUFS + gating, 4x true peak, correlation, spectrum bars)
# -------------------------------------------------------------------------------------------------------
def bench_analyze(rt, args, *, steps, warmup, main):
    torch, eng, world, rank = rt.torch, rt.eng, rt.world, rt.rank
    from mm_b200 import _lib, shard, synth
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
        gate = None
        if rank == 0 and not args.no_cpu:         # clip 0 := a host-synthesised clip the oracle measures too
            x0 = synth.numpy_track(1002, sr, dur)
            src.live()[0:2].copy_(torch.from_numpy(np.ascontiguousarray(x0.T)).to(eng.tdev))
        for k in range(distinct, sub, distinct):   # fill the batch with copies of the 64 distinct clips
            m = min(distinct, sub - k)
            src.t[2 * k:2 * (k + m)].copy_(src.t[:2 * m])
        res = {k: torch.empty(sub * w, dtype=torch.float64, device=eng.tdev) for k, w in
               (("lufs", 1), ("tp", 1), ("corr", 1), ("peak", 1), ("bars0", 64), ("bars1", 64), ("bars2", 64))}
    g = src.geom
    nbatch = (clips + sub - 1) // sub
    ptr = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731

    def analyze():
        _lib.check(eng.lib.mm_dev_measure_lufs(eng.ctx, C.byref(g), src.ptr, ptr(res["lufs"])))
        _lib.check(eng.lib.mm_dev_true_peak_correlation(eng.ctx, C.byref(g), src.ptr, ptr(res["tp"]), ptr(res["corr"]), ptr(res["peak"])))
        for v in range(3):
            _lib.check(eng.lib.mm_dev_spectrum_bars(eng.ctx, C.byref(g), src.ptr, v, ptr(res[f"bars{v}"])))

    def step(i):
        for _ in range(nbatch):
            analyze()

    tm = timed_steps(rt, step, steps, warmup, clocks=main)
    ms = tm["ms"]
    done = nbatch * sub * world
    value = done * dur * steps / (ms * 1e-3)
    lufs_host = res["lufs"].cpu().numpy()
    check = {"lufs_mean": float(np.mean(lufs_host)), "lufs_min": float(np.min(lufs_host)), "lufs_max": float(np.max(lufs_host))}
    if rank == 0 and not args.no_cpu:
        from oracle import chain as oc
        d_l = abs(float(lufs_host[0]) - oc.measure_lufs(x0, sr))
        d_tp = abs(float(res["tp"].cpu()[0]) - oc.true_peak_dbfs(x0, sr))
        corr_ref = oc.measure_stereo_correlation(x0)
        bars_ref = np.array(oc.compute_spectrum_bars(x0, sr))
        d_bars = float(np.max(np.abs(res["bars0"].cpu().numpy()[:64] - bars_ref)))
        check["gate"] = {"what": "clip 0 of the timed batch (30 s, synthetic track 1002) vs the oracle", "dLUFS": d_l, "dTP": d_tp,
                         "dcorr": abs(float(res["corr"].cpu()[0]) - corr_ref), "dbars_db": d_bars,
                         "pass": bool(d_l <= TOL_LU and d_tp <= TOL_DB and d_bars <= 0.011)}
    e2e = None
    if main:
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
                analyze()
                for k in res:
                    hres[k].copy_(res[k], non_blocking=True)
                eng.sync()

            e2e_step()
            rt.barrier()
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            wall = rt.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * sub * dur / wall, "unit": "audio-s/s", "h2d_bytes_per_step": sub * n * 8,
               "d2h_bytes_per_step": int(sum(v.numel() for v in res.values()) * 8), "clips_per_step": sub,
               "api": "mm_dev_measure_lufs / true_peak_correlation / spectrum_bars on clips copied from pinned host memory"}
        del hin, il
    peak, peak_src = measured_peak_gbs()
    ktimes = tm["ktimes"]
    kname, (kms, kcnt, _) = max(ktimes.items(), key=lambda kv: kv[1][0])
    samples = 2.0 * n * sub
    alg = 4.0 * samples                                  # one read of the batch per analyzer kernel
    roofline = {"kernel": kname, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                "hbm": {"achieved": alg / (kms / kcnt * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (kms / kcnt * 1e-3) / 1e9 / peak}}
    if "true_peak" in kname:
        # 4x polyphase 81-tap FIR: 61 multiply-adds per input sample (branch 0 is a pure delay), 2 flop each -- FP32 pipe bound
        tf = 61 * 2.0 * samples / (kms / kcnt * 1e-3) / 1e12
        roofline.update({"bound": "fp32", "achieved": tf, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP32_PEAK_TFLOPS,
                         "traffic": None, "note": "FP32 FMA roofline (148 SMs x 128 lanes x 2 flop x 1.965 GHz nominal boost; no measured FP32 "
                                                  "peak in MEASURED_PEAKS.json); the kernel's HBM figure sits beside it"})
    else:
        roofline.update({"bound": "hbm", "traffic": None, **roofline["hbm"]})
    roofline["path"] = {"algorithmic_bytes_per_stereo_frame": 8.0, "achieved": 8.0 * done * n * steps / (ms * 1e-3) / 1e9 / world,
                        "note": "SURVEY 8d counts ONE fused read; this build runs two reading kernels (the meter, and the true-peak FIR "
                                "with the correlation sums riding on it)"}
    roofline["kernels"] = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / steps,
                               "gbs": alg / (v[0] / v[1] * 1e-3) / 1e9 if ("lufs_kweight" in k or "true_peak" in k) else None}
                           for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:8]}
    out = {"metric": "analyzed audio-seconds per wall-second (integrated LUFS, 4x true peak, sample peak, correlation, 3x64 spectrum bars)",
           "value": value, "unit": "audio-s/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "scaling": "strong",
           "config": {"workload": f"configs[3]: analyzer-only path over {total} synthetic {dur:.0f} s {sr} Hz stereo clips "
                                  f"({clips} per GPU in device batches of {sub}; 64 distinct clips, replicated)", "clips": total, "frames_per_clip": n,
                      "cache": f"batch ({sub * n * 8 / 1e9:.1f} GB) exceeds L2; no flush needed"},
           "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": tm["launches"], "clocks": tm["clocks"], "check": check}
    del src, res
    return out


def compact(r):
    """An extra's record inside the main line: value, time, chain fraction, dominant kernel, gate."""
    rf = r.get("roofline") or {}
    out = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"], "steps": r["steps"], "warmup": r["warmup"],
           "scaling": r["scaling"], "workload": r["config"]["workload"], "gpu_launches_per_step": r["gpu_launches"] / max(r["steps"], 1),
           "roofline": {k: rf.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac") if k in rf},
           "check": r.get("check")}
    for k in ("chain", "path"):
        if k in rf:
            out["roofline"][k] = rf[k]
    if "kernels" in rf:
        out["roofline"]["kernels"] = {k: {kk: vv for kk, vv in v.items() if kk in ("ms_per_launch", "launches_per_step", "gbs")}
                                      for k, v in list(rf["kernels"].items())[:8]}
    return out


def run_ours(args):
    rt = Runtime()
    steps, warmup = args.steps, args.warmup
    xs, xw = max(10, min(steps, 10)), 3                          # the extras: >= 10 timed steps, 3 warm-ups
    if args.workload == "longform":
        main = bench_longform(rt, args, chain=args.chain, steps=steps, warmup=warmup, main=True)
    elif args.workload == "analyze":
        main = bench_analyze(rt, args, steps=steps, warmup=warmup, main=True)
    else:
        main = bench_batch(rt, args, chain=args.chain, mixed=args.workload == "mixed", steps=steps, warmup=warmup, main=True)
    extra = {}
    want = [] if (args.extras == "none" or args.workload != "batch") else args.extras.split(",")
    for name in want:
        rt.free()
        try:
            if name == "v1":
                extra["v1_chain"] = compact(bench_batch(rt, args, chain="v1", mixed=False, steps=xs, warmup=xw, main=False))
            elif name == "envelope":
                extra["v2_envelope"] = compact(bench_batch(rt, args, chain=args.chain, mixed=False, steps=xs, warmup=xw, main=False, envelope=True))
            elif name == "mixed":
                extra["mixed"] = compact(bench_batch(rt, args, chain=args.chain, mixed=True, steps=xs, warmup=xw, main=False))
            elif name == "analyze":
                extra["analyze"] = compact(bench_analyze(rt, args, steps=xs, warmup=xw, main=False))
            elif name == "longform":
                extra["longform"] = compact(bench_longform(rt, args, chain=args.chain, steps=xs, warmup=xw, main=False))
        except Exception as e:                                   # an extra must not take the headline line down with it
            extra[name] = {"error": repr(e)}
            if rt.world > 1:
                raise
    if rt.rank != 0:
        rt.close()
        return 0
    line = {
        "metric": main.get("metric", METRIC), "value": main["value"], "unit": main["unit"], "n_gpus": rt.world, "steps": steps, "warmup": warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": main["scaling"], "vs_baseline": None,
        "dtype": "f64/f32", "data": "synthetic", "config": main["config"], "roofline": main["roofline"], "cpu_baseline": main["cpu_baseline"],
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "check": main["check"],
    }
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    rt.close()
    gates = [main["check"].get("gate")] + [(v.get("check") or {}).get("gate") for v in extra.values() if isinstance(v, dict)]
    failed = [g for g in gates if g and not g.get("pass", True)]
    if failed:
        print("bench.py: ACCURACY GATE FAILED: " + json.dumps(failed), file=sys.stderr, flush=True)
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chain", default="v2", choices=["v1", "v2"])
    ap.add_argument("--tracks", type=int, default=TRACKS)
    ap.add_argument("--sec", type=float, default=DUR)
    ap.add_argument("--e2e-tracks", type=int, default=TRACKS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU oracle legs (cpu_baseline and the accuracy gates)")
    ap.add_argument("--workload", default="batch", choices=["batch", "mixed", "analyze", "longform"],
                    help="batch = BASELINE configs[1] (default, what the driver times); mixed = configs[2] (48 kHz, mixed presets); "
                         "analyze = configs[3] (analyzer-only over 30 s clips); longform = configs[4], one long file split in time")
    ap.add_argument("--extras", default="v1,envelope,mixed,analyze,longform",
                    help="bounded runs of the other configurations appended to the default line under 'extra' ('none' to skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return 0
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
