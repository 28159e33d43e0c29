"""Time-split mastering of one long file (BASELINE config 5): G virtual ranks on ONE GPU (threads, one engine
each, an in-process all-reduce standing in for NCCL) against the whole-file run and the CPU oracle."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _LocalGroup:
    """All-reduce between threads that share a device."""

    def __init__(self, world):
        self.world, self.bufs, self.bar = world, [None] * world, threading.Barrier(world)

    def reducer(self, rank, eng):
        import torch

        def red(t, op):
            eng.sync()                                    # t is produced on this rank's stream
            with torch.cuda.stream(eng.stream):
                self.bufs[rank] = t.clone()
            eng.sync()
            self.bar.wait()
            with torch.cuda.stream(eng.stream):
                s = torch.stack(self.bufs)
                t.copy_(s.sum(0) if op == 0 else (s.min(0).values if op == 1 else s.max(0).values))
            eng.sync()
            self.bar.wait()
        return red


def _run_split(x, sr, style_name, chain, world, want_int16=False, seed=5, margin=None):
    import torch
    from mm_b200 import longform, pipeline as P
    from mm_b200.engine import Engine
    n = x.shape[0]
    margin = longform.slice_margin(sr) if margin is None else margin
    plans = longform.plan_slices(n, world, margin)
    grp = _LocalGroup(world)
    out, errs = [None] * world, []

    def work(rank):
        try:
            eng = Engine(0)
            tdt = {0: "<f8", 1: "<i8", 2: "<f4"}

            def wrap(ptr, count, dtype):
                return torch.as_tensor(longform._DevView(ptr, count, tdt[dtype]), device=eng.tdev)

            cb = longform.make_allreduce(grp.reducer(rank, eng), wrap)
            p = plans[rank]
            out[rank] = longform.master_slice(eng, x[p["start"]:p["stop"]], sr, p, n, P.STYLE_CONFIGS[style_name],
                                              P.STYLE_CONFIGS[style_name]["lufs"], chain, allreduce=cb, want_int16=want_int16,
                                              seed=seed, measure=True)
            out[rank].pop("device_out")
            eng.close()
        except Exception as e:       # pragma: no cover
            errs.append((rank, repr(e)))
            grp.bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    audio = np.concatenate([o["audio"] for o in out], axis=0)
    pcm = np.concatenate([o["pcm"] for o in out], axis=0) if want_int16 else None
    return audio, pcm, [o["stats"] for o in out]


@pytest.mark.parametrize("chain,style", [("v2", "standard"), ("v1", "edm")])
def test_time_split_equals_whole_file_and_oracle(gpu_lib, chain, style):
    from mm_b200 import pipeline as P, synth
    from oracle import chain as oc
    sr = 44100
    x = synth.numpy_track(3, sr, 16.0)                        # 705600 frames; margin at 44.1 kHz: 262144
    n = x.shape[0]
    target = P.STYLE_CONFIGS[style]["lufs"]
    whole = P.master_batch([x], sr, [style], [target], chain=chain, measure=True, want_int16=True, seed=5)
    for world in (2, 3):
        audio, pcm, stats = _run_split(x, sr, style, chain, world, want_int16=True)
        assert audio.shape == (n, 2)
        d = float(np.max(np.abs(audio.astype(np.float64) - whole["audio"][0].astype(np.float64))))
        print(f"[parity] time split {chain}/{style} world {world}: max|split - whole| = {d:.3e}")
        assert d <= 2e-6
        for s in stats:                                       # every rank ends up with the file's global scalars
            for k in ("lufs_in", "lufs_out", "gain_db", "peak_in", "peak_out"):
                assert abs(s[k] - whole["stats"][0][k]) <= 1e-6, (k, s[k], whole["stats"][0][k])
        assert np.mean(pcm != whole["int16"][0]) < 1e-4 and np.max(np.abs(pcm.astype(np.int32) - whole["int16"][0].astype(np.int32))) <= 1
    ref = (oc.run_v1 if chain == "v1" else oc.run_v2)(x.copy(), sr, target, style)
    e = float(np.max(np.abs(audio.astype(np.float64) - ref.astype(np.float64))))
    print(f"[parity] time split {chain}/{style}: max|split - oracle| = {e:.3e}")
    assert e <= 1e-4


def test_margin_is_what_decouples_the_ranks(gpu_lib):
    """With no margin the cut shows (start-up transients of the low cut-off sections); with the library's margin it
    does not: the margin, not luck, is what makes the split exact."""
    from mm_b200 import pipeline as P, synth
    sr = 44100
    x = synth.numpy_track(4, sr, 8.0)
    whole = P.master_batch([x], sr, ["standard"], [-14.0], chain="v2")["audio"][0].astype(np.float64)
    bad, _, _ = _run_split(x, sr, "standard", "v2", 2, margin=0)
    good, _, _ = _run_split(x, sr, "standard", "v2", 2)
    assert np.max(np.abs(bad - whole)) > 1e-3
    assert np.max(np.abs(good - whole)) <= 2e-6


@pytest.mark.parametrize("world", [2, 8])
def test_time_split_96k_against_oracle(gpu_lib, world):
    """BASELINE configs[4] in miniature: a 64 s 96 kHz stereo file split over 2 and 8 virtual ranks with the library's own margin
    (524288 frames per cut side at 96 kHz, W ~ 6-tile halos of the 30-40 Hz sections) against the CPU oracle and the whole-file
    run: samples 1e-4 (measured ~5e-7), every rank ends up with the file's loudness / gain / peak."""
    from mm_b200 import longform, pipeline as P, synth
    from oracle import chain as oc
    sr = 96000
    x = synth.numpy_track(5, sr, 64.0)
    n = x.shape[0]
    assert longform.slice_margin(sr) == 524288
    for chain, style in (("v2", "standard"),) + ((("v1", "edm"),) if world == 2 else ()):
        target = P.STYLE_CONFIGS[style]["lufs"]
        whole = P.master_batch([x], sr, [style], [target], chain=chain, measure=True)
        audio, _, stats = _run_split(x, sr, style, chain, world)
        assert audio.shape == (n, 2)
        d = float(np.max(np.abs(audio.astype(np.float64) - whole["audio"][0].astype(np.float64))))
        ref = (oc.run_v1 if chain == "v1" else oc.run_v2)(x.copy(), sr, target, style)
        e = float(np.max(np.abs(audio.astype(np.float64) - ref.astype(np.float64))))
        print(f"[parity] 96 kHz time split {chain}/{style} world {world}: max|split - whole| = {d:.3e}, max|split - oracle| = {e:.3e}")
        assert d <= 2e-6 and e <= 1e-4
        for s in stats:
            for k in ("lufs_in", "lufs_out", "gain_db", "peak_in", "peak_out"):
                assert abs(s[k] - whole["stats"][0][k]) <= 1e-6, (k, s[k], whole["stats"][0][k])
        assert abs(stats[0]["lufs_out"] - oc.measure_lufs(ref, sr)) <= 0.01
