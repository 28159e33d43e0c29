"""Determinism soak at bench size: the 64-track v2 batch through mm_dev_master with 1, 2 and 4 lanes, several repetitions each --
every float32 sample, every int16 sample and every stats record must be identical across repetitions and lane counts."""
import ctypes as C
import hashlib
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "audio-mastering-web_b200")]
import torch  # noqa: E402
from mm_b200 import _lib, pipeline as P, shard, synth  # noqa: E402
from mm_b200.engine import get_engine, style_struct  # noqa: E402


def main():
    tracks, sr, dur = 64, 44100, 180.0
    eng = get_engine()
    n = int(round(sr * dur))
    src = eng.empty(tracks, 2, n, sr)
    names = list(P.STYLE_CONFIGS)
    with torch.cuda.stream(eng.stream):
        src.t.zero_()
        synth.torch_batch(list(range(tracks)), sr, dur, eng.tdev, out=src.t, row_stride=src.stride, lead=_lib.MM_LEAD)
        pcm = torch.empty((tracks, n, 2), dtype=torch.int16, device=eng.tdev)
        stats = torch.empty((tracks, shard.STATS_DOUBLES), dtype=torch.float64, device=eng.tdev)
    out = eng.like(src)
    arr = (_lib.Style * tracks)(*[style_struct(P.STYLE_CONFIGS[names[t % 8]], P.STYLE_CONFIGS[names[t % 8]]["lufs"]) for t in range(tracks)])
    g = src.geom
    ref = {}
    keep = {}
    for lanes in (1, 2, 4, 2, 1):
        eng.set_lanes(lanes)
        for rep in range(3):
            with torch.cuda.stream(eng.stream):
                out.t.zero_(); pcm.zero_(); stats.zero_()
            _lib.check(eng.lib.mm_dev_master(eng.ctx, C.byref(g), _lib.CHAIN_V2, arr, src.ptr, out.ptr, C.c_void_p(pcm.data_ptr()), None,
                                             77, C.c_void_p(stats.data_ptr()), _lib.FLAG_MEASURE_OUT | _lib.FLAG_MEASURE_IN))
            eng.sync()
            with torch.cuda.stream(eng.stream):
                sig = (int(out.live().view(torch.int32).to(torch.int64).sum().item()), int(pcm.to(torch.int64).sum().item()),
                       int((pcm.to(torch.int64) * 31 % 1000003).sum().item()),
                       hashlib.sha1(stats.cpu().numpy().tobytes()).hexdigest())
            if lanes not in ref:
                ref[lanes] = sig
                with torch.cuda.stream(eng.stream):
                    keep[lanes] = (out.live().clone(), pcm.clone(), stats.clone())
            ok = sig == ref[lanes]
            print("lanes", lanes, "rep", rep, "identical to the first run with this lane count" if ok else f"DIFFERENT {sig} vs {ref[lanes]}", flush=True)
            if not ok:
                sys.exit(1)
    # across lane counts: the sub-batches segment their sweeps differently (segment starts are rebuilt from halos)
    a, pa, sa = keep[1]
    for lanes in (2, 4):
        b, pb, sb = keep[lanes]
        with torch.cuda.stream(eng.stream):
            ndiff = int((a != b).sum().item())
            maxd = float((a.double() - b.double()).abs().max().item())
            ulp = int((a.view(torch.int32).to(torch.int64) - b.view(torch.int32).to(torch.int64)).abs().max().item())
            pd = int((pa != pb).sum().item())
            sd = float((sa - sb).abs().max().item())
        print(f"lanes {lanes} vs 1: float32 samples that differ {ndiff} of {a.numel()} (max |d| {maxd:.3e}, max {ulp} ulp), int16 samples that differ {pd}, "
              f"max |d stats| {sd:.3e}")
    print("soak ok")


if __name__ == "__main__":
    main()
