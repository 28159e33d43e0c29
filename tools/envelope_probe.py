"""Per-kernel times of the envelope-compressor mode on the bench batch (64 x 180 s x 44.1 kHz stereo, v2 chain): an A/B harness
for experiment builds (MM_B200_LIB=.../libmm_b200_<tag>.so).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "audio-mastering-web_b200")]
import torch  # noqa: E402
from mm_b200 import _lib, pipeline as P, shard, synth  # noqa: E402
from mm_b200.engine import get_engine, style_struct  # noqa: E402


def main():
    tracks, sr, dur = int(os.environ.get("TRACKS", "64")), 44100, 180.0
    eng = get_engine()
    n = int(round(sr * dur))
    src = eng.empty(tracks, 2, n, sr)
    with torch.cuda.stream(eng.stream):
        src.t.zero_()
        synth.torch_batch(list(range(tracks)), sr, dur, eng.tdev, out=src.t, row_stride=src.stride, lead=_lib.MM_LEAD)
        pcm = torch.empty((tracks, n, 2), dtype=torch.int16, device=eng.tdev)
        stats = torch.empty((tracks, shard.STATS_DOUBLES), dtype=torch.float64, device=eng.tdev)
    out = eng.like(src)
    arr = (_lib.Style * tracks)(*[style_struct(P.STYLE_CONFIGS["standard"], -14.0)] * tracks)
    g = src.geom
    flags = _lib.FLAG_MEASURE_OUT | (_lib.FLAG_ENVELOPE_COMPRESSOR if os.environ.get("ENVELOPE", "1") == "1" else 0)
    chain_id = _lib.CHAIN_V1 if os.environ.get("CHAIN", "v2") == "v1" else _lib.CHAIN_V2

    def step(i):
        _lib.check(eng.lib.mm_dev_master(eng.ctx, C.byref(g), chain_id, arr, src.ptr, out.ptr, C.c_void_p(pcm.data_ptr()), None,
                                         1234 + i, C.c_void_p(stats.data_ptr()), flags))
    for i in range(3):
        step(i)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for i in range(10):
        step(i)
    e1.record(eng.stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / 10
    eng.timing(True)
    for i in range(5):
        step(i)
    kt = eng.kernel_times()
    eng.timing(False)
    k = kt.get("band_envelope_compress")
    chk = float(out.live()[0:2, 1000:200000].double().abs().sum().item())
    print(json.dumps({"lib": os.path.basename(os.environ.get("MM_B200_LIB", "libmm_b200.so")), "ms_per_step": ms,
                      "audio_s_per_s": tracks * dur / ms * 1e3, "band_envelope_compress_ms": k[0] / k[1] if k else None, "checksum": chk,
                      "followers_ms": {nm: v[0] / v[1] for nm, v in kt.items() if any(t in nm for t in ("env", "deess", "follow"))}}))


if __name__ == "__main__":
    main()
