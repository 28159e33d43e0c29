"""Timeline of the host-buffer pipeline (MM_HOST_TIMELINE=1): 64 PCM_16 uploads of 180 s through mm_master_host_pcm16 with pinned
buffers; the library prints when each chunk's copy-in, layout, chain and copy-out ended.  Prints the wall time of the call too."""
import ctypes as C
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "audio-mastering-web_b200")]
import torch  # noqa: E402
from mm_b200 import _lib, pipeline as P  # noqa: E402
from mm_b200.engine import get_engine, style_struct, TrackStats  # noqa: E402


def main():
    tracks, sr, dur = int(os.environ.get("TRACKS", "64")), 44100, 180.0
    n = int(sr * dur)
    eng = get_engine()
    hin = torch.empty((tracks, n, 2), dtype=torch.int16, pin_memory=True)
    hin.copy_((torch.rand((tracks, n, 2)) * 20000 - 10000).to(torch.int16))
    hout = torch.empty((tracks, n, 2), dtype=torch.int16, pin_memory=True)
    st = (TrackStats * tracks)()
    arr = (_lib.Style * tracks)(*[style_struct(P.STYLE_CONFIGS["standard"], -14.0)] * tracks)

    def call():
        _lib.check(eng.lib.mm_master_host_pcm16(eng.ctx, _lib.CHAIN_V2, tracks, n, 2, sr, arr, C.c_void_p(hin.data_ptr()), None,
                                                C.c_void_p(hout.data_ptr()), 7, st, _lib.FLAG_MEASURE_OUT))
    for _ in range(2):
        call()
    t0 = time.perf_counter()
    for _ in range(3):
        call()
    print("wall ms per call", (time.perf_counter() - t0) / 3 * 1e3, "audio-s/s", tracks * dur * 3 / (time.perf_counter() - t0))


if __name__ == "__main__":
    main()
