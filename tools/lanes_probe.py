"""Experiment: L contexts (own stream, own workspace, persistent grids sized for 1 / MM_GRID_DIV of the device) mastering one-track
batches concurrently from L host threads, against one context doing the same tracks one after the other."""
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "audio-mastering-web_b200")]
import numpy as np  # noqa: E402
from mm_b200 import _lib, pipeline as P, synth  # noqa: E402
from mm_b200.engine import Engine, style_struct  # noqa: E402


def main():
    lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    per_lane = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    tracks_per_launch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sr, dur = 44100, 180.0
    x = synth.numpy_track(1000, sr, dur)
    engs = [Engine(0) for _ in range(lanes)]
    st = [style_struct(P.STYLE_CONFIGS["standard"], -14.0)] * tracks_per_launch
    bs = [e.upload([x] * tracks_per_launch, sr) for e in engs]
    outs = [e.like(b) for e, b in zip(engs, bs)]

    def work(i, reps):
        e = engs[i]
        for _ in range(reps):
            e.master(bs[i], _lib.CHAIN_V2, st, out=outs[i], want_int16=True)
        e.sync()
    for i in range(lanes):
        work(i, 2)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i, per_lane)) for i in range(lanes)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    n = lanes * per_lane * tracks_per_launch
    ref = engs[0].download(outs[0])
    same = all(np.array_equal(ref, engs[i].download(outs[i])) for i in range(1, lanes))
    print(json.dumps({"lanes": lanes, "grid_div": os.environ.get("MM_GRID_DIV", "1"), "tracks_per_launch": tracks_per_launch,
                      "tracks": n, "ms_per_track": dt / n * 1e3, "audio_s_per_s": n * dur / dt, "lanes_equal": same}))


if __name__ == "__main__":
    main()
