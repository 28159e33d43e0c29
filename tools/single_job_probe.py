"""Single-job latency breakdown through the Python drop-in: one 180 s 44.1 kHz stereo track as a pageable numpy array in and out
(what run_mastering_pipeline receives from the job runner, routers/mastering.py:350-441).  Prints one JSON line."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "audio-mastering-web_b200")]
import numpy as np  # noqa: E402
from mm_b200 import _lib, pipeline as P, synth  # noqa: E402
from mm_b200.engine import get_engine, style_struct  # noqa: E402


def med(f, reps=7):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    sr, dur = 44100, 180.0
    x = synth.numpy_track(1000, sr, dur)
    eng = get_engine()
    res = {"threads": os.environ.get("MM_HOST_THREADS", "default")}
    b = eng.upload([x], sr)
    res["upload_ms"] = med(lambda: eng.upload([x], sr))
    res["download_ms"] = med(lambda: eng.download(b))
    st = [style_struct(P.STYLE_CONFIGS["standard"], -14.0)]

    def chain():
        eng.master(b, _lib.CHAIN_V2, st, out=b, want_int16=False)
        eng.sync()
    res["chain_v2_ms"] = med(chain)
    res["master_batch_v2_ms"] = med(lambda: P.master_batch([x], sr, ["standard"], chain="v2", eng=eng))
    res["master_batch_v2_int16_ms"] = med(lambda: P.master_batch([x], sr, ["standard"], chain="v2", eng=eng, want_int16=True))
    res["run_mastering_pipeline_v1_ms"] = med(lambda: P.run_mastering_pipeline(x, sr))
    res["run_mastering_pipeline_v1_stagewise_ms"] = med(lambda: P.run_mastering_pipeline(x, sr, transient_attack=1.3), reps=3)
    res["apply_target_curve_ms"] = med(lambda: P.apply_target_curve(x, sr))
    res["measure_lufs_ms"] = med(lambda: P.measure_lufs(x, sr))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
