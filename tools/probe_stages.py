"""Per-kernel device times of the FFT-class second-wave stages (spectral denoise, FFT resample, oversampled exciter) on a
batch of synthetic tracks: ``python tools/probe_stages.py --tracks 64 --sec 180``."""
import argparse
import ctypes as C
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "audio-mastering-web_b200"))

import torch

from mm_b200 import pipeline as P
from mm_b200.engine import get_engine


def timed(eng, name, fn, frames_in, reps, bytes_per_frame):
    fn()
    eng.sync()
    best = None
    for _ in range(reps):
        eng.timing(True)
        fn()
        eng.sync()
        kt = eng.kernel_times()
        eng.timing(False)
        tot = sum(v[0] for v in kt.values())
        if best is None or tot < best[0]:
            best = (tot, kt)
    tot, kt = best
    print(f"{name}: {tot:.2f} ms of kernels, {frames_in / tot / 1e6:.1f} G channel-samples/s, "
          f"{frames_in * bytes_per_frame / tot / 1e6:.0f} GB/s of algorithmic traffic ({bytes_per_frame} B per channel-sample)")
    for k, (ms, cnt, _smp) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        print(f"    {k:32s} {ms:9.3f} ms  x{cnt:<4d} {ms / tot * 100:5.1f}%")
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=64)
    ap.add_argument("--sec", type=float, default=180.0)
    ap.add_argument("--sr", type=int, default=44100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    eng = get_engine()
    n = int(a.sec * a.sr)
    b = eng.empty(a.tracks, 2, n, a.sr)
    with torch.cuda.stream(eng.stream):
        b.t.normal_(0.0, 0.1)
    out = eng.like(b)
    cs = a.tracks * 2 * n
    want = set(a.only.split(",")) if a.only else None

    def on(k):
        return want is None or k in want

    if on("denoise"):
        # read x, write |Z| (2 words per sample), read |Z| (select, 4 digit passes), read x, write y
        timed(eng, "apply_spectral_denoise", lambda: eng.stage("apply_spectral_denoise", b, C.c_double(0.5), C.c_double(15.0), out=out),
              cs, a.reps, 4 * (1 + 2 + 2 + 1 + 1))
    if on("resample"):
        num = int(round(n * 48000 / a.sr))
        timed(eng, f"fft_resample {n} -> {num}", lambda: eng.fft_resample(b, num, 48000), cs, a.reps, 4 * 2)
    if on("exciter2"):
        timed(eng, "apply_harmonic_exciter oversample=2", lambda: P._exciter_dev(eng, b, 2.0, "tape", 2), cs, a.reps, 4 * 2)
    if on("linphase"):
        timed(eng, "apply_target_curve_linear_phase (4096 taps)", lambda: eng.stage("apply_target_curve_linear_phase", b, 4096, out=out), cs, a.reps, 4 * 2)
    if on("fir8192"):
        import numpy as np
        ir = np.ascontiguousarray((np.hanning(8192) / 4096.0).astype(np.float32))
        timed(eng, "fir_same (8192 taps, reference match)", lambda: eng.stage("fir_same", b, ir.ctypes.data_as(C.c_void_p), 8192, 1, out=out),
              cs, a.reps, 4 * 2)
    if on("truepeak"):
        timed(eng, "true peak + correlation (fused)", lambda: eng.true_peak_correlation(b), cs, a.reps, 4)
    if on("refenv"):
        env = torch.empty(a.tracks * 4097, dtype=torch.float32, device=eng.tdev)
        g = b.geom
        timed(eng, "compute_spectral_envelope", lambda: eng.lib.mm_dev_spectral_envelope(eng.ctx, C.byref(g), b.ptr, C.c_void_p(env.data_ptr())),
              cs, a.reps, 4)
    print("workspace GB", eng.lib.mm_ctx_workspace_bytes(eng.ctx) / 1e9)


if __name__ == "__main__":
    main()
