#!/usr/bin/env python
"""Golden vectors of apply_dynamic_eq with the reference's DEFAULT bands (backend/app/pipeline.py:1616-1700), produced by the
UNMODIFIED reference (build container only):

    python tests/golden/make_golden_dyneq.py      ->  tests/golden/dyneq_default.npz

The default bands are unstable / degenerate iirpeak sections (the reference passes a bandwidth as Q); what the reference
returns for them -- identity for the overflowing ones, ``x * g`` for the q = 1 bands, ``x - const`` for the 12 kHz band at
48 kHz -- is what mm_b200 must reproduce (csrc/deesser.cu st_dynamic_eq).  Inputs are the deterministic synthetic tracks of
``mm_b200.synth.numpy_track`` (reproducible from the track id, so only outputs are stored), scaled up so that the degenerate
bands' compressors engage.  The 20 s cases keep every 16th frame of the output.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "audio-mastering-web_b200"))

from oracle import ref_harness  # noqa: E402
from mm_b200 import synth  # noqa: E402

# name -> (track id, sample rate, seconds, channels, scale, decimation of the stored output)
CASES = {
    "d44_2s": (3, 44100, 2.0, 2, 1.0, 1),
    "d48_2s": (4, 48000, 2.0, 2, 1.0, 1),
    "d48_mono_1s": (5, 48000, 1.0, 1, 1.5, 1),
    "d44_20s": (6, 44100, 20.0, 2, 1.0, 16),
    "d48_20s": (7, 48000, 20.0, 2, 1.0, 16),
    "d96_2s": (8, 96000, 2.0, 2, 1.0, 4),
}


def case_input(name):
    t, sr, dur, ch, scale, _ = CASES[name]
    x = synth.numpy_track(t, sr, dur, channels=ch) * np.float32(scale)
    return (np.ascontiguousarray(x[:, 0]) if ch == 1 else x), sr


def main():
    import warnings
    warnings.filterwarnings("ignore")            # the reference's overflowing lfilter passes warn by design
    P = ref_harness.load().pipeline
    st = {}
    for name, (_, _, _, _, _, dec) in CASES.items():
        x, sr = case_input(name)
        y = P.apply_dynamic_eq(x, sr)
        assert y.dtype == np.float32 and y.shape == x.shape
        st[name] = np.ascontiguousarray(y[::dec])
        print(name, x.shape, "max|y - clip(x)| = %.4f" % float(np.max(np.abs(y - np.clip(x, -1, 1)))))
    path = os.path.join(HERE, "dyneq_default.npz")
    np.savez_compressed(path, **st)
    print("%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
