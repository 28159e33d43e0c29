#!/usr/bin/env python
"""Golden vectors of the FFT-class stages, produced by the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_fft.py      ->  tests/golden/fft_stages.npz

* apply_spectral_denoise (backend/app/pipeline.py:1472-1524): 2048/512 STFT Wiener gain with a per-bin percentile noise floor
* resample_audio (pipeline.py:920-936): scipy.signal.resample, whole-signal FFT
* apply_harmonic_exciter(oversample=2|4) (pipeline.py:1267-1326): FFT up-sampling, side chain at the high rate, FFT down-sampling

Input: tones + a noise floor + bursts (so the percentile floor, the median cap and the gain clip all engage), seeded.
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


def material(n, sr, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    tone = 0.25 * np.sin(2 * np.pi * 220.0 * t) + 0.12 * np.sin(2 * np.pi * 1760.0 * t + 0.3) + 0.05 * np.sin(2 * np.pi * 7040.0 * t)
    gate = ((np.arange(n) % 9000) < 5000).astype(np.float64)          # the tones pause: frames of pure noise floor
    hiss = 0.01 * rng.standard_normal((n, 2))
    x = np.stack([tone * gate, 0.8 * tone * gate], axis=1) + hiss
    return x.astype(np.float32)


def main():
    P = ref_harness.load().pipeline
    sr, n = 48000, 30000
    x = material(n, sr, 7)
    odd = np.ascontiguousarray(x[:20011])                              # not a multiple of the hop: scipy pads the tail
    st = {
        "input": x, "sr": np.int64(sr),
        "denoise_medium": P.apply_spectral_denoise(x, sr, strength=0.5, noise_percentile=15.0),
        "denoise_strong_odd": P.apply_spectral_denoise(odd, sr, strength=0.9, noise_percentile=20.0),
        "denoise_mono_short": P.apply_spectral_denoise(np.ascontiguousarray(x[:2500, 0]), sr, strength=0.35, noise_percentile=10.0),
        "denoise_p37": P.apply_spectral_denoise(np.ascontiguousarray(x[:12345, 1]), sr, strength=1.0, noise_percentile=37.5),
        "resample_48_44": P.resample_audio(x, 48000, 44100),
        "resample_44_48": P.resample_audio(odd, 44100, 48000),
        "resample_mono_96": P.resample_audio(np.ascontiguousarray(x[:9999, 0]), 48000, 96000),
        "resample_down_even": P.resample_audio(np.ascontiguousarray(x[:20000]), 48000, 24000),
        "exciter_os2": P.apply_harmonic_exciter(x * np.float32(2.0), sr, exciter_db=2.0, mode="tape", oversample=2),
        "exciter_os4_mono": P.apply_harmonic_exciter(np.ascontiguousarray(x[:15001, 0]) * np.float32(3.0), sr, exciter_db=1.5,
                                                     mode="warm", oversample=4),
    }
    st = {k: (np.asarray(v, dtype=np.float32) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    # apply_dynamic_eq (pipeline.py:1628-1700) with STABLE bands (q < 1: the reference's iirpeak(w0, bw) call is a stable
    # section only then; its default bands are not -- tests/test_host_design.py)
    st["dyneq_bands"] = np.array([[3000, 0.5, -30, 3.0, 5, 60, -6], [6000, 0.7, -34, 4.0, 2, 40, -8], [200, 0.3, -28, 2.0, 10, 100, -4],
                                  [23900, 0.5, -20, 2.0, 5, 50, -3]], dtype=np.float64)      # the last one is skipped (>= 0.98 nyq)
    keys = ("freq", "q", "threshold_db", "ratio", "attack_ms", "release_ms", "max_cut_db")
    bands = [dict(zip(keys, row)) for row in st["dyneq_bands"]]
    st["dyneq_stereo"] = P.apply_dynamic_eq(x * np.float32(3.0), sr, bands)
    st["dyneq_mono"] = P.apply_dynamic_eq(np.ascontiguousarray(x[:9001, 0]) * np.float32(2.0), sr, bands[:2])
    st["dyneq_skipped"] = P.apply_dynamic_eq(x * np.float32(30.0), sr, bands[3:])
    # the two remaining public stage functions of pipeline.py: the multiband stage on its own (:414-481) and the lookahead
    # maximizer (:548-573)
    st["multiband_only"] = P.apply_multiband_dynamics(x * np.float32(3.0), sr)
    st["multiband_only_custom"] = P.apply_multiband_dynamics(np.ascontiguousarray(x[:, 0]) * np.float32(2.0), sr, knee_db=4.0,
                                                              crossovers_hz=(214.0, 2230.0, 10000.0), band_ratios=(0.8, 2.0, 1.0, 3.0))
    st["lookahead"] = P.apply_maximizer_lookahead(x * np.float32(3.0), sr, lookahead_ms=6.0)
    st["lookahead_mono_3ms"] = P.apply_maximizer_lookahead(np.ascontiguousarray(x[:5000, 1]) * np.float32(4.0), sr, lookahead_ms=3.0)
    st["lookahead_bypass"] = P.apply_maximizer_lookahead(x[:100] * np.float32(4.0), sr, lookahead_ms=6.0)
    # export_audio(auto_blank_sec=...) (pipeline.py:900-918, :976-977): trailing silence cut; kept lengths from the reference
    tail = x.copy()
    tail[17000:] *= np.float32(1e-4)                                     # below -50 dBFS after frame 17000
    st["blank_input"] = tail
    for name, sig, sec in (("blank_len_03", tail, 0.3), ("blank_len_mono_01", np.ascontiguousarray(tail[:, 0]), 0.1),
                           ("blank_len_none", x, 0.2), ("blank_len_all_quiet", tail[17100:], 0.05)):
        wav = P.export_audio(sig, sr, sig.ndim, "wav", auto_blank_sec=sec)
        st[name] = np.int64((len(wav) - 44) // (2 * (sig.shape[1] if sig.ndim > 1 else 1)))
    path = os.path.join(HERE, "fft_stages.npz")
    np.savez_compressed(path, **st)
    print({k: np.shape(v) for k, v in st.items()}, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
