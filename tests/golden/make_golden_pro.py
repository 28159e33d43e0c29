#!/usr/bin/env python
"""Golden vectors of the second-wave ("PRO") stages, produced by the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_pro.py      ->  tests/golden/pro_stages_48k.npz

Same harness as make_golden.py (oracle/ref_harness.py).  Input: the reference's own seeded recipe
(backend/tests/test_mastering_regression_windows.py:32-36: default_rng(42), 48 kHz, sigma 0.04), plain and x6.
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


def main():
    P = ref_harness.load().pipeline
    sr, n = 48000, 24000
    x = (0.04 * np.random.default_rng(42).standard_normal((n, 2))).astype(np.float32)
    x[:, 1] = (0.6 * x[:, 0] + 0.8 * x[:, 1]).astype(np.float32)
    loud = (x * np.float32(6.0)).astype(np.float32)
    # a percussive variant: bursts make the fast and the slow follower part ways
    env = (np.arange(n) % 6000 < 600).astype(np.float32) * 0.9 + 0.1
    perc = (loud * env[:, None]).astype(np.float32)
    st = {
        "input": x, "sr": np.int64(sr), "perc": perc,
        "transient_punch": P.apply_transient_designer(perc, sr, attack_gain=1.6, sustain_gain=0.8),
        "transient_soft": P.apply_transient_designer(perc, sr, attack_gain=0.7, sustain_gain=1.3),
        "transient_mono": P.apply_transient_designer(np.ascontiguousarray(perc[:, 0]), sr, attack_gain=1.4, sustain_gain=1.0),
        "maximizer_ta": P.apply_maximizer_transient_aware(perc, sr, sensitivity=0.5),
        "maximizer_ta_mono": P.apply_maximizer_transient_aware(np.ascontiguousarray(perc[:, 0]), sr, sensitivity=0.8),
        "hf_trim": P.apply_high_freq_trim(loud, sr),
        "hf_trim_custom": P.apply_high_freq_trim(x, sr, crossover_hz=3000.0, high_gain=0.8),
        "haas": P.apply_stereo_imager(x, 1.2, stereoize_delay_ms=8.0, stereoize_mix=0.12, sr=sr),
        "imager4": P.apply_stereo_imager(loud, 1.0, sr=sr, band_widths=(0.8, 1.0, 1.3, 1.6)),
        "imager4_haas": P.apply_stereo_imager(x, 1.0, stereoize_delay_ms=6.0, stereoize_mix=0.2, sr=sr, band_widths=(1.0, 1.2, 1.4, 0.9),
                                              crossovers_hz=(214.0, 2230.0, 10000.0)),
        "linear_phase": P.apply_target_curve(loud, sr, phase_mode="linear_phase"),
        "linear_phase_ms": P.apply_target_curve(x, sr, phase_mode="linear_phase", eq_ms=True),
        "linear_phase_mono_short": P.apply_target_curve_linear_phase(np.ascontiguousarray(loud[:3000, 0]), sr),
        "reverb_plate": P.apply_reverb(loud, sr, "plate", 1.2, 0.15),
        "reverb_hall_ms": P.apply_reverb(perc, sr, "hall", 0.0, 0.2, mix_mid=0.1, mix_side=0.35),
        "reverb_room_mono": P.apply_reverb(np.ascontiguousarray(perc[:, 0]), sr, "room", 0.6, 0.3),
        "haas_loud": P.apply_stereo_imager(loud, 1.0, stereoize_delay_ms=12.0, stereoize_mix=0.3, sr=sr),
    }
    # reference mastering: a brighter, bass-lighter "reference" for the same material
    from scipy import signal as _sg
    bb, aa = _sg.butter(1, 2000 / (sr / 2), "high")
    refsig = (loud + 0.8 * _sg.lfilter(bb, aa, loud, axis=0)).astype(np.float32)
    st["refmatch_reference"] = refsig
    st["refmatch_env_src"] = P.compute_spectral_envelope(loud, sr)
    st["refmatch_env_ref"] = P.compute_spectral_envelope(refsig, sr)
    st["refmatch_out"] = P.apply_reference_match(loud, sr, refsig, sr, strength=0.8)
    st["refmatch_out_mono"] = P.apply_reference_match(np.ascontiguousarray(loud[:, 0]), sr, refsig, sr, strength=1.0)
    st = {k: (np.asarray(v, dtype=np.float32) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    # noise-shaped dither export: the reference draws from the legacy global generator, so seeding it pins the uniforms
    short = loud[:6000]
    for kind, seed in (("ns_e", 321), ("ns_itu", 654)):
        np.random.seed(seed)
        wav = P.export_audio(short, sr, 2, "wav", dither_type=kind)
        st[f"{kind}_int16"] = np.frombuffer(wav[44:], dtype="<i2").reshape(short.shape)
        st[f"{kind}_seed"] = np.int64(seed)
    path = os.path.join(HERE, "pro_stages_48k.npz")
    np.savez_compressed(path, **st)
    print({k: np.shape(v) for k, v in st.items()}, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
