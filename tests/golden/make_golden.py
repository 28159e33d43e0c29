#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/backend/app/{pipeline,chain}.py`` through ``oracle/ref_harness.py``
(stand-ins only for the three absent I/O packages; see that file) and writes
``tests/golden/*.npz`` + ``tests/golden/MANIFEST.json``.  The reference tree does not exist on
the GPU box, so these files are what travels.  Inputs are regenerated from seeds by the tests
(``mm_b200.synth.numpy_track`` / ``default_rng``) AND stored, so a generator change cannot
silently detach the goldens from their inputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "audio-mastering-web_b200"))

from oracle import ref_harness  # noqa: E402
from mm_b200 import synth  # noqa: E402


def stagewise_v1(P, audio, sr, target, style):
    """Same calls, same order as run_mastering_pipeline (pipeline.py:1833-1909), keeping every stage."""
    cfg = P.STYLE_CONFIGS.get(style, P.STYLE_CONFIGS["standard"])
    st = {}
    a = st["dc_offset"] = P.remove_dc_offset(audio)
    a = st["peak_guard_in"] = P.remove_intersample_peaks(a, headroom_db=0.5)
    a = st["target_eq"] = P.apply_target_curve(a, sr)
    a = st["deesser"] = P.apply_deesser(a, sr)
    a = st["dynamics"] = P.apply_dynamics(a, sr)
    if cfg.get("parallel_mix", 0.0) > 0.01:
        a = st["parallel_compress"] = P.apply_parallel_compression(a, sr, mix=cfg["parallel_mix"])
    a = st["normalize_lufs"] = P.normalize_lufs(a, sr, target)
    a = st["final_spectral_balance"] = P.apply_final_spectral_balance(a, sr)
    a = st["style_eq"] = P.apply_style_eq(a, sr, style)
    if cfg.get("exciter_db", 0.0) > 0.05:
        a = st["harmonic_exciter"] = P.apply_harmonic_exciter(a, sr, cfg["exciter_db"])
    if abs(cfg.get("imager_width", 1.0) - 1.0) > 0.01:
        a = st["stereo_imager"] = P.apply_stereo_imager(a, cfg["imager_width"])
    a = st["peak_guard_out"] = P.remove_intersample_peaks(a, headroom_db=0.5)
    a = st["output_fade_in"] = P.apply_output_edge_fade_in(a, sr, fade_ms=6.0)
    return {k: np.asarray(v, dtype=np.float32) for k, v in st.items()}


def main():
    ref = ref_harness.load()
    P, C = ref.pipeline, ref.chain
    import scipy

    manifest = {
        "generator": "tests/golden/make_golden.py",
        "reference": "denisok-ai/audio-mastering-web backend/app (version %s)" % getattr(
            __import__("app.version", fromlist=["__version__"]), "__version__", "?"),
        "numpy": np.__version__, "scipy": scipy.__version__,
        "numba": __import__("numba").__version__,
        "pedalboard": "absent (numpy fallback compressor branch)",
        "pyloudnorm": "absent (oracle/bs1770.py stand-in)",
        "cases": {},
    }

    def save(name, **arrs):
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrs)
        manifest["cases"][name] = {k: list(np.shape(v)) for k, v in arrs.items()}
        print(name, {k: np.shape(v) for k, v in arrs.items()}, "%.0f KB" % (os.path.getsize(path) / 1024))

    # ---- chain cases -------------------------------------------------------------------------
    chain_cases = [
        # name, track, sr, dur, channels, chain, style, target
        ("v1_edm_44k", 0, 44100, 0.6, 2, "v1", "edm", -9.0),
        ("v1_standard_96k", 1, 96000, 0.6, 2, "v1", "standard", -14.0),
        ("v2_standard_48k", 7, 48000, 0.75, 2, "v2", "standard", -14.0),
        ("v2_hiphop_44k_mono", 3, 44100, 0.75, 1, "v2", "hiphop", -13.0),
        ("v1_podcast_48k", 5, 48000, 0.6, 2, "v1", "podcast", -16.0),
        ("v2_house_44k", 11, 44100, 0.6, 2, "v2", "house_basic", -10.0),
    ]
    for case_idx, (name, t, sr, dur, ch, which, style, target) in enumerate(chain_cases):
        x = synth.numpy_track(t, sr, dur, channels=2)
        x = x if ch == 2 else np.ascontiguousarray(x[:, 0])
        arrs = {"input": x, "sr": np.int64(sr), "target": np.float64(target)}
        if which == "v1":
            out = P.run_mastering_pipeline(x.copy(), sr, target_lufs=target, style=style)
            if name == "v1_edm_44k":
                for k, v in stagewise_v1(P, x.copy(), sr, target, style).items():
                    arrs["stage_" + k] = v
                assert np.array_equal(np.clip(arrs["stage_output_fade_in"], -1, 1), out)
        else:
            out = C.MasteringChain.default_chain(target_lufs=target, style=style).process(
                x.copy(), sr, target_lufs=target, style=style)
            arrs["chain_out"] = np.asarray(out, dtype=np.float32)
            out = P.apply_output_edge_fade_in(out, sr, fade_ms=6.0)  # routers/mastering.py:583
        out = np.asarray(out, dtype=np.float32)
        arrs["out"] = out
        arrs["lufs_in"] = np.float64(P.measure_lufs(x, sr))
        arrs["lufs_out"] = np.float64(P.measure_lufs(out, sr))
        arrs["true_peak_out"] = np.float64(ref.true_peak_dbfs(out, sr))
        # dither noise is NOT stored (incompressible): tests rebuild it from this seed and recipe
        rng = np.random.default_rng(7000 + case_idx)
        shape2d = (out.shape[0], ch)   # export_audio reshapes mono to (n, 1) before dithering
        noise = (rng.random(shape2d) + rng.random(shape2d) - 1.0).astype(np.float32)
        arrs["noise_seed"] = np.int64(7000 + case_idx)
        orig = P._dither_noise_tpdf
        P._dither_noise_tpdf = lambda shape, _n=noise: _n
        try:
            wav = P.export_audio(out, sr, ch, "wav", dither_type="tpdf")
        finally:
            P._dither_noise_tpdf = orig
        pcm = np.frombuffer(wav[44:], dtype="<i2")
        arrs["int16"] = pcm.reshape(shape2d)
        save(name, **arrs)

    # ---- single-stage cases on the reference's own seeded recipe ------------------------------
    # backend/tests/test_mastering_regression_windows.py:32-36 (default_rng(42), 48 kHz, sigma 0.04)
    sr = 48000
    n = 24000
    x = (0.04 * np.random.default_rng(42).standard_normal((n, 2))).astype(np.float32)
    x[:, 1] = (0.6 * x[:, 0] + 0.8 * x[:, 1]).astype(np.float32)
    loud = (x * np.float32(6.0)).astype(np.float32)
    st = {
        "input": x, "sr": np.int64(sr),
        "dc": P.remove_dc_offset(x + np.float32(0.01)),
        "peak_guard_loud": P.remove_intersample_peaks(loud, 0.5),
        "target_curve": P.apply_target_curve(x, sr),
        "target_curve_ms": P.apply_target_curve(x, sr, eq_ms=True),
        "deesser_loud": P.apply_deesser(loud, sr),
        "dynamics_v1": P.apply_dynamics(loud, sr),
        "dynamics_v2": P.apply_dynamics(loud, sr, crossovers_hz=(214.0, 2230.0, 10000.0)),
        "dynamics_upward": P.apply_dynamics(x, sr, band_ratios=(0.8, 2.0, 1.0, 0.6)),
        "parallel": P.apply_parallel_compression(loud, sr, mix=0.3),
        "normalize": P.normalize_lufs(x, sr, -14.0),
        "final_balance": P.apply_final_spectral_balance(x, sr),
        "style_eq_edm": P.apply_style_eq(x, sr, "edm"),
        "style_eq_classical": P.apply_style_eq(x, sr, "classical"),
        "exciter": P.apply_harmonic_exciter(loud, sr, 0.8),
        "imager": P.apply_stereo_imager(x, 1.3),
        "fade": P.apply_output_edge_fade_in(x, sr, 6.0),
        "maximizer": P.apply_maximizer(loud),
        "rumble": P.apply_rumble_filter(x, sr, 80.0),
    }
    st = {k: (np.asarray(v, dtype=np.float32) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    save("stages_noise_48k", **st)

    # ---- analyzers ---------------------------------------------------------------------------
    an = {}
    for tag, t, sr, dur in (("a", 2, 44100, 1.0), ("b", 4, 48000, 0.8), ("c", 6, 96000, 0.5)):
        x = synth.numpy_track(t, sr, dur)
        an[f"{tag}_input"] = x
        an[f"{tag}_sr"] = np.int64(sr)
        an[f"{tag}_lufs"] = np.float64(P.measure_lufs(x, sr))
        an[f"{tag}_lufs_mono"] = np.float64(P.measure_lufs(np.ascontiguousarray(x[:, 0]), sr))
        an[f"{tag}_true_peak"] = np.float64(ref.true_peak_dbfs(x, sr))
        an[f"{tag}_bars"] = np.array(P.compute_spectrum_bars(x, sr), dtype=np.float64)
        an[f"{tag}_bars_mid"] = np.array(P.compute_spectrum_bars(((x[:, 0] + x[:, 1]) * 0.5), sr), dtype=np.float64)
        an[f"{tag}_corr"] = np.float64(P.measure_stereo_correlation(x))
        tl, step = P.compute_lufs_timeline(x, sr)
        an[f"{tag}_timeline"] = np.array([np.nan if v is None else v for v in tl], dtype=np.float64)
        an[f"{tag}_timeline_step"] = np.float64(step)
        an[f"{tag}_vscope"] = np.array(P.compute_vectorscope_points(x), dtype=np.float64)
    save("analyzers", **an)

    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
