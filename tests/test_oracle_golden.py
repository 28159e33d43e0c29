"""The oracle restatement (oracle/chain.py, oracle/bs1770.py) against the golden vectors that
tests/golden/make_golden.py produced by running the unmodified reference."""
import numpy as np
import pytest

from conftest import fft_stage_cases, load_golden, tpdf_noise
from oracle import chain as oc
from mm_b200 import synth

CHAIN_CASES = [
    ("v1_edm_44k", 0, 0.6, 2, "v1", "edm"),
    ("v1_standard_96k", 1, 0.6, 2, "v1", "standard"),
    ("v2_standard_48k", 7, 0.75, 2, "v2", "standard"),
    ("v2_hiphop_44k_mono", 3, 0.75, 1, "v2", "hiphop"),
    ("v1_podcast_48k", 5, 0.6, 2, "v1", "podcast"),
    ("v2_house_44k", 11, 0.6, 2, "v2", "house_basic"),
]


@pytest.mark.parametrize("name,track,dur,ch,which,style", CHAIN_CASES)
def test_chain_matches_reference_golden(name, track, dur, ch, which, style):
    g = load_golden(name)
    sr, target = int(g["sr"]), float(g["target"])
    x = synth.numpy_track(track, sr, dur)
    x = x if ch == 2 else np.ascontiguousarray(x[:, 0])
    assert np.array_equal(x, g["input"]), "synthetic generator drifted from the stored golden input"
    stages = {}
    out = (oc.run_v1 if which == "v1" else oc.run_v2)(x.copy(), sr, target, style, stages=stages)
    # float32 samples: the restatement calls the same scipy routines -> expect (near) bit equality
    assert np.max(np.abs(out.astype(np.float64) - g["out"])) <= 1e-6
    for k in g:
        if k.startswith("stage_"):
            assert np.max(np.abs(stages[k[6:]].astype(np.float64) - g[k])) <= 1e-6, k
    assert abs(oc.measure_lufs(out, sr) - float(g["lufs_out"])) < 1e-6
    assert abs(oc.measure_lufs(x, sr) - float(g["lufs_in"])) < 1e-6
    assert abs(oc.true_peak_dbfs(out) - float(g["true_peak_out"])) < 1e-6
    noise = tpdf_noise(g["noise_seed"], g["int16"].shape)
    q = oc.quantize_int16(g["out"].reshape(g["int16"].shape), noise)
    assert np.array_equal(q, g["int16"])


def test_single_stages_match_reference_golden():
    g = load_golden("stages_noise_48k")
    sr = int(g["sr"])
    x = g["input"]
    loud = (x * np.float32(6.0)).astype(np.float32)
    got = {
        "dc": oc.remove_dc_offset(x + np.float32(0.01)),
        "peak_guard_loud": oc.remove_intersample_peaks(loud, 0.5),
        "target_curve": oc.apply_target_curve(x, sr),
        "target_curve_ms": oc.apply_target_curve(x, sr, eq_ms=True),
        "deesser_loud": oc.apply_deesser(loud, sr),
        "dynamics_v1": oc.apply_dynamics(loud, sr),
        "dynamics_v2": oc.apply_dynamics(loud, sr, crossovers_hz=(214.0, 2230.0, 10000.0)),
        "dynamics_upward": oc.apply_dynamics(x, sr, band_ratios=(0.8, 2.0, 1.0, 0.6)),
        "parallel": oc.apply_parallel_compression(loud, sr, mix=0.3),
        "normalize": oc.normalize_lufs(x, sr, -14.0),
        "final_balance": oc.apply_final_spectral_balance(x, sr),
        "style_eq_edm": oc.apply_style_eq(x, sr, "edm"),
        "style_eq_classical": oc.apply_style_eq(x, sr, "classical"),
        "exciter": oc.apply_harmonic_exciter(loud, sr, 0.8),
        "imager": oc.apply_stereo_imager(x, 1.3),
        "fade": oc.apply_output_edge_fade_in(x, sr, 6.0),
        "maximizer": oc.apply_maximizer(loud),
    }
    for k, v in got.items():
        err = np.max(np.abs(np.asarray(v, dtype=np.float64) - g[k]))
        assert err <= 2e-7, (k, err)


def test_pro_stages_match_reference_golden():
    """Second-wave stages against tests/golden/make_golden_pro.py's run of the reference."""
    g = load_golden("pro_stages_48k")
    sr, x, perc = int(g["sr"]), g["input"], g["perc"]
    loud = (x * np.float32(6.0)).astype(np.float32)
    got = {
        "transient_punch": oc.apply_transient_designer(perc, sr, 1.6, 0.8),
        "transient_soft": oc.apply_transient_designer(perc, sr, 0.7, 1.3),
        "transient_mono": oc.apply_transient_designer(np.ascontiguousarray(perc[:, 0]), sr, 1.4, 1.0),
        "maximizer_ta": oc.apply_maximizer_transient_aware(perc, sr, 0.5),
        "maximizer_ta_mono": oc.apply_maximizer_transient_aware(np.ascontiguousarray(perc[:, 0]), sr, 0.8),
        "hf_trim": oc.apply_high_freq_trim(loud, sr),
        "hf_trim_custom": oc.apply_high_freq_trim(x, sr, 3000.0, 0.8),
        "haas": oc.apply_stereoize(x, sr, 1.2, 8.0, 0.12),
        "haas_loud": oc.apply_stereoize(loud, sr, 1.0, 12.0, 0.3),
        "linear_phase": oc.apply_target_curve_linear_phase(loud, sr),
        "linear_phase_mono_short": oc.apply_target_curve_linear_phase(np.ascontiguousarray(loud[:3000, 0]), sr),
        "reverb_plate": oc.apply_reverb(loud, sr, "plate", 1.2, 0.15),
        "reverb_hall_ms": oc.apply_reverb(perc, sr, "hall", 0.0, 0.2, mix_mid=0.1, mix_side=0.35),
        "reverb_room_mono": oc.apply_reverb(np.ascontiguousarray(perc[:, 0]), sr, "room", 0.6, 0.3),
        "imager4": oc.apply_stereo_imager_4band(loud, sr, (0.8, 1.0, 1.3, 1.6)),
        "imager4_haas": oc.apply_stereo_imager_4band(x, sr, (1.0, 1.2, 1.4, 0.9), (214.0, 2230.0, 10000.0), 6.0, 0.2),
    }
    for k, v in got.items():
        assert np.shape(v) == g[k].shape, k
        err = np.max(np.abs(np.asarray(v, dtype=np.float64) - g[k]))
        assert err <= 1e-7, (k, err)      # measured: bit equal (numba's fastmath does not change these roundings here)


def test_reference_match_matches_reference_golden():
    g = load_golden("pro_stages_48k")
    sr = int(g["sr"])
    loud = (g["input"] * np.float32(6.0)).astype(np.float32)
    ref = g["refmatch_reference"]
    assert np.max(np.abs(oc.compute_spectral_envelope(loud, sr) - g["refmatch_env_src"])) <= 1e-6
    assert np.max(np.abs(oc.apply_reference_match(loud, sr, ref, sr, 0.8).astype(np.float64) - g["refmatch_out"])) <= 1e-6
    assert np.max(np.abs(oc.apply_reference_match(np.ascontiguousarray(loud[:, 0]), sr, ref, sr, 1.0).astype(np.float64) - g["refmatch_out_mono"])) <= 1e-6


def test_fft_class_stages_match_reference_golden():
    """apply_spectral_denoise / resample_audio / oversampled exciter (pipeline.py:1472-1524, :920-936, :1294-1320): the
    written-out stft/istft and rfft-resample restatements reproduce the reference's outputs (measured: bit equal)."""
    g = load_golden("fft_stages")
    for k, call in fft_stage_cases(oc, g).items():
        v = call()
        assert np.shape(v) == g[k].shape and v.dtype == np.float32, k
        err = np.max(np.abs(np.asarray(v, dtype=np.float64) - g[k]))
        assert err <= 1e-7, (k, err)
    with pytest.raises(ValueError):
        oc.apply_spectral_denoise(g["input"][:1500], 48000, 0.5)
    x = g["input"]
    assert oc.apply_spectral_denoise(x, 48000, 0.005) is x
    assert oc.resample_audio(x, 48000, 48000).dtype == np.float32


def dyneq_cases(mod, g):
    x, sr = g["input"], int(g["sr"])
    keys = ("freq", "q", "threshold_db", "ratio", "attack_ms", "release_ms", "max_cut_db")
    bands = [dict(zip(keys, row)) for row in g["dyneq_bands"]]
    return {"dyneq_stereo": lambda: mod.apply_dynamic_eq(x * np.float32(3.0), sr, bands),
            "dyneq_mono": lambda: mod.apply_dynamic_eq(np.ascontiguousarray(x[:9001, 0]) * np.float32(2.0), sr, bands[:2]),
            "dyneq_skipped": lambda: mod.apply_dynamic_eq(x * np.float32(30.0), sr, bands[3:])}


def test_dynamic_eq_default_bands_match_reference_golden():
    """apply_dynamic_eq(audio, sr) with the reference's default bands: the oracle makes the same scipy calls as the reference
    (overflowing / degenerate sections included) and must return its outputs bit for bit."""
    import warnings
    from conftest import dyneq_default_cases
    g = load_golden("dyneq_default")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, x, sr, dec in dyneq_default_cases():
            if len(x) > 200000:
                continue                                  # the 20 s cases: GPU test / tests/test_oracle_vs_reference.py
            out = oc.apply_dynamic_eq(x, sr, oc.DYNAMIC_EQ_MASTERING_BANDS)
            assert np.array_equal(out[::dec], g[name]), name


def test_dynamic_eq_stable_bands_match_reference_golden():
    """apply_dynamic_eq (pipeline.py:1628-1700) with bands whose iirpeak(w0, bw) section is stable (q < 1)."""
    g = load_golden("fft_stages")
    for k, call in dyneq_cases(oc, g).items():
        v = call()
        assert v.shape == g[k].shape and v.dtype == np.float32, k
        assert np.max(np.abs(v.astype(np.float64) - g[k])) <= 1e-7, k
    # the gain engages: the result differs from the merely clipped input
    assert np.max(np.abs(g["dyneq_stereo"] - np.clip(g["input"] * np.float32(3.0), -1, 1))) > 1e-2


def tail_stage_cases(mod, g):
    x, sr = g["input"], int(g["sr"])
    return {"multiband_only": lambda: mod.apply_multiband_dynamics(x * np.float32(3.0), sr),
            "multiband_only_custom": lambda: mod.apply_multiband_dynamics(np.ascontiguousarray(x[:, 0]) * np.float32(2.0), sr, knee_db=4.0,
                                                                          crossovers_hz=(214.0, 2230.0, 10000.0), band_ratios=(0.8, 2.0, 1.0, 3.0)),
            "lookahead": lambda: mod.apply_maximizer_lookahead(x * np.float32(3.0), sr, lookahead_ms=6.0),
            "lookahead_mono_3ms": lambda: mod.apply_maximizer_lookahead(np.ascontiguousarray(x[:5000, 1]) * np.float32(4.0), sr, lookahead_ms=3.0),
            "lookahead_bypass": lambda: mod.apply_maximizer_lookahead(x[:100] * np.float32(4.0), sr, lookahead_ms=6.0)}


def test_multiband_only_and_lookahead_maximizer_match_reference_golden():
    """apply_multiband_dynamics (pipeline.py:414-481) and apply_maximizer_lookahead (:548-573)."""
    g = load_golden("fft_stages")
    for k, call in tail_stage_cases(oc, g).items():
        v = call()
        assert v.shape == g[k].shape, k
        assert np.max(np.abs(np.asarray(v, dtype=np.float64) - g[k])) <= 1e-7, k


def test_auto_blank_end_matches_reference_golden():
    """export_audio(auto_blank_sec=...) (pipeline.py:900-918, :976-977): kept lengths from the reference's WAV sizes."""
    g = load_golden("fft_stages")
    tail, x, sr = g["blank_input"], g["input"], int(g["sr"])
    for name, sig, sec in (("blank_len_03", tail, 0.3), ("blank_len_mono_01", np.ascontiguousarray(tail[:, 0]), 0.1),
                           ("blank_len_none", x, 0.2), ("blank_len_all_quiet", tail[17100:], 0.05)):
        assert oc.export_prepare(sig, sr, sec).shape[0] == int(g[name]), name


def test_noise_shaped_dither_export_matches_reference_golden():
    """ns_e / ns_itu (pipeline.py:835-877): the oracle's shaping of the same uniforms + quantiser == the reference's WAV."""
    g = load_golden("pro_stages_48k")
    x = (g["input"] * np.float32(6.0)).astype(np.float32)[:6000]
    for kind in ("ns_e", "ns_itu"):
        np.random.seed(int(g[f"{kind}_seed"]))
        uniform = np.random.rand(*x.shape).astype(np.float32)
        q = oc.quantize_int16(x, oc.dither_noise_shaped(uniform, kind))
        assert np.array_equal(q, g[f"{kind}_int16"]), kind


def test_analyzers_match_reference_golden():
    g = load_golden("analyzers")
    for tag in "abc":
        x, sr = g[f"{tag}_input"], int(g[f"{tag}_sr"])
        assert abs(oc.measure_lufs(x, sr) - float(g[f"{tag}_lufs"])) < 1e-9
        assert abs(oc.measure_lufs(np.ascontiguousarray(x[:, 0]), sr) - float(g[f"{tag}_lufs_mono"])) < 1e-9
        assert abs(oc.true_peak_dbfs(x) - float(g[f"{tag}_true_peak"])) < 1e-9
        assert abs(oc.true_peak_explicit(x) - float(g[f"{tag}_true_peak"])) < 1e-9
        assert np.allclose(oc.compute_spectrum_bars(x, sr), g[f"{tag}_bars"], atol=1e-9)
        assert np.allclose(oc.compute_spectrum_bars((x[:, 0] + x[:, 1]) * 0.5, sr), g[f"{tag}_bars_mid"], atol=1e-9)
        assert abs(oc.measure_stereo_correlation(x) - float(g[f"{tag}_corr"])) < 1e-12
        tl, step = oc.compute_lufs_timeline(x, sr)
        ref_tl = g[f"{tag}_timeline"]
        assert len(tl) == len(ref_tl) and step == float(g[f"{tag}_timeline_step"])
        for a, b in zip(tl, ref_tl):
            assert (a is None and np.isnan(b)) or abs(a - b) < 1e-9
        assert np.allclose(np.array(oc.compute_vectorscope_points(x)), g[f"{tag}_vscope"])


def test_filtfilt_explicit_is_scipy_filtfilt():
    from scipy import signal as sg
    rng = np.random.default_rng(5)
    x = rng.standard_normal(5000)
    for b, a in (sg.butter(2, 40 / 22050, "high"), sg.butter(1, [0.1, 0.2], "band"), sg.butter(2, [0.22, 0.41], "band")):
        assert np.max(np.abs(oc.filtfilt_explicit(b, a, x) - sg.filtfilt(b, a, x))) < 1e-12
