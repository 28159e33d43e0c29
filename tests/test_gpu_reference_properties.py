"""The INVARIANTS the reference's own unit tests pin (backend/tests/test_pipeline.py, test_mastering_regression_windows.py;
SURVEY.md section 4: shapes, dtypes, ranges, finiteness, error wording), restated against the drop-in surface
``mm_b200.pipeline`` / ``mm_b200.chain``.  A maintainer who swaps the module keeps their test suite green; the citations
give the reference test each property comes from.  Values (not just invariants) are pinned elsewhere: tests/test_gpu_chain.py,
test_gpu_stages.py against goldens produced by the reference itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SR = 44100


@pytest.fixture(scope="module")
def P(gpu_lib):
    from mm_b200 import pipeline
    return pipeline


def _tone(freq, amp=0.1, sec=2.0, sr=SR, phase=0.0):
    t = np.linspace(0, sec, int(sr * sec), dtype=np.float32)
    return (amp * np.sin(2 * np.pi * freq * t + phase)).astype(np.float32)


@pytest.fixture(scope="module")
def stereo():
    """test_pipeline.py:13-22: 2 s, 44.1 kHz, a quiet 400 Hz tone, the right channel slightly phase shifted."""
    return np.column_stack((_tone(400), _tone(400, phase=0.1)))


@pytest.fixture(scope="module")
def mono():
    return _tone(440)


def test_dc_offset_is_removed(P):                                    # test_pipeline.py:35-49
    x = (_tone(300, 0.2) + np.float32(0.05))
    assert abs(float(np.mean(P.remove_dc_offset(x)))) < 1e-5
    st = np.column_stack((x, x - np.float32(0.1)))
    out = P.remove_dc_offset(st)
    assert out.shape == st.shape and np.all(np.abs(out.mean(axis=0)) < 1e-5)


def test_measure_lufs_ranges(P, stereo):                             # test_pipeline.py:52-68
    silent = P.measure_lufs(np.zeros((SR * 2, 2), dtype=np.float32), SR)
    assert np.isnan(silent) or silent <= -50.0
    assert -60.0 < P.measure_lufs(stereo, SR) < 0.0
    assert np.isnan(P.measure_lufs(np.zeros(1000, dtype=np.float32), SR))          # shorter than one 400 ms block


def test_spectrum_bars(P, stereo):                                   # test_pipeline.py:71-84
    bars = P.compute_spectrum_bars(stereo, SR)
    assert isinstance(bars, list) and len(bars) == 64 and all(isinstance(b, float) for b in bars)
    short = P.compute_spectrum_bars(np.zeros(100, dtype=np.float32), SR)
    assert len(short) == 64 and all(b <= 0 for b in short)


def test_vectorscope_timeline_correlation(P, stereo, mono):          # test_pipeline.py:87-138
    assert P.compute_vectorscope_points(mono) == []
    pts = P.compute_vectorscope_points(stereo, max_points=500)
    assert 0 < len(pts) <= 500 and all(len(p) == 2 and -1.0 <= p[0] <= 1.0 and -1.0 <= p[1] <= 1.0 for p in pts)
    tl, step = P.compute_lufs_timeline(np.zeros(SR // 2, dtype=np.float32), SR, block_sec=0.4, max_points=300)
    assert isinstance(tl, list) and len(tl) >= 1 and isinstance(step, (int, float))
    tl, step = P.compute_lufs_timeline(stereo.mean(axis=1), SR, block_sec=0.4, max_points=50)
    assert len(tl) >= 1 and step >= 0
    assert P.measure_stereo_correlation(mono) is None
    c = P.measure_stereo_correlation(stereo)
    assert c is not None and -1.0 <= c <= 1.0


def test_export_wav(P, stereo):                                      # test_pipeline.py:141-156, :357-366, :246-256
    for dither in ("tpdf", "ns_e", "ns_itu"):
        wav = P.export_audio(stereo, SR, 2, "wav", dither_type=dither)
        assert isinstance(wav, bytes) and len(wav) > 44 and wav[:4] == b"RIFF" and wav[8:12] == b"WAVE"
        back, sr2 = P.load_audio_from_bytes(wav, "wav")
        assert sr2 == SR and back.shape == stereo.shape and back.dtype == np.float32
        assert np.all(np.isfinite(back)) and np.max(np.abs(back)) <= 1.0
        assert np.max(np.abs(back - stereo)) < 3.0 / 32767.0             # dither + rounding only


def test_run_mastering_pipeline_contract(P, stereo, mono):           # test_pipeline.py:204-243, :301-333, :480-487
    calls = []
    out = P.run_mastering_pipeline(stereo, SR, target_lufs=-14.0, progress_callback=lambda pct, msg: calls.append((pct, msg)))
    assert out.shape == stereo.shape and out.dtype == np.float32 and not np.any(np.isnan(out)) and np.max(np.abs(out)) <= 1.01
    assert len(calls) >= 5 and all(isinstance(p, int) and isinstance(m, str) for p, m in calls)
    for style in ("edm", "dry_vocal", "no_such_style"):
        o = P.run_mastering_pipeline(stereo, SR, target_lufs=-12.0, style=style)
        assert o.shape == stereo.shape and np.all(np.isfinite(o))
    m = P.run_mastering_pipeline(mono, SR, target_lufs=-14.0)
    assert m.shape == mono.shape and np.all(np.isfinite(m)) and np.max(np.abs(m)) <= 1.01
    assert -50.0 < P.measure_lufs(m, SR) < 0.0
    assert abs(float(m[0])) < 1e-6                                        # 6 ms fade-in: the first sample is (near) zero
    # "vocal-like": a few harmonics under a syllable-rate envelope
    t = np.arange(SR * 2) / SR
    voc = (sum(a * np.sin(2 * np.pi * f * t) for f, a in ((220, .2), (440, .12), (880, .06), (3000, .03))) *
           (0.6 + 0.4 * np.sin(2 * np.pi * 3 * t))).astype(np.float32)
    v = P.run_mastering_pipeline(voc, SR, target_lufs=-16.0, style="dry_vocal")
    assert np.all(np.isfinite(v)) and np.max(np.abs(v)) > 1e-3


def test_presets_present(P):                                          # test_pipeline.py:259-281
    for k in ("standard", "edm", "hiphop", "classical", "podcast", "lofi", "house_basic", "dry_vocal"):
        assert k in P.STYLE_CONFIGS and "lufs" in P.STYLE_CONFIGS[k]
    assert set(P.DENOISE_PRESETS) >= {"light", "medium", "aggressive"}
    assert set(P.PRESET_LUFS) >= {"spotify", "youtube", "apple", "club", "broadcast"}


def test_rumble_filter(P):                                            # test_pipeline.py:284-298
    x = (_tone(30, 0.3) + _tone(1000, 0.1)).astype(np.float32)
    out = P.apply_rumble_filter(x, SR, cutoff_hz=80.0)
    assert out.shape == x.shape and np.all(np.isfinite(out))
    lo_in = np.abs(np.fft.rfft(x))[60]                                    # 30 Hz bin of a 2 s signal
    lo_out = np.abs(np.fft.rfft(out.astype(np.float64)))[60]
    assert lo_out < 0.3 * lo_in                                           # the 30 Hz rumble is attenuated ...
    assert abs(np.abs(np.fft.rfft(out.astype(np.float64)))[2000] / np.abs(np.fft.rfft(x))[2000] - 1.0) < 0.05   # ... 1 kHz is not


def test_validate_mastered_not_silent(P, stereo):                     # test_pipeline.py:369-384
    with pytest.raises(ValueError) as e:
        P.validate_mastered_not_silent(np.zeros((1000, 2), dtype=np.float32))
    assert "тишину" in str(e.value) or "Отключите" in str(e.value)
    P.validate_mastered_not_silent(stereo)


def test_pro_modules_not_silent(P, stereo):                           # test_pipeline.py:387-421
    a = P.apply_dynamic_eq(stereo, SR)
    assert a.shape == stereo.shape and a.dtype == np.float32 and np.all(np.isfinite(a)) and np.max(np.abs(a)) > 1e-4
    a = P.apply_transient_designer(a, SR, attack_gain=1.3, sustain_gain=0.9)
    a = P.apply_parallel_compression(a, SR, mix=0.25)
    assert a.shape == stereo.shape and np.all(np.isfinite(a)) and np.max(np.abs(a)) > 1e-4


def test_high_freq_trim_cuts_highs(P):                                # test_pipeline.py:424-436
    hi = _tone(10000, 0.2)
    out = P.apply_high_freq_trim(hi, SR, crossover_hz=5000.0, high_gain=0.9)
    assert out.shape == hi.shape
    assert 0.8 * np.max(np.abs(hi)) <= np.max(np.abs(out)) <= 0.95 * np.max(np.abs(hi))
    lo = _tone(200, 0.2)
    assert abs(np.max(np.abs(P.apply_high_freq_trim(lo, SR))) / np.max(np.abs(lo)) - 1.0) < 0.01


def test_output_edge_fade_in(P, stereo):                              # test_pipeline.py:439-456
    sr = 48000
    x = np.ones(sr, dtype=np.float32) * 0.5
    out = P.apply_output_edge_fade_in(x, sr, fade_ms=5.0)
    assert out[0] < x[0] and np.isclose(out[-1], x[-1], rtol=1e-4)
    n_fade = int(round(sr * 0.005))
    assert np.all(out[n_fade:n_fade + 100] == x[n_fade:n_fade + 100])
    st = P.apply_output_edge_fade_in(stereo, SR, fade_ms=4.0)
    assert st.shape == stereo.shape and np.isclose(st[0, 0], 0.0, atol=1e-6) and np.isclose(st[0, 1], 0.0, atol=1e-6)


def test_v2_chain_on_seeded_noise_is_finite_with_moderate_highs(gpu_lib):   # test_mastering_regression_windows.py:30-49
    """48 s of seeded noise through the v2 default chain: finite, no high-frequency blow-up (> 8 kHz RMS ratio < 80 in
    every window), no sample-to-sample jump above 1.5."""
    from scipy import signal as sg
    from mm_b200 import pipeline as P
    from mm_b200.chain import MasteringChain
    sr = 48000
    x = (0.04 * np.random.default_rng(42).standard_normal(int(sr * 48.0))).astype(np.float32)
    out = MasteringChain.default_chain(target_lufs=-14.0, style="standard").process(x, sr, target_lufs=-14.0, style="standard")
    out = P.apply_output_edge_fade_in(out, sr, fade_ms=6.0)
    assert out.shape == x.shape and np.all(np.isfinite(out))
    b, a = sg.butter(4, 8000 / (sr / 2), "high")
    for w0 in (0.0, 20.0, 40.0):
        s = slice(int(w0 * sr), int((w0 + 8.0) * sr))
        hf_in = np.sqrt(np.mean(sg.lfilter(b, a, x[s].astype(np.float64)) ** 2))
        hf_out = np.sqrt(np.mean(sg.lfilter(b, a, out[s].astype(np.float64)) ** 2))
        assert hf_out / (hf_in + 1e-12) < 80.0
        assert np.max(np.abs(np.diff(out[s].astype(np.float64)))) < 1.5


def test_mastering_trace_lines(P, stereo, monkeypatch, caplog):              # test_mastering_trace.py:17-52
    """With MAGIC_MASTER_MASTERING_TRACE on, every stage logs a `mastering_trace` line carrying the job id, the stage name
    and the device-computed metrics; without it nothing is logged and the fused path runs."""
    import logging
    from mm_b200.chain import MasteringChain
    from mm_b200.mastering_trace import TraceContext, batch_metrics
    ctx = TraceContext(job_id="job-42", filename="../some dir/My Song (final).wav", path="v1")
    caplog.set_level(logging.INFO, logger="magic_master.mastering_trace")
    monkeypatch.delenv("MAGIC_MASTER_MASTERING_TRACE", raising=False)
    quiet = P.run_mastering_pipeline(stereo, SR, trace_ctx=ctx)
    assert not [r for r in caplog.records if "mastering_trace" in r.getMessage()]
    monkeypatch.setenv("MAGIC_MASTER_MASTERING_TRACE", "1")
    monkeypatch.setenv("MAGIC_MASTER_MASTERING_TRACE_LUFS_STAGES", "1")
    traced = P.run_mastering_pipeline(stereo, SR, trace_ctx=ctx)
    lines = [r.getMessage() for r in caplog.records if "mastering_trace" in r.getMessage()]
    stages = [ln.split("stage=")[1].split()[0] for ln in lines]
    assert stages[:3] == ["dc_offset", "peak_guard_in", "target_eq"] and stages[-1] == "finalize_clip" and "normalize_lufs" in stages
    assert all("job_id=job-42" in ln and "filename=My_Song_final_.wav" in ln and "peak_db=" in ln and "nan_count=0" in ln for ln in lines)
    assert any("lufs=" in ln for ln in lines)
    assert np.max(np.abs(traced.astype(np.float64) - quiet)) <= 2e-6        # stage-by-stage == fused (float32 hand-offs aside)
    caplog.clear()
    MasteringChain.default_chain(target_lufs=-14.0, style="standard").process(stereo, SR, target_lufs=-14.0, style="standard",
                                                                            trace_ctx=TraceContext("job-43", "a.wav", "v2"))
    v2 = [r.getMessage().split("stage=")[1].split()[0] for r in caplog.records if "mastering_trace" in r.getMessage()]
    assert v2[0] == "dc_offset" and v2[-1] == "chain_finalize_clip" and "dynamics" in v2
    # the metrics kernel itself: NaN / Inf are counted, the peak ignores them
    bad = stereo.copy()
    bad[10, 0], bad[11, 1], bad[12, 0] = np.nan, np.inf, -np.inf
    eng, b, _ = P._up(bad, SR)
    m = batch_metrics(eng, b)[0]
    assert m["nan_count"] == 1 and m["inf_count"] == 2 and abs(m["peak_linear"] - round(float(np.max(np.abs(stereo))), 6)) < 2e-6


def test_error_conventions_at_the_boundary(P, stereo):
    """SURVEY 8b "error conventions": what degrades quietly in the reference and what fails loudly here.
    * measure_lufs -> NaN and normalize_lufs -> input unchanged for clips shorter than one 400 ms block (pipeline.py:647-664);
    * a missing GPU / library or an option without a kernel raises instead of passing audio through;
    * the C ABI rejects bad geometry with a message instead of reading out of bounds."""
    import ctypes as C
    from mm_b200 import _lib
    from mm_b200.engine import get_engine
    short = stereo[:4000]
    assert np.isnan(P.measure_lufs(short, SR))
    assert np.array_equal(P.normalize_lufs(short, SR, -14.0), short)
    with pytest.raises(NotImplementedError):
        P.export_audio(stereo, SR, 2, "mp3")
    with pytest.raises(ValueError):                       # scipy.signal.stft's own complaint for less than one segment
        P.apply_spectral_denoise(stereo[:1000], SR, 0.5)
    eng = get_engine()
    b = eng.upload([stereo], SR)
    for bad in (_lib.Geom(b.n, b.stride, 1, 3, SR, 0), _lib.Geom(0, b.stride, 1, 2, SR, 0), _lib.Geom(b.n, b.n, 1, 2, SR, 0),
                _lib.Geom(b.n, b.stride, 0, 2, SR, 0)):
        rc = eng.lib.mm_dev_remove_dc_offset(eng.ctx, C.byref(bad), b.ptr, b.ptr)
        assert rc != 0 and _lib.last_error()
    tiny = eng.upload([stereo[:8]], SR)                  # shorter than filtfilt's padlen: degrades to lfilter like the reference
    g = tiny.geom                                        # (values: test_inputs_not_longer_than_padlen_degrade_to_lfilter)
    tout = eng.like(tiny)
    assert eng.lib.mm_dev_apply_target_curve(eng.ctx, C.byref(g), tiny.ptr, tout.ptr, 0) == 0
    # the de-esser on such an input: the reference's own np.convolve smoothing raises (67-tap box against 8 samples); refused by name
    assert eng.lib.mm_dev_apply_deesser(eng.ctx, C.byref(g), tiny.ptr, tout.ptr, -6.0, 3.0, 5000.0, 9000.0, 4.0, 85.0) != 0
    assert "padlen" in _lib.last_error()
    assert eng.lib.mm_dev_master(eng.ctx, C.byref(g), 7, None, tiny.ptr, tiny.ptr, None, None, 0, None, 0) != 0
