"""FFT-class stages through the C ABI against the golden vectors of the unmodified reference
(tests/golden/make_golden_fft.py) and against the oracle on larger seeded inputs."""
import numpy as np
import pytest

from conftest import fft_stage_cases, load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-4          # BASELINE.json north_star: float32 samples within 1e-4 absolute
FFT_TOL = 1e-5      # what the float32 FFT paths are expected to hold (measured ~1e-6)


@pytest.fixture(scope="module")
def P(gpu_lib):
    from mm_b200 import pipeline
    return pipeline


def _err(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def _material(n, sr, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    tone = 0.3 * np.sin(2 * np.pi * 330.0 * t) + 0.1 * np.sin(2 * np.pi * 2500.0 * t + 1.0)
    gate = ((np.arange(n) % 30000) < 17000).astype(np.float64)
    x = np.stack([tone * gate, 0.7 * tone * gate], axis=1) + 0.008 * rng.standard_normal((n, 2))
    return x.astype(np.float32)


def test_spectral_denoise_against_reference_golden(P):
    g = load_golden("fft_stages")
    for k, call in fft_stage_cases(P, g).items():
        if not k.startswith("denoise"):
            continue
        out = call()
        e = _err(out, g[k])
        print(f"[parity] {k}: {e:.3e}")
        assert out.shape == g[k].shape and out.dtype == np.float32 and e <= FFT_TOL, (k, e)


def test_spectral_denoise_edge_behaviour(P):
    x = load_golden("fft_stages")["input"]
    assert P.apply_spectral_denoise(x, 48000, strength=0.005) is x                 # bypass returns the same object
    with pytest.raises(ValueError):
        P.apply_spectral_denoise(x[:1500], 48000, strength=0.5)                    # scipy: noverlap must be less than nperseg
    with pytest.raises(ValueError):
        P.apply_spectral_denoise(x, 48000, strength=0.5, noise_percentile=120.0)
    assert set(P.DENOISE_PRESETS) == {"vocal", "light", "medium", "aggressive", "tape_hiss", "room_tone"}


def test_spectral_denoise_long_against_oracle(P):
    """10 s stereo (940 frames per bin: the radix select walks real histograms) and a 44.1 kHz odd length."""
    from oracle import chain as oc
    for sr, n, st, pc in ((48000, 480000, 0.5, 15.0), (44100, 131071, 0.75, 10.0), (44100, 2048, 0.4, 50.0), (44100, 2049, 1.0, 0.0),
                          (48000, 40000, 0.6, 100.0)):
        x = _material(n, sr, n)
        out = P.apply_spectral_denoise(x, sr, strength=st, noise_percentile=pc)
        ref = oc.apply_spectral_denoise(x, sr, st, pc)
        e = _err(out, ref)
        print(f"[parity] denoise n={n} sr={sr} s={st} p={pc}: {e:.3e}")
        assert e <= FFT_TOL, (n, e)


def test_spectral_denoise_reconstructs_when_floor_is_negligible(P):
    """Size-independent property at the bench length: with a digital-silence noise floor (the tones stop for most of the
    track, percentile 0 picks the silent frames) the gain is 1 everywhere and STFT -> ISTFT must return the input."""
    sr, n = 44100, 180 * 44100
    t = np.arange(n, dtype=np.float64) / sr
    x = (0.4 * np.sin(2 * np.pi * 441.0 * t) * (t > 120.0)).astype(np.float32)
    out = P.apply_spectral_denoise(x, sr, strength=1.0, noise_percentile=0.0)
    e = _err(out, x)
    print(f"[parity] denoise identity, 180 s: {e:.3e}")
    assert out.shape == x.shape and e <= 2e-6


def test_v1_chain_with_denoise_against_oracle(P):
    """run_mastering_pipeline(denoise_strength=...) (pipeline.py:1841-1844): stage-by-stage path == oracle stages in the
    reference's order."""
    from oracle import chain as oc
    sr, n = 48000, 96000
    x = _material(n, sr, 5)
    out = P.run_mastering_pipeline(x, sr, target_lufs=-14.0, style="standard", denoise_strength=0.5)
    ref = oc.run_v1(x, sr, -14.0, "standard", denoise_strength=0.5)
    e = _err(out, ref)
    print(f"[parity] v1 chain with denoise: {e:.3e}")
    assert out.shape == x.shape and e <= TOL


RS_TOL = 2e-5       # float32 Bluestein (four 2^18..2^25-point FFTs per row); measured ~1e-6


def test_fft_resample_against_reference_golden(P):
    g = load_golden("fft_stages")
    for k, call in fft_stage_cases(P, g).items():
        if not k.startswith("resample"):
            continue
        out = call()
        e = _err(out, g[k])
        print(f"[parity] {k}: {e:.3e}")
        assert out.shape == g[k].shape and out.dtype == np.float32 and e <= RS_TOL, (k, e)


def test_oversampled_exciter_against_reference_golden(P):
    g = load_golden("fft_stages")
    for k, call in fft_stage_cases(P, g).items():
        if not k.startswith("exciter"):
            continue
        out = call()
        e = _err(out, g[k])
        print(f"[parity] {k}: {e:.3e}")
        assert out.shape == g[k].shape and out.dtype == np.float32 and e <= RS_TOL, (k, e)


def test_fft_resample_lengths_against_oracle(P):
    """Odd / even / prime lengths either way (the unpaired-bin rule of scipy.signal.resample), identity, errors."""
    from oracle import chain as oc
    rng = np.random.default_rng(3)
    for n, num in ((1000, 1500), (1001, 1500), (1000, 1501), (1500, 1000), (1501, 1000), (1500, 1001), (7919, 7907), (2, 3),
                   (3, 2), (1, 5), (5, 1), (262144, 262145), (300000, 100000)):
        x = (0.5 * rng.standard_normal(n)).astype(np.float32)
        out = P.fft_resample(x, num)
        ref = oc.fft_resample(x, num).astype(np.float32)
        e = _err(out, ref)
        print(f"[parity] resample {n} -> {num}: {e:.3e}")
        assert out.shape == (num,) and e <= RS_TOL * max(1.0, float(np.max(np.abs(ref)))), (n, num, e)
    x = (0.1 * rng.standard_normal((500, 2))).astype(np.float32)
    assert np.array_equal(P.resample_audio(x, 48000, 48000), x)
    with pytest.raises(ValueError):
        P.resample_audio(x, 0, 48000)


def test_fft_resample_three_minute_track(P):
    """BASELINE size: one 180 s stereo track 44.1 -> 48 kHz (7,938,000 -> 8,640,000 frames, two 2^24-point chirp
    convolutions per row) against scipy on the host, and back again (band-limited material survives the round trip)."""
    from oracle import chain as oc
    sr, n = 44100, 180 * 44100
    t = np.arange(n, dtype=np.float64) / sr
    rng = np.random.default_rng(11)
    x = np.stack([0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.2 * np.sin(2 * np.pi * 9000.0 * t * (1 + 1e-3 * t)),
                  0.25 * np.sin(2 * np.pi * 55.0 * t) + 0.02 * rng.standard_normal(n)], axis=1).astype(np.float32)
    up = P.resample_audio(x, 44100, 48000)
    ref = oc.resample_audio(x, 44100, 48000)
    e = _err(up, ref)
    print(f"[parity] resample 180 s 44.1k -> 48k: {e:.3e}")
    assert up.shape == ref.shape and e <= RS_TOL
    back = P.resample_audio(up, 48000, 44100)
    e2 = _err(back, x)
    print(f"[parity] resample round trip 180 s: {e2:.3e}")
    assert back.shape == x.shape and e2 <= 5e-5


def test_reference_match_with_reference_at_another_rate(P):
    """apply_reference_match, ref_sr != sr (pipeline.py:1581-1584): mono mix, truncating length, FFT resample."""
    from oracle import chain as oc
    g = load_golden("pro_stages_48k")
    sr = int(g["sr"])
    loud = (g["input"] * np.float32(6.0)).astype(np.float32)
    ref441 = oc.resample_audio(g["refmatch_reference"], 48000, 44100)
    out = P.apply_reference_match(loud, sr, ref441, 44100, strength=0.8)
    ref_mono = np.mean(ref441, axis=1)
    ref48 = oc.fft_resample(ref_mono.astype(np.float64), int(len(ref_mono) * sr / 44100)).astype(np.float32)
    want = oc.apply_reference_match(loud, sr, ref48, sr, 0.8)
    e = _err(out, want)
    print(f"[parity] reference match, reference at 44.1 kHz: {e:.3e}")
    assert e <= 2e-5


def test_chain_modules_oversampled_exciter_and_multiband_imager(P):
    """v2 chain with ExciterModule(oversample=2) and ImagerModule(band_widths=...) (modules/exciter.py, modules/imaging.py)
    == the same stage functions applied in order."""
    from mm_b200.chain import MasteringChain
    g = load_golden("fft_stages")
    x, sr = (g["input"] * np.float32(2.0)).astype(np.float32), int(g["sr"])
    cfg = {"modules": [{"id": "exciter", "enabled": True, "exciter_db": 2.0, "mode": "tape", "oversample": 2},
                       {"id": "imager", "enabled": True, "width": 1.0, "band_widths": [0.8, 1.0, 1.3, 1.6]}]}
    out = MasteringChain.from_config(cfg).process(x, sr)
    want = P.apply_stereo_imager(P.apply_harmonic_exciter(x, sr, 2.0, "tape", 2), 1.0, sr=sr, band_widths=[0.8, 1.0, 1.3, 1.6])
    want = np.nan_to_num(np.clip(want, -1.0, 1.0))
    e = _err(out, want)
    print(f"[parity] chain exciter(os=2) + imager(4 band): {e:.3e}")
    assert e <= 1e-6
    assert _err(P.apply_harmonic_exciter(x, sr, 2.0, "tape", 2), g["exciter_os2"]) <= RS_TOL


def test_export_auto_blank_and_pcm24(P):
    """export_audio(auto_blank_sec=...) cuts where the reference cuts (golden WAV sizes), the kept int16 frames equal the
    uncut export's under the same dither buffer, and the PCM_24 hand-off equals lrintf(x * 0x7FFFFF)."""
    from oracle import chain as oc
    g = load_golden("fft_stages")
    tail, x, sr = g["blank_input"], g["input"], int(g["sr"])
    for name, sig, sec in (("blank_len_03", tail, 0.3), ("blank_len_mono_01", np.ascontiguousarray(tail[:, 0]), 0.1),
                           ("blank_len_none", x, 0.2), ("blank_len_all_quiet", tail[17100:], 0.05)):
        ch = 1 if sig.ndim == 1 else sig.shape[1]
        wav = P.export_audio(sig, sr, ch, "wav", auto_blank_sec=sec)
        assert (len(wav) - 44) // (2 * ch) == int(g[name]), name
    keep = int(g["blank_len_mono_01"])
    mono = np.ascontiguousarray(tail[:, 0])
    noise = (np.random.default_rng(5).random((len(mono), 1)) + np.random.default_rng(6).random((len(mono), 1)) - 1.0).astype(np.float32)
    cut = P.export_audio(mono, sr, 1, "wav", auto_blank_sec=0.1, noise=noise[:keep])
    full = P.export_audio(mono, sr, 1, "wav", noise=noise)
    assert cut[44:] == full[44:44 + 2 * keep]
    loud = np.concatenate([x * np.float32(40.0), np.array([[np.nan, 1.0], [-1.0, 0.5]], dtype=np.float32)])
    pcm = P.export_pcm24(loud, sr)
    assert pcm.dtype == np.int32 and np.array_equal(pcm, oc.quantize_pcm24(loud))
    assert P.export_pcm24(tail, sr, auto_blank_sec=0.1).shape[0] == keep


def test_dynamic_eq_stable_bands_against_reference_golden(P):
    """apply_dynamic_eq (pipeline.py:1628-1700): float64 zero-phase peaking sections + the float32 follower / gain arithmetic
    of numpy; an unstable band (all of the reference's defaults) is refused by name."""
    from test_oracle_golden import dyneq_cases
    from mm_b200._lib import MMError
    g = load_golden("fft_stages")
    for k, call in dyneq_cases(P, g).items():
        out = call()
        e = _err(out, g[k])
        print(f"[parity] {k}: {e:.3e}")
        assert out.shape == g[k].shape and out.dtype == np.float32 and e <= 1e-5, (k, e)   # |x| up to 3.8: a few float32 ulps over three bands
    assert len(P.DYNAMIC_EQ_MASTERING_BANDS) == 8


def test_dynamic_eq_default_bands_against_reference_golden(P):
    """apply_dynamic_eq(audio, sr) with the reference's DEFAULT bands (pipeline.py:1616-1625): unstable sections the reference
    zeroes (identity), q = 1 sections its lfilter fallback turns into ``x * g``, the 12 kHz band at 48 kHz that filtfilt
    returns as ``x - const`` -- against outputs of the unmodified reference at 44.1 / 48 / 96 kHz, 1 to 20 s."""
    from conftest import dyneq_default_cases
    from mm_b200._lib import MMError
    g = load_golden("dyneq_default")
    for name, x, sr, dec in dyneq_default_cases():
        rep = []
        out = P.apply_dynamic_eq(x, sr, report=rep)
        e = _err(out[::dec], g[name])
        print(f"[parity] dynamic eq defaults {name}: {e:.3e}  {rep}")
        assert out.shape == x.shape and out.dtype == np.float32 and e <= 2e-6, (name, e)
        assert "skipped" not in rep and rep.count("overflow->identity") >= 4
    # an unstable band on a signal too short for its overflow to be certain: passed through (warning), refused under strict
    x, sr, _ = next((x, sr, d) for n_, x, sr, d in dyneq_default_cases() if n_ == "d44_2s")
    short = np.ascontiguousarray(x[:600])
    rep = []
    out = P.apply_dynamic_eq(short, sr, [P.DYNAMIC_EQ_MASTERING_BANDS[1]], report=rep)
    assert rep == ["skipped"] and np.array_equal(out, np.clip(short, -1, 1))
    with pytest.raises(MMError, match="unstable"):
        P.apply_dynamic_eq(short, sr, [P.DYNAMIC_EQ_MASTERING_BANDS[1]], strict=True)


def test_dynamic_eq_long_against_oracle(P):
    from oracle import chain as oc
    sr, n = 44100, 20 * 44100
    x = (_material(n, sr, 21) * np.float32(1.5)).astype(np.float32)
    bands = [{"freq": 150, "q": 0.4, "threshold_db": -30, "ratio": 3.0, "attack_ms": 10, "release_ms": 120, "max_cut_db": -6},
             {"freq": 4000, "q": 0.6, "threshold_db": -36, "ratio": 4.0, "attack_ms": 2, "release_ms": 40, "max_cut_db": -9}]
    out = P.apply_dynamic_eq(x, sr, bands)
    ref = oc.apply_dynamic_eq(x, sr, bands)
    e = _err(out, ref)
    print(f"[parity] dynamic eq 20 s: {e:.3e}")
    assert e <= 5e-6 and _err(ref, np.clip(x, -1, 1)) > 1e-2


def test_fft_stages_from_concurrent_threads(P):
    """SURVEY 8b threading: the job functions run in up to three asyncio.to_thread workers.  Every thread has its own engine
    (stream, workspace, FFT plan cache); concurrent calls return bit for bit what a lone call returns."""
    import threading
    x = load_golden("fft_stages")["input"]
    sr = 48000
    jobs = [lambda: P.resample_audio(x, 48000, 44100),
            lambda: P.apply_spectral_denoise(x, sr, 0.5, 15.0),
            lambda: P.apply_harmonic_exciter(x * np.float32(2.0), sr, 2.0, "tape", 2)]
    alone = [j() for j in jobs]
    got = [[None] * 4 for _ in jobs]
    errs = []

    def work(i):
        try:
            for rep in range(4):
                got[i][rep] = jobs[i]()
        except Exception as e:      # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(len(jobs)):
        for rep in range(4):
            assert np.array_equal(got[i][rep], alone[i]), (i, rep)


def test_long_fir_as_fft_convolution_against_oracle(P):
    """Rows long enough for the FFT-convolution path (n >= 4 K): the 4096-tap linear-phase target curve on a 60 s stereo and
    a mono track, and an 8192-tap reference match, against the oracle's scipy.signal.fftconvolve."""
    from oracle import chain as oc
    sr, n = 44100, 60 * 44100
    x = (_material(n, sr, 31) * np.float32(1.5)).astype(np.float32)
    for name, sig in (("stereo", x), ("mono", np.ascontiguousarray(x[:, 1]))):
        out = P.apply_target_curve(sig, sr, phase_mode="linear_phase")
        ref = oc.apply_target_curve_linear_phase(sig, sr)
        e = _err(out, ref)
        print(f"[parity] linear phase 60 s {name} (FFT convolution): {e:.3e}")
        assert out.shape == ref.shape and e <= FFT_TOL
    from scipy import signal as sg
    bb, aa = sg.butter(1, 2500 / (sr / 2), "high")
    refsig = (x + 0.7 * sg.lfilter(bb, aa, x, axis=0)).astype(np.float32)
    out = P.apply_reference_match(x, sr, refsig, sr, strength=0.8)
    ref = oc.apply_reference_match(x, sr, refsig, sr, 0.8)
    e = _err(out, ref)
    print(f"[parity] reference match 60 s (8192 taps, FFT convolution): {e:.3e}")
    assert e <= 2e-5


def test_big_fft_row_chunking(P):
    """More rows than one 4 GB work area holds (2048 rows of the smallest, 2^18-point transform): FFT resampling of 4200 rows and
    FFT convolution of 2100 stereo tracks are processed in sub-batches; every track must equal its lone result."""
    import ctypes as C
    from mm_b200.engine import get_engine
    from oracle import chain as oc
    eng = get_engine()
    rng = np.random.default_rng(17)
    tracks, n = 2100, 4200
    base = (0.2 * rng.standard_normal((8, n, 2))).astype(np.float32)
    batch = [base[t % 8] * np.float32(1.0 + 0.001 * (t // 8)) for t in range(tracks)]
    b = eng.upload(batch, 48000)
    up = eng.download(eng.fft_resample(b, 6300, 72000))
    for t in (0, 1, 2047, 2048, 2099):
        ref = oc.fft_resample(batch[t][:, 0], 6300).astype(np.float32)
        assert _err(up[t][:, 0], ref) <= RS_TOL, t
    taps = np.ascontiguousarray((np.hanning(1024) * rng.standard_normal(1024) / 64.0).astype(np.float32))
    out = eng.download(eng.stage("fir_same", b, taps.ctypes.data_as(C.c_void_p), 1024, 0))
    from scipy import signal as sg
    for t in (0, 1023, 1024, 2047, 2048, 2099):
        for ch in (0, 1):
            ref = sg.fftconvolve(batch[t][:, ch].astype(np.float64), taps.astype(np.float64), mode="same")
            assert _err(out[t][:, ch], ref) <= RS_TOL, (t, ch)


def test_fft_resample_four_pass_transforms(P):
    """Transforms above 2^24 points run as four passes (64 / 128-point axes): x2 of a 180 s mono track needs 2^25, x4 needs 2^26
    (what the oversampled exciter does at the bench length).  Against scipy on the host."""
    from oracle import chain as oc
    sr, n = 44100, 170 * 44100
    t = np.arange(n, dtype=np.float64) / sr
    rng = np.random.default_rng(23)
    x = (0.3 * np.sin(2 * np.pi * 523.25 * t) + 0.15 * np.sin(2 * np.pi * 11000.0 * t) + 0.02 * rng.standard_normal(n)).astype(np.float32)
    for os_ in (2, 4):
        up = P.fft_resample(x, n * os_)
        ref = oc.fft_resample(x, n * os_).astype(np.float32)
        e = _err(up, ref)
        print(f"[parity] resample 170 s x{os_} (2^{25 if os_ == 2 else 26}-point chirp convolution): {e:.3e}")
        assert up.shape == ref.shape and e <= RS_TOL
        back = P.fft_resample(up, n)
        e2 = _err(back, x)
        print(f"[parity] resample x{os_} and back: {e2:.3e}")
        assert e2 <= 5e-5


def test_multiband_only_and_lookahead_maximizer(P):
    """The last two public stage functions of pipeline.py: apply_multiband_dynamics on its own (:414-481) and
    apply_maximizer_lookahead (:548-573), against the reference's outputs."""
    from test_oracle_golden import tail_stage_cases
    g = load_golden("fft_stages")
    for k, call in tail_stage_cases(P, g).items():
        out = call()
        e = _err(out, g[k])
        print(f"[parity] {k}: {e:.3e}")
        assert out.shape == g[k].shape and out.dtype == np.float32 and e <= 2e-6, (k, e)
    assert (P.MAXIMIZER_THRESHOLD_DB, P.MAXIMIZER_MARGIN_DB, P.FINAL_TRIM_DB) == (-2.5, -0.3, 0.5)


def test_second_wave_stages_batch_equals_single_track(P):
    """Tracks are independent: a batch of different tracks through the STFT denoiser, the dynamic EQ, the spectral envelope and
    the lookahead maximizer returns, bit for bit, what each track returns alone."""
    import ctypes as C
    import torch
    from mm_b200 import _lib
    from mm_b200.engine import get_engine
    eng = get_engine()
    sr, n = 48000, 50000
    tracks = [_material(n, sr, 100 + t) * np.float32(0.5 + 0.4 * t) for t in range(3)]
    b = eng.upload(tracks, sr)
    bands = [{"freq": 3000, "q": 0.5, "threshold_db": -30, "ratio": 3.0, "attack_ms": 5, "release_ms": 60, "max_cut_db": -6}]
    row = _lib.darr([3000 / 24000, (3000 / 24000) / 0.5, -30, 3.0, 5, 60, -6])
    outs = {
        "denoise": eng.download(eng.stage("apply_spectral_denoise", b, C.c_double(0.6), C.c_double(20.0))),
        "dyneq": eng.download(eng.stage("apply_dynamic_eq", b, 1, row)),
        "lookahead": eng.download(eng.stage("apply_maximizer_lookahead", b, C.c_double(6.0))),
    }
    with torch.cuda.stream(eng.stream):
        env = torch.empty(3 * 4097, dtype=torch.float32, device=eng.tdev)
        g = b.geom
        _lib.check(eng.lib.mm_dev_spectral_envelope(eng.ctx, C.byref(g), b.ptr, C.c_void_p(env.data_ptr())))
        eng.sync()
        env = env.cpu().numpy().reshape(3, 4097)
    for t in range(3):
        assert np.array_equal(outs["denoise"][t], P.apply_spectral_denoise(tracks[t], sr, 0.6, 20.0)), t
        assert np.array_equal(outs["dyneq"][t], P.apply_dynamic_eq(tracks[t], sr, bands)), t
        assert np.array_equal(outs["lookahead"][t], P.apply_maximizer_lookahead(tracks[t], sr, 6.0)), t
        assert np.array_equal(env[t], P.compute_spectral_envelope(tracks[t], sr)), t


def test_release_workspace(P):
    """mm_ctx_release_workspace: scratch returns to zero and the next call simply allocates again, same result."""
    from mm_b200.engine import get_engine
    x = load_golden("fft_stages")["input"]
    a = P.resample_audio(x, 48000, 44100)
    eng = get_engine()
    assert eng.lib.mm_ctx_workspace_bytes(eng.ctx) > 0
    eng.release_workspace()
    assert eng.lib.mm_ctx_workspace_bytes(eng.ctx) == 0
    assert np.array_equal(P.resample_audio(x, 48000, 44100), a)
    assert np.array_equal(P.apply_spectral_denoise(x, 48000, 0.5), P.apply_spectral_denoise(x, 48000, 0.5))


def test_fused_true_peak_correlation_equals_separate_kernels(P):
    """mm_dev_true_peak_correlation (one pass) == mm_dev_true_peak + mm_dev_stereo_correlation: true peak and sample peak bit for
    bit, correlation to float64 summation order; mono batches take the separate kernels."""
    from mm_b200.engine import get_engine
    eng = get_engine()
    sr = 44100
    for n in (30 * 44100, 16384 * 3 + 5, 1000):
        tracks = [_material(n, sr, 40 + t) * np.float32(0.7 + 0.3 * t) for t in range(3)]
        b = eng.upload(tracks, sr)
        tp, corr, peak = eng.true_peak_correlation(b)
        tp0 = eng.true_peak(b)
        corr0, peak0 = eng.stereo_correlation(b)
        assert np.array_equal(tp, tp0) and np.array_equal(peak, peak0), n
        assert np.max(np.abs(corr - corr0)) <= 1e-12, (n, corr, corr0)
    mono = eng.upload([_material(5000, sr, 3)[:, 0]], sr)
    tp, corr, peak = eng.true_peak_correlation(mono)
    assert np.array_equal(tp, eng.true_peak(mono)) and np.isnan(corr[0])
    rec = P.analyze_batch([_material(60000, sr, 9)], sr)[0]
    assert set(rec) >= {"lufs", "true_peak_dbfs", "sample_peak", "correlation", "spectrum_bars"}


def test_chirp_plan_cache_eviction(P):
    """More (n, num) pairs than the plan cache keeps (6 chirp filters = 3 pairs): old plans are evicted least-recently-used
    first and every result still matches scipy -- including a pair whose forward plan is the oldest entry when its inverse
    plan is created."""
    from oracle import chain as oc
    rng = np.random.default_rng(29)
    x = (0.3 * rng.standard_normal(5000)).astype(np.float32)
    pairs = [(5000, 5100), (5000, 5200), (5000, 5300), (5000, 5400), (5000, 5100), (5000, 5500), (5000, 5200)]
    for n, num in pairs:
        out = P.fft_resample(x[:n], num)
        assert _err(out, oc.fft_resample(x[:n], num).astype(np.float32)) <= RS_TOL, (n, num)
