"""Analyzer kernels (LUFS, true peak, spectrum bars, correlation) and the dithered export."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(gpu_lib):
    from mm_b200 import pipeline
    return pipeline


def test_analyzers_match_reference_golden(P):
    g = load_golden("analyzers")
    for tag in "abc":
        x, sr = g[f"{tag}_input"], int(g[f"{tag}_sr"])
        lufs = P.measure_lufs(x, sr)
        lufs_m = P.measure_lufs(np.ascontiguousarray(x[:, 0]), sr)
        tp = P.true_peak_dbfs(x, sr)
        corr = P.measure_stereo_correlation(x)
        print(f"[parity] analyzers {tag}: lufs {lufs - float(g[f'{tag}_lufs']):+.2e} mono {lufs_m - float(g[f'{tag}_lufs_mono']):+.2e} "
              f"tp {tp - float(g[f'{tag}_true_peak']):+.2e} corr {corr - float(g[f'{tag}_corr']):+.2e}")
        assert abs(lufs - float(g[f"{tag}_lufs"])) <= 0.01
        assert abs(lufs_m - float(g[f"{tag}_lufs_mono"])) <= 0.01
        assert abs(tp - float(g[f"{tag}_true_peak"])) <= 0.01
        assert abs(corr - float(g[f"{tag}_corr"])) <= 1e-6
        bars = np.array(P.compute_spectrum_bars(x, sr))
        ref = g[f"{tag}_bars"]
        # the reference rounds to 0.01 dB and its float32 FFT has its own noise floor: compare
        # magnitudes (linear) with an absolute floor, plus dB where the bar is well above it
        lin, lin_ref = 10 ** (bars / 20), 10 ** (ref / 20)
        assert np.max(np.abs(lin - lin_ref) - 0.003 * lin_ref) <= 2e-7
        loud = ref > -90
        assert np.max(np.abs(bars[loud] - ref[loud])) <= 0.03


def test_lufs_edge_cases(P):
    sr = 48000
    assert np.isnan(P.measure_lufs(np.zeros((1000, 2), np.float32), sr))        # shorter than one block
    v = P.measure_lufs(np.zeros((sr, 2), np.float32), sr)                        # silence: -inf (tests accept <= -50)
    assert np.isnan(v) or v <= -50
    # BS.1770 known answer: 997 Hz full-scale sine in both channels reads 0.0 LKFS (-3.01 in one);
    # pyloudnorm's RBJ-form K-weighting is ~0.04 dB off the standard's table at 997 Hz, the oracle
    # restates pyloudnorm, and the kernel must follow the oracle closely
    from oracle import chain as oc
    t = np.arange(sr * 3) / sr
    s = np.sin(2 * np.pi * 997 * t).astype(np.float32)
    both, one = np.stack([s, s], axis=1), np.stack([s, np.zeros_like(s)], axis=1)
    assert abs(P.measure_lufs(both, sr) - 0.0) <= 0.06
    assert abs(P.measure_lufs(one, sr) - (-3.01)) <= 0.06
    assert abs(P.measure_lufs(both, sr) - oc.measure_lufs(both, sr)) <= 1e-4
    assert abs(P.measure_lufs(one, sr) - oc.measure_lufs(one, sr)) <= 1e-4


def test_lufs_many_blocks_vs_oracle(P):
    from oracle import chain as oc
    from mm_b200 import synth
    for sr in (44100, 48000, 96000, 22050):
        x = synth.numpy_track(3, sr, 7.3)
        assert abs(P.measure_lufs(x, sr) - oc.measure_lufs(x, sr)) <= 1e-3, sr


def test_lufs_long_signal_cluster_gate_vs_oracle(P):
    """A signal with more than 8192 gating blocks takes the thread-block-cluster gate (gate_long_kernel: 8 CTAs, partial sums
    met through distributed shared memory); 15 minutes with loud, quiet (relatively gated) and silent (absolutely gated) parts."""
    from oracle import chain as oc
    sr, sec = 22050, 900
    rng = np.random.default_rng(7)
    t = np.arange(sr * sec) / sr
    env = np.where((t % 60) < 35, 0.25, np.where((t % 60) < 50, 0.004, 0.0))
    x = (env[:, None] * rng.standard_normal((sr * sec, 2))).astype(np.float32)
    x[:, 1] *= 0.5
    a, b = P.measure_lufs(x, sr), oc.measure_lufs(x, sr)
    assert abs(a - b) <= 1e-4, (a, b)
    m = P.measure_lufs(np.ascontiguousarray(x[:, 0]), sr)
    assert abs(m - oc.measure_lufs(np.ascontiguousarray(x[:, 0]), sr)) <= 1e-4
    assert P.measure_lufs(x, sr) == a                                  # bit-reproducible


def test_true_peak_edges_and_intersample(P):
    from oracle import chain as oc
    sr = 44100
    # a sine at fs/4 sampled at 45 degrees peaks between samples: true peak ~ +3 dB over sample peak
    n = 4000
    x = (0.5 * np.sin(2 * np.pi * (sr / 4) * np.arange(n) / sr + np.pi / 4)).astype(np.float32)
    assert abs(P.true_peak_dbfs(x, sr) - oc.true_peak_dbfs(x)) <= 0.01
    # energy at the very first/last samples exercises the zero-padded edges
    y = np.zeros((5000, 2), np.float32)
    y[0, 0], y[-1, 1], y[2500, 0] = 0.9, -0.8, 0.3
    assert abs(P.true_peak_dbfs(y, sr) - oc.true_peak_dbfs(y)) <= 0.01


def test_quantizer_bit_exact_and_philox_statistics(P):
    from oracle import chain as oc
    from mm_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(3)
    n = 100_003
    x = np.clip(rng.standard_normal((n, 2)) * 0.4, -1.2, 1.2).astype(np.float32)
    x[:10, 0] = [1.0, -1.0, 0.0, np.nan, np.inf, -np.inf, 1.5, -1.5, 32766.5 / 32767, 0.5 / 32767]
    noise = (rng.random((n, 2)) + rng.random((n, 2)) - 1.0).astype(np.float32)
    # ties: samples that land exactly on .5 with zero noise must round half to even
    x[20:30, 1] = (np.arange(10) + 0.5).astype(np.float32) / np.float32(32767)
    noise[20:30, 1] = 0.0
    b = eng.upload([x], 44100)
    q = eng.quantize_int16(b, noise=noise[None])[0]
    assert np.array_equal(q, oc.quantize_int16(x, noise))
    # counter-based Philox TPDF: deterministic per seed, triangular on (-1, 1)
    z = eng.upload([np.zeros((400_000, 2), np.float32)], 44100)
    q1 = eng.quantize_int16(z, seed=7)[0]
    q2 = eng.quantize_int16(z, seed=7)[0]
    q3 = eng.quantize_int16(z, seed=8)[0]
    assert np.array_equal(q1, q2) and not np.array_equal(q1, q3)
    assert set(np.unique(q1)) <= {-1, 0, 1}
    frac = np.mean(q1 == 0)
    assert abs(frac - 0.75) < 0.01           # P(|tri| < 0.5) = 0.75
    assert abs(np.mean(q1 == 1) - 0.125) < 0.005 and abs(np.mean(q1 == -1) - 0.125) < 0.005
    assert abs(np.corrcoef(q1[:, 0], q1[:, 1])[0, 1]) < 0.01


def test_export_audio_wav_bytes(P):
    sr = 44100
    x = (0.25 * np.sin(2 * np.pi * 440 * np.arange(sr) / sr)).astype(np.float32)
    wav = P.export_audio(np.stack([x, x], axis=1), sr, 2, "wav", dither_type="tpdf")
    assert wav[:4] == b"RIFF" and wav[8:12] == b"WAVE"
    back, sr2 = P.load_audio_from_bytes(wav, "wav")
    assert sr2 == sr and back.shape == (sr, 2) and np.max(np.abs(back[:, 0] - x)) < 2.5 / 32768


def test_correlation_conventions(P):
    n = 10000
    t = np.arange(n, dtype=np.float32)
    s = np.sin(t * 0.01).astype(np.float32)
    assert P.measure_stereo_correlation(s) is None
    assert abs(P.measure_stereo_correlation(np.stack([s, s], 1)) - 1.0) < 1e-9
    assert abs(P.measure_stereo_correlation(np.stack([s, -s], 1)) + 1.0) < 1e-9
    assert P.measure_stereo_correlation(np.zeros((n, 2), np.float32)) is None
    assert P.measure_stereo_correlation(np.stack([s, np.full(n, 0.25, np.float32)], 1)) == 0.0


def test_timeline_lra_vectorscope_and_batch_analyzer(P):
    from oracle import chain as oc
    from mm_b200 import synth
    g = load_golden("analyzers")
    for tag in "abc":
        x, sr = g[f"{tag}_input"], int(g[f"{tag}_sr"])
        tl, step = P.compute_lufs_timeline(x, sr)
        ref = g[f"{tag}_timeline"]
        assert len(tl) == len(ref) and step == float(g[f"{tag}_timeline_step"])
        for a, b in zip(tl, ref):
            assert (a is None and np.isnan(b)) or abs(a - b) <= 0.011          # both sides round to 0.01
        assert np.allclose(np.array(P.compute_vectorscope_points(x)), g[f"{tag}_vscope"])
    # a longer track: many sliding segments, 3 s blocks for the loudness range
    sr = 44100
    x = synth.numpy_track(9, sr, 14.0)
    tl, step = P.compute_lufs_timeline(x, sr)
    rtl, rstep = oc.compute_lufs_timeline(x, sr)
    assert step == rstep and len(tl) == len(rtl)
    assert max(abs(a - b) for a, b in zip(tl, rtl)) <= 0.011
    tl3, _ = oc.compute_lufs_timeline(x, sr, block_sec=3.0, max_points=200)
    vals = np.array([v for v in tl3 if v is not None and v > -70])
    lra_ref = max(0.0, float(np.percentile(vals, 95) - np.percentile(vals, 10)))
    assert abs(P.loudness_range_lu(x, sr) - lra_ref) <= 0.03
    # batch analyzer == per-track calls
    tracks = [synth.numpy_track(30 + i, sr, 1.2) for i in range(5)]
    recs = P.analyze_batch(tracks, sr)
    for t, r in zip(tracks, recs):
        assert abs(r["lufs"] - oc.measure_lufs(t, sr)) <= 0.01
        assert abs(r["true_peak_dbfs"] - oc.true_peak_dbfs(t)) <= 0.01
        assert abs(r["correlation"] - oc.measure_stereo_correlation(t)) <= 1e-6
        assert abs(r["sample_peak"] - float(np.max(np.abs(t)))) <= 1e-7
        assert len(r["spectrum_bars"]) == 64 and len(r["spectrum_bars_side"]) == 64
