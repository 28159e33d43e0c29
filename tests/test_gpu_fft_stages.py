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
