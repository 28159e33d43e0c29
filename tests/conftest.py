import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "audio-mastering-web_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def fft_stage_cases(mod, g):
    """The calls behind tests/golden/fft_stages.npz (make_golden_fft.py), against any module with the reference's names."""
    x, sr = g["input"], int(g["sr"])
    odd = np.ascontiguousarray(x[:20011])
    return {
        "denoise_medium": lambda: mod.apply_spectral_denoise(x, sr, strength=0.5, noise_percentile=15.0),
        "denoise_strong_odd": lambda: mod.apply_spectral_denoise(odd, sr, strength=0.9, noise_percentile=20.0),
        "denoise_mono_short": lambda: mod.apply_spectral_denoise(np.ascontiguousarray(x[:2500, 0]), sr, strength=0.35, noise_percentile=10.0),
        "denoise_p37": lambda: mod.apply_spectral_denoise(np.ascontiguousarray(x[:12345, 1]), sr, strength=1.0, noise_percentile=37.5),
        "resample_48_44": lambda: mod.resample_audio(x, 48000, 44100),
        "resample_44_48": lambda: mod.resample_audio(odd, 44100, 48000),
        "resample_mono_96": lambda: mod.resample_audio(np.ascontiguousarray(x[:9999, 0]), 48000, 96000),
        "resample_down_even": lambda: mod.resample_audio(np.ascontiguousarray(x[:20000]), 48000, 24000),
        "exciter_os2": lambda: mod.apply_harmonic_exciter(x * np.float32(2.0), sr, exciter_db=2.0, mode="tape", oversample=2),
        "exciter_os4_mono": lambda: mod.apply_harmonic_exciter(np.ascontiguousarray(x[:15001, 0]) * np.float32(3.0), sr, exciter_db=1.5,
                                                               mode="warm", oversample=4),
    }


def dyneq_default_cases():
    """(name, input, sr, decimation) of tests/golden/dyneq_default.npz (make_golden_dyneq.py): inputs are regenerated from the
    deterministic track recipe, only the reference's outputs are stored."""
    sys.path.insert(0, GOLDEN)
    import make_golden_dyneq as mg
    for name, spec in mg.CASES.items():
        x, sr = mg.case_input(name)
        yield name, x, sr, spec[5]


def tpdf_noise(seed, shape2d):
    """Recipe of tests/golden/make_golden.py for the dither buffer of a chain case."""
    rng = np.random.default_rng(int(seed))
    return (rng.random(shape2d) + rng.random(shape2d) - 1.0).astype(np.float32)


@pytest.fixture(scope="session")
def gpu_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mm_b200
    return mm_b200
