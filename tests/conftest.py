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
