"""BS.1770 known answers and properties for the oracle's restatement of pyloudnorm's meter
(oracle/bs1770.py): the only pinning available for the third-party arithmetic."""
import numpy as np
import pytest

from oracle import bs1770


def _sine(sr, f, dur, amp=1.0):
    t = np.arange(int(sr * dur)) / sr
    return (amp * np.sin(2 * np.pi * f * t)).astype(np.float32)


@pytest.mark.parametrize("sr", [44100, 48000, 96000])
def test_997hz_full_scale_sine(sr):
    s = _sine(sr, 997.0, 3.0)
    m = bs1770.Meter(sr)
    # ITU-R BS.1770: 0 dBFS 997 Hz sine -> -3.01 LKFS in one channel, 0.0 in two; pyloudnorm's RBJ
    # coefficients land within 0.06 dB of the table values
    assert abs(m.integrated_loudness(np.stack([s, np.zeros_like(s)], axis=1)) + 3.01) < 0.06
    assert abs(m.integrated_loudness(np.stack([s, s], axis=1)) - 0.0) < 0.06
    assert abs(m.integrated_loudness(s) + 3.01) < 0.06


def test_level_linearity_and_gating():
    sr = 48000
    s = _sine(sr, 1000.0, 4.0, 0.25)
    m = bs1770.Meter(sr)
    l0 = m.integrated_loudness(s)
    assert abs(m.integrated_loudness((s * np.float32(0.5)).astype(np.float32)) - (l0 - 6.0206)) < 1e-3
    # silence appended: the absolute / relative gates remove it
    padded = np.concatenate([s, np.zeros(sr * 4, np.float32)])
    # (three straddling blocks with partial energy survive the relative gate: ~0.17 dB lower)
    assert 0.0 <= l0 - m.integrated_loudness(padded) < 0.25
    # quiet tail 20 dB down is removed by the relative gate (-10 LU)
    tail = np.concatenate([s, (s * np.float32(0.05)).astype(np.float32)])
    assert 0.0 <= l0 - m.integrated_loudness(tail) < 0.25


def test_short_input_raises_and_block_bounds_are_hop_multiples():
    with pytest.raises(ValueError):
        bs1770.Meter(48000).integrated_loudness(np.zeros(1000, np.float32))
    for sr in (22050, 44100, 48000, 88200, 96000, 192000):
        lo, hi = bs1770.block_bounds(sr * 5, sr)
        hop = int(round(sr * 0.1))
        assert np.all(lo % hop == 0) and np.all((hi - lo) == 4 * hop)
