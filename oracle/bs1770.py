"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU restatement of the loudness meter the reference calls through the third-party
package ``pyloudnorm`` (requirement ``pyloudnorm>=0.2.0``, unpinned, NOT vendored under
/root/reference and NOT installed in this image).  Reference call sites:
``backend/app/pipeline.py:646-648`` (normalize_lufs), ``:660-662`` (measure_lufs),
``:686-692`` (compute_lufs_timeline).

What is restated here is pyloudnorm's published algorithm (pyloudnorm ``meter.py`` /
``iirfilter.py`` / ``util.py``, which implement ITU-R BS.1770-4):

* K-weighting = two causal biquads run with ``scipy.signal.lfilter`` from zero state,
  coefficients from the RBJ cookbook evaluated at the actual sample rate
  (high-shelf +4 dB, Q=1/sqrt(2), 1500 Hz; high-pass Q=0.5, 38 Hz);
* the filtered signal is written back into a copy of the input, i.e. it keeps the
  input dtype (float32 buffers are rounded to float32 after each of the two stages);
* 400 ms blocks, 75 % overlap, block j spans ``[int(0.4*(0.25 j)*sr), int(0.4*(0.25 j+1)*sr))``;
* absolute gate -70 LKFS (``>=``), relative gate -10 LU below the abs-gated mean (``>``);
* channel weights [1, 1, 1, 1.41, 1.41].

PARITY PINNING: no reference test pins an LUFS value (only ranges,
``backend/tests/test_pipeline.py:52-68``), and pyloudnorm itself is absent, so against the
third-party package this module is "parity unpinned".  It is pinned instead against
BS.1770 known answers (``tests/test_oracle_bs1770.py``): a 997 Hz full-scale sine in one
channel reads -3.01 LKFS, both channels 0.0 LKFS, and level/linearity properties.
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _sg

CHANNEL_GAINS = (1.0, 1.0, 1.0, 1.41, 1.41)
BLOCK_SEC = 0.4
OVERLAP = 0.75
ABS_GATE = -70.0


def k_weighting_coeffs(rate: float):
    """Return [(b, a), (b, a)] for the shelf and the high-pass stage (float64, a[0] == 1)."""
    out = []
    # stage 1: RBJ high shelf, G = +4 dB, Q = 1/sqrt(2), fc = 1500 Hz
    G, Q, fc = 4.0, 1.0 / np.sqrt(2.0), 1500.0
    A = 10.0 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    cw = np.cos(w0)
    sA = np.sqrt(A)
    b0 = A * ((A + 1) + (A - 1) * cw + 2 * sA * alpha)
    b1 = -2 * A * ((A - 1) + (A + 1) * cw)
    b2 = A * ((A + 1) + (A - 1) * cw - 2 * sA * alpha)
    a0 = (A + 1) - (A - 1) * cw + 2 * sA * alpha
    a1 = 2 * ((A - 1) - (A + 1) * cw)
    a2 = (A + 1) - (A - 1) * cw - 2 * sA * alpha
    out.append((np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0))
    # stage 2: RBJ high pass, Q = 0.5, fc = 38 Hz
    Q, fc = 0.5, 38.0
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    cw = np.cos(w0)
    b0 = (1 + cw) / 2
    b1 = -(1 + cw)
    b2 = (1 + cw) / 2
    a0 = 1 + alpha
    a1 = -2 * cw
    a2 = 1 - alpha
    out.append((np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0))
    return out


def block_bounds(n_samples: int, rate: float, block_sec: float = BLOCK_SEC):
    """(lower, upper) int64 arrays of the gating blocks, pyloudnorm's int() truncation."""
    T_g = block_sec
    step = 1.0 - OVERLAP
    T = n_samples / rate
    num_blocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
    lo = np.empty(max(num_blocks, 0), dtype=np.int64)
    hi = np.empty(max(num_blocks, 0), dtype=np.int64)
    for j in range(max(num_blocks, 0)):
        lo[j] = int(T_g * (j * step) * rate)
        hi[j] = int(T_g * (j * step + 1) * rate)
    return lo, hi


def gate_blocks(z: np.ndarray) -> float:
    """Two-pass gating of block mean squares ``z`` (channels, blocks) -> LUFS."""
    n_ch = z.shape[0]
    G = CHANNEL_GAINS
    with np.errstate(divide="ignore", invalid="ignore"):
        wsum = np.zeros(z.shape[1])
        for i in range(n_ch):
            wsum = wsum + G[i] * z[i]
        l = -0.691 + 10.0 * np.log10(wsum)
        j1 = np.nonzero(l >= ABS_GATE)[0]
        if j1.size:
            zavg = [np.mean(z[i, j1]) for i in range(n_ch)]
        else:
            zavg = [np.nan] * n_ch
        gamma_r = -0.691 + 10.0 * np.log10(np.sum([G[i] * zavg[i] for i in range(n_ch)])) - 10.0
        j2 = np.nonzero((l > gamma_r) & (l > ABS_GATE))[0]
        if j2.size:
            zavg = np.array([np.mean(z[i, j2]) for i in range(n_ch)])
        else:
            zavg = np.zeros(n_ch)  # nan_to_num(mean of empty)
        zavg = np.nan_to_num(zavg)
        return float(-0.691 + 10.0 * np.log10(np.sum([G[i] * zavg[i] for i in range(n_ch)])))


class Meter:
    """Drop-in for ``pyloudnorm.Meter(rate)`` restricted to what the reference uses."""

    def __init__(self, rate, filter_class="K-weighting", block_size=BLOCK_SEC):
        self.rate = rate
        self.block_size = block_size
        self._coeffs = k_weighting_coeffs(rate)

    def _validate(self, data):
        if not isinstance(data, np.ndarray):
            raise ValueError("Data must be of type numpy.ndarray.")
        if not np.issubdtype(data.dtype, np.floating):
            raise ValueError("Data must be floating point.")
        if data.ndim == 2 and data.shape[1] > 5:
            raise ValueError("Audio must have five channels or less.")
        if data.shape[0] < self.block_size * self.rate:
            raise ValueError("Audio must have length greater than the block size.")

    def k_weight(self, data):
        buf = data.copy()
        self._validate(buf)
        if buf.ndim == 1:
            buf = buf.reshape(buf.shape[0], 1)
        for b, a in self._coeffs:
            for ch in range(buf.shape[1]):
                buf[:, ch] = _sg.lfilter(b, a, buf[:, ch])
        return buf

    def block_mean_squares(self, data):
        buf = self.k_weight(data)
        lo, hi = block_bounds(buf.shape[0], self.rate, self.block_size)
        z = np.zeros((buf.shape[1], lo.size))
        scale = 1.0 / (self.block_size * self.rate)
        for i in range(buf.shape[1]):
            for j in range(lo.size):
                z[i, j] = scale * np.sum(np.square(buf[lo[j]:hi[j], i]))
        return z

    def integrated_loudness(self, data):
        return gate_blocks(self.block_mean_squares(data))
