"""ORACLE (test infrastructure, never shipped or timed as the product).

numpy/scipy restatement of the reference's mastering hot path -- ``backend/app/pipeline.py``
(v1 chain), ``backend/app/chain.py`` + ``backend/app/modules/*.py`` (v2 chain),
``backend/app/routers/tools.py:44-54`` (true peak) -- written from the behaviour described in
SURVEY.md section 8a.  Every function cites the reference lines it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it.

Third-party arithmetic: ``scipy.signal`` (butter / filtfilt / lfilter / resample_poly) is
installed in this image (scipy 1.18.1; the reference asks for ``scipy>=1.11``, unpinned) and is
called directly, as the reference does.  ``pyloudnorm`` is absent: see ``oracle/bs1770.py``.
``pedalboard`` is absent: like the reference itself in that situation
(``backend/app/pipeline.py:442-446``) the multiband stage uses the memoryless soft-knee branch;
the JUCE-compressor branch is PARITY UNPINNED and not restated.

PINNING: ``tests/golden/make_golden.py`` runs the unmodified reference (through
``oracle/ref_harness.py``) in the build container and commits its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against them,
``tests/test_oracle_vs_reference.py`` compares live when ``/root/reference`` is present.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import signal as sg

from . import bs1770

# -- presets / constants (values from backend/app/pipeline.py:56-110) ---------------------------
PRESET_LUFS = {"spotify": -14.0, "youtube": -14.0, "apple": -16.0, "club": -9.0, "broadcast": -24.0}

_STYLE_FIELDS = ("lufs", "sub", "bass", "mids", "presence", "air", "comp_mult", "exciter_db", "imager_width", "parallel_mix")
_STYLE_ROWS = {
    "standard":    (-14.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0),
    "edm":         (-9.0, 1.8, 0.9, -0.3, 0.6, 0.9, 1.3, 0.6, 1.25, 0.3),
    "hiphop":      (-13.0, 1.4, 0.7, 0.5, 0.3, 0.2, 1.2, 0.3, 1.1, 0.35),
    "classical":   (-18.0, -0.5, 0.0, 0.0, 0.3, 0.6, 0.45, 0.0, 1.05, 0.0),
    "podcast":     (-16.0, -1.2, -0.4, 0.9, 0.7, 0.0, 1.1, 0.0, 1.0, 0.2),
    "lofi":        (-18.0, 0.4, 0.6, -0.6, -1.0, -1.8, 0.65, 0.2, 0.9, 0.0),
    "house_basic": (-10.0, 1.8, 0.9, -0.5, 0.8, 1.0, 1.35, 0.8, 1.3, 0.3),
    "dry_vocal":   (-14.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0),
}
STYLE_CONFIGS = {k: dict(zip(_STYLE_FIELDS, v)) for k, v in _STYLE_ROWS.items()}

TRUE_PEAK_LIMIT_DB = -1.5
MULTIBAND_CROSSOVERS_HZ = (214.0, 3500.0, 10000.0)
V2_CROSSOVERS_HZ = (214.0, 2230.0, 10000.0)          # backend/app/chain.py:116
MULTIBAND_CONFIG = ((-7.2, 1.0, -7.2, 1.5), (-18.5, 2.2, -18.5, 1.8), (-17.0, 1.55, -17.0, 1.65), (-15.0, 1.35, -15.0, 1.2))
MAXIMIZER_THRESHOLD_DB = -2.5
MAXIMIZER_MARGIN_DB = -0.3
FINAL_TRIM_DB = 0.5


def _cols(x):
    """View (n,) as (n,1); returns (2-D array, was_mono)."""
    x = np.asarray(x)
    return (x[:, None], True) if x.ndim == 1 else (x, False)


def _uncols(y, mono):
    return y[:, 0] if mono else y


# -- zero-phase filtering -----------------------------------------------------------------------
def zero_phase(b, a, x):
    """``_safe_filtfilt`` (pipeline.py:36-52): filtfilt, else causal lfilter, else identity."""
    try:
        return sg.filtfilt(b, a, x)
    except (np.linalg.LinAlgError, ValueError):
        try:
            return sg.lfilter(b, a, x)
        except Exception:
            return x


def filtfilt_explicit(b, a, x):
    """What ``scipy.signal.filtfilt(b, a, x)`` does with its defaults, spelled out.

    This is the contract the CUDA sweep kernels implement: odd extension by
    ``3*max(len(a), len(b))`` samples on each side, a forward DF2T pass started from
    ``lfilter_zi(b, a) * ext[0]``, a backward pass over the whole extended forward output
    started from ``zi * y_fwd[-1]``, then the pads are dropped.
    """
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    x = np.asarray(x, dtype=np.float64)
    pad = 3 * max(len(a), len(b))
    if x.shape[0] <= pad:
        raise ValueError("input shorter than padlen")
    ext = np.concatenate([2.0 * x[0] - x[pad:0:-1], x, 2.0 * x[-1] - x[-2:-pad - 2:-1]])
    zi = sg.lfilter_zi(b, a)
    y, _ = sg.lfilter(b, a, ext, zi=zi * ext[0])
    yr, _ = sg.lfilter(b, a, y[::-1], zi=zi * y[-1])
    return yr[::-1][pad:-pad]


# -- pointwise / reduction stages ---------------------------------------------------------------
def remove_dc_offset(audio):
    """pipeline.py:134-138."""
    audio = np.asarray(audio)
    if audio.ndim == 1:
        return audio - np.mean(audio)
    return audio - np.mean(audio, axis=0, keepdims=True)


def remove_intersample_peaks(audio, headroom_db=0.5):
    """pipeline.py:141-149: one peak over all channels; scale to -headroom if above; clip."""
    pk = np.nanmax(np.abs(audio))
    if (not np.isfinite(pk)) or pk <= 1e-12:
        return np.nan_to_num(audio, nan=0.0, posinf=1.0, neginf=-1.0)
    lim = 10 ** (-headroom_db / 20)
    if pk > lim:
        audio = audio * (lim / pk)
    return np.clip(audio, -1.0, 1.0)


def apply_output_edge_fade_in(audio, sr, fade_ms=6.0):
    """pipeline.py:152-167."""
    if fade_ms <= 0 or sr <= 0 or audio.size == 0:
        return audio
    nf = int(round(sr * (fade_ms / 1000.0)))
    nf = max(2, min(nf, int(sr * 0.1)))
    out = np.array(audio, dtype=np.float32, copy=True, order="C")
    k = min(nf, out.shape[0])
    ramp = np.linspace(0.0, 1.0, k, dtype=np.float32)
    if out.ndim == 1:
        out[:k] *= ramp
    else:
        out[:k, :] *= ramp[:, None]
    return out


# -- studio EQ ----------------------------------------------------------------------------------
def target_curve_designs(sr):
    """pipeline.py:170-184."""
    nyq = sr / 2.0
    hp = sg.butter(2, min(40.0 / nyq, 0.99), btype="high")
    lp = sg.butter(2, min(18000.0 / nyq, 0.99), btype="low")
    fp = min(3000.0 / nyq, 0.99)
    pres = sg.butter(1, [fp * 0.7, fp * 1.3], btype="band")
    fm = min(300.0 / nyq, 0.99)
    mud = sg.butter(1, [fm * 0.7, fm * 1.3], btype="band")
    return hp, lp, pres, mud, 10 ** (0.35 / 20), 10 ** (-0.25 / 20)


def apply_target_curve(audio, sr, phase_mode="minimum", eq_ms=False):
    """pipeline.py:238-273 (IIR path; ``eq_ms`` M/S variant :248-255). linear_phase: second wave."""
    audio = np.asarray(audio)
    if audio.ndim == 2 and audio.shape[1] == 2 and eq_ms:
        mid = ((audio[:, 0] + audio[:, 1]) * 0.5).astype(np.float32)
        side = ((audio[:, 0] - audio[:, 1]) * 0.5).astype(np.float32)
        m = apply_target_curve(mid, sr, phase_mode, False)
        s = apply_target_curve(side, sr, phase_mode, False)
        return np.stack([np.clip(m + s, -1, 1).astype(np.float32), np.clip(m - s, -1, 1).astype(np.float32)], axis=1)
    if phase_mode == "linear_phase":
        raise NotImplementedError("linear-phase EQ is second-wave scope (SURVEY 8f)")
    x, mono = _cols(audio)
    hp, lp, pres, mud, gp, gm = target_curve_designs(sr)
    out = np.zeros_like(x)
    for c in range(x.shape[1]):
        v = zero_phase(*lp, zero_phase(*hp, x[:, c]))
        out[:, c] = v + (gp - 1.0) * zero_phase(*pres, v) + (gm - 1.0) * zero_phase(*mud, v)
    return _uncols(out, mono)


# -- envelope follower / de-esser ----------------------------------------------------------------
def _follow(x32, atk, rel):
    env = np.empty(x32.shape[0], dtype=np.float32)
    e = np.float32(abs(x32[0]))
    env[0] = e
    for i in range(1, x32.shape[0]):
        v = abs(x32[i])
        c = atk if v > e else rel
        e = np.float32(c * float(e) + (1.0 - c) * float(v))
        env[i] = e
    return env


try:  # speed only; same arithmetic (no fastmath here)
    import numba as _nb

    @_nb.njit(cache=False)
    def _follow_nb(x32, atk, rel):
        n = x32.shape[0]
        env = np.empty(n, dtype=np.float32)
        env[0] = abs(x32[0])
        for i in range(1, n):
            v = abs(x32[i])
            if v > env[i - 1]:
                env[i] = atk * env[i - 1] + (1.0 - atk) * v
            else:
                env[i] = rel * env[i - 1] + (1.0 - rel) * v
        return env
except Exception:  # pragma: no cover
    _follow_nb = None


def envelope_follower(x, sr, attack_sec, release_sec):
    """pipeline.py:495-518: one-pole attack/release follower, f64 coefficients, f32 state."""
    if len(x) == 0:
        return x
    atk = float(np.exp(-1.0 / max(1e-6, sr * attack_sec)))
    rel = float(np.exp(-1.0 / max(1e-6, sr * release_sec)))
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    if _follow_nb is not None:
        return _follow_nb(x32, atk, rel)
    return _follow(x32, atk, rel)


def apply_deesser(audio, sr, threshold_db=-6.0, ratio=3.0, freq_lo=5000.0, freq_hi=9000.0, attack_ms=4.0, release_ms=85.0):
    """pipeline.py:1200-1264."""
    x, mono = _cols(audio)
    nyq = sr / 2.0
    lo, hi = min(freq_lo / nyq, 0.97), min(freq_hi / nyq, 0.97)
    if lo >= hi:
        return _uncols(x, mono)
    b, a = sg.butter(2, [lo, hi], btype="band")
    thr = 10 ** (threshold_db / 20.0)
    k = max(3, int(sr * 0.0015))
    k += 1 - (k % 2)
    ker = np.ones(k, dtype=np.float32) / float(k)
    out = x.copy().astype(np.float32)
    for c in range(x.shape[1]):
        xc = x[:, c].astype(np.float32)
        sc = zero_phase(b, a, xc).astype(np.float32)
        env = envelope_follower(np.abs(sc), float(sr), attack_ms / 1000.0, release_ms / 1000.0)
        red = np.where(env > thr, thr + (env - thr) / ratio, env)
        g = np.where(env > 1e-10, red / (env + 1e-12), 1.0)
        g = np.clip(g, 0.35, 1.0).astype(np.float32)
        g = np.clip(np.convolve(g, ker, mode="same").astype(np.float32), 0.35, 1.0)
        out[:, c] = xc - sc + sc * g
    return _uncols(out, mono)


# -- dynamics -----------------------------------------------------------------------------------
def compress_soft_knee(audio, threshold_db=-18.0, ratio=2.5, knee_db=6.0, max_upward_boost_db=12.0):
    """pipeline.py:282-330 (memoryless fallback compressor)."""
    if ratio <= 0.0:
        return audio
    t = 10 ** (threshold_db / 20.0)
    ax = np.abs(audio)
    sgn = np.sign(audio)
    if ratio < 1.0:
        lvl = np.where(ax > 1e-12, 20.0 * np.log10(np.maximum(ax, 1e-12)), -100.0)
        boost = np.clip((threshold_db - lvl) * (1.0 - ratio), 0.0, max(0.1, float(max_upward_boost_db)))
        return (sgn * np.clip(ax * 10 ** (boost / 20.0), 0.0, 1.0)).astype(np.float32)
    if ratio == 1.0:
        return audio
    knee_db = max(0.0, float(knee_db))
    if knee_db < 0.5:
        return (sgn * np.minimum(ax, t + np.maximum(ax - t, 0.0) / ratio)).astype(np.float32)
    lo = t * 10 ** (-knee_db / 20.0)
    hi = t * 10 ** (knee_db / 20.0)
    slope = (t + (hi - t) / ratio - lo) / (hi - lo)
    y = np.where(ax <= lo, ax, np.where(ax >= hi, t + (ax - t) / ratio, lo + (ax - lo) * slope))
    return (sgn * np.clip(y, 0.0, None)).astype(np.float32)


def hard_limit(audio, threshold_db=-1.0):
    """pipeline.py:276-279."""
    lim = 10 ** (threshold_db / 20.0)
    return np.clip(audio, -lim, lim).astype(np.float32)


def split_bands(audio, sr, crossovers_hz):
    """pipeline.py:333-364: four zero-phase Butterworth-2 bands, float64."""
    nyq = sr / 2.0
    f1, f2, f3 = (min(c / nyq, 0.99) for c in crossovers_hz)
    x, mono = _cols(audio)
    lp1, hp1 = sg.butter(2, f1, "low"), sg.butter(2, f1, "high")
    lp2, hp2 = sg.butter(2, f2, "low"), sg.butter(2, f2, "high")
    lp3, hp3 = sg.butter(2, f3, "low"), sg.butter(2, f3, "high")
    bands = [np.empty(x.shape, dtype=np.float64) for _ in range(4)]
    for c in range(x.shape[1]):
        v = x[:, c]
        bands[0][:, c] = zero_phase(*lp1, v)
        bands[1][:, c] = zero_phase(*lp2, zero_phase(*hp1, v))
        bands[2][:, c] = zero_phase(*lp3, zero_phase(*hp2, v))
        bands[3][:, c] = zero_phase(*hp3, v)
    return [_uncols(bd, mono) for bd in bands]


# -- envelope-compressor branch (pedalboard / JUCE) ---------------------------------------------------------------
# PARITY UNPINNED: pedalboard is not installed here and nothing in the reference's tests touches this branch
# (pipeline.py:373-411, :442-465).  What follows restates the published JUCE sources that pedalboard.Compressor wraps --
# juce::dsp::Compressor<float>::processSample over juce::dsp::BallisticsFilter<float> (peak level type):
#     a = |x|;  cte = a > y_prev ? cteAT : cteRL;  y = a + cte * (y_prev - a)        cte = exp(-2 pi 1000 / (sr * t_ms)), y(-1) = 0
#     gain = y < thr ? 1 : pow(y * (1 / thr), 1 / ratio - 1);   out = gain * x       thr = 10^(dB / 20), all float32
# JUCE's per-block snap-to-zero of states below 1e-8 is not modelled.
def _ballistics(x32, cat, crl):
    env = np.empty(x32.shape[0], dtype=np.float32)
    y = np.float32(0.0)
    for i in range(x32.shape[0]):
        a = np.float32(abs(x32[i]))
        d = np.float32(y - a)
        m = np.float32(cat * d) if d < 0 else np.float32(crl * d)
        y = np.float32(a + m)
        env[i] = y
    return env


try:
    @_nb.njit(cache=False)
    def _ballistics_nb(x32, cat, crl):
        n = x32.shape[0]
        env = np.empty(n, dtype=np.float32)
        y = np.float32(0.0)
        for i in range(n):
            a = np.float32(abs(x32[i]))
            d = np.float32(y - a)
            if d < np.float32(0.0):
                m = np.float32(cat * d)
            else:
                m = np.float32(crl * d)
            y = np.float32(a + m)
            env[i] = y
        return env
except Exception:  # pragma: no cover
    _ballistics_nb = None


BAND_TIMES_MS = ((10.0, 80.0), (10.0, 80.0), (12.0, 130.0), (18.0, 180.0))      # pipeline.py:451-456


def compress_band_envelope(band, sr, threshold_db, ratio, lim_db, gain, attack_ms=10.0, release_ms=80.0):
    """_compress_band_pedalboard (pipeline.py:373-411) with the JUCE compressor restated (parity unpinned, see above):
    compressor -> hard clip at lim_db -> x gain, float32."""
    b, mono = _cols(np.asarray(band))
    x = np.ascontiguousarray(b, dtype=np.float32)
    cat = np.float32(np.exp(-2.0 * np.pi * 1000.0 / (float(sr) * attack_ms)))
    crl = np.float32(np.exp(-2.0 * np.pi * 1000.0 / (float(sr) * release_ms)))
    thr = np.float32(10.0 ** (threshold_db / 20.0))
    thr_inv = np.float32(1.0) / thr
    pw = np.float32(1.0) / np.float32(max(ratio, 1.0)) - np.float32(1.0)
    out = np.empty_like(x)
    for c in range(x.shape[1]):
        col = np.ascontiguousarray(x[:, c])
        env = (_ballistics_nb or _ballistics)(col, cat, crl)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(env < thr, np.float32(1.0), np.power(env * thr_inv, pw, dtype=np.float32)).astype(np.float32)
        out[:, c] = g * col
    lim = 10 ** (lim_db / 20.0)
    out = (np.clip(out, -lim, lim).astype(np.float32) * gain).astype(np.float32)
    return _uncols(out, mono)


def apply_multiband_dynamics(samples, sr, knee_db=6.0, crossovers_hz=None, band_ratios=None, max_upward_boost_db=12.0,
                             compressor="soft_knee"):
    """pipeline.py:414-481: ``compressor="soft_knee"`` is the numpy branch (:466-474, what the reference runs without
    pedalboard -- pinned); ``"envelope"`` the pedalboard branch (:457-465) on the restated JUCE compressor (unpinned)."""
    samples = np.asarray(samples)
    x = samples.reshape(-1, 1) if samples.ndim == 1 else samples
    cross = crossovers_hz if crossovers_hz and len(crossovers_hz) == 3 else MULTIBAND_CROSSOVERS_HZ
    cross = tuple(float(np.clip(c, 20.0, 20000.0)) for c in cross)
    if cross[0] >= cross[1] or cross[1] >= cross[2]:
        cross = MULTIBAND_CROSSOVERS_HZ
    bands = split_bands(x, float(sr), cross)
    over = tuple(float(r) for r in band_ratios) if band_ratios is not None and len(band_ratios) == 4 else None
    acc = None
    for i, (lim_db, ratio, thr_db, gain) in enumerate(MULTIBAND_CONFIG):
        r = over[i] if over else ratio
        if compressor == "envelope" and r >= 1.0:
            bd = compress_band_envelope(bands[i], sr, thr_db, r, lim_db, gain, *BAND_TIMES_MS[i])
        else:
            bd = compress_soft_knee(bands[i], thr_db, r, knee_db, max_upward_boost_db)
            bd = hard_limit(bd, lim_db) * gain
        acc = bd if acc is None else acc + bd
    out = acc.astype(np.float32)
    return out[:, 0] if x.shape[1] == 1 else out


def apply_maximizer(audio):
    """pipeline.py:484-492."""
    ceil_ = 10 ** (MAXIMIZER_MARGIN_DB / 20.0)
    thr = 10 ** (MAXIMIZER_THRESHOLD_DB / 20.0)
    ax = np.abs(audio)
    y = np.where(ax <= thr, ax, thr + (ax - thr) * (ceil_ - thr) / (1.0 - thr))
    return (np.sign(audio) * np.minimum(y, ceil_)).astype(np.float32)


def apply_maximizer_lookahead(audio, sr, lookahead_ms=6.0):
    """pipeline.py:548-573."""
    delay_n = int(sr * (lookahead_ms / 1000.0))
    if delay_n <= 0 or delay_n >= audio.shape[0]:
        return apply_maximizer(audio)
    a, mono = _cols(audio)
    delayed = np.concatenate([np.zeros((delay_n, a.shape[1]), dtype=a.dtype), a[:-delay_n]], axis=0)
    limited = apply_maximizer(delayed)
    out = np.concatenate([a[:delay_n], limited[delay_n:]], axis=0).astype(np.float32)
    cf = min(delay_n, max(2, int(sr * 0.002)))
    for i in range(cf):
        idx = delay_n - cf + i
        if 0 <= idx < out.shape[0]:
            w = (i + 1) / float(cf)
            out[idx, :] = (1.0 - w) * a[idx, :] + w * limited[idx, :]
    return _uncols(out, mono)


def apply_dynamics(samples, sr, knee_db=6.0, crossovers_hz=None, band_ratios=None, max_upward_boost_db=12.0, compressor="soft_knee"):
    """pipeline.py:610-641."""
    samples = np.asarray(samples)
    x = samples.reshape(-1, 1) if samples.ndim == 1 else samples
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = apply_multiband_dynamics(x, sr, knee_db, crossovers_hz, band_ratios, max_upward_boost_db, compressor=compressor)
    if y.ndim == 1:
        y = y.reshape(-1, 1)
    y = hard_limit(apply_maximizer(y), TRUE_PEAK_LIMIT_DB)
    return y[:, 0] if y.shape[1] == 1 else y


def apply_parallel_compression(audio, sr, mix=0.3, ratio=8.0, threshold_db=-20.0):
    """pipeline.py:1771-1797."""
    mix = float(np.clip(mix, 0.0, 1.0))
    if mix < 0.01:
        return audio
    comp = compress_soft_knee(audio, threshold_db, ratio, 6.0, 0.0)
    return np.clip((audio * (1.0 - mix) + comp * mix).astype(np.float32), -1.0, 1.0).astype(np.float32)


# -- loudness -----------------------------------------------------------------------------------
def measure_lufs(audio, sr):
    """pipeline.py:658-664."""
    try:
        return float(bs1770.Meter(sr).integrated_loudness(audio))
    except Exception:
        return float("nan")


def normalize_lufs(audio, sr, target_lufs):
    """pipeline.py:644-655."""
    try:
        loud = bs1770.Meter(sr).integrated_loudness(audio)
    except Exception:
        return audio
    delta = np.clip(target_lufs - loud, -20.0, 20.0)
    return (audio * 10 ** (delta / 20.0)).astype(np.float32)


def compute_lufs_timeline(audio, sr, block_sec=0.4, max_points=300):
    """pipeline.py:667-697."""
    dur = len(audio) / sr
    blk = int(sr * block_sec)
    if dur <= block_sec or audio.size < blk:
        v = measure_lufs(audio, sr)
        return ([round(v, 2)] if not np.isnan(v) else [None], 0.0)
    npts = min(max_points, max(1, int((dur - block_sec) / (block_sec * 0.25)) + 1))
    step_sec = (dur - block_sec) / max(npts - 1, 1)
    step = int(sr * step_sec)
    meter = bs1770.Meter(sr)
    res = []
    pos = 0
    while pos + blk <= len(audio) and len(res) < max_points:
        try:
            res.append(round(float(meter.integrated_loudness(audio[pos:pos + blk])), 2))
        except Exception:
            res.append(None)
        pos += step
    return res, round(step_sec, 4)


# -- post-normalisation EQ, exciter, imager ------------------------------------------------------
def apply_final_spectral_balance(audio, sr):
    """pipeline.py:576-607."""
    x, mono = _cols(audio)
    nyq = sr / 2.0
    f3 = min(3000.0 / nyq, 0.99)
    f8 = min(8000.0 / nyq, 0.99)
    d3 = sg.butter(1, [f3 * 0.8, f3 * 1.2], "band")
    d16 = sg.butter(2, min(16000.0 / nyq, 0.99), "high")
    dlo = sg.butter(2, min(180.0 / nyq, 0.99), "low")
    d8 = sg.butter(1, [f8 * 0.8, f8 * 1.2], "band")
    g3, g16, glo, g8 = 10 ** (-0.5 / 20), 10 ** (-0.3 / 20), 10 ** (0.3 / 20), 10 ** (0.2 / 20)
    out = x.copy()
    for c in range(x.shape[1]):
        v = x[:, c]
        w = v + (g3 - 1.0) * zero_phase(*d3, v) * 0.25 + (g16 - 1.0) * zero_phase(*d16, v) * 0.25
        w = w + (glo - 1.0) * zero_phase(*dlo, v) * 0.25 + (g8 - 1.0) * zero_phase(*d8, v) * 0.25
        out[:, c] = w * 10 ** (FINAL_TRIM_DB / 20.0)
    return _uncols(out, mono)


def style_eq_bands(sr, style):
    """Active (b, a, g) triples of ``apply_style_eq`` (pipeline.py:1412-1428)."""
    cfg = STYLE_CONFIGS.get(style, STYLE_CONFIGS["standard"])
    nyq = sr / 2.0
    spec = [(30.0, 90.0, cfg["sub"]), (90.0, 280.0, cfg["bass"]), (700.0, 2800.0, cfg["mids"]),
            (2800.0, 9000.0, cfg["presence"]), (10000.0, min(sr * 0.46, 18000.0), cfg["air"])]
    res = []
    for lo, hi, gdb in spec:
        if abs(gdb) < 0.05:
            continue
        lo_n, hi_n = min(lo / nyq, 0.98), min(hi / nyq, 0.98)
        if lo_n >= hi_n:
            continue
        b, a = sg.butter(1, [lo_n, hi_n], "band")
        res.append((b, a, 10 ** (gdb / 20.0)))
    return res


def apply_style_eq(audio, sr, style="standard"):
    """pipeline.py:1401-1434: bands applied one after another, float32 between bands."""
    x, mono = _cols(audio)
    out = x.copy().astype(np.float32)
    for b, a, g in style_eq_bands(sr, style):
        for c in range(out.shape[1]):
            out[:, c] = out[:, c] + (g - 1.0) * zero_phase(b, a, out[:, c])
    return _uncols(out, mono)


def exciter_saturate(x, mode, k=2.0):
    """pipeline.py:1179-1197."""
    x = np.clip(x, -1.0, 1.0)
    if mode == "transistor":
        return x - (x ** 3) / 3.0
    if mode == "tape":
        return np.tanh(k * x) / (k + 1e-8)
    if mode == "tube":
        return x + 0.3 * (x ** 2)
    if mode == "warm":
        return 0.5 * (np.tanh(k * x) / (k + 1e-8) + x + 0.3 * (x ** 2))
    if mode == "digital":
        return np.where(np.abs(x) <= 1.0, x, np.sign(x) * (2.0 - np.abs(x)))
    return np.tanh(k * x) / (k + 1e-8)


def fft_resample(x, num):
    """scipy.signal.resample(x, num) for real 1-D ``x`` (scipy 1.18 ``_signaltools.resample``, time domain, no window), which
    pipeline.py:920-936 and :1294-1320 call: one-sided spectrum, the first min(n, num)//2 + 1 bins kept, the unpaired bin
    at m/2 doubled (down-sampling) or halved (up-sampling), inverse real FFT of length ``num`` scaled by num / n."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    m = min(num, n)
    X = np.fft.rfft(x)[: m // 2 + 1].copy()
    if m % 2 == 0 and num != n:
        X[m // 2] *= 2.0 if num < n else 0.5
    return np.fft.irfft(X * (num / n), n=num)


def resample_audio(audio, sr, target_sr):
    """pipeline.py:920-936."""
    if target_sr <= 0 or sr <= 0:
        raise ValueError("Invalid sample rate")
    if target_sr == sr:
        return np.asarray(audio, dtype=np.float32)
    a, mono = _cols(np.asarray(audio, dtype=np.float64))
    n_out = int(round(a.shape[0] * target_sr / sr))
    out = np.empty((n_out, a.shape[1]), dtype=np.float32)
    for c in range(a.shape[1]):
        out[:, c] = fft_resample(a[:, c], n_out).astype(np.float32)
    return _uncols(out, mono)


def apply_harmonic_exciter(audio, sr, exciter_db=0.0, mode="warm", oversample=1):
    """pipeline.py:1267-1326; ``oversample`` 2..4 runs the side chain on an FFT-up-sampled float32 copy and
    FFT-down-samples the sum (:1294-1320)."""
    if abs(exciter_db) < 0.05:
        return audio
    os_ = max(1, min(4, int(oversample)))
    x, mono = _cols(audio)
    n = x.shape[0]
    if os_ > 1:
        work = np.stack([fft_resample(x[:, c].astype(np.float64), n * os_).astype(np.float32) for c in range(x.shape[1])], axis=1)
    else:
        work = x
    nyq = sr * os_ / 2.0
    b, a = sg.butter(2, min(6000.0 / nyq, 0.97), "high")
    gain = 10 ** (exciter_db / 20.0) - 1.0
    m = mode if mode in ("warm", "tape", "tube", "transistor", "digital") else "warm"
    k = 2.5 if m == "warm" else 2.0
    out = work.copy()
    for c in range(work.shape[1]):
        hf = zero_phase(b, a, work[:, c])
        out[:, c] = work[:, c] + (exciter_saturate(hf, m, k) - hf) * gain * 0.25
    if os_ > 1:
        out = np.stack([fft_resample(out[:, c].astype(np.float64), n).astype(np.float32) for c in range(out.shape[1])], axis=1)
    return _uncols(out.astype(np.float32), mono)


def apply_spectral_denoise(audio, sr, strength=0.5, noise_percentile=15.0):
    """pipeline.py:1472-1524, with scipy's stft / istft (legacy ``_spectral_helper``: boundary='zeros', padded=True, periodic
    Hann, scaling='spectrum') written out: 1024 zeros either side, the tail padded to a whole hop, frames of 2048 every 512,
    rfft(window * frame) / sum(window); per bin the ``noise_percentile`` percentile over the frames capped by 0.85 x the
    median; gain clip(1 - strength (noise / (|Z| + 1e-10))^2, 0.25, 1); inverse: irfft * sum(window), windowed overlap-add
    divided by the overlap-added squared window, the 1024-sample boundary removed."""
    strength = float(np.clip(strength, 0.0, 1.0))
    if strength < 0.01:
        return audio
    a, mono = _cols(audio)
    n = a.shape[0]
    nfft, hop = 2048, 512
    if n < nfft:
        raise ValueError("noverlap must be less than nperseg.")          # what scipy raises once nperseg is cut to n <= 1536
    win = sg.get_window("hann", nfft)
    nadd = (-n) % hop
    F = (n + nadd) // hop + 1
    out = np.zeros_like(a, dtype=np.float32)
    idx = hop * np.arange(F)[:, None] + np.arange(nfft)[None, :]
    for ch in range(a.shape[1]):
        ext = np.concatenate([np.zeros(nfft // 2), a[:, ch].astype(np.float64), np.zeros(nfft // 2 + nadd)])
        Z = np.fft.rfft(ext[idx] * win, axis=1).T / win.sum()               # (1025, F)
        mag = np.abs(Z)
        noise = np.percentile(mag, noise_percentile, axis=1, keepdims=True)
        med = np.median(mag, axis=1, keepdims=True)
        cap = np.minimum(np.maximum(noise, 1e-12), 0.85 * np.maximum(med, 1e-12))
        gain = np.clip(1.0 - strength * (cap / (mag + 1e-10)) ** 2, 0.25, 1.0)
        y = np.fft.irfft((mag * gain * np.exp(1j * np.angle(Z))).T, n=nfft, axis=1) * win.sum()
        acc = np.zeros(ext.shape[0])
        norm = np.zeros(ext.shape[0])
        for f in range(F):
            acc[f * hop: f * hop + nfft] += y[f] * win
            norm[f * hop: f * hop + nfft] += win ** 2
        acc, norm = acc[nfft // 2: -(nfft // 2)], norm[nfft // 2: -(nfft // 2)]
        xo = acc / np.where(norm > 1e-10, norm, 1.0)
        out[:, ch] = np.clip(xo[:n], -1.0, 1.0).astype(np.float32)
    return _uncols(out, mono)


def apply_stereo_imager(audio, width=1.0):
    """pipeline.py:1339-1398, plain width mode (:1329-1336); 4-band / Haas are second-wave."""
    if audio.ndim == 1 or audio.shape[1] == 1:
        return audio
    l = audio[:, 0].astype(np.float32)
    r = audio[:, 1].astype(np.float32)
    mid = (l + r) * 0.5
    side = (l - r) * 0.5 * width
    return np.column_stack([np.clip(mid + side, -1.0, 1.0), np.clip(mid - side, -1.0, 1.0)]).astype(np.float32)


# -- chains -------------------------------------------------------------------------------------
# -- second-wave stages on the same primitives ----------------------------------
def apply_transient_designer(audio, sr, attack_gain=1.0, sustain_gain=1.0):
    """pipeline.py:1736-1768: fast (0.5 / 5 ms) and slow (5 / 100 ms) followers of |x| per channel."""
    attack_gain = float(np.clip(attack_gain, 0.1, 3.0))
    sustain_gain = float(np.clip(sustain_gain, 0.1, 3.0))
    if abs(attack_gain - 1.0) < 0.02 and abs(sustain_gain - 1.0) < 0.02:
        return audio
    a, mono = _cols(audio)
    out = np.zeros_like(a, dtype=np.float32)
    for ch in range(a.shape[1]):
        x = a[:, ch].astype(np.float32)
        ax = np.abs(x)
        fast = envelope_follower(ax, float(sr), 0.0005, 0.005)
        slow = envelope_follower(ax, float(sr), 0.005, 0.1)
        transient = np.maximum(fast - slow, 0.0)
        new_env = transient * attack_gain + slow * sustain_gain
        gain = np.clip(new_env / (fast + 1e-12), 0.0, 4.0).astype(np.float32)
        out[:, ch] = np.clip(x * gain, -1.0, 1.0)
    return _uncols(out, mono)


def apply_maximizer_transient_aware(audio, sr, sensitivity=0.5):
    """pipeline.py:521-545: the maximizer backs off where a fast follower of mean |x| runs ahead of a slow one."""
    a, mono = _cols(audio)
    a = a.astype(np.float32)
    limited = np.array(apply_maximizer(a), dtype=np.float32).reshape(a.shape)
    det = np.mean(np.abs(a), axis=1).astype(np.float32)
    fast = envelope_follower(det, float(sr), 0.0005, 0.002)
    slow = envelope_follower(det, float(sr), 0.01, 0.04)
    diff = np.maximum(fast - slow, 0.0)
    mask = np.clip(diff / (slow + 1e-12) * float(sensitivity), 0.0, 1.0)
    mask = np.minimum(mask, 1.0)
    for ch in range(a.shape[1]):
        limited[:, ch] = limited[:, ch] * (1.0 - mask) + a[:, ch] * mask
    return _uncols(np.clip(limited, -1.0, 1.0).astype(np.float32), mono)


# pipeline.py:1616-1625
DYNAMIC_EQ_MASTERING_BANDS = [
    {"freq": 120, "q": 1.0, "threshold_db": -14, "ratio": 2.0, "attack_ms": 10, "release_ms": 100, "max_cut_db": -4},
    {"freq": 250, "q": 1.2, "threshold_db": -12, "ratio": 2.5, "attack_ms": 8, "release_ms": 80, "max_cut_db": -5},
    {"freq": 400, "q": 1.0, "threshold_db": -12, "ratio": 2.0, "attack_ms": 8, "release_ms": 80, "max_cut_db": -4},
    {"freq": 800, "q": 1.2, "threshold_db": -12, "ratio": 2.0, "attack_ms": 5, "release_ms": 60, "max_cut_db": -4},
    {"freq": 2500, "q": 1.4, "threshold_db": -12, "ratio": 2.5, "attack_ms": 5, "release_ms": 60, "max_cut_db": -5},
    {"freq": 5000, "q": 1.4, "threshold_db": -14, "ratio": 3.0, "attack_ms": 3, "release_ms": 50, "max_cut_db": -6},
    {"freq": 8000, "q": 1.2, "threshold_db": -16, "ratio": 4.0, "attack_ms": 2, "release_ms": 40, "max_cut_db": -8},
    {"freq": 12000, "q": 0.8, "threshold_db": -18, "ratio": 2.0, "attack_ms": 5, "release_ms": 60, "max_cut_db": -4},
]


def apply_dynamic_eq(audio, sr, bands=None):
    """pipeline.py:1628-1700.  ``sg.iirpeak(w0, bw)`` is called with the reference's own (bandwidth-as-Q) arguments, so the
    default bands are unstable / degenerate sections (tests/test_host_design.py); the same scipy calls overflow, raise and
    fall back exactly as they do in the reference."""
    if bands is None:
        bands = DYNAMIC_EQ_MASTERING_BANDS
    a, mono = _cols(audio)
    nyq = sr / 2.0
    out = a.copy().astype(np.float32)
    for band in bands:
        freq, q = float(band.get("freq", 1000)), float(band.get("q", 1.4))
        thr = 10 ** (float(band.get("threshold_db", -12)) / 20.0)
        ratio = float(band.get("ratio", 3.0))
        atk, rel = float(band.get("attack_ms", 5)) / 1000.0, float(band.get("release_ms", 80)) / 1000.0
        max_cut = 10 ** (float(band.get("max_cut_db", -6)) / 20.0)
        if freq <= 0 or freq >= nyq * 0.98:
            continue
        w0 = float(np.clip(freq / nyq, 0.001, 0.98))
        bw = float(np.clip(w0 / max(q, 0.1), 0.001, 0.5))
        bb, aa = sg.iirpeak(w0, bw)
        for ch in range(a.shape[1]):
            x = out[:, ch].copy()
            bs = np.nan_to_num(zero_phase(bb, aa, x.astype(np.float64)).astype(np.float32), nan=0.0, posinf=0.0, neginf=0.0)
            env = np.nan_to_num(envelope_follower(np.abs(bs), float(sr), atk, rel), nan=0.0, posinf=0.0, neginf=0.0)
            g = np.where(env > thr, np.clip((thr + (env - thr) / ratio) / (env + 1e-12), max_cut, 1.0), 1.0).astype(np.float32)
            g = np.clip(np.nan_to_num(g, nan=1.0, posinf=1.0, neginf=1.0), 0.3, 1.0)
            out[:, ch] = x - bs + bs * g
    bad = ~np.isfinite(out)
    if np.any(bad):
        out = np.where(bad, a.astype(np.float32), out)
    return _uncols(np.clip(out, -1.0, 1.0).astype(np.float32), mono)


def apply_high_freq_trim(audio, sr, crossover_hz=5000.0, high_gain=0.9):
    """pipeline.py:1705-1733: low = filtfilt(butter(2, fc)), out = low + high_gain * (x - low), clip."""
    if abs(high_gain - 1.0) < 0.001:
        return audio
    a, mono = _cols(audio)
    b_lp, a_lp = sg.butter(2, min(crossover_hz / (sr / 2.0), 0.98), btype="low", output="ba")
    out = a.copy().astype(np.float32)
    for ch in range(out.shape[1]):
        low = zero_phase(b_lp, a_lp, out[:, ch].astype(np.float64)).astype(np.float32)
        high = out[:, ch].astype(np.float32) - low
        out[:, ch] = low + high_gain * high
    return _uncols(np.clip(out, -1.0, 1.0).astype(np.float32), mono)


def apply_stereoize(audio, sr, width=1.0, delay_ms=8.0, mix=0.12):
    """apply_stereo_imager with the Haas cross-delay (pipeline.py:1339-1398, single-band width)."""
    if audio.ndim == 1 or audio.shape[1] == 1:
        return audio
    wide = np.asarray(apply_stereo_imager(audio, width) if True else audio, dtype=np.float32)
    if width == 1.0:      # the reference still runs the mid/side arithmetic (and its clip) for width 1
        left, right = audio[:, 0].astype(np.float32), audio[:, 1].astype(np.float32)
        mid = (left + right) * np.float32(0.5)
        side = (left - right) * np.float32(0.5) * np.float32(width)
        wide = np.column_stack([np.clip(mid + side, -1.0, 1.0), np.clip(mid - side, -1.0, 1.0)]).astype(np.float32)
    out_l, out_r = wide[:, 0], wide[:, 1]
    delay_n = max(0, min(int(sr * delay_ms / 1000.0), audio.shape[0] - 1))
    m = min(0.35, max(0.0, float(mix)))
    if delay_ms > 0 and m > 0 and delay_n > 0:
        dr = np.concatenate([np.zeros(delay_n, dtype=out_r.dtype), out_r[:-delay_n]])
        dl = np.concatenate([np.zeros(delay_n, dtype=out_l.dtype), out_l[:-delay_n]])
        out_l, out_r = np.clip(out_l + m * dr, -1.0, 1.0), np.clip(out_r + m * dl, -1.0, 1.0)
    return np.column_stack([out_l, out_r]).astype(np.float32)


def apply_stereo_imager_4band(audio, sr, band_widths, crossovers_hz=None, delay_ms=0.0, mix=0.12):
    """apply_stereo_imager in its 4-band mode (pipeline.py:1360-1398): _split_bands, per-band mid/side width, merge."""
    if audio.ndim == 1 or audio.shape[1] == 1:
        return audio
    left, right = audio[:, 0].astype(np.float32), audio[:, 1].astype(np.float32)
    cross = tuple(float(v) for v in crossovers_hz) if crossovers_hz is not None and len(crossovers_hz) == 3 else (214.0, 3500.0, 10000.0)
    cross = tuple(np.clip(c, 20.0, 20000.0) for c in cross)
    if cross[0] >= cross[1] or cross[1] >= cross[2]:
        cross = (214.0, 3500.0, 10000.0)
    bands = split_bands(np.column_stack([left, right]), float(sr), cross)
    out_l, out_r = np.zeros_like(left), np.zeros_like(right)
    for i in range(4):
        bl, br = bands[i][:, 0], bands[i][:, 1]
        mid, side = (bl + br) * 0.5, (bl - br) * 0.5 * float(band_widths[i])
        out_l += np.clip(mid + side, -1.0, 1.0)
        out_r += np.clip(mid - side, -1.0, 1.0)
    out_l, out_r = np.clip(out_l, -1.0, 1.0), np.clip(out_r, -1.0, 1.0)
    delay_n = max(0, min(int(sr * delay_ms / 1000.0), audio.shape[0] - 1))
    m = min(0.35, max(0.0, float(mix)))
    if delay_ms > 0 and m > 0 and delay_n > 0:
        dr = np.concatenate([np.zeros(delay_n, dtype=out_r.dtype), out_r[:-delay_n]])
        dl = np.concatenate([np.zeros(delay_n, dtype=out_l.dtype), out_l[:-delay_n]])
        out_l, out_r = np.clip(out_l + m * dr, -1.0, 1.0), np.clip(out_r + m * dl, -1.0, 1.0)
    return np.column_stack([out_l, out_r]).astype(np.float32)


def build_linear_phase_ir(sr, n_fft=4096):
    """_build_linear_phase_ir (pipeline.py:187-217)."""
    (b_hp, a_hp), (b_lp, a_lp), (b_pres, a_pres), (b_mud, a_mud), g_presence, g_mud = target_curve_designs(sr)
    w = np.pi * np.arange(n_fft // 2 + 1) / (n_fft // 2)
    H = sg.freqz(b_hp, a_hp, worN=w)[1] * sg.freqz(b_lp, a_lp, worN=w)[1] * (
        1.0 + (g_presence - 1.0) * sg.freqz(b_pres, a_pres, worN=w)[1] + (g_mud - 1.0) * sg.freqz(b_mud, a_mud, worN=w)[1])
    mag = np.clip(np.abs(H), 1e-8, 1e8)
    N = n_fft
    phase = -2.0 * np.pi * np.arange(N // 2 + 1, dtype=np.float64) * (N - 1) / (2.0 * N)
    full = np.zeros(N, dtype=np.complex128)
    full[: N // 2 + 1] = mag * np.exp(1j * phase)
    for k in range(1, N // 2):
        full[N - k] = np.conj(full[k])
    full[N // 2] = np.real(full[N // 2])
    return np.ascontiguousarray(np.fft.ifft(full).real.astype(np.float32))


def apply_target_curve_linear_phase(audio, sr, n_fft=4096):
    """pipeline.py:220-235: fftconvolve(x, ir, mode="same") per channel, clip."""
    a, mono = _cols(audio)
    ir = build_linear_phase_ir(sr, n_fft)
    out = np.zeros_like(a, dtype=np.float32)
    for ch in range(a.shape[1]):
        out[:, ch] = sg.fftconvolve(a[:, ch], ir, mode="same")
    return _uncols(np.clip(out, -1.0, 1.0).astype(np.float32), mono)


_REVERB_PRESETS = {
    "plate": (1.2, [29, 37, 41, 53], [0.7, 0.65, 0.6, 0.55], [5, 7], [0.5, 0.4]),
    "room": (0.6, [23, 31, 43, 47], [0.5, 0.45, 0.4, 0.35], [3, 5], [0.4, 0.3]),
    "hall": (2.2, [47, 53, 61, 71], [0.75, 0.7, 0.65, 0.6], [8, 11], [0.5, 0.45]),
    "theater": (3.5, [59, 67, 73, 83], [0.78, 0.73, 0.68, 0.63], [10, 14], [0.52, 0.45]),
    "cathedral": (5.0, [97, 103, 109, 127], [0.82, 0.78, 0.74, 0.7], [15, 19], [0.55, 0.48]),
}


def _comb(x, d, g):
    """y[n] = x[n] + g y[n - d] (pipeline.py:1065-1078) = lfilter([1], [1, 0, ..., -g])."""
    if d <= 0 or d >= len(x):
        return x
    a = np.zeros(d + 1)
    a[0], a[d] = 1.0, -g
    return sg.lfilter([1.0], a, x)


def _allpass(x, d, g):
    """y[n] = -g x[n] + x[n - d] + g y[n - d] (pipeline.py:1082-1094)."""
    if d <= 0 or d >= len(x):
        return x
    b = np.zeros(d + 1)
    a = np.zeros(d + 1)
    b[0], b[d] = -g, 1.0
    a[0], a[d] = 1.0, -g
    return sg.lfilter(b, a, x)


def _reverb_mono(x, sr, reverb_type, decay_sec, mix):
    preset = _REVERB_PRESETS.get(reverb_type, _REVERB_PRESETS["plate"])
    decay = decay_sec if decay_sec > 0 else preset[0]
    dps = 0.001 ** (1.0 / max(0.1, decay))
    n = len(x)
    x = np.asarray(x, dtype=np.float64)
    wet = np.zeros(n)
    for d_ms, g in zip(preset[1], preset[2]):
        d = min(int(sr * d_ms / 1000.0), n - 1)
        if d < 1:
            continue
        wet += _comb(x, d, g * (dps ** (d_ms / 1000.0)))
    wet /= max(len(preset[1]), 1)
    for d_ms, g in zip(preset[3], preset[4]):
        d = min(int(sr * d_ms / 1000.0), n - 1)
        if d < 1:
            continue
        wet = _allpass(wet, d, g)
    peak = np.max(np.abs(wet))
    if peak > 1e-6:
        wet = wet / min(peak, 2.0)
    return (x * (1.0 - mix) + wet * mix).astype(np.float32)


def apply_reverb(audio, sr, reverb_type="plate", decay_sec=1.2, mix=0.15, mix_mid=None, mix_side=None):
    """pipeline.py:1119-1176."""
    a, mono = _cols(audio)
    if a.shape[1] == 2 and (mix_mid is not None or mix_side is not None):
        mid = ((a[:, 0] + a[:, 1]) * 0.5).astype(np.float64)
        side = ((a[:, 0] - a[:, 1]) * 0.5).astype(np.float64)
        mm = max(0.0, min(1.0, float(mix_mid) if mix_mid is not None else mix))
        ms = max(0.0, min(1.0, float(mix_side) if mix_side is not None else mix))
        mo, so = _reverb_mono(mid, sr, reverb_type, decay_sec, mm), _reverb_mono(side, sr, reverb_type, decay_sec, ms)
        return np.stack([np.clip(mo + so, -1.0, 1.0).astype(np.float32), np.clip(mo - so, -1.0, 1.0).astype(np.float32)], axis=1)
    out = np.zeros_like(a)
    for ch in range(a.shape[1]):
        out[:, ch] = _reverb_mono(a[:, ch].astype(np.float64), sr, reverb_type, decay_sec, mix)
    return _uncols(np.clip(out, -1.0, 1.0).astype(np.float32), mono)


def compute_spectral_envelope(audio, sr, n_fft=8192):
    """pipeline.py:1527-1551."""
    mono = np.mean(audio, axis=1).astype(np.float32) if audio.ndim > 1 else np.asarray(audio, dtype=np.float32)
    hop = n_fft // 4
    window = np.hanning(n_fft).astype(np.float32)
    accum = np.zeros(n_fft // 2 + 1, dtype=np.float64)
    count = 0
    for i in range((len(mono) - n_fft) // hop + 1):
        frame = mono[i * hop: i * hop + n_fft]
        if len(frame) < n_fft:
            break
        accum += np.abs(np.fft.rfft(frame * window)) ** 2
        count += 1
    if count == 0:
        return np.ones(n_fft // 2 + 1, dtype=np.float32)
    return np.sqrt(accum / count).astype(np.float32)


def apply_reference_match(audio, sr, reference_audio, ref_sr, strength=1.0, n_fft=8192):
    """pipeline.py:1554-1612 (same sample rate)."""
    from scipy.signal import savgol_filter
    strength = float(np.clip(strength, 0.0, 1.0))
    if strength < 0.01:
        return audio
    a, mono = _cols(audio)
    src_env, ref_env = compute_spectral_envelope(a, sr, n_fft), compute_spectral_envelope(reference_audio, sr, n_fft)
    ratio = (ref_env.astype(np.float64) + 1e-8) / (src_env.astype(np.float64) + 1e-8)
    win_len = min(51, (len(ratio) // 4) * 2 + 1)
    win_len = max(5, win_len if win_len % 2 == 1 else win_len + 1)
    rs = np.clip(savgol_filter(ratio, win_len, 3), 0.1, 10.0)
    ra = np.clip(1.0 + (rs - 1.0) * strength, 0.1, 10.0)
    H = np.zeros(n_fft, dtype=np.complex128)
    H[: n_fft // 2 + 1] = ra
    H[n_fft // 2 + 1:] = ra[1: n_fft // 2][::-1]
    ir = (np.fft.ifft(H).real * np.hanning(n_fft)).astype(np.float32)
    out = np.zeros_like(a, dtype=np.float32)
    for ch in range(a.shape[1]):
        out[:, ch] = sg.fftconvolve(a[:, ch].astype(np.float64), ir.astype(np.float64), mode="same")
    return _uncols(np.clip(out, -1.0, 1.0).astype(np.float32), mono)


def _finalize(a):
    out = np.ascontiguousarray(np.clip(a, -1.0, 1.0).astype(np.float32))
    np.nan_to_num(out, copy=False, nan=0.0, posinf=1.0, neginf=-1.0)
    return out


def run_v1(audio, sr, target_lufs=-14.0, style="standard", stages=None, denoise_strength=0.0, compressor="soft_knee"):
    """``run_mastering_pipeline`` default path (pipeline.py:1800-1909; optional spectral denoise :1841-1844; no
    reference/transient).

    ``stages``: optional dict that receives a copy of the buffer after each stage.
    """
    cfg = STYLE_CONFIGS.get(style, STYLE_CONFIGS["standard"])

    def keep(name, v):
        if stages is not None:
            stages[name] = np.array(v, copy=True)
        return v

    a = keep("dc_offset", remove_dc_offset(audio))
    a = keep("peak_guard_in", remove_intersample_peaks(a, 0.5))
    if denoise_strength > 0.01:
        a = keep("spectral_denoise", apply_spectral_denoise(a, sr, strength=denoise_strength))
    a = keep("target_eq", apply_target_curve(a, sr))
    a = keep("deesser", apply_deesser(a, sr))
    a = keep("dynamics", apply_dynamics(a, sr, compressor=compressor))
    if cfg["parallel_mix"] > 0.01:
        a = keep("parallel_compress", apply_parallel_compression(a, sr, mix=cfg["parallel_mix"]))
    a = keep("normalize_lufs", normalize_lufs(a, sr, target_lufs))
    a = keep("final_spectral_balance", apply_final_spectral_balance(a, sr))
    a = keep("style_eq", apply_style_eq(a, sr, style))
    if cfg["exciter_db"] > 0.05:
        a = keep("harmonic_exciter", apply_harmonic_exciter(a, sr, cfg["exciter_db"]))
    if abs(cfg["imager_width"] - 1.0) > 0.01:
        a = keep("stereo_imager", apply_stereo_imager(a, cfg["imager_width"]))
    a = keep("peak_guard_out", remove_intersample_peaks(a, 0.5))
    a = keep("output_fade_in", apply_output_edge_fade_in(a, sr, 6.0))
    return keep("finalize_clip", _finalize(a))


def run_v2(audio, sr, target_lufs=-14.0, style="standard", stages=None, job_fade=True, compressor="soft_knee"):
    """``MasteringChain.default_chain(...).process`` (chain.py:66-98, :101-134) followed, when
    ``job_fade``, by the job function's 6 ms fade-in (routers/mastering.py:583)."""
    cfg = STYLE_CONFIGS.get(style, STYLE_CONFIGS["standard"])

    def keep(name, v):
        if stages is not None:
            stages[name] = np.array(v, copy=True)
        return v

    a = keep("dc_offset", remove_dc_offset(audio))
    a = keep("peak_guard", remove_intersample_peaks(a, 0.5))
    a = keep("target_curve", apply_target_curve(a, sr))
    a = keep("dynamics", apply_dynamics(a, sr, knee_db=6.0, crossovers_hz=V2_CROSSOVERS_HZ, compressor=compressor))
    a = keep("normalize_lufs", normalize_lufs(a, sr, float(target_lufs)))
    a = keep("final_spectral_balance", apply_final_spectral_balance(a, sr))
    a = keep("style_eq", apply_style_eq(a, sr, style))
    if abs(cfg["exciter_db"]) >= 0.05:
        a = keep("exciter", apply_harmonic_exciter(a, sr, cfg["exciter_db"], "warm", 1))
    if abs(cfg["imager_width"] - 1.0) >= 0.01:
        a = keep("imager", apply_stereo_imager(a, cfg["imager_width"]))
    a = keep("peak_guard_2", remove_intersample_peaks(a, 0.5))
    a = keep("chain_finalize_clip", _finalize(a))
    if job_fade:
        a = keep("v2_output_fade_in", apply_output_edge_fade_in(a, sr, 6.0))
    return a


# -- export -------------------------------------------------------------------------------------
def quantize_int16(samples, noise):
    """pipeline.py:880-898 with the dither noise supplied by the caller (float32, same shape)."""
    s = np.nan_to_num(np.asarray(samples, dtype=np.float32), nan=0.0, posinf=1.0, neginf=-1.0)
    s = np.clip(s, -1.0, 1.0).astype(np.float64)
    d = s * 32767.0 + np.asarray(noise, dtype=np.float32)
    d = np.nan_to_num(d, nan=0.0, posinf=32767.0, neginf=-32768.0)
    return np.clip(np.round(d), -32768, 32767).astype(np.int16)


# -- analyzers ----------------------------------------------------------------------------------
def dither_noise_shaped(uniform, kind):
    """_dither_noise_ns_e / _dither_noise_ns_itu (pipeline.py:835-877) from the float32 uniforms the reference draws."""
    white = (2.0 * np.asarray(uniform, dtype=np.float32) - 1.0).astype(np.float32)
    w2 = white.reshape(white.shape[0], -1)
    out = np.empty_like(w2)
    if kind == "ns_e":
        out[0] = w2[0]
        for i in range(1, w2.shape[0]):
            out[i] = w2[i] - w2[i - 1] + 0.99 * out[i - 1]
    else:
        for c in range(w2.shape[1]):
            out[:, c] = sg.lfilter(np.array([1.0, -2.0, 1.0]), np.array([1.0, -1.96, 0.9604]), w2[:, c])
    return (out * 0.9).astype(np.float32).reshape(white.shape)


def auto_blank_end(samples, sr, threshold_dbfs=-60.0, min_silence_sec=0.5):
    """pipeline.py:900-918: cut after the last frame whose peak over the channels exceeds the threshold, keeping
    ``min_silence_sec`` of tail."""
    if samples.size == 0 or min_silence_sec <= 0:
        return samples
    thr = 10 ** (threshold_dbfs / 20.0)
    n_silence = int(sr * min_silence_sec)
    if n_silence <= 0:
        return samples
    peak = np.max(np.abs(samples), axis=1) if samples.ndim > 1 else np.abs(samples)
    above = np.nonzero(peak > thr)[0]
    if above.size == 0:
        return samples
    return samples[: min(samples.shape[0], int(above[-1]) + 1 + n_silence)]


def export_prepare(samples, sr, auto_blank_sec=0.0):
    """Head of export_audio (pipeline.py:972-977): float32, 2-D, clip, optional trailing-silence cut at -50 dBFS."""
    a = np.asarray(samples, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    a = np.clip(a, -1.0, 1.0)
    return auto_blank_end(a, sr, -50.0, auto_blank_sec) if auto_blank_sec > 0 else a


def quantize_pcm24(samples):
    """What libsndfile does with float32 written as PCM_24 (FLAC branch, pipeline.py:981-985; flac.c f2flac24_array with
    normalisation on): lrintf(x * 0x7FFFFF) in float32.  PARITY UNPINNED: libsndfile itself is not available here."""
    a = np.clip(np.nan_to_num(np.asarray(samples, dtype=np.float32), nan=0.0), -1.0, 1.0)
    return np.rint(a * np.float32(8388607.0)).astype(np.int32)


def true_peak_dbfs(audio, sr=None):
    """routers/tools.py:44-54: 4x ``resample_poly`` then sample peak in dBFS."""
    audio = np.asarray(audio)
    if audio.size == 0:
        return -120.0
    x, _ = _cols(audio.astype(np.float64))
    pk = max(float(np.max(np.abs(sg.resample_poly(x[:, c], 4, 1)))) for c in range(x.shape[1]))
    return float(20 * np.log10(max(pk, 1e-12)))


def true_peak_fir():
    """The 81-tap filter ``resample_poly(x, 4, 1)`` builds (scipy ``_upfirdn``/``firwin`` defaults):
    ``firwin(81, 1/4, window=('kaiser', 5.0)) * 4``; output m = sum_k h[k] * xup[m + 40 - k]."""
    return sg.firwin(81, 0.25, window=("kaiser", 5.0)) * 4.0


def true_peak_explicit(audio):
    """Same number as :func:`true_peak_dbfs` by zero-stuffing + direct convolution (documents the
    polyphase structure the CUDA kernel uses)."""
    x, _ = _cols(np.asarray(audio, dtype=np.float64))
    h = true_peak_fir()
    pk = 0.0
    for c in range(x.shape[1]):
        up = np.zeros(4 * x.shape[0])
        up[::4] = x[:, c]
        y = np.convolve(up, h)[40:40 + 4 * x.shape[0]]
        pk = max(pk, float(np.max(np.abs(y))))
    return float(20 * np.log10(max(pk, 1e-12)))


def compute_spectrum_bars(audio, sr, n_fft=4096, n_bars=64, min_hz=20.0, max_hz=20000.0):
    """pipeline.py:700-739."""
    audio = np.asarray(audio)
    if audio.size < n_fft:
        return [-80.0] * n_bars
    mono = np.mean(audio, axis=1) if audio.ndim > 1 else np.asarray(audio, dtype=np.float64)
    n = len(mono)
    s = max(0, n // 2 - n_fft // 2)
    frame = mono[s:s + n_fft].copy()
    frame *= np.hanning(n_fft)
    mag = np.abs(np.fft.rfft(frame)) * (2.0 / n_fft)
    nyq = sr / 2.0
    bars = []
    for b in range(n_bars):
        f0 = min_hz * (max_hz / min_hz) ** (b / max(n_bars - 1, 1))
        f1 = min_hz * (max_hz / min_hz) ** ((b + 1) / max(n_bars - 1, 1))
        k0 = max(0, int((f0 / nyq) * (n_fft // 2)))
        k1 = min(len(mag) - 1, int(np.ceil((f1 / nyq) * (n_fft // 2))))
        pk = 1e-12 if k0 > k1 else float(np.max(mag[k0:k1 + 1]))
        bars.append(round(20.0 * np.log10(max(pk, 1e-12)), 2))
    return bars


def measure_stereo_correlation(audio):
    """pipeline.py:766-791."""
    if audio.ndim != 2 or audio.shape[1] != 2 or audio.size < 4:
        return None
    l = np.asarray(audio[:, 0], dtype=np.float64)
    r = np.asarray(audio[:, 1], dtype=np.float64)
    n = l.size
    sl, sr_, slr, sll, srr = np.sum(l), np.sum(r), np.sum(l * r), np.sum(l * l), np.sum(r * r)
    if np.sqrt(max(sll * srr, 0.0)) < 1e-20:
        return None
    den = np.sqrt(max(n * sll - sl * sl, 0.0)) * np.sqrt(max(n * srr - sr_ * sr_, 0.0))
    if den < 1e-20:
        return 0.0
    return float(np.clip((n * slr - sl * sr_) / den, -1.0, 1.0))


def compute_vectorscope_points(audio, max_points=1000):
    """pipeline.py:742-763."""
    if audio.ndim != 2 or audio.shape[1] != 2 or audio.size < 4:
        return []
    n = audio.shape[0]
    step = max(1, n // max_points)
    idx = np.arange(0, n, step)[:max_points]
    l = np.clip(audio[idx, 0].astype(np.float64), -1.0, 1.0)
    r = np.clip(audio[idx, 1].astype(np.float64), -1.0, 1.0)
    return [[round(float(a), 5), round(float(b), 5)] for a, b in zip(l, r)]


def validate_not_silent(mastered):
    """pipeline.py:939-962 (returns the failure reason or None instead of raising)."""
    if mastered.size == 0:
        return "empty_buffer"
    if not np.all(np.isfinite(mastered)):
        return "nan_or_inf"
    if float(np.max(np.abs(mastered))) < 1e-5:
        return "near_silence_peak"
    return None
