"""Drop-in for the reference's ``backend/app/pipeline.py`` function surface, hosted on the GPU.

Same names, positional/keyword signatures, defaults, return shapes and error conventions as the
reference functions cited in each docstring; arrays are numpy ``(n,)`` mono or ``(n, ch)``
interleaved float32 at this edge, planar float32 on the device.  Every function launches CUDA
kernels through ``libmm_b200.so``; there is no CPU implementation behind any of them.  The
``*_batch`` helpers at the bottom are additive (the reference has no batching).
"""
from __future__ import annotations

import ctypes as C
import io
from typing import Callable, Optional

import numpy as np

from . import _lib, wavio
from .engine import Batch, Engine, get_engine, style_struct

# ---- presets / constants (backend/app/pipeline.py:56-110) -----------------------------------------
PRESET_LUFS = {"spotify": -14.0, "youtube": -14.0, "apple": -16.0, "club": -9.0, "broadcast": -24.0}

STYLE_CONFIGS = {
    "standard": {"lufs": -14.0, "sub": 0.0, "bass": 0.0, "mids": 0.0, "presence": 0.0, "air": 0.0, "comp_mult": 1.0,
                 "exciter_db": 0.0, "imager_width": 1.0, "parallel_mix": 0.0},
    "edm": {"lufs": -9.0, "sub": 1.8, "bass": 0.9, "mids": -0.3, "presence": 0.6, "air": 0.9, "comp_mult": 1.3,
            "exciter_db": 0.6, "imager_width": 1.25, "parallel_mix": 0.3},
    "hiphop": {"lufs": -13.0, "sub": 1.4, "bass": 0.7, "mids": 0.5, "presence": 0.3, "air": 0.2, "comp_mult": 1.2,
               "exciter_db": 0.3, "imager_width": 1.1, "parallel_mix": 0.35},
    "classical": {"lufs": -18.0, "sub": -0.5, "bass": 0.0, "mids": 0.0, "presence": 0.3, "air": 0.6, "comp_mult": 0.45,
                  "exciter_db": 0.0, "imager_width": 1.05, "parallel_mix": 0.0},
    "podcast": {"lufs": -16.0, "sub": -1.2, "bass": -0.4, "mids": 0.9, "presence": 0.7, "air": 0.0, "comp_mult": 1.1,
                "exciter_db": 0.0, "imager_width": 1.0, "parallel_mix": 0.2},
    "lofi": {"lufs": -18.0, "sub": 0.4, "bass": 0.6, "mids": -0.6, "presence": -1.0, "air": -1.8, "comp_mult": 0.65,
             "exciter_db": 0.2, "imager_width": 0.9, "parallel_mix": 0.0},
    "house_basic": {"lufs": -10.0, "sub": 1.8, "bass": 0.9, "mids": -0.5, "presence": 0.8, "air": 1.0, "comp_mult": 1.35,
                    "exciter_db": 0.8, "imager_width": 1.3, "parallel_mix": 0.3},
    "dry_vocal": {"lufs": -14.0, "sub": 0.0, "bass": 0.0, "mids": 0.0, "presence": 0.0, "air": 0.0, "comp_mult": 1.0,
                  "exciter_db": 0.0, "imager_width": 1.0, "parallel_mix": 0.0},
}

TRUE_PEAK_LIMIT_DB = -1.5
MULTIBAND_CROSSOVERS_HZ = (214.0, 3500.0, 10000.0)
MAXIMIZER_THRESHOLD_DB = -2.5
MAXIMIZER_MARGIN_DB = -0.3
FINAL_TRIM_DB = 0.5
MULTIBAND_CONFIG = [(-7.2, 1.0, -7.2, 1.5), (-18.5, 2.2, -18.5, 1.8), (-17.0, 1.55, -17.0, 1.65), (-15.0, 1.35, -15.0, 1.2)]
# (strength, noise_percentile) -- backend/app/pipeline.py:1439-1446
DENOISE_PRESETS = {"vocal": (0.15, 25.0), "light": (0.20, 22.0), "medium": (0.5, 15.0), "aggressive": (0.75, 10.0),
                   "tape_hiss": (0.25, 22.0), "room_tone": (0.40, 18.0)}

_EXCITER_MODES = {"warm": 0, "tape": 1, "tube": 2, "transistor": 3, "digital": 4}

# Which compressor apply_multiband_dynamics runs per band.  The reference decides by whether ``pedalboard`` imports
# (backend/app/pipeline.py:442-446): with it, the envelope compressor (:373-411); without it, the memoryless soft-knee branch
# (:466-474).  "soft_knee" is the default here because it is the branch pinned against the unmodified reference; "envelope"
# (PARITY UNPINNED: restated JUCE arithmetic, csrc/bandcomp.cu) is what a production install of the reference runs.  Set
# ``MM_COMPRESSOR=envelope`` (or assign COMPRESSOR_MODE) to make it the default of every entry point, or pass ``compressor=``.
import os as _os
COMPRESSOR_MODE = _os.environ.get("MM_COMPRESSOR", "soft_knee")


def _compressor_id(compressor=None) -> int:
    mode = COMPRESSOR_MODE if compressor is None else compressor
    if mode in ("soft_knee", "numpy", 0):
        return _lib.COMPRESSOR_SOFT_KNEE
    if mode in ("envelope", "pedalboard", 1):
        return _lib.COMPRESSOR_ENVELOPE
    raise ValueError(f"unknown compressor mode {mode!r} (soft_knee | envelope)")


# ---- host <-> device edge -------------------------------------------------------------------------
# elementwise reference functions return their input's shape; the filter stages squeeze an (n, 1) input to (n,) ("[:, 0]")
_SHAPE_PRESERVING = {"remove_dc_offset", "remove_intersample_peaks", "fade_in", "apply_maximizer", "apply_parallel_compression",
                     "normalize_lufs", "finalize_clip", "apply_rumble_filter"}


# Resident scope (internal): run_mastering_pipeline's stage-by-stage path strings ~16 of the stage functions below together.  Each
# of them takes and returns a host array -- an upload and a download of the whole track per stage (~6 ms each for a 3-minute track,
# against ~1 ms of kernels).  Inside a `_ResidentScope` `_down` hands out a PLACEHOLDER instead: an uninitialised host array of the
# right shape (untouched virtual memory) whose buffer address is registered against the device batch that holds the samples, and
# `_up` of a registered array returns that batch without any copy.  Every stage function only looks at an input's shape on the
# host and reaches the samples through `_up`, so the chain runs device-resident through the very same functions;
# `_materialize` downloads the final result.  Placeholders never leave the scope (the registry keeps them alive, so no other
# array can take over a registered address while the scope is open).
import threading as _threading

_resident = _threading.local()


class _ResidentScope:
    def __enter__(self):
        self.prev = getattr(_resident, "reg", None)
        _resident.reg = {}
        return self

    def __exit__(self, *exc):
        _resident.reg = self.prev
        return False


def _resident_hit(a: np.ndarray):
    reg = getattr(_resident, "reg", None)
    if not reg or a.size == 0:
        return None
    hit = reg.get(a.__array_interface__["data"][0])
    if hit is None:
        return None
    b = hit[1]
    if a.dtype != np.float32 or a.size != b.n * b.channels or a.shape[0] != b.n or not a.flags.c_contiguous:
        return None
    return b


def _materialize(a):
    """Placeholder -> the real samples (anything else is returned as it is)."""
    arr = np.asarray(a) if isinstance(a, np.ndarray) else None
    b = _resident_hit(arr) if arr is not None else None
    if b is None:
        return a
    out = get_engine().download(b)[0]
    return out[:, 0] if arr.ndim == 1 else out


# additive, opt-in for an integrator who strings stage functions together on one track (the v2 job function calls up to ten PRO
# stages between the chain and the export, routers/mastering.py:443-609): `with resident_scope(): ...; out = materialize(out)`.
# Inside the scope every array a stage returns is a placeholder -- it must only be passed to other functions of this module (or
# to `materialize`), never read on the host.
resident_scope = _ResidentScope
materialize = _materialize


def _up(audio, sr, eng: Optional[Engine] = None, squeeze: bool = True):
    eng = eng or get_engine()
    a = np.asarray(audio)
    mono = a.ndim == 1 or (squeeze and a.shape[1] == 1)
    b = _resident_hit(a) if isinstance(audio, np.ndarray) else None
    if b is not None:                                      # samples already on the device (resident scope)
        return eng, (b if b.sr == int(sr) else Batch(b.t, b.tracks, b.channels, b.n, int(sr))), mono
    a2 = np.ascontiguousarray(a.reshape(a.shape[0], -1), dtype=np.float32)
    if a2.shape[1] not in (1, 2):
        raise ValueError("mm_b200 supports mono or stereo audio")
    return eng, eng.upload([a2], int(sr)), mono


def _down(eng: Engine, b: Batch, mono: bool) -> np.ndarray:
    reg = getattr(_resident, "reg", None)
    if reg is not None and b.tracks == 1:
        ph = np.empty((b.n, b.channels), np.float32)     # never read: its address stands for the device batch
        reg[ph.__array_interface__["data"][0]] = (ph, b)
        return ph[:, 0] if mono else ph
    out = eng.download(b)[0]
    return out[:, 0] if mono else out


def _stage(name, audio, sr, *args):
    eng, b, mono = _up(audio, sr, squeeze=name not in _SHAPE_PRESERVING)
    return _down(eng, eng.stage(name, b, *args, out=b), mono)


# ---- stages -----------------------------------------------------------------------------------------
def remove_dc_offset(audio: np.ndarray) -> np.ndarray:
    """backend/app/pipeline.py:134-138."""
    return _stage("remove_dc_offset", audio, 44100)


def remove_intersample_peaks(audio: np.ndarray, headroom_db: float = 0.5) -> np.ndarray:
    """backend/app/pipeline.py:141-149."""
    return _stage("remove_intersample_peaks", audio, 44100, C.c_double(headroom_db))


def apply_output_edge_fade_in(audio: np.ndarray, sr: int, fade_ms: float = 6.0) -> np.ndarray:
    """backend/app/pipeline.py:152-167."""
    if fade_ms <= 0 or sr <= 0 or np.size(audio) == 0:
        return audio
    return _stage("fade_in", audio, sr, C.c_double(fade_ms))


def apply_target_curve(audio: np.ndarray, sr: int, phase_mode: str = "minimum", eq_ms: bool = False) -> np.ndarray:
    """backend/app/pipeline.py:238-273: zero-phase IIR path, or ``phase_mode="linear_phase"`` (:187-235, 4096-tap FIR)."""
    a = np.asarray(audio)
    ms = bool(eq_ms) and a.ndim == 2 and a.shape[1] == 2
    if phase_mode == "linear_phase":
        return _stage("apply_target_curve_linear_phase", audio, sr, 1 if ms else 0)
    return _stage("apply_target_curve", audio, sr, 1 if ms else 0)


def apply_target_curve_linear_phase(audio: np.ndarray, sr: int, n_fft: int = 4096) -> np.ndarray:
    """backend/app/pipeline.py:220-235."""
    if n_fft != 4096:
        raise NotImplementedError("the linear-phase target curve is built for the reference's n_fft = 4096")
    return _stage("apply_target_curve_linear_phase", audio, sr, 0)


def apply_deesser(audio: np.ndarray, sr: int, threshold_db: float = -6.0, ratio: float = 3.0, freq_lo: float = 5000.0,
                  freq_hi: float = 9000.0, attack_ms: float = 4.0, release_ms: float = 85.0) -> np.ndarray:
    """backend/app/pipeline.py:1200-1264."""
    return _stage("apply_deesser", audio, sr, *(C.c_double(v) for v in (threshold_db, ratio, freq_lo, freq_hi, attack_ms, release_ms)))


def apply_dynamics(samples: np.ndarray, sr: int, knee_db: float = 6.0, crossovers_hz=None, band_ratios=None,
                   max_upward_boost_db: float = 12.0, compressor: Optional[str] = None) -> np.ndarray:
    """backend/app/pipeline.py:610-641.  ``compressor`` (additive): "soft_knee" = the numpy branch (:466-474), "envelope" = the
    pedalboard-style envelope compressor (:373-411, parity unpinned); default ``COMPRESSOR_MODE``."""
    cx = _lib.darr(crossovers_hz) if crossovers_hz is not None and len(crossovers_hz) == 3 else None
    br = _lib.darr(band_ratios) if band_ratios is not None and len(band_ratios) == 4 else None
    return _stage("apply_dynamics_mode", samples, sr, C.c_double(knee_db), cx, br, C.c_double(max_upward_boost_db), 0,
                  _compressor_id(compressor))


def apply_multiband_dynamics(samples: np.ndarray, sr: int, knee_db: float = 6.0, crossovers_hz=None, band_ratios=None,
                             max_upward_boost_db: float = 12.0, compressor: Optional[str] = None) -> np.ndarray:
    """backend/app/pipeline.py:414-481: the four-band split / compress / limit / gain / sum on its own -- apply_dynamics without
    the maximizer and limiter behind it.  ``compressor`` as in apply_dynamics."""
    cx = _lib.darr(crossovers_hz) if crossovers_hz is not None and len(crossovers_hz) == 3 else None
    br = _lib.darr(band_ratios) if band_ratios is not None and len(band_ratios) == 4 else None
    return _stage("apply_dynamics_mode", samples, sr, C.c_double(knee_db), cx, br, C.c_double(max_upward_boost_db), 1,
                  _compressor_id(compressor))


def apply_maximizer(audio: np.ndarray) -> np.ndarray:
    """backend/app/pipeline.py:484-492 (module constants -2.5 / -0.3 dB)."""
    return _stage("apply_maximizer", audio, 44100)


def apply_maximizer_lookahead(audio: np.ndarray, sr: int, lookahead_ms: float = 6.0) -> np.ndarray:
    """backend/app/pipeline.py:548-573: the first ``lookahead_ms`` pass unlimited, the rest is the maximizer of the delayed
    signal, with a 2 ms cross-fade before the seam."""
    eng, b, mono = _up(audio, sr)
    return _down(eng, eng.stage("apply_maximizer_lookahead", b, C.c_double(float(lookahead_ms))), mono)


def apply_parallel_compression(audio: np.ndarray, sr: int, mix: float = 0.3, ratio: float = 8.0,
                               threshold_db: float = -20.0) -> np.ndarray:
    """backend/app/pipeline.py:1771-1797."""
    if float(np.clip(mix, 0.0, 1.0)) < 0.01:
        return audio
    return _stage("apply_parallel_compression", audio, sr, C.c_double(mix), C.c_double(ratio), C.c_double(threshold_db))


def measure_lufs(audio: np.ndarray, sr: int) -> float:
    """backend/app/pipeline.py:658-664: NaN when the meter cannot run (shorter than one 400 ms block)."""
    a = np.asarray(audio)
    if a.size == 0 or a.shape[0] < 0.4 * sr:
        return float("nan")
    eng, b, _ = _up(a, sr)       # device/driver errors propagate: no silent fallback
    return float(eng.measure_lufs(b)[0])


def normalize_lufs(audio: np.ndarray, sr: int, target_lufs: float) -> np.ndarray:
    """backend/app/pipeline.py:644-655: input returned unchanged when the meter cannot run."""
    a = np.asarray(audio)
    if a.shape[0] < 0.4 * sr:
        return audio
    return _stage("normalize_lufs", audio, sr, _lib.darr([target_lufs]))


def apply_final_spectral_balance(audio: np.ndarray, sr: int) -> np.ndarray:
    """backend/app/pipeline.py:576-607."""
    return _stage("apply_final_spectral_balance", audio, sr)


def apply_style_eq(audio: np.ndarray, sr: int, style: str = "standard") -> np.ndarray:
    """backend/app/pipeline.py:1401-1434."""
    cfg = STYLE_CONFIGS.get(style, STYLE_CONFIGS["standard"])
    return _stage("apply_style_eq", audio, sr, _lib.darr([cfg[k] for k in ("sub", "bass", "mids", "presence", "air")]))


def _exciter_dev(eng, b, exciter_db, mode, oversample):
    """apply_harmonic_exciter on a device batch (returns a batch; may be ``b`` itself when bypassed)."""
    if abs(exciter_db) < 0.05:
        return b
    os_ = max(1, min(4, int(oversample)))
    if os_ == 1:
        return eng.stage("apply_harmonic_exciter", b, C.c_double(exciter_db), _EXCITER_MODES.get(mode, 0))
    work = eng.fft_resample(b, b.n * os_, b.sr * os_)
    work = eng.stage("apply_harmonic_exciter", work, C.c_double(exciter_db), _EXCITER_MODES.get(mode, 0), out=work)
    return eng.fft_resample(work, b.n, b.sr)


def apply_harmonic_exciter(audio: np.ndarray, sr: int, exciter_db: float = 0.0, mode: str = "warm", oversample: int = 1) -> np.ndarray:
    """backend/app/pipeline.py:1267-1326.  ``oversample`` 2..4: whole-signal FFT up-sampling (scipy.signal.resample
    semantics, csrc/bigfft.cu), the side chain at the high rate, FFT down-sampling -- three device calls, no host trip."""
    if abs(exciter_db) < 0.05:
        return audio
    eng, b, mono = _up(audio, sr)
    return _down(eng, _exciter_dev(eng, b, exciter_db, mode, oversample), mono)


def fft_resample(audio: np.ndarray, num: int) -> np.ndarray:
    """``scipy.signal.resample(audio, num, axis=0).astype(float32)`` on the device."""
    eng, b, mono = _up(audio, 44100, squeeze=False)
    if int(num) == b.n:
        return _down(eng, b, mono)
    return _down(eng, eng.fft_resample(b, int(num)), mono)


def resample_audio(audio: np.ndarray, sr: int, target_sr: int) -> np.ndarray:
    """backend/app/pipeline.py:920-936."""
    if target_sr <= 0 or sr <= 0:
        raise ValueError("Invalid sample rate")
    if target_sr == sr:
        return np.asarray(audio, dtype=np.float32)
    a = np.asarray(audio)
    return fft_resample(a, int(round(a.shape[0] * target_sr / sr)))


def _imager_dev(eng, b, width=1.0, stereoize_delay_ms=0.0, stereoize_mix=0.12, band_widths=None, crossovers_hz=None):
    """apply_stereo_imager on a device batch (backend/app/pipeline.py:1339-1398); returns a batch."""
    if b.channels != 2:
        return b
    sr = b.sr
    four = band_widths is not None and len(band_widths) == 4 and sr and sr > 0
    haas = bool(stereoize_delay_ms and stereoize_delay_ms > 0 and sr and sr > 0 and stereoize_mix > 0 and
                min(int(sr * stereoize_delay_ms / 1000.0), b.n - 1) > 0)
    if four:
        cx = _lib.darr([float(v) for v in crossovers_hz]) if crossovers_hz is not None and len(crossovers_hz) == 3 else None
        wide = eng.stage("apply_stereo_imager_4band", b, _lib.darr([float(w) for w in band_widths]), cx)
        if not haas:
            return wide
        # the cross-delay then acts on the merged pair as it is (width = NaN: no mid/side step)
        return eng.stage("apply_stereoize", wide, C.c_double(float("nan")), C.c_double(stereoize_delay_ms), C.c_double(stereoize_mix))
    if haas:                                  # the delayed tap reads behind the writer: not in place
        return eng.stage("apply_stereoize", b, C.c_double(width), C.c_double(stereoize_delay_ms), C.c_double(stereoize_mix))
    return eng.stage("apply_stereo_imager", b, C.c_double(width))


def apply_stereo_imager(audio: np.ndarray, width: float = 1.0, stereoize_delay_ms: float = 0.0, stereoize_mix: float = 0.12,
                        sr=None, band_widths=None, crossovers_hz=None) -> np.ndarray:
    """backend/app/pipeline.py:1339-1398: plain width, 4-band width (``band_widths``) and the Haas "stereoize"
    cross-delay."""
    a = np.asarray(audio)
    if a.ndim == 1 or a.shape[1] == 1:
        return audio
    if not (sr and sr > 0):                   # without a rate only the plain width mode exists (:1360, :1385)
        return _stage("apply_stereo_imager", audio, 44100, C.c_double(width))
    eng, b, mono = _up(audio, sr)
    return _down(eng, _imager_dev(eng, b, width, stereoize_delay_ms, stereoize_mix, band_widths, crossovers_hz), mono)


def apply_transient_designer(audio: np.ndarray, sr: int, attack_gain: float = 1.0, sustain_gain: float = 1.0) -> np.ndarray:
    """backend/app/pipeline.py:1736-1768."""
    ag, sg_ = float(np.clip(attack_gain, 0.1, 3.0)), float(np.clip(sustain_gain, 0.1, 3.0))
    if abs(ag - 1.0) < 0.02 and abs(sg_ - 1.0) < 0.02:
        return audio                      # the reference returns the very same object
    return _stage("apply_transient_designer", audio, sr, C.c_double(ag), C.c_double(sg_))


def apply_maximizer_transient_aware(audio: np.ndarray, sr: int, sensitivity: float = 0.5) -> np.ndarray:
    """backend/app/pipeline.py:521-545."""
    return _stage("apply_maximizer_transient_aware", audio, sr, C.c_double(float(sensitivity)))


# backend/app/pipeline.py:1616-1625
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


DYNEQ_CLASS_NAMES = ("stable", "overflow->identity", "lfilter-fallback", "marginal-filtfilt", "skipped")


def apply_dynamic_eq(audio: np.ndarray, sr: int, bands=None, strict: bool = False, report: Optional[list] = None) -> np.ndarray:
    """backend/app/pipeline.py:1628-1700: per band a zero-phase ``iirpeak`` section, the attack/release follower of the band,
    a downward gain above the threshold, ``x - band + band * g``; bands act one after the other.

    The reference hands ``scipy.signal.iirpeak`` a *bandwidth* in its ``Q`` slot (:1661-1663), so a band is a stable section
    only for ``q < 1`` (roughly), and most bands of ``DYNAMIC_EQ_MASTERING_BANDS`` are unstable or degenerate.  What the
    reference then returns is reproduced per band class (``csrc/deesser.cu`` ``st_dynamic_eq``): an unstable section whose
    forward pass must overflow is zeroed by the reference's ``nan_to_num`` and is therefore the identity; the degenerate
    ``[1, 0, -1] / [1, ~0, ~-1]`` sections (q = 1, or the bandwidth clipped at w0 = 0.5) come back as ``x`` (``lfilter``
    fallback) or ``x - const`` (``filtfilt`` with poles at +-1).  An unstable band whose overflow is not certain (a few
    hundred samples, a constant signal) is passed through with a warning -- ``strict=True`` raises ``MMError`` instead.
    ``report`` (additive): a list that receives one class name per band handed to the device."""
    if bands is None:
        bands = DYNAMIC_EQ_MASTERING_BANDS
    nyq = sr / 2.0
    rows = []
    for band in bands:
        freq = float(band.get("freq", 1000))
        q = float(band.get("q", 1.4))
        if freq <= 0 or freq >= nyq * 0.98:
            continue
        w0 = float(np.clip(freq / nyq, 0.001, 0.98))
        bw = float(np.clip(w0 / max(q, 0.1), 0.001, 0.5))
        rows.append([w0, bw, float(band.get("threshold_db", -12)), float(band.get("ratio", 3.0)), float(band.get("attack_ms", 5)),
                     float(band.get("release_ms", 80)), float(band.get("max_cut_db", -6))])
    flat = _lib.darr([v for r in rows for v in r]) if rows else None
    classes = (C.c_int32 * max(len(rows), 1))()
    out = _stage("apply_dynamic_eq2", audio, sr, len(rows), flat, _lib.DYNEQ_STRICT if strict else 0, classes)
    kinds = [DYNEQ_CLASS_NAMES[classes[i]] for i in range(len(rows))]
    if report is not None:
        report.extend(kinds)
    if "skipped" in kinds:
        import logging
        logging.getLogger(__name__).warning(
            "apply_dynamic_eq: %d unstable band(s) passed through (overflow of the reference's filter not certain for %d frames)",
            kinds.count("skipped"), int(np.asarray(audio).shape[0]))
    return out


HIGH_FREQ_TRIM_CROSSOVER_HZ = 5000.0
HIGH_FREQ_TRIM_GAIN = 0.9


def apply_high_freq_trim(audio: np.ndarray, sr: int, crossover_hz: float = HIGH_FREQ_TRIM_CROSSOVER_HZ,
                         high_gain: float = HIGH_FREQ_TRIM_GAIN) -> np.ndarray:
    """backend/app/pipeline.py:1705-1733."""
    if abs(high_gain - 1.0) < 0.001:
        return audio
    return _stage("apply_high_freq_trim", audio, sr, C.c_double(crossover_hz), C.c_double(high_gain))


def apply_spectral_denoise(audio: np.ndarray, sr: int, strength: float = 0.5, noise_percentile: float = 15.0) -> np.ndarray:
    """backend/app/pipeline.py:1472-1524: STFT (2048 / 512, scipy.signal.stft conventions) Wiener gain against a per-bin
    percentile noise floor, all on the device (csrc/denoise.cu).  Inputs shorter than one 2048-sample segment raise
    ``ValueError`` as the reference's scipy call does."""
    strength = float(np.clip(strength, 0.0, 1.0))
    if strength < 0.01:
        return audio
    if not 0.0 <= float(noise_percentile) <= 100.0:
        raise ValueError("Percentiles must be in the range [0, 100]")
    if np.asarray(audio).shape[0] < 2048:
        raise ValueError("noverlap must be less than nperseg.")
    return _stage("apply_spectral_denoise", audio, sr, C.c_double(strength), C.c_double(float(noise_percentile)))


def compute_spectral_envelope(audio: np.ndarray, sr: int, n_fft: int = 8192) -> np.ndarray:
    """backend/app/pipeline.py:1527-1551: averaged RMS spectrum (8192-sample Hann frames, hop 2048) -> float32 (4097,)."""
    if n_fft != 8192:
        raise NotImplementedError("the spectral-envelope kernel is built for the reference's n_fft = 8192")
    a = np.asarray(audio)
    if a.shape[0] < n_fft:
        return np.ones(n_fft // 2 + 1, dtype=np.float32)
    import torch
    eng, b, _ = _up(audio, sr)
    with torch.cuda.stream(eng.stream):
        env = torch.empty(b.tracks * (n_fft // 2 + 1), dtype=torch.float32, device=eng.tdev)
        g = b.geom
        _lib.check(eng.lib.mm_dev_spectral_envelope(eng.ctx, C.byref(g), b.ptr, C.c_void_p(env.data_ptr())))
        eng.sync()
        return env.cpu().numpy()[: n_fft // 2 + 1].copy()


def reference_match_ir(src_env: np.ndarray, ref_env: np.ndarray, strength: float, n_fft: int = 8192) -> np.ndarray:
    """The design half of apply_reference_match (backend/app/pipeline.py:1590-1604): ratio curve, Savitzky-Golay smoothing,
    strength, clipping, Hann-windowed impulse response.  4097 numbers on the host, exactly the reference's numpy/scipy calls."""
    from scipy.signal import savgol_filter
    eps = 1e-8
    ratio = (ref_env.astype(np.float64) + eps) / (src_env.astype(np.float64) + eps)
    win_len = min(51, (len(ratio) // 4) * 2 + 1)
    win_len = max(5, win_len if win_len % 2 == 1 else win_len + 1)
    ratio_smooth = np.clip(savgol_filter(ratio, win_len, 3), 0.1, 10.0)
    ratio_applied = np.clip(1.0 + (ratio_smooth - 1.0) * strength, 0.1, 10.0)
    n_bins = n_fft // 2 + 1
    H = np.zeros(n_fft, dtype=np.complex128)
    H[:n_bins] = ratio_applied
    H[n_bins:] = ratio_applied[1: n_fft // 2][::-1]
    return (np.fft.ifft(H).real * np.hanning(n_fft)).astype(np.float32)


def apply_reference_match(audio: np.ndarray, sr: int, reference_audio: np.ndarray, ref_sr: int, strength: float = 1.0,
                          n_fft: int = 8192) -> np.ndarray:
    """backend/app/pipeline.py:1554-1612: both spectral envelopes and the 8192-tap convolution run on the device, the filter
    design on the host.  A reference at another sample rate is mixed to mono and FFT-resampled first (:1581-1584)."""
    strength = float(np.clip(strength, 0.0, 1.0))
    if strength < 0.01:
        return audio
    if ref_sr != sr:
        # the reference mixes to mono, then resamples; resampling is linear, so both channels are resampled on the device (one
        # complex transform, the cost of a mono row) and the envelope kernel forms the channel mean itself -- no host arithmetic
        ref = np.asarray(reference_audio)
        reference_audio = fft_resample(ref, int(ref.shape[0] * sr / ref_sr))
    src_env = compute_spectral_envelope(audio, sr, n_fft)
    ref_env = compute_spectral_envelope(reference_audio, sr, n_fft)
    ir = np.ascontiguousarray(reference_match_ir(src_env, ref_env, strength, n_fft))
    return _stage("fir_same", audio, sr, ir.ctypes.data_as(C.c_void_p), int(ir.shape[0]), 1)


_REVERB_TYPES = {"plate": 0, "room": 1, "hall": 2, "theater": 3, "cathedral": 4}


def apply_reverb(audio: np.ndarray, sr: int, reverb_type: str = "plate", decay_sec: float = 1.2, mix: float = 0.15,
                 mix_mid: Optional[float] = None, mix_side: Optional[float] = None) -> np.ndarray:
    """backend/app/pipeline.py:1119-1176 (Schroeder comb + allpass; optional separate mid / side mixes)."""
    a = np.asarray(audio)
    use_ms = a.ndim == 2 and a.shape[1] == 2 and (mix_mid is not None or mix_side is not None)
    m_mid = float(mix_mid) if mix_mid is not None else float(mix)
    m_side = float(mix_side) if mix_side is not None else float(mix)
    return _stage("apply_reverb", audio, sr, _REVERB_TYPES.get(reverb_type, 0), C.c_double(float(decay_sec)), C.c_double(float(mix)),
                  1 if use_ms else 0, C.c_double(m_mid), C.c_double(m_side))


def apply_rumble_filter(audio: np.ndarray, sr: int, cutoff_hz: float = 80.0) -> np.ndarray:
    """backend/app/pipeline.py:1449-1469."""
    return _stage("apply_rumble_filter", audio, sr, C.c_double(cutoff_hz))


# user-facing texts of the reference (pipeline.py:948-961); its tests match "тишину" / "Отключите"
_SILENT_MSG = ("Обработка дала тишину. Отключите часть доп. настроек (Spectral Denoiser, De-esser, "
               "Transient Designer, Parallel Compression, Dynamic EQ) и попробуйте снова.")
_NONFINITE_MSG = "Обработка дала недопустимые значения (NaN/Inf). Отключите Dynamic EQ или другие доп. модули и попробуйте снова."


def validate_mastered_not_silent(mastered: np.ndarray, *, trace_ctx=None, trace_sr: int = 44100) -> None:
    """backend/app/pipeline.py:939-962: ValueError on an empty, non-finite or silent (< 1e-5 peak) result
    (the reference's tests match "тишину" / "Отключите", backend/tests/test_pipeline.py:369-384)."""
    m = np.asarray(mastered)
    if m.size == 0:
        raise ValueError(_SILENT_MSG)
    from .mastering_trace import batch_metrics
    eng, b, _ = _up(m, trace_sr)
    met = batch_metrics(eng, b)[0]            # one device reduction: peak over the finite samples, NaN and Inf counts
    if met["nan_count"] or met["inf_count"]:
        raise ValueError(_NONFINITE_MSG)
    if met["peak_raw"] < 1e-5:
        raise ValueError(_SILENT_MSG)


# ---- analyzers ----------------------------------------------------------------------------------------
def compute_spectrum_bars(audio: np.ndarray, sr: int, n_fft: int = 4096, n_bars: int = 64, min_hz: float = 20.0,
                          max_hz: float = 20000.0) -> list:
    """backend/app/pipeline.py:700-739 (the kernel is specialised to the reference's defaults)."""
    if (n_fft, n_bars, min_hz, max_hz) != (4096, 64, 20.0, 20000.0):
        raise NotImplementedError("spectrum kernel is built for n_fft=4096, 64 bars, 20 Hz..20 kHz")
    if np.size(audio) < n_fft:
        return [-80.0] * n_bars
    eng, b, _ = _up(audio, sr)
    return [float(v) for v in eng.spectrum_bars(b, 0)[0]]


def measure_stereo_correlation(audio: np.ndarray) -> Optional[float]:
    """backend/app/pipeline.py:766-791."""
    a = np.asarray(audio)
    if a.ndim != 2 or a.shape[1] != 2 or a.size < 4:
        return None
    eng, b, _ = _up(a, 44100)
    corr, _ = eng.stereo_correlation(b)
    return None if np.isnan(corr[0]) else float(corr[0])


def true_peak_dbfs(audio: np.ndarray, sr: int = 44100) -> float:
    """backend/app/routers/tools.py:44-54 (``_true_peak_dbfs``)."""
    if np.size(audio) == 0:
        return -120.0
    eng, b, _ = _up(audio, sr)
    return float(eng.true_peak(b)[0])


def compute_lufs_timeline(audio: np.ndarray, sr: int, block_sec: float = 0.4, max_points: int = 300):
    """backend/app/pipeline.py:667-697: integrated loudness of sliding ``block_sec`` segments.  The segments
    are gathered on the device into one batch (one row set per segment) and metered by one kernel launch."""
    import torch
    a = np.asarray(audio)
    duration_sec = len(a) / sr
    block = int(sr * block_sec)
    if duration_sec <= block_sec or a.size < block:
        v = measure_lufs(a, sr)
        return ([round(v, 2)] if not np.isnan(v) else [None], 0.0)
    n_points = min(max_points, max(1, int((duration_sec - block_sec) / (block_sec * 0.25)) + 1))
    step_sec = (duration_sec - block_sec) / max(n_points - 1, 1)
    step = int(sr * step_sec)
    count = 0
    pos = 0
    while pos + block <= len(a) and count < max_points:     # same loop bounds as the reference
        count += 1
        pos += step
        if step == 0:
            count = max_points
    if block < 0.4 * sr:                                       # pyloudnorm raises for every segment -> None
        return [None] * count, round(step_sec, 4)
    eng, b, _ = _up(a, sr)
    seg = eng.empty(count, b.channels, block, sr)
    with torch.cuda.stream(eng.stream):
        live = b.live()                                        # (channels, n)
        if step > 0:
            win = live.unfold(1, block, step)[:, :count]       # (channels, count, block) view
        else:
            win = live[:, None, :block].expand(-1, count, -1)
        seg.live().view(count, b.channels, block).copy_(win.permute(1, 0, 2))
    vals = eng.measure_lufs(seg)
    return [None if np.isnan(v) else round(float(v), 2) for v in vals], round(step_sec, 4)


def loudness_range_lu(audio: np.ndarray, sr: int) -> float:
    """backend/app/routers/tools.py:57-65 (``_loudness_range_lu``): p95 - p10 of the 3 s block loudness values."""
    timeline, _ = compute_lufs_timeline(audio, sr, block_sec=3.0, max_points=200)
    vals = [v for v in timeline if v is not None and v > -70]
    if len(vals) < 2:
        return 0.0
    arr = np.array(vals, dtype=np.float64)
    p10, p95 = np.percentile(arr, 10), np.percentile(arr, 95)
    return float(max(0.0, p95 - p10))


def compute_vectorscope_points(audio: np.ndarray, max_points: int = 1000) -> list:
    """backend/app/pipeline.py:742-763: at most ``max_points`` decimated [l, r] pairs, clipped, 5 decimals.
    A pure index gather of <= 1000 frames of the caller's own host array: no kernel involved."""
    a = np.asarray(audio)
    if a.ndim != 2 or a.shape[1] != 2 or a.size < 4:
        return []
    n = a.shape[0]
    step = max(1, n // max_points)
    idx = np.arange(0, n, step)[:max_points]
    pts = np.clip(a[idx].astype(np.float64), -1.0, 1.0)
    return [[round(float(l), 5), round(float(r), 5)] for l, r in pts]


def analyze_batch(tracks, sr, spectrum=True, eng: Optional[Engine] = None) -> list:
    """Additive: the analyzer set of /api/v2/analyze and /api/tools/lufs-analyze (routers/mastering.py:1198-1302,
    routers/tools.py:85-152) for equally-shaped tracks as one device batch: one upload, one kernel per metric."""
    eng = eng or get_engine()
    b = eng.upload(tracks, int(sr))
    lufs = eng.measure_lufs(b)
    tp, corr, peak = eng.true_peak_correlation(b)
    bars = [eng.spectrum_bars(b, v) for v in ((0, 1, 2) if (spectrum and b.channels == 2) else ((0,) if spectrum else ()))]
    out = []
    for t in range(b.tracks):
        rec = {"lufs": float(lufs[t]), "true_peak_dbfs": float(tp[t]), "sample_peak": float(peak[t]),
               "peak_dbfs": float(20 * np.log10(max(float(peak[t]), 1e-12))),
               "correlation": None if np.isnan(corr[t]) else float(corr[t])}
        if bars:
            rec["spectrum_bars"] = [float(v) for v in bars[0][t]]
            if len(bars) == 3:
                rec["spectrum_bars_mid"] = [float(v) for v in bars[1][t]]
                rec["spectrum_bars_side"] = [float(v) for v in bars[2][t]]
        out.append(rec)
    return out


# ---- chains -----------------------------------------------------------------------------------------
def _v1_progress_ticks(style, cfg, denoise_strength, transient_attack, transient_sustain, reference, reference_strength):
    """The (percent, message) pairs run_mastering_pipeline reports, in order, with its conditional ticks 22 / 57 / 78 / 86 /
    89 / 92 (backend/app/pipeline.py:1833-1907).  The router copies the message into the job record shown to the user
    (routers/mastering.py:379-381), so the texts are the reference's, keyed by the stage that has just finished."""
    t = [("start", 5, "Подготовка…"), ("dc_offset", 10, "Удаление DC-смещения"), ("peak_guard_in", 15, "Защита от пиков")]
    if denoise_strength > 0.01:
        t.append(("spectral_denoise", 22, f"Шумоподавление · strength={denoise_strength:.2f}"))
    t += [("target_eq", 32, "Студийный EQ"), ("deesser", 38, "De-esser (5–9 kHz)"),
          ("dynamics", 52, "Многополосная динамика и максимайзер")]
    if cfg.get("parallel_mix", 0.0) > 0.01:
        t.append(("parallel_compress", 57, f"Параллельная компрессия · mix={cfg['parallel_mix']:.2f}"))
    t += [("normalize_lufs", 65, "Нормализация LUFS"), ("final_spectral_balance", 72, "Финальная частотная коррекция")]
    if reference:
        t.append(("reference_match", 78, f"Reference mastering · strength={reference_strength:.2f}"))
    t.append(("style_eq", 82, f"Жанровый EQ · {style}"))
    if abs(transient_attack - 1.0) > 0.02 or abs(transient_sustain - 1.0) > 0.02:
        t.append(("transient_designer", 86, f"Транзиентный дизайнер · punch={transient_attack:.2f} sustain={transient_sustain:.2f}"))
    if cfg.get("exciter_db", 0.0) > 0.05:
        t.append(("harmonic_exciter", 89, f"Гармонический эксайтер · +{cfg['exciter_db']:.1f} dB"))
    if abs(cfg.get("imager_width", 1.0) - 1.0) > 0.01:
        t.append(("stereo_imager", 92, f"Стерео-расширение · width={cfg['imager_width']:.2f}"))
    t += [("peak_guard_out", 95, "Финальная защита пиков"), ("finalize_clip", 97, "Готово")]
    return t


def run_mastering_pipeline(audio: np.ndarray, sr: int, target_lufs: float = -14.0, style: str = "standard",
                           progress_callback: Optional[Callable[[int, str], None]] = None, denoise_strength: float = 0.0,
                           transient_attack: float = 1.0, transient_sustain: float = 1.0, reference_audio=None,
                           reference_sr=None, reference_strength: float = 0.8, trace_ctx=None) -> np.ndarray:
    """backend/app/pipeline.py:1800-1909.  The default path is one fused device call; a transient-designer request
    (which sits between the style EQ and the exciter, :1879-1889), spectral denoise (:1841-1844) or a reference track
    (:1868-1871) takes the stage-by-stage path.

    ``progress_callback(percent, message)``: the reference's percentages and messages in its order, conditional ticks
    included.  On the stage-by-stage path each fires when its stage has finished, as in the reference; on the fused path the
    first ("Подготовка…", 5) fires before the device call and the rest right after it (the stages have no host-visible
    boundaries inside one launch sequence -- a 3-minute track is ~10 ms of device time)."""
    name = style                                               # the reference's "Жанровый EQ" message shows the caller's string
    style = style if style in STYLE_CONFIGS else "standard"
    from . import mastering_trace as _mt
    tracing = trace_ctx is not None and _mt.trace_enabled()
    refm = reference_audio is not None and reference_sr is not None
    ticks = _v1_progress_ticks(name, STYLE_CONFIGS[style], denoise_strength, transient_attack, transient_sustain, refm, reference_strength)

    def report(stage):
        if progress_callback is not None:
            for key, pct, msg in ticks:
                if key == stage:
                    progress_callback(pct, msg)

    report("start")
    if tracing or refm or denoise_strength > 0.01 or abs(transient_attack - 1.0) > 0.02 or abs(transient_sustain - 1.0) > 0.02:
        out = _run_v1_stagewise(audio, sr, target_lufs, style, transient_attack, transient_sustain,
                                trace_ctx=trace_ctx if tracing else None, denoise_strength=denoise_strength,
                                reference=(reference_audio, reference_sr, reference_strength) if refm else None, report=report)
    else:
        out = master_batch([audio], sr, [style], [target_lufs], chain="v1")["audio"][0]
        for key, _, _ in ticks[1:]:
            report(key)
    _trim_engine()
    return out


def _trim_engine():
    """Give the calling thread's device scratch back when it exceeds ``MM_WORKSPACE_KEEP_MB`` (default 2048): worker threads
    of a web service (asyncio.to_thread: up to min(32, cpu + 4) of them) would otherwise each keep their high-water mark."""
    import os
    eng = get_engine()
    keep = float(os.environ.get("MM_WORKSPACE_KEEP_MB", "2048")) * (1 << 20)
    if eng.lib.mm_ctx_workspace_bytes(eng.ctx) > keep:
        eng.release_workspace()


def _run_v1_stagewise(audio, sr, target_lufs, style, transient_attack, transient_sustain, trace_ctx=None, reference=None,
                      denoise_strength=0.0, report=None):
    """run_mastering_pipeline stage by stage (pipeline.py:1833-1909), for the options the fused chain does not carry:
    the transient designer, and the per-stage trace (mastering_trace.trace_stage after every stage, same stage names)."""
    if _os.environ.get("MM_RESIDENT", "1") == "0":           # A/B switch: every stage function uploads and downloads (round-2 start)
        return _run_v1_stages(audio, sr, target_lufs, style, transient_attack, transient_sustain, trace_ctx, reference,
                              denoise_strength, report)
    with _ResidentScope():        # the samples stay on the device between the stage functions; one download at the end
        return _materialize(_run_v1_stages(audio, sr, target_lufs, style, transient_attack, transient_sustain, trace_ctx, reference,
                                           denoise_strength, report))


def _run_v1_stages(audio, sr, target_lufs, style, transient_attack, transient_sustain, trace_ctx, reference, denoise_strength, report):
    from .mastering_trace import trace_stage
    cfg = STYLE_CONFIGS[style]

    def tr(name, a, **extra):
        trace_stage(trace_ctx, name, a, sr, **extra)
        if report is not None:
            report(name)
        return a

    a = tr("dc_offset", remove_dc_offset(audio))
    a = tr("peak_guard_in", remove_intersample_peaks(a, headroom_db=0.5))
    if denoise_strength > 0.01:                                          # pipeline.py:1841-1844
        a = tr("spectral_denoise", apply_spectral_denoise(a, sr, strength=denoise_strength), denoise_strength=denoise_strength)
    a = tr("target_eq", apply_target_curve(a, sr))
    a = tr("deesser", apply_deesser(a, sr))
    a = tr("dynamics", apply_dynamics(a, sr))
    if cfg.get("parallel_mix", 0.0) > 0.01:
        a = tr("parallel_compress", apply_parallel_compression(a, sr, mix=cfg["parallel_mix"]), parallel_mix=cfg["parallel_mix"])
    a = tr("normalize_lufs", normalize_lufs(a, sr, target_lufs), target_lufs=target_lufs)
    a = tr("final_spectral_balance", apply_final_spectral_balance(a, sr))
    if reference is not None:                                            # pipeline.py:1868-1871
        a = tr("reference_match", apply_reference_match(a, sr, reference[0], reference[1], strength=reference[2]),
               reference_strength=reference[2])
    a = tr("style_eq", apply_style_eq(a, sr, style), style=style)
    if abs(transient_attack - 1.0) > 0.02 or abs(transient_sustain - 1.0) > 0.02:
        a = tr("transient_designer", apply_transient_designer(a, sr, attack_gain=transient_attack, sustain_gain=transient_sustain),
               transient_attack=transient_attack, transient_sustain=transient_sustain)
    if cfg.get("exciter_db", 0.0) > 0.05:
        a = tr("harmonic_exciter", apply_harmonic_exciter(a, sr, cfg["exciter_db"]), exciter_db=cfg["exciter_db"])
    if abs(cfg.get("imager_width", 1.0) - 1.0) > 0.01:
        a = tr("stereo_imager", apply_stereo_imager(a, cfg["imager_width"]), imager_width=cfg["imager_width"])
    a = tr("peak_guard_out", remove_intersample_peaks(a, headroom_db=0.5))
    a = apply_output_edge_fade_in(a, sr, fade_ms=6.0)
    trace_stage(trace_ctx, "output_fade_in", a, sr)
    return tr("finalize_clip", _stage("finalize_clip", a, sr))         # clip + nan_to_num (pipeline.py:1904-1906), on the device


def export_audio(samples: np.ndarray, sr: int, channels: int, out_format: str = "wav", dither_type: str = "tpdf",
                 auto_blank_sec: float = 0.0, bitrate=None, noise: Optional[np.ndarray] = None, seed: int = 0) -> bytes:
    """backend/app/pipeline.py:965-991, WAV branch with TPDF or noise-shaped ("ns_e", "ns_itu") dither.  ``noise=`` is
    additive (bit-exact test hook): for TPDF the float32 dither buffer itself, for the shaped types the float32 uniforms
    ``np.random.rand(n, ch).astype(np.float32)`` the reference would have drawn (:843 / :864)."""
    if out_format.lower() != "wav":
        raise NotImplementedError("only the WAV branch is on the hot path (codecs are host-side I/O, SURVEY L0); "
                                  "export_pcm24 hands 24-bit samples to a host FLAC encoder")
    eng, b, _ = _up(samples, sr)
    if auto_blank_sec > 0:                                 # pipeline.py:976-977
        b = eng.auto_blank_end(b, -50.0, auto_blank_sec)
    shape = {"ns_e": 1, "ns_itu": 2}.get(dither_type or "tpdf", 0)
    if shape and b.n >= (4 if shape == 1 else 8):         # shorter buffers: the reference falls back to TPDF (:840, :861)
        pcm = eng.quantize_int16_shaped(b, shape, uniform=noise, seed=seed)[0]
    else:
        pcm = eng.quantize_int16(b, noise=noise, seed=seed)[0]
    return wavio.pack_wav_pcm16(pcm, sr)


def export_pcm24(samples: np.ndarray, sr: int, auto_blank_sec: float = 0.0) -> np.ndarray:
    """The sample conversion of export_audio's FLAC branch (backend/app/pipeline.py:981-985: libsndfile PCM_24 from
    float32 = clip, lrintf(x * 0x7FFFFF)) on the device -> int32 (n, channels) for a host-side FLAC encoder."""
    eng, b, _ = _up(samples, sr)
    if auto_blank_sec > 0:
        b = eng.auto_blank_end(b, -50.0, auto_blank_sec)
    return eng.quantize_pcm24(b)[0]


def load_audio_from_bytes(data: bytes, fmt: str = "wav"):
    """backend/app/pipeline.py:802-827, WAV only (container I/O is outside the hot path)."""
    if fmt.lower().lstrip(".") != "wav":
        raise NotImplementedError("only WAV decoding is provided")
    return wavio.unpack_wav(data)


# ---- additive: batched entry point ------------------------------------------------------------------
def master_batch(tracks, sr, styles, targets=None, chain="v2", want_int16=False, noise=None, seed=0, job_fade=True,
                 measure=False, eng: Optional[Engine] = None, compressor: Optional[str] = None):
    """Master equally-shaped tracks as one device batch.

    tracks: list of (n,) / (n, ch) float32 arrays; styles: list of style names; targets: list of LUFS
    (default: the style's own).  Returns dict(audio=[...], int16=[...] or None, stats=[dict,...])."""
    eng = eng or get_engine()
    mono = np.asarray(tracks[0]).ndim == 1
    b = eng.upload(tracks, int(sr))
    names = [s if s in STYLE_CONFIGS else "standard" for s in styles]
    if targets is None:
        targets = [STYLE_CONFIGS[s]["lufs"] for s in names]
    sts = [style_struct(STYLE_CONFIGS[s], t) for s, t in zip(names, targets)]
    flags = (0 if job_fade else _lib.FLAG_NO_JOB_FADE) | ((_lib.FLAG_MEASURE_IN | _lib.FLAG_MEASURE_OUT) if measure else 0)
    if _compressor_id(compressor) == _lib.COMPRESSOR_ENVELOPE:
        flags |= _lib.FLAG_ENVELOPE_COMPRESSOR
    out, pcm, stats = eng.master(b, _lib.CHAIN_V1 if chain == "v1" else _lib.CHAIN_V2, sts, out=b, want_int16=want_int16,
                                 noise=noise, seed=seed, flags=flags)
    audio = eng.download(out)
    res = {"audio": [a[:, 0] if mono else a for a in audio], "int16": None, "stats": []}
    if pcm is not None:
        res["int16"] = list(eng.to_host(pcm))
    for s in stats:
        res["stats"].append({k: (list(getattr(s, k)) if k == "mean" else getattr(s, k)) for k, _ in s._fields_})
    return res


def master_wav_jobs(wavs, styles, targets=None, chain="v2", seed=0, eng: Optional[Engine] = None) -> list:
    """Additive, job level: what ``_run_mastering_job(_v2)`` does per upload (routers/mastering.py:350-637) -- and what
    ``/api/v2/batch`` (routers/mastering.py:855-1037) does for up to ten uploads one after the other -- for a whole list of PCM_16
    WAV uploads: decode, master, dither, encode, with the PCM frames crossing PCIe as int16 in both directions.  Uploads may differ
    in length, channel count and sample rate: ``mm_master_host_jobs`` takes the list as it is (one ``mm_host_job`` per upload, its
    frames read in place from the WAV bytes), merges runs of equal shape into chunks and sends every chunk through ONE copy-in /
    chain / copy-out pipeline, so an upload's transfer overlaps its neighbour's chain.  Every track's result -- dither stream
    included, keyed by the track's index in ``wavs`` -- equals what the same upload gives alone at that index.  Returns a list of
    dicts ``wav`` (bytes), ``stats``."""
    eng = eng or get_engine()
    metas = [wavio.pcm16_view(w) for w in wavs]
    names = [s if s in STYLE_CONFIGS else "standard" for s in styles]
    if targets is None:
        targets = [STYLE_CONFIGS[s]["lufs"] for s in names]
    flags = (_lib.FLAG_MEASURE_IN | _lib.FLAG_MEASURE_OUT |
             (_lib.FLAG_ENVELOPE_COMPRESSOR if _compressor_id() == _lib.COMPRESSOR_ENVELOPE else 0))
    jobs = (_lib.HostJob * len(wavs))()
    keep = []                                                            # the views / result arrays the job list points into
    for i, (frames, n, ch, sr) in enumerate(metas):
        if n == 0:
            raise ValueError("master_wav_jobs: an upload holds no frames")
        src = np.ascontiguousarray(frames)                               # a view of the upload's bytes (no copy) unless it was strided
        dst = np.empty((n, ch), np.int16)
        keep.append((src, dst))
        j = jobs[i]
        j.n, j.channels, j.sr = n, ch, sr
        j.pcm16_in, j.pcm16_out = src.ctypes.data, dst.ctypes.data
        j.style = style_struct(STYLE_CONFIGS[names[i]], targets[i])
        j.dither_id = i
    _lib.check(eng.lib.mm_master_host_jobs(eng.ctx, _lib.CHAIN_V1 if chain == "v1" else _lib.CHAIN_V2, len(wavs), jobs, int(seed), flags))
    out = []
    for i, (_, dst) in enumerate(keep):
        st = jobs[i].stats
        rec = {f: (list(getattr(st, f)) if f == "mean" else getattr(st, f)) for f, _ in st._fields_}
        out.append({"wav": wavio.pack_wav_pcm16(dst, metas[i][3]), "stats": rec})
    _trim_engine()
    return out
