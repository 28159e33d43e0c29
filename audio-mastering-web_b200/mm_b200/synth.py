"""Deterministic synthetic tracks for benchmarks and parity tests.

Track ``t`` is defined by parameters drawn from ``np.random.default_rng(1000 + t)``: six partials
(log-uniform 40 Hz..12 kHz, the right channel shares four of them with the left), a slow
amplitude envelope, low-passed noise, a small DC offset, and a target sample peak in
[0.3, 0.98] so that some tracks trip the -0.5 dB input peak guard.

The partial/envelope/DC/peak parameters are identical on every backend.  The noise term is
white noise through a one-pole 800 Hz low-pass normalised to unit RMS; ``numpy_track`` draws it
from the same numpy generator, ``torch_batch`` draws it on the device (same statistics, different
samples) -- parity tests always copy the very array one side mastered to the other side, so
only the shape of the workload, not the noise realisation, has to agree.
"""
from __future__ import annotations

import math

import numpy as np


def track_params(t: int):
    rng = np.random.default_rng(1000 + int(t))
    p = {}
    p["f_l"] = np.exp(rng.uniform(math.log(40.0), math.log(12000.0), 6))
    p["a_l"] = rng.uniform(0.02, 0.15, 6)
    p["ph_l"] = rng.uniform(0.0, 2 * math.pi, 6)
    # right channel: partials 0..3 shared with left, 4..5 its own
    p["f_r"] = p["f_l"].copy()
    p["a_r"] = p["a_l"].copy()
    p["ph_r"] = p["ph_l"].copy()
    p["f_r"][4:] = np.exp(rng.uniform(math.log(40.0), math.log(12000.0), 2))
    p["a_r"][4:] = rng.uniform(0.02, 0.15, 2)
    p["ph_r"][4:] = rng.uniform(0.0, 2 * math.pi, 2)
    p["env_rate"] = rng.uniform(0.1, 0.5)
    p["env_phase"] = rng.uniform(0.0, 2 * math.pi)
    p["dc"] = rng.uniform(-2e-3, 2e-3, 2)
    p["peak"] = rng.uniform(0.3, 0.98)
    p["noise_seed"] = int(rng.integers(0, 2**31 - 1))
    return p


def _onepole_coef(sr: float) -> float:
    return math.exp(-2.0 * math.pi * 800.0 / sr)


def numpy_track(t: int, sr: int, dur_sec: float, channels: int = 2) -> np.ndarray:
    """float32 ``(n, channels)`` track ``t`` synthesised on the host."""
    from scipy.signal import lfilter

    p = track_params(t)
    n = int(round(sr * dur_sec))
    tt = np.arange(n, dtype=np.float64) / sr
    env = 0.55 + 0.45 * np.sin(2 * math.pi * p["env_rate"] * tt + p["env_phase"])
    rng = np.random.default_rng(p["noise_seed"])
    k = _onepole_coef(sr)
    out = np.empty((n, channels), dtype=np.float64)
    for c in range(channels):
        f, a, ph = (p["f_l"], p["a_l"], p["ph_l"]) if c == 0 else (p["f_r"], p["a_r"], p["ph_r"])
        s = np.zeros(n)
        for i in range(6):
            s += a[i] * np.sin(2 * math.pi * f[i] * tt + ph[i])
        w = lfilter([1.0 - k], [1.0, -k], rng.standard_normal(n))
        w /= math.sqrt(float(np.mean(w * w)) + 1e-30)
        out[:, c] = env * (s + 0.08 * w) + p["dc"][c % 2]
    out *= p["peak"] / max(float(np.max(np.abs(out))), 1e-30)
    return out.astype(np.float32)


def torch_batch(track_ids, sr: int, dur_sec: float, device, channels: int = 2, out=None, row_stride=None, lead=0):
    """Synthesise tracks on ``device`` into planar rows ``[track*channels + c]``.

    Returns a float32 tensor ``(len(track_ids)*channels, n)`` (or fills ``out`` rows, whose
    first sample sits ``lead`` floats into a row of ``row_stride`` floats).
    """
    import torch

    n = int(round(sr * dur_sec))
    rows = len(track_ids) * channels
    if out is None:
        out = torch.empty((rows, n), dtype=torch.float32, device=device)
        view = out
    else:
        view = out.view(-1)[: rows * row_stride].view(rows, row_stride)[:, lead:lead + n]
    tt = torch.arange(n, dtype=torch.float64, device=device) / sr
    k = _onepole_coef(sr)
    # |H| of the one-pole low-pass on the rfft grid (noise is shaped in the frequency domain)
    w = torch.arange(n // 2 + 1, dtype=torch.float64, device=device) * (2 * math.pi / n)
    hmag = ((1.0 - k) / torch.sqrt(1.0 - 2.0 * k * torch.cos(w) + k * k)).to(torch.float32)
    for j, t in enumerate(track_ids):
        p = track_params(t)
        env = 0.55 + 0.45 * torch.sin(2 * math.pi * p["env_rate"] * tt + p["env_phase"])
        g = torch.Generator(device=device)
        g.manual_seed(p["noise_seed"])
        chans = []
        for c in range(channels):
            f, a, ph = (p["f_l"], p["a_l"], p["ph_l"]) if c == 0 else (p["f_r"], p["a_r"], p["ph_r"])
            s = torch.zeros(n, dtype=torch.float64, device=device)
            for i in range(6):
                s += float(a[i]) * torch.sin(2 * math.pi * float(f[i]) * tt + float(ph[i]))
            wn = torch.randn(n, dtype=torch.float32, device=device, generator=g)
            wn = torch.fft.irfft(torch.fft.rfft(wn) * hmag, n=n)
            wn = wn / torch.sqrt(torch.mean(wn * wn) + 1e-30)
            chans.append(env * (s + 0.08 * wn.to(torch.float64)) + float(p["dc"][c % 2]))
        pk = max(float(torch.max(torch.abs(ch))) for ch in chans)
        for c, ch in enumerate(chans):
            view[j * channels + c].copy_((ch * (p["peak"] / max(pk, 1e-30))).to(torch.float32))
    return out


def torch_long_slice(t: int, sr: int, start: int, stop: int, device, out_rows, lead: int = 0, block: int = 1 << 22):
    """Frames [start, stop) of the LONG-FORM variant of track ``t`` (BASELINE config 5) written into the two planar rows
    ``out_rows`` (float32 tensor (2, >= lead + stop - start)).  Same partials / envelope / DC as ``numpy_track``; so that
    every rank of a time-split run sees identical samples where slices overlap, the noise is white noise keyed by
    (seed, 2^20-frame block index, channel) -- generated on the device per block -- and the level is fixed analytically
    (peak target / bound on the partial sum) instead of by a pass over the whole file."""
    import torch

    p = track_params(t)
    nb = 1 << 20
    for c in range(2):
        f, a, ph = (p["f_l"], p["a_l"], p["ph_l"]) if c == 0 else (p["f_r"], p["a_r"], p["ph_r"])
        gain = float(p["peak"]) / (float(np.sum(a)) + 0.35 + 2e-3)
        for b0 in range(start, stop, block):
            b1 = min(stop, b0 + block)
            tt = torch.arange(b0, b1, dtype=torch.float64, device=device) / sr
            s = torch.zeros(b1 - b0, dtype=torch.float64, device=device)
            for i in range(6):
                s += float(a[i]) * torch.sin(2 * math.pi * float(f[i]) * tt + float(ph[i]))
            env = 0.55 + 0.45 * torch.sin(2 * math.pi * p["env_rate"] * tt + p["env_phase"])
            wn = torch.empty(b1 - b0, dtype=torch.float32, device=device)
            for k in range(b0 // nb, (b1 - 1) // nb + 1):
                g = torch.Generator(device=device)
                g.manual_seed((int(p["noise_seed"]) * 2 + c) * 1_000_003 + k)
                blk = torch.randn(nb, dtype=torch.float32, device=device, generator=g)
                lo, hi = max(b0, k * nb), min(b1, (k + 1) * nb)
                wn[lo - b0:hi - b0] = blk[lo - k * nb:hi - k * nb]
            x = (env * (s + 0.08 * wn.to(torch.float64)) + float(p["dc"][c])) * gain
            out_rows[c, lead + b0 - start:lead + b1 - start].copy_(x.to(torch.float32))
    return out_rows
