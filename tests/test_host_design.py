"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol;
filter design matches scipy; the scan tables reproduce scipy.lfilter when the tile decomposition
of csrc/sweep.cuh is replayed in numpy."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy import signal as sg

from conftest import REPO


@pytest.fixture(scope="module")
def lib():
    from mm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from mm_b200 import build
        build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from mm_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "mm_b200.h")).read()
    declared = set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mm_ctx", "mm_geom", "mm_style", "mm_track_stats", "mm_ktime"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mm_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.mm_abi_version() == 1
    assert lib.mm_row_stride(1) % 32 == 0 and lib.mm_row_stride(7_938_000) >= 7_938_000 + 64


def test_no_cuda_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mm_b200 import _lib
    ctx = C.c_void_p()
    assert lib.mm_ctx_create(0, None, C.byref(ctx)) != 0
    assert "no CPU fallback" in _lib.last_error() or "CUDA" in _lib.last_error()
    from mm_b200.engine import Engine, MMError
    with pytest.raises(MMError):
        Engine(0)


def _butter(lib, order, btype, wn):
    from mm_b200 import _lib
    b = (C.c_double * 5)()
    a = (C.c_double * 5)()
    n = lib.mm_design_butter(order, btype, _lib.darr(wn), b, a)
    assert n > 0, _lib.last_error()
    return np.array(b[:n]), np.array(a[:n])


@pytest.mark.parametrize("sr", [22050, 44100, 48000, 88200, 96000, 192000])
def test_butter_and_zi_match_scipy(lib, sr):
    from mm_b200 import _lib
    nyq = sr / 2
    cases = [(2, 1, [40 / nyq], "high"), (2, 0, [min(18000 / nyq, 0.99)], "low"), (2, 0, [180 / nyq], "low"),
             (2, 1, [min(16000 / nyq, .99)], "high"), (1, 2, [30 / nyq, 90 / nyq], "band"),
             (1, 2, [min(3000 / nyq, .99) * 0.7, min(3000 / nyq, .99) * 1.3], "band"),
             (2, 2, [min(5000 / nyq, .97) * 0.999, min(9000 / nyq, .97)], "band"),
             (2, 0, [214 / nyq], "low"), (2, 1, [min(10000 / nyq, .99)], "high")]
    for order, bt, wn, name in cases:
        if len(wn) == 2 and wn[0] >= wn[1]:
            continue
        b, a = _butter(lib, order, bt, wn)
        rb, ra = sg.butter(order, wn if len(wn) == 2 else wn[0], name)
        assert np.allclose(b, rb, rtol=1e-12, atol=1e-16), (sr, order, name, b, rb)
        assert np.allclose(a, ra, rtol=1e-12, atol=1e-15), (sr, order, name, a, ra)
        zi = (C.c_double * 4)()
        assert lib.mm_design_lfilter_zi(_lib.darr(rb), _lib.darr(ra), len(rb), zi) == 0
        assert np.allclose(np.array(zi[:len(rb) - 1]), sg.lfilter_zi(rb, ra), rtol=1e-9, atol=1e-12)


def test_k_weighting_matches_oracle(lib):
    from oracle import bs1770
    for sr in (22050, 44100, 48000, 96000, 192000):
        ref = bs1770.k_weighting_coeffs(sr)
        for stage in (0, 1):
            b = (C.c_double * 3)()
            a = (C.c_double * 3)()
            assert lib.mm_design_k_weighting(stage, float(sr), b, a) == 0
            assert np.allclose(np.array(b[:]), ref[stage][0], rtol=1e-13)
            assert np.allclose(np.array(a[:]), ref[stage][1], rtol=1e-13)


def test_k_weighting_highpass_as_state_variable_filter(lib):
    """The loudness kernel runs the K-weighting high-pass as a Chamberlin state-variable filter (4 float32 operations per sample,
    csrc/lufs_kernel.cuh): the realization has the biquad's transfer function, and its float32 recurrence restarted from the exact
    state every 64 samples (what the kernel's scan provides) stays within 1e-6 of scipy's float64 lfilter."""
    import scipy.signal as sg
    from mm_b200 import _lib
    rng = np.random.default_rng(3)
    for sr in (8000, 22050, 44100, 48000, 96000, 192000):
        b = (C.c_double * 3)()
        a = (C.c_double * 3)()
        assert lib.mm_design_k_weighting(1, float(sr), b, a) == 0
        fqg = (C.c_double * 3)()
        abcd = (C.c_double * 9)()
        assert lib.mm_design_svf_highpass(b, a, fqg, abcd) == 0
        f, q, g = fqg[:]
        A = np.array(abcd[:4]).reshape(2, 2)
        B, Cm, D = np.array(abcd[4:6]), np.array(abcd[6:8]), abcd[8]
        assert np.allclose(A, [[1.0, f], [-f, 1.0 - f * (f + q)]]) and np.allclose(B, [0.0, f])
        assert np.allclose(Cm, [-g, -g * (f + q)]) and D == g
        n = 1 << 15
        x = (0.3 * rng.standard_normal(n) + 0.2 + 0.5 * np.sin(2 * np.pi * 30.0 * np.arange(n) / sr)).astype(np.float32)
        ref = sg.lfilter(np.array(b[:]), np.array(a[:]), x.astype(np.float64))
        # float64 state-space replay: same transfer function
        s = np.zeros(2)
        y = np.empty(n)
        S = np.empty((n + 1, 2))
        S[0] = s
        xd = x.astype(np.float64)
        for i in range(n):
            y[i] = Cm @ s + D * xd[i]
            s = A @ s + B * xd[i]
            S[i + 1] = s
        assert np.max(np.abs(y - ref)) <= 1e-11
        # the kernel's float32 recurrence, restarted from the exact state every 64 samples
        f32 = np.float32
        ff, nq = f32(f), f32(-q)
        got = np.empty(n)
        for c0 in range(0, n, 64):
            lp, bp = f32(S[c0, 0]), f32(S[c0, 1])
            for i in range(c0, min(n, c0 + 64)):
                lp = f32(np.float64(ff) * np.float64(bp) + np.float64(lp))           # fmaf: one rounding
                hp = f32(np.float64(nq) * np.float64(bp) + np.float64(f32(x[i] - lp)))
                bp = f32(np.float64(ff) * np.float64(hp) + np.float64(bp))
                got[i] = g * float(hp)
        assert np.max(np.abs(got - ref)) <= 1e-6, sr
        assert abs(np.sum(got ** 2) / np.sum(ref ** 2) - 1.0) <= 1e-7
    # a biquad that is not a double-zero high-pass is refused by name
    bl, al = sg.butter(2, 0.1)
    assert lib.mm_design_svf_highpass(_lib.darr(bl), _lib.darr(al), fqg, abcd) != 0
    assert "double zero" in _lib.last_error()


def _tables(lib, b, a):
    from mm_b200 import _lib
    m = len(b) - 1
    mm2 = m * m
    capw = 64
    g = (C.c_double * (32 * m))()
    Pw = (C.c_double * (5 * mm2))()
    Plane = (C.c_double * (32 * mm2))()
    Qpow = (C.c_double * (5 * mm2))()
    Mpow = (C.c_double * (capw * mm2))()
    Apow = (C.c_double * (33 * mm2))()
    zi = (C.c_double * m)()
    S, T = C.c_int(), C.c_int()
    W = lib.mm_design_scan_tables(_lib.darr(b), _lib.darr(a), len(b), g, Pw, Plane, Qpow, Mpow, capw, Apow, zi, C.byref(S), C.byref(T))
    assert W > 0, _lib.last_error()
    assert (S.value, T.value) == (32, 128)
    r = lambda arr, k: np.array(arr[:]).reshape(k, m, m)  # noqa: E731
    return dict(m=m, W=W, g=np.array(g[:]).reshape(32, m), Pw=r(Pw, 5), Plane=r(Plane, 32), Qpow=r(Qpow, 5),
                Mpow=r(Mpow, capw)[:min(W, capw)], Apow=r(Apow, 33), zi=np.array(zi[:]))


def _replay_tile_scan(tb, b, a, x, zi_scale):
    """numpy replay of csrc/sweep.cuh tile_scan (forward direction, one row) over len(x) samples that
    start `dead` = 0 samples into tile 0; initial state zi * zi_scale injected at sample 0."""
    m, S, T = tb["m"], 32, 128
    L = S * T
    n = len(x)
    ntiles = (n + L - 1) // L
    xp = np.zeros(ntiles * L)
    xp[:n] = x
    A = tb["Apow"][1]
    y = np.zeros(ntiles * L)
    aggs = []
    for tile in range(ntiles):
        ch = xp[tile * L:(tile + 1) * L].reshape(T, S)
        E = ch @ tb["g"]                                  # zero-state end state per thread  (T, m)
        if tile == 0:
            s_init = tb["zi"] * zi_scale
            E[0] = E[0] + tb["Apow"][S] @ s_init          # dead = 0: state injected before sample 0
        # warp scans (Kogge-Stone with Pw) == sequential inclusive prefix inside each warp
        inc = np.zeros_like(E)
        for w in range(T // 32):
            e = E[w * 32:(w + 1) * 32].copy()
            for d in range(5):
                sh = np.zeros_like(e)
                sh[1 << d:] = e[:-(1 << d)]
                e = e + sh @ tb["Pw"][d].T * (np.arange(32)[:, None] >= (1 << d))
            inc[w * 32:(w + 1) * 32] = e
        tot = inc[31::32]
        base = np.zeros((T // 32, m))
        for w in range(1, T // 32):
            base[w] = tot[w - 1] + tb["Qpow"][1] @ base[w - 1]
        agg = tot[-1] + tb["Qpow"][1] @ base[-1]
        aggs.append(agg)
        Cin = np.zeros(m)
        for j in range(min(tb["W"], tile, len(tb["Mpow"]))):
            Cin = Cin + tb["Mpow"][j] @ aggs[tile - 1 - j]
        for t in range(T):
            w, l = divmod(t, 32)
            bw = base[w] + tb["Qpow"][w] @ Cin
            z = (inc[t - 1] if l > 0 else np.zeros(m)) + tb["Plane"][l] @ bw
            if tile == 0 and t == 0:
                z = tb["zi"] * zi_scale
            z = z.copy()
            for j in range(S):                            # DF2T, as df2t_step
                xv = ch[t, j]
                yv = b[0] * xv + z[0]
                for i in range(m):
                    nxt = z[i + 1] if i + 1 < m else 0.0
                    z[i] = b[i + 1] * xv - a[i + 1] * yv + nxt
                y[tile * L + t * S + j] = yv
        assert np.allclose(A, tb["Apow"][1])
    return y[:n]


@pytest.mark.parametrize("design", ["hp40_96k", "bp30_90_96k", "lp18k_44k", "bp4_deess_48k", "kw_hp38_192k"])
def test_scan_tables_replay_equals_lfilter(lib, design):
    from oracle import bs1770
    b, a = {
        "hp40_96k": sg.butter(2, 40 / 48000, "high"),
        "bp30_90_96k": sg.butter(1, [30 / 48000, 90 / 48000], "band"),
        "lp18k_44k": sg.butter(2, 18000 / 22050, "low"),
        "bp4_deess_48k": sg.butter(2, [5000 / 24000, 9000 / 24000], "band"),
        "kw_hp38_192k": bs1770.k_weighting_coeffs(192000)[1],
    }[design]
    tb = _tables(lib, b, a)
    rng = np.random.default_rng(11)
    n = 4096 * 3 + 1717 if tb["W"] <= 3 else 4096 * (tb["W"] + 2) + 5
    n = min(n, 4096 * 9)
    x = 0.2 * rng.standard_normal(n) + 0.3
    y = _replay_tile_scan(tb, b, a, x, zi_scale=x[0])
    ref, _ = sg.lfilter(b, a, x, zi=sg.lfilter_zi(b, a) * x[0])
    assert tb["W"] >= 1
    assert np.max(np.abs(y - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), (design, tb["W"], np.max(np.abs(y - ref)))


# ---- float32 pass 2 on the balanced realization (csrc/sweep3.cuh NF32 sections, csrc/lufs_kernel.cuh) ----------
def _tables2(lib, b, a, mode):
    from mm_b200 import _lib
    m = len(b) - 1
    mm2 = m * m
    g = (C.c_double * (32 * m))()
    Plane = (C.c_double * (32 * mm2))()
    Apow = (C.c_double * (33 * mm2))()
    zi = (C.c_double * m)()
    ss = (C.c_double * (mm2 + 2 * m + 2))()
    W = lib.mm_design_scan_tables2(_lib.darr(b), _lib.darr(a), len(b), mode, g, None, Plane, None, None, 0, Apow, zi, ss)
    assert W > 0, _lib.last_error()
    ssv = np.array(ss[:])
    return dict(m=m, W=W, g=np.array(g[:]).reshape(32, m), Plane=np.array(Plane[:]).reshape(32, m, m),
                Apow=np.array(Apow[:]).reshape(33, m, m), zi=np.array(zi[:]), A=ssv[:mm2].reshape(m, m), B=ssv[mm2:mm2 + m],
                C=ssv[mm2 + m:mm2 + 2 * m], D=ssv[mm2 + 2 * m], norm2=ssv[mm2 + 2 * m + 1])


def _chunked_f32_pass2(tb, x, zi_scale):
    """Chunk-start states exactly as the float64 scan resolves them (here: a float64 run of the realization),
    then the float32 recurrence of ss32_step inside every 32-sample chunk: the balanced realization with its states
    rescaled by d_i = 1 / B_i (B = 1: 7 operations per biquad sample), as csrc/stages.cu fill_filter prepares it."""
    A, B, Cv, D = tb["A"], tb["B"], tb["C"], tb["D"]
    n = len(x)
    m = tb["m"]
    s = tb["zi"] * zi_scale
    starts = np.zeros((n // 32 + 1, m))
    for i in range(n):
        if i % 32 == 0:
            starts[i // 32] = s
        s = A @ s + B * x[i]
    d = 1.0 / B
    A32 = (d[:, None] * A / d[None, :]).astype(np.float32)
    C32, D32, d32 = (Cv / d).astype(np.float32), np.float32(D), d.astype(np.float32)
    y = np.zeros(n, np.float32)
    x32 = x.astype(np.float32)
    for c0 in range(0, n, 32):
        sf = (starts[c0 // 32] * d32.astype(np.float64)).astype(np.float32)
        for i in range(c0, min(c0 + 32, n)):
            acc = np.float32(D32 * x32[i])
            for k in range(m - 1, -1, -1):
                acc = np.float32(np.float64(C32[k]) * np.float64(sf[k]) + np.float64(acc))      # fmaf
            y[i] = acc
            sn = np.zeros_like(sf)
            for r in range(m):
                t = x32[i]
                for k in range(m - 1, -1, -1):
                    t = np.float32(np.float64(A32[r, k]) * np.float64(sf[k]) + np.float64(t))   # fmaf
                sn[r] = t
            sf = sn
    return y


@pytest.mark.parametrize("design,tol", [("lp18k_44k", 2.5e-7), ("hp2230_96k", 4e-7), ("hp10k_48k", 2.5e-7), ("bp_pres_44k", 3e-7),
                                        ("hp40_96k", 2.5e-6), ("lp180_48k", 2.5e-6), ("kw_shelf_48k", 5e-7), ("kw_hp38_44k", 2.5e-6)])
def test_balanced_realization_float32_pass2(lib, design, tol):
    """The balanced realization is the same filter (float64), a contraction (||A||_2 <= 1), and its float32
    per-chunk recurrence stays within `tol` of scipy's float64 lfilter on a loud mix of 55 Hz + 3 kHz + noise:
    ~1e-7 where state errors die inside a chunk, ~1e-6 for the low cut-offs (which the kernels only run in
    float32 behind recombination weights <= 0.3 or inside the loudness meter)."""
    from oracle import bs1770
    b, a = {
        "lp18k_44k": sg.butter(2, 18000 / 22050, "low"), "hp2230_96k": sg.butter(2, 2230 / 48000, "high"),
        "hp10k_48k": sg.butter(2, 10000 / 24000, "high"), "bp_pres_44k": sg.butter(1, [2100 / 22050, 3900 / 22050], "band"),
        "hp40_96k": sg.butter(2, 40 / 48000, "high"), "lp180_48k": sg.butter(2, 180 / 24000, "low"),
        "kw_shelf_48k": bs1770.k_weighting_coeffs(48000)[0], "kw_hp38_44k": bs1770.k_weighting_coeffs(44100)[1],
    }[design]
    sr = {"44k": 44100, "48k": 48000, "96k": 96000}[design.rsplit("_", 1)[1]]
    tb = _tables2(lib, b, a, 1)
    td = _tables2(lib, b, a, 0)
    assert tb["norm2"] <= 1.0 + 1e-12 and abs(td["norm2"] - tb["norm2"]) < 1e-9
    assert np.allclose(np.linalg.eigvals(tb["A"]), np.linalg.eigvals(td["A"]), atol=1e-9) or \
        np.allclose(sorted(np.linalg.eigvals(tb["A"]), key=lambda v: (v.real, v.imag)),
                    sorted(np.linalg.eigvals(td["A"]), key=lambda v: (v.real, v.imag)), atol=1e-7)
    # g and Plane are expressed in the balanced coordinates
    assert np.allclose(tb["g"][31], tb["B"], rtol=1e-12, atol=1e-15)
    assert np.allclose(tb["g"][30], tb["A"] @ tb["B"], rtol=1e-10, atol=1e-15)
    assert np.allclose(tb["Plane"][1], np.linalg.matrix_power(tb["A"], 32), rtol=1e-9, atol=1e-14)
    rng = np.random.default_rng(5)
    n = 32 * 900
    t = np.arange(n) / sr
    x = (0.5 * np.sin(2 * np.pi * 55 * t) + 0.3 * np.sin(2 * np.pi * 3000 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32).astype(np.float64)
    ref, _ = sg.lfilter(b, a, x, zi=sg.lfilter_zi(b, a) * x[0])
    # float64 run of the realization == the filter (including the mapped zi)
    s = tb["zi"] * x[0]
    y64 = np.zeros(n)
    for i in range(n):
        y64[i] = tb["C"] @ s + tb["D"] * x[i]
        s = tb["A"] @ s + tb["B"] * x[i]
    assert np.max(np.abs(y64 - ref)) < 1e-10
    y32 = _chunked_f32_pass2(tb, x, x[0])
    err = np.max(np.abs(y32.astype(np.float64) - ref))
    assert err <= tol, (design, err)


_DYNEQ_DEFAULT = [(120, 1.0), (250, 1.2), (400, 1.0), (800, 1.2), (2500, 1.4), (5000, 1.4), (8000, 1.2), (12000, 0.8)]


def _scipy_band_kind(w0, bw):
    """What _safe_filtfilt (backend/app/pipeline.py:36-52) does with sg.iirpeak(w0, bw): 0 stable filtfilt, 2 ValueError ->
    lfilter of an H(z) = b0 section, 3 filtfilt with poles at +-1, -1 unstable filtfilt."""
    b, a = sg.iirpeak(w0, bw)
    degenerate = abs(a[1]) <= 1e-12 and abs(a[2] + 1.0) <= 1e-12
    try:
        sg.lfilter_zi(b, a)
    except ValueError:
        assert degenerate, (w0, bw, a)          # the only way an iirpeak section gets a pole at exactly z = 1
        return 2, b, a
    if degenerate:
        return 3, b, a
    return (0 if np.max(np.abs(np.roots(a))) < 1.0 else -1), b, a


def test_dynamic_eq_band_classes_match_scipy(lib):
    """mm_design_iirpeak reproduces scipy 1.18's iirpeak coefficients bit for bit and sorts the reference's bandwidth-as-Q
    sections (backend/app/pipeline.py:1657-1663) into the classes scipy's own filtfilt / lfilter_zi calls fall into: default
    bands (:1616-1625) at every common rate plus a sweep of q = 1 bands across the sum(a) == 0 boundary."""
    cases = []
    for sr in (22050, 32000, 44100, 48000, 88200, 96000, 192000):
        nyq = sr / 2.0
        for freq, q in _DYNEQ_DEFAULT + [(f, 1.0) for f in (1000, 3000, 6000, 7000, 8000, 9000, 15000)] + [(1000, 0.7), (3000, 0.5)]:
            if freq >= nyq * 0.98:
                continue
            w0 = float(np.clip(freq / nyq, 0.001, 0.98))
            cases.append((sr, freq, q, w0, float(np.clip(w0 / max(q, 0.1), 0.001, 0.5))))
    seen = set()
    for sr, freq, q, w0, bw in cases:
        kind_ref, b_ref, a_ref = _scipy_band_kind(w0, bw)
        b, a = (C.c_double * 3)(), (C.c_double * 3)()
        kind, rmax = C.c_int(99), C.c_double(0)
        assert lib.mm_design_iirpeak(w0, bw, b, a, C.byref(kind), C.byref(rmax)) == 0
        assert list(b) == list(b_ref) and list(a) == list(a_ref), (sr, freq, q)
        assert kind.value == kind_ref, (sr, freq, q, kind.value, kind_ref, list(a_ref))
        if kind_ref == -1:
            assert abs(rmax.value - np.max(np.abs(np.roots(a_ref)))) < 1e-9 * rmax.value and rmax.value > 1.0
        seen.add(kind_ref)
    assert seen == {0, 2, 3, -1}


def test_reference_dynamic_eq_default_bands_are_unstable():
    """The reference calls ``sg.iirpeak(w0, bw)`` with a bandwidth in the Q slot (backend/app/pipeline.py:1659-1663); for all
    eight default bands (:1616-1625) that is an unstable or marginal section at 44.1 and 48 kHz (how each
    class is reproduced)."""
    bands = _DYNEQ_DEFAULT
    for sr in (44100, 48000):
        nyq = sr / 2.0
        for freq, q in bands:
            w0 = float(np.clip(freq / nyq, 0.001, 0.98))
            bw = float(np.clip(w0 / max(q, 0.1), 0.001, 0.5))
            _, a = sg.iirpeak(w0, bw)
            assert np.max(np.abs(np.roots(a))) >= 1.0 - 1e-9, (sr, freq, q)
