"""Host side of the CUDA path: a context, device-resident planar batches, stage calls.

PyTorch is used for device memory and streams only; every computation is a call into
``libmm_b200.so`` through the C ABI (``include/mm_b200.h``).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import Geom, Style, TrackStats, MMError  # noqa: F401  (re-exported)


def _torch_np_dtype(dt):
    torch = _torch()
    return {torch.float32: np.float32, torch.float64: np.float64, torch.int16: np.int16, torch.int32: np.int32,
            torch.int64: np.int64, torch.uint8: np.uint8}[dt]


def _torch():
    import torch
    return torch


class Batch:
    """Planar device batch: rows ``[track*channels + c]`` of ``stride`` floats, sample i at ``MM_LEAD + i``."""

    def __init__(self, tensor, tracks, channels, n, sr):
        self.t = tensor                       # torch float32 (rows, stride)
        self.tracks, self.channels, self.n, self.sr = int(tracks), int(channels), int(n), int(sr)

    @property
    def stride(self):
        return int(self.t.shape[1])

    @property
    def geom(self):
        return Geom(self.n, self.stride, self.tracks, self.channels, self.sr, 0)

    @property
    def ptr(self):
        return C.c_void_p(self.t.data_ptr())

    def live(self):
        """View of the samples only: (rows, n)."""
        return self.t[:, _lib.MM_LEAD:_lib.MM_LEAD + self.n]


class Engine:
    """One CUDA context (stream + workspace) on one device.  Not shared between threads: use
    :func:`get_engine` for a per-thread instance."""

    def __init__(self, device: int = 0):
        torch = _torch()
        if not torch.cuda.is_available():
            raise MMError("no CUDA device: mm_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self.stream = torch.cuda.Stream(device=self.tdev)
        ctx = C.c_void_p()
        _lib.check(self.lib.mm_ctx_create(self.device, C.c_void_p(self.stream.cuda_stream), C.byref(ctx)))
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.mm_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- memory ---------------------------------------------------------------------------------
    def row_stride(self, n: int) -> int:
        return int(self.lib.mm_row_stride(int(n)))

    def empty(self, tracks, channels, n, sr) -> Batch:
        torch = _torch()
        with torch.cuda.stream(self.stream):
            t = torch.empty((tracks * channels, self.row_stride(n)), dtype=torch.float32, device=self.tdev)
        return Batch(t, tracks, channels, n, sr)

    def like(self, b: Batch) -> Batch:
        return self.empty(b.tracks, b.channels, b.n, b.sr)

    def upload(self, tracks_np, sr) -> Batch:
        """list of (n,) / (n, ch) float32 arrays of identical shape (or one 3-D array) -> Batch."""
        torch = _torch()
        views = [np.asarray(a, dtype=np.float32).reshape(len(a), -1) for a in tracks_np]
        if any(v.shape != views[0].shape for v in views):
            raise ValueError("all input arrays must have the same shape")
        # no host-side gathering: every caller's array goes to the device as it is
        views = [np.ascontiguousarray(v) for v in views]
        tracks, (n, ch) = len(views), views[0].shape
        b = self.empty(tracks, ch, n, sr)
        with torch.cuda.stream(self.stream):
            il = torch.empty((tracks, n, ch), dtype=torch.float32, device=self.tdev)
            # a caller's (pageable) array is staged into pinned memory by several threads inside the library while the DMA engine
            # already moves the finished blocks (mm_ctx_copy_in): a 3-minute track took ~8 ms through one numpy copy + one DMA
            for t, v in enumerate(views):
                _lib.check(self.lib.mm_ctx_copy_in(self.ctx, C.c_void_p(il[t].data_ptr()), C.c_void_p(v.ctypes.data), v.nbytes))
            g = b.geom
            _lib.check(self.lib.mm_dev_deinterleave(self.ctx, C.byref(g), C.c_void_p(il.data_ptr()), b.ptr))
            self.sync()
        return b

    def download(self, b: Batch) -> np.ndarray:
        """Batch -> float32 array (tracks, n, channels)."""
        torch = _torch()
        with torch.cuda.stream(self.stream):
            il = torch.empty((b.tracks, b.n, b.channels), dtype=torch.float32, device=self.tdev)
            g = b.geom
            _lib.check(self.lib.mm_dev_interleave(self.ctx, C.byref(g), b.ptr, C.c_void_p(il.data_ptr())))
            host = np.empty((b.tracks, b.n, b.channels), np.float32)       # the caller owns its result
            _lib.check(self.lib.mm_ctx_copy_out(self.ctx, C.c_void_p(host.ctypes.data), C.c_void_p(il.data_ptr()), host.nbytes))
        return host

    def to_host(self, t) -> np.ndarray:
        """A device tensor of this engine's stream -> a numpy array the caller owns (staged, see mm_ctx_copy_out)."""
        if not t.is_contiguous():
            with _torch().cuda.stream(self.stream):
                t = t.contiguous()
        host = np.empty(tuple(t.shape), dtype=_torch_np_dtype(t.dtype))
        if host.nbytes:
            _lib.check(self.lib.mm_ctx_copy_out(self.ctx, C.c_void_p(host.ctypes.data), C.c_void_p(t.data_ptr()), host.nbytes))
        return host

    def release_workspace(self):
        """Give the device scratch and the pinned staging buffers back (they grow on demand and stay at their high-water mark
        otherwise)."""
        _lib.check(self.lib.mm_ctx_release_workspace(self.ctx))

    def set_lanes(self, lanes: int):
        """0 = automatic, 1 = one stream, up to 8 (mm_ctx_set_lanes): concurrent sub-batches / chunks of one call."""
        _lib.check(self.lib.mm_ctx_set_lanes(self.ctx, int(lanes)))

    def sync(self):
        _lib.check(self.lib.mm_ctx_sync(self.ctx))

    # ---- generic stage call -----------------------------------------------------------------------
    def stage(self, name: str, src: Batch, *args, out: Batch | None = None) -> Batch:
        """Call ``mm_dev_<name>(ctx, geom, in, out, *args)``; returns the output batch."""
        dst = out if out is not None else self.like(src)
        g = src.geom
        fn = getattr(self.lib, "mm_dev_" + name)
        _lib.check(fn(self.ctx, C.byref(g), src.ptr, dst.ptr, *args))
        return dst

    def fft_resample(self, src: Batch, num: int, sr: int | None = None) -> Batch:
        """scipy.signal.resample(row, num) of every row (csrc/bigfft.cu); ``sr`` labels the new batch."""
        num = int(num)
        if num == src.n:
            return src
        dst = self.empty(src.tracks, src.channels, num, sr if sr is not None else src.sr)
        gi, go = src.geom, dst.geom
        _lib.check(self.lib.mm_dev_fft_resample(self.ctx, C.byref(gi), src.ptr, C.byref(go), dst.ptr))
        return dst

    # ---- reductions / analyzers -------------------------------------------------------------------
    def _doubles(self, count):
        torch = _torch()
        with torch.cuda.stream(self.stream):
            return torch.empty(count, dtype=torch.float64, device=self.tdev)

    def _to_host(self, t):
        self.sync()
        with _torch().cuda.stream(self.stream):
            return t.cpu().numpy()

    def measure_lufs(self, b: Batch) -> np.ndarray:
        out = self._doubles(b.tracks)
        g = b.geom
        _lib.check(self.lib.mm_dev_measure_lufs(self.ctx, C.byref(g), b.ptr, C.c_void_p(out.data_ptr())))
        return self._to_host(out)

    def true_peak(self, b: Batch) -> np.ndarray:
        out = self._doubles(b.tracks)
        g = b.geom
        _lib.check(self.lib.mm_dev_true_peak(self.ctx, C.byref(g), b.ptr, C.c_void_p(out.data_ptr())))
        return self._to_host(out)

    def spectrum_bars(self, b: Batch, view: int = 0) -> np.ndarray:
        out = self._doubles(b.tracks * 64)
        g = b.geom
        _lib.check(self.lib.mm_dev_spectrum_bars(self.ctx, C.byref(g), b.ptr, int(view), C.c_void_p(out.data_ptr())))
        return self._to_host(out).reshape(b.tracks, 64)

    def stereo_correlation(self, b: Batch):
        corr, peak = self._doubles(b.tracks), self._doubles(b.tracks)
        g = b.geom
        _lib.check(self.lib.mm_dev_stereo_correlation(self.ctx, C.byref(g), b.ptr, C.c_void_p(corr.data_ptr()),
                                                      C.c_void_p(peak.data_ptr())))
        return self._to_host(corr), self._to_host(peak)

    def auto_blank_end(self, b: Batch, threshold_dbfs: float, min_silence_sec: float) -> Batch:
        """_auto_blank_end (pipeline.py:900-918) for a one-track batch: a view of ``b`` cut after the last frame above the
        threshold plus ``min_silence_sec`` (the scan runs on the device; the cut is a shorter geometry over the same rows)."""
        torch = _torch()
        n_silence = int(b.sr * min_silence_sec)
        if b.n == 0 or min_silence_sec <= 0 or n_silence <= 0 or b.tracks != 1:
            return b
        with torch.cuda.stream(self.stream):
            idx = torch.empty(b.tracks, dtype=torch.int64, device=self.tdev)
            g = b.geom
            _lib.check(self.lib.mm_dev_last_above(self.ctx, C.byref(g), b.ptr, float(10 ** (threshold_dbfs / 20.0)), C.c_void_p(idx.data_ptr())))
            self.sync()
            last = int(idx.cpu()[0])
        keep = b.n if last < 0 else min(b.n, last + 1 + n_silence)
        return Batch(b.t, b.tracks, b.channels, keep, b.sr)

    def true_peak_correlation(self, b: Batch):
        """true peak (dBTP), stereo correlation and sample peak per track from one pass over the samples."""
        tp, corr, peak = self._doubles(b.tracks), self._doubles(b.tracks), self._doubles(b.tracks)
        g = b.geom
        _lib.check(self.lib.mm_dev_true_peak_correlation(self.ctx, C.byref(g), b.ptr, C.c_void_p(tp.data_ptr()), C.c_void_p(corr.data_ptr()),
                                                         C.c_void_p(peak.data_ptr())))
        return self._to_host(tp), self._to_host(corr), self._to_host(peak)

    def quantize_pcm24(self, b: Batch) -> np.ndarray:
        """-> int32 (tracks, n, channels) holding 24-bit samples (libsndfile's float -> PCM_24 conversion)."""
        torch = _torch()
        with torch.cuda.stream(self.stream):
            pcm = torch.empty((b.tracks, b.n, b.channels), dtype=torch.int32, device=self.tdev)
            g = b.geom
            _lib.check(self.lib.mm_dev_quantize_pcm24(self.ctx, C.byref(g), b.ptr, C.c_void_p(pcm.data_ptr())))
            return self.to_host(pcm)

    def quantize_int16(self, b: Batch, noise: np.ndarray | None = None, seed: int = 0) -> np.ndarray:
        """-> int16 (tracks, n, channels); ``noise`` float32 of that shape selects the bit-exact mode."""
        torch = _torch()
        with torch.cuda.stream(self.stream):
            pcm = torch.empty((b.tracks, b.n, b.channels), dtype=torch.int16, device=self.tdev)
            nz = None
            if noise is not None:
                nz = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float32).reshape(b.tracks, b.n, b.channels)).to(self.tdev)
            g = b.geom
            _lib.check(self.lib.mm_dev_quantize_int16(self.ctx, C.byref(g), b.ptr, C.c_void_p(pcm.data_ptr()),
                                                      C.c_void_p(nz.data_ptr()) if nz is not None else None, int(seed)))
            return self.to_host(pcm)

    def quantize_int16_shaped(self, b: Batch, shape: int, uniform: np.ndarray | None = None, seed: int = 0) -> np.ndarray:
        """Noise-shaped dither export (shape 1 = ns_e, 2 = ns_itu) -> int16 (tracks, n, channels); ``uniform`` =
        ``np.random.rand(n, ch).astype(float32)`` per track selects the bit-exact mode."""
        torch = _torch()
        with torch.cuda.stream(self.stream):
            pcm = torch.empty((b.tracks, b.n, b.channels), dtype=torch.int16, device=self.tdev)
            un = None
            if uniform is not None:
                un = torch.from_numpy(np.ascontiguousarray(uniform, dtype=np.float32).reshape(b.tracks, b.n, b.channels)).to(self.tdev)
            g = b.geom
            _lib.check(self.lib.mm_dev_quantize_int16_shaped(self.ctx, C.byref(g), b.ptr, C.c_void_p(pcm.data_ptr()),
                                                             C.c_void_p(un.data_ptr()) if un is not None else None, int(seed), int(shape)))
            return self.to_host(pcm)

    # ---- whole chains -----------------------------------------------------------------------------
    def master(self, src: Batch, chain: int, styles, *, out: Batch | None = None, want_int16=False, noise=None,
               seed: int = 0, flags: int = 0, want_stats=True):
        """Run ``mm_dev_master``.  ``styles``: list of ``Style`` (one per track).
        Returns (out_batch, pcm_tensor_or_None, stats_array_or_None)."""
        torch = _torch()
        dst = out if out is not None else self.like(src)
        arr = (Style * src.tracks)(*styles)
        with torch.cuda.stream(self.stream):
            pcm = torch.empty((src.tracks, src.n, src.channels), dtype=torch.int16, device=self.tdev) if want_int16 else None
            nz = None
            if noise is not None:
                nz = noise if torch.is_tensor(noise) else torch.from_numpy(
                    np.ascontiguousarray(noise, dtype=np.float32).reshape(src.tracks, src.n, src.channels)).to(self.tdev)
            st = torch.empty(src.tracks * C.sizeof(TrackStats), dtype=torch.uint8, device=self.tdev) if want_stats else None
            g = src.geom
            _lib.check(self.lib.mm_dev_master(
                self.ctx, C.byref(g), int(chain), arr, src.ptr, dst.ptr,
                C.c_void_p(pcm.data_ptr()) if pcm is not None else None,
                C.c_void_p(nz.data_ptr()) if nz is not None else None, int(seed),
                C.c_void_p(st.data_ptr()) if st is not None else None, int(flags)))
            stats = None
            if st is not None:
                self.sync()
                raw = st.cpu().numpy().tobytes()
                stats = (TrackStats * src.tracks).from_buffer_copy(raw)
        return dst, pcm, stats

    def kernel_times(self):
        cap = 256
        buf = (_lib.KTime * cap)()
        cnt = C.c_int(0)
        _lib.check(self.lib.mm_ctx_kernel_times(self.ctx, buf, cap, C.byref(cnt)))
        return {buf[i].name.decode(): (buf[i].ms, buf[i].launches, buf[i].samples) for i in range(min(cnt.value, cap))}

    def timing(self, on: bool):
        _lib.check(self.lib.mm_ctx_timing(self.ctx, 1 if on else 0))

    def launch_count(self) -> int:
        return int(self.lib.mm_ctx_launch_count(self.ctx))


_tls = threading.local()


def get_engine(device: int = 0) -> Engine:
    """Per-thread engine (the reference's job functions run in worker threads)."""
    engines = getattr(_tls, "engines", None)
    if engines is None:
        engines = _tls.engines = {}
    if device not in engines:
        engines[device] = Engine(device)
    return engines[device]


def style_struct(cfg: dict, target_lufs: float) -> Style:
    """STYLE_CONFIGS row (dict) + target -> C struct."""
    s = Style()
    s.target_lufs = float(target_lufs)
    for i, k in enumerate(("sub", "bass", "mids", "presence", "air")):
        s.eq_gain_db[i] = float(cfg.get(k, 0.0))
    s.exciter_db = float(cfg.get("exciter_db", 0.0))
    s.imager_width = float(cfg.get("imager_width", 1.0))
    s.parallel_mix = float(cfg.get("parallel_mix", 0.0))
    return s
