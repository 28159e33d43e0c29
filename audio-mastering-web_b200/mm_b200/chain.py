"""Drop-in for the reference's ``backend/app/chain.py`` + ``backend/app/modules/*.py`` (v2 chain).

``MasteringChain.from_config`` / ``default_config`` / ``default_chain`` / ``process`` keep the
reference's signatures and semantics (chain.py:40-134): modules run in list order, disabled modules
pass audio through, ``amount < 1`` blends processed with dry (modules/base.py:32-46), ``target_lufs`` /
``style`` keyword arguments override the module's own parameters (modules/normalize_lufs.py:30-32,
modules/equalizer.py:83-85), and the result is clipped to +-1 with nan_to_num (chain.py:93-94).

The audio is uploaded once, every module is a CUDA stage call on the device-resident batch, and the
result is downloaded once.  When the configuration is exactly ``default_config(target, style)`` the
whole chain runs as the fused sweep plan (``mm_dev_master``), which is what ``bench.py`` times.
Every module option of the reference has a device stage (oversampled exciter, 4-band / Haas imager, reverb included).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Optional

import numpy as np

from . import _lib
from . import pipeline as P
from .engine import Batch, Engine, MMError, get_engine, style_struct


class BaseModule:
    """modules/base.py:7-50 on a device batch."""

    module_id = "base"

    def __init__(self, enabled: bool = True, amount: float = 1.0, ms_mode: str = "both", **kwargs: Any):
        self.enabled = bool(enabled)
        self.amount = float(np.clip(amount, 0.0, 1.0))
        self.ms_mode = str(ms_mode)
        self.params = kwargs

    @classmethod
    def from_config(cls, config: dict) -> "BaseModule":
        return cls(**config)

    # -- numpy edge (same call shape as the reference) ------------------------------------------------
    def process(self, audio: np.ndarray, sr: int, **kwargs: Any) -> np.ndarray:
        if not self.enabled:
            return audio
        eng, b, mono = P._up(audio, sr)
        out = self.process_batch(eng, b, **kwargs)
        return audio if out is b else P._down(eng, out, mono)

    # -- device path ------------------------------------------------------------------------------------
    def process_batch(self, eng: Engine, b: Batch, **kwargs: Any) -> Batch:
        if not self.enabled:
            return b
        try:
            processed = self._process(eng, b, **kwargs)
        except (NotImplementedError, MMError, ImportError):
            raise                      # never hide a missing kernel or a device failure
        except Exception:
            return b                   # modules/base.py:40-43: any other failure passes the audio through
        if self.amount >= 1.0 or processed is b:
            return processed
        g = b.geom
        _lib.check(eng.lib.mm_dev_blend(eng.ctx, C.byref(g), b.ptr, processed.ptr, processed.ptr, C.c_double(self.amount)))
        return processed

    def _process(self, eng: Engine, b: Batch, **kwargs: Any) -> Batch:
        return b


def _d(v):
    return C.c_double(float(v))


class DCOffsetModule(BaseModule):
    module_id = "dc_offset"

    def _process(self, eng, b, **kw):
        return eng.stage("remove_dc_offset", b)


class PeakGuardModule(BaseModule):
    module_id = "peak_guard"

    def __init__(self, enabled=True, amount=1.0, headroom_db=0.5, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.headroom_db = float(headroom_db)

    def _process(self, eng, b, **kw):
        return eng.stage("remove_intersample_peaks", b, _d(self.headroom_db))


class TargetCurveModule(BaseModule):
    module_id = "target_curve"

    def __init__(self, enabled=True, amount=1.0, ms_mode="both", phase_mode="minimum", eq_ms=False, **kwargs):
        super().__init__(enabled=enabled, amount=amount, ms_mode=ms_mode, **kwargs)
        self.phase_mode, self.eq_ms = str(phase_mode), bool(eq_ms)

    def _process(self, eng, b, **kw):
        ms = bool(kw.get("eq_ms", self.eq_ms)) and b.channels == 2
        if kw.get("phase_mode", self.phase_mode) == "linear_phase":
            return eng.stage("apply_target_curve_linear_phase", b, 1 if ms else 0)
        return eng.stage("apply_target_curve", b, 1 if ms else 0)


class DynamicsModule(BaseModule):
    module_id = "dynamics"

    def __init__(self, enabled=True, amount=1.0, knee_db=6.0, crossovers_hz=None, band_ratios=None, max_upward_boost_db=12.0, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.knee_db = float(knee_db)
        self.crossovers_hz = tuple(float(x) for x in crossovers_hz) if crossovers_hz else None
        self.band_ratios = tuple(float(x) for x in band_ratios) if band_ratios else None
        self.max_upward_boost_db = float(max_upward_boost_db)

    def _process(self, eng, b, **kw):
        cx = _lib.darr(self.crossovers_hz) if self.crossovers_hz and len(self.crossovers_hz) == 3 else None
        br = _lib.darr(self.band_ratios) if self.band_ratios and len(self.band_ratios) == 4 else None
        return eng.stage("apply_dynamics_mode", b, _d(self.knee_db), cx, br, _d(self.max_upward_boost_db), 0, P._compressor_id())


class MaximizerModule(BaseModule):
    module_id = "maximizer"

    def __init__(self, enabled=True, amount=1.0, sensitivity=0.5, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.sensitivity = float(sensitivity)

    def _process(self, eng, b, **kw):
        return eng.stage("apply_maximizer_transient_aware", b, _d(self.sensitivity))      # modules/maximizer.py -> pipeline.py:521-545


class NormalizeLUFSModule(BaseModule):
    module_id = "normalize_lufs"

    def __init__(self, enabled=True, amount=1.0, target_lufs=-14.0, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.target_lufs = float(target_lufs)

    def _process(self, eng, b, **kw):
        if b.n < 0.4 * b.sr:
            return b                                   # pipeline.py:647-650: the meter cannot run
        target = float(kw.get("target_lufs", self.target_lufs))
        return eng.stage("normalize_lufs", b, _lib.darr([target] * b.tracks))


class FinalSpectralBalanceModule(BaseModule):
    module_id = "final_spectral_balance"

    def _process(self, eng, b, **kw):
        return eng.stage("apply_final_spectral_balance", b)


class StyleEQModule(BaseModule):
    module_id = "style_eq"

    def __init__(self, enabled=True, amount=1.0, style="standard", **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.style = str(style)

    def _process(self, eng, b, **kw):
        cfg = P.STYLE_CONFIGS.get(kw.get("style", self.style), P.STYLE_CONFIGS["standard"])
        return eng.stage("apply_style_eq", b, _lib.darr([cfg[k] for k in ("sub", "bass", "mids", "presence", "air")]))


class ExciterModule(BaseModule):
    module_id = "exciter"

    def __init__(self, enabled=True, amount=1.0, exciter_db=0.0, mode="warm", oversample=1, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.exciter_db, self.mode, self.oversample = float(exciter_db), str(mode), int(oversample)

    def _process(self, eng, b, **kw):
        return P._exciter_dev(eng, b, self.exciter_db, self.mode, self.oversample)     # modules/exciter.py -> pipeline.py:1267-1326


class ImagerModule(BaseModule):
    module_id = "imager"

    def __init__(self, enabled=True, amount=1.0, width=1.0, stereoize_delay_ms=0.0, stereoize_mix=0.12, band_widths=None,
                 crossovers_hz=None, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.width, self.stereoize_delay_ms, self.stereoize_mix = float(width), float(stereoize_delay_ms or 0.0), float(stereoize_mix)
        self.band_widths = list(band_widths) if band_widths else None
        self.crossovers_hz = tuple(float(x) for x in crossovers_hz) if crossovers_hz else None

    def _process(self, eng, b, **kw):      # modules/imaging.py:44-53 -> pipeline.py:1339-1398
        return P._imager_dev(eng, b, self.width, self.stereoize_delay_ms, self.stereoize_mix, self.band_widths, self.crossovers_hz)


class ReverbModule(BaseModule):
    module_id = "reverb"

    def __init__(self, enabled=False, amount=1.0, reverb_type="plate", decay_sec=1.2, mix=0.15, mix_mid=None, mix_side=None, **kwargs):
        super().__init__(enabled=enabled, amount=amount, **kwargs)
        self.reverb_type, self.decay_sec, self.mix = str(reverb_type), float(decay_sec), float(mix)
        self.mix_mid = float(mix_mid) if mix_mid is not None else None
        self.mix_side = float(mix_side) if mix_side is not None else None

    def _process(self, eng, b, **kw):      # modules/reverb.py -> pipeline.py:1119-1176
        use_ms = b.channels == 2 and (self.mix_mid is not None or self.mix_side is not None)
        return eng.stage("apply_reverb", b, P._REVERB_TYPES.get(self.reverb_type, 0), _d(self.decay_sec), _d(self.mix), 1 if use_ms else 0,
                         _d(self.mix_mid if self.mix_mid is not None else self.mix), _d(self.mix_side if self.mix_side is not None else self.mix))


MODULE_REGISTRY = {m.module_id: m for m in (
    DCOffsetModule, PeakGuardModule, TargetCurveModule, DynamicsModule, MaximizerModule, NormalizeLUFSModule,
    FinalSpectralBalanceModule, StyleEQModule, ExciterModule, ImagerModule, ReverbModule)}


class MasteringChain:
    """chain.py:40-134."""

    def __init__(self, modules, config: Optional[dict] = None):
        self.modules = modules
        self._config = config

    @classmethod
    def from_config(cls, config: dict) -> "MasteringChain":
        modules = []
        for item in config.get("modules", []):
            item = dict(item)
            mid = item.pop("id", None)
            if not mid or mid not in MODULE_REGISTRY:
                continue
            modules.append(MODULE_REGISTRY[mid].from_config(item))
        return cls(modules, config=config)

    @classmethod
    def default_config(cls, target_lufs: float = -14.0, style: str = "standard") -> dict:
        cfg = P.STYLE_CONFIGS.get(style, P.STYLE_CONFIGS["standard"])
        exciter_db, width = cfg.get("exciter_db", 0.0), cfg.get("imager_width", 1.0)
        return {"modules": [
            {"id": "dc_offset", "enabled": True, "amount": 1.0},
            {"id": "peak_guard", "enabled": True, "headroom_db": 0.5, "amount": 1.0},
            {"id": "target_curve", "enabled": True, "phase_mode": "minimum", "eq_ms": False, "amount": 1.0},
            {"id": "dynamics", "enabled": True, "knee_db": 6.0, "crossovers_hz": [214.0, 2230.0, 10000.0], "amount": 1.0},
            {"id": "normalize_lufs", "enabled": True, "target_lufs": target_lufs, "amount": 1.0},
            {"id": "final_spectral_balance", "enabled": True, "amount": 1.0},
            {"id": "style_eq", "enabled": True, "style": style, "amount": 1.0},
            {"id": "exciter", "enabled": abs(exciter_db) >= 0.05, "exciter_db": exciter_db, "mode": "warm", "oversample": 1, "amount": 1.0},
            {"id": "imager", "enabled": abs(width - 1.0) >= 0.01, "width": width, "stereoize_delay_ms": 0.0, "stereoize_mix": 0.12,
             "band_widths": None, "crossovers_hz": [214.0, 2230.0, 10000.0], "amount": 1.0},
            {"id": "reverb", "enabled": False, "reverb_type": "plate", "decay_sec": 1.2, "mix": 0.15, "mix_mid": None, "mix_side": None, "amount": 1.0},
            {"id": "peak_guard", "enabled": True, "headroom_db": 0.5, "amount": 1.0},
        ]}

    @classmethod
    def default_chain(cls, target_lufs: float = -14.0, style: str = "standard") -> "MasteringChain":
        return cls.from_config(cls.default_config(target_lufs=target_lufs, style=style))

    def _fused_plan(self, target_lufs, style):
        """(target, style) when this chain is exactly the default one for them, else None."""
        if self._config is None:
            return None
        try:
            mods = self._config["modules"]
            st = style if style is not None else next(m["style"] for m in mods if m.get("id") == "style_eq")
            tg = target_lufs if target_lufs is not None else next(m["target_lufs"] for m in mods if m.get("id") == "normalize_lufs")
        except (KeyError, StopIteration):
            return None
        if st not in P.STYLE_CONFIGS:
            return None
        ref = self.default_config(target_lufs=float(tg), style=st)
        mine = {"modules": [dict(m) for m in mods]}
        for m in mine["modules"]:                 # process() kwargs override these two fields anyway
            if m.get("id") == "normalize_lufs":
                m["target_lufs"] = float(tg)
            if m.get("id") == "style_eq":
                m["style"] = st
        # the exciter / imager entries of the default config depend on the style the chain was built with
        return (float(tg), st) if mine == ref else None

    def process(self, audio: np.ndarray, sr: int, *, target_lufs: Optional[float] = None, style: Optional[str] = None,
                progress_callback: Optional[Callable[[int, str], None]] = None, trace_ctx=None, **kwargs: Any) -> np.ndarray:
        total = len(self.modules)
        from . import mastering_trace as _mt
        tracing = trace_ctx is not None and _mt.trace_enabled()      # the trace needs the module boundaries: no fusion
        plan = self._fused_plan(target_lufs, style) if (not kwargs and not tracing) else None
        if plan is not None:
            # the reference reports one progress tick per module (chain.py:80-82); same ticks, fused execution
            if progress_callback:
                for i, mod in enumerate(self.modules):
                    progress_callback(5 + int(90 * (i / total)), getattr(mod, "module_id", "module"))
            out = P.master_batch([audio], sr, [plan[1]], [plan[0]], chain="v2", job_fade=False)["audio"][0]
        else:
            eng, b, mono = P._up(audio, sr)
            for i, mod in enumerate(self.modules):
                if progress_callback and total > 0:
                    progress_callback(5 + int(90 * (i / total)), getattr(mod, "module_id", "module"))
                kw = dict(kwargs)
                if target_lufs is not None:
                    kw["target_lufs"] = target_lufs
                if style is not None:
                    kw["style"] = style
                b = mod.process_batch(eng, b, **kw)
                if tracing:       # chain.py:92: device reduction over the resident batch, no copy
                    _mt.trace_stage(trace_ctx, getattr(mod, "module_id", "module"), b, sr, eng=eng)
            b = eng.stage("finalize_clip", b, out=b)            # chain.py:93-94: clip + nan_to_num, on the device
            out = P._down(eng, b, mono)
            if tracing:
                _mt.trace_stage(trace_ctx, "chain_finalize_clip", out, sr)
        if progress_callback:
            progress_callback(98, "Готово")
        P._trim_engine()          # a worker thread does not keep more than MM_WORKSPACE_KEEP_MB of scratch between jobs
        return out
