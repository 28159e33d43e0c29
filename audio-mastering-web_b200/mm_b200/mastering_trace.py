"""Per-stage mastering trace (backend/app/mastering_trace.py): one structured log line per stage when
``MAGIC_MASTER_MASTERING_TRACE`` is on (``..._LUFS_STAGES`` adds the stage's integrated loudness).  Same logger name,
same ``mastering_trace key=value ...`` format and keys as the reference, so log tooling keeps working; the metrics
(peak over finite samples, NaN / Inf counts) are device reductions over the resident batch, not a copy to the host.
Tracing forces the stage-by-stage path (the fused chain has no stage boundaries to report)."""
from __future__ import annotations

import ctypes as C
import logging
import os
import re
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

_LOG = logging.getLogger("magic_master.mastering_trace")
if not _LOG.handlers:
    _h = logging.StreamHandler()
    _h.setLevel(logging.INFO)
    _h.setFormatter(logging.Formatter("%(levelname)s %(name)s: %(message)s"))
    _LOG.addHandler(_h)
_LOG.setLevel(logging.INFO)

_ENV_TRACE = "MAGIC_MASTER_MASTERING_TRACE"
_ENV_LUFS = "MAGIC_MASTER_MASTERING_TRACE_LUFS_STAGES"


def _on(name: str) -> bool:
    return os.environ.get(name, "").strip().lower() in ("1", "true", "yes", "on")


def trace_enabled() -> bool:
    return _on(_ENV_TRACE)


@dataclass
class TraceContext:
    job_id: str
    filename: str
    path: str  # v1 | v2 | telegram
    style: str = "standard"
    user_id: Optional[int] = None
    target_lufs: Optional[float] = None
    pro_flags: str = ""

    def __post_init__(self):
        base = os.path.basename(self.filename or "unknown")
        base = re.sub(r"[^\w.\-]+", "_", base, flags=re.UNICODE)
        self.filename = base[:120]


def _fmt_kv(**kwargs: Any) -> str:
    parts = []
    for k, v in kwargs.items():
        if v is None:
            continue
        parts.append(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}")
    return " ".join(parts)


def batch_metrics(eng, b) -> list:
    """signal_metrics for every track of a device batch (one reduction kernel)."""
    import torch
    from . import _lib
    with torch.cuda.stream(eng.stream):
        out = torch.empty(b.tracks * 3, dtype=torch.float64, device=eng.tdev)
        g = b.geom
        _lib.check(eng.lib.mm_dev_signal_metrics(eng.ctx, C.byref(g), b.ptr, C.c_void_p(out.data_ptr())))
        eng.sync()
        v = out.cpu().numpy().reshape(b.tracks, 3)
    res = []
    for t in range(b.tracks):
        peak = float(v[t, 0])
        res.append({"channels": b.channels, "samples": b.n, "duration_sec": round(b.n / float(b.sr), 4) if b.sr else 0.0,
                    "peak_linear": round(peak, 6), "peak_db": round(float(20.0 * np.log10(max(peak, 1e-12))), 2),
                    "nan_count": int(v[t, 1]) - int(v[t, 2]), "inf_count": int(v[t, 2]), "peak_raw": peak})      # kernel: non-finite, Inf
    return res


def trace_stage(ctx: Optional[TraceContext], stage: str, audio, sr: int, *, eng=None, **extra: Any) -> None:
    """``audio``: a device ``Batch`` (no copy) or a host array (uploaded for the reduction)."""
    if ctx is None or not trace_enabled():
        return
    from . import pipeline as P
    from .engine import Batch
    if isinstance(audio, Batch):
        b = audio
        eng = eng or P.get_engine()
    else:
        if np.asarray(audio).size == 0:
            return
        eng, b, _ = P._up(audio, sr)
    m = dict(batch_metrics(eng, b)[0])
    m.pop("peak_raw", None)                 # the unrounded peak is for validate_mastered_not_silent, not for the log line
    payload = {"job_id": ctx.job_id, "path": ctx.path, "filename": ctx.filename, "stage": stage, **m,
               **{k: v for k, v in extra.items() if v is not None}}
    if _on(_ENV_LUFS):
        lu = float(eng.measure_lufs(b)[0])
        if lu == lu:
            payload["lufs"] = round(lu, 3)
    _LOG.info("mastering_trace %s", _fmt_kv(**payload))
