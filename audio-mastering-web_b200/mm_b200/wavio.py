"""Minimal RIFF/WAVE packing for the export edge of the hot path.

The reference hands an already-quantised int16 array to libsndfile
(``backend/app/pipeline.py:899`` ``sf.write(buf, int16, sr, format="WAV", subtype="PCM_16")``)
and reads uploads with ``sf.read(..., dtype="float32", always_2d=True)`` (``:816``).
Container I/O is out of the hot path's scope; this module only packs
and unpacks canonical PCM WAV so that ``export_audio`` / ``load_audio_from_bytes`` keep their
byte-level contract without libsndfile.
"""
from __future__ import annotations

import struct

import numpy as np


def pack_wav_pcm16(int16_frames: np.ndarray, sr: int) -> bytes:
    """int16 array ``(n,)`` or ``(n, ch)`` -> canonical 44-byte-header PCM_16 WAV bytes."""
    a = np.ascontiguousarray(int16_frames, dtype="<i2")
    ch = 1 if a.ndim == 1 else a.shape[1]
    payload = a.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHH", 16, 1, ch, int(sr), int(sr) * ch * 2, ch * 2, 16)
    hdr += b"data" + struct.pack("<I", len(payload))
    return hdr + payload


def unpack_wav(data: bytes):
    """WAV bytes -> (float32 ``(n, ch)``, sr).  PCM 8/16/24/32 and IEEE float 32/64."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE stream")
    pos = 12
    fmt = None
    body = None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        (size,) = struct.unpack("<I", data[pos + 4:pos + 8])
        start = pos + 8
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", data[start:start + 16])
            if tag == 0xFFFE and size >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID's first 2 bytes
                (tag,) = struct.unpack("<H", data[start + 24:start + 26])
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            body = memoryview(data)[start:start + size]      # no copy: the frames are read in place
        pos = start + size + (size & 1)
    if fmt is None or body is None:
        raise ValueError("WAV stream lacks fmt or data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:
        if bits == 16:
            x = np.frombuffer(body[: len(body) // 2 * 2], dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 8:
            x = (np.frombuffer(body, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 24:
            raw = np.frombuffer(body[: len(body) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
            v = np.where(v & 0x800000, v - (1 << 24), v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(body[: len(body) // 4 * 4], dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"unsupported PCM width {bits}")
    elif tag == 3:
        dt = "<f4" if bits == 32 else "<f8"
        w = bits // 8
        x = np.frombuffer(body[: len(body) // w * w], dtype=dt).astype(np.float32)
    else:
        raise ValueError(f"unsupported WAV format tag {tag}")
    n = x.size // ch
    return np.ascontiguousarray(x[: n * ch].reshape(n, ch)), int(sr)


def pcm16_view(data: bytes):
    """Canonical PCM_16 WAV bytes -> (int16 ``(n, ch)`` view of the data chunk, n, ch, sr) without widening."""
    import struct as _st
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE stream")
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        (size,) = _st.unpack("<I", data[pos + 4:pos + 8])
        start = pos + 8
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = _st.unpack("<HHIIHH", data[start:start + 16])
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            body = memoryview(data)[start:start + size]      # no copy: the frames are read in place
        pos = start + size + (size & 1)
    if fmt is None or body is None or fmt[0] != 1 or fmt[3] != 16:
        raise ValueError("pcm16_view: a PCM_16 WAV is required")
    ch, sr = fmt[1], fmt[2]
    if ch not in (1, 2) or sr <= 0:
        raise ValueError(f"pcm16_view: {ch} channels at {sr} Hz (mono or stereo PCM_16 required)")
    x = np.frombuffer(body[: len(body) // (2 * ch) * 2 * ch], dtype="<i2").reshape(-1, ch)
    return x, x.shape[0], ch, sr
