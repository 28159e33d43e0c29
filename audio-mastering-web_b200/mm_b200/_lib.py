"""ctypes binding of ``include/mm_b200.h`` (the C ABI of the CUDA library).

There is no CPU fallback: if ``libmm_b200.so`` is missing the import fails loudly, and
``mm_ctx_create`` fails when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MM_B200_LIB") or os.path.join(HERE, "libmm_b200.so")     # MM_B200_LIB: an experiment build (build.py --tag)


class MMError(RuntimeError):
    """A C-ABI call returned non-zero; the message is ``mm_last_error()``."""


class Geom(C.Structure):
    _fields_ = [("n", C.c_int64), ("stride", C.c_int64), ("tracks", C.c_int32), ("channels", C.c_int32),
                ("sr", C.c_int32), ("track_base", C.c_int32)]


class Style(C.Structure):
    _fields_ = [("target_lufs", C.c_double), ("eq_gain_db", C.c_double * 5), ("exciter_db", C.c_double),
                ("imager_width", C.c_double), ("parallel_mix", C.c_double)]


class TrackStats(C.Structure):
    _fields_ = [("lufs_in", C.c_double), ("lufs_mid", C.c_double), ("lufs_out", C.c_double), ("gain_db", C.c_double),
                ("peak_in", C.c_double), ("peak_out", C.c_double), ("mean", C.c_double * 2), ("nonfinite", C.c_double)]


class HostJob(C.Structure):
    """mm_host_job: one upload of mm_master_host_jobs (its own shape, host buffers, style and dither stream)."""
    _fields_ = [("n", C.c_int64), ("channels", C.c_int32), ("sr", C.c_int32), ("audio_in", C.c_void_p), ("pcm16_in", C.c_void_p),
                ("audio_out", C.c_void_p), ("pcm16_out", C.c_void_p), ("style", Style), ("dither_id", C.c_int32),
                ("reserved", C.c_int32), ("stats", TrackStats)]


class KTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_double), ("launches", C.c_int64), ("samples", C.c_double)]


MM_LEAD = 32
CHAIN_V1, CHAIN_V2 = 1, 2
FLAG_MEASURE_IN, FLAG_MEASURE_OUT, FLAG_NO_JOB_FADE, FLAG_ENVELOPE_COMPRESSOR = 1, 2, 4, 8
COMPRESSOR_SOFT_KNEE, COMPRESSOR_ENVELOPE = 0, 1
LOWPASS, HIGHPASS, BANDPASS = 0, 1, 2
DYNEQ_STRICT = 1

_vp, _i, _i64, _u64, _u32, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double
_dp = C.POINTER(C.c_double)
_gp = C.POINTER(Geom)

# name -> (restype, argtypes); every symbol include/mm_b200.h declares
SIGNATURES = {
    "mm_abi_version": (_i, []),
    "mm_last_error": (C.c_char_p, []),
    "mm_row_stride": (_i64, [_i64]),
    "mm_ctx_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "mm_ctx_destroy": (None, [_vp]),
    "mm_ctx_sync": (_i, [_vp]),
    "mm_ctx_release_workspace": (_i, [_vp]),
    "mm_ctx_launch_count": (_i64, [_vp]),
    "mm_ctx_timing": (_i, [_vp, _i]),
    "mm_ctx_kernel_times": (_i, [_vp, C.POINTER(KTime), _i, C.POINTER(_i)]),
    "mm_dev_deinterleave": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_interleave": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_remove_dc_offset": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_remove_intersample_peaks": (_i, [_vp, _gp, _vp, _vp, _d]),
    "mm_dev_fade_in": (_i, [_vp, _gp, _vp, _vp, _d]),
    "mm_dev_blend": (_i, [_vp, _gp, _vp, _vp, _vp, _d]),
    "mm_dev_apply_target_curve": (_i, [_vp, _gp, _vp, _vp, _i]),
    "mm_dev_apply_deesser": (_i, [_vp, _gp, _vp, _vp, _d, _d, _d, _d, _d, _d]),
    "mm_dev_apply_dynamics": (_i, [_vp, _gp, _vp, _vp, _d, _dp, _dp, _d]),
    "mm_dev_apply_dynamics_mode": (_i, [_vp, _gp, _vp, _vp, _d, _dp, _dp, _d, _i, _i]),
    "mm_dev_apply_multiband_dynamics": (_i, [_vp, _gp, _vp, _vp, _d, _dp, _dp, _d]),
    "mm_dev_apply_maximizer_lookahead": (_i, [_vp, _gp, _vp, _vp, _d]),
    "mm_dev_apply_maximizer": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_apply_parallel_compression": (_i, [_vp, _gp, _vp, _vp, _d, _d, _d]),
    "mm_dev_measure_lufs": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_normalize_lufs": (_i, [_vp, _gp, _vp, _vp, _dp]),
    "mm_dev_apply_final_spectral_balance": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_apply_style_eq": (_i, [_vp, _gp, _vp, _vp, _dp]),
    "mm_dev_apply_harmonic_exciter": (_i, [_vp, _gp, _vp, _vp, _d, _i]),
    "mm_dev_apply_stereo_imager": (_i, [_vp, _gp, _vp, _vp, _d]),
    "mm_dev_apply_rumble_filter": (_i, [_vp, _gp, _vp, _vp, _d]),
    "mm_dev_iir": (_i, [_vp, _gp, _vp, _vp, _dp, _dp, _i, _i]),
    "mm_dev_finalize_clip": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_last_above": (_i, [_vp, _gp, _vp, _d, _vp]),
    "mm_dev_quantize_pcm24": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_quantize_int16": (_i, [_vp, _gp, _vp, _vp, _vp, _u64]),
    "mm_dev_quantize_int16_shaped": (_i, [_vp, _gp, _vp, _vp, _vp, _u64, _i]),
    "mm_dev_true_peak": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_spectrum_bars": (_i, [_vp, _gp, _vp, _i, _vp]),
    "mm_dev_stereo_correlation": (_i, [_vp, _gp, _vp, _vp, _vp]),
    "mm_dev_true_peak_correlation": (_i, [_vp, _gp, _vp, _vp, _vp, _vp]),
    "mm_dev_signal_metrics": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_master": (_i, [_vp, _gp, _i, C.POINTER(Style), _vp, _vp, _vp, _vp, _u64, _vp, _u32]),
    "mm_dev_apply_target_curve_linear_phase": (_i, [_vp, _gp, _vp, _vp, _i]),
    "mm_design_linear_phase_ir": (_i, [_i, _i, _vp]),
    "mm_dev_apply_dynamic_eq": (_i, [_vp, _gp, _vp, _vp, _i, _vp]),
    "mm_dev_apply_dynamic_eq2": (_i, [_vp, _gp, _vp, _vp, _i, _vp, _u32, C.POINTER(C.c_int32)]),
    "mm_dev_fft_resample": (_i, [_vp, _gp, _vp, _gp, _vp]),
    "mm_dev_apply_spectral_denoise": (_i, [_vp, _gp, _vp, _vp, _d, _d]),
    "mm_dev_spectral_envelope": (_i, [_vp, _gp, _vp, _vp]),
    "mm_dev_fir_same": (_i, [_vp, _gp, _vp, _vp, _vp, _i, _i]),
    "mm_dev_apply_reverb": (_i, [_vp, _gp, _vp, _vp, _i, _d, _d, _i, _d, _d]),
    "mm_dev_apply_transient_designer": (_i, [_vp, _gp, _vp, _vp, _d, _d]),
    "mm_dev_apply_maximizer_transient_aware": (_i, [_vp, _gp, _vp, _vp, _d]),
    "mm_dev_apply_high_freq_trim": (_i, [_vp, _gp, _vp, _vp, _d, _d]),
    "mm_dev_apply_stereo_imager_4band": (_i, [_vp, _gp, _vp, _vp, _dp, _dp]),
    "mm_dev_apply_stereoize": (_i, [_vp, _gp, _vp, _vp, _d, _d, _d]),
    "mm_slice_margin": (_i64, [C.c_int32]),
    "mm_dev_master_slice": (_i, [_vp, _gp, _i, C.POINTER(Style), _vp, _vp, _vp, _vp, _u64, _vp, _u32, _vp]),
    "mm_nccl_unique_id": (_i, [_vp]),
    "mm_nccl_comm_create": (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    "mm_nccl_comm_destroy": (_i, [_vp]),
    "mm_nccl_version": (_i, []),
    "mm_master_host": (_i, [_vp, _i, C.c_int32, _i64, C.c_int32, C.c_int32, C.POINTER(Style), _vp, _vp, _vp, _vp, _u64,
                            C.POINTER(TrackStats), _u32]),
    "mm_master_host_pcm16": (_i, [_vp, _i, C.c_int32, _i64, C.c_int32, C.c_int32, C.POINTER(Style), _vp, _vp, _vp, _u64,
                                  C.POINTER(TrackStats), _u32]),
    "mm_master_host_ids": (_i, [_vp, _i, C.c_int32, _i64, C.c_int32, C.c_int32, C.POINTER(Style), _vp, _vp, _vp, _vp, _u64,
                                C.POINTER(TrackStats), _u32, C.POINTER(C.c_int32)]),
    "mm_master_host_jobs": (_i, [_vp, _i, C.c_int32, C.POINTER(HostJob), _u64, _u32]),
    "mm_host_alloc": (_i, [C.POINTER(_vp), _i64]),
    "mm_host_free": (_i, [_vp]),
    "mm_ctx_set_lanes": (_i, [_vp, _i]),
    "mm_ctx_copy_in": (_i, [_vp, _vp, _vp, _i64]),
    "mm_ctx_copy_out": (_i, [_vp, _vp, _vp, _i64]),
    "mm_ctx_workspace_bytes": (_i64, [_vp]),
    "mm_master_workspace_bytes": (_i64, [_gp, _i]),
    "mm_design_butter": (_i, [_i, _i, _dp, _dp, _dp]),
    "mm_design_iirpeak": (_i, [_d, _d, _dp, _dp, C.POINTER(_i), _dp]),
    "mm_design_lfilter_zi": (_i, [_dp, _dp, _i, _dp]),
    "mm_design_k_weighting": (_i, [_i, _d, _dp, _dp]),
    "mm_design_svf_highpass": (_i, [_dp, _dp, _dp, _dp]),
    "mm_design_scan_tables": (_i, [_dp, _dp, _i, _dp, _dp, _dp, _dp, _dp, _i, _dp, _dp, C.POINTER(_i), C.POINTER(_i)]),
    "mm_design_scan_tables2": (_i, [_dp, _dp, _i, _i, _dp, _dp, _dp, _dp, _dp, _i, _dp, _dp, _dp]),
}

_lib = None


def load():
    """Load ``libmm_b200.so`` (building nothing): raises ImportError with instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m mm_b200.build` (nvcc, sm_100a). "
            "mm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mm_abi_version() != 1:
        raise ImportError("libmm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return (load().mm_last_error() or b"").decode("utf-8", "replace")


def check(rc: int):
    if rc != 0:
        raise MMError(last_error())


def darr(values):
    """Python sequence -> ctypes double array (or None)."""
    if values is None:
        return None
    vals = [float(v) for v in values]
    return (C.c_double * len(vals))(*vals)
