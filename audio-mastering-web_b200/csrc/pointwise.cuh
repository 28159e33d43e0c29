// Pointwise stage bodies that fold into sweep epilogues or run in the finalise pass.
// Arithmetic widths follow the reference's numpy dtypes (NEP 50 promotion, numpy >= 2):
// float64 where the operand came out of scipy.filtfilt, float32 elsewhere.  Compiled with
// --fmad=false so every rounding below is the one written.
#pragma once
#include "common.cuh"

namespace mm {

// _compress_soft_knee on a float64 band (backend/app/pipeline.py:282-330) -> float32
__device__ __forceinline__ float compress_band_f64(double x, const DynBand& b, double* bypass_f64) {
    *bypass_f64 = x;
    if (b.mode == 0) return (float)x;
    const double ax = fabs(x);
    const double sgn = (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
    double o;
    if (b.mode == 3) {          // upward (ratio < 1), dB domain
        const double lvl = ax > 1e-12 ? 20.0 * log10(fmax(ax, 1e-12)) : -100.0;
        double boost = (b.thr_db - lvl) * (1.0 - b.ratio);
        boost = fmin(fmax(boost, 0.0), b.max_boost_db);
        o = fmin(fmax(ax * pow(10.0, boost / 20.0), 0.0), 1.0);
    } else if (b.mode == 1) {   // knee < 0.5 dB
        o = fmin(ax, b.thr + fmax(ax - b.thr, 0.0) / b.ratio);
    } else {                    // soft knee, piecewise linear
        o = ax <= b.lower ? ax : (ax >= b.upper ? b.thr + (ax - b.thr) / b.ratio : b.lower + (ax - b.lower) * b.slope);
        o = fmax(o, 0.0);
    }
    return (float)(sgn * o);
}

// band -> compress -> hard clip at lim_db -> * gain, all as the numpy branch does it
// (backend/app/pipeline.py:466-474).  The downward soft/hard knee (the default configuration) runs in
// float32: the band itself is float32 here (HBM storage), the reference's float64 knee arithmetic
// followed by its float32 cast differs from this by at most one float32 ulp of the band sample.
// The upward (ratio < 1) branch needs log10/pow and stays in float64.
__device__ __forceinline__ float band_chain(float y, const DynBand& b) {
    if (b.mode == 0) {          // ratio == 1: the float64 band is only clipped
        return __fmul_rn(fminf(fmaxf(y, -b.lim), b.lim), b.gain);
    }
    if (b.mode == 3) {
        double raw;
        const float c = compress_band_f64((double)y, b, &raw);
        return __fmul_rn(fminf(fmaxf(c, -b.lim), b.lim), b.gain);
    }
    const float ax = fabsf(y);
    float o;
    if (b.mode == 1) {
        o = fminf(ax, fmaf(fmaxf(ax - b.thr_f, 0.f), b.inv_ratio_f, b.thr_f));
    } else {
        const float hi = fmaf(ax - b.thr_f, b.inv_ratio_f, b.thr_f);
        const float mid = fmaf(ax - b.lower_f, b.slope_f, b.lower_f);
        o = ax <= b.lower_f ? ax : (ax >= b.upper_f ? hi : mid);
        o = fmaxf(o, 0.f);
    }
    o = fminf(o, b.lim);        // clip(+-lim) of sign * o
    return __fmul_rn(copysignf(o, y), b.gain);
}

// apply_maximizer + hard limiter at TRUE_PEAK_LIMIT_DB on float32 (pipeline.py:484-492, :636)
__device__ __forceinline__ float maximize_limit(float s, const DynParams& d) {
    const float ax = fabsf(s);
    const float sgn = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
    float o = ax;
    if (!(ax <= d.max_thr)) {
        o = __fadd_rn(d.max_thr, __fdiv_rn(__fmul_rn(__fsub_rn(ax, d.max_thr), d.max_num), d.max_den));
    }
    o = fminf(o, d.max_ceil);
    const float v = __fmul_rn(sgn, o);
    return fminf(fmaxf(v, -d.tp_lim), d.tp_lim);
}

// apply_parallel_compression on float32 (pipeline.py:1771-1797; soft knee 6 dB, float32 arithmetic)
__device__ __forceinline__ float parallel_compress(float x, double mixd, const DynParams& d) {
    const float mix = (float)mixd;
    const float one_minus = (float)(1.0 - mixd);          // python float (1.0 - mix), then weak-cast to float32
    const float ax = fabsf(x);
    const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
    float o;
    if (ax <= d.par_lower) o = ax;
    else if (ax >= d.par_upper) o = __fadd_rn(d.par_thr, __fdiv_rn(__fsub_rn(ax, d.par_thr), d.par_ratio));
    else o = __fadd_rn(d.par_lower, __fmul_rn(__fsub_rn(ax, d.par_lower), d.par_slope));
    o = fmaxf(o, 0.f);
    const float comp = __fmul_rn(sgn, o);
    const float dry = __fmul_rn(x, one_minus);
    const float out = __fadd_rn(dry, __fmul_rn(comp, mix));
    return fminf(fmaxf(out, -1.f), 1.f);
}

// _exciter_saturate "warm"/tape/tube/transistor/digital in float64 (pipeline.py:1179-1197)
__device__ __forceinline__ double exciter_sat(double x, int mode, double k) {
    x = fmin(fmax(x, -1.0), 1.0);
    switch (mode) {
        case 1: return tanh(k * x) / (k + 1e-8);                         // tape
        case 2: return x + 0.3 * (x * x);                                // tube
        case 3: return x - (x * x * x) / 3.0;                            // transistor
        case 4: return x;                                                // digital (|x| <= 1 after the clip)
        default: return 0.5 * (tanh(k * x) / (k + 1e-8) + x + 0.3 * (x * x));  // warm
    }
}

// ordered-int encoding so atomicMax/atomicMin work on floats of either sign
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace mm
