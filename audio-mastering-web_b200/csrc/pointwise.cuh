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
// (backend/app/pipeline.py:466-474).  The downward knee (ratio >= 1, the default configuration) is one
// branch-free piecewise-linear map in float32,
//     o = |y| <= lower ? |y| : (|y| >= upper ? thr + (|y| - thr)/ratio : lower + (|y| - lower) * slope)
// = min(|y|, knee line, above-knee line) (the map is concave); a bypassed band (ratio == 1) uses the
// identity for all three lines, a hard knee the above-knee line twice.  The band
// itself is float32 here (HBM storage); the reference's float64 knee followed by its float32 cast differs
// from this by at most one float32 ulp of the band sample.  The upward branch (ratio < 1) needs
// log10/pow, stays in float64 and out of line so that the hot loop stays small in the instruction cache.
static __device__ __noinline__ float band_chain_upward(float y, const DynBand& b) {
    double raw;
    const float c = compress_band_f64((double)y, b, &raw);
    return __fmul_rn(fminf(fmaxf(c, -b.lim), b.lim), b.gain);
}
// downward knee only (modes 0..2): branch free.  A downward knee is concave piecewise linear, i.e. the
// minimum of its three lines (identity, knee segment, above-knee segment).
__device__ __forceinline__ float band_chain(float y, const DynBand& b) {
    const float ax = fabsf(y);
    float o = fminf(ax, fminf(fmaf(ax, b.s_mid, b.c_mid), fmaf(ax, b.s_hi, b.c_hi)));
    o = fminf(fmaxf(o, 0.f), b.lim);        // clip(+-lim) of sign * o
    return __fmul_rn(copysignf(o, y), b.gain);
}
// any mode, including the upward branch (ratio < 1)
__device__ __forceinline__ float band_chain_gen(float y, const DynBand& b) {
    if (b.mode == 3) return band_chain_upward(y, b);
    return band_chain(y, b);
}

// apply_maximizer + hard limiter at TRUE_PEAK_LIMIT_DB on float32 (pipeline.py:484-492, :636):
//   |s| <= thr ? |s| : thr + (|s| - thr) * (ceil - thr) / (1 - thr), then min(ceil), sign, clip(+-tp_lim)
// = copysign(min(|s|, line(|s|), ceil, tp_lim), s): concave again (the line has slope < 1).
__device__ __forceinline__ float maximize_limit(float s, const DynParams& d) {
    const float ax = fabsf(s);
    const float o = fminf(fminf(ax, fmaf(ax, d.max_k, d.max_c)), d.max_top);
    return copysignf(o, s);
}

// apply_parallel_compression on float32 (pipeline.py:1771-1797; soft knee 6 dB): x (1 - mix) + comp mix, clip +-1
__device__ __forceinline__ float parallel_compress(float x, float mix, float one_minus, const DynParams& d) {
    const float ax = fabsf(x);
    float o = fminf(ax, fminf(fmaf(ax, d.par_slope, d.par_cmid), fmaf(ax, d.par_shi, d.par_chi)));
    o = fmaxf(o, 0.f);
    const float comp = copysignf(o, x);
    const float out = __fadd_rn(__fmul_rn(x, one_minus), __fmul_rn(comp, mix));
    return fminf(fmaxf(out, -1.f), 1.f);
}

// _exciter_saturate "warm"/tape/tube/transistor/digital (pipeline.py:1179-1197).  The reference evaluates
// it in float64; here the saturator runs in float32 (tanhf, full-precision variant): what reaches the
// output is (sat - hf) * gain * 0.25 with gain = 10^(dB/20) - 1 <= ~0.1, so a float32 ulp of `sat`
// (6e-8 |hf|) moves the output by < 2e-9.
__device__ __forceinline__ float exciter_sat(float x, int mode, float k) {
    x = fminf(fmaxf(x, -1.f), 1.f);
    switch (mode) {
        case 1: return tanhf(k * x) / (k + 1e-8f);                          // tape
        case 2: return x + 0.3f * (x * x);                                  // tube
        case 3: return x - (x * x * x) / 3.0f;                              // transistor
        case 4: return x;                                                   // digital (|x| <= 1 after the clip)
        default: return 0.5f * (tanhf(k * x) / (k + 1e-8f) + x + 0.3f * (x * x));  // warm
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
