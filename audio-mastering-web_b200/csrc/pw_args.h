// Argument blocks of the reduction / pointwise kernels (misc_kernels.cuh), shared with capi.cu.
#pragma once
#include "common.cuh"

namespace mm {

constexpr int kPwThreads = 256;
constexpr int kPwFramesPerBlock = kPwThreads * 4 * 4;   // 4 float4 per thread

struct RowStats { double sum; unsigned mn, mx, pad; };   // mn/mx in ordered-uint encoding

struct InScalarArgs {
    const RowStats* st;
    long long n;
    int tracks, channels, use_dc, use_guard;
    float limit;            // float32(10 ** (-headroom_db / 20))
    double* sub;            // [rows]
    double* mul;            // [rows]
    double* peak_track;     // [tracks] or null
    double* mean_row;       // [rows] or null
};

struct OutScalarArgs {
    const float* peak_bits;   // [tracks]
    int tracks, channels;
    float limit;
    double* mul;              // [rows]
    double* peak_track;       // [tracks] or null
};

enum { PW_AFFINE = 0, PW_MAXIMIZER = 1, PW_PARALLEL = 2, PW_IMAGER = 3, PW_FINALIZE = 4, PW_PEAK = 5, PW_FADE = 6,
       PW_MS_ENCODE = 7, PW_MS_DECODE = 8, PW_GAIN_F64 = 9, PW_BLEND = 10 };

struct PwArgs {
    const float* in;
    const float* in2;           // PW_BLEND: the processed signal
    float blend;                // PW_BLEND: amount in [0, 1]
    float* out;                 // planar (may be null for PW_PEAK)
    long long n, stride;
    int tracks, channels, mode;
    const double* sub;          // per row (PW_AFFINE / FINALIZE: unused) or null
    const double* mul;          // per row scale or null
    int clip;                   // clip to +-1 after the affine map
    const double* width;        // per track imager width (null = none)
    int force_imager;           // run the mid/side arithmetic even for width == 1 (standalone apply_stereo_imager)
    int skip_unity;             // PW_PEAK over a mixed batch: CTAs of tracks whose width is 1 return at once
    DynParams dyn;              // maximizer / parallel constants
    double par_mix;             // PW_PARALLEL uniform mix
    int n_fade;                 // fade-in length (0 = none)
    double fade_step;           // 1 / (n_fade - 1)
    float* peak;                // per track float-bits max (PW_PEAK / FINALIZE input side)
    // int16 export fused into FINALIZE
    int16_t* pcm;               // interleaved [tracks][n][ch] or null
    const float* noise;         // interleaved float noise or null (-> Philox)
    unsigned long long seed;
    double* nonfinite;          // per track count of non-finite samples seen before nan_to_num (or null)
    int track_base;             // mm_geom::track_base (dither counter)
};

struct QuantArgs {
    const float* in; long long n, stride; int tracks, channels;
    int16_t* pcm; const float* noise; unsigned long long seed;
    int track_base;
    const float* noise_planar;   // shaped dither: planar rows (batch layout), scaled by noise_scale in float32
    float noise_scale;
};
struct WhiteArgs {               // white noise in [-1, 1) for the noise-shaped dithers, planar rows
    const float* uniform;        // interleaved float32 uniforms in [0, 1) (np.random.rand(...).astype(float32)) or null -> Philox
    float* out; long long n, stride; int tracks, channels; unsigned long long seed; int track_base;
};

}  // namespace mm
