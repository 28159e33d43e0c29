// Reductions and pointwise passes of the chain: input statistics, peak guards, imager,
// finalise (+ TPDF dither to int16), layout conversion.
#pragma once
#include "pointwise.cuh"
#include "pw_args.h"

namespace mm {


// ---- per-row sum / min / max (remove_dc_offset + remove_intersample_peaks, pipeline.py:134-149) ----

// A CTA reduces kRsFramesPerBlock samples of one row: eight 16-byte loads per thread in flight at a time (round 1: four loads, one
// barrier and three atomics per 4096 samples -- 248 000 CTAs per batch, 0.66 of the copy peak with `long_scoreboard` 31 cycles per
// issued instruction).
constexpr int kRsFramesPerBlock = kPwFramesPerBlock * 8;
__global__ void __launch_bounds__(kPwThreads) row_stats_kernel(const float* __restrict__ in, long long n, long long stride,
                                                               RowStats* __restrict__ st) {
    const int row = blockIdx.y;
    const float* src = in + (size_t)row * (size_t)stride + kLead;
    const long long base = (long long)blockIdx.x * kRsFramesPerBlock;
    double s = 0.0;
    float mn = __int_as_float(0x7f800000), mx = -__int_as_float(0x7f800000);
    if (base + kRsFramesPerBlock <= n) {
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = __ldcs(reinterpret_cast<const float4*>(src + base + 4LL * (threadIdx.x + kPwThreads * (8 * it + r))));
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                s += ((double)v[r].x + (double)v[r].y) + ((double)v[r].z + (double)v[r].w);
                mn = fminf(fminf(mn, v[r].x), fminf(v[r].y, fminf(v[r].z, v[r].w)));
                mx = fmaxf(fmaxf(mx, v[r].x), fmaxf(v[r].y, fmaxf(v[r].z, v[r].w)));
            }
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < 32; ++r) {
            const long long i = base + 4LL * (threadIdx.x + kPwThreads * r);
            if (i + 3 < n) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src + i));
                s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
                mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
                mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
            } else {
                for (int c = 0; c < 4; ++c)
                    if (i + c < n) { const float x = src[i + c]; s += (double)x; mn = fminf(mn, x); mx = fmaxf(mx, x); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += shfl_xor_d(s, o);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ double ss[kPwThreads / 32];
    __shared__ float smn[kPwThreads / 32], smx[kPwThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ss[warp] = s; smn[warp] = mn; smx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kPwThreads / 32; ++w) { s += ss[w]; mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
        atomicAdd(&st[row].sum, s);
        atomicMin(&st[row].mn, f2ord(mn));
        atomicMax(&st[row].mx, f2ord(mx));
    }
}

// RowStats <-> three float64 vectors [sum | min | max] for the cross-rank reduction of a time-sliced file
__global__ void row_stats_pack_kernel(const RowStats* st, int rows, double* buf) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    // [sums | negated minima | maxima]: one SUM and one MAX all-reduce (min x = -max(-x))
    if (r < rows) { buf[r] = st[r].sum; buf[rows + r] = -(double)ord2f(st[r].mn); buf[2 * rows + r] = (double)ord2f(st[r].mx); }
}
__global__ void row_stats_unpack_kernel(RowStats* st, int rows, const double* buf) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) { st[r].sum = buf[r]; st[r].mn = f2ord((float)-buf[rows + r]); st[r].mx = f2ord((float)buf[2 * rows + r]); }
}

// the tracks with an active imager re-track their output peak after the imager (peak_after_imager): zero those entries first
__global__ void reset_imager_peaks_kernel(float* peak, const double* width, int tracks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < tracks && width[t] != 1.0) peak[t] = 0.f;
}

__global__ void row_stats_init_kernel(RowStats* st, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) { st[r].sum = 0.0; st[r].mn = 0xffffffffu; st[r].mx = 0u; st[r].pad = 0; }
}

// mean / peak -> per-row (sub, mul) of the float32 prologue  y = (x - mean) * scale
//   use_dc: subtract the channel mean (remove_dc_offset); use_guard: scale to -headroom if the
//   track peak exceeds it (remove_intersample_peaks).  Float32 steps as numpy takes them.
__global__ void in_scalars_kernel(const InScalarArgs P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.tracks) return;
    float peak = 0.f;
    bool nan = false;
    for (int c = 0; c < P.channels; ++c) {
        const int row = t * P.channels + c;
        const float mean = P.use_dc ? (float)(P.st[row].sum / (double)P.n) : 0.f;
        const float hi = __fsub_rn(ord2f(P.st[row].mx), mean);
        const float lo = __fsub_rn(ord2f(P.st[row].mn), mean);
        peak = fmaxf(peak, fmaxf(fabsf(hi), fabsf(lo)));
        nan = nan || !(hi == hi) || !(lo == lo);
        P.sub[row] = (double)mean;
        if (P.mean_row) P.mean_row[row] = (double)mean;
    }
    float scale = 1.f;
    if (P.use_guard && !nan && peak <= 3.0e38f && peak > 1e-12f && peak > P.limit) scale = __fdiv_rn(P.limit, peak);
    for (int c = 0; c < P.channels; ++c) P.mul[t * P.channels + c] = (double)scale;
    if (P.peak_track) P.peak_track[t] = (double)peak;
}

// Output guard scalars from an already accumulated per-track |x| max (float bits).
__global__ void out_scalars_kernel(const OutScalarArgs P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.tracks) return;
    const float peak = P.peak_bits[t];
    float scale = 1.f;
    if (peak == peak && peak <= 3.0e38f && peak > 1e-12f && peak > P.limit) scale = __fdiv_rn(P.limit, peak);
    for (int c = 0; c < P.channels; ++c) P.mul[t * P.channels + c] = (double)scale;
    if (P.peak_track) P.peak_track[t] = (double)peak;
}

// ---- generic pointwise pass over (track, frames); both channels of a track in one thread -------------


__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// _write_wav_16bit_dithered quantiser (pipeline.py:887-898): float64(x) * 32767 + noise, round half to
// even, clip.  x * 32767 needs 39 significant bits, hence float64.  The rounding is the 1.5 * 2^52 magic
// add on the FP64 pipe (round-to-nearest-even, exactly numpy's np.round for |d| < 2^31) instead of
// rint + a double->int conversion: conversions run at 15 lanes/clk/SM on this part.
__device__ __forceinline__ int16_t quantize16(float x, double noise) {
    if (x != x) x = 0.f;
    x = fminf(fmaxf(x, -1.f), 1.f);
    const double d = __dadd_rn(__dmul_rn((double)x, 32767.0), noise);
    const int q = __double2loint(__dadd_rn(d, 6755399441055744.0));
    return (int16_t)min(max(q, -32768), 32767);
}

// The same quantiser for a sample the caller has already cleaned (finite, clipped to +-1): x * 32767 is exact in float64 (24 + 15
// bits), so ONE fused multiply-add rounds exactly where the product followed by the add does; with |noise| < 1 (TPDF) the sum
// cannot round below -32768 and only the upper clamp is needed.
template <bool TPDF> __device__ __forceinline__ int16_t quantize16_clean(float x, double noise) {
    const double d = fma((double)x, 32767.0, noise);
    const int q = __double2loint(__dadd_rn(d, 6755399441055744.0));
    return (int16_t)(TPDF ? min(q, 32767) : min(max(q, -32768), 32767));
}

// TPDF dither noise (rand + rand - 1.0, pipeline.py:830-832) from counter-based random bits.
// One Philox4x32-10 call serves TWO frames: each 32-bit word gives one TPDF sample from its two 16-bit halves,
//   n = (hi16 + lo16) / 65536 - 1  in (-1, 1)   (triangular on a 2^-16 LSB lattice).
// Counter = (frame >> 1, track); word (frame & 1) * 2 + channel.  The 10 rounds are ~100 integer instructions, which
// made the finalise pass ALU bound at one call per frame.
__device__ __forceinline__ double tpdf16(unsigned w) {
    const unsigned s = (w >> 16) + (w & 0xffffu);                       // 0 .. 131070
    const double d = __hiloint2double(0x43300000, (int)s) - 4503599627370496.0;   // exactly s (2^52 bit placement)
    return fma(d, 1.0 / 65536.0, -1.0);
}
__device__ __forceinline__ void dither_words(unsigned long long frame, int track, unsigned long long seed, unsigned (&rnd)[4]) {
    const unsigned long long pair = frame >> 1;
    philox4x32_10((unsigned)pair, (unsigned)(pair >> 32), (unsigned)track, 0u, (unsigned)seed, (unsigned)(seed >> 32), rnd);
}

__global__ void __launch_bounds__(kPwThreads) pointwise_kernel(const PwArgs P) {
    const int track = blockIdx.y;
    const int C = P.channels;
    const long long base = (long long)blockIdx.x * kPwFramesPerBlock;
    const size_t r0 = (size_t)(track * C) * (size_t)P.stride + kLead;
    const size_t r1 = r0 + (C > 1 ? (size_t)P.stride : 0);
    float sub0 = 0.f, sub1 = 0.f, mul0 = 1.f, mul1 = 1.f;
    if (P.sub) { sub0 = (float)P.sub[track * C]; sub1 = (float)P.sub[track * C + (C > 1)]; }
    if (P.mul) { mul0 = (float)P.mul[track * C]; mul1 = (float)P.mul[track * C + (C > 1)]; }
    const bool imager = (P.width != nullptr) && C == 2 && (P.force_imager || fabs(P.width[track] - 1.0) > 0.0);
    if (P.mode == PW_PEAK && P.skip_unity && !imager) return;      // mixed batch: only the tracks with an active imager are re-scanned
    const double muld0 = (P.mode == PW_GAIN_F64 && P.mul) ? P.mul[track * C] : 1.0;
    const double muld1 = (P.mode == PW_GAIN_F64 && P.mul) ? P.mul[track * C + (C > 1)] : 1.0;
    const float wf = imager ? (float)P.width[track] : 1.f;
    float pk = 0.f;
    double bad = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const long long i = base + 4LL * (threadIdx.x + kPwThreads * r);
        if (i >= P.n) continue;
        const bool full = i + 3 < P.n;
        float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        if (full) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(P.in + r0 + i));
            a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
            if (C > 1) { const float4 w = __ldcs(reinterpret_cast<const float4*>(P.in + r1 + i)); b[0] = w.x; b[1] = w.y; b[2] = w.z; b[3] = w.w; }
        } else {
            for (int c = 0; c < 4; ++c) if (i + c < P.n) { a[c] = P.in[r0 + i + c]; if (C > 1) b[c] = P.in[r1 + i + c]; }
        }
        int16_t q[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float l = a[c], rr = b[c];
            if (P.mode == PW_AFFINE) {
                l = __fmul_rn(__fsub_rn(l, sub0), mul0); rr = __fmul_rn(__fsub_rn(rr, sub1), mul1);
                if (P.clip) { l = fminf(fmaxf(l, -1.f), 1.f); rr = fminf(fmaxf(rr, -1.f), 1.f); }
            } else if (P.mode == PW_MAXIMIZER) {
                // apply_maximizer alone (no limiter): tp_lim is set to +inf by the host
                l = maximize_limit(l, P.dyn); rr = maximize_limit(rr, P.dyn);
            } else if (P.mode == PW_PARALLEL) {
                l = parallel_compress(l, (float)P.par_mix, (float)(1.0 - P.par_mix), P.dyn); rr = parallel_compress(rr, (float)P.par_mix, (float)(1.0 - P.par_mix), P.dyn);
            } else if (P.mode == PW_BLEND) {
                // BaseModule.process (modules/base.py:44-46): audio * (1 - amount) + processed * amount, float32
                const size_t o = (size_t)i + c;
                const float pl = (i + c < P.n) ? P.in2[r0 + o] : 0.f, pr = (C > 1 && i + c < P.n) ? P.in2[r1 + o] : 0.f;
                const float one_m = (float)(1.0 - (double)P.blend);
                l = __fadd_rn(__fmul_rn(l, one_m), __fmul_rn(pl, P.blend));
                rr = __fadd_rn(__fmul_rn(rr, one_m), __fmul_rn(pr, P.blend));
            } else if (P.mode == PW_GAIN_F64) {
                // normalize_lufs (pipeline.py:654-655): float32 array * float64 scalar -> float64 -> float32
                l = (float)((double)l * muld0); rr = (float)((double)rr * muld1);
            } else if (P.mode == PW_MS_ENCODE) {
                // pipeline.py:249-250: mid = (L + R) * 0.5, side = (L - R) * 0.5 (float32)
                const float m = __fmul_rn(__fadd_rn(l, rr), 0.5f), sd = __fmul_rn(__fsub_rn(l, rr), 0.5f);
                l = m; rr = sd;
            } else if (P.mode == PW_MS_DECODE) {
                // pipeline.py:253-254: L = clip(m + s), R = clip(m - s)
                const float lo = fminf(fmaxf(__fadd_rn(l, rr), -1.f), 1.f), ro = fminf(fmaxf(__fsub_rn(l, rr), -1.f), 1.f);
                l = lo; rr = ro;
            } else if (P.mode == PW_FADE) {
                if (i + c < P.n_fade) {
                    const float ramp = (i + c == P.n_fade - 1) ? 1.f : (float)((double)(i + c) * P.fade_step);
                    l = __fmul_rn(l, ramp); rr = __fmul_rn(rr, ramp);
                }
            } else {   // PW_IMAGER, PW_PEAK, PW_FINALIZE share the imager front end
                if (imager) {
                    const float mid = __fmul_rn(__fadd_rn(l, rr), 0.5f);
                    const float side = __fmul_rn(__fmul_rn(__fsub_rn(l, rr), 0.5f), wf);
                    l = fminf(fmaxf(__fadd_rn(mid, side), -1.f), 1.f);
                    rr = fminf(fmaxf(__fsub_rn(mid, side), -1.f), 1.f);
                }
                if (P.mode == PW_PEAK) {
                    if (i + c < P.n) { pk = fmaxf(pk, fabsf(l)); if (C > 1) pk = fmaxf(pk, fabsf(rr)); }
                } else if (P.mode == PW_FINALIZE) {
                    if (i + c < P.n) { if (!(fabsf(l) <= 3.4e38f)) bad += 1.0; if (C > 1 && !(fabsf(rr) <= 3.4e38f)) bad += 1.0; }
                    l = __fmul_rn(l, mul0); rr = __fmul_rn(rr, mul1);
                    l = (l != l) ? 0.f : fminf(fmaxf(l, -1.f), 1.f);
                    rr = (rr != rr) ? 0.f : fminf(fmaxf(rr, -1.f), 1.f);
                    if (i + c < P.n_fade) {
                        const float ramp = (i + c == P.n_fade - 1) ? 1.f : (float)((double)(i + c) * P.fade_step);
                        l = __fmul_rn(l, ramp); rr = __fmul_rn(rr, ramp);
                    }
                    if (P.pcm) {
                        double n0, n1 = 0.0;
                        const size_t fi = (size_t)track * (size_t)P.n + (size_t)(i + c);
                        if (P.noise) {
                            if (i + c < P.n) { n0 = (double)P.noise[fi * C]; if (C > 1) n1 = (double)P.noise[fi * C + 1]; } else n0 = 0.0;
                        } else {
                            unsigned rnd[4];
                            const unsigned long long fr = (unsigned long long)(i + c);
                            dither_words(fr, track + P.track_base, P.seed, rnd);
                            n0 = tpdf16(rnd[(fr & 1) * 2]); n1 = tpdf16(rnd[(fr & 1) * 2 + 1]);
                        }
                        q[c * C] = quantize16(l, n0);
                        if (C > 1) q[c * C + 1] = quantize16(rr, n1);
                    }
                }
            }
            a[c] = l; b[c] = rr;
        }
        if (P.out) {
            if (full) {
                __stcs(reinterpret_cast<float4*>(P.out + r0 + i), make_float4(a[0], a[1], a[2], a[3]));
                if (C > 1) __stcs(reinterpret_cast<float4*>(P.out + r1 + i), make_float4(b[0], b[1], b[2], b[3]));
            } else {
                for (int c = 0; c < 4; ++c) if (i + c < P.n) { P.out[r0 + i + c] = a[c]; if (C > 1) P.out[r1 + i + c] = b[c]; }
            }
        }
        if (P.mode == PW_FINALIZE && P.pcm) {
            int16_t* dst = P.pcm + ((size_t)track * (size_t)P.n + (size_t)i) * C;
            if (full && C == 2) {
                // 4 frames x 2 channels = 16 bytes; base offset is 4-frame aligned but the track origin
                // need not be 16-byte aligned (odd n), so fall back to 4-byte stores when it is not
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    int4 pk4;
                    pk4.x = (int)(uint16_t)q[0] | ((int)(uint16_t)q[1] << 16);
                    pk4.y = (int)(uint16_t)q[2] | ((int)(uint16_t)q[3] << 16);
                    pk4.z = (int)(uint16_t)q[4] | ((int)(uint16_t)q[5] << 16);
                    pk4.w = (int)(uint16_t)q[6] | ((int)(uint16_t)q[7] << 16);
                    *reinterpret_cast<int4*>(dst) = pk4;
                } else {
                    for (int c = 0; c < 8; ++c) dst[c] = q[c];
                }
            } else {
                for (int c = 0; c < 4; ++c) if (i + c < P.n) for (int k = 0; k < C; ++k) dst[c * C + k] = q[c * C + k];
            }
        }
    }
    if (P.mode == PW_PEAK && P.peak) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        if ((threadIdx.x & 31) == 0 && pk > 0.f) atomicMax(reinterpret_cast<int*>(P.peak + track), __float_as_int(pk));
    }
    if (P.mode == PW_FINALIZE && P.nonfinite) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bad += shfl_xor_d(bad, o);
        if ((threadIdx.x & 31) == 0 && bad > 0.0) atomicAdd(P.nonfinite + track, bad);
    }
}

// ---- final pass of a chain: imager -> output peak guard scale -> clip / nan_to_num -> fade-in -> float32 out
//      (+ TPDF dither to interleaved int16).  remove_intersample_peaks (pipeline.py:141-149), the chain's final
//      clip (:1906-1908 / chain.py:93-94), apply_output_edge_fade_in (:152-167), _write_wav_16bit_dithered (:880-898).
struct FinalArgs {
    const float* in;
    float* out;
    long long n, stride;
    int tracks;
    const double* mul;          // per row output-guard scale
    const double* width;        // per track imager width or null
    int n_fade;
    double fade_step;
    int16_t* pcm;               // interleaved [tracks][n][C] or null
    const float* noise;         // interleaved float32 noise or null (-> Philox)
    unsigned long long seed;
    double* nonfinite;          // per track or null
    int track_base;             // mm_geom::track_base: keeps the dither stream independent of host-side chunking
    long long frame_base;       // index of frame 0 in the whole file (time slices): dither counter
    const int* track_ids;       // optional: the dither stream of track t is keyed by track_ids[t] (uploads of different shapes are
                                // mastered in groups; every track keeps the stream of its index in the caller's list)
};

constexpr int kFinThreads = 256;
constexpr int kFinVec = 4;                                   // float4 per thread per channel
constexpr int kFinFrames = kFinThreads * kFinVec * 4;        // frames per block

template <int C, bool PCM, bool NOISE>
__global__ void __launch_bounds__(kFinThreads, 3) finalize_kernel(const FinalArgs P) {
    const int track = blockIdx.y;
    const long long base = (long long)blockIdx.x * kFinFrames;
    const size_t r0 = (size_t)(track * C) * (size_t)P.stride + kLead;
    const size_t r1 = r0 + (C > 1 ? (size_t)P.stride : 0);
    const float mul0 = (float)__ldg(P.mul + track * C), mul1 = (float)__ldg(P.mul + track * C + (C > 1));
    const bool imager = C == 2 && P.width != nullptr && __ldg(P.width + track) != 1.0;
    const float wf = imager ? (float)__ldg(P.width + track) : 1.f;
    float4 a[kFinVec], b[kFinVec];
    // all loads first: 2 * kFinVec independent 16-byte requests per thread in flight
#pragma unroll
    for (int r = 0; r < kFinVec; ++r) {
        const long long i = base + 4LL * (threadIdx.x + kFinThreads * r);
        a[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        b[r] = a[r];
        if (i + 3 < P.n) {
            a[r] = __ldcs(reinterpret_cast<const float4*>(P.in + r0 + i));
            if (C > 1) b[r] = __ldcs(reinterpret_cast<const float4*>(P.in + r1 + i));
        } else {
            for (int c = 0; c < 4; ++c)
                if (i + c < P.n) { setcomp4(a[r], c, P.in[r0 + i + c]); if (C > 1) setcomp4(b[r], c, P.in[r1 + i + c]); }
        }
    }
    int bad = 0;                                             // non-finite samples seen by this thread (at most 32)
#pragma unroll
    for (int r = 0; r < kFinVec; ++r) {
        const long long i = base + 4LL * (threadIdx.x + kFinThreads * r);
        if (i >= P.n) continue;
        const bool full = i + 3 < P.n;
        const bool fading = i < P.n_fade;                    // uniform per vector: the 6 ms ramp touches the first CTA of a track only
        int16_t q[8];
        unsigned rnd[4] = {0u, 0u, 0u, 0u};
        float4 nz0 = make_float4(0.f, 0.f, 0.f, 0.f), nz1 = nz0;
        if (PCM && NOISE) {
            const float* np_ = P.noise + ((size_t)track * (size_t)P.n + (size_t)i) * C;
            if (full && ((reinterpret_cast<uintptr_t>(np_) & 15) == 0)) {
                nz0 = __ldcs(reinterpret_cast<const float4*>(np_));
                if (C > 1) nz1 = __ldcs(reinterpret_cast<const float4*>(np_ + 4));
            } else {
                for (int k = 0; k < 4 * C; ++k)
                    if (i + k / C < P.n) { if (k < 4) setcomp4(nz0, k, np_[k]); else setcomp4(nz1, k - 4, np_[k]); }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float l = comp4(a[r], c), rr = comp4(b[r], c);
            if (imager) {
                const float mid = __fmul_rn(__fadd_rn(l, rr), 0.5f);
                const float side = __fmul_rn(__fmul_rn(__fsub_rn(l, rr), 0.5f), wf);
                l = fminf(fmaxf(__fadd_rn(mid, side), -1.f), 1.f);
                rr = fminf(fmaxf(__fsub_rn(mid, side), -1.f), 1.f);
            }
            if (P.nonfinite && i + c < P.n) {
                if (!(fabsf(l) <= 3.4e38f)) ++bad;
                if (C > 1 && !(fabsf(rr) <= 3.4e38f)) ++bad;
            }
            l = __fmul_rn(l, mul0); rr = __fmul_rn(rr, mul1);
            l = (l != l) ? 0.f : fminf(fmaxf(l, -1.f), 1.f);
            rr = (rr != rr) ? 0.f : fminf(fmaxf(rr, -1.f), 1.f);
            if (fading && i + c < P.n_fade) {
                const float ramp = (i + c == P.n_fade - 1) ? 1.f : (float)((double)(i + c) * P.fade_step);
                l = __fmul_rn(l, ramp); rr = __fmul_rn(rr, ramp);
            }
            setcomp4(a[r], c, l);
            setcomp4(b[r], c, rr);
            if (PCM) {
                double n0, n1 = 0.0;
                if (NOISE) {
                    // interleaved: frame c of this vector holds elements c*C .. c*C + C - 1
                    const int e0 = c * C, e1 = c * C + 1;
                    n0 = (double)(e0 < 4 ? comp4(nz0, e0) : comp4(nz1, e0 - 4));
                    if (C > 1) n1 = (double)(e1 < 4 ? comp4(nz0, e1) : comp4(nz1, e1 - 4));
                } else {
                    // i and frame_base are even: frames (c, c + 1) of an even c share one Philox call
                    const unsigned long long fr = (unsigned long long)(i + c + P.frame_base);
                    if ((c & 1) == 0) dither_words(fr, P.track_ids ? __ldg(P.track_ids + track) : track + P.track_base, P.seed, rnd);
                    n0 = tpdf16(rnd[(c & 1) * 2]);
                    n1 = tpdf16(rnd[(c & 1) * 2 + 1]);
                }
                q[c * C] = quantize16_clean<!NOISE>(l, n0);           // l, rr are finite and within +-1 here
                if (C > 1) q[c * C + 1] = quantize16_clean<!NOISE>(rr, n1);
            }
        }
        if (full) {
            __stcs(reinterpret_cast<float4*>(P.out + r0 + i), a[r]);
            if (C > 1) __stcs(reinterpret_cast<float4*>(P.out + r1 + i), b[r]);
        } else {
            for (int c = 0; c < 4; ++c)
                if (i + c < P.n) { P.out[r0 + i + c] = comp4(a[r], c); if (C > 1) P.out[r1 + i + c] = comp4(b[r], c); }
        }
        if (PCM) {
            int16_t* dst = P.pcm + ((size_t)track * (size_t)P.n + (size_t)i) * C;
            if (full && C == 2 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                int4 pk4;
                pk4.x = (int)(uint16_t)q[0] | ((int)(uint16_t)q[1] << 16);
                pk4.y = (int)(uint16_t)q[2] | ((int)(uint16_t)q[3] << 16);
                pk4.z = (int)(uint16_t)q[4] | ((int)(uint16_t)q[5] << 16);
                pk4.w = (int)(uint16_t)q[6] | ((int)(uint16_t)q[7] << 16);
                __stcs(reinterpret_cast<int4*>(dst), pk4);
            } else {
                for (int c = 0; c < 4; ++c)
                    if (i + c < P.n) for (int k = 0; k < C; ++k) dst[c * C + k] = q[c * C + k];
            }
        }
    }
    if (P.nonfinite) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        if ((threadIdx.x & 31) == 0 && bad > 0) atomicAdd(P.nonfinite + track, (double)bad);
    }
}

// standalone quantiser for an already final float32 buffer (export_audio on its own)
__global__ void __launch_bounds__(kPwThreads) quantize_kernel(const QuantArgs P) {
    const int track = blockIdx.y, C = P.channels;
    const long long i = ((long long)blockIdx.x * kPwThreads + threadIdx.x);
    if (i >= P.n) return;
    const size_t fi = (size_t)track * (size_t)P.n + (size_t)i;
    unsigned rnd[4] = {0, 0, 0, 0};
    if (!P.noise && !P.noise_planar) dither_words((unsigned long long)i, track + P.track_base, P.seed, rnd);
    for (int c = 0; c < C; ++c) {
        const float x = P.in[(size_t)(track * C + c) * (size_t)P.stride + kLead + i];
        double nz;
        if (P.noise_planar) nz = (double)__fmul_rn(P.noise_planar[(size_t)(track * C + c) * (size_t)P.stride + kLead + i], P.noise_scale);
        else nz = P.noise ? (double)P.noise[fi * C + c] : tpdf16(rnd[(i & 1) * 2 + c]);
        P.pcm[fi * C + c] = quantize16(x, nz);
    }
}

// white = 2 * rand - 1 in float32 (pipeline.py:843 / :864), planar, for the shaping filter
__global__ void __launch_bounds__(kPwThreads) white_noise_kernel(const WhiteArgs P) {
    const int track = blockIdx.y, C = P.channels;
    const long long i = ((long long)blockIdx.x * kPwThreads + threadIdx.x);
    if (i >= P.n) return;
    unsigned rnd[4] = {0, 0, 0, 0};
    if (!P.uniform) philox4x32_10((unsigned)i, (unsigned)((unsigned long long)i >> 32), (unsigned)(track + P.track_base), 1u,
                                  (unsigned)P.seed, (unsigned)(P.seed >> 32), rnd);
    for (int c = 0; c < C; ++c) {
        const float u = P.uniform ? P.uniform[((size_t)track * (size_t)P.n + (size_t)i) * C + c] : (float)(rnd[c] >> 8) * 5.9604644775390625e-08f;
        P.out[(size_t)(track * C + c) * (size_t)P.stride + kLead + i] = __fsub_rn(__fmul_rn(2.0f, u), 1.0f);
    }
}

// mastering_trace.signal_metrics (backend/app/mastering_trace.py:115-149): per track, max |x| over the finite samples,
// count of non-finite samples, count of infinities.  out[track * 3 + {0, 1, 2}] (doubles, zeroed by the host).
__global__ void __launch_bounds__(kPwThreads) signal_metrics_kernel(const float* __restrict__ in, long long n, long long stride,
                                                                    int channels, double* __restrict__ out) {
    const int row = blockIdx.y;
    const float* src = in + (size_t)row * (size_t)stride + kLead;
    float pk = 0.f;
    double bad = 0.0, inf = 0.0;
    for (long long i = (long long)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kPwThreads) {
        const float v = src[i];
        const float a = fabsf(v);
        if (a <= 3.4028235e38f) pk = fmaxf(pk, a);
        else { bad += 1.0; if (a == a) inf += 1.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        bad += shfl_xor_d(bad, o);
        inf += shfl_xor_d(inf, o);
    }
    if ((threadIdx.x & 31) == 0) {
        double* o = out + (size_t)(row / channels) * 3;
        // |x| max as double bits is monotone for non-negative values: 64-bit atomicMax on the bit pattern
        atomicMax(reinterpret_cast<unsigned long long*>(o), (unsigned long long)__double_as_longlong((double)pk));
        if (bad > 0.0) atomicAdd(o + 1, bad);
        if (inf > 0.0) atomicAdd(o + 2, inf);
    }
}

// ---- layout conversion ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPwThreads) deinterleave_kernel(const float* __restrict__ il, float* __restrict__ pl,
                                                                  long long n, long long stride, int channels) {
    const int track = blockIdx.y;
    const long long i = (long long)blockIdx.x * kPwThreads + threadIdx.x;
    if (i >= n) return;
    const float* src = il + ((size_t)track * (size_t)n + (size_t)i) * channels;
    for (int c = 0; c < channels; ++c) pl[(size_t)(track * channels + c) * (size_t)stride + kLead + i] = src[c];
}
// PCM_16 frames -> planar float32, x / 32768 (libsndfile's float conversion of 16-bit PCM)
__global__ void __launch_bounds__(kPwThreads) deinterleave_pcm16_kernel(const int16_t* __restrict__ il, float* __restrict__ pl,
                                                                        long long n, long long stride, int channels) {
    const int track = blockIdx.y;
    const long long i = (long long)blockIdx.x * kPwThreads + threadIdx.x;
    if (i >= n) return;
    const int16_t* src = il + ((size_t)track * (size_t)n + (size_t)i) * channels;
    for (int c = 0; c < channels; ++c)
        pl[(size_t)(track * channels + c) * (size_t)stride + kLead + i] = (float)src[c] * (1.0f / 32768.0f);
}
__global__ void __launch_bounds__(kPwThreads) interleave_kernel(const float* __restrict__ pl, float* __restrict__ il,
                                                                long long n, long long stride, int channels) {
    const int track = blockIdx.y;
    const long long i = (long long)blockIdx.x * kPwThreads + threadIdx.x;
    if (i >= n) return;
    float* dst = il + ((size_t)track * (size_t)n + (size_t)i) * channels;
    for (int c = 0; c < channels; ++c) dst[c] = pl[(size_t)(track * channels + c) * (size_t)stride + kLead + i];
}

}  // namespace mm
