// apply_reverb (backend/app/pipeline.py:1055-1176): Schroeder reverb -- four parallel comb filters
//   y[n] = x[n] + g y[n - D]                          (_comb_filter, :1065-1078)
// averaged, two allpass filters in series
//   y[n] = -g x[n] + x[n - D] + g y[n - D]            (_allpass_filter, :1082-1094)
// peak normalisation of the wet signal, dry/wet mix, clip; optionally on mid / side with separate mixes.  Float64
// throughout, like the reference's numba loops.
//
// A recurrence with delay D is D independent first-order recurrences, one per phase n mod D (D = 130 ... 12000 samples).
// One thread owns one (row, phase) and walks it sequentially; the threads of a warp hold consecutive phases, so every step
// of the walk is a coalesced access.  rows x D threads (10^5 ... 10^6) keep the memory system busy although each thread is
// a dependent chain.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "context.h"
#include "stages_internal.h"

namespace mm {

struct RevSrc {                  // the signal a row of the reverb works on
    const float* in;
    long long n, stride;
    int channels;
    int ms;                      // 1: rows are (mid, side) of the stereo pair instead of (L, R)
};
__device__ __forceinline__ double rev_x(const RevSrc& S, int row, long long i) {
    const float* r = S.in + (size_t)row * (size_t)S.stride + kLead;
    if (!S.ms) return (double)r[i];
    const int track = row >> 1;
    const float l = S.in[(size_t)(2 * track) * (size_t)S.stride + kLead + i], rr = S.in[(size_t)(2 * track + 1) * (size_t)S.stride + kLead + i];
    // ((L + R) * 0.5).astype(float64) on float32 arrays (pipeline.py:1137-1138)
    return (double)((row & 1) ? __fmul_rn(__fsub_rn(l, rr), 0.5f) : __fmul_rn(__fadd_rn(l, rr), 0.5f));
}

struct CombArgs { RevSrc S; double* wet; long long D; double g; int accumulate; int rows; };
__global__ void __launch_bounds__(128) reverb_comb_kernel(const CombArgs P) {
    const long long p = (long long)blockIdx.x * 128 + threadIdx.x;
    const int row = blockIdx.y;
    if (p >= P.D || p >= P.S.n) return;
    double* w = P.wet + (size_t)row * (size_t)P.S.stride;
    double y = 0.0;
    for (long long i = p; i < P.S.n; i += P.D) {
        y = rev_x(P.S, row, i) + P.g * y;
        w[i] = P.accumulate ? w[i] + y : y;
    }
}
// the degenerate comb (delay >= n): _comb_filter returns x itself
__global__ void __launch_bounds__(256) reverb_addx_kernel(const CombArgs P) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int row = blockIdx.y;
    if (i >= P.S.n) return;
    double* w = P.wet + (size_t)row * (size_t)P.S.stride;
    const double x = rev_x(P.S, row, i);
    w[i] = P.accumulate ? w[i] + x : x;
}

struct ApArgs { const double* src; double* dst; long long n, stride, D; double g, scale; };
__global__ void __launch_bounds__(128) reverb_allpass_kernel(const ApArgs P) {
    const long long p = (long long)blockIdx.x * 128 + threadIdx.x;
    const int row = blockIdx.y;
    if (p >= P.D || p >= P.n) return;
    const double* s = P.src + (size_t)row * (size_t)P.stride;
    double* d = P.dst + (size_t)row * (size_t)P.stride;
    double xp = 0.0, yp = 0.0;
    for (long long i = p; i < P.n; i += P.D) {
        const double x = s[i] * P.scale;
        const double y = -P.g * x + xp + P.g * yp;
        d[i] = y;
        xp = x; yp = y;
    }
}
__global__ void __launch_bounds__(256) reverb_scale_kernel(const ApArgs P) {      // no allpass ran: dst = src * scale
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int row = blockIdx.y;
    if (i < P.n) P.dst[(size_t)row * (size_t)P.stride + i] = P.src[(size_t)row * (size_t)P.stride + i] * P.scale;
}

__global__ void __launch_bounds__(256) reverb_peak_kernel(const double* wet, long long n, long long stride, unsigned long long* peak_bits) {
    const int row = blockIdx.y;
    double pk = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        pk = fmax(pk, fabs(wet[(size_t)row * (size_t)stride + i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pk = fmax(pk, shfl_xor_d(pk, o));
    if ((threadIdx.x & 31) == 0) atomicMax(peak_bits + row, (unsigned long long)__double_as_longlong(pk));
}

struct MixArgs { RevSrc S; const double* wet; const unsigned long long* peak_bits; float* out; double mix[2]; };
__global__ void __launch_bounds__(256) reverb_mix_kernel(const MixArgs P) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int track = blockIdx.y;
    if (i >= P.S.n) return;
    const int C = P.S.channels;
    float o[2] = {0.f, 0.f};
    for (int c = 0; c < C; ++c) {
        const int row = track * C + c;
        const double peak = __longlong_as_double((long long)P.peak_bits[row]);
        double w = P.wet[(size_t)row * (size_t)P.S.stride + i];
        if (peak > 1e-6) w = w / fmin(peak, 2.0);                                  // pipeline.py:1114-1116
        const double m = P.mix[P.S.ms ? c : 0];
        o[c] = (float)(rev_x(P.S, row, i) * (1.0 - m) + w * m);                     // (...).astype(float32)
    }
    float* dst = P.out + (size_t)(track * C) * (size_t)P.S.stride + kLead + i;
    if (P.S.ms) {                                                                   // pipeline.py:1147-1149, float32
        dst[0] = fminf(fmaxf(__fadd_rn(o[0], o[1]), -1.f), 1.f);
        dst[P.S.stride] = fminf(fmaxf(__fsub_rn(o[0], o[1]), -1.f), 1.f);
    } else {
        for (int c = 0; c < C; ++c) dst[(size_t)c * (size_t)P.S.stride] = fminf(fmaxf(o[c], -1.f), 1.f);
    }
}

int st_reverb(mm_ctx* c, const mm_geom* g, const float* in, float* out, int type, double decay_sec, double mix, int use_ms,
              double mix_mid, double mix_side) {
    static const double decay0[5] = {1.2, 0.6, 2.2, 3.5, 5.0};
    static const double comb_ms[5][4] = {{29, 37, 41, 53}, {23, 31, 43, 47}, {47, 53, 61, 71}, {59, 67, 73, 83}, {97, 103, 109, 127}};
    static const double comb_g[5][4] = {{0.7, 0.65, 0.6, 0.55}, {0.5, 0.45, 0.4, 0.35}, {0.75, 0.7, 0.65, 0.6}, {0.78, 0.73, 0.68, 0.63},
                                        {0.82, 0.78, 0.74, 0.7}};
    static const double ap_ms[5][2] = {{5, 7}, {3, 5}, {8, 11}, {10, 14}, {15, 19}};
    static const double ap_g[5][2] = {{0.5, 0.4}, {0.4, 0.3}, {0.5, 0.45}, {0.52, 0.45}, {0.55, 0.48}};
    if (type < 0 || type > 4) type = 0;                                            // unknown name -> "plate"
    const double decay = decay_sec > 0 ? decay_sec : decay0[type];
    const double decay_per_sec = std::pow(0.001, 1.0 / std::max(0.1, decay));
    const int rows = g->tracks * g->channels;
    const bool ms = use_ms && g->channels == 2;
    double *wa, *wb;
    unsigned long long* pk;
    MM_TRY(arena(c, SL_REV0, (size_t)rows * (size_t)g->stride, &wa));
    MM_TRY(arena(c, SL_REV1, (size_t)rows * (size_t)g->stride, &wb));
    MM_TRY(arena(c, SL_XCHG, (size_t)rows, &pk));
    RevSrc S;
    S.in = in; S.n = g->n; S.stride = g->stride; S.channels = g->channels; S.ms = ms ? 1 : 0;
    bool first = true;
    for (int k = 0; k < 4; ++k) {
        const long long d = std::min<long long>((long long)((double)g->sr * comb_ms[type][k] / 1000.0), g->n - 1);
        if (d < 1) continue;
        CombArgs A;
        A.S = S; A.wet = wa; A.D = d; A.g = comb_g[type][k] * std::pow(decay_per_sec, comb_ms[type][k] / 1000.0);
        A.accumulate = first ? 0 : 1; A.rows = rows;
        KernelScope ks(c, "reverb_comb");
        if (d >= g->n) reverb_addx_kernel<<<dim3((unsigned)((g->n + 255) / 256), rows), 256, 0, c->stream>>>(A);
        else reverb_comb_kernel<<<dim3((unsigned)((d + 127) / 128), rows), 128, 0, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
        first = false;
    }
    if (first) MM_CUDA(cudaMemsetAsync(wa, 0, (size_t)rows * (size_t)g->stride * sizeof(double), c->stream));
    double *src = wa, *dst = wb;
    double scale = 1.0 / 4.0;                                                       // wet /= max(len(comb_delays_ms), 1)
    for (int k = 0; k < 2; ++k) {
        const long long d = std::min<long long>((long long)((double)g->sr * ap_ms[type][k] / 1000.0), g->n - 1);
        if (d < 1 || d >= g->n) continue;                                           // _allpass_filter returns x
        ApArgs A;
        A.src = src; A.dst = dst; A.n = g->n; A.stride = g->stride; A.D = d; A.g = ap_g[type][k]; A.scale = scale;
        KernelScope ks(c, "reverb_allpass");
        reverb_allpass_kernel<<<dim3((unsigned)((d + 127) / 128), rows), 128, 0, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
        std::swap(src, dst);
        scale = 1.0;
    }
    if (scale != 1.0) {
        ApArgs A;
        A.src = src; A.dst = dst; A.n = g->n; A.stride = g->stride; A.D = 0; A.g = 0; A.scale = scale;
        reverb_scale_kernel<<<dim3((unsigned)((g->n + 255) / 256), rows), 256, 0, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
        std::swap(src, dst);
    }
    MM_CUDA(cudaMemsetAsync(pk, 0, (size_t)rows * sizeof(unsigned long long), c->stream));
    {
        KernelScope ks(c, "reverb_peak");
        reverb_peak_kernel<<<dim3((unsigned)std::min<long long>((g->n + 255) / 256, 512), rows), 256, 0, c->stream>>>(src, g->n, g->stride, pk);
        MM_CUDA(cudaGetLastError());
    }
    MixArgs M;
    M.S = S; M.wet = src; M.peak_bits = pk; M.out = out;
    auto clamp01 = [](double v) { return std::max(0.0, std::min(1.0, v)); };
    M.mix[0] = ms ? clamp01(mix_mid) : mix;
    M.mix[1] = ms ? clamp01(mix_side) : mix;
    KernelScope ks(c, "reverb_mix");
    reverb_mix_kernel<<<dim3((unsigned)((g->n + 255) / 256), g->tracks), 256, 0, c->stream>>>(M);
    MM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mm
