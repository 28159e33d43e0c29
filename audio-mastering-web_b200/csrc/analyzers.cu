// Analyzer kernels: 4x-oversampled true peak, 4096-point spectrum bars, stereo correlation.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "context.h"
#include "stages_internal.h"

namespace mm {

// ---------------------------------------------------------------------------------------------------
// _true_peak_dbfs (backend/app/routers/tools.py:44-54): resample_poly(x, 4, 1) then max |.|
//
// scipy builds h = firwin(81, 1/4, window=('kaiser', 5.0)) * 4 and evaluates
//   y[4 t + p] = sum_d h[4 d + p] * x[t + 10 - d],   d = 0..20 (p = 0) or 0..19 (p = 1..3),
// with x = 0 outside [0, n): four polyphase branches of 21/20/20/20 taps.  Staged through shared
// memory with a 10-sample halo on both sides and fused with the max reduction.
// ---------------------------------------------------------------------------------------------------
constexpr int kTpThreads = 256;
constexpr int kTpPer = 8;                         // input samples per thread
constexpr int kTpTile = kTpThreads * kTpPer;      // 2048 input samples per CTA
constexpr int kTpHalo = 10;
constexpr int kTpTaps = 21;

struct TpCoef { float h[4][kTpTaps]; };
struct CorrAcc { double sl, sr, slr, sll, srr; unsigned pk; unsigned pad; };           // h[p][d] = h81[4 d + p] (0 where 4 d + p > 80)

__global__ void __launch_bounds__(kTpThreads) true_peak_kernel(const float* __restrict__ in, long long n, long long stride,
                                                               int channels, const __grid_constant__ TpCoef K,
                                                               float* __restrict__ peak_bits) {
    __shared__ float sx[kTpTile + 2 * kTpHalo + 2];
    const int row = blockIdx.y;
    const float* src = in + (size_t)row * (size_t)stride + kLead;
    const long long base = (long long)blockIdx.x * kTpTile;
    for (int j = threadIdx.x; j < kTpTile + 2 * kTpHalo; j += kTpThreads) {
        const long long i = base - kTpHalo + j;
        sx[j] = (i >= 0 && i < n) ? __ldcs(src + i) : 0.f;
    }
    __syncthreads();
    // thread handles t = base + tid + kTpThreads * u  (stride-1 across the warp: conflict-free LDS)
    float pk = 0.f;
#pragma unroll
    for (int u = 0; u < kTpPer; ++u) {
        const int tl = threadIdx.x + kTpThreads * u;          // local t
        if (base + tl >= n) break;
        // x[t + 10 - d] lives at sx[tl + 20 - d]
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int d = 0; d < kTpTaps; ++d) {
            const float xv = sx[tl + 2 * kTpHalo - d];
            a0 = fmaf(K.h[0][d], xv, a0);
            if (d < kTpTaps - 1) {
                a1 = fmaf(K.h[1][d], xv, a1);
                a2 = fmaf(K.h[2][d], xv, a2);
                a3 = fmaf(K.h[3][d], xv, a3);
            }
        }
        pk = fmaxf(pk, fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fmaxf(fabsf(a2), fabsf(a3))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    __shared__ float wp[kTpThreads / 32];
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = pk;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kTpThreads / 32; ++w) pk = fmaxf(pk, wp[w]);
        if (pk > 0.f) atomicMax(reinterpret_cast<int*>(peak_bits + row / channels), __float_as_int(pk));
    }
}

// Register-tiled variant: a thread owns 8 consecutive inputs (32 outputs of the 4x stream) and keeps their 27-sample window
// x[t0 - 9 .. t0 + 17] in registers, so one shared-memory load feeds up to 24 FMAs instead of 4.  Branch 0 of this
// Nyquist(4) filter is a pure delay -- h[4 d] = 0 for d != 10 (sinc zeros; scipy's taps there are ~1e-17) -- and is
// evaluated as the single product h[40] x[t]: 61 instead of 81 multiply-adds per input sample.  Shared-memory index m
// lives at m + m / 8, which makes the per-thread window loads (stride 8 floats across a warp) conflict free.
constexpr int kTp8Lead = 16;                      // local index of the tile's first sample
constexpr int kTp8Span = kTpTile + 2 * kTp8Lead;
constexpr int kTp8Buf = kTp8Span + kTp8Span / 8 + 8;
constexpr int kTp8Loads = (kTp8Span + kTpThreads - 1) / kTpThreads;
constexpr int kTp8Tiles = 8;                      // consecutive tiles per CTA: the next tile's samples are in flight (registers)
                                                  // while the current one is filtered, one barrier per tile
__global__ void __launch_bounds__(kTpThreads) true_peak_kernel8(const float* __restrict__ in, long long n, long long stride,
                                                                int channels, const __grid_constant__ TpCoef K,
                                                                float* __restrict__ peak_bits) {
    __shared__ float sx[2][kTp8Buf];
    const int row = blockIdx.y;
    const float* src = in + (size_t)row * (size_t)stride + kLead;
    const long long tile0 = (long long)blockIdx.x * kTp8Tiles;
    const long long ntiles = (n + kTpTile - 1) / kTpTile;
    const int T = (int)min((long long)kTp8Tiles, ntiles - tile0);
    float pre[kTp8Loads];
    auto fetch = [&](long long tile) {
        const long long base = tile * kTpTile - kTp8Lead;
#pragma unroll
        for (int r = 0; r < kTp8Loads; ++r) {
            const int m = threadIdx.x + kTpThreads * r;
            const long long i = base + m;
            pre[r] = (m < kTp8Span && i >= 0 && i < n) ? __ldcs(src + i) : 0.f;
        }
    };
    fetch(tile0);
    float pk = 0.f;
    const float h0 = K.h[0][10];
#pragma unroll 1
    for (int k = 0; k < T; ++k) {
        float* buf = sx[k & 1];
#pragma unroll
        for (int r = 0; r < kTp8Loads; ++r) {
            const int m = threadIdx.x + kTpThreads * r;
            if (m < kTp8Span) buf[m + (m >> 3)] = pre[r];
        }
        __syncthreads();
        if (k + 1 < T) fetch(tile0 + k + 1);
        const long long base = (tile0 + k) * kTpTile;
        // window element j = x[t0 - 9 + j] has local index m = 8 tid + 7 + j
        float xw[27];
        const float* wp0 = buf + 9 * threadIdx.x;
#pragma unroll
        for (int j = 0; j < 27; ++j) xw[j] = wp0[((7 + j) >> 3) * 9 + ((7 + j) & 7)];
        float a1[kTpPer], a2[kTpPer], a3[kTpPer];
#pragma unroll
        for (int u = 0; u < kTpPer; ++u) a1[u] = a2[u] = a3[u] = 0.f;
#pragma unroll
        for (int d = 0; d < kTpTaps - 1; ++d) {
            const float h1 = K.h[1][d], h2 = K.h[2][d], h3 = K.h[3][d];
#pragma unroll
            for (int u = 0; u < kTpPer; ++u) {
                const float xv = xw[u + 19 - d];              // x[t + 10 - d]
                a1[u] = fmaf(h1, xv, a1[u]);
                a2[u] = fmaf(h2, xv, a2[u]);
                a3[u] = fmaf(h3, xv, a3[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kTpPer; ++u) {
            const float a0 = h0 * xw[u + 9];                  // x[t]
            // outputs exist for t < n only (resample_poly emits 4 n samples)
            if (base + 8 * (long long)threadIdx.x + u < n)
                pk = fmaxf(pk, fmaxf(fmaxf(fabsf(a0), fabsf(a1[u])), fmaxf(fabsf(a2[u]), fabsf(a3[u]))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    __shared__ float wp[kTpThreads / 32];
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = pk;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kTpThreads / 32; ++w) pk = fmaxf(pk, wp[w]);
        if (pk > 0.f) atomicMax(reinterpret_cast<int*>(peak_bits + row / channels), __float_as_int(pk));
    }
}

// True peak and the stereo-correlation sums of a stereo track in ONE pass over its samples: the FIR is FP32 bound and leaves the
// memory system idle, the five float64 sums + sample peak ride on the centre samples the register windows already hold.
__global__ void __launch_bounds__(kTpThreads) true_peak_corr_kernel(const float* __restrict__ in, long long n, long long stride,
                                                                    const __grid_constant__ TpCoef K, float* __restrict__ peak_bits,
                                                                    CorrAcc* __restrict__ acc) {
    __shared__ float sx[2][2][kTp8Buf];
    const int track = blockIdx.y;
    const float* src0 = in + (size_t)(track * 2) * (size_t)stride + kLead;
    const float* src1 = src0 + stride;
    const long long tile0 = (long long)blockIdx.x * kTp8Tiles;
    const long long ntiles = (n + kTpTile - 1) / kTpTile;
    const int T = (int)min((long long)kTp8Tiles, ntiles - tile0);
    float pre[2][kTp8Loads];
    auto fetch = [&](long long tile) {
        const long long base = tile * kTpTile - kTp8Lead;
#pragma unroll
        for (int r = 0; r < kTp8Loads; ++r) {
            const int m = threadIdx.x + kTpThreads * r;
            const long long i = base + m;
            const bool ok = m < kTp8Span && i >= 0 && i < n;
            pre[0][r] = ok ? __ldcs(src0 + i) : 0.f;
            pre[1][r] = ok ? __ldcs(src1 + i) : 0.f;
        }
    };
    fetch(tile0);
    float pk = 0.f, spk = 0.f;
    double v[5] = {0, 0, 0, 0, 0};
    const float h0 = K.h[0][10];
#pragma unroll 1
    for (int k = 0; k < T; ++k) {
#pragma unroll
        for (int r = 0; r < kTp8Loads; ++r) {
            const int m = threadIdx.x + kTpThreads * r;
            if (m < kTp8Span) { sx[k & 1][0][m + (m >> 3)] = pre[0][r]; sx[k & 1][1][m + (m >> 3)] = pre[1][r]; }
        }
        __syncthreads();
        if (k + 1 < T) fetch(tile0 + k + 1);
        const long long base = (tile0 + k) * kTpTile;
        float xc[2][kTpPer];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float xw[27];
            const float* wp0 = sx[k & 1][c] + 9 * threadIdx.x;
#pragma unroll
            for (int j = 0; j < 27; ++j) xw[j] = wp0[((7 + j) >> 3) * 9 + ((7 + j) & 7)];
            float a1[kTpPer], a2[kTpPer], a3[kTpPer];
#pragma unroll
            for (int u = 0; u < kTpPer; ++u) a1[u] = a2[u] = a3[u] = 0.f;
#pragma unroll
            for (int d = 0; d < kTpTaps - 1; ++d) {
                const float h1 = K.h[1][d], h2 = K.h[2][d], h3 = K.h[3][d];
#pragma unroll
                for (int u = 0; u < kTpPer; ++u) {
                    const float xv = xw[u + 19 - d];
                    a1[u] = fmaf(h1, xv, a1[u]);
                    a2[u] = fmaf(h2, xv, a2[u]);
                    a3[u] = fmaf(h3, xv, a3[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kTpPer; ++u) {
                xc[c][u] = xw[u + 9];
                const float a0 = h0 * xw[u + 9];
                if (base + 8 * (long long)threadIdx.x + u < n)
                    pk = fmaxf(pk, fmaxf(fmaxf(fabsf(a0), fabsf(a1[u])), fmaxf(fabsf(a2[u]), fabsf(a3[u]))));
            }
        }
#pragma unroll
        for (int u = 0; u < kTpPer; ++u) {                   // samples past n were staged as zeros: they add nothing
            const double l = (double)xc[0][u], rr = (double)xc[1][u];
            v[0] += l; v[1] += rr; v[2] = fma(l, rr, v[2]); v[3] = fma(l, l, v[3]); v[4] = fma(rr, rr, v[4]);
            spk = fmaxf(spk, fmaxf(fabsf(xc[0][u]), fabsf(xc[1][u])));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        spk = fmaxf(spk, __shfl_xor_sync(0xffffffffu, spk, o));
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] += shfl_xor_d(v[q], o);
    }
    __shared__ float wp[kTpThreads / 32], ws[kTpThreads / 32];
    __shared__ double sv[kTpThreads / 32][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { wp[warp] = pk; ws[warp] = spk; for (int q = 0; q < 5; ++q) sv[warp][q] = v[q]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kTpThreads / 32; ++w) {
            pk = fmaxf(pk, wp[w]); spk = fmaxf(spk, ws[w]);
            for (int q = 0; q < 5; ++q) v[q] += sv[w][q];
        }
        if (pk > 0.f) atomicMax(reinterpret_cast<int*>(peak_bits + track), __float_as_int(pk));
        CorrAcc* a = acc + track;
        atomicAdd(&a->sl, v[0]); atomicAdd(&a->sr, v[1]); atomicAdd(&a->slr, v[2]); atomicAdd(&a->sll, v[3]); atomicAdd(&a->srr, v[4]);
        atomicMax(&a->pk, __float_as_uint(spk));
    }
}

__global__ void peak_to_db_kernel(const float* peak_bits, int tracks, double* db) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < tracks) db[t] = 20.0 * log10(fmax((double)peak_bits[t], 1e-12));
}

// firwin(81, 0.25, window=('kaiser', 5.0)) * 4 in float64: windowed sinc scaled to unit DC gain.
static double bessel_i0(double x) {
    double s = 1.0, t = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 200; ++k) {
        t *= q / ((double)k * (double)k);
        s += t;
        if (t < 1e-17 * s) break;
    }
    return s;
}
static void true_peak_fir(double* h81) {
    const double pi = 3.14159265358979323846;
    const int N = 81;
    const double alpha = 0.5 * (N - 1), cutoff = 0.25, beta = 5.0;
    double sum = 0.0;
    for (int j = 0; j < N; ++j) {
        const double m = j - alpha;
        // firwin: h = right * sinc(right * m) - left * sinc(left * m) with left = 0, right = cutoff
        const double arg = cutoff * m;
        const double sinc = (arg == 0.0) ? 1.0 : std::sin(pi * arg) / (pi * arg);
        const double r = 2.0 * j / (N - 1) - 1.0;
        const double win = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(beta);
        h81[j] = cutoff * sinc * win;
        sum += h81[j];
    }
    // scale_frequency = 0 for a low-pass: unit gain at DC; resample_poly then multiplies by `up`
    for (int j = 0; j < N; ++j) h81[j] = h81[j] / sum * 4.0;
}

static TpCoef g_tp_coef;
static bool g_tp_delay0 = false;
static const TpCoef* tp_coef() { return &g_tp_coef; }
// builds the polyphase coefficients once; returns whether branch 0 is the pure delay the register-tiled kernels assume
static bool tp_build_coef() {
    static std::once_flag once;
    std::call_once(once, [] {
        TpCoef& K = g_tp_coef;
        bool& delay0 = g_tp_delay0;
        {
        double h[81];
        true_peak_fir(h);
        memset(&K, 0, sizeof(K));
        for (int p = 0; p < 4; ++p)
            for (int d = 0; d < kTpTaps; ++d)
                if (4 * d + p <= 80) K.h[p][d] = (float)h[4 * d + p];
        delay0 = true;                                  // branch 0 is a pure delay (all other taps are sinc zeros)?
        for (int d = 0; d < kTpTaps; ++d)
            if (d != 10 && std::fabs(h[4 * d]) > 1e-12) delay0 = false;
        }
    });
    return g_tp_delay0;
}

int st_true_peak(mm_ctx* c, const mm_geom* g, const float* in, double* tp_dev) {
    const bool delay0 = tp_build_coef();
    const TpCoef& K = g_tp_coef;
    float* bits;
    MM_TRY(arena(c, SL_PEAKBITS, (size_t)g->tracks, &bits));
    MM_CUDA(cudaMemsetAsync(bits, 0, (size_t)g->tracks * sizeof(float), c->stream));
    const int rows = g->tracks * g->channels;
    dim3 grid((unsigned)((g->n + kTpTile - 1) / kTpTile), (unsigned)rows);
    {
        KernelScope ks(c, "true_peak_fir4x_max");
        const dim3 grid8((unsigned)((g->n + (long long)kTpTile * kTp8Tiles - 1) / ((long long)kTpTile * kTp8Tiles)), (unsigned)rows);
        if (delay0) true_peak_kernel8<<<grid8, kTpThreads, 0, c->stream>>>(in, g->n, g->stride, g->channels, K, bits);
        else true_peak_kernel<<<grid, kTpThreads, 0, c->stream>>>(in, g->n, g->stride, g->channels, K, bits);
    }
    MM_CUDA(cudaGetLastError());
    {
        KernelScope ks(c, "peak_to_db");
        peak_to_db_kernel<<<(g->tracks + 127) / 128, 128, 0, c->stream>>>(bits, g->tracks, tp_dev);
    }
    MM_CUDA(cudaGetLastError());
    return 0;
}

int mm_true_peak_fir_host(double* h81) {
    true_peak_fir(h81);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// compute_spectrum_bars (backend/app/pipeline.py:700-739): Hann-windowed 4096-point FFT of the
// mid-file frame, 64 log-spaced bars of max |X| * 2/N in dB.  One CTA per track: radix-4 Stockham
// autosort FFT (6 passes of 1024 butterflies) ping-ponging between two shared-memory buffers.
// ---------------------------------------------------------------------------------------------------
constexpr int kFftN = 4096;
constexpr int kFftThreads = 256;
constexpr int kBars = 64;

struct SpecArgs {
    const float* in;
    long long n, stride;
    int channels, view, sr;
    double* bars;        // [tracks][64]
};

__global__ void __launch_bounds__(kFftThreads) spectrum_kernel(const SpecArgs P) {
    extern __shared__ float2 fsm[];
    float2* A = fsm;
    float2* B = fsm + kFftN;
    __shared__ float mag[kFftN / 2 + 1];
    const int track = blockIdx.x;
    double* out = P.bars + (size_t)track * kBars;
    // pipeline.py:708-709: fewer than n_fft samples in total -> all bars at -80
    if (P.n * P.channels < kFftN || P.n < kFftN) {
        if (threadIdx.x < kBars) out[threadIdx.x] = -80.0;
        return;
    }
    const float* r0 = P.in + (size_t)(track * P.channels) * (size_t)P.stride + kLead;
    const float* r1 = r0 + (P.channels > 1 ? (size_t)P.stride : 0);
    const long long start = max(0LL, P.n / 2 - kFftN / 2);
    for (int j = threadIdx.x; j < kFftN; j += kFftThreads) {
        const float l = r0[start + j], r = r1[start + j];
        float m;
        if (P.channels == 1) m = l;
        else if (P.view == 2) m = __fmul_rn(__fsub_rn(l, r), 0.5f);
        else m = __fmul_rn(__fadd_rn(l, r), 0.5f);            // mean of two channels == mid
        // np.hanning(M): 0.5 - 0.5 cos(2 pi j / (M - 1))
        const float w = 0.5f - 0.5f * cospif(2.0f * (float)j / (float)(kFftN - 1));
        A[j] = make_float2(m * w, 0.f);
    }
    __syncthreads();
    // Stockham radix-4, decimation in frequency: n = 4096 = 4^6
    int Ns = 1;
    float2* src = A;
    float2* dst = B;
#pragma unroll 1
    for (int pass = 0; pass < 6; ++pass) {
        for (int j = threadIdx.x; j < kFftN / 4; j += kFftThreads) {
            const int k = j & (Ns - 1);                         // index inside the current sub-transform
            float2 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = src[j + r * (kFftN / 4)];
            // twiddle: w^(r k), w = exp(-2 pi i / (4 Ns))
            const float ang = -2.0f * (float)k / (float)(4 * Ns);   // in units of pi
#pragma unroll
            for (int r = 1; r < 4; ++r) {
                float s, cth;
                sincospif(ang * (float)r, &s, &cth);
                const float2 t = v[r];
                v[r] = make_float2(t.x * cth - t.y * s, t.x * s + t.y * cth);
            }
            // 4-point DFT
            const float2 a02 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
            const float2 s02 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
            const float2 a13 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
            const float2 s13 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
            const float2 y0 = make_float2(a02.x + a13.x, a02.y + a13.y);
            const float2 y2 = make_float2(a02.x - a13.x, a02.y - a13.y);
            const float2 y1 = make_float2(s02.x + s13.y, s02.y - s13.x);    // s02 - i s13
            const float2 y3 = make_float2(s02.x - s13.y, s02.y + s13.x);    // s02 + i s13
            const int j0 = ((j - k) << 2) + k;                  // expand: (j / Ns) * 4 Ns + k
            dst[j0] = y0;
            dst[j0 + Ns] = y1;
            dst[j0 + 2 * Ns] = y2;
            dst[j0 + 3 * Ns] = y3;
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
        Ns <<= 2;
    }
    for (int k = threadIdx.x; k <= kFftN / 2; k += kFftThreads) {
        const float2 X = src[k];
        mag[k] = sqrtf(X.x * X.x + X.y * X.y) * (2.0f / (float)kFftN);
    }
    __syncthreads();
    if (threadIdx.x < kBars) {
        const int b = threadIdx.x;
        const double nyq = (double)P.sr / 2.0;
        const double f0 = 20.0 * pow(1000.0, (double)b / 63.0);
        const double f1 = 20.0 * pow(1000.0, (double)(b + 1) / 63.0);
        const int k0 = max(0, (int)((f0 / nyq) * (double)(kFftN / 2)));
        const int k1 = min(kFftN / 2, (int)ceil((f1 / nyq) * (double)(kFftN / 2)));
        double pk = 1e-12;
        if (k0 <= k1) {
            float m = mag[k0];
            for (int k = k0 + 1; k <= k1; ++k) m = fmaxf(m, mag[k]);
            pk = (double)m;
        }
        const double db = 20.0 * log10(fmax(pk, 1e-12));
        out[b] = rint(db * 100.0) / 100.0;
    }
}

int st_spectrum_bars(mm_ctx* c, const mm_geom* g, const float* in, int view, double* bars_dev) {
    if (view < 0 || view > 2) { set_error("spectrum view must be 0 (mean), 1 (mid) or 2 (side)"); return 1; }
    const size_t smem = 2 * kFftN * sizeof(float2);
    MM_TRY(kernel_setup(c, (const void*)spectrum_kernel, kFftThreads, smem, false, nullptr));
    SpecArgs A;
    A.in = in; A.n = g->n; A.stride = g->stride; A.channels = g->channels; A.view = view; A.sr = g->sr; A.bars = bars_dev;
    KernelScope ks(c, "spectrum_fft4096_bars");
    spectrum_kernel<<<g->tracks, kFftThreads, smem, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// measure_stereo_correlation (backend/app/pipeline.py:766-791) + sample peak: five float64 sums
// ---------------------------------------------------------------------------------------------------

constexpr int kCorrThreads = 256;
constexpr int kCorrPerBlock = kCorrThreads * 4 * 4;

__global__ void __launch_bounds__(kCorrThreads) corr_kernel(const float* __restrict__ in, long long n, long long stride,
                                                            int channels, CorrAcc* __restrict__ acc) {
    const int track = blockIdx.y;
    const float* r0 = in + (size_t)(track * channels) * (size_t)stride + kLead;
    const float* r1 = r0 + (channels > 1 ? (size_t)stride : 0);
    const long long base = (long long)blockIdx.x * kCorrPerBlock;
    double v[5] = {0, 0, 0, 0, 0};
    float pk = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const long long i = base + 4LL * (threadIdx.x + kCorrThreads * r);
        float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        if (i + 3 < n) {
            const float4 x = __ldcs(reinterpret_cast<const float4*>(r0 + i));
            a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w;
            if (channels > 1) { const float4 y = __ldcs(reinterpret_cast<const float4*>(r1 + i)); b[0] = y.x; b[1] = y.y; b[2] = y.z; b[3] = y.w; }
        } else {
            for (int k = 0; k < 4; ++k) if (i + k < n) { a[k] = r0[i + k]; if (channels > 1) b[k] = r1[i + k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double l = (double)a[k], rr = (double)b[k];
            v[0] += l; v[1] += rr; v[2] = fma(l, rr, v[2]); v[3] = fma(l, l, v[3]); v[4] = fma(rr, rr, v[4]);
            pk = fmaxf(pk, fmaxf(fabsf(a[k]), fabsf(b[k])));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] += shfl_xor_d(v[q], o);
        pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    }
    __shared__ double sv[kCorrThreads / 32][5];
    __shared__ float sp[kCorrThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { for (int q = 0; q < 5; ++q) sv[warp][q] = v[q]; sp[warp] = pk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kCorrThreads / 32; ++w) { for (int q = 0; q < 5; ++q) v[q] += sv[w][q]; pk = fmaxf(pk, sp[w]); }
        CorrAcc* a = acc + track;
        atomicAdd(&a->sl, v[0]); atomicAdd(&a->sr, v[1]); atomicAdd(&a->slr, v[2]); atomicAdd(&a->sll, v[3]); atomicAdd(&a->srr, v[4]);
        atomicMax(&a->pk, __float_as_uint(pk));
    }
}

__global__ void corr_final_kernel(const CorrAcc* acc, int tracks, int channels, long long n, double* corr, double* peak) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tracks) return;
    const CorrAcc a = acc[t];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (peak) peak[t] = (double)__uint_as_float(a.pk);
    if (!corr) return;
    double r;
    if (channels != 2 || n * 2 < 4) r = nan;                                 // None
    else if (sqrt(fmax(a.sll * a.srr, 0.0)) < 1e-20) r = nan;                // None
    else {
        const double N = (double)n;
        const double den = sqrt(fmax(N * a.sll - a.sl * a.sl, 0.0)) * sqrt(fmax(N * a.srr - a.sr * a.sr, 0.0));
        if (den < 1e-20) r = 0.0;
        else r = fmin(fmax((N * a.slr - a.sl * a.sr) / den, -1.0), 1.0);
    }
    corr[t] = r;
}

int st_correlation(mm_ctx* c, const mm_geom* g, const float* in, double* corr_dev, double* peak_dev) {
    CorrAcc* acc;
    MM_TRY(arena(c, SL_ENV0, (size_t)g->tracks, &acc));
    MM_CUDA(cudaMemsetAsync(acc, 0, (size_t)g->tracks * sizeof(CorrAcc), c->stream));
    dim3 grid((unsigned)((g->n + kCorrPerBlock - 1) / kCorrPerBlock), (unsigned)g->tracks);
    {
        KernelScope ks(c, "stereo_corr_sums");
        corr_kernel<<<grid, kCorrThreads, 0, c->stream>>>(in, g->n, g->stride, g->channels, acc);
    }
    MM_CUDA(cudaGetLastError());
    {
        KernelScope ks(c, "stereo_corr_final");
        corr_final_kernel<<<(g->tracks + 127) / 128, 128, 0, c->stream>>>(acc, g->tracks, g->channels, g->n, corr_dev, peak_dev);
    }
    MM_CUDA(cudaGetLastError());
    return 0;
}


// true peak + correlation + sample peak of stereo tracks in one pass (mono: the two separate kernels; no correlation there)
int st_true_peak_corr(mm_ctx* c, const mm_geom* g, const float* in, double* tp_dev, double* corr_dev, double* peak_dev) {
    if (g->channels != 2 || !tp_build_coef()) {
        MM_TRY(st_true_peak(c, g, in, tp_dev));
        return st_correlation(c, g, in, corr_dev, peak_dev);
    }
    float* bits;
    CorrAcc* acc;
    MM_TRY(arena(c, SL_PEAKBITS, (size_t)g->tracks, &bits));
    MM_TRY(arena(c, SL_ENV0, (size_t)g->tracks, &acc));
    MM_CUDA(cudaMemsetAsync(bits, 0, (size_t)g->tracks * sizeof(float), c->stream));
    MM_CUDA(cudaMemsetAsync(acc, 0, (size_t)g->tracks * sizeof(CorrAcc), c->stream));
    const dim3 grid((unsigned)((g->n + (long long)kTpTile * kTp8Tiles - 1) / ((long long)kTpTile * kTp8Tiles)), (unsigned)g->tracks);
    {
        KernelScope ks(c, "true_peak_fir4x_corr");
        true_peak_corr_kernel<<<grid, kTpThreads, 0, c->stream>>>(in, g->n, g->stride, *tp_coef(), bits, acc);
    }
    MM_CUDA(cudaGetLastError());
    {
        KernelScope ks(c, "peak_to_db");
        peak_to_db_kernel<<<(g->tracks + 127) / 128, 128, 0, c->stream>>>(bits, g->tracks, tp_dev);
    }
    {
        KernelScope ks(c, "stereo_corr_final");
        corr_final_kernel<<<(g->tracks + 127) / 128, 128, 0, c->stream>>>(acc, g->tracks, g->channels, g->n, corr_dev, peak_dev);
    }
    MM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mm
