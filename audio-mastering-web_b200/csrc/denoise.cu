// apply_spectral_denoise (backend/app/pipeline.py:1472-1524): STFT Wiener gain with a per-bin percentile noise floor.
//
// The reference calls scipy.signal.stft / istft (legacy _spectral_helper): nperseg 2048, hop 512, periodic Hann,
// boundary='zeros' (1024 zeros either side), padded=True (tail padded to a whole hop), scaling='spectrum' (divide by
// sum(window) = 1024).  Frame f covers samples [512 f - 1024, 512 f + 1024); F = ceil(n / 512) + 1 frames per row.
//
//   dn_mag_kernel   : |Z[k, f]| for every frame, written bin-major ([row][bin][frame]) so that a bin's time series is
//                     contiguous.  Each 2048-point real FFT is one packed 1024-point complex Stockham FFT (fft.cuh).
//   dn_floor_kernel : per (row, bin) the order statistics numpy's percentile (linear interpolation) and median need, by
//                     an 8-bit-digit radix select over the float bit patterns (magnitudes are >= 0, so the patterns sort
//                     like the values); noise = min(max(percentile, 1e-12), 0.85 max(median, 1e-12)).
//   dn_apply_kernel : a CTA owns 13 hops of output; it recomputes the 16 frames that touch them (cheaper than storing the
//                     complex STFT: 4 n floats per row), applies gain = clip(1 - strength (noise / (|Z| + 1e-10))^2,
//                     0.25, 1) between the forward and inverse untangling steps, inverse-transforms, windows and
//                     overlap-adds in shared memory, then divides by the overlap-added squared window and clips.
#include <algorithm>
#include <cmath>

#include "context.h"
#include "fft.cuh"
#include "stages_internal.h"

namespace mm {

constexpr int kDnN = 2048;              // nperseg
constexpr int kDnM = kDnN / 2;          // packed complex FFT size
constexpr int kDnHop = 512;
constexpr int kDnBins = kDnN / 2 + 1;
constexpr int kDnThreads = 256;         // four groups of 64 threads, one frame per group at a time
constexpr int kDnGroups = kDnThreads / 64;
constexpr int kDnBuf = kDnM + kDnM / 16;    // padded FFT buffer: index i lives at i + i / 16
constexpr int kDnMagFrames = 8;         // frames per CTA of the magnitude pass (32-byte runs per bin)
constexpr int kDnHops = 13;             // hops of output per CTA of the apply pass
constexpr int kDnApplyFrames = kDnHops + 3;
constexpr int kDnAccLen = (kDnHops + 6) * kDnHop;

struct DnSmem {                          // carved from dynamic shared memory, in this order
    float* win;        // [kDnN]   periodic Hann rounded to float32 (the same values window, synthesise and normalise)
    float2* tw;        // [kDnM]   e^{-2 pi i k / 1024}
    float2* twh;       // [kDnM/2 + 1]  e^{-2 pi i k / 2048}
    float2* buf;       // [kDnGroups][kDnBuf]
    unsigned char* rest;
};
constexpr size_t kDnSmemHead = kDnN * sizeof(float) + (kDnM + kDnM / 2 + 1 + 1 + kDnGroups * kDnBuf) * sizeof(float2);

__device__ __forceinline__ DnSmem dn_carve(unsigned char* base) {
    DnSmem S;
    S.win = reinterpret_cast<float*>(base);
    S.tw = reinterpret_cast<float2*>(S.win + kDnN);
    S.twh = S.tw + kDnM;
    S.buf = S.twh + kDnM / 2 + 2;
    S.rest = reinterpret_cast<unsigned char*>(S.buf + kDnGroups * kDnBuf);
    return S;
}

__device__ __forceinline__ void dn_fill_tables(const DnSmem& S) {
    fft_fill_twiddles<kDnM, kDnThreads>(S.tw);
    for (int k = threadIdx.x; k <= kDnM / 2; k += kDnThreads) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)kDnN, &s, &c);
        S.twh[k] = make_float2((float)c, (float)s);
    }
    for (int j = threadIdx.x; j < kDnN; j += kDnThreads) S.win[j] = (float)(0.5 - 0.5 * cospi(2.0 * (double)j / (double)kDnN));
}

__device__ __forceinline__ void gbar(int group) { asm volatile("bar.sync %0, 64;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ int padx(int i) { return i + (i >> 4); }

// 1024-point forward FFT by one 64-thread group, 16 points per thread: radix 16 (inputs v[q] = z[gt + 64 q]), exchange,
// radix 16, exchange, radix 4.  On return v[m] = Z[gt + 64 m] (natural order).  `buf` is free to be overwritten on return
// only after a further gbar().
__device__ __forceinline__ void fft1024_group(float2 (&v)[16], float2* buf, const float2* tw, int gt, int group) {
    dft_reg<16>(v);
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[gt * 17 + k] = v[dft_pos<16>(k)];                 // padx(16 gt + k)
    gbar(group);
    const int kk = gt & 15;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        float2 x = buf[padx(gt + 64 * q)];
        if (q) x = cmulf(x, tw[4 * kk * q]);
        v[q] = x;
    }
    gbar(group);
    dft_reg<16>(v);
    const int o = (gt - kk) * 16 + kk;
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[padx(o + 16 * k)] = v[dft_pos<16>(k)];
    gbar(group);
    float2 w[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int jj = gt + 64 * i;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 x = buf[padx(jj + 256 * q)];
            if (q) x = cmulf(x, tw[jj * q]);
            w[4 * i + q] = x;
        }
        dft4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[i + 4 * q] = w[4 * i + q];                        // Z[gt + 64 (i + 4 q)]
}

// windowed frame f of one row, packed even / odd: v[q] = z[gt + 64 q] (zeros outside [0, n))
__device__ __forceinline__ void dn_load_frame(const float* row, long long n, long long f, const float* win, int gt, float2 (&v)[16]) {
    const long long t0 = f * kDnHop - kDnN / 2;
    if (t0 >= 0 && t0 + kDnN <= n) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int m = gt + 64 * q;
            const float2 x = *reinterpret_cast<const float2*>(row + t0 + 2 * m);
            const float2 w = *reinterpret_cast<const float2*>(win + 2 * m);
            v[q] = make_float2(x.x * w.x, x.y * w.y);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int m = gt + 64 * q;
            const long long t = t0 + 2 * m;
            const float x0 = (t >= 0 && t < n) ? row[t] : 0.0f;
            const float x1 = (t + 1 >= 0 && t + 1 < n) ? row[t + 1] : 0.0f;
            v[q] = make_float2(x0 * win[2 * m], x1 * win[2 * m + 1]);
        }
    }
}

// X[k] and X[M - k] of the 2M-point real FFT from the packed transform Z (k = 0 .. M/2; k = 0 yields X[0] and X[M]); Z padded
__device__ __forceinline__ void dn_untangle(const float2* Z, const float2* twh, int k, float2& xk, float2& xm) {
    if (k == 0) {
        xk = make_float2(Z[0].x + Z[0].y, 0.0f);
        xm = make_float2(Z[0].x - Z[0].y, 0.0f);
        return;
    }
    const float2 zk = Z[padx(k)], zm = Z[padx(kDnM - k)];
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    const float2 d = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));        // (Z[k] - conj Z[M-k]) / 2
    const float2 wo = cmulf(make_float2(d.y, -d.x), twh[k]);                          // W^k * (-i d)
    xk = make_float2(e.x + wo.x, e.y + wo.y);
    xm = make_float2(e.x - wo.x, -(e.y - wo.y));                                      // conj(E - W^k O)
}

struct DnMagArgs {
    const float* in;
    long long n, stride, frames, fpad;
    float* mag;        // [rows][kDnBins][fpad]
};

__global__ void __launch_bounds__(kDnThreads) dn_mag_kernel(const DnMagArgs P) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const DnSmem S = dn_carve(dsm);
    float* stage = reinterpret_cast<float*>(S.rest);                                  // [kDnMagFrames][kDnBins]
    dn_fill_tables(S);
    __syncthreads();
    const int group = threadIdx.x >> 6, gt = threadIdx.x & 63;
    float2* buf = S.buf + group * kDnBuf;
    const int rowi = blockIdx.y;
    const float* row = P.in + (size_t)rowi * (size_t)P.stride + kLead;
    const long long f0 = (long long)blockIdx.x * kDnMagFrames;
    const int nf = (int)min((long long)kDnMagFrames, P.frames - f0);
    for (int fl = group; fl < nf; fl += kDnGroups) {                                  // group-uniform trip count
        float2 v[16];
        dn_load_frame(row, P.n, f0 + fl, S.win, gt, v);
        gbar(group);                                                                  // the previous frame's readers are done
        fft1024_group(v, buf, S.tw, gt, group);
        gbar(group);
#pragma unroll
        for (int m = 0; m < 16; ++m) buf[padx(gt + 64 * m)] = v[m];
        gbar(group);
        for (int k = gt; k <= kDnM / 2; k += 64) {
            float2 xk, xm;
            dn_untangle(buf, S.twh, k, xk, xm);
            stage[fl * kDnBins + k] = sqrtf(xk.x * xk.x + xk.y * xk.y) * (1.0f / 1024.0f);
            stage[fl * kDnBins + kDnM - k] = sqrtf(xm.x * xm.x + xm.y * xm.y) * (1.0f / 1024.0f);
        }
    }
    __syncthreads();
    float* dst = P.mag + (size_t)rowi * kDnBins * (size_t)P.fpad + f0;
    for (int i = threadIdx.x; i < kDnBins * kDnMagFrames; i += kDnThreads) {
        const int k = i / kDnMagFrames, fl = i % kDnMagFrames;
        if (fl < nf) dst[(size_t)k * (size_t)P.fpad + fl] = stage[fl * kDnBins + k];
    }
}

struct DnFloorArgs {
    const float* mag;
    long long frames, fpad;
    long long rank[4];      // percentile lo, hi; median lo, hi (0-based order statistics)
    double gamma;           // percentile interpolation weight
    float* noise;           // [rows][kDnBins]
};

__global__ void __launch_bounds__(kDnThreads) dn_floor_kernel(const DnFloorArgs P) {
    __shared__ unsigned hist[4][256];
    __shared__ unsigned prefix[4];
    __shared__ long long remain[4];
    __shared__ int owner[4];
    const unsigned* series = reinterpret_cast<const unsigned*>(P.mag + ((size_t)blockIdx.y * kDnBins + blockIdx.x) * (size_t)P.fpad);
    if (threadIdx.x < 4) {
        prefix[threadIdx.x] = 0;
        remain[threadIdx.x] = P.rank[threadIdx.x];
    }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 4 * 256; i += kDnThreads) (&hist[0][0])[i] = 0;
        if (threadIdx.x < 4) {
            int o = threadIdx.x;                      // targets that share a prefix share a histogram
            for (int t = threadIdx.x - 1; t >= 0; --t)
                if (prefix[t] == prefix[threadIdx.x]) o = t;
            owner[threadIdx.x] = o;
        }
        __syncthreads();
        unsigned pf[4];
        bool own[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { pf[t] = prefix[t]; own[t] = owner[t] == t; }
        for (long long i = threadIdx.x; i < P.frames; i += kDnThreads) {
            const unsigned key = series[i];
            const unsigned hi = shift == 24 ? 0u : key >> (shift + 8);
            const unsigned digit = (key >> shift) & 255u;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (own[t] && hi == pf[t]) atomicAdd(&hist[t][digit], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            const unsigned* h = hist[owner[threadIdx.x]];
            long long r = remain[threadIdx.x];
            int d = 0;
            for (; d < 255; ++d) {
                if (r < (long long)h[d]) break;
                r -= h[d];
            }
            remain[threadIdx.x] = r;
            prefix[threadIdx.x] = (prefix[threadIdx.x] << 8) | (unsigned)d;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double a = (double)__uint_as_float(prefix[0]), b = (double)__uint_as_float(prefix[1]);
        const double m0 = (double)__uint_as_float(prefix[2]), m1 = (double)__uint_as_float(prefix[3]);
        // numpy _lerp: a + (b - a) t, and b - (b - a)(1 - t) for t >= 0.5
        const double diff = b - a, t = P.gamma;
        const double pct = t >= 0.5 ? b - diff * (1.0 - t) : a + diff * t;
        const double med = P.rank[2] == P.rank[3] ? m0 : (m0 + m1) / 2.0;
        const double cap = fmin(fmax(pct, 1e-12), 0.85 * fmax(med, 1e-12));
        P.noise[(size_t)blockIdx.y * kDnBins + blockIdx.x] = (float)cap;
    }
}

struct DnApplyArgs {
    const float* in;
    float* out;
    long long n, stride, frames;
    const float* noise;     // [rows][kDnBins]
    float strength;
};

__global__ void __launch_bounds__(kDnThreads) dn_apply_kernel(const DnApplyArgs P) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const DnSmem S = dn_carve(dsm);
    float* noise = reinterpret_cast<float*>(S.rest);                                  // [kDnBins]
    float* acc = noise + kDnBins + 3;                                                 // [kDnAccLen]
    dn_fill_tables(S);
    const int group = threadIdx.x >> 6, gt = threadIdx.x & 63;
    float2* buf = S.buf + group * kDnBuf;
    const int rowi = blockIdx.y;
    const float* row = P.in + (size_t)rowi * (size_t)P.stride + kLead;
    float* orow = P.out + (size_t)rowi * (size_t)P.stride + kLead;
    for (int k = threadIdx.x; k < kDnBins; k += kDnThreads) noise[k] = P.noise[(size_t)rowi * kDnBins + k];
    for (int i = threadIdx.x; i < kDnAccLen; i += kDnThreads) acc[i] = 0.0f;
    __syncthreads();
    const long long hb = (long long)blockIdx.x * kDnHops;          // first hop owned: output samples [512 hb, 512 (hb + 13))
    const long long fa = hb - 1;                                    // frames fa .. fa + 15 touch them
    const long long acc0 = fa * kDnHop - kDnN / 2;                  // sample index of acc[0]
    for (int round = 0; round < kDnApplyFrames / kDnGroups; ++round) {
        const int fl = round * kDnGroups + group;
        const long long f = fa + fl;
        const bool valid = f >= 0 && f < P.frames;                  // uniform over the group
        float2 v[16];
        if (valid) {
            dn_load_frame(row, P.n, f, S.win, gt, v);
            fft1024_group(v, buf, S.tw, gt, group);
            gbar(group);
#pragma unroll
            for (int m = 0; m < 16; ++m) buf[padx(gt + 64 * m)] = v[m];
            gbar(group);
            // gain between the forward and the inverse untangling; the pair {k, M - k} belongs to this thread alone, so the
            // conjugated inverse-packed spectrum overwrites Z in place
            for (int k = gt; k <= kDnM / 2; k += 64) {
                float2 xk, xm;
                dn_untangle(buf, S.twh, k, xk, xm);
                // Wiener gain on |Zxx| = |X| / 1024 (pipeline.py:1504-1510)
                const float mk = sqrtf(xk.x * xk.x + xk.y * xk.y) * (1.0f / 1024.0f);
                const float mm_ = sqrtf(xm.x * xm.x + xm.y * xm.y) * (1.0f / 1024.0f);
                const float rk = noise[k] / (mk + 1e-10f), rm = noise[kDnM - k] / (mm_ + 1e-10f);
                const float gk = fminf(fmaxf(1.0f - P.strength * (rk * rk), 0.25f), 1.0f);
                const float gm = fminf(fmaxf(1.0f - P.strength * (rm * rm), 0.25f), 1.0f);
                const float2 yk = make_float2(gk * xk.x, gk * xk.y), ym = make_float2(gm * xm.x, gm * xm.y);
                // inverse packing: E' = (Y[k] + conj Y[M-k]) / 2, O' = (Y[k] - conj Y[M-k]) / 2 * conj(W^k), Z'[k] = E' + i O',
                // Z'[M-k] = conj(E') + i conj(O'); stored conjugated so that the forward transform inverts
                if (k == 0) {
                    buf[0] = make_float2(0.5f * (yk.x + ym.x), -0.5f * (yk.x - ym.x));
                } else {
                    const float2 e = make_float2(0.5f * (yk.x + ym.x), 0.5f * (yk.y - ym.y));
                    const float2 d = make_float2(0.5f * (yk.x - ym.x), 0.5f * (yk.y + ym.y));
                    const float2 o = cmulf(d, cconj(S.twh[k]));
                    buf[padx(k)] = cconj(make_float2(e.x - o.y, e.y + o.x));
                    buf[padx(kDnM - k)] = cconj(make_float2(e.x + o.y, -e.y + o.x));
                }
            }
            gbar(group);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = buf[padx(gt + 64 * q)];
            gbar(group);
            fft1024_group(v, buf, S.tw, gt, group);
        }
        // overlap-add in frame order (the reference's order): the four frames of a round overlap one another
#pragma unroll 1
        for (int g = 0; g < kDnGroups; ++g) {
            if (g == group && valid) {
                float* dst = acc + fl * kDnHop;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int j = gt + 64 * m;                      // y[2j] = Re z[j] / M, y[2j+1] = -Im conj-trick
                    const float2 w = *reinterpret_cast<const float2*>(S.win + 2 * j);
                    float2 a = *reinterpret_cast<float2*>(dst + 2 * j);
                    a.x += (v[m].x * (1.0f / kDnM)) * w.x;
                    a.y += (-v[m].y * (1.0f / kDnM)) * w.y;
                    *reinterpret_cast<float2*>(dst + 2 * j) = a;
                }
            }
            __syncthreads();
        }
    }
    const long long s0 = hb * kDnHop;
    for (int i = threadIdx.x; i < kDnHops * kDnHop; i += kDnThreads) {
        const long long t = s0 + i;
        if (t >= P.n) break;
        // overlap-added squared window over the frames that exist (scipy istft's `norm`), ascending f
        const long long p = t + kDnN / 2;                           // position in the zero-extended signal
        long long flo = (p - kDnN) / kDnHop + 1;                    // first f with 512 f + 2048 > p
        if (p < kDnN) flo = 0;
        const long long fhi = min(P.frames - 1, p / kDnHop);
        double norm = 0.0;
        for (long long f = flo; f <= fhi; ++f) {
            const double w = (double)S.win[p - f * kDnHop];
            norm += w * w;
        }
        const float v = (float)((double)acc[t - acc0] / (norm > 1e-10 ? norm : 1.0));
        orow[t] = fminf(fmaxf(v, -1.0f), 1.0f);
    }
}

int st_spectral_denoise(mm_ctx* c, const mm_geom* g, const float* in, float* out, double strength, double noise_percentile) {
    const int rows = g->tracks * g->channels;
    const long long n = g->n;
    if (n < kDnN) { set_error("apply_spectral_denoise: fewer than 2048 frames (scipy.signal.stft refuses noverlap >= nperseg)"); return 2; }
    const long long frames = (n + kDnHop - 1) / kDnHop + 1;
    const long long fpad = (frames + kDnMagFrames - 1) / kDnMagFrames * kDnMagFrames;
    float* mag;
    float* noise;
    MM_TRY(arena(c, SL_DN_MAG, (size_t)rows * kDnBins * (size_t)fpad, &mag));
    MM_TRY(arena(c, SL_DN_NOISE, (size_t)rows * kDnBins, &noise));
    const size_t tables = kDnSmemHead;
    {
        const size_t smem = tables + (size_t)kDnMagFrames * kDnBins * sizeof(float);
        MM_CUDA(cudaFuncSetAttribute(dn_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device; cheap
        DnMagArgs A;
        A.in = in; A.n = n; A.stride = g->stride; A.frames = frames; A.fpad = fpad; A.mag = mag;
        dim3 grid((unsigned)(fpad / kDnMagFrames), (unsigned)rows);
        KernelScope ks(c, "denoise_stft_mag");
        dn_mag_kernel<<<grid, kDnThreads, smem, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
    }
    {
        DnFloorArgs A;
        A.mag = mag; A.frames = frames; A.fpad = fpad; A.noise = noise;
        // numpy percentile, method 'linear': virtual index (n - 1) q with q = percentile / 100
        const double q = noise_percentile / 100.0;
        const double virt = (double)(frames - 1) * q;
        long long lo = (long long)std::floor(virt);
        double gamma = virt - (double)lo;
        lo = std::max<long long>(0, std::min(lo, frames - 1));
        const long long hi = std::min(lo + 1, frames - 1);
        if (virt >= (double)(frames - 1)) gamma = 0.0;
        A.rank[0] = lo; A.rank[1] = hi; A.gamma = gamma;
        A.rank[2] = (frames - 1) / 2; A.rank[3] = frames / 2;          // np.median: mean of the two middle elements
        dim3 grid(kDnBins, (unsigned)rows);
        KernelScope ks(c, "denoise_noise_floor_select");
        dn_floor_kernel<<<grid, kDnThreads, 0, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
    }
    {
        const size_t smem = tables + (size_t)(kDnBins + 3 + kDnAccLen) * sizeof(float);
        MM_CUDA(cudaFuncSetAttribute(dn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DnApplyArgs A;
        A.in = in; A.out = out; A.n = n; A.stride = g->stride; A.frames = frames; A.noise = noise; A.strength = (float)strength;
        const long long hops = (n + kDnHop - 1) / kDnHop;
        dim3 grid((unsigned)((hops + kDnHops - 1) / kDnHops), (unsigned)rows);
        KernelScope ks(c, "denoise_wiener_istft");
        dn_apply_kernel<<<grid, kDnThreads, smem, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace mm
