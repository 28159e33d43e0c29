// Whole-signal FFT resampling: scipy.signal.resample(x, num) for real rows, which the reference calls from resample_audio
// (backend/app/pipeline.py:920-936), the oversampled exciter (:1294-1320) and apply_reference_match for a reference track
// at another rate (:1581-1584).
//
//   x_r[t] = (1/n) Re sum_{k < m/2 + 1} g_k X[k] e^{+2 pi i k t / num},   X = DFT_n(x),  m = min(n, num),
//   g_0 = 1, g_k = 2, and for even m the unpaired bin g_{m/2} = 1 (up-sampling) or 2 (down-sampling)
//
// which is scipy's "rfft -> keep m/2 + 1 bins -> x2 / x0.5 on the unpaired bin -> irfft(n = num) * num / n" written as one
// one-sided sum.  n and num are arbitrary (7,938,000 -> 8,640,000 for a 180 s track), so both DFTs are evaluated as
// Bluestein chirp convolutions, e^{s 2 pi i j k / N} = u[j] u[k] conj(u[k - j]) with u[j] = e^{s i pi j^2 / N} (phase from
// j^2 mod 2N in 64-bit integers, float64 sincospi), over power-of-two circular lengths L >= n_in + n_out - 1.
//
// The length-L complex FFT (float32, L = 2^18 .. 2^27) is three in-place passes over the row viewed as [N1][N2][N3]:
// pass 1 and 2 transform a strided axis for a tile of adjacent columns (coalesced 32..128-byte runs, transposed into
// shared memory), pass 3 the contiguous axis; every pass is a shared-memory Stockham FFT (radix 4 + one radix-2 step for
// odd log2) of up to 4096 points per CTA, with the inter-pass twiddles W_L^e read from two L2-resident tables
// (W_L^e = lo[e & 4095] * hi[e >> 12], both rounded from float64).  The spectrum stays in the digit-scrambled order
// [k1][k2][k3]; the chirp filter's spectrum is computed by the same passes, its product rides the store of pass 3, and the
// inverse runs the passes backwards -- no transposes, no bit reversal.
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>

#include "context.h"
#include "fft.cuh"
#include "stages_internal.h"

namespace mm {

constexpr int kBfThreads = 256;
constexpr int kBfTile = 4096;           // complex points per CTA
constexpr int kBfLoBits = 12;

struct BigFft {
    int p = 0;
    long long L = 0;
    int lg[3] = {0, 0, 0};              // log2 of N1, N2, N3
    float2* twR[3] = {nullptr, nullptr, nullptr};
    float2* tlo = nullptr;
    float2* thi = nullptr;
};

struct ChirpPlan {
    long long N = 0, nin = 0, nout = 0;
    int sign = 0;
    BigFft* fft = nullptr;
    float2* FW = nullptr;               // scrambled spectrum of the chirp filter, scaled by 1 / L
};

struct BigFftCache {
    std::map<int, std::unique_ptr<BigFft>> ffts;
    std::map<std::string, ChirpPlan> chirps;
    std::vector<std::string> order;
};

static std::map<mm_ctx*, BigFftCache>& caches() {
    static std::map<mm_ctx*, BigFftCache> m;
    return m;
}

void bigfft_release(mm_ctx* c) {
    auto it = caches().find(c);
    if (it == caches().end()) return;
    for (auto& kv : it->second.ffts) {
        for (int i = 0; i < 3; ++i) cudaFree(kv.second->twR[i]);
        cudaFree(kv.second->tlo);
        cudaFree(kv.second->thi);
    }
    for (auto& kv : it->second.chirps) cudaFree(kv.second.FW);
    caches().erase(it);
}

// e^{sign i pi j^2 / N}
__device__ __forceinline__ double2 chirp(long long j, long long N, int sign) {
    const unsigned long long q = ((unsigned long long)j * (unsigned long long)j) % (unsigned long long)(2 * N);
    double s, c;
    sincospi((double)q / (double)N, &s, &c);
    return make_double2(c, sign > 0 ? s : -s);
}

__global__ void bf_table_kernel(float2* t, long long count, long long mul, long long L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s, c;
    sincospi(-2.0 * (double)(i * mul) / (double)L, &s, &c);
    t[i] = make_float2((float)c, (float)s);
}

struct BfPass {
    float2* data;
    long long pitch;        // float2 per row
    int R, lgR;             // FFT length of this pass
    int G, lgG;             // FFTs per CTA
    long long S;            // element stride of the transformed axis (1: contiguous pass)
    long long tmul;         // twiddle exponent = inner * k * tmul (0: no twiddle)
    const float2* twR;
    const float2* tlo;
    const float2* thi;
    const float2* mul;      // forward only: pointwise multiplier applied on store
    int inverse;
};

__global__ void __launch_bounds__(kBfThreads) bf_pass_kernel(const BfPass P) {
    extern __shared__ __align__(16) unsigned char bsm[];
    const int R = P.R, G = P.G, Rp = R + 1;
    float2* A = reinterpret_cast<float2*>(bsm);
    float2* B = A + (size_t)G * Rp;
    float2* tw = B + (size_t)G * Rp;
    for (int t = threadIdx.x; t < R; t += kBfThreads) tw[t] = P.twR[t];
    float2* row = P.data + (size_t)blockIdx.y * (size_t)P.pitch;
    const bool strided = P.S > 1;
    long long base, inner0 = 0;
    if (strided) {
        const long long tiles_per_outer = P.S >> P.lgG;
        const long long outer = blockIdx.x / tiles_per_outer, it = blockIdx.x % tiles_per_outer;
        base = outer * (long long)R * P.S + it * G;
        inner0 = it * G;
    } else {
        base = (long long)blockIdx.x * kBfTile;
    }
    // ---- load (inverse: conj(x) * W^e, so that the forward transform below inverts) ----
    for (int i = threadIdx.x; i < kBfTile; i += kBfThreads) {
        int g, r;
        long long off;
        if (strided) { r = i >> P.lgG; g = i & (G - 1); off = base + (long long)r * P.S + g; }
        else { g = i >> P.lgR; r = i & (R - 1); off = base + i; }
        float2 v = row[off];
        if (P.inverse) {
            v = cconj(v);
            if (P.tmul) {
                const long long e = (inner0 + g) * (long long)r * P.tmul;
                v = cmulf(v, cmulf(P.tlo[e & ((1 << kBfLoBits) - 1)], P.thi[e >> kBfLoBits]));
            }
        }
        A[g * Rp + r] = v;
    }
    __syncthreads();
    // ---- G Stockham FFTs of length R ----
    float2* src = A;
    float2* dst = B;
    const int quarter = R >> 2, lgQ = P.lgR - 2;
    int Ns = 1;
    for (; Ns * 4 <= R; Ns <<= 2) {
        const int tstep = R / (4 * Ns);
        for (int b = threadIdx.x; b < kBfTile / 4; b += kBfThreads) {
            const int g = b >> lgQ, j = b & (quarter - 1);
            const int k = j & (Ns - 1);
            const float2* s = src + g * Rp;
            float2* d = dst + g * Rp;
            float2 v0 = s[j], v1 = s[j + quarter], v2 = s[j + 2 * quarter], v3 = s[j + 3 * quarter];
            if (Ns > 1) {
                v1 = cmulf(v1, tw[k * tstep]);
                v2 = cmulf(v2, tw[2 * k * tstep]);
                v3 = cmulf(v3, tw[3 * k * tstep]);
            }
            const float2 a02 = make_float2(v0.x + v2.x, v0.y + v2.y), s02 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 a13 = make_float2(v1.x + v3.x, v1.y + v3.y), s13 = make_float2(v1.x - v3.x, v1.y - v3.y);
            const int j0 = ((j - k) << 2) + k;
            d[j0] = make_float2(a02.x + a13.x, a02.y + a13.y);
            d[j0 + Ns] = make_float2(s02.x + s13.y, s02.y - s13.x);
            d[j0 + 2 * Ns] = make_float2(a02.x - a13.x, a02.y - a13.y);
            d[j0 + 3 * Ns] = make_float2(s02.x - s13.y, s02.y + s13.x);
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    if (Ns < R) {                                   // odd log2(R): one radix-2 step with Ns = R / 2
        const int half = R >> 1, lgH = P.lgR - 1;
        for (int b = threadIdx.x; b < kBfTile / 2; b += kBfThreads) {
            const int g = b >> lgH, k = b & (half - 1);
            const float2* s = src + g * Rp;
            float2* d = dst + g * Rp;
            const float2 v0 = s[k], v1 = cmulf(s[k + half], tw[k]);
            d[k] = make_float2(v0.x + v1.x, v0.y + v1.y);
            d[k + half] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    // ---- store (forward: * W^e, * multiplier; inverse: conj) ----
    for (int i = threadIdx.x; i < kBfTile; i += kBfThreads) {
        int g, r;
        long long off;
        if (strided) { r = i >> P.lgG; g = i & (G - 1); off = base + (long long)r * P.S + g; }
        else { g = i >> P.lgR; r = i & (R - 1); off = base + i; }
        float2 v = src[g * Rp + r];
        if (P.inverse) {
            v = cconj(v);
        } else {
            if (P.tmul) {
                const long long e = (inner0 + g) * (long long)r * P.tmul;
                v = cmulf(v, cmulf(P.tlo[e & ((1 << kBfLoBits) - 1)], P.thi[e >> kBfLoBits]));
            }
            if (P.mul) v = cmulf(v, P.mul[off]);
        }
        row[off] = v;
    }
}

static int bf_run(mm_ctx* c, const BigFft* F, float2* data, long long pitch, int rows, int inverse, const float2* mul) {
    static bool attr = false;
    const size_t smem_max = (2 * (size_t)(kBfTile + 256) + 1024) * sizeof(float2);
    if (!attr) {
        MM_CUDA(cudaFuncSetAttribute(bf_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        attr = true;
    }
    const long long N1 = 1LL << F->lg[0], N2 = 1LL << F->lg[1], N3 = 1LL << F->lg[2];
    for (int step = 0; step < 3; ++step) {
        const int ps = inverse ? 2 - step : step;
        BfPass P;
        P.data = data; P.pitch = pitch; P.inverse = inverse;
        P.lgR = F->lg[ps]; P.R = 1 << P.lgR;
        P.lgG = 12 - P.lgR; P.G = 1 << P.lgG;
        P.twR = F->twR[ps]; P.tlo = F->tlo; P.thi = F->thi;
        P.mul = (!inverse && ps == 2) ? mul : nullptr;
        if (ps == 0) { P.S = N2 * N3; P.tmul = 1; }
        else if (ps == 1) { P.S = N3; P.tmul = N1; }
        else { P.S = 1; P.tmul = 0; }
        const size_t smem = (2 * (size_t)P.G * (P.R + 1) + P.R) * sizeof(float2);
        dim3 grid((unsigned)(F->L / kBfTile), (unsigned)rows);
        KernelScope ks(c, inverse ? "bigfft_pass_inverse" : "bigfft_pass_forward");
        bf_pass_kernel<<<grid, kBfThreads, smem, c->stream>>>(P);
        MM_CUDA(cudaGetLastError());
    }
    return 0;
}

static int bf_get_fft(mm_ctx* c, int p, BigFft** out) {
    auto& cache = caches()[c];
    auto it = cache.ffts.find(p);
    if (it != cache.ffts.end()) { *out = it->second.get(); return 0; }
    std::unique_ptr<BigFft> F(new BigFft);
    F->p = p;
    F->L = 1LL << p;
    // three factors, each between 2^4 and 2^9, the contiguous one the largest
    F->lg[0] = p / 3; F->lg[1] = (p - F->lg[0]) / 2; F->lg[2] = p - F->lg[0] - F->lg[1];
    for (int i = 0; i < 3; ++i) {
        const long long R = 1LL << F->lg[i];
        MM_CUDA(cudaMalloc(&F->twR[i], R * sizeof(float2)));
        bf_table_kernel<<<(unsigned)((R + 255) / 256), 256, 0, c->stream>>>(F->twR[i], R, 1, R);
    }
    const long long nlo = 1LL << kBfLoBits, nhi = F->L >> kBfLoBits;
    MM_CUDA(cudaMalloc(&F->tlo, nlo * sizeof(float2)));
    MM_CUDA(cudaMalloc(&F->thi, nhi * sizeof(float2)));
    bf_table_kernel<<<(unsigned)((nlo + 255) / 256), 256, 0, c->stream>>>(F->tlo, nlo, 1, F->L);
    bf_table_kernel<<<(unsigned)((nhi + 255) / 256), 256, 0, c->stream>>>(F->thi, nhi, nlo, F->L);
    MM_CUDA(cudaGetLastError());
    *out = F.get();
    cache.ffts[p] = std::move(F);
    return 0;
}

// filter sequence conj(u) laid out circularly: index j for 0 <= j < nout, index L - j for 1 <= j < nin
__global__ void bf_chirp_filter_kernel(float2* w, long long L, long long N, long long nin, long long nout, int sign) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    long long j = -1;
    if (i < nout) j = i;
    else if (L - i < nin) j = L - i;
    float2 v = make_float2(0.0f, 0.0f);
    if (j >= 0) {
        const double2 u = chirp(j, N, -sign);
        v = make_float2((float)u.x, (float)u.y);
    }
    w[i] = v;
}

__global__ void bf_scale_kernel(float2* w, long long L, float s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) w[i] = make_float2(w[i].x * s, w[i].y * s);
}

static int bf_get_chirp(mm_ctx* c, long long N, long long nin, long long nout, int sign, const ChirpPlan** out) {
    auto& cache = caches()[c];
    char key[96];
    snprintf(key, sizeof key, "%lld/%lld/%lld/%d", N, nin, nout, sign);
    auto it = cache.chirps.find(key);
    if (it != cache.chirps.end()) { *out = &it->second; return 0; }
    if (cache.order.size() >= 6) {                  // a handful of (n, num) pairs is all a service sees; bound the filters kept
        MM_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(cache.chirps[cache.order.front()].FW);
        cache.chirps.erase(cache.order.front());
        cache.order.erase(cache.order.begin());
    }
    int p = 18;                                     // smallest length whose passes tile as laid out below (2 MB per row)
    while ((1LL << p) < nin + nout - 1) ++p;
    if (p > 27) { set_error("fft resample: %lld + %lld points exceed the 2^27-point transform", nin, nout); return 2; }
    ChirpPlan P;
    P.N = N; P.nin = nin; P.nout = nout; P.sign = sign;
    MM_TRY(bf_get_fft(c, p, &P.fft));
    const long long L = P.fft->L;
    MM_CUDA(cudaMalloc(&P.FW, L * sizeof(float2)));
    bf_chirp_filter_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(P.FW, L, N, nin, nout, sign);
    MM_CUDA(cudaGetLastError());
    MM_TRY(bf_run(c, P.fft, P.FW, L, 1, 0, nullptr));
    bf_scale_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(P.FW, L, (float)(1.0 / (double)L));
    MM_CUDA(cudaGetLastError());
    cache.chirps[key] = P;
    cache.order.push_back(key);
    *out = &cache.chirps[key];
    return 0;
}

struct RsArgs {
    const float* in;
    float* out;
    long long n, num, m2, in_stride, out_stride, pitch, L1, L2;
    int up, m_even;
    float2* work;
};

// A1[j] = x[j] u1[j]   (u1[j] = e^{-i pi j^2 / n}), zero up to L1
__global__ void rs_pre_kernel(const RsArgs P) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.L1) return;
    float2 v = make_float2(0.0f, 0.0f);
    if (j < P.n) {
        const double x = (double)P.in[(size_t)blockIdx.y * (size_t)P.in_stride + kLead + j];
        const double2 u = chirp(j, P.n, -1);
        v = make_float2((float)(x * u.x), (float)(x * u.y));
    }
    P.work[(size_t)blockIdx.y * (size_t)P.pitch + j] = v;
}

// X[k] = u1[k] conv1[k];  A2[k] = (g_k / n) X[k] u2[k]   (u2[k] = e^{+i pi k^2 / num}), zero up to L2
__global__ void rs_mid_kernel(const RsArgs P) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.L2) return;
    float2* w = P.work + (size_t)blockIdx.y * (size_t)P.pitch + k;
    float2 v = make_float2(0.0f, 0.0f);
    if (k < P.m2) {
        const float2 cv = *w;
        const double2 u1 = chirp(k, P.n, -1), u2 = chirp(k, P.num, +1);
        const double xr = (double)cv.x * u1.x - (double)cv.y * u1.y, xi = (double)cv.x * u1.y + (double)cv.y * u1.x;
        double g = k == 0 ? 1.0 : 2.0;
        if (P.m_even && k == P.m2 - 1) g = P.up ? 1.0 : 2.0;
        g /= (double)P.n;
        v = make_float2((float)(g * (xr * u2.x - xi * u2.y)), (float)(g * (xr * u2.y + xi * u2.x)));
    }
    *w = v;
}

// out[t] = Re(u2[t] conv2[t])
__global__ void rs_post_kernel(const RsArgs P) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.num) return;
    const float2 cv = P.work[(size_t)blockIdx.y * (size_t)P.pitch + t];
    const double2 u = chirp(t, P.num, +1);
    P.out[(size_t)blockIdx.y * (size_t)P.out_stride + kLead + t] = (float)((double)cv.x * u.x - (double)cv.y * u.y);
}

int st_fft_resample(mm_ctx* c, const mm_geom* gi, const float* in, const mm_geom* go, float* out) {
    const long long n = gi->n, num = go->n;
    const int rows = gi->tracks * gi->channels;
    if (rows != go->tracks * go->channels) { set_error("fft resample: input and output batches differ in rows"); return 2; }
    if (n < 1 || num < 1) { set_error("fft resample: empty signal"); return 2; }
    if (n == num) { set_error("fft resample: equal lengths (copy instead)"); return 2; }
    const long long m = std::min(n, num), m2 = m / 2 + 1;
    const ChirpPlan *F1, *F2;
    MM_TRY(bf_get_chirp(c, n, n, m2, -1, &F1));
    MM_TRY(bf_get_chirp(c, num, m2, num, +1, &F2));
    RsArgs A;
    A.n = n; A.num = num; A.m2 = m2; A.in_stride = gi->stride; A.out_stride = go->stride;
    A.L1 = F1->fft->L; A.L2 = F2->fft->L; A.pitch = std::max(A.L1, A.L2);
    A.up = num > n; A.m_even = (m % 2 == 0);
    // rows per sub-batch: keep the complex work area near 4 GB
    const int chunk = (int)std::max<long long>(1, std::min<long long>(rows, (4LL << 30) / (A.pitch * (long long)sizeof(float2))));
    MM_TRY(arena(c, SL_BIGFFT, (size_t)chunk * (size_t)A.pitch, &A.work));
    for (int r0 = 0; r0 < rows; r0 += chunk) {
        const int nr = std::min(chunk, rows - r0);
        A.in = in + (size_t)r0 * (size_t)gi->stride;
        A.out = out + (size_t)r0 * (size_t)go->stride;
        {
            KernelScope ks(c, "resample_chirp_pre");
            rs_pre_kernel<<<dim3((unsigned)((A.L1 + 255) / 256), (unsigned)nr), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
        MM_TRY(bf_run(c, F1->fft, A.work, A.pitch, nr, 0, F1->FW));
        MM_TRY(bf_run(c, F1->fft, A.work, A.pitch, nr, 1, nullptr));
        {
            KernelScope ks(c, "resample_chirp_mid");
            rs_mid_kernel<<<dim3((unsigned)((A.L2 + 255) / 256), (unsigned)nr), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
        MM_TRY(bf_run(c, F2->fft, A.work, A.pitch, nr, 0, F2->FW));
        MM_TRY(bf_run(c, F2->fft, A.work, A.pitch, nr, 1, nullptr));
        {
            KernelScope ks(c, "resample_chirp_post");
            rs_post_kernel<<<dim3((unsigned)((num + 255) / 256), (unsigned)nr), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

}  // namespace mm
