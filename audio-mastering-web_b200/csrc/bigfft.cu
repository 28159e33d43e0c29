// Whole-signal FFT resampling: scipy.signal.resample(x, num) for real rows, which the reference calls from resample_audio
// (backend/app/pipeline.py:920-936), the oversampled exciter (:1294-1320) and apply_reference_match for a reference track
// at another rate (:1581-1584).
//
//   x_r[t] = (1/n) sum_{|k| <= m/2} w_k X[k mod n] e^{+2 pi i k t / num},   X = DFT_n(x),  m = min(n, num)
//
// which is scipy's "fft -> keep the m lowest bins -> split / fold the unpaired bin -> ifft(n = num) * num / n" written as one
// sum (its rfft path for real input is the same thing); a track's two channels go through as ONE complex signal L + i R.  n and num are arbitrary (7,938,000 -> 8,640,000 for a 180 s track), so both DFTs are evaluated as
// Bluestein chirp convolutions, e^{s 2 pi i j k / N} = u[j] u[k] conj(u[k - j]) with u[j] = e^{s i pi j^2 / N} (phase from
// j^2 mod 2N in 64-bit integers, reduced in float64), over power-of-two circular lengths L >= n_in + n_out - 1.
//
// The length-L complex FFT (float32, L = 2^18 .. 2^27) is three (four above 2^24) in-place passes over the row viewed as
// [N1][N2][N3]([N4]) with axes of 64, 128 or 256 points: the strided axes are transformed for a tile of adjacent columns
// (coalesced 128..512-byte runs), the last pass the contiguous axis.  A CTA owns 4096 points, 16 per thread: a radix-16
// (radix-8) step in registers straight from global memory, one exchange through padded shared memory, a second in-register
// step straight back to global memory; the inter-pass twiddles W_L^e come from two L2-resident tables
// (W_L^e = lo[e & 4095] * hi[e >> 12], both rounded from float64).  The spectrum stays in the digit-scrambled order
// [k1][k2][k3]; the chirp filter's spectrum is computed by the same passes, its product rides the load of the inverse's
// first pass, and the inverse runs the passes backwards -- no transposes, no bit reversal.
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
    int np = 3;                         // passes (axes)
    int lg[4] = {0, 0, 0, 0};           // log2 of the axis lengths, outermost first
    float2* twR[4] = {nullptr, nullptr, nullptr, nullptr};
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

static BigFftCache& cache_of(mm_ctx* c) {
    if (!c->bigfft) c->bigfft = new BigFftCache;
    return *static_cast<BigFftCache*>(c->bigfft);
}

void bigfft_release(mm_ctx* c) {
    if (!c->bigfft) return;
    BigFftCache* cache = static_cast<BigFftCache*>(c->bigfft);
    for (auto& kv : cache->ffts) {
        for (int i = 0; i < 4; ++i) cudaFree(kv.second->twR[i]);
        cudaFree(kv.second->tlo);
        cudaFree(kv.second->thi);
    }
    for (auto& kv : cache->chirps) cudaFree(kv.second.FW);
    delete cache;
    c->bigfft = nullptr;
}

// e^{sign i pi j^2 / N} to float32 accuracy: j^2 mod 2N exactly (Barrett reduction with m = floor(2^64 / 2N)), the phase
// reduced in float64 to [-1/4, 1/4] half-turns around the nearest quadrant, float32 sincospi there, quadrant rotation
struct ChirpMod {
    unsigned long long twoN, m;
    double invN;
};
static ChirpMod chirp_mod(long long N) {
    ChirpMod c;
    c.twoN = 2ULL * (unsigned long long)N;
    c.m = (unsigned long long)((((unsigned __int128)1) << 64) / c.twoN);
    c.invN = 1.0 / (double)N;
    return c;
}
__device__ __forceinline__ float2 chirp(unsigned long long j, const ChirpMod& M, int sign) {
    const unsigned long long q = j * j;
    unsigned long long rem = q - __umul64hi(q, M.m) * M.twoN;
    while (rem >= M.twoN) rem -= M.twoN;
    const double t = (double)(long long)rem * M.invN;          // half-turns in [0, 2)
    const double kq = rint(t * 2.0);
    const float a = (float)(t - 0.5 * kq);
    float s, c;
    sincospif(a, &s, &c);
    const int k = (int)kq & 3;
    const float cr = k == 0 ? c : k == 1 ? -s : k == 2 ? -c : s;
    const float sr = k == 0 ? s : k == 1 ? c : k == 2 ? -s : -c;
    return make_float2(cr, sign > 0 ? sr : -sr);
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
    long long S;            // element stride of the transformed axis (1: contiguous pass)
    unsigned tmul;          // twiddle exponent = inner * k * tmul (0: no twiddle)
    const float2* twR;      // [R] e^{-2 pi i t / R}
    const float2* tlo;
    const float2* thi;
    const float2* mul;      // pointwise multiplier applied on load (the inverse's first pass)
    int inverse;
};

// W_L^{c (j + q STEP)}, q = 0 .. N-1, from four table reads: base W^{c j}, step D = W^{c STEP}, D^2, D^4, D^8 by squaring and
// every power as a product of at most five factors (error <= 5 float32 roundings, instead of 2 N dependent L2 reads)
template <int N> __device__ __forceinline__ void twiddle_run(const BfPass& P, unsigned c, unsigned j, unsigned step, float2* w) {
    const unsigned lomask = (1u << kBfLoBits) - 1u;
    const unsigned e0 = c * j, e1 = c * step;
    w[0] = cmulf(P.tlo[e0 & lomask], P.thi[e0 >> kBfLoBits]);
    float2 d = cmulf(P.tlo[e1 & lomask], P.thi[e1 >> kBfLoBits]);
#pragma unroll
    for (int h = 1; h < N; h <<= 1) {
#pragma unroll
        for (int q = 0; q < h; ++q) w[h + q] = cmulf(w[q], d);
        d = cmulf(d, d);
    }
}

// One pass = G = 4096 / R transforms of length R = RA * RB per CTA, 16 points per thread: a radix-RA step straight from
// global memory into registers, one exchange through shared memory, a radix-RB step straight back to global memory.
template <int RA, int RB, bool STRIDED> __global__ void __launch_bounds__(kBfThreads, 3) bf_pass_kernel(const BfPass P) {
    constexpr int R = RA * RB, G = kBfTile / R, Rp = R + RB + 1;          // index i of a transform lives at i + i / RA
    constexpr int NA = 16 / RA, NB = 16 / RB;
    __shared__ float2 Y[G * Rp];
    __shared__ float2 tw[R];
    for (int t = threadIdx.x; t < R; t += kBfThreads) tw[t] = P.twR[t];
    float2* row = P.data + (size_t)blockIdx.y * (size_t)P.pitch;
    const unsigned S = (unsigned)P.S;                 // offsets inside a row fit 32 bits (L <= 2^27)
    unsigned base, inner0 = 0;
    if (STRIDED) {
        const unsigned tiles_per_outer = S / G;
        const unsigned outer = blockIdx.x / tiles_per_outer, it = blockIdx.x % tiles_per_outer;
        base = outer * (unsigned)R * S + it * G;
        inner0 = it * G;
    } else {
        base = blockIdx.x * (unsigned)kBfTile;
    }
    float2 v[16];
    // ---- radix RA (Stockham step with Ns = 1) ----
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const int b = threadIdx.x + kBfThreads * i;
        const int g = STRIDED ? b % G : b / RB, j = STRIDED ? b / G : b % RB;
#pragma unroll
        for (int q = 0; q < RA; ++q) {
            const int r = j + q * RB;
            const unsigned off = STRIDED ? base + (unsigned)r * S + g : base + g * R + r;
            float2 x = row[off];
            if (P.mul) x = cmulf(x, P.mul[off]);    // inverse, first pass: the chirp filter's spectrum
            v[i * RA + q] = x;
        }
        if (P.inverse) {                            // conj(x) W^e: the forward transform below then inverts
            if (STRIDED && P.tmul) {
                float2 w[RA];
                twiddle_run<RA>(P, (inner0 + g) * P.tmul, j, RB, w);
#pragma unroll
                for (int q = 0; q < RA; ++q) v[i * RA + q] = cmulf(cconj(v[i * RA + q]), w[q]);
            } else {
#pragma unroll
                for (int q = 0; q < RA; ++q) v[i * RA + q] = cconj(v[i * RA + q]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const int b = threadIdx.x + kBfThreads * i;
        const int g = STRIDED ? b % G : b / RB, j = STRIDED ? b / G : b % RB;
        dft_reg<RA>(v + i * RA);
#pragma unroll
        for (int k = 0; k < RA; ++k) Y[g * Rp + j * (RA + 1) + k] = v[i * RA + dft_pos<RA>(k)];
    }
    __syncthreads();
    // ---- radix RB (Ns = RA): inputs j' + q RA, twiddle W_R^{j' q}, outputs j' + q' RA ----
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = threadIdx.x + kBfThreads * i;
        const int g = STRIDED ? b % G : b / RA, j = STRIDED ? b / G : b % RA;
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            float2 x = Y[g * Rp + j + q * (RA + 1)];
            if (q) x = cmulf(x, tw[j * q]);
            v[i * RB + q] = x;
        }
        dft_reg<RB>(v + i * RB);
        float2 w[RB];
        const bool twid = STRIDED && !P.inverse && P.tmul;
        if (twid) twiddle_run<RB>(P, (inner0 + g) * P.tmul, j, RA, w);
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            const int r = j + k * RA;
            const unsigned off = STRIDED ? base + (unsigned)r * S + g : base + g * R + r;
            float2 x = v[i * RB + dft_pos<RB>(k)];
            if (P.inverse) x = cconj(x);
            else if (twid) x = cmulf(x, w[k]);
            row[off] = x;
        }
    }
}

template <bool STRIDED> static void bf_launch(int lgR, dim3 grid, cudaStream_t st, const BfPass& P) {
    if (lgR == 6) bf_pass_kernel<8, 8, STRIDED><<<grid, kBfThreads, 0, st>>>(P);
    else if (lgR == 7) bf_pass_kernel<16, 8, STRIDED><<<grid, kBfThreads, 0, st>>>(P);
    else bf_pass_kernel<16, 16, STRIDED><<<grid, kBfThreads, 0, st>>>(P);
}

static int bf_run(mm_ctx* c, const BigFft* F, float2* data, long long pitch, int rows, int inverse, const float2* mul) {
    for (int step = 0; step < F->np; ++step) {
        const int ps = inverse ? F->np - 1 - step : step;
        BfPass P;
        P.data = data; P.pitch = pitch; P.inverse = inverse;
        P.twR = F->twR[ps]; P.tlo = F->tlo; P.thi = F->thi;
        P.mul = (inverse && ps == F->np - 1) ? mul : nullptr;
        int below = 0, above = 0;                   // log2 of the product of the axes after / before this one
        for (int m = ps + 1; m < F->np; ++m) below += F->lg[m];
        for (int m = 0; m < ps; ++m) above += F->lg[m];
        P.S = 1LL << below;
        P.tmul = ps == F->np - 1 ? 0u : 1u << above;
        dim3 grid((unsigned)(F->L / kBfTile), (unsigned)rows);
        KernelScope ks(c, inverse ? "bigfft_pass_inverse" : "bigfft_pass_forward");
        if (P.S > 1) bf_launch<true>(F->lg[ps], grid, c->stream, P);
        else bf_launch<false>(F->lg[ps], grid, c->stream, P);
        MM_CUDA(cudaGetLastError());
    }
    return 0;
}

static int bf_get_fft(mm_ctx* c, int p, BigFft** out) {
    auto& cache = cache_of(c);
    auto it = cache.ffts.find(p);
    if (it != cache.ffts.end()) { *out = it->second.get(); return 0; }
    std::unique_ptr<BigFft> F(new BigFft);
    F->p = p;
    F->L = 1LL << p;
    // three (L <= 2^24) or four axes of 64, 128 or 256 points, the longer ones last
    F->np = p <= 24 ? 3 : 4;
    for (int i = 0; i < F->np; ++i) F->lg[i] = p / F->np + (i >= F->np - p % F->np ? 1 : 0);
    for (int i = 0; i < F->np; ++i) {
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
__global__ void bf_chirp_filter_kernel(float2* w, long long L, const ChirpMod M, long long nin, long long nout, int sign) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    long long j = -1;
    if (i < nout) j = i;
    else if (L - i < nin) j = L - i;
    float2 v = make_float2(0.0f, 0.0f);
    if (j >= 0) v = chirp((unsigned long long)j, M, -sign);
    w[i] = v;
}

__global__ void bf_scale_kernel(float2* w, long long L, float s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) w[i] = make_float2(w[i].x * s, w[i].y * s);
}

static int bf_get_chirp(mm_ctx* c, long long N, long long nin, long long nout, int sign, const ChirpPlan** out) {
    auto& cache = cache_of(c);
    char key[96];
    snprintf(key, sizeof key, "%lld/%lld/%lld/%d", N, nin, nout, sign);
    auto it = cache.chirps.find(key);
    if (it != cache.chirps.end()) {
        // least recently used goes first: a plan handed out a moment ago (the forward half of a resample) must not be the one
        // the next miss evicts
        auto pos = std::find(cache.order.begin(), cache.order.end(), std::string(key));
        if (pos != cache.order.end()) { cache.order.erase(pos); cache.order.push_back(key); }
        *out = &it->second;
        return 0;
    }
    if (cache.order.size() >= 6) {                  // a handful of (n, num) pairs is all a service sees; bound the filters kept
        MM_CUDA(cudaStreamSynchronize(c->stream));
        c->workspace_bytes -= (int64_t)(cache.chirps[cache.order.front()].fft->L * sizeof(float2));
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
    c->workspace_bytes += (int64_t)(L * sizeof(float2));
    bf_chirp_filter_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(P.FW, L, chirp_mod(N), nin, nout, sign);
    MM_CUDA(cudaGetLastError());
    MM_TRY(bf_run(c, P.fft, P.FW, L, 1, 0, nullptr));
    bf_scale_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(P.FW, L, (float)(1.0 / (double)L));
    MM_CUDA(cudaGetLastError());
    cache.chirps[key] = P;
    cache.order.push_back(key);
    *out = &cache.chirps[key];
    return 0;
}

// Both channels of a track ride ONE complex signal z = L + i R (resampling is linear with real coefficients, so the real
// and imaginary parts of the resampled z are the resampled channels; mono: R = 0).  For a complex signal scipy's rule is
// the two-sided one:
//   y[t] = (1/n) sum_{k = -Kn}^{Kn} w_k X[k mod n] e^{2 pi i k t / num},   Kn = m / 2 (m even) or (m - 1) / 2 (m odd),
//   w = 1, except w_{+-m/2} = 1/2 when up-sampling an even n (the unpaired bin split in two); when down-sampling to an
//   even num both +-num/2 terms fold onto the new Nyquist bin with weight 1 -- which is the same sum.
// The window of bins starts at -Kn, so both chirps carry a linear phase: the forward one e^{-i pi (j^2 - 2 Kn j) / n}
// yields X'[k'] = X[k' - Kn] for k' = 0 .. nk - 1 directly, the inverse one e^{+i pi (t^2 - 2 Kn t) / num} undoes the shift.
struct RsArgs {
    const float* in;
    float* out;
    long long n, num, nk, in_stride, out_stride, pitch, L1, L2;
    int channels;
    float w_end;            // weight of the first and the last bin of the window
    float2* work;
    ChirpMod Mn, Mnum;
    unsigned long long lin_n, lin_num;   // (-2 Kn) mod 2n, (-2 Kn) mod 2 num
};

// e^{sign i pi (j^2 + lin j) / N},  lin given mod 2N
__device__ __forceinline__ float2 chirp_lin(unsigned long long j, unsigned long long lin, const ChirpMod& M, int sign) {
    const unsigned long long q = j * (j + lin);                        // < 2^27 * 2^29
    unsigned long long rem = q - __umul64hi(q, M.m) * M.twoN;
    while (rem >= M.twoN) rem -= M.twoN;
    const double t = (double)(long long)rem * M.invN;
    const double kq = rint(t * 2.0);
    const float a = (float)(t - 0.5 * kq);
    float s, c;
    sincospif(a, &s, &c);
    const int k = (int)kq & 3;
    const float cr = k == 0 ? c : k == 1 ? -s : k == 2 ? -c : s;
    const float sr = k == 0 ? s : k == 1 ? c : k == 2 ? -s : -c;
    return make_float2(cr, sign > 0 ? sr : -sr);
}

// A1[j] = z[j] e^{-i pi (j^2 - 2 Kn j) / n}, zero up to L1
__global__ void rs_pre_kernel(const RsArgs P) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.L1) return;
    float2 v = make_float2(0.0f, 0.0f);
    if (j < P.n) {
        const float* r0 = P.in + (size_t)(blockIdx.y * P.channels) * (size_t)P.in_stride + kLead;
        const float2 z = make_float2(r0[j], P.channels > 1 ? r0[P.in_stride + j] : 0.0f);
        v = cmulf(z, chirp_lin((unsigned long long)j, P.lin_n, P.Mn, -1));
    }
    P.work[(size_t)blockIdx.y * (size_t)P.pitch + j] = v;
}

// X'[k'] = u1[k'] conv1[k'];  A2[k'] = (w / n) X'[k'] v[k']   (u1 = e^{-i pi k'^2 / n}, v = e^{+i pi k'^2 / num}), zero up to L2
__global__ void rs_mid_kernel(const RsArgs P) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.L2) return;
    float2* w = P.work + (size_t)blockIdx.y * (size_t)P.pitch + k;
    float2 v = make_float2(0.0f, 0.0f);
    if (k < P.nk) {
        const float2 cv = *w;
        const float2 u1 = chirp((unsigned long long)k, P.Mn, -1), u2 = chirp((unsigned long long)k, P.Mnum, +1);
        const float wk = (k == 0 || k == P.nk - 1) ? P.w_end : 1.0f;
        const float gf = (float)((double)wk / (double)P.n);
        const float2 y = cmulf(cmulf(cv, u1), u2);
        v = make_float2(gf * y.x, gf * y.y);
    }
    *w = v;
}

// (L, R)[t] = e^{+i pi (t^2 - 2 Kn t) / num} conv2[t]
__global__ void rs_post_kernel(const RsArgs P) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.num) return;
    const float2 cv = P.work[(size_t)blockIdx.y * (size_t)P.pitch + t];
    const float2 y = cmulf(cv, chirp_lin((unsigned long long)t, P.lin_num, P.Mnum, +1));
    float* r0 = P.out + (size_t)(blockIdx.y * P.channels) * (size_t)P.out_stride + kLead;
    r0[t] = y.x;
    if (P.channels > 1) r0[P.out_stride + t] = y.y;
}

int st_fft_resample(mm_ctx* c, const mm_geom* gi, const float* in, const mm_geom* go, float* out) {
    const long long n = gi->n, num = go->n;
    if (gi->tracks != go->tracks || gi->channels != go->channels) { set_error("fft resample: input and output batches differ in tracks / channels"); return 2; }
    if (n < 1 || num < 1) { set_error("fft resample: empty signal"); return 2; }
    if (n == num) { set_error("fft resample: equal lengths (copy instead)"); return 2; }
    const long long m = std::min(n, num);
    const long long Kn = (m % 2 == 0) ? m / 2 : (m - 1) / 2;
    const long long nk = 2 * Kn + 1;
    const ChirpPlan *F1, *F2;
    MM_TRY(bf_get_chirp(c, n, n, nk, -1, &F1));
    MM_TRY(bf_get_chirp(c, num, nk, num, +1, &F2));
    RsArgs A;
    A.n = n; A.num = num; A.nk = nk; A.in_stride = gi->stride; A.out_stride = go->stride; A.channels = gi->channels;
    A.L1 = F1->fft->L; A.L2 = F2->fft->L; A.pitch = std::max(A.L1, A.L2);
    A.w_end = (m % 2 == 0 && num > n) ? 0.5f : 1.0f;
    A.Mn = chirp_mod(n); A.Mnum = chirp_mod(num);
    A.lin_n = (unsigned long long)((2 * n - (2 * Kn) % (2 * n)) % (2 * n));
    A.lin_num = (unsigned long long)((2 * num - (2 * Kn) % (2 * num)) % (2 * num));
    const int tracks = gi->tracks;
    // tracks per sub-batch: keep the complex work area near 4 GB
    const int chunk = (int)std::max<long long>(1, std::min<long long>(tracks, (4LL << 30) / (A.pitch * (long long)sizeof(float2))));
    MM_TRY(arena(c, SL_BIGFFT, (size_t)chunk * (size_t)A.pitch, &A.work));
    for (int t0 = 0; t0 < tracks; t0 += chunk) {
        const int nr = std::min(chunk, tracks - t0);
        A.in = in + (size_t)(t0 * gi->channels) * (size_t)gi->stride;
        A.out = out + (size_t)(t0 * go->channels) * (size_t)go->stride;
        {
            KernelScope ks(c, "resample_chirp_pre");
            rs_pre_kernel<<<dim3((unsigned)((A.L1 + 255) / 256), (unsigned)nr), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
        MM_TRY(bf_run(c, F1->fft, A.work, A.pitch, nr, 0, nullptr));
        MM_TRY(bf_run(c, F1->fft, A.work, A.pitch, nr, 1, F1->FW));
        {
            KernelScope ks(c, "resample_chirp_mid");
            rs_mid_kernel<<<dim3((unsigned)((A.L2 + 255) / 256), (unsigned)nr), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
        MM_TRY(bf_run(c, F2->fft, A.work, A.pitch, nr, 0, nullptr));
        MM_TRY(bf_run(c, F2->fft, A.work, A.pitch, nr, 1, F2->FW));
        {
            KernelScope ks(c, "resample_chirp_post");
            rs_post_kernel<<<dim3((unsigned)((num + 255) / 256), (unsigned)nr), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
    }
    return 0;
}


// ---- long FIR as one circular convolution per row -----------------------------------------------------------------------
// out[i] = sum_k taps[k] x[i + (K-1)/2 - k]  (scipy.signal.fftconvolve(x, taps, mode="same"): the 4096-tap linear-phase
// target curve, pipeline.py:220-235, and the 8192-tap reference match, :1600-1606).  The two channels of a track ride one
// complex transform -- z = L + i R convolved with real taps is (L * h) + i (R * h), no untangling -- of length
// L >= n + K - 1 (a power of two), so a 180 s stereo track costs two 2^24-point FFTs instead of 2 x 4096 MACs per sample.
struct CvArgs {
    const float* in;
    float* out;
    long long n, stride, L;
    int channels, K, center, clip;
    float2* work;
    const float* taps;
};

__global__ void cv_taps_kernel(float2* h, long long L, const float* taps, int K) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) h[i] = make_float2(i < K ? taps[i] : 0.0f, 0.0f);
}

__global__ void cv_pre_kernel(const CvArgs P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.L) return;
    float2 v = make_float2(0.0f, 0.0f);
    if (i < P.n) {
        const float* r0 = P.in + (size_t)(blockIdx.y * P.channels) * (size_t)P.stride + kLead;
        v.x = r0[i];
        if (P.channels > 1) v.y = r0[P.stride + i];
    }
    P.work[(size_t)blockIdx.y * (size_t)P.L + i] = v;
}

__global__ void cv_post_kernel(const CvArgs P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    float2 v = P.work[(size_t)blockIdx.y * (size_t)P.L + i + P.center];
    if (P.clip) { v.x = fminf(fmaxf(v.x, -1.f), 1.f); v.y = fminf(fmaxf(v.y, -1.f), 1.f); }
    float* r0 = P.out + (size_t)(blockIdx.y * P.channels) * (size_t)P.stride + kLead;
    r0[i] = v.x;
    if (P.channels > 1) r0[P.stride + i] = v.y;
}

bool fft_convolve_fits(const mm_geom* g, int K) { return g->n + K - 1 <= (1LL << 27); }

int st_fft_convolve_same(mm_ctx* c, const mm_geom* g, const float* in, float* out, const float* taps_dev, int K, int clip) {
    int p = 18;
    while ((1LL << p) < g->n + K - 1) ++p;
    if (p > 27) { set_error("FFT convolution: %lld + %d points exceed the 2^27-point transform", (long long)g->n, K); return 2; }
    BigFft* F;
    MM_TRY(bf_get_fft(c, p, &F));
    const long long L = F->L;
    float2* H;
    MM_TRY(arena(c, SL_BIGFFT_H, (size_t)L, &H));
    cv_taps_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(H, L, taps_dev, K);
    MM_CUDA(cudaGetLastError());
    MM_TRY(bf_run(c, F, H, L, 1, 0, nullptr));
    bf_scale_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(H, L, (float)(1.0 / (double)L));
    MM_CUDA(cudaGetLastError());
    CvArgs A;
    A.n = g->n; A.stride = g->stride; A.L = L; A.channels = g->channels; A.K = K; A.center = (K - 1) / 2; A.clip = clip;
    A.taps = taps_dev;
    const int tracks = g->tracks;
    const int chunk = (int)std::max<long long>(1, std::min<long long>(tracks, (4LL << 30) / (L * (long long)sizeof(float2))));
    MM_TRY(arena(c, SL_BIGFFT, (size_t)chunk * (size_t)L, &A.work));
    for (int t0 = 0; t0 < tracks; t0 += chunk) {
        const int nt = std::min(chunk, tracks - t0);
        A.in = in + (size_t)(t0 * g->channels) * (size_t)g->stride;
        A.out = out + (size_t)(t0 * g->channels) * (size_t)g->stride;
        {
            KernelScope ks(c, "fftconv_pack");
            cv_pre_kernel<<<dim3((unsigned)((L + 255) / 256), (unsigned)nt), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
        MM_TRY(bf_run(c, F, A.work, L, nt, 0, nullptr));
        MM_TRY(bf_run(c, F, A.work, L, nt, 1, H));
        {
            KernelScope ks(c, "fftconv_unpack");
            cv_post_kernel<<<dim3((unsigned)((g->n + 255) / 256), (unsigned)nt), 256, 0, c->stream>>>(A);
            MM_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

}  // namespace mm
