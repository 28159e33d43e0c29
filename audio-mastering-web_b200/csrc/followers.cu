// Stages built on a pair of attack/release envelope followers (_envelope_follower_core, backend/app/pipeline.py:495-518):
//   apply_transient_designer        (pipeline.py:1736-1768)   per channel: fast (0.5 / 5 ms) and slow (5 / 100 ms) follower of |x|
//   apply_maximizer_transient_aware (pipeline.py:521-545)     per track: fast (0.5 / 2 ms) and slow (10 / 40 ms) follower of mean |x|
// plus two small first-wave neighbours that needed no new kernel class:
//   apply_high_freq_trim            (pipeline.py:1705-1733)   zero-phase low-pass + weighted recombination + clip
//   Haas "stereoize" branch of apply_stereo_imager (pipeline.py:1388-1398)
//
// The follower has no associative operator; as in deesser.cu a unit (row or track) is cut into chunks that start
// `halo` samples early from the state |v|: one step contracts the distance between two states by at least the
// release coefficient, so after halo = 17.5 / -ln(slowest coefficient) samples the start-up error is e^-17.5 = 2.5e-8
// of the signal peak.  Chunk 0 starts at sample 0 with the reference's own initial state and is exact.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "context.h"
#include "pointwise.cuh"
#include "stages_internal.h"

namespace mm {

enum { FOL_TRANSIENT = 0, FOL_MAXIMIZER_TA = 1 };

struct FolArgs {
    const float* x;
    float* out;
    long long n, stride;
    int units;             // rows (transient designer) or tracks (maximizer)
    long long chunk, halo; // multiples of 32 samples
    int nchunks;           // per unit
    double f_atk, f_1matk, f_rel, f_1mrel;   // fast follower: float64 coefficients and products, float32 state --
    double s_atk, s_1matk, s_rel, s_1mrel;   // slow follower     what numba compiles pipeline.py:495-507 to
    float attack_gain, sustain_gain;         // transient designer
    float sensitivity;                       // maximizer
    DynParams dyn;                           // maximizer line (max_k, max_c, max_top)
};

__device__ __forceinline__ float fol_step_branch(float e, float v, double atk, double omatk, double rel, double omrel) {
    const double ed = (double)e, vd = (double)v;
    return v > e ? (float)(atk * ed + omatk * vd) : (float)(rel * ed + omrel * vd);
}

constexpr int kFolDepth = 4;
constexpr int kFolThreads = 32;

template <int MODE, int C>
__global__ void __launch_bounds__(kFolThreads) dual_follower_kernel(const FolArgs P) {
    __shared__ __align__(128) float ring[C][kFolDepth][kFolThreads][32];
    const int lane = threadIdx.x;
    const long long gid = (long long)blockIdx.x * kFolThreads + lane;
    const long long total = (long long)P.units * P.nchunks;
    const bool active = gid < total;
    const int unit = active ? (int)(gid / P.nchunks) : 0;
    const int chunk = active ? (int)(gid % P.nchunks) : 0;
    const size_t r0 = (size_t)(unit * C) * (size_t)P.stride + kLead;
    const long long live0 = (long long)chunk * P.chunk;
    const long long live1 = active ? min(live0 + P.chunk, P.n) : live0;
    const long long pos0 = live0 - P.halo;                     // position of line 0; negative: the chunk starts at sample 0
    const long long start = max(pos0, 0LL);
    const int halo_lines = (int)(P.halo / 32);
    const int nlines = halo_lines + (int)(P.chunk / 32);       // uniform across the warp (P.halo, P.chunk: multiples of 32)
    const int my_lo = pos0 < 0 ? (int)((-pos0) / 32) : 0;      // this chunk's lines with samples of [0, n): [my_lo, my_hi)
    const int my_hi = (active && live1 > pos0) ? (int)((live1 - pos0 + 31) / 32) : 0;
    const int sx = lane & 7;
    // cooperative fetch (see deesser.cu): eight lanes copy the eight 16-byte units of one chunk's line -- a warp-level cp.async
    // touches 4 lines completely instead of 32 partially
    const int fu = lane & 7;
    long long p_off[8];
    int p_lo[8], p_hi[8], p_dst[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int cl = 4 * q + (lane >> 3);
        p_off[q] = __shfl_sync(0xffffffffu, (long long)r0 + pos0, cl) + 4 * fu;
        p_lo[q] = __shfl_sync(0xffffffffu, my_lo, cl);
        p_hi[q] = __shfl_sync(0xffffffffu, my_hi, cl);
        p_dst[q] = cl * 32 + 4 * (fu ^ (cl & 7));
    }
    auto fetch = [&](int line) {
        if (line < nlines) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float* g = P.x + (long long)c * P.stride + 32LL * line;
                float* s = &ring[c][line % kFolDepth][0][0];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (line >= p_lo[q] && line < p_hi[q]) {
                        const unsigned d = (unsigned)__cvta_generic_to_shared(s + p_dst[q]);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g + p_off[q]) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
#pragma unroll
    for (int l = 0; l < kFolDepth - 1; ++l) fetch(l);
    auto detector = [&](float a, float b) -> float {
        if (MODE == FOL_TRANSIENT || C == 1) return fabsf(a);
        return __fmul_rn(__fadd_rn(fabsf(a), fabsf(b)), 0.5f);      // np.mean(np.abs(audio), axis=1) in float32
    };
    float ef = 0.f, es = 0.f;
    if (active) {
        const float a = __ldg(P.x + r0 + start), b = C > 1 ? __ldg(P.x + r0 + (size_t)P.stride + start) : 0.f;
        ef = es = detector(a, b);                                  // env[0] = |v0| (pipeline.py:499)
    }
    // the very first sample of the unit keeps env[0] = |v0| exactly: the recurrence starts at sample 1 there
    const bool first_exact = active && start == 0;
#pragma unroll 1
    for (int line = 0; line < nlines; ++line) {
        __syncwarp();                                          // every lane is done with the slot the next fetch refills
        fetch(line + kFolDepth - 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kFolDepth - 1) : "memory");
        __syncwarp();                                          // the partners' copies of this lane's line have landed
        if (line < my_lo || line >= my_hi) continue;
        const bool livel = line >= halo_lines;
        const long long i0 = pos0 + 32LL * line;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 va = *reinterpret_cast<const float4*>(&ring[0][line % kFolDepth][lane][4 * (u ^ sx)]);
            float4 vb = va;
            if (C > 1) vb = *reinterpret_cast<const float4*>(&ring[C > 1 ? 1 : 0][line % kFolDepth][lane][4 * (u ^ sx)]);
            float4 oa, ob;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float xa = comp4(va, c), xb = comp4(vb, c);
                const float v = detector(xa, xb);
                if (!(first_exact && line == my_lo && u == 0 && c == 0)) {
                    ef = fol_step_branch(ef, v, P.f_atk, P.f_1matk, P.f_rel, P.f_1mrel);
                    es = fol_step_branch(es, v, P.s_atk, P.s_1matk, P.s_rel, P.s_1mrel);
                }
                if (MODE == FOL_TRANSIENT) {
                    // pipeline.py:1761-1765, float32 throughout (Python-float gains are weak scalars)
                    const float tr = fmaxf(__fsub_rn(ef, es), 0.f);
                    const float ne = __fadd_rn(__fmul_rn(tr, P.attack_gain), __fmul_rn(es, P.sustain_gain));
                    const float gain = fminf(fmaxf(__fdiv_rn(ne, __fadd_rn(ef, 1e-12f)), 0.f), 4.f);
                    setcomp4(oa, c, fminf(fmaxf(__fmul_rn(xa, gain), -1.f), 1.f));
                } else {
                    // pipeline.py:533-541
                    const float diff = fmaxf(__fsub_rn(ef, es), 0.f);
                    const float mask = fminf(fmaxf(__fmul_rn(__fdiv_rn(diff, __fadd_rn(es, 1e-12f)), P.sensitivity), 0.f), 1.f);
                    const float om = __fsub_rn(1.0f, mask);
                    const float la = maximize_limit(xa, P.dyn);
                    setcomp4(oa, c, fminf(fmaxf(__fadd_rn(__fmul_rn(la, om), __fmul_rn(xa, mask)), -1.f), 1.f));
                    if (C > 1) {
                        const float lb = maximize_limit(xb, P.dyn);
                        setcomp4(ob, c, fminf(fmaxf(__fadd_rn(__fmul_rn(lb, om), __fmul_rn(xb, mask)), -1.f), 1.f));
                    }
                }
            }
            if (livel) {
                const long long i = i0 + 4 * u;
                float* da = P.out + r0 + i;
                if (i + 3 < P.n) {
                    *reinterpret_cast<float4*>(da) = oa;
                    if (MODE == FOL_MAXIMIZER_TA && C > 1) *reinterpret_cast<float4*>(da + P.stride) = ob;
                } else {
                    for (int c = 0; c < 4; ++c)
                        if (i + c < P.n) { da[c] = comp4(oa, c); if (MODE == FOL_MAXIMIZER_TA && C > 1) da[P.stride + c] = comp4(ob, c); }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

static double fol_coef(double sr, double t) { return std::exp(-1.0 / std::max(1e-6, sr * t)); }

static void fol_geometry(FolArgs& A, const mm_geom* g, double slowest) {
    const long long nceil = ((g->n + 31) / 32) * 32;
    long long halo = slowest < 1.0 ? (long long)std::ceil(17.5 / -std::log(slowest)) : nceil;
    halo = std::min<long long>(((halo + 31) / 32) * 32, nceil);
    A.halo = halo;
    long long chunk = std::max<long long>(((halo / 2 + 31) / 32) * 32, 4096);
    while ((long long)A.units * ((g->n + chunk - 1) / chunk) > 148LL * 8 * 32 && chunk < nceil) chunk *= 2;
    A.chunk = chunk;
    A.nchunks = (int)((g->n + chunk - 1) / chunk);
}

static void fol_set(FolArgs& A, double sr, double fa, double fr, double sa, double srl) {
    const double c0 = fol_coef(sr, fa), c1 = fol_coef(sr, fr), c2 = fol_coef(sr, sa), c3 = fol_coef(sr, srl);
    A.f_atk = c0; A.f_1matk = 1.0 - c0; A.f_rel = c1; A.f_1mrel = 1.0 - c1;
    A.s_atk = c2; A.s_1matk = 1.0 - c2; A.s_rel = c3; A.s_1mrel = 1.0 - c3;
}

int st_transient_designer(mm_ctx* c, const mm_geom* g, const float* in, float* out, double attack_gain, double sustain_gain) {
    FolArgs A;
    memset(&A, 0, sizeof(A));
    A.x = in; A.out = out; A.n = g->n; A.stride = g->stride; A.units = g->tracks * g->channels;
    fol_set(A, (double)g->sr, 0.0005, 0.005, 0.005, 0.1);
    A.attack_gain = (float)attack_gain; A.sustain_gain = (float)sustain_gain;
    fol_geometry(A, g, fol_coef((double)g->sr, 0.1));
    const long long total = (long long)A.units * A.nchunks;
    KernelScope ks(c, "transient_designer");
    dual_follower_kernel<FOL_TRANSIENT, 1><<<(unsigned)((total + kFolThreads - 1) / kFolThreads), kFolThreads, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int st_maximizer_transient_aware(mm_ctx* c, const mm_geom* g, const float* in, float* out, double sensitivity) {
    FolArgs A;
    memset(&A, 0, sizeof(A));
    A.x = in; A.out = out; A.n = g->n; A.stride = g->stride; A.units = g->tracks;
    fol_set(A, (double)g->sr, 0.0005, 0.002, 0.01, 0.04);
    A.sensitivity = (float)sensitivity;
    fill_dyn(&A.dyn, 6.0, nullptr, 12.0);
    A.dyn.max_top = (float)std::pow(10.0, -0.3 / 20.0);          // apply_maximizer alone: the ceiling, no limiter behind it
    fol_geometry(A, g, fol_coef((double)g->sr, 0.04));
    const long long total = (long long)A.units * A.nchunks;
    const unsigned grid = (unsigned)((total + kFolThreads - 1) / kFolThreads);
    KernelScope ks(c, "maximizer_transient_aware");
    if (g->channels == 2) dual_follower_kernel<FOL_MAXIMIZER_TA, 2><<<grid, kFolThreads, 0, c->stream>>>(A);
    else dual_follower_kernel<FOL_MAXIMIZER_TA, 1><<<grid, kFolThreads, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

// ---- Haas "stereoize" (pipeline.py:1388-1398) on the width-processed pair, one pass -------------------------------
struct HaasArgs {
    const float* in;
    float* out;
    long long n, stride;
    int tracks;
    float width, mix;
    long long delay;
    int apply_width;
};
__global__ void __launch_bounds__(256) haas_kernel(const HaasArgs P) {
    const int track = blockIdx.y;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= P.n) return;
    const size_t r0 = (size_t)(track * 2) * (size_t)P.stride + kLead, r1 = r0 + (size_t)P.stride;
    auto widened = [&](long long k, float& l, float& r) {
        l = P.in[r0 + k]; r = P.in[r1 + k];
        if (!P.apply_width) return;        // the pair is already the merged output of the 4-band mode
        // _imager_apply_width_stereo (pipeline.py:1329-1336)
        const float mid = __fmul_rn(__fadd_rn(l, r), 0.5f);
        const float side = __fmul_rn(__fmul_rn(__fsub_rn(l, r), 0.5f), P.width);
        l = fminf(fmaxf(__fadd_rn(mid, side), -1.f), 1.f);
        r = fminf(fmaxf(__fsub_rn(mid, side), -1.f), 1.f);
    };
    float l, r, dl = 0.f, dr = 0.f;
    widened(i, l, r);
    if (i >= P.delay) widened(i - P.delay, dl, dr);
    P.out[r0 + i] = fminf(fmaxf(__fadd_rn(l, __fmul_rn(P.mix, dr)), -1.f), 1.f);
    P.out[r1 + i] = fminf(fmaxf(__fadd_rn(r, __fmul_rn(P.mix, dl)), -1.f), 1.f);
}

int st_haas_imager(mm_ctx* c, const mm_geom* g, const float* in, float* out, double width, double delay_ms, double mix) {
    if (in == out) { set_error("stereoize: in-place operation is not supported (the delayed tap reads behind the writer)"); return 1; }
    HaasArgs A;
    A.in = in; A.out = out; A.n = g->n; A.stride = g->stride; A.tracks = g->tracks;
    A.apply_width = width == width;       // NaN: skip the mid/side step
    A.width = A.apply_width ? (float)width : 1.f;
    A.mix = (float)std::min(0.35, std::max(0.0, mix));
    long long d = std::min<long long>((long long)((double)g->sr * delay_ms / 1000.0), g->n - 1);
    A.delay = std::max<long long>(0, d);
    dim3 grid((unsigned)((g->n + 255) / 256), (unsigned)g->tracks);
    KernelScope ks(c, "stereoize_haas");
    haas_kernel<<<grid, 256, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

// ---- apply_target_curve_linear_phase (pipeline.py:220-235): fftconvolve(x, ir, mode="same") with the 4096-tap IR of
// design.cpp: linear_phase_target_ir, then clip.  Evaluated as a direct-form FIR: x window and the (reversed) taps in
// shared memory, 8 outputs per thread from registers, float32 products summed per 64-tap block and carried in float64
// (the reference's pocketfft runs in float32: both sides sit ~1e-6 from the exact convolution).  4096 MAC per sample make
// this stage FP32-FMA bound (it is an option of the v2 target-curve module, not on the default chains); an
// overlap-save FFT version on the spectrum kernel's Stockham passes is the next step.
constexpr int kFirThreads = 256;
constexpr int kFirPer = 8;
constexpr int kFirTile = kFirThreads * kFirPer;     // 2048 outputs per CTA
struct FirArgs {
    const float* in;
    float* out;
    const float* taps;       // device, K floats
    long long n, stride;
    int K, center;           // out[i] = sum_k h[k] x[i + center - k]
    int clip;
};
__global__ void __launch_bounds__(kFirThreads) fir_same_kernel(const FirArgs P) {
    extern __shared__ __align__(16) float fsm[];
    float* sh = fsm;                                  // hr[j] = h[K - 1 - j]  (so that x and taps walk the same way)
    float* sx = fsm + P.K;                            // x[base + center - (K - 1) + j], j = 0 .. kFirTile + K - 2 (+ pad)
    const int row = blockIdx.y;
    const float* src = P.in + (size_t)row * (size_t)P.stride + kLead;
    const long long base = (long long)blockIdx.x * kFirTile;
    for (int j = threadIdx.x; j < P.K; j += kFirThreads) sh[j] = __ldg(P.taps + (P.K - 1 - j));
    const long long x0 = base + P.center - (P.K - 1);
    for (int j = threadIdx.x; j < kFirTile + P.K + 8; j += kFirThreads) {
        const long long i = x0 + j;
        sx[j] = (i >= 0 && i < P.n) ? __ldcs(src + i) : 0.f;
    }
    __syncthreads();
    // output o = base + t0 + u:  sum_j hr[j] x[o + center - (K-1) + j] = sum_j sh[j] sx[t0 + u + j]
    const int t0 = threadIdx.x * kFirPer;
    double acc[kFirPer];
#pragma unroll
    for (int u = 0; u < kFirPer; ++u) acc[u] = 0.0;
#pragma unroll 1
    for (int jb = 0; jb < P.K; jb += 64) {
        float part[kFirPer];
#pragma unroll
        for (int u = 0; u < kFirPer; ++u) part[u] = 0.f;
#pragma unroll
        for (int j8 = 0; j8 < 64; j8 += 8) {
            float w[16], hv[8];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 q = *reinterpret_cast<const float4*>(&sx[t0 + jb + j8 + 4 * v]);
                w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const float4 q = *reinterpret_cast<const float4*>(&sh[jb + j8 + 4 * v]);
                hv[4 * v] = q.x; hv[4 * v + 1] = q.y; hv[4 * v + 2] = q.z; hv[4 * v + 3] = q.w;
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                for (int u = 0; u < kFirPer; ++u) part[u] = fmaf(hv[jj], w[u + jj], part[u]);
        }
#pragma unroll
        for (int u = 0; u < kFirPer; ++u) acc[u] += (double)part[u];
    }
    float* dst = P.out + (size_t)row * (size_t)P.stride + kLead;
#pragma unroll
    for (int u = 0; u < kFirPer; ++u) {
        const long long o = base + t0 + u;
        if (o < P.n) {
            float r = (float)acc[u];
            if (P.clip) r = fminf(fmaxf(r, -1.f), 1.f);
            dst[o] = r;
        }
    }
}

// Long filters run as one FFT convolution per track (bigfft.cu: 2 x 2^24-point transforms for a 180 s stereo track instead of
// K multiply-adds per sample); the direct kernel keeps short filters, rows beyond the 2^27-point transform (a 2-hour 96 kHz
// file) and MM_FIR=direct.
static bool fir_use_fft(const mm_geom* g, int K) {
    static const bool direct = [] { const char* e = getenv("MM_FIR"); return e && !strcmp(e, "direct"); }();
    return !direct && K >= 1024 && g->n >= 4 * (long long)K && fft_convolve_fits(g, K);
}

int st_target_curve_linear_phase(mm_ctx* c, const mm_geom* g, const float* in, float* out) {
    if (in == out) { set_error("linear-phase target curve: in-place operation is not supported"); return 1; }
    const int K = 4096;
    std::map<int, float*>& cache = c->lp_taps;        // per context: device pointers never cross devices or threads
    float* taps = nullptr;
    auto it = cache.find(g->sr);
    if (it != cache.end()) taps = it->second;
    else {
        std::vector<float> ir(K);
        if (!linear_phase_target_ir(g->sr, K, ir.data())) { set_error("linear-phase target curve: IR design failed for %d Hz", g->sr); return 1; }
        MM_CUDA(cudaMalloc(&taps, K * sizeof(float)));
        MM_CUDA(cudaMemcpyAsync(taps, ir.data(), K * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        MM_CUDA(cudaStreamSynchronize(c->stream));
        cache[g->sr] = taps;
    }
    if (fir_use_fft(g, K)) return st_fft_convolve_same(c, g, in, out, taps, K, 1);
    FirArgs A;
    A.in = in; A.out = out; A.taps = taps; A.n = g->n; A.stride = g->stride; A.K = K; A.center = (K - 1) / 2; A.clip = 1;
    const size_t smem = (size_t)(K + kFirTile + K + 8 + 8) * sizeof(float);
    MM_CUDA(cudaFuncSetAttribute(fir_same_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((g->n + kFirTile - 1) / kFirTile), (unsigned)(g->tracks * g->channels));
    KernelScope ks(c, "target_curve_linear_phase_fir4096");
    fir_same_kernel<<<grid, kFirThreads, smem, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int st_fir_same(mm_ctx* c, const mm_geom* g, const float* in, float* out, const float* taps_dev, int K, int clip) {
    if (in == out) { set_error("FIR: in-place operation is not supported"); return 1; }
    if (K < 64 || (K % 64) != 0 || K > 16384) { set_error("FIR: the tap count must be a multiple of 64 in [64, 16384]"); return 1; }
    if (fir_use_fft(g, K)) return st_fft_convolve_same(c, g, in, out, taps_dev, K, clip);
    FirArgs A;
    A.in = in; A.out = out; A.taps = taps_dev; A.n = g->n; A.stride = g->stride; A.K = K; A.center = (K - 1) / 2; A.clip = clip;
    const size_t smem = (size_t)(K + kFirTile + K + 8 + 8) * sizeof(float);
    MM_CUDA(cudaFuncSetAttribute(fir_same_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 98304)));
    dim3 grid((unsigned)((g->n + kFirTile - 1) / kFirTile), (unsigned)(g->tracks * g->channels));
    KernelScope ks(c, "fir_same");
    fir_same_kernel<<<grid, kFirThreads, smem, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

// ---- 4-band stereo imager (pipeline.py:1360-1386): _split_bands, per-band mid/side width, sum, clip ------------------
// The split is the dynamics stage's (same designs, same sweeps); the per-band width and the merge are one pointwise pass
// over the four stored band pairs, in float64 like the reference's (its bands come out of filtfilt as float64; ours are
// read back from float32 storage), accumulated into float32 band by band as numpy's in-place `out_l += ol` does.
struct Imager4Args {
    const float* band[4];
    float* out;
    long long n, stride;
    double width[4];
};
__global__ void __launch_bounds__(256) imager4_kernel(const Imager4Args P) {
    const int track = blockIdx.y;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= P.n) return;
    const size_t r0 = (size_t)(track * 2) * (size_t)P.stride + kLead + (size_t)i, r1 = r0 + (size_t)P.stride;
    float ol = 0.f, orr = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double l = (double)P.band[b][r0], r = (double)P.band[b][r1];
        const double mid = (l + r) * 0.5, side = (l - r) * 0.5 * P.width[b];
        ol = (float)((double)ol + fmin(fmax(mid + side, -1.0), 1.0));
        orr = (float)((double)orr + fmin(fmax(mid - side, -1.0), 1.0));
    }
    P.out[r0] = fminf(fmaxf(ol, -1.f), 1.f);
    P.out[r1] = fminf(fmaxf(orr, -1.f), 1.f);
}

int st_imager4(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* widths, const double* crossovers_hz) {
    double cross[3] = {214.0, 3500.0, 10000.0};                    // MULTIBAND_CROSSOVERS_HZ (pipeline.py:94)
    if (crossovers_hz) {
        double t[3];
        for (int i = 0; i < 3; ++i) t[i] = std::min(std::max(crossovers_hz[i], 20.0), 20000.0);
        if (!(t[0] >= t[1] || t[1] >= t[2])) { cross[0] = t[0]; cross[1] = t[1]; cross[2] = t[2]; }
    }
    float* bands[4];
    MM_TRY(st_split_bands(c, g, in, cross, bands));
    Imager4Args A;
    for (int b = 0; b < 4; ++b) { A.band[b] = bands[b]; A.width[b] = widths[b]; }
    A.out = out; A.n = g->n; A.stride = g->stride;
    dim3 grid((unsigned)((g->n + 255) / 256), (unsigned)g->tracks);
    KernelScope ks(c, "imager_4band_merge");
    imager4_kernel<<<grid, 256, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mm
