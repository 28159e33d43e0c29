// The envelope-compressor branch of apply_multiband_dynamics (backend/app/pipeline.py:373-411, selected at :442-465 whenever
// `pedalboard` imports): per band  pedalboard.Compressor(threshold_db, ratio, attack_ms, release_ms)  ->  hard clip at lim_db
// ->  x gain, the four bands summed, then (apply_dynamics, :610-641) maximizer and limiter.
//
// PARITY UNPINNED: pedalboard is not installed in the build image and no test of the reference touches this branch, so the
// arithmetic below restates the published JUCE sources pedalboard wraps (juce::dsp::Compressor<float> over
// juce::dsp::BallisticsFilter<float>, peak level type) and is checked against its own CPU restatement
// (oracle/chain.py compress_band_envelope), not against pedalboard itself:
//     a    = |x|                                   c = a > y_prev ? cteAT : cteRL,   cte = exp(-2 pi 1000 / (sr t_ms))
//     y    = a + c (y_prev - a)                    (float32; y_prev = 0 at the start of a channel)
//     gain = y < thr ? 1 : (y / thr)^(1/ratio - 1) (thr = 10^(dB/20))
//     out  = gain x
// (JUCE's per-block snap-to-zero of states below 1e-8 is not reproduced.)
//
// The follower has no associative operator, but one step is a monotone piecewise-linear map of y_prev with slopes cteAT / cteRL
// < 1, so two runs started from different states converge at least as fast as the slower coefficient decays: a row is cut into
// chunks that each start `halo` = 17.5 time constants early from state 0 (start-up error e^-17.5 = 2.5e-8 of the peak), exactly
// like the de-esser's follower (deesser.cu).  One thread walks one chunk of one row through the three (or four) followers, the
// band limiters, the sum, the maximizer and the limiter, and writes the finished sample: the bands are read once (plus the halo
// of the follower bands) and nothing but the output is written -- 4 R + 1 W per channel-sample after a 2 R / 2 W backward sweep
// that materialises the two middle bands, against the soft-knee mode's single 4 R / 1 W epilogue (+4 words).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "context.h"
#include "pointwise.cuh"
#include "stages_internal.h"

namespace mm {

constexpr int kBcThreads = 32;
// ring geometry, measured with the cooperative fetch (64 x 180 s, ms per launch; MM_BC_DEPTH x MM_BC_LINE samples): 2 x 16: 7.2,
// 3 x 16: 7.4, 4 x 16: 8.1, 2 x 32: 8.0, 3 x 32: 10.3, 2..4 x 8: 8.8 -- 16 KB per one-warp CTA (13 warps per SM) wins
#ifndef MM_BC_DEPTH
#define MM_BC_DEPTH 2
#endif
#ifndef MM_BC_LINE
#define MM_BC_LINE 16
#endif
constexpr int kBcDepth = MM_BC_DEPTH;       // lines in flight per thread and stream
constexpr int kBcLine = MM_BC_LINE;         // samples per line (64 bytes)

struct BandCompArgs {
    const float* band[4];
    float* out;
    long long n, stride;
    int rows;
    long long chunk, halo;                  // multiples of kBcLine
    int nchunks;
    int env[4];                             // 1: envelope compressor (ratio >= 1); 0: memoryless soft-knee chain (ratio < 1, upward)
    float cat[4], crl[4], thr[4], thr_inv[4], pw[4];
    DynParams dyn;
};

__device__ __forceinline__ void bc_cp16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    // L2::256B: every thread walks its own sequential stream (tens of thousands of streams at once, far more than HBM has open
    // rows); asking L2 to pull the whole 256-byte block on the first touch turns four row activations per block into one
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// one BallisticsFilter step (peak rectifier): y = a + c (y_prev - a), c by the sign of (y_prev - a); both products are formed,
// the select sits behind them (dependency chain: FSUB, FMUL, FSEL, FADD)
__device__ __forceinline__ float bc_step(float e, float a, float cat, float crl) {
    const float d = __fsub_rn(e, a);
    const float m = d < 0.f ? __fmul_rn(cat, d) : __fmul_rn(crl, d);
    return __fadd_rn(a, m);
}
// juce::dsp::Compressor::processSample: the VCA gain of one envelope value, (e / thr)^(1/ratio - 1) above the threshold.
// MUFU.LG2 / MUFU.EX2 (2^-22 absolute on the logarithm, 2 ulp on the power): the gain is good to ~3e-7 relative, far inside
// what an unpinned restatement of std::pow can claim, and an order of magnitude cheaper than powf; branch free.
// The .ftz forms are single MUFU instructions; without them the compiler wraps each in denormal-range scaling (several FSETP /
// FMUL / FSEL per call).  The arguments never come near that range here: max(e, thr) / thr is in [1, ~1e3], the exponent in
// [-10, 0] -- so the values are bit for bit those of __log2f / exp2f.
__device__ __forceinline__ float bc_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float bc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float bc_gain(float e, float thr, float thr_inv, float pw) {
#ifdef MM_BC_LIBM_POW
    const float g = exp2f(__fmul_rn(pw, __log2f(__fmul_rn(fmaxf(e, thr), thr_inv))));
#else
    const float g = bc_ex2(__fmul_rn(pw, bc_lg2(__fmul_rn(fmaxf(e, thr), thr_inv))));
#endif
    return e < thr ? 1.f : g;
}

// ENV: bit k set = band k runs the envelope compressor (compile-time for the default configuration 0b1110: band 0 has ratio 1)
//
// Fetch pattern.  Lane l walks chunk l of the warp, so the 32 lanes read 32 different places of a row; the first version let
// every lane copy its OWN line (cp.async, 16 bytes at a time): each warp-level copy instruction touched 32 different 128-byte
// lines = 32 trips through the L1 / LSU pipeline, one instruction per band and four samples -- about one LSU wavefront per
// sample and SM cycle, which is what bounded the kernel (9.5 ms per 64-track batch whatever the occupancy: 6 or 13 warps per SM,
// 32- or 64-byte lines all measured the same).  Now the lanes fetch COOPERATIVELY: kUnits consecutive lanes copy the kUnits
// 16-byte units of ONE chunk's line, so an instruction touches 32 / kUnits lines, each completely (8x fewer wavefronts with
// 128-byte lines); a __syncwarp() hands the landed lines to their owners.  Trip counts are uniform across the warp (a chunk
// without a left neighbourhood or with a short tail just skips the lines outside [0, n)).
template <int ENV>
__global__ void __launch_bounds__(kBcThreads) band_compress_kernel(const __grid_constant__ BandCompArgs P) {
    __shared__ __align__(128) float ring[kBcDepth][4][kBcThreads][kBcLine];
    constexpr int kUnits = kBcLine / 4;                          // 16-byte units per line = lanes that fetch one line together
    constexpr int kLPI = 32 / kUnits;                            // lines one warp-level copy instruction covers
    constexpr int kSxShift = kUnits >= 8 ? 0 : (kUnits == 4 ? 1 : 2);
    const int lane = threadIdx.x;
    const long long gid = (long long)blockIdx.x * kBcThreads + lane;
    const long long total = (long long)P.rows * P.nchunks;
    const bool active = gid < total;
    const int row = active ? (int)(gid / P.nchunks) : 0;
    const int chunk = active ? (int)(gid % P.nchunks) : 0;
    const long long ro = (long long)row * P.stride + kLead;
    const long long live0 = (long long)chunk * P.chunk;
    const long long pos0 = live0 - P.halo;                       // position of line 0 (negative: no left neighbourhood, lines skipped)
    const long long hi = active ? min(live0 + P.chunk, P.n) : pos0;
    const int halo_lines = (int)(P.halo / kBcLine);
    const int nlines = halo_lines + (int)(P.chunk / kBcLine);    // uniform; P.halo and P.chunk are multiples of kBcLine
    // lines of this lane's chunk that hold samples of [0, n): [my_lo, my_hi)
    const int my_lo = pos0 < 0 ? (int)((-pos0) / kBcLine) : 0;
    const int my_hi = hi > pos0 ? (int)((hi - pos0 + kBcLine - 1) / kBcLine) : 0;
    // 16-byte unit swizzle: a quarter-warp's float4 reads of its own rows hit 8 different bank groups
    auto sx_of = [&](int l) -> int { return (l >> kSxShift) & (kUnits - 1); };
    const int sx = sx_of(lane);
    auto env_on = [&](int k) -> bool { return ENV >= 0 ? ((ENV >> k) & 1) != 0 : P.env[k] != 0; };
    // the chunks this lane helps to fetch: in copy instruction q, chunk-lane q * kLPI + lane / kUnits, unit lane % kUnits
    const int fu = lane % kUnits;
    long long p_off[kUnits];
    int p_lo[kUnits], p_hi[kUnits], p_dst[kUnits];
#pragma unroll
    for (int q = 0; q < kUnits; ++q) {
        const int cl = q * kLPI + lane / kUnits;
        p_off[q] = __shfl_sync(0xffffffffu, ro + pos0, cl) + 4 * fu;
        p_lo[q] = __shfl_sync(0xffffffffu, my_lo, cl);
        p_hi[q] = __shfl_sync(0xffffffffu, my_hi, cl);
        p_dst[q] = cl * kBcLine + 4 * (fu ^ sx_of(cl));
    }
    auto fetch = [&](int line) {
        if (line < nlines) {
            const bool live = line >= halo_lines;
            float* slot = &ring[line % kBcDepth][0][0][0];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!live && !env_on(k)) continue;              // a halo only feeds the followers
                const float* g = P.band[k] + (long long)kBcLine * line;
#pragma unroll
                for (int q = 0; q < kUnits; ++q)
                    if (line >= p_lo[q] && line < p_hi[q]) bc_cp16(slot + k * (kBcThreads * kBcLine) + p_dst[q], g + p_off[q]);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
#pragma unroll
    for (int l = 0; l < kBcDepth - 1; ++l) fetch(l);
    float e[4] = {0.f, 0.f, 0.f, 0.f};                           // BallisticsFilter::reset(): yold = 0
    // halo: only the states matter.  Loops over the units of a line are rolled: the body has to stay inside the instruction cache
#pragma unroll 1
    for (int line = 0; line < halo_lines; ++line) {
        __syncwarp();                                            // every lane is done with the slot the next fetch refills
        fetch(line + kBcDepth - 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kBcDepth - 1) : "memory");
        __syncwarp();                                            // the partners' copies of this lane's line have landed
        if (line < my_lo || line >= my_hi) continue;
#pragma unroll 1
        for (int u = 0; u < kUnits; ++u) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (env_on(k)) v[k] = *reinterpret_cast<const float4*>(&ring[line % kBcDepth][k][lane][4 * (u ^ sx)]);
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (env_on(k)) e[k] = bc_step(e[k], fabsf(comp4(v[k], c)), P.cat[k], P.crl[k]);
        }
    }
    float par_mix = 0.f, par_one_minus = 1.f;
    if (P.dyn.par_mix) {
        const double mixd = __ldg(P.dyn.par_mix + row);
        par_mix = (float)mixd;
        par_one_minus = (float)(1.0 - mixd);
    }
    float* dst = P.out + ro;
#pragma unroll 1
    for (int line = halo_lines; line < nlines; ++line) {
        __syncwarp();
        fetch(line + kBcDepth - 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kBcDepth - 1) : "memory");
        __syncwarp();
        if (line >= my_hi) continue;
        const long long i0 = pos0 + (long long)kBcLine * line;
#pragma unroll 1
        for (int u = 0; u < kUnits; ++u) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(&ring[line % kBcDepth][k][lane][4 * (u ^ sx)]);
            float4 o;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float x = comp4(v[k], c);
                    float y;
                    if (env_on(k)) {
                        e[k] = bc_step(e[k], fabsf(x), P.cat[k], P.crl[k]);
                        const float g = bc_gain(e[k], P.thr[k], P.thr_inv[k], P.pw[k]);
                        const DynBand& b = P.dyn.band[k];
                        y = __fmul_rn(fminf(fmaxf(__fmul_rn(g, x), -b.lim), b.lim), b.gain);     // compressor -> limiter -> gain (:404-409)
                    } else {
                        y = band_chain_gen(x, P.dyn.band[k]);                                      // :466-474 (ratio < 1: upward)
                    }
                    acc = k == 0 ? y : __fadd_rn(acc, y);                                          // _merge_bands: float32 sums in band order
                }
                float res = maximize_limit(acc, P.dyn);
                if (par_mix >= 0.01f) res = parallel_compress(res, par_mix, par_one_minus, P.dyn);
                setcomp4(o, c, res);
            }
            const long long i = i0 + 4 * u;
            if (i + 3 < P.n) __stcs(reinterpret_cast<float4*>(dst + i), o);
            else {
                for (int c = 0; c < 4; ++c) if (i + c < P.n) dst[i + c] = comp4(o, c);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// attack / release per band as the reference picks them (pipeline.py:451-456)
static void band_times(int i, double* atk_ms, double* rel_ms) {
    if (i >= 3) { *atk_ms = 18.0; *rel_ms = 180.0; }
    else if (i == 2) { *atk_ms = 12.0; *rel_ms = 130.0; }
    else { *atk_ms = 10.0; *rel_ms = 80.0; }
}

// bands[0..3]: the four zero-phase bands as float32 rows (the reference casts each band to float32 before pedalboard, :398);
// d: fill_dyn()'s band table (limiters, gains, the soft-knee lines of bands that stay memoryless) and maximizer constants
int launch_band_compress(mm_ctx* c, const mm_geom* g, const float* const* bands, float* out, const DynParams& d) {
    BandCompArgs A;
    memset(&A, 0, sizeof(A));
    const int rows = g->tracks * g->channels;
    for (int k = 0; k < 4; ++k) A.band[k] = bands[k];
    A.out = out; A.n = g->n; A.stride = g->stride; A.rows = rows; A.dyn = d;
    double slow = 0.0;
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < 4; ++k) {
        const DynBand& b = d.band[k];
        // `use_pedalboard and ratio >= 1.0` (:457); at ratio == 1 the VCA exponent is 0 and the compressor returns its input,
        // which is also what the memoryless chain does with a bypassed band (mode 0): no follower needed
        A.env[k] = b.ratio > 1.0 ? 1 : 0;
        double atk_ms, rel_ms;
        band_times(k, &atk_ms, &rel_ms);
        // BallisticsFilter: cte = exp(expFactor / t_ms), expFactor = -2 pi 1000 / sr; t < 1e-3 ms -> 0
        const double cat = std::exp(-2.0 * pi * 1000.0 / ((double)g->sr * atk_ms)), crl = std::exp(-2.0 * pi * 1000.0 / ((double)g->sr * rel_ms));
        A.cat[k] = (float)cat; A.crl[k] = (float)crl;
        A.thr[k] = (float)std::pow(10.0, b.thr_db / 20.0);
        A.thr_inv[k] = 1.0f / A.thr[k];
        A.pw[k] = 1.0f / (float)std::max(b.ratio, 1.0) - 1.0f;
        if (A.env[k]) slow = std::max(slow, std::max((double)A.cat[k], (double)A.crl[k]));
    }
    const long long nceil = ((g->n + kBcLine - 1) / kBcLine) * kBcLine;
    long long halo = (slow > 0.0 && slow < 1.0) ? (long long)std::ceil(17.5 / -std::log(slow)) : 0;
    halo = std::min<long long>(((halo + kBcLine - 1) / kBcLine) * kBcLine, nceil);
    A.halo = halo;
    int mask = 0;
    for (int k = 0; k < 4; ++k) mask |= A.env[k] << k;
    auto kern = mask == 0xE ? band_compress_kernel<0xE> : band_compress_kernel<-1>;
    int bps = 0;
    MM_TRY(kernel_setup(c, (const void*)kern, kBcThreads, 0, true, &bps));
    const long long capacity = (long long)std::max(1, bps) * c->num_sms * kBcThreads;     // threads of one full wave
    long long chunk = std::max<long long>(halo, ((long long)rows * g->n + capacity - 1) / capacity);
    chunk = std::max<long long>(((chunk + kBcLine - 1) / kBcLine) * kBcLine, 1024);
    chunk = std::min(chunk, nceil);
    A.chunk = chunk;
    A.nchunks = (int)((g->n + chunk - 1) / chunk);
    const long long total = (long long)rows * A.nchunks;
    KernelScope ks(c, "band_envelope_compress");
    ks.samples = (double)rows * (double)g->n;
    kern<<<(unsigned)((total + kBcThreads - 1) / kBcThreads), kBcThreads, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mm
