// apply_deesser (backend/app/pipeline.py:1200-1264) and its attack/release envelope follower
// (_envelope_follower_core, pipeline.py:495-507).
//
//   sc   = filtfilt(butter(2, [lo, hi], 'band'), x)          4th-order ba section, padlen 15  -> sweep kernels (M = 4)
//   env  = follower(|sc|, attack, release)                    nonlinear one-pole recurrence     -> envelope_gain_kernel
//   g    = clip(where(env > thr, thr + (env - thr)/ratio, env) / (env + 1e-12), 0.35, 1)
//   g    = clip(box_k(g), 0.35, 1)     k = odd(max(3, int(sr * 0.0015))), zero-padded edges    -> deesser_apply_kernel
//   out  = x - sc + sc * g
//
// The follower has no associative operator, but the map e -> e' of one step,
//   e' = max(atk e + (1 - atk) v, rel e + (1 - rel) v)       (atk < rel; equals the reference's branch on v > e)
// is monotone and contracts the distance between any two states by at least `rel` per sample.  A
// row is therefore cut into chunks that each start `halo` samples early from the state |v|: after
// halo = ceil(17.5 * sr * release) samples the start-up error has shrunk by e^-17.5 = 2.5e-8 of
// the signal peak, below float32 resolution of the envelope.  Chunk 0 starts at sample 0 with the
// reference's own initial state env[0] = |v[0]| and is exact.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.h"
#include "stages_internal.h"

namespace mm {

__device__ __forceinline__ void cp_async16_env(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    // L2::256B: one DRAM row activation per 256-byte block of a thread's private sequential stream instead of two
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}

struct EnvArgs {
    const float* sc;       // planar side-chain rows
    float* gain;           // planar pre-smoothing gain rows (same geometry)
    long long n, stride;
    int rows;
    long long chunk, halo; // multiples of 32 samples (one 128-byte line)
    int nchunks;           // per row
    float atk, one_m_atk, rel, one_m_rel;
    float atk_lo, rel_lo;  // float64 coefficient = (atk + atk_lo): without the low part a float32-rounded release coefficient
                           // biases a 4000-sample decay by up to 1e-4 of the envelope
    float thr, inv_ratio;
    int mode;              // 0: de-esser gain, 1: dynamic-EQ band gain
    float ratio, max_cut;
    double d_atk, d_1matk, d_rel, d_1mrel;   // mode 1: float64 coefficients and products, float32 state (numba's arithmetic,
                                             // pipeline.py:495-507): a float32 coefficient biases long release tails by ~1e-4
};

__device__ __forceinline__ float deess_gain(float env, float thr, float inv_ratio) {
    // pipeline.py:1247-1252 in float32 (thr / ratio are Python floats -> weak -> float32 arithmetic)
    const float red = env > thr ? fmaf(env - thr, inv_ratio, thr) : env;
    const float g = env > 1e-10f ? __fdividef(red, env + 1e-12f) : 1.0f;
    return fminf(fmaxf(g, 0.35f), 1.0f);
}

// apply_dynamic_eq (pipeline.py:1680-1689), float32 as numpy evaluates it: where(env > thr, clip((thr + (env - thr) / ratio)
// / (env + 1e-12), max_cut, 1), 1), then clip(., 0.3, 1)
__device__ __forceinline__ float dyneq_gain(float env, float thr, float ratio, float max_cut) {
    float g = 1.0f;
    if (env > thr) {
        const float red = __fadd_rn(thr, __fdiv_rn(__fsub_rn(env, thr), ratio));
        g = fminf(fmaxf(__fdiv_rn(red, __fadd_rn(env, 1e-12f)), max_cut), 1.0f);
    }
    return fminf(fmaxf(g, 0.3f), 1.0f);
}

// One follower step; `ep` is the envelope one sample earlier.  The low parts of the coefficients multiply `ep` instead of `e`
// (the difference, lo * (e - ep), is far below float32 resolution), which keeps them off the e -> e' dependency chain.
template <int MODE> __device__ __forceinline__ float env_step(float e, float ep, float v, const EnvArgs& P) {
    if (MODE) {
        const double ed = (double)e, vd = (double)v;
        return v > e ? (float)(P.d_atk * ed + P.d_1matk * vd) : (float)(P.d_rel * ed + P.d_1mrel * vd);
    }
    // max of the attack and the release update == the reference's branch on v > e (atk < rel)
    const float a = fmaf(P.atk, e, fmaf(P.atk_lo, ep, P.one_m_atk * v));
    const float r = fmaf(P.rel, e, fmaf(P.rel_lo, ep, P.one_m_rel * v));
    return fmaxf(a, r);
}
#define MM_ENV_STEP(val) do { const float _en = env_step<MODE>(e, ep, fabsf(val), P); ep = e; e = _en; } while (0)

// One thread = one chunk (+ its halo) of one row: a strictly sequential recurrence, so the kernel is
// latency bound by design; each thread streams its samples through a private ring of 128-byte lines in
// shared memory (cp.async, kEnvDepth lines in flight) so that HBM latency never sits on the chain.
#ifndef MM_ENV_DEPTH
#define MM_ENV_DEPTH 4
#endif
constexpr int kEnvDepth = MM_ENV_DEPTH;
constexpr int kEnvThreads = 32;

// Lines are fetched COOPERATIVELY (as in bandcomp.cu): lane l walks chunk l, but eight consecutive lanes copy the eight 16-byte
// units of ONE chunk's line, so a warp-level cp.async touches 4 lines completely instead of 32 lines partially (one trip through
// the L1 / LSU pipeline per touched line: the per-lane pattern spent 32 trips per instruction).  Trip counts are uniform across
// the warp; a chunk with a truncated left neighbourhood or a short tail skips the lines outside its range.
template <int MODE> __global__ void __launch_bounds__(kEnvThreads) envelope_gain_kernel(const EnvArgs P) {
    __shared__ __align__(128) float ring[kEnvDepth][kEnvThreads][32];
    const int lane = threadIdx.x;
    const long long gid = (long long)blockIdx.x * kEnvThreads + lane;
    const long long total = (long long)P.rows * P.nchunks;
    const bool active = gid < total;
    const int row = active ? (int)(gid / P.nchunks) : 0;
    const int chunk = active ? (int)(gid % P.nchunks) : 0;
    const long long ro = (long long)row * P.stride + kLead;
    const float* src = P.sc + ro;
    float* dst = P.gain + ro;
    const long long live0 = (long long)chunk * P.chunk;
    const long long live1 = active ? min(live0 + P.chunk, P.n) : live0;
    const long long pos0 = live0 - P.halo;                     // position of line 0; negative: the chunk starts at sample 0
    const long long start = max(pos0, 0LL);
    const int halo_lines = (int)(P.halo / 32);
    const int nlines = halo_lines + (int)(P.chunk / 32);       // uniform (P.halo, P.chunk: multiples of 32)
    const int my_lo = pos0 < 0 ? (int)((-pos0) / 32) : 0;      // this chunk's lines with samples of [0, n): [my_lo, my_hi)
    const int my_hi = (active && live1 > pos0) ? (int)((live1 - pos0 + 31) / 32) : 0;
    const int sx = lane & 7;                                   // 16-byte unit swizzle of this thread's lines
    const int fu = lane & 7;                                   // the unit this lane copies, of chunk-lane 4 q + lane / 8 in instruction q
    long long p_off[8];
    int p_lo[8], p_hi[8], p_dst[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int cl = 4 * q + (lane >> 3);
        p_off[q] = __shfl_sync(0xffffffffu, ro + pos0, cl) + 4 * fu;
        p_lo[q] = __shfl_sync(0xffffffffu, my_lo, cl);
        p_hi[q] = __shfl_sync(0xffffffffu, my_hi, cl);
        p_dst[q] = cl * 32 + 4 * (fu ^ (cl & 7));
    }
    auto fetch = [&](int line) {
        if (line < nlines) {
            const float* g = P.sc + 32LL * line;
            float* s = &ring[line % kEnvDepth][0][0];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (line >= p_lo[q] && line < p_hi[q]) cp_async16_env(s + p_dst[q], g + p_off[q]);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
#pragma unroll
    for (int l = 0; l < kEnvDepth - 1; ++l) fetch(l);
    // env[0] = |v0|: starting from e = |v0| the first update returns |v0| again (to within one ulp), so the
    // recurrence below needs no special case for the chunk's first sample
    float e = active ? fabsf(__ldg(src + start)) : 0.f;
    float ep = e;
    // halo: only the state matters -- 5 instructions per sample, all but two of them off the dependency chain
#pragma unroll 1
    for (int line = 0; line < halo_lines; ++line) {
        __syncwarp();                                          // every lane is done with the slot the next fetch refills
        fetch(line + kEnvDepth - 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kEnvDepth - 1) : "memory");
        __syncwarp();                                          // the partners' copies of this lane's line have landed
        if (line < my_lo || line >= my_hi) continue;
        const float* s = &ring[line % kEnvDepth][lane][0];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 v = *reinterpret_cast<const float4*>(s + 4 * (u ^ sx));
            MM_ENV_STEP(v.x);
            MM_ENV_STEP(v.y);
            MM_ENV_STEP(v.z);
            MM_ENV_STEP(v.w);
        }
    }
#pragma unroll 1
    for (int line = halo_lines; line < nlines; ++line) {
        __syncwarp();
        fetch(line + kEnvDepth - 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kEnvDepth - 1) : "memory");
        __syncwarp();
        if (line >= my_hi) continue;
        const float* s = &ring[line % kEnvDepth][lane][0];
        const long long i0 = pos0 + 32LL * line;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 v = *reinterpret_cast<const float4*>(s + 4 * (u ^ sx));
            float4 g;
            MM_ENV_STEP(v.x); g.x = MODE ? dyneq_gain(e, P.thr, P.ratio, P.max_cut) : deess_gain(e, P.thr, P.inv_ratio);
            MM_ENV_STEP(v.y); g.y = MODE ? dyneq_gain(e, P.thr, P.ratio, P.max_cut) : deess_gain(e, P.thr, P.inv_ratio);
            MM_ENV_STEP(v.z); g.z = MODE ? dyneq_gain(e, P.thr, P.ratio, P.max_cut) : deess_gain(e, P.thr, P.inv_ratio);
            MM_ENV_STEP(v.w); g.w = MODE ? dyneq_gain(e, P.thr, P.ratio, P.max_cut) : deess_gain(e, P.thr, P.inv_ratio);
            const long long i = i0 + 4 * u;
            if (i + 3 < P.n) *reinterpret_cast<float4*>(dst + i) = g;
            else {
                for (int c = 0; c < 4; ++c) if (i + c < P.n) dst[i + c] = comp4(g, c);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// box smoothing of the gain (np.convolve(g, ones(k)/k, mode="same"), zero padding) + recombination
struct ApplyArgs {
    const float* x;
    const float* sc;
    const float* gain;
    float* out;
    long long n, stride;
    int k;            // odd
    float kerf;       // float32(1 / k)
};
constexpr int kApplyThreads = 256;
constexpr int kApplyPer = 8;                                   // consecutive outputs per thread (two float4)
constexpr int kApplyTile = kApplyThreads * kApplyPer;          // 2048 outputs per CTA
constexpr int kApplyMaxHalf = 256;                             // supports k <= 513 (sr <= 342 kHz)

// padded index: one spare word per 8 so that thread t's window (base 8 t) walks distinct banks
__device__ __forceinline__ int padi(int i) { return i + (i >> 3); }

__global__ void __launch_bounds__(kApplyThreads) deesser_apply_kernel(const ApplyArgs P) {
    __shared__ float sg[kApplyTile + 2 * kApplyMaxHalf + (kApplyTile + 2 * kApplyMaxHalf) / 8 + 8];
    const int row = blockIdx.y;
    const long long base = (long long)blockIdx.x * kApplyTile;
    const int half = P.k / 2;
    const size_t ro = (size_t)row * (size_t)P.stride + kLead;
    const int span = kApplyTile + 2 * half;
    // sg[padi(j)] = g[base - lead + j] with lead = half rounded up to a multiple of 4 (keeps float4 alignment)
    const int lead = (half + 3) & ~3;
    for (int j4 = threadIdx.x * 4; j4 < span + (lead - half) + 3; j4 += kApplyThreads * 4) {
        const long long i = base - lead + j4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= 0 && i + 3 < P.n) v = __ldcs(reinterpret_cast<const float4*>(P.gain + ro + i));
        else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (i + c >= 0 && i + c < P.n) setcomp4(v, c, P.gain[ro + i + c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) sg[padi(j4 + c)] = comp4(v, c);
    }
    // x and side chain of this thread's 8 outputs: issued before the barrier so they overlap the window sums
    const int o0 = threadIdx.x * kApplyPer;
    const long long i0 = base + o0;
    float xs[8], ss[8];
    if (i0 + 7 < P.n) {
        const float4 x0 = __ldcs(reinterpret_cast<const float4*>(P.x + ro + i0)), x1 = __ldcs(reinterpret_cast<const float4*>(P.x + ro + i0 + 4));
        const float4 s0 = __ldcs(reinterpret_cast<const float4*>(P.sc + ro + i0)), s1 = __ldcs(reinterpret_cast<const float4*>(P.sc + ro + i0 + 4));
        xs[0] = x0.x; xs[1] = x0.y; xs[2] = x0.z; xs[3] = x0.w; xs[4] = x1.x; xs[5] = x1.y; xs[6] = x1.z; xs[7] = x1.w;
        ss[0] = s0.x; ss[1] = s0.y; ss[2] = s0.z; ss[3] = s0.w; ss[4] = s1.x; ss[5] = s1.y; ss[6] = s1.z; ss[7] = s1.w;
    } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool ok = i0 + u < P.n;
            xs[u] = ok ? P.x[ro + i0 + u] : 0.f;
            ss[u] = ok ? P.sc[ro + i0 + u] : 0.f;
        }
    }
    __syncthreads();
    // window of output o covers g[o - half .. o + half] = sg index (o - half + lead) .. (+ k - 1)
    const int w0 = o0 + lead - half;
    double acc = 0.0;
    for (int j = 0; j < P.k; ++j) acc += (double)sg[padi(w0 + j)];
    float res[8];
#pragma unroll
    for (int u = 0; u < kApplyPer; ++u) {
        float g = (float)(acc * (double)P.kerf);
        g = fminf(fmaxf(g, 0.35f), 1.0f);
        res[u] = __fadd_rn(__fsub_rn(xs[u], ss[u]), __fmul_rn(ss[u], g));
        acc += (double)sg[padi(w0 + u + P.k)] - (double)sg[padi(w0 + u)];
    }
    if (i0 + 7 < P.n) {
        __stcs(reinterpret_cast<float4*>(P.out + ro + i0), make_float4(res[0], res[1], res[2], res[3]));
        __stcs(reinterpret_cast<float4*>(P.out + ro + i0 + 4), make_float4(res[4], res[5], res[6], res[7]));
    } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i0 + u < P.n) P.out[ro + i0 + u] = res[u];
    }
}

// follower launch shared by the de-esser and the dynamic EQ: chunking from the slower of the two time constants
static int launch_envelope(mm_ctx* c, const mm_geom* g, EnvArgs& A, double attack_ms, double release_ms, const char* name) {
    const int rows = g->tracks * g->channels;
    A.n = g->n; A.stride = g->stride; A.rows = rows;
    // pipeline.py:509-518: coef = exp(-1 / max(1e-6, sr * t)); Python floats, float32 state
    const double atk = std::exp(-1.0 / std::max(1e-6, (double)g->sr * (attack_ms / 1000.0)));
    const double rel = std::exp(-1.0 / std::max(1e-6, (double)g->sr * (release_ms / 1000.0)));
    A.atk = (float)atk; A.one_m_atk = (float)(1.0 - atk);
    A.rel = (float)rel; A.one_m_rel = (float)(1.0 - rel);
    A.atk_lo = (float)(atk - (double)A.atk); A.rel_lo = (float)(rel - (double)A.rel);
    A.d_atk = atk; A.d_1matk = 1.0 - atk; A.d_rel = rel; A.d_1mrel = 1.0 - rel;
    const double slow = std::max(atk, rel);
    const long long nceil = ((g->n + 31) / 32) * 32;
    long long halo = slow < 1.0 ? (long long)std::ceil(17.5 / -std::log(slow)) : nceil;
    halo = std::min<long long>(((halo + 31) / 32) * 32, nceil);
    A.halo = halo;
    // chunk = halo / 2 (three times the work, short critical path) until the grid exceeds ~14 warps per SM.  Measured at the
    // bench size (972 one-warp CTAs): ring depth 8 (32 KB per CTA, 6 CTAs/SM = 888 slots) left a tail wave, 6.5 ms; depth 4
    // (12 CTAs/SM, one wave) 4.85 ms; chunk = halo / 4 (five times the work, twice the warps) 7.9 ms
    static const int chunk_q = [] { const char* e = getenv("MM_ENV_CHUNK_Q"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 2; }();   // chunk = halo * chunk_q / 4
    long long chunk = std::max<long long>(((halo * chunk_q / 4 + 31) / 32) * 32, 4096);
    while ((long long)rows * ((g->n + chunk - 1) / chunk) > 148LL * 14 * 32 && chunk < nceil) chunk *= 2;
    A.chunk = chunk;
    A.nchunks = (int)((g->n + A.chunk - 1) / A.chunk);
    const long long total = (long long)rows * A.nchunks;
    KernelScope ks(c, name);
    const unsigned nblk = (unsigned)((total + kEnvThreads - 1) / kEnvThreads);
    if (A.mode) envelope_gain_kernel<1><<<nblk, kEnvThreads, 0, c->stream>>>(A);
    else envelope_gain_kernel<0><<<nblk, kEnvThreads, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

// out = x - band + band * gain (float32, numpy's order), optional final clip (pipeline.py:1690, :1696)
__global__ void __launch_bounds__(256) dyneq_apply_kernel(const float* x, const float* band, const float* gain, float* out, long long n,
                                                         long long stride, int clip) {
    const size_t ro = (size_t)blockIdx.y * (size_t)stride + kLead;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 3 < n) {
        const float4 xv = *reinterpret_cast<const float4*>(x + ro + i), bv = *reinterpret_cast<const float4*>(band + ro + i);
        const float4 gv = *reinterpret_cast<const float4*>(gain + ro + i);
        float4 r;
        r.x = __fadd_rn(__fsub_rn(xv.x, bv.x), __fmul_rn(bv.x, gv.x));
        r.y = __fadd_rn(__fsub_rn(xv.y, bv.y), __fmul_rn(bv.y, gv.y));
        r.z = __fadd_rn(__fsub_rn(xv.z, bv.z), __fmul_rn(bv.z, gv.z));
        r.w = __fadd_rn(__fsub_rn(xv.w, bv.w), __fmul_rn(bv.w, gv.w));
        if (clip) {
            r.x = fminf(fmaxf(r.x, -1.f), 1.f); r.y = fminf(fmaxf(r.y, -1.f), 1.f);
            r.z = fminf(fmaxf(r.z, -1.f), 1.f); r.w = fminf(fmaxf(r.w, -1.f), 1.f);
        }
        *reinterpret_cast<float4*>(out + ro + i) = r;
    } else {
        for (int c = 0; c < 4 && i + c < n; ++c) {
            const float xs = x[ro + i + c], b = band[ro + i + c];
            float r = __fadd_rn(__fsub_rn(xs, b), __fmul_rn(b, gain[ro + i + c]));
            if (clip) r = fminf(fmaxf(r, -1.f), 1.f);
            out[ro + i + c] = r;
        }
    }
}

__global__ void __launch_bounds__(256) clip_rows_kernel(const float* in, float* out, long long n, long long stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const size_t o = (size_t)blockIdx.y * (size_t)stride + kLead + i;
        out[o] = fminf(fmaxf(in[o], -1.f), 1.f);
    }
}

// band = (float)(kk * ((double)x - sub[row])): what the reference's filter call returns for a band whose iirpeak section
// degenerates to b = k [1, 0, -1], a = [1, ~0, ~-1] (H(z) = k; see st_dynamic_eq)
__global__ void __launch_bounds__(256) dyneq_affine_band_kernel(const float* x, float* band, long long n, long long stride, double kk,
                                                               const double* sub) {
    const size_t ro = (size_t)blockIdx.y * (size_t)stride + kLead;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) band[ro + i] = (float)(kk * ((double)x[ro + i] - (sub ? sub[blockIdx.y] : 0.0)));
}
// last sample of scipy's odd extension (padlen 9): 2 x[n-1] - x[n-10], per row, float64
__global__ void dyneq_right_end_kernel(const float* x, long long n, long long stride, int rows, double* sub) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        const float* p = x + (size_t)r * (size_t)stride + kLead;
        sub[r] = 2.0 * (double)p[n - 1] - (double)p[n - 10];
    }
}
// flag[row] = 1 when the row's first `cnt` samples are not all equal (a constant row never excites an unstable section)
__global__ void __launch_bounds__(256) dyneq_activity_kernel(const float* x, long long cnt, long long stride, int* flag) {
    const float* p = x + (size_t)blockIdx.y * (size_t)stride + kLead;
    const float x0 = p[0];
    bool any = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) any |= p[i] != x0;
    if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(flag + blockIdx.y, 1);
}

// apply_dynamic_eq (backend/app/pipeline.py:1628-1700): per band a zero-phase peaking section (scipy.signal.iirpeak with the
// reference's arguments), the attack/release follower of |band|, a downward gain above the threshold, x - band + band * g;
// bands run one after the other on the running signal.  params[b] = {w0, bw, threshold_db, ratio, attack_ms, release_ms,
// max_cut_db} with w0, bw already clipped as the reference does (:1657-1658).
//
// The reference hands iirpeak a *bandwidth* in its Q slot (:1661-1663), so the section's -3 dB width is pi * q radians whatever
// the frequency, and what `_safe_filtfilt` (:36-52) returns falls into one of five classes, decided here from (b, a) exactly as
// scipy 1.18 computes them (double arithmetic, libm tan / cos) and reported through `classes`:
//   MM_DYNEQ_STABLE (0)    poles inside the unit circle: the device path (float64 zero-phase section + follower)
//   MM_DYNEQ_OVERFLOW (1)  a pole outside the unit circle and a signal long / lively enough that the forward lfilter pass must
//                          leave float32 range long before the end: every sample of filtfilt's output is NaN, +-Inf, or beyond
//                          float32 -> nan_to_num zeroes the whole band (:1677) -> x - 0 + 0 * g = x: the band is the identity
//                          (every default band at 44.1 kHz, six of eight at 48 kHz)
//   MM_DYNEQ_LFILTER (2)   q = 1 below ~0.35 nyq: a = [1, -1.2e-16 cos, -(1 - 1.1e-16)] sums to 0.0 in float64, lfilter_zi
//                          raises ValueError, the reference falls back to the causal lfilter whose transfer function is
//                          b0 (1 - z^-2) / (1 - z^-2) = b0: band = b0 x (the 1e-16 feedback term moves it by < 1e-12)
//   MM_DYNEQ_MARGINAL (3)  same degenerate section but sum(a) != 0 (bw clipped to 0.5 at w0 = 0.5, or q = 1 above 0.35 nyq):
//                          filtfilt runs with poles at +-(1 - 5.5e-17); started from zi * ext[0] = -b0 ext[0] [1, 1] the forward
//                          pass returns b0 (ext - ext[0]), the backward pass b0^2 (ext - ext[last]):
//                          band = b0^2 (x - (2 x[n-1] - x[n-10]))
//   MM_DYNEQ_SKIPPED (4)   unstable but overflow is not certain (short or constant signal, pole barely outside): the reference's
//                          output is rounding-noise-seeded garbage and cannot be a parity target; the band is passed through
//                          (flags & MM_DYNEQ_STRICT: refused by name with return code 3)
int dyneq_band_kind(const Ba& ba, double* rmax_out) {
    const double a1 = ba.a[1], a2 = ba.a[2];
    const bool stable = std::fabs(a2) < 1.0 && std::fabs(a1) < 1.0 + a2;
    const bool degenerate = std::fabs(a1) <= 1e-12 && std::fabs(a2 + 1.0) <= 1e-12;     // H(z) = b0
    volatile double sum_a = 1.0 + a1;            // numpy's left-to-right float64 sum of [1, a1, a2] (lfilter_zi's pole-at-1 test)
    sum_a = sum_a + a2;
    const double disc = a1 * a1 - 4.0 * a2;
    if (rmax_out)
        *rmax_out = disc >= 0.0 ? std::max(std::fabs((-a1 + std::sqrt(disc)) / 2.0), std::fabs((-a1 - std::sqrt(disc)) / 2.0))
                                : std::sqrt(std::fabs(a2));
    if (degenerate) return sum_a == 0.0 ? MM_DYNEQ_LFILTER : MM_DYNEQ_MARGINAL;
    if (stable && sum_a != 0.0) return MM_DYNEQ_STABLE;
    return -1;
}

int st_dynamic_eq(mm_ctx* c, const mm_geom* g, const float* in, float* out, int nbands, const double* params, unsigned flags,
                  int* classes) {
    const int rows = g->tracks * g->channels;
    std::vector<const FilterPlan*> plans((size_t)nbands, nullptr);
    std::vector<int> cls((size_t)nbands, MM_DYNEQ_STABLE);
    std::vector<double> kk((size_t)nbands, 1.0);
    int activity_known = 0;                      // 0 unknown, 1 every row lively, -1 some row constant
    for (int b = 0; b < nbands; ++b) {
        const double* q = params + 7 * b;
        Ba ba;
        if (!iirpeak(q[0], q[1], &ba)) { set_error("apply_dynamic_eq: band %d: iirpeak(%g, %g) is not a valid design", b, q[0], q[1]); return 3; }
        const double a1 = ba.a[1], a2 = ba.a[2];
        double rmax = 0.0;
        const int kind = dyneq_band_kind(ba, &rmax);
        if (kind == MM_DYNEQ_LFILTER || kind == MM_DYNEQ_MARGINAL) {
            if (g->n <= 9 && kind == MM_DYNEQ_MARGINAL) { set_error("apply_dynamic_eq: %lld frames is not longer than filtfilt's padlen 9", (long long)g->n); return 1; }
            cls[b] = kind;
            kk[b] = kind == MM_DYNEQ_LFILTER ? ba.b[0] : ba.b[0] * ba.b[0];
        } else if (kind == MM_DYNEQ_STABLE) {
            const FilterPlan* p = get_plan(c, ba, PREC_F64);
            if (!p) return 1;
            if (g->n <= p->pad) { set_error("apply_dynamic_eq: %lld frames is not longer than filtfilt's padlen %d", (long long)g->n, p->pad); return 1; }
            plans[b] = p;
        } else {
            // largest pole radius; the unstable mode excited by the first change of the signal (>= one float32 denormal, 1e-45)
            // must pass 1e48 (float32 range times 1e10 for the zero crossings of a complex pair) before the forward pass ends
            bool certain = false;
            if (rmax > 1.0 + 1e-9) {
                const double need = std::ceil((48.0 + 45.0 + 12.0) * std::log(10.0) / std::log(rmax)) + 64.0;   // + 1e12 of residue / margin
                if ((double)g->n > 2.0 * need) {
                    if (activity_known == 0) {
                        int* flag;
                        MM_TRY(arena(c, SL_MISC, (size_t)rows, &flag));
                        MM_CUDA(cudaMemsetAsync(flag, 0, (size_t)rows * sizeof(int), c->stream));
                        // lively within the first half of the signal covers every band that passes the 2 * need test
                        const long long cnt = g->n / 2;
                        dyneq_activity_kernel<<<dim3((unsigned)std::min<long long>((cnt + 255) / 256, 512), (unsigned)rows), 256, 0, c->stream>>>(in, cnt, g->stride, flag);
                        MM_CUDA(cudaGetLastError());
                        std::vector<int> h((size_t)rows);
                        MM_CUDA(cudaMemcpyAsync(h.data(), flag, (size_t)rows * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                        MM_CUDA(cudaStreamSynchronize(c->stream));
                        activity_known = 1;
                        for (int r = 0; r < rows; ++r) if (!h[r]) activity_known = -1;
                    }
                    certain = activity_known == 1;
                }
            }
            if (certain) cls[b] = MM_DYNEQ_OVERFLOW;
            else if (flags & MM_DYNEQ_STRICT) {
                set_error("apply_dynamic_eq: band %d: iirpeak(w0=%g, Q=%g) is an unstable section (a = [1, %g, %g], pole radius %g) whose "
                          "overflow over %lld frames is not certain; the reference passes a bandwidth where scipy expects Q "
                          "(pipeline.py:1661-1663)", b, q[0], q[1], a1, a2, rmax, (long long)g->n);
                return 3;
            } else cls[b] = MM_DYNEQ_SKIPPED;
        }
    }
    if (classes) for (int b = 0; b < nbands; ++b) classes[b] = cls[b];
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    float* sc = B.T[1];
    float* gain = B.T[2];
    const float* cur = in;
    int last = -1;                                   // the last band that touches the samples carries the closing clip
    for (int b = 0; b < nbands; ++b) if (cls[b] != MM_DYNEQ_OVERFLOW && cls[b] != MM_DYNEQ_SKIPPED) last = b;
    for (int b = 0; b < nbands; ++b) {
        if (cls[b] == MM_DYNEQ_OVERFLOW || cls[b] == MM_DYNEQ_SKIPPED) continue;
        const double* q = params + 7 * b;
        if (cls[b] == MM_DYNEQ_STABLE) {
            const FilterPlan* p[1] = {plans[b]};
            const float* i1[1] = {cur};
            float* o1[1] = {B.E[0]};
            Pro none;
            MM_TRY(sweep_fwd(c, g, 1, 1, p, i1, o1, none, plans[b]->pad));
            const float* i2[1] = {B.E[0]};
            float* o2[1] = {sc};
            Epi store;
            MM_TRY(sweep_bwd(c, g, 1, p, i2, o2, 1, store, plans[b]->pad));
        } else {
            double* sub = nullptr;
            if (cls[b] == MM_DYNEQ_MARGINAL) {
                MM_TRY(arena(c, SL_XCHG, (size_t)rows, &sub));
                dyneq_right_end_kernel<<<(rows + 127) / 128, 128, 0, c->stream>>>(cur, g->n, g->stride, rows, sub);
                MM_CUDA(cudaGetLastError());
            }
            KernelScope ks(c, "dyneq_degenerate_band");
            dyneq_affine_band_kernel<<<dim3((unsigned)((g->n + 255) / 256), (unsigned)rows), 256, 0, c->stream>>>(cur, sc, g->n, g->stride, kk[b], sub);
            MM_CUDA(cudaGetLastError());
        }
        EnvArgs A;
        memset(&A, 0, sizeof(A));
        A.sc = sc; A.gain = gain; A.mode = 1;
        A.thr = (float)std::pow(10.0, q[2] / 20.0);
        A.ratio = (float)q[3];
        A.max_cut = (float)std::pow(10.0, q[6] / 20.0);
        MM_TRY(launch_envelope(c, g, A, q[4], q[5], "dyneq_envelope_gain"));
        KernelScope ks(c, "dyneq_apply");
        dyneq_apply_kernel<<<dim3((unsigned)((g->n + 1023) / 1024), (unsigned)rows), 256, 0, c->stream>>>(cur, sc, gain, out, g->n, g->stride,
                                                                                                         b == last);
        MM_CUDA(cudaGetLastError());
        cur = out;
    }
    if (last < 0) {                                  // no band touched the samples: clip(input) (pipeline.py:1696)
        KernelScope ks(c, "dyneq_clip");
        clip_rows_kernel<<<dim3((unsigned)((g->n + 255) / 256), (unsigned)rows), 256, 0, c->stream>>>(in, out, g->n, g->stride);
        MM_CUDA(cudaGetLastError());
    }
    return 0;
}

int st_deesser(mm_ctx* c, const mm_geom* g, const float* in, float* out, double threshold_db, double ratio, double freq_lo,
               double freq_hi, double attack_ms, double release_ms) {
    const double nyq = g->sr / 2.0;
    const double lo = std::min(freq_lo / nyq, 0.97), hi = std::min(freq_hi / nyq, 0.97);
    const size_t bytes = (size_t)g->tracks * g->channels * (size_t)g->stride * sizeof(float);
    if (lo >= hi) {   // pipeline.py:1227-1228: unchanged
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    const FilterPlan* bp = plan_butter(c, 2, kBand, lo, hi);
    if (!bp) return 1;
    if (g->n <= bp->pad) { set_error("apply_deesser: %lld frames is not longer than filtfilt's padlen %d", (long long)g->n, bp->pad); return 1; }
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    float* sc = B.T[1];
    float* gain = B.T[2];
    {
        const FilterPlan* p[1] = {bp};
        const float* i1[1] = {in};
        float* o1[1] = {B.E[0]};
        Pro none;
        MM_TRY(sweep_fwd(c, g, 1, 1, p, i1, o1, none, bp->pad));
        const float* i2[1] = {B.E[0]};
        float* o2[1] = {sc};
        Epi store;
        MM_TRY(sweep_bwd(c, g, 1, p, i2, o2, 1, store, bp->pad));
    }
    const int rows = g->tracks * g->channels;
    {
        EnvArgs A;
        memset(&A, 0, sizeof(A));
        A.sc = sc; A.gain = gain;
        A.thr = (float)std::pow(10.0, threshold_db / 20.0);
        A.inv_ratio = (float)(1.0 / ratio);
        MM_TRY(launch_envelope(c, g, A, attack_ms, release_ms, "envelope_gain"));
    }
    {
        ApplyArgs A;
        A.x = in; A.sc = sc; A.gain = gain; A.out = out; A.n = g->n; A.stride = g->stride;
        int k = std::max(3, (int)((double)g->sr * 0.0015));   // pipeline.py:1256-1258
        k += 1 - (k % 2);
        if (k / 2 > kApplyMaxHalf) { set_error("apply_deesser: smoothing kernel %d too long for this build", k); return 1; }
        A.k = k;
        A.kerf = (float)(1.0 / (double)k);
        // np.ones(k, float32) / float(k): float32 array / Python float -> float32 division
        A.kerf = 1.0f / (float)k;
        dim3 grid((unsigned)((g->n + kApplyTile - 1) / kApplyTile), (unsigned)rows);
        KernelScope ks(c, "deesser_smooth_apply");
        deesser_apply_kernel<<<grid, kApplyThreads, 0, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace mm
