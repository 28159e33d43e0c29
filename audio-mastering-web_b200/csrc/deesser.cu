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

#include "context.h"
#include "stages_internal.h"

namespace mm {

struct EnvArgs {
    const float* sc;       // planar side-chain rows
    float* gain;           // planar pre-smoothing gain rows (same geometry)
    long long n, stride;
    int rows;
    long long chunk, halo; // multiples of 4
    int nchunks;
    float atk, one_m_atk, rel, one_m_rel;
    float thr, ratio;
};

__device__ __forceinline__ float deess_gain(float env, float thr, float ratio) {
    // pipeline.py:1247-1252 in float32 (thr / ratio are Python floats -> weak -> float32 arithmetic)
    const float red = env > thr ? __fadd_rn(thr, __fdiv_rn(__fsub_rn(env, thr), ratio)) : env;
    float g = env > 1e-10f ? __fdiv_rn(red, __fadd_rn(env, 1e-12f)) : 1.0f;
    return fminf(fmaxf(g, 0.35f), 1.0f);
}

__device__ __forceinline__ float env_step(float e, float v, const EnvArgs& P) {
    const float a = __fmaf_rn(P.atk, e, __fmul_rn(P.one_m_atk, v));
    const float r = __fmaf_rn(P.rel, e, __fmul_rn(P.one_m_rel, v));
    return fmaxf(a, r);
}

// one thread = one chunk of one row; rows in blockIdx.y
__global__ void __launch_bounds__(128) envelope_gain_kernel(const EnvArgs P) {
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= P.nchunks) return;
    const int row = blockIdx.y;
    const float* src = P.sc + (size_t)row * (size_t)P.stride + kLead;
    float* dst = P.gain + (size_t)row * (size_t)P.stride + kLead;
    const long long live0 = (long long)chunk * P.chunk;
    const long long live1 = min(live0 + P.chunk, P.n);
    long long i = max(live0 - P.halo, 0LL);
    float e = fabsf(src[i]);                 // env[0] = |v0| (exact for i == 0, start-up guess otherwise)
    // warm-up over the halo: no stores
    for (; i + 3 < live0; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        if (i != 0) e = env_step(e, fabsf(v.x), P);
        e = env_step(e, fabsf(v.y), P);
        e = env_step(e, fabsf(v.z), P);
        e = env_step(e, fabsf(v.w), P);
    }
    // live part (live0 is a multiple of 4, so is i here)
#pragma unroll 2
    for (; i + 3 < live1; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        float4 g;
        if (i != 0) e = env_step(e, fabsf(v.x), P);
        g.x = deess_gain(e, P.thr, P.ratio);
        e = env_step(e, fabsf(v.y), P);
        g.y = deess_gain(e, P.thr, P.ratio);
        e = env_step(e, fabsf(v.z), P);
        g.z = deess_gain(e, P.thr, P.ratio);
        e = env_step(e, fabsf(v.w), P);
        g.w = deess_gain(e, P.thr, P.ratio);
        *reinterpret_cast<float4*>(dst + i) = g;
    }
    for (; i < live1; ++i) {
        if (i != 0) e = env_step(e, fabsf(src[i]), P);
        dst[i] = deess_gain(e, P.thr, P.ratio);
    }
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
constexpr int kApplyThreads = 128;
constexpr int kApplyPer = 16;                                  // outputs per thread
constexpr int kApplyTile = kApplyThreads * kApplyPer;          // 2048 outputs per CTA
constexpr int kApplyMaxHalf = 256;                             // supports k <= 513 (sr <= 342 kHz)

__global__ void __launch_bounds__(kApplyThreads) deesser_apply_kernel(const ApplyArgs P) {
    __shared__ float sg[kApplyTile + 2 * kApplyMaxHalf];
    const int row = blockIdx.y;
    const long long base = (long long)blockIdx.x * kApplyTile;
    const int half = P.k / 2;
    const size_t ro = (size_t)row * (size_t)P.stride + kLead;
    const int span = kApplyTile + 2 * half;
    for (int j = threadIdx.x; j < span; j += kApplyThreads) {
        const long long i = base - half + j;
        sg[j] = (i >= 0 && i < P.n) ? P.gain[ro + i] : 0.f;
    }
    __syncthreads();
    const int o0 = threadIdx.x * kApplyPer;
    double acc = 0.0;
    for (int j = 0; j < P.k; ++j) acc += (double)sg[o0 + j];
#pragma unroll
    for (int u = 0; u < kApplyPer; ++u) {
        const long long i = base + o0 + u;
        if (i < P.n) {
            float g = (float)(acc * (double)P.kerf);
            g = fminf(fmaxf(g, 0.35f), 1.0f);
            const float x = P.x[ro + i], s = P.sc[ro + i];
            P.out[ro + i] = __fadd_rn(__fsub_rn(x, s), __fmul_rn(s, g));
        }
        acc += (double)sg[o0 + u + P.k] - (double)sg[o0 + u];
    }
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
        A.sc = sc; A.gain = gain; A.n = g->n; A.stride = g->stride; A.rows = rows;
        // pipeline.py:509-518: coef = exp(-1 / max(1e-6, sr * t)); Python floats, float32 state
        const double atk = std::exp(-1.0 / std::max(1e-6, (double)g->sr * (attack_ms / 1000.0)));
        const double rel = std::exp(-1.0 / std::max(1e-6, (double)g->sr * (release_ms / 1000.0)));
        A.atk = (float)atk; A.one_m_atk = (float)(1.0 - atk);
        A.rel = (float)rel; A.one_m_rel = (float)(1.0 - rel);
        A.thr = (float)std::pow(10.0, threshold_db / 20.0);
        A.ratio = (float)ratio;
        const double slow = std::max(atk, rel);
        long long halo = slow < 1.0 ? (long long)std::ceil(17.5 / -std::log(slow)) : g->n;
        halo = std::min<long long>(((halo + 3) / 4) * 4, ((g->n + 3) / 4) * 4);
        A.halo = halo;
        A.chunk = std::max<long long>(halo, 4096);
        A.nchunks = (int)((g->n + A.chunk - 1) / A.chunk);
        dim3 grid((unsigned)((A.nchunks + 127) / 128), (unsigned)rows);
        KernelScope ks(c, "envelope_gain");
        envelope_gain_kernel<<<grid, 128, 0, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
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
