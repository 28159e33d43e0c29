// Shared definitions for the mastering kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

// ---- tile geometry -------------------------------------------------------------------------
// One CTA = one tile of one row: kT threads, each owning kS consecutive samples (in sweep order).
constexpr int kS = 32;                   // samples per thread
constexpr int kT = 128;                  // threads per tile
constexpr int kNW = kT / 32;             // warps per tile
constexpr int kL = kS * kT;              // samples per tile (4096)
constexpr int kLead = 32;                // == MM_LEAD: float offset of sample 0 inside a row
// Backward sweeps count their tiles down from `qend`, one past the last position of tile 0.  Rounding qend up to a whole 128-byte
// line (32 floats) makes every backward tile origin line aligned -- bulk tensor loads and the coalesced stores then move whole
// lines (with the minimal rounding, to 4 floats, a tile's 128-byte rows each straddled two lines); tile 0 just starts with up to
// 31 dead samples instead of up to 3 (dead < kS either way).
#ifndef MM_BWD_ALIGN
#define MM_BWD_ALIGN 32
#endif
__host__ __device__ inline long long bwd_qend(long long q_last) { return (q_last + MM_BWD_ALIGN) & ~(long long)(MM_BWD_ALIGN - 1); }

// Offsets (in doubles) inside a per-filter device table, see design.h ScanTables.
template <int M> struct Tab {
    static constexpr int MM = M * M;
    static constexpr int Pw = 0;                         // [5][MM]
    static constexpr int Plane = Pw + 5 * MM;            // [32][MM]
    static constexpr int Qpow = Plane + 32 * MM;         // [kNW+1][MM]   Qpow[kNW] = A^kL
    static constexpr int Zi = Qpow + (kNW + 1) * MM;     // [M]
    static constexpr int Apow = Zi + M;                  // [kS+1][MM]
    static constexpr int Mpow = Apow + (kS + 1) * MM;    // [W][MM]  (A^kL)^j, kept for host-side verification
};

// Per-filter constants that ride in the kernel parameter block (constant bank): the compiler folds
// them straight into DFMA operands because every index below is a compile-time constant.
template <int M> struct FiltK {
    double b[M + 1];
    double a[M];          // a[1..M]
    double g[kS][M];      // g[j] = A^(kS-1-j) B   (zero-state end state = sum_j g[j] x_j)
    // float32 pass 2: the balanced realization of the section with its states rescaled by d_i = 1 / B_i, so that B = (1, .., 1)
    // (floating-point round-off is scale invariant, the realization stays as well conditioned as the balanced one):
    //   s' = A s + (x, .., x),  y = C s + D x      -- 7 operations per sample for a biquad instead of 9
    // dn32 maps a state in the tables' (balanced) coordinates into the rescaled ones
    float A32[M][M], dn32[M], C32[M], D32;
};

// Two float32 sections evaluated in lock step with Blackwell's packed FFMA2 (fma.rn.f32x2): lane .x belongs to
// the first section of the pair, .y to the second.  Same realization as FiltK::A32.. (balanced coordinates).
struct PairK {
    float2 A[2][2], B[2], C[2], D;
    float2 g[kS][2];      // float32 pass 1: g[j][i] = (A^(kS-1-j) B)_i of both sections
};
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
// one packed step of a pair of 2-state sections: returns (y_a, y_b), advances (s0, s1)
__device__ __forceinline__ float2 pair_step(const PairK& k, float2 x, float2& s0, float2& s1) {
    const float2 y = ffma2(k.C[0], s0, ffma2(k.C[1], s1, fmul2(k.D, x)));
    const float2 n0 = ffma2(k.A[0][0], s0, ffma2(k.A[0][1], s1, fmul2(k.B[0], x)));
    const float2 n1 = ffma2(k.A[1][0], s0, ffma2(k.A[1][1], s1, fmul2(k.B[1], x)));
    s0 = n0;
    s1 = n1;
    return y;
}

// The same step with the two lanes' inputs given as scalars: the input-dependent products (D x, B x) are scalar FMULs written
// straight into the halves of the accumulator pairs, so no register moves are needed to pack (xa, xb) -- 6 FMUL + 6 FFMA2 where
// packing cost 9 packed instructions plus ~10 moves per step (SASS of the round-1 loudness kernel).  Same association, same bits.
__device__ __forceinline__ float2 pair_step_xy(const PairK& k, float xa, float xb, float2& s0, float2& s1) {
    float2 t, m0, m1;
    t.x = __fmul_rn(k.D.x, xa);     t.y = __fmul_rn(k.D.y, xb);
    m0.x = __fmul_rn(k.B[0].x, xa); m0.y = __fmul_rn(k.B[0].y, xb);
    m1.x = __fmul_rn(k.B[1].x, xa); m1.y = __fmul_rn(k.B[1].y, xb);
    const float2 y = ffma2(k.C[0], s0, ffma2(k.C[1], s1, t));
    const float2 n0 = ffma2(k.A[0][0], s0, ffma2(k.A[0][1], s1, m0));
    const float2 n1 = ffma2(k.A[1][0], s0, ffma2(k.A[1][1], s1, m1));
    s0 = n0;
    s1 = n1;
    return y;
}

// prologue applied to samples as they are loaded
enum { PRO_NONE = 0, PRO_SUBMUL_F32 = 1, PRO_MUL_F64 = 2 };
// epilogue of a sweep
enum { EPI_STORE = 0, EPI_COMBINE = 1, EPI_DYNAMICS = 2, EPI_EXCITER = 3, EPI_DYNAMICS_GEN = 4 };   // _GEN: upward bands and/or parallel mix

struct DynBand {           // one band of MULTIBAND_CONFIG after host-side preparation
    double thr_db, thr, ratio, lower, upper, slope, max_boost_db;
    float lim, gain;
    float s_mid, c_mid, s_hi, c_hi;   // float32 lines of the downward knee: knee segment and above-knee segment
    int mode;              // 0 bypass (ratio == 1 or <= 0), 1 hard knee, 2 soft knee, 3 upward
};
struct DynParams {
    DynBand band[4];
    float max_k, max_c, max_top;                  // maximizer line k |s| + c (pipeline.py:484-492) and min(ceil, TRUE_PEAK_LIMIT) cap
    // optional parallel compression folded behind the limiter (v1, pipeline.py:1771-1797)
    const double* par_mix;                        // per-row mix (device) or nullptr
    float par_slope, par_cmid, par_shi, par_chi;  // knee / above-knee lines of the parallel compressor
};

template <int M, int NF> struct SweepArgs {
    FiltK<M> f[NF];
    const double* tab[NF];
    int W[NF];
    const float* in[NF];
    float* out[NF];
    const float* aux[2];
    long long n, stride;
    int rows, ntiles, pad, channels;
    int pro_mode;
    const double* pro_sub;   // per row (may be null)
    const double* pro_mul;   // per row (may be null)
    int aux_pro;             // apply the prologue to aux[0] too
    int epi;
    int epi_clip;            // EPI_COMBINE: clip the result to +-1
    double w[NF], wc, trim;
    float w32[NF], wc32, trim32;
    PairK pr[(NF + 1) / 2];  // packed float32 sections (pairs 2p, 2p+1 with 2p+1 < NF32)
    DynParams dyn;
    double exc_gain, exc_k;
    int exc_mode;
    float* peak;             // per track |out| max (float bits, atomicMax) or null
    const int* row_map;      // optional: the sweep visits rows row_map[0 .. rows) of the batch (tracks whose style fires this stage)
    // per-row parameters of a mixed-preset batch (indexed by the row's position in the batch; null = the scalars above)
    const double* w_row;     // EPI_COMBINE with one section: recombination weight (a style-EQ band's 10^(dB/20) - 1)
    const double* exc_row;   // EPI_EXCITER: 10^(dB/20) - 1
    const unsigned char* peak_row;   // the row's output counts towards its track's peak (this is the last stage that touches it)
    long long pk_lo, pk_hi;  // row positions (inclusive) whose outputs count towards the peak (a time slice's own frames)
};

__device__ __forceinline__ double shfl_up_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, d);
    hi = __shfl_up_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, d);
    hi = __shfl_xor_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ float comp4(const float4& v, int c) {
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}
__device__ __forceinline__ void setcomp4(float4& v, int c, float x) {
    if (c == 0) v.x = x; else if (c == 1) v.y = x; else if (c == 2) v.z = x; else v.w = x;
}

// One DF2T step: y = b0 x + z0 ; z_i = b_{i+1} x - a_{i+1} y + z_{i+1}
template <int M> __device__ __forceinline__ double df2t_step(const FiltK<M>& fk, double x, double (&z)[M]) {
    const double y = fma(fk.b[0], x, z[0]);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double t = (i + 1 < M) ? z[i + 1] : 0.0;
        t = fma(fk.b[i + 1], x, t);
        z[i] = fma(-fk.a[i], y, t);
    }
    return y;
}

// One float32 step of a section in its (balanced) state-space form.  The start state of every 32-sample chunk
// comes from the float64 scan, so round-off only accumulates inside a chunk.
template <int M> __device__ __forceinline__ float ss32_step(const FiltK<M>& fk, float x, float (&s)[M]) {
    float y = fk.D32 * x;
#pragma unroll
    for (int i = M - 1; i >= 0; --i) y = fmaf(fk.C32[i], s[i], y);
    float sn[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        float t = x;                                                    // B = 1 in the rescaled coordinates
#pragma unroll
        for (int k = M - 1; k >= 0; --k) t = fmaf(fk.A32[i][k], s[k], t);
        sn[i] = t;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = sn[i];
    return y;
}

// y = Mat(MxM, row-major at p) * v, added into acc
template <int M> __device__ __forceinline__ void matvec_acc(const double* __restrict__ p, const double (&v)[M], double (&acc)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double s = acc[i];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(__ldg(p + i * M + k), v[k], s);
        acc[i] = s;
    }
}

}  // namespace mm
