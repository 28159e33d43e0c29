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
};

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
    double w[NF], wc, trim;
    DynParams dyn;
    double exc_gain, exc_k;
    int exc_mode;
    float* peak;             // per track |out| max (float bits, atomicMax) or null
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
