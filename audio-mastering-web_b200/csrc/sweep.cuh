// Chunked linear-recurrence scan over one tile (kL samples of one row) for NF IIR sections.
//
// Work decomposition (the whole file is the "kernel (1)" of BASELINE.json's north_star):
//   thread  : kS consecutive samples (sweep order); zero-state end state by a dot product with
//             g[j] = A^(kS-1-j) B (no dependency chain), later the true DF2T recurrence from the
//             resolved incoming state
//   warp    : Kogge-Stone scan of the 2x2 (or 4x4) state-transfer recurrence with shuffles,
//             multipliers (A^kS)^(2^d) precomputed in float64
//   tile    : Horner combine of the kNW warp totals through shared memory
//   grid    : truncated decoupled look-back: tile k adds M^j * aggregate(k-1-j) for j < W, where
//             M = A^kL and W is the first power with |M^W| < 1e-18 (fp64-negligible).  Aggregates
//             are ZERO-STATE end states, so no tile ever waits on a chain of predecessors.
// Tiles are handed out by an atomic ticket, tile-major across rows, so every predecessor of a
// tile has already been claimed by a running CTA (forward progress without co-residency games).
#pragma once
#include "common.cuh"

namespace mm {

template <int M, int NF> struct ScanScratch {
    double tot[NF][kNW][M];
    double carry[NF][M];
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    return *reinterpret_cast<const volatile unsigned*>(p);
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

// Args must expose: f[], tab[], W[], agg, flag, epoch, rows, ntiles, err.
// Filters F0..F0+NF-1 of P are applied; filter f reads shared stream (NIN == 1 ? 0 : f) and
// overwrites shared stream f with its float32 output.
//   row, tile      : this CTA's work item
//   inject         : tile == 0 and the section starts from zi * first_live_sample
//   dead           : number of dead (zero) samples at the head of tile 0 in sweep order
template <int M, int NF, int NIN, int DIR, int F0, class Args>
__device__ __forceinline__ void tile_scan(const Args& P, float* buf, ScanScratch<M, NF>& sh,
                                          int row, int tile, bool inject, int dead) {
    constexpr int MM = M * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (DIR > 0) ? tid : (kT - 1 - tid);
    const float* cb = buf + chunk * kChunk;

    const int t0 = dead / kS, d0 = dead - t0 * kS;
    const bool inj_thread = inject && tid == t0;

    // initial DF2T state zi * x_first (scipy filtfilt / lfilter_zi semantics)
    double s_init[NF][M];
    if (inj_thread) {
        const int s_first = dead;                                   // sweep position of first live sample
        const int mi = (DIR > 0) ? s_first : (kL - 1 - s_first);
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const float x0 = buf[(NIN == 1 ? 0 : f) * kTileFloats + pm(mi)];
#pragma unroll
            for (int i = 0; i < M; ++i) s_init[f][i] = __ldg(P.tab[F0 + f] + Tab<M>::Zi + i) * (double)x0;
        }
    }

    // ---- pass 1: zero-state end state of this thread's chunk ------------------------------------
    double E[NF][M];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int i = 0; i < M; ++i) E[f][i] = 0.0;
#pragma unroll
    for (int u = 0; u < kS / 4; ++u) {
        const int uu = (DIR > 0) ? u : (kS / 4 - 1 - u);
        float4 xv[NIN];
#pragma unroll
        for (int s = 0; s < NIN; ++s) xv[s] = *reinterpret_cast<const float4*>(cb + s * kTileFloats + 4 * uu);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cc = (DIR > 0) ? c : (3 - c);
            const int j = 4 * u + c;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double x = (double)comp4(xv[NIN == 1 ? 0 : f], cc);
#pragma unroll
                for (int i = 0; i < M; ++i) E[f][i] = fma(P.f[F0 + f].g[j][i], x, E[f][i]);
            }
        }
    }
    if (inj_thread) {
#pragma unroll
        for (int f = 0; f < NF; ++f) matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Apow + (kS - d0) * MM, s_init[f], E[f]);
    }

    // ---- warp scan ------------------------------------------------------------------------------
#pragma unroll
    for (int d = 0; d < 5; ++d) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            double pe[M];
#pragma unroll
            for (int i = 0; i < M; ++i) pe[i] = shfl_up_d(E[f][i], 1 << d);
            if (lane >= (1 << d)) matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Pw + d * MM, pe, E[f]);
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int i = 0; i < M; ++i) sh.tot[f][warp][i] = E[f][i];
    }
    __syncthreads();

    // ---- zero-state prefix over warps (Horner in Q = A^(32 kS)) ------------------------------------
    double base[NF][M];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
#pragma unroll
        for (int i = 0; i < M; ++i) base[f][i] = 0.0;
        for (int v = 0; v < warp; ++v) {
            double nb[M];
#pragma unroll
            for (int i = 0; i < M; ++i) nb[i] = sh.tot[f][v][i];
            matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Qpow + MM, base[f], nb);
#pragma unroll
            for (int i = 0; i < M; ++i) base[f][i] = nb[i];
        }
    }

    // ---- publish this tile's zero-state aggregate, then look back ---------------------------------
    const size_t slot0 = ((size_t)F0 * P.rows + row) * (size_t)P.ntiles + tile;   // filter stride = rows*ntiles
    const size_t fstride = (size_t)P.rows * (size_t)P.ntiles;
    if (tid == kT - 1) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            double ag[M];
#pragma unroll
            for (int i = 0; i < M; ++i) ag[i] = E[f][i];                 // lane 31 of last warp: its warp total
            matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Qpow + MM, base[f], ag);
            double* dst = P.agg + (slot0 + f * fstride) * M;
#pragma unroll
            for (int i = 0; i < M; ++i) __stcg(dst + i, ag[i]);
        }
        __threadfence();
#pragma unroll
        for (int f = 0; f < NF; ++f) *reinterpret_cast<volatile unsigned*>(P.flag + slot0 + f * fstride) = P.epoch;
    }
    if (warp < NF) {
        // warp f resolves filter f's carry-in
        int f = warp;
        double C[M];
#pragma unroll
        for (int i = 0; i < M; ++i) C[i] = 0.0;
        const int Wf = P.W[F0 + f];
        for (int j0 = 0; j0 < Wf && j0 < tile; j0 += 32) {
            const int j = j0 + lane;
            double c[M];
#pragma unroll
            for (int i = 0; i < M; ++i) c[i] = 0.0;
            if (j < Wf && j < tile) {
                const size_t slot = slot0 + f * fstride - 1 - j;
                unsigned spins = 0;
                while (ld_volatile_u32(P.flag + slot) != P.epoch) {
                    if (++spins > (1u << 21)) { atomicExch(P.err, 1); break; }
                    __nanosleep(32);
                }
                __threadfence();
                double a[M];
#pragma unroll
                for (int i = 0; i < M; ++i) a[i] = __ldcg(P.agg + slot * M + i);
                matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Mpow + j * MM, a, c);
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double v = c[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
                C[i] += v;
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < M; ++i) sh.carry[f][i] = C[i];
        }
    }
    __syncthreads();

    // ---- incoming state of this thread: J_{l-1} + Plane[l] * (Wx_w + Q^w C) --------------------------
    double z[NF][M];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double C[M];
#pragma unroll
        for (int i = 0; i < M; ++i) C[i] = sh.carry[f][i];
        matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Qpow + warp * MM, C, base[f]);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const double up = shfl_up_d(E[f][i], 1);
            z[f][i] = (lane > 0) ? up : 0.0;
        }
        matvec_acc<M>(P.tab[F0 + f] + Tab<M>::Plane + lane * MM, base[f], z[f]);
    }

    // ---- pass 2: the recurrence proper, float32 results back into shared memory ----------------------
    float* wb = buf + chunk * kChunk;
#pragma unroll
    for (int u = 0; u < kS / 4; ++u) {
        const int uu = (DIR > 0) ? u : (kS / 4 - 1 - u);
        float4 xv[NIN];
#pragma unroll
        for (int s = 0; s < NIN; ++s) xv[s] = *reinterpret_cast<const float4*>(cb + s * kTileFloats + 4 * uu);
        float4 yv[NF];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cc = (DIR > 0) ? c : (3 - c);
            const int j = 4 * u + c;
            if (inj_thread && j == d0) {
#pragma unroll
                for (int f = 0; f < NF; ++f)
#pragma unroll
                    for (int i = 0; i < M; ++i) z[f][i] = s_init[f][i];
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double x = (double)comp4(xv[NIN == 1 ? 0 : f], cc);
                const double y = df2t_step<M>(P.f[F0 + f], x, z[f]);
                setcomp4(yv[f], cc, (float)y);
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) *reinterpret_cast<float4*>(wb + f * kTileFloats + 4 * uu) = yv[f];
    }
}

}  // namespace mm
