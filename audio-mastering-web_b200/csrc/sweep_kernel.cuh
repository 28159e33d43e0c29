// The generic zero-phase/causal IIR sweep kernel: load a tile (with prologue and scipy's odd
// extension), run tile_scan for NF sections, apply the epilogue and store.
#pragma once
#include "pointwise.cuh"
#include "sweep.cuh"

namespace mm {

template <int M, int NF>
__device__ __forceinline__ float apply_prologue(const SweepArgs<M, NF>& P, float x, float subf, float mulf, double muld) {
    if (P.pro_mode == PRO_SUBMUL_F32) return __fmul_rn(__fsub_rn(x, subf), mulf);
    if (P.pro_mode == PRO_MUL_F64) return (float)((double)x * muld);
    return x;
}

// FWD (DIR=+1): inputs are x-domain rows (sample i at q = kLead + i); the sweep covers the odd
// extension q in [kLead - pad, kLead + n + pad) and writes ext-domain rows at the same q.
// BWD (DIR=-1): inputs are ext-domain rows; the sweep starts at q_last = kLead + n + pad - 1 and
// the epilogue writes x-domain samples q in [kLead, kLead + n).
template <int M, int NF, int NIN, int DIR>
__global__ void __launch_bounds__(kT) sweep_kernel(const __grid_constant__ SweepArgs<M, NF> P) {
    extern __shared__ __align__(16) float smem[];
    __shared__ ScanScratch<M, NF> sh;
    __shared__ unsigned s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = atomicAdd(P.ticket, 1u) - P.ticket_base;
    __syncthreads();
    const unsigned ticket = s_ticket;
    const int tile = (int)(ticket / (unsigned)P.rows);
    const int row = (int)(ticket - (unsigned)tile * (unsigned)P.rows);

    const long long q_first = kLead - P.pad;
    const long long q_last = kLead + P.n + P.pad - 1;
    long long tile_lo;
    int dead = 0;
    if (DIR > 0) {
        tile_lo = (long long)tile * kL;
        dead = (int)q_first;
    } else {
        const long long qend = (q_last + 4) & ~3LL;     // roundup4(q_last + 1)
        tile_lo = qend - (long long)(tile + 1) * kL;
        dead = (int)(qend - 1 - q_last);
    }
    const size_t rowoff = (size_t)row * (size_t)P.stride;

    float subf = 0.f, mulf = 1.f;
    double muld = 1.0;
    if (P.pro_mode != PRO_NONE) {
        if (P.pro_sub) subf = (float)__ldg(P.pro_sub + row);
        if (P.pro_mul) { muld = __ldg(P.pro_mul + row); mulf = (float)muld; }
    }

    // ---- load phase ---------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < NIN; ++s) {
        const float* src = P.in[s] + rowoff;
        float* dst = smem + s * kTileFloats;
#pragma unroll
        for (int r = 0; r < kL / (4 * kT); ++r) {
            const int mi = 4 * (tid + kT * r);
            const long long q = tile_lo + mi;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DIR > 0) {
                if (q >= kLead && q + 3 < kLead + P.n) {
                    v = __ldcs(reinterpret_cast<const float4*>(src + q));
                    if (P.pro_mode != PRO_NONE) {
                        v.x = apply_prologue<M, NF>(P, v.x, subf, mulf, muld);
                        v.y = apply_prologue<M, NF>(P, v.y, subf, mulf, muld);
                        v.z = apply_prologue<M, NF>(P, v.z, subf, mulf, muld);
                        v.w = apply_prologue<M, NF>(P, v.w, subf, mulf, muld);
                    }
                } else if (q + 3 >= q_first && q <= q_last) {
                    // edge group: x-domain samples, odd extension, or dead
                    const float x_lo = apply_prologue<M, NF>(P, src[kLead], subf, mulf, muld);
                    const float x_hi = apply_prologue<M, NF>(P, src[kLead + P.n - 1], subf, mulf, muld);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const long long i = q + c - kLead;
                        float e = 0.f;
                        if (i >= 0 && i < P.n) e = apply_prologue<M, NF>(P, src[kLead + i], subf, mulf, muld);
                        else if (i < 0 && i >= -(long long)P.pad)
                            e = __fsub_rn(__fmul_rn(2.f, x_lo), apply_prologue<M, NF>(P, src[kLead - i], subf, mulf, muld));
                        else if (i >= P.n && i < P.n + P.pad)
                            e = __fsub_rn(__fmul_rn(2.f, x_hi), apply_prologue<M, NF>(P, src[kLead + 2 * (P.n - 1) - i], subf, mulf, muld));
                        setcomp4(v, c, e);
                    }
                }
            } else {
                if (q >= q_first && q + 3 <= q_last) {
                    v = __ldcs(reinterpret_cast<const float4*>(src + q));
                } else if (q + 3 >= q_first && q <= q_last) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const long long qq = q + c;
                        setcomp4(v, c, (qq >= q_first && qq <= q_last) ? src[qq] : 0.f);
                    }
                }
            }
            *reinterpret_cast<float4*>(dst + pm(mi)) = v;
        }
    }
    __syncthreads();

    tile_scan<M, NF, NIN, DIR, 0>(P, smem, sh, row, tile, tile == 0 && P.pad > 0, dead);
    __syncthreads();

    // ---- epilogue / store phase ------------------------------------------------------------------------
    const long long st_lo = (DIR > 0) ? q_first : (long long)kLead;
    const long long st_hi = (DIR > 0) ? q_last : (long long)(kLead + P.n - 1);   // inclusive
    float pk = 0.f;
#pragma unroll 2
    for (int r = 0; r < kL / (4 * kT); ++r) {
        const int mi = 4 * (tid + kT * r);
        const long long q = tile_lo + mi;
        if (q + 3 < st_lo || q > st_hi) continue;
        const bool full = (q >= st_lo && q + 3 <= st_hi);
        float4 y[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) y[f] = *reinterpret_cast<const float4*>(smem + f * kTileFloats + pm(mi));
        if (P.epi == EPI_STORE) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float* dst = P.out[f] + rowoff;
                if (full) __stcs(reinterpret_cast<float4*>(dst + q), y[f]);
                else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (q + c >= st_lo && q + c <= st_hi) dst[q + c] = comp4(y[f], c);
                }
            }
        } else {
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, o;
            const float* x0 = P.aux[0] ? P.aux[0] + rowoff : nullptr;
            const float* x1 = P.aux[1] ? P.aux[1] + rowoff : nullptr;
            if (full) {
                if (x0) a0 = __ldcs(reinterpret_cast<const float4*>(x0 + q));
                if (x1) a1 = __ldcs(reinterpret_cast<const float4*>(x1 + q));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (q + c >= st_lo && q + c <= st_hi) {
                        if (x0) setcomp4(a0, c, x0[q + c]);
                        if (x1) setcomp4(a1, c, x1[q + c]);
                    }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float xa = comp4(a0, c);
                if (P.aux_pro) xa = apply_prologue<M, NF>(P, xa, subf, mulf, muld);
                float res;
                if (P.epi == EPI_COMBINE) {
                    double acc = P.wc * (double)xa;
#pragma unroll
                    for (int f = 0; f < NF; ++f) acc += P.w[f] * (double)comp4(y[f], c);
                    res = (float)(acc * P.trim);
                } else if (P.epi == EPI_EXCITER) {
                    const double hf = (double)comp4(y[0], c);
                    const double sat = exciter_sat(hf, P.exc_mode, P.exc_k);
                    res = (float)((double)xa + (sat - hf) * P.exc_gain * 0.25);
                } else {   // EPI_DYNAMICS: aux0 = band 1, y0 = band 2, y1 = band 3, aux1 = band 4
                    float s = band_chain(xa, P.dyn.band[0]);
                    s = __fadd_rn(s, band_chain(comp4(y[0], c), P.dyn.band[1]));
                    s = __fadd_rn(s, band_chain(comp4(y[NF > 1 ? 1 : 0], c), P.dyn.band[2]));
                    s = __fadd_rn(s, band_chain(comp4(a1, c), P.dyn.band[3]));
                    res = maximize_limit(s, P.dyn);
                    if (P.dyn.par_mix) {
                        const double mix = __ldg(P.dyn.par_mix + row);
                        if (mix >= 0.01) res = parallel_compress(res, mix, P.dyn);
                    }
                }
                setcomp4(o, c, res);
                if (q + c >= st_lo && q + c <= st_hi) pk = fmaxf(pk, fabsf(res));
            }
            float* dst = P.out[0] + rowoff;
            if (full) __stcs(reinterpret_cast<float4*>(dst + q), o);
            else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (q + c >= st_lo && q + c <= st_hi) dst[q + c] = comp4(o, c);
            }
        }
    }
    if (P.peak != nullptr && P.epi != EPI_STORE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        if ((tid & 31) == 0 && pk > 0.f) atomicMax(reinterpret_cast<int*>(P.peak + row / P.channels), __float_as_int(pk));
    }
}

}  // namespace mm
