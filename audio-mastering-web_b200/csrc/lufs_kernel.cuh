// ITU-R BS.1770 K-weighting (two causal biquads, zero initial state) fused with the 400 ms block
// mean-square partial sums, followed by the two-pass gating -- pyloudnorm.Meter.integrated_loudness
// as called at backend/app/pipeline.py:646-648 / :660-662.
#pragma once
#include "sweep.cuh"

namespace mm {

struct LufsArgs {
    FiltK<2> f[2];
    const double* tab[2];
    int W[2];
    const float* in;
    long long n, stride;
    int rows, ntiles, channels;
    int pro_mode;
    const double* pro_sub;
    const double* pro_mul;
    // segment bookkeeping: sample i belongs to segment s iff bnd[s] <= i < bnd[s+1]
    const long long* bnd;    // [nseg + 1]
    int nseg;
    const int* tile_seg;     // [ntiles] segment of max(first sample of tile, 0), clamped to nseg
    double* segsum;          // [rows][nseg]
    double* agg;
    unsigned* flag;
    unsigned epoch, ticket_base;
    unsigned* ticket;
    int* err;
};

__global__ void __launch_bounds__(kT) lufs_kernel(const __grid_constant__ LufsArgs P) {
    __shared__ __align__(16) float smem[kTileFloats];
    __shared__ ScanScratch<2, 1> sh;
    __shared__ unsigned s_ticket;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_ticket = atomicAdd(P.ticket, 1u) - P.ticket_base;
    __syncthreads();
    const unsigned ticket = s_ticket;
    const int tile = (int)(ticket / (unsigned)P.rows);
    const int row = (int)(ticket - (unsigned)tile * (unsigned)P.rows);
    const long long tile_lo = (long long)tile * kL;
    const float* src = P.in + (size_t)row * (size_t)P.stride;

    float subf = 0.f, mulf = 1.f;
    double muld = 1.0;
    if (P.pro_mode != PRO_NONE) {
        if (P.pro_sub) subf = (float)__ldg(P.pro_sub + row);
        if (P.pro_mul) { muld = __ldg(P.pro_mul + row); mulf = (float)muld; }
    }
#pragma unroll
    for (int r = 0; r < kL / (4 * kT); ++r) {
        const int mi = 4 * (tid + kT * r);
        const long long q = tile_lo + mi;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q >= kLead && q + 3 < kLead + P.n) {
            v = __ldcs(reinterpret_cast<const float4*>(src + q));
        } else if (q + 3 >= kLead && q < kLead + P.n) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (q + c >= kLead && q + c < kLead + P.n) setcomp4(v, c, src[q + c]);
        }
        if (P.pro_mode == PRO_SUBMUL_F32) {
            v.x = __fmul_rn(__fsub_rn(v.x, subf), mulf); v.y = __fmul_rn(__fsub_rn(v.y, subf), mulf);
            v.z = __fmul_rn(__fsub_rn(v.z, subf), mulf); v.w = __fmul_rn(__fsub_rn(v.w, subf), mulf);
        } else if (P.pro_mode == PRO_MUL_F64) {
            v.x = (float)((double)v.x * muld); v.y = (float)((double)v.y * muld);
            v.z = (float)((double)v.z * muld); v.w = (float)((double)v.w * muld);
        }
        // dead positions must stay exactly zero after the prologue
        if (!(q >= kLead && q + 3 < kLead + P.n)) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (!(q + c >= kLead && q + c < kLead + P.n)) setcomp4(v, c, 0.f);
        }
        *reinterpret_cast<float4*>(smem + pm(mi)) = v;
    }
    __syncthreads();
    // shelf, then high-pass; the float32 round trip between the stages is pyloudnorm's own
    // (it writes each stage back into a copy of the float32 input)
    tile_scan<2, 1, 1, +1, 0>(P, smem, sh, row, tile, false, 0);
    tile_scan<2, 1, 1, +1, 1>(P, smem, sh, row, tile, false, 0);

    // ---- squared sums per segment ------------------------------------------------------------------------
    const long long i0 = tile_lo + (long long)tid * kS - kLead;     // first sample index of this thread
    const long long iw = tile_lo + (long long)(tid & ~31) * kS - kLead;   // first sample of this warp
    int sw = __ldg(P.tile_seg + tile);
    while (sw < P.nseg && __ldg(P.bnd + sw + 1) <= iw) ++sw;
    int s = sw;
    while (s < P.nseg && __ldg(P.bnd + s + 1) <= i0) ++s;
    long long nb = (s < P.nseg) ? __ldg(P.bnd + s + 1) : (long long)0x7fffffffffffffffLL;
    double accA = 0.0, accB = 0.0, acc = 0.0;
    const float* cb = smem + tid * kChunk;
    double* dst = P.segsum + (size_t)row * (size_t)P.nseg;
#pragma unroll 4
    for (int j = 0; j < kS; ++j) {
        const long long i = i0 + j;
        if (i >= nb) {
            if (s == sw) accA += acc; else if (s == sw + 1) accB += acc; else if (s < P.nseg && acc != 0.0) atomicAdd(dst + s, acc);
            acc = 0.0;
            while (s < P.nseg && __ldg(P.bnd + s + 1) <= i) ++s;
            nb = (s < P.nseg) ? __ldg(P.bnd + s + 1) : (long long)0x7fffffffffffffffLL;
        }
        const double y = (double)cb[j];
        if (i >= 0) acc = fma(y, y, acc);
    }
    if (s == sw) accA += acc; else if (s == sw + 1) accB += acc; else if (s < P.nseg && acc != 0.0) atomicAdd(dst + s, acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { accA += shfl_xor_d(accA, o); accB += shfl_xor_d(accB, o); }
    if (lane == 0) {
        if (sw < P.nseg && accA != 0.0) atomicAdd(dst + sw, accA);
        if (sw + 1 < P.nseg && accB != 0.0) atomicAdd(dst + sw + 1, accB);
    }
}

// Two-pass gating; one CTA per track.
struct GateArgs {
    const double* segsum;    // [rows][nseg]
    int nseg, nblocks, channels, tracks;
    const int* blk_lo;       // [nblocks] first segment of block j
    const int* blk_hi;       // [nblocks] one past the last segment of block j
    double scale;            // 1 / (0.4 * rate)
    int valid;               // 0: signal shorter than one block (pyloudnorm raises)
    double* lufs;            // [tracks]
    const double* target;    // [tracks] or null
    double* gain_row;        // [rows] linear gain written for the next prologue, or null
    double* gain_db;         // [tracks] or null
};

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;
}

__global__ void __launch_bounds__(256) gate_kernel(const GateArgs P) {
    __shared__ double red[8];
    const int track = blockIdx.x;
    const int C = P.channels;
    double lufs;
    if (!P.valid) {
        lufs = __longlong_as_double(0x7ff8000000000000LL);
    } else {
        const double* s0 = P.segsum + (size_t)(track * C) * P.nseg;
        const double* s1 = s0 + (C > 1 ? P.nseg : 0);
        double gamma_r = 0.0;
        double result = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double a0 = 0.0, a1 = 0.0, cnt = 0.0;
            for (int j = threadIdx.x; j < P.nblocks; j += blockDim.x) {
                double z0 = 0.0, z1 = 0.0;
                for (int s = P.blk_lo[j]; s < P.blk_hi[j]; ++s) { z0 += s0[s]; if (C > 1) z1 += s1[s]; }
                z0 *= P.scale; z1 *= P.scale;
                const double l = -0.691 + 10.0 * log10(z0 + (C > 1 ? z1 : 0.0));
                const bool keep = pass == 0 ? (l >= -70.0) : (l > gamma_r && l > -70.0);
                if (keep) { a0 += z0; a1 += z1; cnt += 1.0; }
            }
            a0 = block_reduce_sum(a0, red);
            a1 = block_reduce_sum(a1, red);
            cnt = block_reduce_sum(cnt, red);
            double m0, m1;
            if (cnt > 0.0) { m0 = a0 / cnt; m1 = a1 / cnt; }
            else if (pass == 0) { m0 = m1 = __longlong_as_double(0x7ff8000000000000LL); }   // mean of empty -> nan
            else { m0 = m1 = 0.0; }                                                           // nan_to_num
            const double v = -0.691 + 10.0 * log10(m0 + (C > 1 ? m1 : 0.0));
            if (pass == 0) gamma_r = v - 10.0; else result = v;
        }
        lufs = result;
    }
    if (threadIdx.x == 0) {
        P.lufs[track] = lufs;
        if (P.target != nullptr) {
            double g = 1.0, gdb = 0.0;
            if (P.valid) {
                gdb = P.target[track] - lufs;
                gdb = fmin(fmax(gdb, -20.0), 20.0);      // np.clip; +inf (silence) -> +20
                g = pow(10.0, gdb / 20.0);
            }
            if (P.gain_row) for (int c = 0; c < C; ++c) P.gain_row[track * C + c] = g;
            if (P.gain_db) P.gain_db[track] = gdb;
        }
    }
}

}  // namespace mm
