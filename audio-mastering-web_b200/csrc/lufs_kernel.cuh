// ITU-R BS.1770 K-weighting (two causal biquads, zero initial state) fused with the 400 ms block
// mean-square partial sums, followed by the two-pass gating -- pyloudnorm.Meter.integrated_loudness
// as called at backend/app/pipeline.py:646-648 / :660-662.
//
// Same decomposition as the sweep kernel (sweep3.cuh): a CTA walks a segment of tiles of one row with
// the filter states carried in shared memory; the state at the segment start is rebuilt from a halo of
// W(shelf) + W(high-pass) tiles.  Per tile the shelf is scanned in place, rounded to float32 (pyloudnorm
// writes each stage back into a copy of its float32 input), then the high-pass; the squares of the
// result are summed per 100 ms hop into 64-bit fixed-point accumulators (integer atomics are
// associative, so the loudness -- and the gain derived from it -- is bit-reproducible).
#pragma once
#include "sweep3.cuh"

namespace mm {

constexpr double kSqScale = 1099511627776.0;        // 2^40: fixed-point scale of the square sums

struct LufsArgs {
    FiltK<2> f[2];
    const double* tab[2];
    const float* in;
    long long n, stride;
    int rows, ntiles, channels;
    int seglen, nseg, whalo;
    int pro_mode;
    const double* pro_sub;
    const double* pro_mul;
    // hop bookkeeping: sample i belongs to hop s iff bnd[s] <= i < bnd[s+1]
    const long long* bnd;    // [nhop + 1]
    int nhop;
    const int* tile_seg;     // [ntiles] hop of max(first sample of tile, 0), clamped to nhop
    unsigned long long* segsum;   // [rows][nhop] fixed-point sums of squares
};

struct LufsScratch {
    double tot[kNW][2];
    double carry[2][2][2];   // [tile parity][filter][state]
};

// pyloudnorm writes each stage's float64 lfilter output back into its float32 buffer.  Those two roundings
// (6e-8 relative) move the loudness by ~1e-7 dB against a +-0.01 LU tolerance; reproducing them costs either
// two conversions or three FP64 operations per sample and stage in a kernel that is bound by exactly those
// pipes, so the stages are chained in float64 here.
__device__ __forceinline__ double round_to_f32(double y) { return y; }

constexpr int kLufsSmem = kL * (int)sizeof(float) + kL * (int)sizeof(double) + 2 * (int)sizeof(SmemTab<2>);

__global__ void __launch_bounds__(kT) lufs_kernel(const __grid_constant__ LufsArgs P) {
    // a float32 staging tile receives the NEXT tile (cp.async) while the current one is scanned; the shelf's
    // pass 2 writes its float32-rounded output as float64 into a second buffer, so the high-pass stage needs
    // no conversions at all
    extern __shared__ __align__(128) unsigned char lufs_smem[];
    float* tile_s = reinterpret_cast<float*>(lufs_smem);
    double* tile_d = reinterpret_cast<double*>(lufs_smem + kL * sizeof(float));
    SmemTab<2>* tab = reinterpret_cast<SmemTab<2>*>(lufs_smem + kL * sizeof(float) + kL * sizeof(double));
    __shared__ LufsScratch sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int MM = 4;
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
        double* d = reinterpret_cast<double*>(&tab[f]);
        for (int i = tid; i < (5 + 32 + kNW + 1) * MM; i += kT) d[i] = __ldg(P.tab[f] + i);
    }
    const int cbase = tid * 32, cx = (tid & 7) << 2;
    // float64 layout of this thread's chunk: 16 vectors of 2 doubles at dbase + ((2 w) ^ dx), conflict free
    // for 16-byte accesses (8 consecutive threads cover all 8 16-byte bank groups)
    const int dbase = tid * 32, dx = (tid & 7) << 1;
    const int items = P.rows * P.nseg;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int row = item % P.rows, seg = item / P.rows;
        const int t_live = seg * P.seglen;
        const int t_end = min(P.ntiles, t_live + P.seglen);
        const int t_first = max(0, t_live - P.whalo);
        const float* src = P.in + (size_t)row * (size_t)P.stride;
        float subf = 0.f, mulf = 1.f;
        double muld = 1.0;
        if (P.pro_mode != PRO_NONE) {
            if (P.pro_sub) subf = (float)__ldg(P.pro_sub + row);
            if (P.pro_mul) { muld = __ldg(P.pro_mul + row); mulf = (float)muld; }
        }
        __syncthreads();
        if (tid < 4) sh.carry[t_first & 1][tid >> 1][tid & 1] = 0.0;
        unsigned long long* dst = P.segsum + (size_t)row * (size_t)P.nhop;
        auto load_tile = [&](int t) {
            const long long lo = (long long)t * kL;
            if (lo >= kLead && lo + kL <= kLead + P.n) {
                const float* s4 = src + lo + 4 * tid;
                float* d4 = tile_s + 4 * swz(tid);
#pragma unroll
                for (int r = 0; r < kTileVecs / kT; ++r) cp_async16(d4 + 4 * kT * r, s4 + 4 * kT * r);
            } else {
#pragma unroll 1
                for (int r = 0; r < kTileVecs / kT; ++r) {
                    const int v = tid + kT * r;
                    const long long q = lo + 4 * v;
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (q + c >= kLead && q + c < kLead + P.n) setcomp4(val, c, src[q + c]);
                    *reinterpret_cast<float4*>(tile_s + 4 * swz(v)) = val;
                }
            }
            cp_async_commit();
        };
        load_tile(t_first);
#pragma unroll 1
        for (int tile = t_first; tile < t_end; ++tile) {
            const bool live = tile >= t_live;
            const long long tile_lo = (long long)tile * kL;
            const bool interior = tile_lo >= kLead && tile_lo + kL <= kLead + P.n;
            cp_async_wait<0>();
            __syncthreads();                               // staged floats of this tile visible; previous tile done

            // ================= stage 0: high shelf, float32 in -> float32-rounded float64 out ===============
            // this thread's 32 input samples leave shared memory here: every thread reads its float chunk
            // before anybody overwrites the buffer with doubles (barrier below)
            float xin[32];
#pragma unroll
            for (int u = 0; u < kS / 4; ++u) {
                float4 xv = *reinterpret_cast<const float4*>(tile_s + cbase + ((4 * u) ^ cx));
                if (P.pro_mode != PRO_NONE) {
                    const long long q0 = tile_lo + cbase + 4 * u;      // dead positions of edge tiles stay exactly zero
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool in_range = interior || (q0 + c >= kLead && q0 + c < kLead + P.n);
                        if (in_range) setcomp4(xv, c, pro1(P.pro_mode, comp4(xv, c), subf, mulf, muld));
                    }
                }
                xin[4 * u] = xv.x; xin[4 * u + 1] = xv.y; xin[4 * u + 2] = xv.z; xin[4 * u + 3] = xv.w;
            }
            __syncthreads();                               // every thread holds its inputs: the staging tile is free
            if (tile + 1 < t_end) load_tile(tile + 1);     // lands while this tile is scanned
            double E[2] = {0.0, 0.0};
#pragma unroll
            for (int j = 0; j < kS; ++j) {
                const double x = (double)xin[j];
                E[0] = fma(P.f[0].g[j][0], x, E[0]);
                E[1] = fma(P.f[0].g[j][1], x, E[1]);
            }
            double base[2], cin[2], z[2];
            auto resolve = [&](int f) {
                // warp scan, tile Horner, carry update; leaves the state entering this thread's chunk in z
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    double pe[2];
                    pe[0] = shfl_up_d(E[0], 1 << d);
                    pe[1] = shfl_up_d(E[1], 1 << d);
                    if (lane >= (1 << d)) matvec_acc_s<2>(tab[f].Pw[d], pe, E);
                }
                if (lane == 31) { sh.tot[warp][0] = E[0]; sh.tot[warp][1] = E[1]; }
                __syncthreads();
                base[0] = 0.0; base[1] = 0.0;
                for (int v = 0; v < warp; ++v) {
                    double nb[2] = {sh.tot[v][0], sh.tot[v][1]};
                    matvec_acc_s<2>(tab[f].Qpow[1], base, nb);
                    base[0] = nb[0]; base[1] = nb[1];
                }
                cin[0] = sh.carry[tile & 1][f][0]; cin[1] = sh.carry[tile & 1][f][1];
                if (tid == kT - 1) {
                    double ag[2] = {E[0], E[1]};
                    matvec_acc_s<2>(tab[f].Qpow[1], base, ag);
                    matvec_acc_s<2>(tab[f].Qpow[kNW], cin, ag);
                    sh.carry[(tile + 1) & 1][f][0] = ag[0];
                    sh.carry[(tile + 1) & 1][f][1] = ag[1];
                }
                matvec_acc_s<2>(tab[f].Qpow[warp], cin, base);
                z[0] = shfl_up_d(E[0], 1);
                z[1] = shfl_up_d(E[1], 1);
                if (lane == 0) { z[0] = 0.0; z[1] = 0.0; }
                matvec_acc_s<2>(tab[f].Plane[lane], base, z);
            };
            resolve(0);
            double E1[2] = {0.0, 0.0};
#pragma unroll
            for (int w = 0; w < kS / 2; ++w) {
                const double y0 = round_to_f32(df2t_step<2>(P.f[0], (double)xin[2 * w], z));
                const double y1 = round_to_f32(df2t_step<2>(P.f[0], (double)xin[2 * w + 1], z));
                // stage 1's pass 1 rides along: zero-state end state of the high-pass over this chunk
                E1[0] = fma(P.f[1].g[2 * w][0], y0, E1[0]);
                E1[1] = fma(P.f[1].g[2 * w][1], y0, E1[1]);
                E1[0] = fma(P.f[1].g[2 * w + 1][0], y1, E1[0]);
                E1[1] = fma(P.f[1].g[2 * w + 1][1], y1, E1[1]);
                *reinterpret_cast<double2*>(tile_d + dbase + ((2 * w) ^ dx)) = make_double2(y0, y1);
            }
            // ================= stage 1: high-pass on the float64 tile ======================================
            E[0] = E1[0]; E[1] = E1[1];
            __syncthreads();     // sh.tot of stage 0 has been consumed by everybody
            resolve(1);
            if (!live) continue;                           // halo: only the carried states were needed

            // pass 2 fused with the squared sums per hop (the squares use the float32-rounded output)
            const long long i0 = tile_lo + (long long)tid * kS - kLead;          // first sample index of this thread
            const long long iw = tile_lo + (long long)(tid & ~31) * kS - kLead;   // first sample of this warp
            int sw = __ldg(P.tile_seg + tile);
            while (sw < P.nhop && __ldg(P.bnd + sw + 1) <= iw) ++sw;
            int s = sw;
            while (s < P.nhop && __ldg(P.bnd + s + 1) <= i0) ++s;
            long long nb = (s < P.nhop) ? __ldg(P.bnd + s + 1) : (long long)0x7fffffffffffffffLL;
            double accA = 0.0, accB = 0.0, acc = 0.0;
            auto flush = [&](int hop, double v) {
                if (hop == sw) accA += v;
                else if (hop == sw + 1) accB += v;
                else if (hop < P.nhop && v != 0.0) atomicAdd(dst + hop, (unsigned long long)__double2ll_rn(v * kSqScale));
            };
            const bool whole = (i0 >= 0) && (i0 + kS <= P.n) && (i0 + kS <= nb);   // chunk inside the row and one hop
#pragma unroll
            for (int w = 0; w < kS / 2; ++w) {
                const double2 xv = *reinterpret_cast<const double2*>(tile_d + dbase + ((2 * w) ^ dx));
                const double y0 = round_to_f32(df2t_step<2>(P.f[1], xv.x, z));
                const double y1 = round_to_f32(df2t_step<2>(P.f[1], xv.y, z));
                if (whole) {
                    acc = fma(y0, y0, acc);
                    acc = fma(y1, y1, acc);
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const long long i = i0 + 2 * w + c;
                        if (i >= nb) {
                            flush(s, acc);
                            acc = 0.0;
                            while (s < P.nhop && __ldg(P.bnd + s + 1) <= i) ++s;
                            nb = (s < P.nhop) ? __ldg(P.bnd + s + 1) : (long long)0x7fffffffffffffffLL;
                        }
                        const double y = c == 0 ? y0 : y1;
                        if (i >= 0 && i < P.n) acc = fma(y, y, acc);
                    }
                }
            }
            flush(s, acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { accA += shfl_xor_d(accA, o); accB += shfl_xor_d(accB, o); }
            if (lane == 0) {
                if (sw < P.nhop && accA != 0.0) atomicAdd(dst + sw, (unsigned long long)__double2ll_rn(accA * kSqScale));
                if (sw + 1 < P.nhop && accB != 0.0) atomicAdd(dst + sw + 1, (unsigned long long)__double2ll_rn(accB * kSqScale));
            }
        }
    }
}

// Two-pass gating; one CTA per track.
struct GateArgs {
    const unsigned long long* segsum;   // [rows][nseg] fixed-point
    int nseg, nblocks, channels, tracks;
    const int* blk_lo;       // [nblocks] first hop of block j
    const int* blk_hi;       // [nblocks] one past the last hop of block j
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
        const unsigned long long* s0 = P.segsum + (size_t)(track * C) * P.nseg;
        const unsigned long long* s1 = s0 + (C > 1 ? P.nseg : 0);
        const double inv = 1.0 / kSqScale;
        double gamma_r = 0.0;
        double result = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double a0 = 0.0, a1 = 0.0, cnt = 0.0;
            for (int j = threadIdx.x; j < P.nblocks; j += blockDim.x) {
                unsigned long long u0 = 0, u1 = 0;
                for (int s = P.blk_lo[j]; s < P.blk_hi[j]; ++s) { u0 += s0[s]; if (C > 1) u1 += s1[s]; }
                const double z0 = (double)u0 * inv * P.scale, z1 = (double)u1 * inv * P.scale;
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
