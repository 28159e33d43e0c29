// ITU-R BS.1770 K-weighting (two causal biquads, zero initial state) fused with the 400 ms block
// mean-square partial sums, followed by the two-pass gating -- pyloudnorm.Meter.integrated_loudness
// as called at backend/app/pipeline.py:646-648 / :660-662.
//
// Same decomposition as the sweep kernel (sweep3.cuh): a CTA walks a segment of tiles of one row with
// the filter state carried in shared memory; the state at the segment start is rebuilt from a halo.
// The shelf -> high-pass cascade is ONE 4-state linear system here: pass 1 (float64) forms the zero-state
// end state of each 32-sample chunk straight from the input samples, one 4x4 scan resolves the state
// entering every chunk, and pass 2 runs the two sections in float32 on their balanced realizations
// (design.h) from that exact start state, squaring and summing as it goes.  Round-off therefore never
// accumulates beyond 32 samples; the measured loudness moves by < 1e-5 LU against the +-0.01 LU
// tolerance.  (pyloudnorm also rounds each stage's output to float32; that 6e-8 relative rounding is
// not reproduced, as before.)  The squares are summed per 100 ms hop into 64-bit fixed-point
// accumulators (integer atomics are associative, so the loudness -- and the gain derived from it -- is
// bit-reproducible).
#pragma once
#include "sweep3.cuh"

namespace mm {

constexpr double kSqScale = 1099511627776.0;        // 2^40: fixed-point scale of the square sums

struct LufsArgs {
    PairK k;                 // .x = shelf, .y = high-pass (balanced realizations); g[j] = cascade pass-1 weights:
                             // g[j][0] = states (0, 1) of the shelf, g[j][1] = states (2, 3) of the high-pass
    const double* tab;       // Tab<4> of the cascade (device): Plane[lane] is read from here
    // float32 scan tables, column pairs: M[k][0] = rows (0, 1) of column k, M[k][1] = rows (2, 3)
    float2 Pw2[5][4][2];     // (A^32)^(2^d)
    float2 Qw2[kNW + 1][4][2];   // (A^1024)^w; [kNW] carries a state across one tile
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
    long long goff;          // index of the row's sample 0 in the signal the hops are laid over (time slices; else 0)
    long long own_lo, own_hi;   // samples of the row that are counted (time slices; else 0, n)
    unsigned long long* segsum;   // [rows][nhop] fixed-point sums of squares
};

// The scan runs in float32 as well: in the balanced coordinates every combine is well conditioned, a state error of
// 6e-8 dies with the filters' own memory (A^4096 is ~1e-10 for the 38 Hz high-pass at 44.1 kHz), and the meter only
// needs the block powers to ~1e-6 relative (0.01 LU = 2.3e-3).  No float64 instruction is left in this kernel.
struct LufsTab {             // per-lane scan table of the 4-state cascade in shared memory
    float PlaneT[16][32];    // Plane[lane][k] transposed: lane-contiguous, conflict free
};
struct LufsScratch {
    float4 tot[kNW];
    float4 carry[2];         // [tile parity]
};
// acc (rows 0,1 | rows 2,3) += M v
__device__ __forceinline__ void mv4(const float2 (&M)[4][2], const float (&v)[4], float2& a01, float2& a23) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 vk = make_float2(v[k], v[k]);
        a01 = ffma2(M[k][0], vk, a01);
        a23 = ffma2(M[k][1], vk, a23);
    }
}

constexpr int kLufsSmem = 2 * kL * (int)sizeof(float) + (int)sizeof(LufsTab);

__global__ void __launch_bounds__(kT, 6) lufs_kernel(const __grid_constant__ LufsArgs P) {
    // two float32 staging tiles: the NEXT tile lands (cp.async) while the current one is scanned.  Both passes read their
    // samples from shared memory in rolled loops over float4 groups: the whole kernel stays small enough for the
    // instruction cache (the fully unrolled register version spent 30 % of its issue slots waiting for instructions)
    extern __shared__ __align__(128) unsigned char lufs_smem[];
    float* tiles = reinterpret_cast<float*>(lufs_smem);
    LufsTab* tab = reinterpret_cast<LufsTab*>(lufs_smem + 2 * kL * sizeof(float));
    __shared__ LufsScratch sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * 16; i += kT) tab->PlaneT[i % 16][i / 16] = (float)__ldg(P.tab + Tab<4>::Plane + i);
    const int cbase = tid * 32, cx = (tid & 7) << 2;
    const int items = P.rows * P.nseg;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int row = item % P.rows, seg = item / P.rows;
        const int t_live = seg * P.seglen;
        const int t_end = min(P.ntiles, t_live + P.seglen);
        const int t_first = max(0, t_live - P.whalo);
        const float* src = P.in + (size_t)row * (size_t)P.stride;
        __syncthreads();
        if (tid == 0) sh.carry[t_first & 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned long long* dst = P.segsum + (size_t)row * (size_t)P.nhop;
        auto load_tile = [&](int t, int slot) {
            float* tile_s = tiles + (size_t)slot * kL;
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
        load_tile(t_first, 0);
#pragma unroll 1
        for (int tile = t_first; tile < t_end; ++tile) {
            const bool live = tile >= t_live;
            const long long tile_lo = (long long)tile * kL;
            const int slot = (tile - t_first) & 1;
            cp_async_wait<0>();
            __syncthreads();                               // this tile's floats visible; everybody is done with the other slot
            if (tile + 1 < t_end) load_tile(tile + 1, slot ^ 1);     // lands while this tile is scanned
            const float* mine = tiles + (size_t)slot * kL + cbase;  // this thread's 32 samples: float4 group u at ((4 u) ^ cx)

            // ---- pass 1 (packed float32): zero-state end state of the 4-state cascade over this chunk ----
            float2 E01 = make_float2(0.f, 0.f), E23 = E01;
#pragma unroll 2
            for (int u = 0; u < kS / 4; ++u) {
                const float4 xv = *reinterpret_cast<const float4*>(mine + ((4 * u) ^ cx));
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float x = comp4(xv, c);
                    const float2 X = make_float2(x, x);
                    E01 = ffma2(P.k.g[4 * u + c][0], X, E01);
                    E23 = ffma2(P.k.g[4 * u + c][1], X, E23);
                }
            }
            // ---- warp scan, tile Horner, carry update (float32, packed) ----
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                float pe[4];
                pe[0] = __shfl_up_sync(0xffffffffu, E01.x, 1 << d);
                pe[1] = __shfl_up_sync(0xffffffffu, E01.y, 1 << d);
                pe[2] = __shfl_up_sync(0xffffffffu, E23.x, 1 << d);
                pe[3] = __shfl_up_sync(0xffffffffu, E23.y, 1 << d);
                if (lane < (1 << d)) { pe[0] = 0.f; pe[1] = 0.f; pe[2] = 0.f; pe[3] = 0.f; }
                mv4(P.Pw2[d], pe, E01, E23);
            }
            if (lane == 31) sh.tot[warp] = make_float4(E01.x, E01.y, E23.x, E23.y);
            __syncthreads();
            float2 b01 = make_float2(0.f, 0.f), b23 = b01;        // state contributed by the warps before this one
#pragma unroll
            for (int v = 0; v < kNW - 1; ++v) {
                if (v < warp) {
                    const float4 t4 = sh.tot[v];
                    const float bv[4] = {b01.x, b01.y, b23.x, b23.y};
                    float2 n01 = make_float2(t4.x, t4.y), n23 = make_float2(t4.z, t4.w);
                    mv4(P.Qw2[1], bv, n01, n23);
                    b01 = n01; b23 = n23;
                }
            }
            const float4 c4 = sh.carry[tile & 1];
            const float cin[4] = {c4.x, c4.y, c4.z, c4.w};
            if (tid == kT - 1) {
                const float bv[4] = {b01.x, b01.y, b23.x, b23.y};
                float2 a01 = E01, a23 = E23;
                mv4(P.Qw2[1], bv, a01, a23);
                mv4(P.Qw2[kNW], cin, a01, a23);
                sh.carry[(tile + 1) & 1] = make_float4(a01.x, a01.y, a23.x, a23.y);
            }
            if (!live) continue;                           // halo: only the carried state was needed
            mv4(P.Qw2[warp], cin, b01, b23);               // + the state entering the tile, carried to this warp
            float z[4];
            {
                const float u0 = __shfl_up_sync(0xffffffffu, E01.x, 1), u1 = __shfl_up_sync(0xffffffffu, E01.y, 1);
                const float u2 = __shfl_up_sync(0xffffffffu, E23.x, 1), u3 = __shfl_up_sync(0xffffffffu, E23.y, 1);
                z[0] = lane > 0 ? u0 : 0.f; z[1] = lane > 0 ? u1 : 0.f; z[2] = lane > 0 ? u2 : 0.f; z[3] = lane > 0 ? u3 : 0.f;
            }
            const float bs[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float acc = z[i];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fmaf(tab->PlaneT[i * 4 + k][lane], bs[k], acc);
                z[i] = acc;
            }

            // ---- pass 2 (packed float32, the high-pass runs one sample behind the shelf) + squared sums per hop ----
            const long long i0 = tile_lo + (long long)tid * kS - kLead;          // first sample index of this thread
            const long long iw = tile_lo + (long long)(tid & ~31) * kS - kLead;   // first sample of this warp
            const long long ig0 = i0 + P.goff, igw = iw + P.goff;                    // the same in hop coordinates
            int sw = __ldg(P.tile_seg + tile);
            while (sw < P.nhop && __ldg(P.bnd + sw + 1) <= igw) ++sw;
            int s = sw;
            while (s < P.nhop && __ldg(P.bnd + s + 1) <= ig0) ++s;
            const long long kBig = 0x3fffffffffffffffLL;
            const long long nb1 = (s < P.nhop) ? __ldg(P.bnd + s + 1) - P.goff : kBig;     // row-local hop ends
            const long long nb2 = (s + 1 < P.nhop) ? __ldg(P.bnd + s + 2) - P.goff : kBig;
            const bool whole = (i0 >= P.own_lo) && (i0 + kS <= P.own_hi) && (i0 + kS <= nb1);   // chunk inside the counted range and one hop
            float2 S0 = make_float2(z[0], z[2]);
            float2 S1 = make_float2(z[1], z[3]);
            // the high-pass lane runs one sample behind the shelf lane: step 0 is the shelf alone, the last output comes
            // from the high-pass state after the loop
            float u_prev = 0.f;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
            auto first_step = [&](float x0) {
                const float u = fmaf(P.k.C[0].x, S0.x, fmaf(P.k.C[1].x, S1.x, P.k.D.x * x0));
                const float n0 = fmaf(P.k.A[0][0].x, S0.x, fmaf(P.k.A[0][1].x, S1.x, P.k.B[0].x * x0));
                const float n1 = fmaf(P.k.A[1][0].x, S0.x, fmaf(P.k.A[1][1].x, S1.x, P.k.B[1].x * x0));
                S0.x = n0; S1.x = n1;
                u_prev = u;
            };
            if (__all_sync(0xffffffffu, whole)) {
#pragma unroll 2
                for (int u = 0; u < kS / 4; ++u) {
                    const float4 xv = *reinterpret_cast<const float4*>(mine + ((4 * u) ^ cx));
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c == 0 && u == 0) { first_step(xv.x); continue; }
                        const float2 Y = pair_step(P.k, make_float2(comp4(xv, c), u_prev), S0, S1);   // Y.x = shelf(j), Y.y = K-weighted(j-1)
                        u_prev = Y.x;
                        if (c & 1) acc0 = fmaf(Y.y, Y.y, acc0); else acc1 = fmaf(Y.y, Y.y, acc1);
                    }
                }
                const float yl = fmaf(P.k.C[0].y, S0.y, fmaf(P.k.C[1].y, S1.y, P.k.D.y * u_prev));
                acc1 = fmaf(yl, yl, acc1);
                acc0 += acc1; acc1 = 0.f;
            } else {
                // chunks cut by hop boundaries (at most two: the host plan guarantees it) or by the ends of the counted range:
                // samples [jlo, j1) -> hop s, [j1, j2) -> hop s+1, [j2, jv) -> hop s+2; the rest is not counted
                const int jv = (int)max(0LL, min((long long)kS, P.own_hi - i0));
                const int jlo = (int)max(0LL, min((long long)kS, P.own_lo - i0));
                const int j1 = (int)max(0LL, min((long long)jv, nb1 - i0));
                const int j2 = (int)max((long long)j1, min((long long)jv, nb2 - i0));
                auto count = [&](float yk, int js) {       // js: the sample this K-weighted output belongs to
                    const float t = yk * yk;
                    acc0 += (js >= jlo && js < j1) ? t : 0.f;
                    acc1 += (js >= jlo && js >= j1 && js < j2) ? t : 0.f;
                    acc2 += (js >= jlo && js >= j2 && js < jv) ? t : 0.f;
                };
#pragma unroll 1
                for (int u = 0; u < kS / 4; ++u) {
                    const float4 xv = *reinterpret_cast<const float4*>(mine + ((4 * u) ^ cx));
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c == 0 && u == 0) { first_step(xv.x); continue; }
                        const float2 Y = pair_step(P.k, make_float2(comp4(xv, c), u_prev), S0, S1);
                        u_prev = Y.x;
                        count(Y.y, 4 * u + c - 1);
                    }
                }
                count(fmaf(P.k.C[0].y, S0.y, fmaf(P.k.C[1].y, S1.y, P.k.D.y * u_prev)), kS - 1);
            }
            double accA = 0.0, accB = 0.0;
            auto flush = [&](int hop, float v) {
                if (v == 0.f) return;
                if (hop == sw) accA += (double)v;
                else if (hop == sw + 1) accB += (double)v;
                else if (hop < P.nhop) atomicAdd(dst + hop, (unsigned long long)__double2ll_rn((double)v * kSqScale));
            };
            flush(s, acc0);
            flush(s + 1, acc1);
            flush(s + 2, acc2);
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
