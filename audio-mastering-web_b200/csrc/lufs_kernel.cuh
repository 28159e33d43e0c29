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
#include <cooperative_groups.h>
#include "sweep4.cuh"

namespace mm {

constexpr double kSqScale = 1099511627776.0;        // 2^40: fixed-point scale of the square sums

constexpr int kLufsMaxS = 64;
struct LufsArgs {
    // pass-1 weights of the cascade for a lane's chunk of S samples: g[j][0] = states (0, 1) of the shelf, g[j][1] = states (2, 3) of
    // the high-pass (balanced coordinates)
    float2 g[kLufsMaxS][2];
    const double* plane;     // [32][16] (device): (A^S)^lane of the cascade
    // float32 scan tables, column pairs: M[k][0] = rows (0, 1) of column k, M[k][1] = rows (2, 3)
    float2 Pw2[5][4][2];     // (A^S)^(2^d)
    float2 Q1[4][2];         // A^(32 S): carries a state across one warp-tile
    // pass 2.  The shelf as a float32 DF2T section: b, negated a1 / a2, and the map from the balanced state to the DF2T state.
    // The high-pass (b = g [1, -2, 1]) as a Chamberlin state-variable filter on the scan's states (lp, bp):
    //   lp += f bp;  hp = (u - lp) - q bp;  bp += f hp  -- 4 operations (the balanced realization needed 7); its gain g is applied
    // to the hop sums (g^2, in float64) instead of to every sample.
    float sh_b[3], sh_na[2], shT[2][2];
    float hp_f, hp_nq;
    double hp_g2;
    const float* in;
    long long n, stride;
    int rows, ntiles, channels;   // ntiles, seglen, whalo: in WARP-tiles of 1024 samples
    int seglen, nseg, whalo;
    int pro_mode;
    const double* pro_sub;
    const double* pro_mul;
    // hop bookkeeping: sample i belongs to hop s iff bnd[s] <= i < bnd[s+1]
    const long long* bnd;    // [nhop + 1]
    int nhop;
    const int* tile_seg;     // [4096-sample tiles] hop of max(first sample of tile, 0), clamped to nhop
    long long goff;          // index of the row's sample 0 in the signal the hops are laid over (time slices; else 0)
    long long own_lo, own_hi;   // samples of the row that are counted (time slices; else 0, n)
    unsigned long long* segsum;   // [rows][nhop] fixed-point sums of squares
};

// The scan runs in float32 as well: in the balanced coordinates every combine is well conditioned, a state error of
// 6e-8 dies with the filters' own memory (A^4096 is ~1e-10 for the 38 Hz high-pass at 44.1 kHz), and the meter only
// needs the block powers to ~1e-6 relative (0.01 LU = 2.3e-3).  No float64 instruction is left in the sample loops.
//
// Round 2: every WARP is autonomous.  A warp owns a segment of 1024-sample warp-tiles of one row (32 lanes x 32 samples), keeps
// the state entering its next warp-tile in registers (identical in all lanes), double-buffers its own 4 KB staging slots with
// cp.async and never meets a block barrier: the round-1 kernel spent 2.5 stall cycles per issued instruction at its two
// __syncthreads per tile.  The hop bookkeeping (which 100 ms hop a sample's square belongs to) is done once per warp-tile in
// 32-bit offsets relative to the warp-tile -- a warp-tile holds at most two hop boundaries (host check: bnd[s+2] - bnd[s] >= 1024)
// -- instead of per chunk in 64-bit arithmetic, and the per-hop partial sums are reduced with float32 shuffles (fixed order:
// bit-reproducible) before one fixed-point atomic per warp-tile and hop.  A lane's chunk is S = 64 samples (one warp scan per
// 2048 samples; S = 32 for sample rates whose 100 ms hops are shorter than 1024 samples).
struct LufsTab {                                 // per-lane scan table of the 4-state cascade in shared memory
    float PlaneT[16][32];                        // Plane[lane][k] transposed: lane-contiguous, conflict free
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

template <int S> struct LufsCfg {
    static constexpr int kWT = 32 * S;             // samples per warp-tile
    static constexpr int kSmem = kNW * 2 * kWT * (int)sizeof(float) + (int)sizeof(LufsTab);
    static constexpr int kMinBlocks = S == 32 ? 6 : 3;
};

template <int S>
__global__ void __launch_bounds__(kT, LufsCfg<S>::kMinBlocks) lufs_kernel(const __grid_constant__ LufsArgs P) {
    constexpr int kWT = LufsCfg<S>::kWT;
    constexpr int kU = S / 4;                      // float4 groups per chunk; group u of lane l sits at vector l * kU + u
    extern __shared__ __align__(128) unsigned char lufs_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* slots = reinterpret_cast<float*>(lufs_smem) + (size_t)warp * 2 * kWT;      // this warp's two staging slots
    LufsTab* tab = reinterpret_cast<LufsTab*>(lufs_smem + (size_t)kNW * 2 * kWT * sizeof(float));
    for (int i = tid; i < 32 * 16; i += kT) tab->PlaneT[i % 16][i / 16] = (float)__ldg(P.plane + i);
    __syncthreads();                                         // the only block barrier: the table is read-only from here on
    float pl[16];                                            // this lane's Plane rows (A^(32 lane)): registers for the whole kernel
#pragma unroll
    for (int i = 0; i < 16; ++i) pl[i] = tab->PlaneT[i][lane];
    // a chunk is kU consecutive 16-byte vectors; vector v of the warp-tile is stored at sw(v): the low three bits XORed with the
    // chunk index (v / kU) & 7, so that eight neighbouring lanes walking their chunks, and eight neighbouring vectors of the
    // coalesced copy, both touch eight different 16-byte bank groups
    auto sw = [](int v) -> int { return (v & ~7) | ((v ^ (v >> (S == 64 ? 4 : 3))) & 7); };
    const int items = P.rows * P.nseg;
    const int wstride = gridDim.x * kNW;
#pragma unroll 1
    for (int item = blockIdx.x * kNW + warp; item < items; item += wstride) {
        const int row = item % P.rows, seg = item / P.rows;
        const int t_live = seg * P.seglen;
        const int t_end = min(P.ntiles, t_live + P.seglen);
        const int t_first = max(0, t_live - P.whalo);
        const float* src = P.in + (size_t)row * (size_t)P.stride;
        unsigned long long* dst = P.segsum + (size_t)row * (size_t)P.nhop;
        auto load_tile = [&](int t, int slot) {
            float* tile_s = slots + (size_t)slot * kWT;
            const long long lo = (long long)t * kWT;
            if (lo >= kLead && lo + kWT <= kLead + P.n) {
                const float* s4 = src + lo + 4 * lane;
#pragma unroll
                for (int r = 0; r < kWT / 128; ++r) cp_async16(tile_s + 4 * sw(lane + 32 * r), s4 + 128 * r);
            } else {
#pragma unroll 1
                for (int r = 0; r < kWT / 128; ++r) {
                    const int v = lane + 32 * r;
                    const long long q = lo + 4 * v;
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (q + c >= kLead && q + c < kLead + P.n) setcomp4(val, c, src[q + c]);
                    *reinterpret_cast<float4*>(tile_s + 4 * sw(v)) = val;
                }
            }
            cp_async_commit();
        };
        float cin[4] = {0.f, 0.f, 0.f, 0.f};                 // state entering the next warp-tile (same in every lane)
        __syncwarp();                                        // the previous item's last reads of slot 0 are done
        load_tile(t_first, 0);
#pragma unroll 1
        for (int tile = t_first; tile < t_end; ++tile) {
            const bool live = tile >= t_live;
            const int slot = (tile - t_first) & 1;
            cp_async_wait<0>();
            __syncwarp();                                    // this warp-tile's floats visible; the other slot is free
            if (tile + 1 < t_end) load_tile(tile + 1, slot ^ 1);     // lands while this warp-tile is scanned
            const float* tile_s = slots + (size_t)slot * kWT;
            auto group = [&](int u) -> float4 { return *reinterpret_cast<const float4*>(tile_s + 4 * sw(lane * kU + u)); };

            // ---- pass 1 (packed float32): zero-state end state of the 4-state cascade over this lane's chunk ----
            float2 E01 = make_float2(0.f, 0.f), E23 = E01;
#pragma unroll
            for (int u = 0; u < kU; ++u) {                       // fully unrolled: g[j] are immediate constant-bank operands
                const float4 xv = group(u);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float x = comp4(xv, c);
                    const float2 X = make_float2(x, x);
                    E01 = ffma2(P.g[4 * u + c][0], X, E01);
                    E23 = ffma2(P.g[4 * u + c][1], X, E23);
                }
            }
            // ---- warp scan (float32, packed): inclusive prefix of the chunk end states ----
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
            // state entering this lane's chunk = zero-state prefix of the lanes before it + A^(32 lane) * (state entering the tile)
            float z[4];
            {
                const float u0 = __shfl_up_sync(0xffffffffu, E01.x, 1), u1 = __shfl_up_sync(0xffffffffu, E01.y, 1);
                const float u2 = __shfl_up_sync(0xffffffffu, E23.x, 1), u3 = __shfl_up_sync(0xffffffffu, E23.y, 1);
                z[0] = lane > 0 ? u0 : 0.f; z[1] = lane > 0 ? u1 : 0.f; z[2] = lane > 0 ? u2 : 0.f; z[3] = lane > 0 ? u3 : 0.f;
            }
            // state leaving the warp-tile = lane 31's inclusive prefix + A^1024 * (state entering it)
            float2 n01 = make_float2(__shfl_sync(0xffffffffu, E01.x, 31), __shfl_sync(0xffffffffu, E01.y, 31));
            float2 n23 = make_float2(__shfl_sync(0xffffffffu, E23.x, 31), __shfl_sync(0xffffffffu, E23.y, 31));
            mv4(P.Q1, cin, n01, n23);
            if (live) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float acc = z[i];
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc = fmaf(pl[i * 4 + k], cin[k], acc);
                    z[i] = acc;
                }
            }
            cin[0] = n01.x; cin[1] = n01.y; cin[2] = n23.x; cin[3] = n23.y;
            if (!live) continue;                             // halo: only the carried state was needed

            // ---- hop bookkeeping, once per warp-tile, in offsets relative to its first sample ----
            const long long iw = (long long)tile * kWT - kLead;          // sample index of the warp-tile's first position
            const long long igw = iw + P.goff;                            // the same in hop coordinates
            int hs = __ldg(P.tile_seg + (int)(((long long)tile * kWT) / kL));   // hop of the enclosing 4096-tile's first sample
            while (hs < P.nhop && __ldg(P.bnd + hs + 1) <= igw) ++hs;
            const long long kFar = 1 << 20;
            const int b1 = (int)min(kFar, max(0LL, (hs < P.nhop ? __ldg(P.bnd + hs + 1) : igw + kFar) - igw));      // first hop end, relative
            const int b2 = (int)min(kFar, max(0LL, (hs + 1 < P.nhop ? __ldg(P.bnd + hs + 2) : igw + kFar) - igw));  // second hop end
            const int lo_rel = (int)min(kFar, max(-kFar, P.own_lo - iw));      // counted range of the row, relative
            const int hi_rel = (int)min(kFar, max(-kFar, P.own_hi - iw));
            const bool fast = lo_rel <= 0 && hi_rel >= kWT && b1 >= kWT;       // whole warp-tile counted, one hop (warp-uniform)

            // ---- pass 2 (scalar float32) + squared sums per hop ----
            // Packing (x_j, shelf_{j-1}) into register pairs for FFMA2 costs more moves than the packed arithmetic saves (SASS: 26
            // instructions per sample packed, 15 here).  The shelf (poles at radius ~0.9) runs as a float32 DF2T section -- 5
            // operations -- from the scan's state mapped into DF2T coordinates; the 38 Hz high-pass (poles at 0.995) keeps its
            // balanced realization (9 operations).  Every chunk restarts from the scan-resolved state, so round-off lives 32 samples.
            float zs0 = fmaf(P.shT[0][0], z[0], P.shT[0][1] * z[1]), zs1 = fmaf(P.shT[1][0], z[0], P.shT[1][1] * z[1]);
            float lp = z[2], bp = z[3];
            auto kweight = [&](float x) -> float {
                const float u = fmaf(P.sh_b[0], x, zs0);                                   // shelf, DF2T
                zs0 = fmaf(P.sh_b[1], x, fmaf(P.sh_na[0], u, zs1));
                zs1 = fmaf(P.sh_b[2], x, P.sh_na[1] * u);
                lp = fmaf(P.hp_f, bp, lp);                                                 // high-pass, state-variable form (y / g)
                const float hp = fmaf(P.hp_nq, bp, u - lp);
                bp = fmaf(P.hp_f, hp, bp);
                return hp;
            };
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;                // sums of this lane's squares in hops hs, hs + 1, hs + 2
            if (fast) {
                float acc1 = 0.f;
#pragma unroll 2
                for (int u = 0; u < kU; ++u) {
                    const float4 xv = group(u);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float y = kweight(comp4(xv, c));
                        if (c & 1) a0 = fmaf(y, y, a0); else acc1 = fmaf(y, y, acc1);
                    }
                }
                a0 += acc1;
            } else {
                // this lane's samples [jlo, j1) belong to hop hs, [j1, j2) to hs + 1, [j2, jv) to hs + 2; the rest is not counted.
                // One running sum, flushed into its hop's slot whenever the sample index reaches the next boundary: a compare per
                // sample instead of three masked accumulations (46 % of the 2048-sample warp-tiles hold a hop boundary at 44.1 kHz,
                // so this path is half of the kernel)
                const int j0 = S * lane;
                const int jv = max(0, min(S, hi_rel - j0));
                const int jlo = max(0, min(jv, lo_rel - j0));
                const int j1 = max(jlo, min(jv, b1 - j0));
                const int j2 = max(j1, min(jv, b2 - j0));
                float acc = 0.f;
                int seg = 0, nb = jlo;                           // seg 0: before the counted range, 1..3: hops hs..hs + 2, 4: after it
                auto flush = [&]() {
                    if (seg == 1) a0 += acc; else if (seg == 2) a1 += acc; else if (seg == 3) a2 += acc;
                    acc = 0.f;
                    ++seg;
                    nb = seg == 1 ? j1 : (seg == 2 ? j2 : (seg == 3 ? jv : S + 1));
                };
#pragma unroll 1
                for (int u = 0; u < kU; ++u) {
                    const float4 xv = group(u);
                    if (__all_sync(0xffffffffu, nb >= 4 * u + 4)) {          // no lane has a boundary inside this group of four
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float y = kweight(comp4(xv, c));
                            acc = fmaf(y, y, acc);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            while (4 * u + c == nb) flush();                 // boundaries may coincide
                            const float y = kweight(comp4(xv, c));
                            acc = fmaf(y, y, acc);
                        }
                    }
                }
                while (seg < 4) flush();
            }
            // ---- per-hop sums of the warp-tile: fixed-order float32 butterfly, one fixed-point atomic per hop ----
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            if (!fast) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
            }
            if (lane == 0) {
                const double sc = P.hp_g2 * kSqScale;                  // the high-pass gain squared and the fixed-point scale
                if (hs < P.nhop && a0 != 0.f) atomicAdd(dst + hs, (unsigned long long)__double2ll_rn((double)a0 * sc));
                if (!fast) {
                    if (hs + 1 < P.nhop && a1 != 0.f) atomicAdd(dst + hs + 1, (unsigned long long)__double2ll_rn((double)a1 * sc));
                    if (hs + 2 < P.nhop && a2 != 0.f) atomicAdd(dst + hs + 2, (unsigned long long)__double2ll_rn((double)a2 * sc));
                }
            }
        }
        cp_async_wait<0>();
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

// The same gating for LONG signals (a two-hour file has 72 000 blocks: one 256-thread CTA walks them in 2 x 281 dependent rounds
// of L2 loads and float64 logarithms -- 0.53 ms, 11 % of a time slice's step on 8 GPUs).  A thread-block CLUSTER of kGateCluster
// CTAs x 1024 threads shares the blocks; the per-CTA partial sums meet through distributed shared memory in rank order (fixed
// order: bit-reproducible), every CTA ends with the same totals, rank 0 writes.  Chosen by the block count alone, so a signal's
// loudness never depends on what it is batched with.
constexpr int kGateCluster = 8;
constexpr int kGateLongThreads = 1024;
constexpr int kGateLongMinBlocks = 8192;

__global__ void __cluster_dims__(kGateCluster, 1, 1) __launch_bounds__(kGateLongThreads) gate_long_kernel(const GateArgs P) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double red[kGateLongThreads / 32];
    __shared__ double part[2][3];                                 // this CTA's (a0, a1, cnt) of pass 0 / pass 1
    const int track = blockIdx.x / kGateCluster;
    const int crank = (int)cluster.block_rank();
    const int C = P.channels;
    double lufs;
    if (!P.valid) {                                               // uniform over the grid: nobody waits at a cluster barrier
        lufs = __longlong_as_double(0x7ff8000000000000LL);
    } else {
        const unsigned long long* s0 = P.segsum + (size_t)(track * C) * P.nseg;
        const unsigned long long* s1 = s0 + (C > 1 ? P.nseg : 0);
        const double inv = 1.0 / kSqScale;
        double gamma_r = 0.0, result = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double a0 = 0.0, a1 = 0.0, cnt = 0.0;
#pragma unroll 3
            for (int j = crank * kGateLongThreads + threadIdx.x; j < P.nblocks; j += kGateCluster * kGateLongThreads) {
                unsigned long long u0 = 0, u1 = 0;
                const int lo = __ldg(P.blk_lo + j), hi = __ldg(P.blk_hi + j);
                for (int s = lo; s < hi; ++s) { u0 += s0[s]; if (C > 1) u1 += s1[s]; }
                const double z0 = (double)u0 * inv * P.scale, z1 = (double)u1 * inv * P.scale;
                const double l = -0.691 + 10.0 * log10(z0 + (C > 1 ? z1 : 0.0));
                const bool keep = pass == 0 ? (l >= -70.0) : (l > gamma_r && l > -70.0);
                if (keep) { a0 += z0; a1 += z1; cnt += 1.0; }
            }
            a0 = block_reduce_sum(a0, red);
            a1 = block_reduce_sum(a1, red);
            cnt = block_reduce_sum(cnt, red);
            if (threadIdx.x == 0) { part[pass][0] = a0; part[pass][1] = a1; part[pass][2] = cnt; }
            cluster.sync();
            a0 = a1 = cnt = 0.0;
            for (int r = 0; r < kGateCluster; ++r) {
                const double* q = cluster.map_shared_rank(&part[pass][0], r);
                a0 += q[0]; a1 += q[1]; cnt += q[2];
            }
            double m0, m1;
            if (cnt > 0.0) { m0 = a0 / cnt; m1 = a1 / cnt; }
            else if (pass == 0) { m0 = m1 = __longlong_as_double(0x7ff8000000000000LL); }
            else { m0 = m1 = 0.0; }
            const double v = -0.691 + 10.0 * log10(m0 + (C > 1 ? m1 : 0.0));
            if (pass == 0) gamma_r = v - 10.0; else result = v;
        }
        lufs = result;
        cluster.sync();                                           // no CTA leaves while its shared memory may still be read
    }
    if (crank == 0 && threadIdx.x == 0) {
        P.lufs[track] = lufs;
        if (P.target != nullptr) {
            double g = 1.0, gdb = 0.0;
            if (P.valid) {
                gdb = P.target[track] - lufs;
                gdb = fmin(fmax(gdb, -20.0), 20.0);
                g = pow(10.0, gdb / 20.0);
            }
            if (P.gain_row) for (int c = 0; c < C; ++c) P.gain_row[track * C + c] = g;
            if (P.gain_db) P.gain_db[track] = gdb;
        }
    }
}

}  // namespace mm
