// Persistent, software-pipelined zero-phase / causal IIR sweep kernel ("kernel (1)" of the north star), round 2: every WARP is
// an autonomous worker.
//
// A row is cut into SEGMENTS of `seglen` warp-tiles (kWT = 1024 samples each, in sweep order: 32 lanes x 32 consecutive
// samples).  One warp owns one segment at a time and walks its warp-tiles in order, so the state entering a warp-tile is simply
// the state that left the previous one (kept in the warp's slice of shared memory): no flags, no spinning, no inter-warp
// traffic, and -- new in round 2 -- no block barrier anywhere in the tile loop.  Round 1 ran 128-thread CTAs over 4096-sample
// tiles with four __syncthreads per tile and a single-stage input ring: every CTA exposed the full DRAM latency of its next
// tile's input once per tile, all four warps sat in the same phase, and the 4-section sweeps reached 0.36 issue slots per cycle
// at 12 warps per SM (ncu, profiles/r01_ncu_full_fwd_f4_dynamics.md).  Now a warp double-buffers its own 4 KB input slots
// (cp.async for tile t + 1 is in flight while tile t is scanned), warps drift apart freely, and loads, arithmetic and stores of
// different warps overlap on their own.
// The state entering a segment is rebuilt from a HALO: the `whalo` warp-tiles before the segment are read
// once more and only their zero-state end states are formed (pass 1) and chained,
//       s <- A^kWT s + aggregate(tile),
// which after the halo differs from the true state by A^(kWT*whalo) times the unknown older state --
// whalo is chosen on the host so that every entry of that matrix is below 1e-18 (the filters are
// stable: 4..12 warp-tiles for most sections, more for the lowest cut-offs at high sample rates).
//
// Per warp-tile:
//   cp.async ring (ST stages)   : the NIN input streams of the NEXT warp-tile land in shared memory while the
//                                  current one is being scanned; the x-domain aux streams an epilogue needs
//                                  are prefetched into L2 at tile start and read in the store phase
//   pass 1   (per lane)          : zero-state end state of its 32 samples, E = sum_j g[j] x_j   (2 DFMA/sample)
//   warp scan                    : 2x2 (4x4) state-transfer powers composed with warp shuffles, tables in smem
//   pass 2   (per lane)          : the DF2T recurrence from the resolved state, float32 results into smem
//   epilogue                     : coalesced float4 stores / recombination / dynamics / exciter, peak tracking
//
// Shared-memory warp-tiles are stored as 256 16-byte vectors with the 128-byte XOR swizzle
//   phys(chunk, u) = chunk * 8 + (u ^ (chunk & 7))
// so that both access patterns are bank-conflict free: the coalesced one (8 consecutive lanes touch one
// chunk) and the scan one (lane l walks chunk l).
#pragma once
#include "pointwise.cuh"
#include "common.cuh"

namespace mm {

#ifndef MM_COMBINE_F32
#define MM_COMBINE_F32 1
#endif
constexpr int kWT = 1024;            // samples per warp-tile (32 lanes x kS)
// Warps per sweep CTA.  The warps of a CTA share nothing but the read-only scan tables, so the CTA size only sets the granularity
// at which an SM's 228 KB of shared memory is handed out: the count that fits the most warps on an SM wins (ties: the smaller
// CTA).  `slice` = bytes of one warp's tiles, `tables` = bytes of the CTA's scan tables; see sweep_warps_on_sm for what else a CTA holds.
constexpr int sweep_warps_on_sm(int k, size_t slice, size_t tables) {
    // per CTA besides the tiles and tables: 1 KB reserved by the driver, up to 1 KB of padding in front of the 1024-byte aligned
    // dynamic part, and the static carries + mbarriers (<= 160 bytes per warp)
    const size_t cta = (size_t)k * slice + tables + 2304 + 160 * (size_t)k;
    if (cta > 232448) return 0;
    const int ctas = (int)(232448 / cta) > 32 ? 32 : (int)(232448 / cta);
    return k * ctas > 16 ? 16 : k * ctas;      // beyond ~16 warps the register file (90 to 200 registers per thread) is the limit anyway
}
// the smallest CTA that reaches the best warp count
constexpr int sweep_warps_per_cta(size_t slice, size_t tables) {
    int best_warps = 0;
    for (int k = 1; k <= 12; ++k) { const int w = sweep_warps_on_sm(k, slice, tables); if (w > best_warps) best_warps = w; }
    for (int k = 1; k <= 12; ++k) if (sweep_warps_on_sm(k, slice, tables) >= best_warps) return k;
    return 1;
}

constexpr int kTileVecs = kWT / 4;  // float4 per stream warp-tile
constexpr int kVecsPerLane = kTileVecs / 32;

// float -> double widening.  Measured on B200 (tools/microbench.cu): F2F runs at ~15 lanes/clk/SM, an
// integer-pipe emulation (shift/add/select, 6-7 instructions) is slower at saturation and costs issue
// slots this kernel does not have, so the conversion unit it is.
__device__ __forceinline__ double f2d_bits(float x) { return (double)x; }

template <int M> struct SmemTab {
    double Pw[5][M * M];
    double Plane[M * M][32];           // (A^32)^lane, TRANSPOSED: entry e of lane l at Plane[e][l] -- lane-contiguous, conflict free
                                       // (lane-major rows of 32 bytes cost a 4-way bank conflict per read: the 70-100 M conflicts
                                       // per launch ncu counted in both rounds)
    double Qpow[kNW + 1][M * M];       // Qpow[kNW] = A^kL carries a state across one tile
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ int swz(int v) { return (v & ~7) | ((v ^ (v >> 3)) & 7); }   // v = chunk * 8 + u

// ---- TMA (cp.async.bulk.tensor) + mbarrier: the input tiles of interior warp-tiles -----------------------------------------
// A stream is viewed as a 2-D tensor [128-byte lines][32 floats]; one warp-tile is the box {32 floats, 32 lines} = 4 KB, and the
// hardware's SWIZZLE_128B (16-byte chunk index XOR line index mod 8) is exactly swz() above, so a bulk tensor copy issued by ONE
// lane lands the tile in the layout both the coalesced and the per-lane accesses want -- instead of 8 LDGSTS per lane and stream.
struct alignas(64) TmaDesc { unsigned char bytes[128]; };     // CUtensorMap (opaque; filled by cuTensorMapEncodeTiled on the host)

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// waits for the phase with the given parity; bounded: a tensor map / byte-count mistake traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// smem -> global (bulk async-group completion): the finished float32 tiles of a storing sweep leave by one instruction per stream
__device__ __forceinline__ void tma_store_2d(const TmaDesc* desc, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n"
                 ::"l"(desc), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const TmaDesc* desc, unsigned long long* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(desc), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

template <int M> __device__ __forceinline__ void matvec_acc_s(const double* p, const double (&v)[M], double (&acc)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double s = acc[i];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(p[i * M + k], v[k], s);
        acc[i] = s;
    }
}

template <int M, int NF, int NSET = 1> struct Sweep2Args {
    TmaDesc tmap[NF];           // tensor maps of the input streams (first NIN used); valid when use_tma != 0
    TmaDesc tmap_out[NF];       // ... and of the output streams of a storing sweep (EPI_STORE); valid when use_tma & 2
    SweepArgs<M, NF> a[NSET];   // NSET is always 1 (the split-section experiment of round 1 is gone)
    int seglen;           // live warp-tiles per segment
    int nseg;             // segments per row
    int whalo;            // halo warp-tiles read before a segment (max over the sweep's sections)
    int use_tma;          // bit 0: interior input tiles arrive by cp.async.bulk.tensor (else per-lane cp.async); bit 1: interior
                          // output tiles of a storing sweep leave the same way
    int tma_shift;        // floats the tensor maps' base is shifted by (backward sweeps: tile origins are = qend mod 32)
    long long tma_row_lines;   // 128-byte lines per row (stride / 32)
};

template <int M, int NF, int NIN, int DIR, int EPI, int NAUX, int ST>
struct Sweep2Cfg {
    static constexpr int kExtra = NF > NIN ? NF - NIN : 0;
    static constexpr size_t kTileBytes = (size_t)kWT * sizeof(float);
    // per warp: ST input slots of NIN streams, then the staging tiles of the outputs that do not fit the input slots
    static constexpr size_t kWarpBytes = ((size_t)ST * NIN + kExtra) * kTileBytes;
    static constexpr size_t kExtraOff = (size_t)ST * NIN * kTileBytes;     // inside a warp's slice
    static constexpr int kSW = sweep_warps_per_cta(kWarpBytes, NF * sizeof(SmemTab<M>));   // warps per CTA
    static constexpr int kThreads = 32 * kSW;
    static constexpr size_t kTabOff = kSW * kWarpBytes;  // the x-domain aux streams of an epilogue are read straight from
                                                         // global memory in the (coalesced) store phase: no smem tiles
    static constexpr size_t kBytes = kTabOff + NF * sizeof(SmemTab<M>);
    // CTAs (of kSW warps) per SM the shared-memory footprint allows (227 KB usable): the register allocator is held to it
    static constexpr int kFit = (int)((227u * 1024u) / (kBytes + 1024u));
    // the dynamics epilogues (four band chains + maximizer per sample) need ~128 registers: 16 warps/SM beat 24 with spills
    static constexpr int kCapW = ((EPI == EPI_DYNAMICS || EPI == EPI_DYNAMICS_GEN) ? 16 : 24) / kSW;
    static constexpr int kCap = kCapW < 1 ? 1 : kCapW;
    static constexpr int kMinBlocks = kFit < 1 ? 1 : (kFit > kCap ? kCap : kFit);
};
// input ring depth per instantiation: double-buffered while a warp's slice stays within 20 KB (10+ warps per SM by shared memory).
// The four-input backward sweep would need 32 KB per warp (6 warps per SM: measured 8.7 ms against 6.5 ms single-stage with 10
// warps); it stays single-stage and pulls the next tile's lines into L2 while the current one is scanned.
template <int NF, int NIN> struct SweepStages {
    static constexpr int kExtra = NF > NIN ? NF - NIN : 0;
    static constexpr int value = (2 * NIN + kExtra) * 4 <= 20 ? 2 : 1;
};

template <int M, int NF, int SW> struct Scratch2 {
    double carry[SW][2][NF][M];      // per warp: the state entering warp-tile t lives in carry[warp][t & 1]
};

// prologue helpers (see common.cuh PRO_*)
__device__ __forceinline__ float pro1(int mode, float x, float subf, float mulf, double muld) {
    if (mode == PRO_SUBMUL_F32) return __fmul_rn(__fsub_rn(x, subf), mulf);
    if (mode == PRO_MUL_F64) return (float)(f2d_bits(x) * muld);
    return x;
}

// NF32: the first NF32 sections run pass 2 in float32 on their balanced realization (tables and g in those
// coordinates), the others in float64 DF2T.  Pass 1 and the scan are float64 for both.
template <int M, int NF, int NIN, int DIR, int EPI, int NAUX, int ST, int NF32, int NSET = 1>
__global__ void __launch_bounds__(Sweep2Cfg<M, NF, NIN, DIR, EPI, NAUX, ST>::kThreads, Sweep2Cfg<M, NF, NIN, DIR, EPI, NAUX, ST>::kMinBlocks) sweep2_kernel(const __grid_constant__ Sweep2Args<M, NF, NSET> PP) {
    typedef Sweep2Cfg<M, NF, NIN, DIR, EPI, NAUX, ST> Cfg;
    constexpr int kSW = Cfg::kSW, kSweepThreads = Cfg::kThreads;
    const SweepArgs<M, NF>& P = PP.a[0];
    extern __shared__ __align__(1024) unsigned char smraw[];       // SWIZZLE_128B destinations need 1024-byte alignment
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(8) unsigned long long mbar[kSW][ST];
    unsigned mphase = 0;                                           // bit s: parity the next wait on this warp's slot s expects
    if ((PP.use_tma & 1) && lane == 0) {
        if ((unsigned)__cvta_generic_to_shared(smraw) & 1023u) __trap();      // the swizzled boxes would land shifted: fail loudly
#pragma unroll
        for (int s = 0; s < ST; ++s) mbar_init(&mbar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    float* ring = reinterpret_cast<float*>(smraw + (size_t)warp * Cfg::kWarpBytes);            // this warp's input slots
    float* extra = reinterpret_cast<float*>(smraw + (size_t)warp * Cfg::kWarpBytes + Cfg::kExtraOff);
    SmemTab<M>* tab = reinterpret_cast<SmemTab<M>*>(smraw + Cfg::kTabOff);
    __shared__ Scratch2<M, NF, kSW> sh;
    constexpr int MM = M * M;

    // ---- one-time: scan tables into shared memory (shared by the CTA's warps, read-only afterwards) ----
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
        const double* g = P.tab[f];
        double* d = reinterpret_cast<double*>(&tab[f]);
        // device table: Pw [5][MM], Plane [32][MM], Qpow [kNW + 1][MM], contiguous; Plane is transposed on the way in
        for (int i = tid; i < (5 + 32 + kNW + 1) * MM; i += kSweepThreads) {
            const int j = i - 5 * MM;
            const int dsti = (j >= 0 && j < 32 * MM) ? 5 * MM + (j % MM) * 32 + j / MM : i;
            d[dsti] = __ldg(g + i);
        }
    }
    __syncthreads();                                        // the only block barrier of the kernel

    const long long q_first = kLead - P.pad;
    const long long q_last = kLead + P.n + P.pad - 1;
    const long long qend = bwd_qend(q_last);                // BWD: one past the last tile-0 position (line aligned)
    const int dead0 = (DIR > 0) ? (int)q_first : (int)(qend - 1 - q_last);
    // store range of this sweep (inclusive)
    const long long st_lo = (DIR > 0) ? q_first : (long long)kLead;
    const long long st_hi = (DIR > 0) ? q_last : (long long)(kLead + P.n - 1);

    auto tile_origin = [&](int tile) -> long long {
        return (DIR > 0) ? (long long)tile * kWT : qend - (long long)(tile + 1) * kWT;
    };
    // can the tile's inputs be fetched with unconditional 16-byte async copies?
    auto fast_in = [&](long long lo) -> bool {
        return (DIR > 0) ? (lo >= kLead && lo + kWT <= kLead + P.n) : (lo >= q_first && lo + kWT - 1 <= q_last);
    };
    auto fast_out = [&](long long lo) -> bool { return lo >= st_lo && lo + kWT - 1 <= st_hi; };

    // ---- loaders ------------------------------------------------------------------------------------
    auto issue_inputs = [&](int row, int tile, int slot) {
        const long long lo = tile_origin(tile);
        const size_t rowoff = (size_t)row * (size_t)P.stride;
        if (fast_in(lo) && (PP.use_tma & 1)) {
            // one lane arms the slot's mbarrier with the byte count and issues one bulk tensor copy per stream
            if (lane == 0) {
                fence_proxy_async();                       // earlier generic-proxy accesses of this slot (ordered by the caller's __syncwarp)
                mbar_expect_tx(&mbar[warp][slot], NIN * (unsigned)Cfg::kTileBytes);
                const int line = (int)((long long)row * PP.tma_row_lines + (lo - PP.tma_shift) / 32);
#pragma unroll
                for (int s = 0; s < NIN; ++s) tma_load_2d(ring + ((size_t)slot * NIN + s) * kWT, &PP.tmap[s], &mbar[warp][slot], 0, line);
            }
        } else if (fast_in(lo)) {
#pragma unroll
            for (int s = 0; s < NIN; ++s) {
                const float* src = P.in[s] + rowoff + lo + 4 * lane;
                float* dst = ring + ((size_t)slot * NIN + s) * kWT;
#pragma unroll
                for (int r = 0; r < kVecsPerLane; ++r) cp_async16(dst + 4 * swz(lane + 32 * r), src + 128 * r);
            }
        } else {
            // edge tile: synchronous, with prologue, scipy's odd extension and dead zeros applied here
            float subf = 0.f, mulf = 1.f;
            double muld = 1.0;
            if (P.pro_mode != PRO_NONE) {
                if (P.pro_sub) subf = (float)__ldg(P.pro_sub + row);
                if (P.pro_mul) { muld = __ldg(P.pro_mul + row); mulf = (float)muld; }
            }
#pragma unroll 1
            for (int s = 0; s < NIN; ++s) {
                const float* src = P.in[s] + rowoff;
                float* dst = ring + ((size_t)slot * NIN + s) * kWT;
#pragma unroll 1
                for (int r = 0; r < kVecsPerLane; ++r) {
                    const int v = lane + 32 * r;
                    const long long q = lo + 4 * v;
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q + 3 >= q_first && q <= q_last) {
                        if (DIR > 0) {
                            const float x_lo = pro1(P.pro_mode, src[kLead], subf, mulf, muld);
                            const float x_hi = pro1(P.pro_mode, src[kLead + P.n - 1], subf, mulf, muld);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const long long i = q + c - kLead;
                                float e = 0.f;
                                if (i >= 0 && i < P.n) e = pro1(P.pro_mode, src[kLead + i], subf, mulf, muld);
                                else if (i < 0 && i >= -(long long)P.pad)
                                    e = __fsub_rn(__fmul_rn(2.f, x_lo), pro1(P.pro_mode, src[kLead - i], subf, mulf, muld));
                                else if (i >= P.n && i < P.n + P.pad)
                                    e = __fsub_rn(__fmul_rn(2.f, x_hi), pro1(P.pro_mode, src[kLead + 2 * (P.n - 1) - i], subf, mulf, muld));
                                setcomp4(val, c, e);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const long long qq = q + c;
                                setcomp4(val, c, (qq >= q_first && qq <= q_last) ? src[qq] : 0.f);
                            }
                        }
                    }
                    *reinterpret_cast<float4*>(dst + 4 * swz(v)) = val;
                }
            }
        }
    };
    // ---- segment loop: every warp is its own worker ------------------------------------------------------
    const int items = P.rows * PP.nseg;
#pragma unroll 1
    for (int item = blockIdx.x * kSW + warp; item < items; item += gridDim.x * kSW) {
    const int row = P.row_map ? __ldg(P.row_map + item % P.rows) : item % P.rows;
    const int seg = item / P.rows;
    const int t_live = seg * PP.seglen;
    const int t_end = min(P.ntiles, t_live + PP.seglen);
    const int t_first = max(0, t_live - PP.whalo);
    const size_t rowoff = (size_t)row * (size_t)P.stride;
    // per-row parameters of a mixed-preset batch (one launch per stage over the rows whose style fires it)
    double w0d = P.w[0], excg = P.exc_gain;
    float w0f = P.w32[0];
    bool pk_on = P.peak != nullptr;
    if (EPI == EPI_COMBINE && NF == 1 && P.w_row) { w0d = __ldg(P.w_row + row); w0f = (float)w0d; }
    if (EPI == EPI_EXCITER && P.exc_row) excg = __ldg(P.exc_row + row);
    if (EPI != EPI_STORE && pk_on && P.peak_row) pk_on = P.peak_row[row] != 0;
    __syncwarp();                                      // the previous item is completely done (this warp's smem, carry)
    if (lane < NF * M) sh.carry[warp][t_first & 1][lane / M][lane % M] = 0.0;
    int slot = 0;
    issue_inputs(row, t_first, 0);
    cp_async_commit();

#pragma unroll 1
    for (int tile = t_first; tile < t_end; ++tile) {
        const bool live = tile >= t_live;
        const long long tile_lo = tile_origin(tile);
        if (EPI == EPI_STORE && (PP.use_tma & 2) && lane == 0) tma_store_wait_read();   // last tile's bulk stores have read their tiles
        if (ST > 1) {
            __syncwarp();                              // every lane is done with the other slot (stored from it last tile)
            if (tile + 1 < t_end) issue_inputs(row, tile + 1, slot ^ 1);
            cp_async_commit();                         // group: inputs(tile + 1)
            cp_async_wait<1>();                        // inputs(tile) have landed (this lane's part)
            if ((PP.use_tma & 1) && fast_in(tile_lo)) { mbar_wait(&mbar[warp][slot], (mphase >> slot) & 1u); mphase ^= 1u << slot; }
        } else {
            cp_async_wait<0>();
            if ((PP.use_tma & 1) && fast_in(tile_lo)) { mbar_wait(&mbar[warp][0], mphase & 1u); mphase ^= 1u; }
            if (tile + 1 < t_end && (lane & 7) == 0) {
                // single-stage ring: the next tile's inputs can only be fetched once this tile has left the buffer; pull their
                // lines into L2 now so that fetch is short (one request per 128 bytes)
                const long long nlo = tile_origin(tile + 1);
                if (fast_in(nlo)) {
#pragma unroll
                    for (int s = 0; s < NIN; ++s) {
                        const float* np_ = P.in[s] + rowoff + nlo + 4 * lane;
#pragma unroll
                        for (int r = 0; r < kVecsPerLane; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + 128 * r));
                    }
                }
            }
        }
        __syncwarp();                                  // inputs(tile) and the carried state visible to all lanes
        if (EPI != EPI_STORE && live && (lane & 7) == 0) {
            // the store phase will read this tile's aux samples: pull their lines into L2 now (one request per 128 bytes)
#pragma unroll
            for (int s = 0; s < NAUX; ++s) {
                const float* ap = P.aux[s] + rowoff + tile_lo + 4 * lane;
#pragma unroll
                for (int r = 0; r < kVecsPerLane; ++r)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 128 * r));
            }
        }

        float* tin = ring + (size_t)slot * NIN * kWT;
        const bool in_fast = fast_in(tile_lo);
        const bool inject = (tile == 0) && (P.pad > 0);
        const int dead = (tile == 0) ? dead0 : 0;
        const int d0 = dead;                           // dead < kS always (pad <= 15, lead 32)
        const bool inj_thread = inject && lane == 0;

        // prologue constants: edge tiles were transformed by their loader already
        int pmode = PRO_NONE;
        float subf = 0.f, mulf = 1.f;
        double muld = 1.0;
        if (DIR > 0 && in_fast && P.pro_mode != PRO_NONE) {
            pmode = P.pro_mode;
            if (P.pro_sub) subf = (float)__ldg(P.pro_sub + row);
            if (P.pro_mul) { muld = __ldg(P.pro_mul + row); mulf = (float)muld; }
        }

        const int chunk = (DIR > 0) ? lane : (31 - lane);
        const int cbase = chunk * 32;                  // float index of this lane's chunk
        const int cx = (chunk & 7) << 2;               // float-index XOR of the swizzle
        // address of logical vec u of this chunk in stream buffer b: b + cbase + ((4u) ^ cx)

        // ---- pass 1 -----------------------------------------------------------------------------------
        // packed float32 sections (pairs 2p, 2p+1 below 2 NP) accumulate their zero-state end states in float32
        // too: the balanced coordinates keep that as accurate as the float32 pass 2 itself (design.h)
        constexpr int NP = (M == 2) ? NF32 / 2 : 0;
        double E[NF][M];
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int i = 0; i < M; ++i) E[f][i] = 0.0;
        float2 Ep[NP > 0 ? NP : 1][2];
#pragma unroll
        for (int p = 0; p < NP; ++p) Ep[p][0] = Ep[p][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kS / 4; ++u) {
            const int uu = (DIR > 0) ? u : (kS / 4 - 1 - u);
            float4 xv[NIN];
#pragma unroll
            for (int s = 0; s < NIN; ++s) {
                float* p = tin + (size_t)s * kWT + cbase + ((4 * uu) ^ cx);
                xv[s] = *reinterpret_cast<const float4*>(p);
                if (pmode == PRO_SUBMUL_F32) {
                    xv[s].x = __fmul_rn(__fsub_rn(xv[s].x, subf), mulf); xv[s].y = __fmul_rn(__fsub_rn(xv[s].y, subf), mulf);
                    xv[s].z = __fmul_rn(__fsub_rn(xv[s].z, subf), mulf); xv[s].w = __fmul_rn(__fsub_rn(xv[s].w, subf), mulf);
                    *reinterpret_cast<float4*>(p) = xv[s];
                } else if (pmode == PRO_MUL_F64) {
                    xv[s].x = (float)((double)xv[s].x * muld); xv[s].y = (float)((double)xv[s].y * muld);
                    xv[s].z = (float)((double)xv[s].z * muld); xv[s].w = (float)((double)xv[s].w * muld);
                    *reinterpret_cast<float4*>(p) = xv[s];
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = (DIR > 0) ? c : (3 - c);
                const int j = 4 * u + c;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const float2 X = make_float2(comp4(xv[NIN == 1 ? 0 : 2 * p], cc), comp4(xv[NIN == 1 ? 0 : 2 * p + 1], cc));
                    Ep[p][0] = ffma2(P.pr[p].g[j][0], X, Ep[p][0]);
                    Ep[p][1] = ffma2(P.pr[p].g[j][1], X, Ep[p][1]);
                }
#pragma unroll
                for (int f = 2 * NP; f < NF; ++f) {
                    const double xd = f2d_bits(comp4(xv[NIN == 1 ? 0 : f], cc));
#pragma unroll
                    for (int i = 0; i < M; ++i) E[f][i] = fma(P.f[f].g[j][i], xd, E[f][i]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            E[2 * p][0] = (double)Ep[p][0].x; E[2 * p + 1][0] = (double)Ep[p][0].y;
            E[2 * p][M > 1 ? 1 : 0] = (double)Ep[p][1].x; E[2 * p + 1][M > 1 ? 1 : 0] = (double)Ep[p][1].y;
        }
        if (inj_thread) {
            // scipy's filtfilt start: state zi * x_first, injected `d0` samples into this chunk
            const int mi = (DIR > 0) ? d0 : (kS - 1 - d0);
            const int off = cbase + (((mi >> 2) << 2) ^ cx) + (mi & 3);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const float x0 = tin[(size_t)(NIN == 1 ? 0 : f) * kWT + off];
                double si[M];
#pragma unroll
                for (int i = 0; i < M; ++i) si[i] = __ldg(P.tab[f] + Tab<M>::Zi + i) * (double)x0;
                matvec_acc<M>(P.tab[f] + Tab<M>::Apow + (kS - d0) * MM, si, E[f]);
            }
        }

        // ---- warp scan: branch-free (lanes below the stride shuffle in zeros), sections interleaved --------
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const bool act = lane >= (1 << d);
            double pe[NF][M];
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double v = shfl_up_d(E[f][i], 1 << d);
                    pe[f][i] = act ? v : 0.0;
                }
#pragma unroll
            for (int f = 0; f < NF; ++f) matvec_acc_s<M>(tab[f].Pw[d], pe[f], E[f]);
        }

        // ---- state leaving this warp-tile = zero-state aggregate + A^kWT * state entering it -------------------
        if (lane == 31) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                double ag[M], cin[M];
#pragma unroll
                for (int i = 0; i < M; ++i) { ag[i] = E[f][i]; cin[i] = sh.carry[warp][tile & 1][f][i]; }
                matvec_acc_s<M>(tab[f].Qpow[1], cin, ag);
#pragma unroll
                for (int i = 0; i < M; ++i) sh.carry[warp][(tile + 1) & 1][f][i] = ag[i];
            }
        }
        if (!live) {   // halo tile: only its contribution to the state was needed
            if (ST > 1) slot ^= 1;
            else {
                __syncwarp();
                if (tile + 1 < t_end) issue_inputs(row, tile + 1, 0);
                cp_async_commit();
            }
            continue;
        }

        // ---- incoming state of this lane ------------------------------------------------------------------------
        double z[NF][M];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            double C[M];
#pragma unroll
            for (int i = 0; i < M; ++i) C[i] = sh.carry[warp][tile & 1][f][i];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const double up = shfl_up_d(E[f][i], 1);
                z[f][i] = (lane > 0) ? up : 0.0;
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double acc = z[f][i];
#pragma unroll
                for (int k = 0; k < M; ++k) acc = fma(tab[f].Plane[i * M + k][lane], C[k], acc);
                z[f][i] = acc;
            }
        }

        // ---- pass 2 (+ the recombining epilogue, evaluated on the float64 section outputs) --------------------
        constexpr int NOUT = (EPI == EPI_STORE) ? NF : 1;
        constexpr int NWR = (EPI == EPI_STORE) ? NF : ((EPI == EPI_DYNAMICS || EPI == EPI_DYNAMICS_GEN) ? 2 : 1);   // tiles pass 2 writes
        float* tout[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) tout[f] = (f < NIN) ? (tin + (size_t)f * kWT) : (extra + (size_t)(f - NIN) * kWT);
        const bool out_fast = fast_out(tile_lo);
        const bool pk_fast = tile_lo >= P.pk_lo && tile_lo + kWT - 1 <= P.pk_hi;
        float aux_subf = 0.f, aux_mulf = 1.f;
        double aux_muld = 1.0;
        const int aux_pmode = (EPI != EPI_STORE && P.aux_pro) ? P.pro_mode : PRO_NONE;
        if (EPI != EPI_STORE && P.aux_pro && P.pro_mode != PRO_NONE) {
            if (P.pro_sub) aux_subf = (float)__ldg(P.pro_sub + row);
            if (P.pro_mul) { aux_muld = __ldg(P.pro_mul + row); aux_mulf = (float)aux_muld; }
        }
        float par_mix = 0.f, par_one_minus = 1.f;
        if (EPI == EPI_DYNAMICS_GEN && P.dyn.par_mix) {
            const double mixd = __ldg(P.dyn.par_mix + row);
            par_mix = (float)mixd;
            par_one_minus = (float)(1.0 - mixd);      // Python float (1.0 - mix), then weak-cast to float32
        }
        float pk = 0.f;
        // A recombining epilogue is split in two.  Pass 2 STAGES, per sample, what depends on the section outputs
        // (float32, into the tile buffers the inputs came from); the store phase reads the staged values back
        // coalesced, fetches the x-domain aux samples of the same positions straight from global memory and
        // finishes the arithmetic.  Nothing of an aux stream ever sits in shared memory.
        //   EPI_COMBINE     stage S = sum_f w_f y_f                      finish (wc xa + S) * trim [clip]
        //   EPI_EXCITER     stage t = (sat(hf) - hf) * gain / 4          finish xa + t
        //   EPI_DYNAMICS*   stage chain(band 2), chain(band 3)           finish chain(b1) + . + . + chain(b4), maximize, limit
        constexpr int NSTAGE = (EPI == EPI_DYNAMICS || EPI == EPI_DYNAMICS_GEN) ? 2 : 1;
        auto stage_value = [&](const float (&yf)[NF], const double (&yd)[NF], float (&st)[2]) {
            auto yflt = [&](int f) -> float { return f < NF32 ? yf[f] : (float)yd[f]; };
            st[1] = 0.f;
            if (EPI == EPI_COMBINE) {
                // pipeline.py:273 / :603-606 / :1431: the weighted sum of the filtered components; float32 sections in
                // float32, float64 sections in float64, rounded to float32 once
                float accf = 0.f;
#pragma unroll
                for (int f = 0; f < NF32; ++f) accf = fmaf(NF == 1 ? w0f : P.w32[f], yf[f], accf);
                if (NF32 < NF) {
                    double acc = (double)accf;
#pragma unroll
                    for (int f = NF32; f < NF; ++f) acc = fma(NF == 1 ? w0d : P.w[f], yd[f], acc);
                    accf = (float)acc;
                }
                st[0] = accf;
            } else if (EPI == EPI_EXCITER) {
                const float hf = yflt(0);
                const float sat = exciter_sat(hf, P.exc_mode, (float)P.exc_k);
                st[0] = (float)((double)(sat - hf) * (excg * 0.25));
            } else if (EPI == EPI_DYNAMICS) {   // y0 = band 2, y1 = band 3; downward knees only
                st[0] = band_chain(yflt(0), P.dyn.band[1]);
                st[1] = band_chain(yflt(NF > 1 ? 1 : 0), P.dyn.band[2]);
            } else if (EPI == EPI_DYNAMICS_GEN) {
                st[0] = band_chain_gen(yflt(0), P.dyn.band[1]);
                st[1] = band_chain_gen(yflt(NF > 1 ? 1 : 0), P.dyn.band[2]);
            }
        };
        // xa = aux0 (prologue applied), a1c = aux1
        auto final_value = [&](float s0, float s1, float xa, float a1c) -> float {
            if (EPI == EPI_COMBINE) {
                float res;
                if (MM_COMBINE_F32 && NF32 == NF) {
                    // every section ran in float32 (the staged sum is a float32 value already): recombine in float32 as well --
                    // one fused multiply-add and one multiply, within two float32 ulps of the float64 expression below, without
                    // its three float<->double conversions per sample (the conversion unit runs at 16 lanes per clock and SM)
                    res = __fmul_rn(fmaf(P.wc32, xa, s0), P.trim32);
                } else {
                    res = (float)(fma(P.wc, (double)xa, (double)s0) * P.trim);              // float64 recombination, one cast
                }
                return P.epi_clip ? fminf(fmaxf(res, -1.f), 1.f) : res;
            } else if (EPI == EPI_EXCITER) {
                return (float)((double)s0 + (double)xa);
            } else if (EPI == EPI_DYNAMICS) {   // aux0 = band 1, aux1 = band 4 (pipeline.py:466-481, :484-492, :636)
                float sacc = band_chain(xa, P.dyn.band[0]);
                sacc = __fadd_rn(sacc, s0);
                sacc = __fadd_rn(sacc, s1);
                sacc = __fadd_rn(sacc, band_chain(a1c, P.dyn.band[3]));
                return maximize_limit(sacc, P.dyn);
            } else {                             // EPI_DYNAMICS_GEN: upward bands and/or the v1 parallel compressor
                float sacc = band_chain_gen(xa, P.dyn.band[0]);
                sacc = __fadd_rn(sacc, s0);
                sacc = __fadd_rn(sacc, s1);
                sacc = __fadd_rn(sacc, band_chain_gen(a1c, P.dyn.band[3]));
                float res = maximize_limit(sacc, P.dyn);
                if (par_mix >= 0.01f) res = parallel_compress(res, par_mix, par_one_minus, P.dyn);
                return res;
            }
        };
        // float32 sections start every chunk from the float64-resolved state, mapped into the rescaled coordinates (B = 1) of
        // ss32_step.  Scalar steps: packing the two sections of a pair into FFMA2 operands costs more register moves than the
        // packed arithmetic saves (SASS of the loudness kernel: 26 instructions per sample packed, 15 scalar).
        // Groups of four samples unrolled in pass 2.  The recombining kernels are 4000 instructions long and their warps drift apart
        // through it: ncu showed `no_instruction` (instruction-cache misses) at 0.9 cycles per issued instruction in the dynamics
        // sweep.  Measured per launch (64 x 180 s): dynamics 4.41 / 4.04 / 3.96 ms at 4 / 2 / 1 groups, four-section combine
        // 4.77 / 4.64 / 4.75, two-section combine 2.88 / 2.91 / 2.96.
        constexpr int kP2Unroll = (EPI == EPI_DYNAMICS || EPI == EPI_DYNAMICS_GEN) ? 1 : (NF >= 4 ? 2 : 4);
        float sf[NF32 > 0 ? NF32 : 1][M];
#pragma unroll
        for (int f = 0; f < NF32; ++f)
#pragma unroll
            for (int i = 0; i < M; ++i) sf[f][i] = (float)(z[f][i] * (double)P.f[f].dn32[i]);
        if (!inj_thread) {
#pragma unroll (EPI == EPI_STORE ? kS / 4 : kP2Unroll)     // storing sweeps: full unroll measured best (4-section forward: 7.51 / 7.59 / 7.76 ms at 8 / 4 / 2)
            for (int u = 0; u < kS / 4; ++u) {
                const int uu = (DIR > 0) ? u : (kS / 4 - 1 - u);
                const int off = cbase + ((4 * uu) ^ cx);
                float4 xv[NIN];
#pragma unroll
                for (int s = 0; s < NIN; ++s) xv[s] = *reinterpret_cast<const float4*>(tin + (size_t)s * kWT + off);
                float4 yv[NWR];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cc = (DIR > 0) ? c : (3 - c);
                    float yf[NF];
                    double yd[NF];
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        const float xs = comp4(xv[NIN == 1 ? 0 : f], cc);
                        if (f < NF32) { yf[f] = ss32_step<M>(P.f[f], xs, sf[f < NF32 ? f : 0]); yd[f] = 0.0; }
                        else { yd[f] = df2t_step<M>(P.f[f], f2d_bits(xs), z[f]); yf[f] = 0.f; }
                    }
                    if (EPI == EPI_STORE) {
#pragma unroll
                        for (int f = 0; f < NF; ++f) setcomp4(yv[f < NOUT ? f : 0], cc, f < NF32 ? yf[f] : (float)yd[f]);
                    } else {
                        float st[2];
                        stage_value(yf, yd, st);
                        setcomp4(yv[0], cc, st[0]);
                        if (NSTAGE > 1) setcomp4(yv[NWR > 1 ? 1 : 0], cc, st[1]);
                    }
                }
#pragma unroll
                for (int f = 0; f < NWR; ++f) *reinterpret_cast<float4*>(tout[f] + off) = yv[f];
            }
        } else {
            // the one lane of tile 0 that starts from zi * x_first after `d0` dead samples
#pragma unroll 1
            for (int j = 0; j < kS; ++j) {
                const int mi = (DIR > 0) ? j : (kS - 1 - j);
                const int off = cbase + (((mi >> 2) << 2) ^ cx) + (mi & 3);
                float xs[NIN];
#pragma unroll
                for (int s = 0; s < NIN; ++s) xs[s] = tin[(size_t)s * kWT + off];
                if (j == d0) {
#pragma unroll
                    for (int f = 0; f < NF; ++f)
#pragma unroll
                        for (int i = 0; i < M; ++i) {
                            z[f][i] = __ldg(P.tab[f] + Tab<M>::Zi + i) * (double)xs[NIN == 1 ? 0 : f];
                            if (f < NF32) sf[f < NF32 ? f : 0][i] = (float)(z[f][i] * (double)P.f[f].dn32[i]);
                        }
                }
                float yf[NF];
                double yd[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float xq = xs[NIN == 1 ? 0 : f];
                    if (f < NF32) { yf[f] = ss32_step<M>(P.f[f], xq, sf[f < NF32 ? f : 0]); yd[f] = 0.0; }
                    else { yd[f] = df2t_step<M>(P.f[f], (double)xq, z[f]); yf[f] = 0.f; }
                }
                if (EPI == EPI_STORE) {
#pragma unroll
                    for (int f = 0; f < NF; ++f) tout[f][off] = f < NF32 ? yf[f] : (float)yd[f];
                } else {
                    float st[2];
                    stage_value(yf, yd, st);
                    tout[0][off] = st[0];
                    if (NSTAGE > 1) tout[NWR > 1 ? 1 : 0][off] = st[1];
                }
            }
        }
        const bool bulk_out = EPI == EPI_STORE && (PP.use_tma & 2) && out_fast;
        if (bulk_out) fence_proxy_async();             // this lane's staged results become visible to the async proxy
        __syncwarp();                                  // results visible to the whole warp

        // ---- store: coalesced.  EPI_STORE: NOUT finished float32 streams.  Recombining epilogues: finish here ------------
        if (EPI != EPI_STORE) {
            const float* ax0 = P.aux[0] + rowoff;
            const float* ax1 = (NAUX > 1) ? (P.aux[1] + rowoff) : ax0;
            if (out_fast) {
                const size_t go0 = (size_t)tile_lo + 4 * lane;
                // the dynamics epilogue finishes four band chains and the maximizer per sample: its store phase runs as two rolled
                // halves (half the code; 4.17 -> 4.04 ms per launch), the lighter epilogues stay unrolled (rolling cost them 4 %)
                constexpr int kStoreSplit = (EPI == EPI_DYNAMICS || EPI == EPI_DYNAMICS_GEN) ? 2 : 1;
                constexpr int kPart = kVecsPerLane / kStoreSplit;
#pragma unroll 1
                for (int part = 0; part < kStoreSplit; ++part) {
                float4 xa[kPart], xb[kPart];
#pragma unroll
                for (int rr = 0; rr < kPart; ++rr) {           // all aux loads of the part first: independent requests in flight
                    const int r = part * kPart + rr;
                    xa[rr] = __ldcs(reinterpret_cast<const float4*>(ax0 + go0 + 128 * r));
                    if (NAUX > 1) xb[rr] = __ldcs(reinterpret_cast<const float4*>(ax1 + go0 + 128 * r));
                }
#pragma unroll
                for (int rr = 0; rr < kPart; ++rr) {
                    const int r = part * kPart + rr;
                    float4 a0 = xa[rr];
                    if (aux_pmode == PRO_SUBMUL_F32) {
                        a0.x = __fmul_rn(__fsub_rn(a0.x, aux_subf), aux_mulf); a0.y = __fmul_rn(__fsub_rn(a0.y, aux_subf), aux_mulf);
                        a0.z = __fmul_rn(__fsub_rn(a0.z, aux_subf), aux_mulf); a0.w = __fmul_rn(__fsub_rn(a0.w, aux_subf), aux_mulf);
                    } else if (aux_pmode == PRO_MUL_F64) {
                        a0.x = (float)((double)a0.x * aux_muld); a0.y = (float)((double)a0.y * aux_muld);
                        a0.z = (float)((double)a0.z * aux_muld); a0.w = (float)((double)a0.w * aux_muld);
                    }
                    const int so = 4 * swz(lane + 32 * r);
                    const float4 s0 = *reinterpret_cast<const float4*>(tout[0] + so);
                    float4 s1 = s0;
                    if (NSTAGE > 1) s1 = *reinterpret_cast<const float4*>(tout[NWR > 1 ? 1 : 0] + so);
                    const float4 a1 = (NAUX > 1) ? xb[rr] : a0;
                    float4 o;
                    o.x = final_value(s0.x, s1.x, a0.x, a1.x); o.y = final_value(s0.y, s1.y, a0.y, a1.y);
                    o.z = final_value(s0.z, s1.z, a0.z, a1.z); o.w = final_value(s0.w, s1.w, a0.w, a1.w);
                    if (pk_fast) {
                        pk = fmaxf(fmaxf(pk, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
                    } else {
                        const long long q = tile_lo + 4 * (lane + 32 * r);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (q + c >= P.pk_lo && q + c <= P.pk_hi) pk = fmaxf(pk, fabsf(comp4(o, c)));
                    }
                    __stcs(reinterpret_cast<float4*>(P.out[0] + rowoff + go0 + 128 * r), o);
                }
                }
            } else {
#pragma unroll 1
                for (int r = 0; r < kVecsPerLane; ++r) {
                    const int v = lane + 32 * r;
                    const long long q = tile_lo + 4 * v;
                    if (q + 3 < st_lo || q > st_hi) continue;
                    const int so = 4 * swz(v);
                    const float4 s0 = *reinterpret_cast<const float4*>(tout[0] + so);
                    float4 s1 = s0;
                    if (NSTAGE > 1) s1 = *reinterpret_cast<const float4*>(tout[NWR > 1 ? 1 : 0] + so);
                    float* dst = P.out[0] + rowoff;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (q + c < st_lo || q + c > st_hi) continue;
                        float xa0 = ax0[q + c];
                        if (aux_pmode != PRO_NONE) xa0 = pro1(aux_pmode, xa0, aux_subf, aux_mulf, aux_muld);
                        const float res = final_value(comp4(s0, c), comp4(s1, c), xa0, NAUX > 1 ? ax1[q + c] : 0.f);
                        if (q + c >= P.pk_lo && q + c <= P.pk_hi) pk = fmaxf(pk, fabsf(res));
                        dst[q + c] = res;
                    }
                }
            }
        } else if (bulk_out) {
            // interior tile of a storing sweep: one bulk tensor store per stream, issued by one lane
            if (lane == 0) {
                const int line = (int)((long long)row * PP.tma_row_lines + (tile_lo - PP.tma_shift) / 32);
#pragma unroll
                for (int f = 0; f < NOUT; ++f) tma_store_2d(&PP.tmap_out[f], tout[f], 0, line);
                tma_store_commit();
            }
        } else if (out_fast) {
            // interior tile: every vector is complete
            const size_t go0 = rowoff + (size_t)tile_lo + 4 * lane;
#pragma unroll
            for (int r = 0; r < kVecsPerLane; ++r) {
                const int so = 4 * swz(lane + 32 * r);
#pragma unroll
                for (int f = 0; f < NOUT; ++f)
                    __stcs(reinterpret_cast<float4*>(P.out[f] + go0 + 128 * r), *reinterpret_cast<const float4*>(tout[f] + so));
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < kVecsPerLane; ++r) {
                const int v = lane + 32 * r;
                const long long q = tile_lo + 4 * v;
                if (q + 3 < st_lo || q > st_hi) continue;
                const int so = 4 * swz(v);
#pragma unroll
                for (int f = 0; f < NOUT; ++f) {
                    const float4 yv = *reinterpret_cast<const float4*>(tout[f] + so);
                    float* dst = P.out[f] + rowoff;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (q + c >= st_lo && q + c <= st_hi) dst[q + c] = comp4(yv, c);
                }
            }
        }
        if (EPI != EPI_STORE && pk_on) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
            if (lane == 0 && pk > 0.f) atomicMax(reinterpret_cast<int*>(P.peak + row / P.channels), __float_as_int(pk));
        }

        if (ST > 1) slot ^= 1;
        else {
            if (bulk_out && lane == 0) tma_store_wait_read();
            __syncwarp();                              // single stage: every lane is done with the buffer
            if (tile + 1 < t_end) issue_inputs(row, tile + 1, 0);
            cp_async_commit();
        }
    }   // tiles
    cp_async_wait<0>();
    if (EPI == EPI_STORE && (PP.use_tma & 2) && lane == 0) tma_store_wait_all();    // the segment's bulk stores are complete
    }   // items
}

}  // namespace mm
