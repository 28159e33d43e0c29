// Stage functions and chains: host orchestration of the sweep / reduction / pointwise kernels.
// Every stage cites the reference function it reproduces (paths relative to the reference tree).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cuda.h>

#include "context.h"
#include "sweep4.cuh"
#include "lufs_kernel.cuh"
#include "misc_kernels.cuh"
#include "stages_internal.h"

namespace mm {

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
// warp-tiles (kWT samples) of a sweep
static inline int tiles_fwd(long long n, int pad) { return (int)((kLead + n + pad - 1 + kWT) / kWT); }
static inline int tiles_bwd(long long n, int pad) {
    const long long q_last = kLead + n + pad - 1, q_first = kLead - pad;
    const long long qend = bwd_qend(q_last);
    return (int)((qend - q_first + kWT - 1) / kWT);
}

template <int M> static void fill_filter(FiltK<M>& fk, const FilterPlan* p) {
    for (int i = 0; i <= M; ++i) fk.b[i] = p->ba.b[i];
    for (int i = 0; i < M; ++i) fk.a[i] = p->ba.a[i + 1];
    for (int j = 0; j < kS; ++j)
        for (int i = 0; i < M; ++i) fk.g[j][i] = p->tabs.g[(size_t)j * M + i];
    // float32 pass 2 in rescaled coordinates s' = d s, d_i = 1 / B_i (balanced plans guarantee B_i != 0; other plans never run it)
    double d[M];
    for (int i = 0; i < M; ++i) d[i] = (p->tabs.mode == kBalancedF32 && p->tabs.B[i] != 0.0) ? 1.0 / p->tabs.B[i] : 1.0;
    for (int i = 0; i < M; ++i) {
        for (int k = 0; k < M; ++k) fk.A32[i][k] = (float)(d[i] * p->tabs.A[i * M + k] / d[k]);
        fk.dn32[i] = (float)d[i];
        fk.C32[i] = (float)(p->tabs.C[i] / d[i]);
    }
    fk.D32 = (float)p->tabs.D;
}

// Segments per row: minimise  waves * (tiles per segment + halo)  over the segment count.
static void choose_segments(int rows, int ntiles, int whalo, int capacity, int* nseg_out, int* seglen_out) {
    long long best_cost = -1;
    int best = 1;
    const int max_seg = std::min(ntiles, std::max(1, (capacity * 4 + rows - 1) / rows));
    for (int ns = 1; ns <= max_seg; ++ns) {
        const int len = (ntiles + ns - 1) / ns;
        const int real_ns = (ntiles + len - 1) / len;
        if (real_ns != ns) continue;
        const long long items = (long long)rows * ns;
        const long long waves = (items + capacity - 1) / capacity;
        const long long cost = waves * (long long)(len + (ns > 1 ? whalo : 0)) * 64 + ns;   // mild preference for fewer segments
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ns; }
    }
    *nseg_out = best;
    *seglen_out = (ntiles + best - 1) / best;
}

// ---- TMA tensor maps ---------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
static int tma_policy() {            // MM_TMA: 0 = per-lane cp.async loads and float4 stores, 1 = bulk tensor loads, 2 (default) = + bulk stores
    static int pol = -1;
    if (pol < 0) { const char* e = getenv("MM_TMA"); pol = e ? std::max(0, std::min(2, atoi(e))) : 2; }
    return pol;
}
// A stream as a 2-D tensor of 128-byte lines: {32 floats, lines}; box = one warp-tile {32, 32}; SWIZZLE_128B.  `base` points at
// float 0 of row 0 plus `shift` floats (backward sweeps: their tile origins are = qend mod 32), `lines` 128-byte lines follow.
static bool make_stream_map(TmaDesc* out, const float* base, long long shift, long long lines) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || lines < 32 || (((uintptr_t)(base + shift)) & 15)) return false;
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap is 128 bytes");
    const cuuint64_t dims[2] = {32, (cuuint64_t)lines};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    return enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(base + shift), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int M, int NF, int NIN, int DIR, int EPI, int NAUX, int NF32>
static int launch_sweep2(mm_ctx* c, SweepArgs<M, NF>* As, int whalo, const char* name, int batch_rows) {
    constexpr int ST = SweepStages<NF, NIN>::value;
    typedef Sweep2Cfg<M, NF, NIN, DIR, EPI, NAUX, ST> Cfg;
    auto kern = sweep2_kernel<M, NF, NIN, DIR, EPI, NAUX, ST, NF32, 1>;
    SweepArgs<M, NF>& A = As[0];
    const size_t smem = Cfg::kBytes;
    int blocks_per_sm = 0;
    {
        const bool first = c->occupancy.find((const void*)kern) == c->occupancy.end();
        MM_TRY(kernel_setup(c, (const void*)kern, Cfg::kThreads, smem, true, &blocks_per_sm));
        if (first && getenv("MM_DEBUG")) fprintf(stderr, "[mm] %s (%d float32 sections, %d-stage ring): %zu B smem, %d CTAs x %d warps per SM x %d SMs\n", name, NF32, ST, smem, blocks_per_sm, Cfg::kSW, c->num_sms);
    }
    A.ntiles = DIR > 0 ? tiles_fwd(A.n, A.pad) : tiles_bwd(A.n, A.pad);
    if ((size_t)A.rows * (size_t)A.ntiles == 0) return 0;
    constexpr int kSW = Cfg::kSW, kSweepThreads = Cfg::kThreads;
    // warps in flight = independent workers (a lane context sizes its grid for its share of the device)
    const int capacity = std::max(kSW, c->num_sms * blocks_per_sm / std::max(1, c->grid_div) * kSW);
    Sweep2Args<M, NF, 1> PP;
    PP.whalo = whalo;                                // warp-tiles (ScanTables::Wq)
    choose_segments(A.rows, A.ntiles, PP.whalo, capacity, &PP.nseg, &PP.seglen);
    const long long items = (long long)A.rows * PP.nseg;
    const unsigned grid = (unsigned)std::min<long long>((items + kSW - 1) / kSW, capacity / kSW);
    PP.a[0] = A;
    // interior input tiles by TMA: one tensor map per input stream over the whole batch buffer
    PP.use_tma = 0;
    PP.tma_shift = 0;
    PP.tma_row_lines = A.stride / 32;
    if (tma_policy() && (A.stride % 32) == 0 && batch_rows > 0) {
        const long long q_last = kLead + A.n + A.pad - 1;
        const long long qend = bwd_qend(q_last);
        const long long shift = DIR > 0 ? 0 : (qend % 32);
        const long long lines = ((long long)batch_rows * A.stride - shift) / 32;
        bool ok = true;
        for (int s = 0; s < NIN && ok; ++s) ok = make_stream_map(&PP.tmap[s], A.in[s], shift, lines);
        if (ok) { PP.use_tma = 1; PP.tma_shift = (int)shift; }
        // bulk stores pay for the four-output forward sweep (4.6 -> 3.9 ms per launch: 64 LDS + STG per lane and tile become four
        // instructions of one lane); with one or two outputs the issuing lane's wait before the next tile costs more than the
        // stores it saves (measured +2..3 %), so those keep their coalesced float4 stores
        static const int store_min_nf = [] { const char* e = getenv("MM_TMA_STORE_MIN_NF"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 4; }();
        if (ok && EPI == EPI_STORE && NF >= store_min_nf && tma_policy() >= 2) {
            // (the halo start position is the same for loads and stores: both are tile origins)
            bool oko = true;
            for (int f = 0; f < NF && oko; ++f) oko = make_stream_map(&PP.tmap_out[f], A.out[f], shift, lines);
            if (oko) PP.use_tma |= 2;
        }
    }
    {
        KernelScope ks(c, name);
        ks.samples = (double)A.rows * (double)A.n;
        kern<<<grid, kSweepThreads, smem, c->stream>>>(PP);
    }
    MM_CUDA(cudaGetLastError());
    return 0;
}

static int halo_tiles(const FilterPlan* const* plans, int nf) {
    int w = 1;
    // 1e-18: results do not depend on the segmentation.  Counted in warp-tiles (round 1 and early round 2: in 4096-sample tiles,
    // i.e. rounded up to a multiple of four warp-tiles -- most sections forget within one or two)
    for (int f = 0; f < nf; ++f) w = std::max(w, plans[f]->tabs.Wq);
    return w;
}

template <int M, int NF>
static void fill_common(const mm_ctx* c, SweepArgs<M, NF>& A, const mm_geom* g, const FilterPlan* const* plans, const float* const* in, int nin,
                        float* const* out, int nout, const Pro& pro, const Epi& epi, int pad) {
    memset(&A, 0, sizeof(A));
    A.row_map = c->row_map;
    A.pk_lo = kLead + (c->slice ? c->slice->own_lo : 0);
    A.pk_hi = kLead + (c->slice ? c->slice->own_hi : g->n) - 1;
    for (int f = 0; f < NF; ++f) {
        fill_filter<M>(A.f[f], plans[f]);
        A.tab[f] = plans[f]->dev;
        A.W[f] = plans[f]->tabs.W;
        A.in[f] = in[f < nin ? f : nin - 1];
        A.out[f] = out[f < nout ? f : nout - 1];
        A.w[f] = epi.w[f];
        A.w32[f] = (float)epi.w[f];
    }
    if (M == 2) {
        for (int p = 0; 2 * p + 1 < NF; ++p) {
            const ScanTables &ta = plans[2 * p]->tabs, &tb = plans[2 * p + 1]->tabs;
            PairK& k = A.pr[p];
            for (int i = 0; i < 2; ++i) {
                for (int j = 0; j < 2; ++j) k.A[i][j] = make_float2((float)ta.A[i * 2 + j], (float)tb.A[i * 2 + j]);
                k.B[i] = make_float2((float)ta.B[i], (float)tb.B[i]);
                k.C[i] = make_float2((float)ta.C[i], (float)tb.C[i]);
                for (int j = 0; j < kS; ++j) k.g[j][i] = make_float2((float)ta.g[(size_t)j * 2 + i], (float)tb.g[(size_t)j * 2 + i]);
            }
            k.D = make_float2((float)ta.D, (float)tb.D);
        }
    }
    A.aux[0] = epi.aux0;
    A.aux[1] = epi.aux1;
    A.n = g->n;
    A.stride = g->stride;
    A.rows = c->row_map ? c->row_map_rows : g->tracks * g->channels;
    A.pad = pad;
    A.channels = g->channels;
    A.pro_mode = pro.mode;
    A.pro_sub = pro.sub;
    A.pro_mul = pro.mul;
    A.aux_pro = epi.aux_pro;
    A.epi = epi.mode;
    A.epi_clip = epi.clip;
    A.wc = epi.wc;
    A.trim = epi.trim;
    A.wc32 = (float)epi.wc;
    A.trim32 = (float)epi.trim;
    if (epi.dyn) A.dyn = *epi.dyn;
    A.exc_gain = epi.exc_gain;
    A.exc_k = epi.exc_k;
    A.exc_mode = epi.exc_mode;
    A.peak = epi.peak;
    A.w_row = epi.w_row;
    A.exc_row = epi.exc_row;
    A.peak_row = epi.peak_row;
}

// One sweep's sections put in launch order: the float32-pass-2 sections first (the kernel's NF32 counts a
// prefix).  `allowed` is the bit set of prefix lengths the kernel family is instantiated for; surplus float32
// sections fall back to their float64 plan.  Streams and weights follow their section unless the epilogue
// binds positions (dynamics: y0 = band 2, y1 = band 3), in which case it is all or nothing.
struct Arranged {
    const FilterPlan* plans[4];
    const float* in[4];
    float* out[4];
    Epi epi;
    int n32 = 0;
};
static int arrange(mm_ctx* c, int nf, const FilterPlan* const* plans, const float* const* in, int nin, float* const* out, int nout,
                   const Epi& epi, unsigned allowed, bool positional, Arranged* R) {
    const FilterPlan* pl[4];
    int n32 = 0;
    for (int f = 0; f < nf; ++f) { pl[f] = plans[f]; n32 += plans[f]->tabs.mode == kBalancedF32; }
    int want = n32;
    while (want > 0 && !((allowed >> want) & 1u)) --want;
    if (positional && want != nf) want = 0;
    // demote the float32 sections with the longest-lived state errors first
    while (n32 > want) {
        int worst = -1;
        for (int f = 0; f < nf; ++f)
            if (pl[f]->tabs.mode == kBalancedF32 && (worst < 0 || pl[f]->tabs.norm2 >= pl[worst]->tabs.norm2)) worst = f;
        pl[worst] = get_plan_mode(c, pl[worst]->ba, kDf2tF64);
        if (!pl[worst]) return 1;
        --n32;
    }
    int order[4], k = 0;
    for (int f = 0; f < nf; ++f) if (pl[f]->tabs.mode == kBalancedF32) order[k++] = f;
    for (int f = 0; f < nf; ++f) if (pl[f]->tabs.mode != kBalancedF32) order[k++] = f;
    R->epi = epi;
    for (int j = 0; j < nf; ++j) {
        const int f = order[j];
        R->plans[j] = pl[f];
        R->in[j] = in[nin == nf ? f : 0];
        R->out[j] = out[nout == nf ? f : 0];
        R->epi.w[j] = epi.w[f];
    }
    R->n32 = n32;
    return 0;
}

template <int M, int NF, int NIN, int DIR, int EPI, int NAUX, int NF32>
static int run_sweep(mm_ctx* c, const mm_geom* g, const Arranged& R, int nout, const Pro& pro, int pad, const char* name) {
    SweepArgs<M, NF> A;
    fill_common<M, NF>(c, A, g, R.plans, R.in, NIN, R.out, nout, pro, R.epi, pad);
    return launch_sweep2<M, NF, NIN, DIR, EPI, NAUX, NF32>(c, &A, halo_tiles(R.plans, NF), name, g->tracks * g->channels);
}

// `_safe_filtfilt` (pipeline.py:36-52): scipy's filtfilt refuses an input of <= padlen samples (ValueError) and the reference then
// returns the CAUSAL lfilter(b, a, x) from a zero state.  Here a zero-phase pair is a forward sweep followed by a backward sweep
// with the same pad; for such an input the forward sweep runs without extension and start state (pad 0: plain lfilter) and the
// backward sweep runs IDENTITY sections (y = x) under the same epilogue -- no extra kernels.
static bool degrades_to_lfilter(const mm_geom* g, int pad) { return pad > 0 && g->n <= pad; }
static const FilterPlan* identity_plan(mm_ctx* c, int m) {
    Ba ba;
    memset(&ba, 0, sizeof(ba));
    ba.m = m;
    ba.b[0] = 1.0;
    ba.a[0] = 1.0;
    return get_plan_mode(c, ba, kDf2tF64);
}

int sweep_fwd(mm_ctx* c, const mm_geom* g, int nf, int nin, const FilterPlan* const* plans, const float* const* in,
              float* const* out, const Pro& pro, int pad) {
    if (degrades_to_lfilter(g, pad)) pad = 0;
    Epi epi;
    const int m = plans[0]->ba.m;
    for (int f = 0; f < nf; ++f)
        if (plans[f]->ba.m != m) { set_error("mixed section orders in one sweep"); return 1; }
    if (nf > 4 || (nin != 1 && nin != nf)) { set_error("forward sweep: %d filters / %d inputs unsupported", nf, nin); return 1; }
    unsigned allowed = 1u;
    if (m == 2 && nf == 1) allowed = 0x3;
    if (m == 2 && nf == 2) allowed = 0x7;
    if (m == 2 && nf == 4) allowed = 0x15;
    Arranged R;
    MM_TRY(arrange(c, nf, plans, in, nin, out, nf, epi, allowed, false, &R));
#define MM_FWD(M_, NF_, NIN_, N32_) \
    if (m == M_ && nf == NF_ && nin == NIN_ && R.n32 == N32_) \
        return run_sweep<M_, NF_, NIN_, +1, EPI_STORE, 0, N32_>(c, g, R, nf, pro, pad, "sweep_fwd_m" #M_ "_f" #NF_ "_i" #NIN_);
    MM_FWD(2, 1, 1, 0) MM_FWD(2, 1, 1, 1)
    MM_FWD(2, 2, 1, 0) MM_FWD(2, 2, 1, 1) MM_FWD(2, 2, 1, 2)
    MM_FWD(2, 2, 2, 0) MM_FWD(2, 2, 2, 1) MM_FWD(2, 2, 2, 2)
    MM_FWD(2, 4, 1, 0) MM_FWD(2, 4, 1, 2) MM_FWD(2, 4, 1, 4)
    MM_FWD(4, 1, 1, 0)
#undef MM_FWD
    set_error("no forward sweep instantiation for order %d, %d filters, %d inputs", m, nf, nin);
    return 1;
}

int sweep_bwd(mm_ctx* c, const mm_geom* g, int nf, const FilterPlan* const* plans, const float* const* in,
              float* const* out, int nout, const Epi& epi, int pad) {
    const FilterPlan* ident[4];
    if (degrades_to_lfilter(g, pad) && nf <= 4) {
        const FilterPlan* id = identity_plan(c, plans[0]->ba.m);
        if (!id) return 1;
        for (int f = 0; f < nf; ++f) ident[f] = id;
        plans = ident;
        pad = 0;
    }
    const int m = plans[0]->ba.m;
    const int naux = (epi.mode == EPI_STORE) ? 0 : ((epi.aux1 != nullptr) ? 2 : 1);
    if (epi.mode != EPI_STORE && epi.aux0 == nullptr) { set_error("backward sweep epilogue needs its x-domain stream"); return 1; }
    if (nf > 4) { set_error("backward sweep: %d filters unsupported", nf); return 1; }
    if (m == 2 && nf == 4 && epi.mode == EPI_STORE && nout == 4) {
        // four independent sections: two 2-section sweeps move the same bytes and fit more CTAs per SM
        MM_TRY(sweep_bwd(c, g, 2, plans, in, out, 2, epi, pad));
        return sweep_bwd(c, g, 2, plans + 2, in + 2, out + 2, 2, epi, pad);
    }
    const bool positional = epi.mode == EPI_DYNAMICS || epi.mode == EPI_DYNAMICS_GEN;
    unsigned allowed = 1u;
    if (m == 2 && nf == 1) allowed = 0x3;
    if (m == 2 && nf == 2) allowed = (epi.mode == EPI_STORE) ? 0x7 : 0x5;
    if (m == 2 && nf == 4) allowed = 0x11;
    Arranged R;
    MM_TRY(arrange(c, nf, plans, in, nf, out, nout, epi, allowed, positional, &R));
#define MM_BWD(M_, NF_, EPI_, NAUX_, N32_, TAG_) \
    if (m == M_ && nf == NF_ && epi.mode == EPI_ && naux == NAUX_ && R.n32 == N32_) \
        return run_sweep<M_, NF_, NF_, -1, EPI_, NAUX_, N32_>(c, g, R, nout, epi.auxp, pad, "sweep_bwd_m" #M_ "_f" #NF_ TAG_);
    MM_BWD(2, 1, EPI_STORE, 0, 0, "_store") MM_BWD(2, 1, EPI_STORE, 0, 1, "_store")
    MM_BWD(2, 1, EPI_COMBINE, 1, 0, "_combine") MM_BWD(2, 1, EPI_COMBINE, 1, 1, "_combine")
    MM_BWD(2, 1, EPI_EXCITER, 1, 0, "_exciter") MM_BWD(2, 1, EPI_EXCITER, 1, 1, "_exciter")
    MM_BWD(2, 2, EPI_STORE, 0, 0, "_store") MM_BWD(2, 2, EPI_STORE, 0, 1, "_store") MM_BWD(2, 2, EPI_STORE, 0, 2, "_store")
    MM_BWD(2, 2, EPI_COMBINE, 1, 0, "_combine") MM_BWD(2, 2, EPI_COMBINE, 1, 2, "_combine")
    MM_BWD(2, 2, EPI_DYNAMICS, 2, 0, "_dynamics") MM_BWD(2, 2, EPI_DYNAMICS, 2, 2, "_dynamics")
    MM_BWD(2, 2, EPI_DYNAMICS_GEN, 2, 0, "_dynamics_gen") MM_BWD(2, 2, EPI_DYNAMICS_GEN, 2, 2, "_dynamics_gen")
    MM_BWD(2, 4, EPI_COMBINE, 1, 0, "_combine") MM_BWD(2, 4, EPI_COMBINE, 1, 4, "_combine")
    MM_BWD(4, 1, EPI_STORE, 0, 0, "_store")
#undef MM_BWD
    set_error("no backward sweep instantiation for order %d, %d filters, epilogue %d, %d aux", m, nf, epi.mode, naux);
    return 1;
}

// ---------------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------------
const FilterPlan* plan_butter(mm_ctx* c, int order, BType bt, double w0, double w1, int prec) {
    Ba ba;
    memset(&ba, 0, sizeof(ba));
    double wn[2] = {w0, w1};
    if (!butter(order, bt, wn, &ba)) { set_error("butter(%d, [%g, %g], type %d): bad critical frequencies", order, w0, w1, (int)bt); return nullptr; }
    return get_plan(c, ba, prec);
}

int get_bufs(mm_ctx* c, const mm_geom* g, Bufs* B) {
    const size_t fl = (size_t)g->tracks * g->channels * (size_t)g->stride;
    for (int i = 0; i < 4; ++i) MM_TRY(arena(c, SL_E0 + i, fl, &B->E[i]));
    for (int i = 0; i < 5; ++i) MM_TRY(arena(c, SL_T0 + i, fl, &B->T[i]));
    return 0;
}

int check_geom(const mm_geom* g) {
    if (!g || g->n <= 0 || g->tracks <= 0 || (g->channels != 1 && g->channels != 2) || g->sr <= 0) {
        set_error("bad geometry (n > 0, tracks > 0, channels in {1,2}, sr > 0 required)");
        return 1;
    }
    if (g->stride < mm_row_stride(g->n) || (g->stride & 3)) { set_error("stride must be >= mm_row_stride(n) and a multiple of 4"); return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// reductions / scalars
// ---------------------------------------------------------------------------------------------------
int run_row_stats(mm_ctx* c, const mm_geom* g, const float* in, RowStats** st_out) {
    const int rows = g->tracks * g->channels;
    RowStats* st;
    MM_TRY(arena(c, SL_ROWSTATS, (size_t)rows, &st));
    {
        KernelScope ks(c, "row_stats_init");
        row_stats_init_kernel<<<(rows + 255) / 256, 256, 0, c->stream>>>(st, rows);
    }
    // a time slice contributes its own frames only (own_lo is a multiple of 4: the float4 loads stay aligned)
    const long long lo = c->slice ? c->slice->own_lo : 0, cnt = c->slice ? c->slice->own_hi - c->slice->own_lo : g->n;
    dim3 grid((unsigned)((cnt + kRsFramesPerBlock - 1) / kRsFramesPerBlock), (unsigned)rows);
    {
        KernelScope ks(c, "row_stats");
        row_stats_kernel<<<grid, kPwThreads, 0, c->stream>>>(in + lo, cnt, g->stride, st);
    }
    MM_CUDA(cudaGetLastError());
    *st_out = st;
    return 0;
}

// One all-reduce of a time slice's exchange step, in stream order: NCCL enqueued from C on the context's stream when the slice
// carries a communicator (nccl_shim.cu), else the caller's callback (tests: threads sharing a GPU, gloo on the CPU).
int slice_allreduce(mm_ctx* c, void* ptr, int64_t count, int dtype, int op, const char* what) {
    const mm_slice* sl = c->slice;
    if (!sl) return 0;
    if (sl->nccl_comm) {
        c->launches += 1;            // NCCL's kernel, not ours -- counted so that the launch total covers everything on the stream
        return nccl_allreduce(c, sl->nccl_comm, ptr, count, dtype, op);
    }
    if (!sl->allreduce) return 0;
    if (sl->allreduce(sl->user, ptr, count, dtype, op) != 0) { set_error("allreduce of %s failed", what); return 1; }
    return 0;
}
static inline bool slice_exchanges(const mm_ctx* c) { return c->slice && (c->slice->nccl_comm || c->slice->allreduce); }

int exchange_row_stats(mm_ctx* c, RowStats* st, int rows) {
    if (!slice_exchanges(c)) return 0;
    double* xb;
    MM_TRY(arena(c, SL_XCHG, (size_t)rows * 3, &xb));
    row_stats_pack_kernel<<<(rows + 127) / 128, 128, 0, c->stream>>>(st, rows, xb);
    MM_CUDA(cudaGetLastError());
    MM_TRY(slice_allreduce(c, xb, rows, 0, 0, "the channel sums"));
    MM_TRY(slice_allreduce(c, xb + rows, 2 * (int64_t)rows, 0, 2, "the channel minima / maxima"));
    row_stats_unpack_kernel<<<(rows + 127) / 128, 128, 0, c->stream>>>(st, rows, xb);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_in_scalars(mm_ctx* c, const mm_geom* g, const RowStats* st, int use_dc, int use_guard, double headroom_db,
                   double* sub, double* mul, double* peak_track, double* mean_row) {
    InScalarArgs A;
    A.st = st; A.n = c->slice ? c->slice->global_n : g->n; A.tracks = g->tracks; A.channels = g->channels;
    A.use_dc = use_dc; A.use_guard = use_guard;
    A.limit = (float)std::pow(10.0, -headroom_db / 20.0);
    A.sub = sub; A.mul = mul; A.peak_track = peak_track; A.mean_row = mean_row;
    KernelScope ks(c, "in_scalars");
    in_scalars_kernel<<<(g->tracks + 127) / 128, 128, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_pointwise(mm_ctx* c, const mm_geom* g, PwArgs& A, const char* name) {
    A.n = g->n; A.stride = g->stride; A.tracks = g->tracks; A.channels = g->channels; A.track_base = g->track_base;
    dim3 grid((unsigned)((g->n + kPwFramesPerBlock - 1) / kPwFramesPerBlock), (unsigned)g->tracks);
    KernelScope ks(c, name);
    ks.samples = (double)g->n * g->tracks * g->channels;
    pointwise_kernel<<<grid, kPwThreads, 0, c->stream>>>(A);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int reset_imager_peaks(mm_ctx* c, float* peak, const double* width, int tracks) {
    reset_imager_peaks_kernel<<<(tracks + 127) / 128, 128, 0, c->stream>>>(peak, width, tracks);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_out_scalars(mm_ctx* c, const OutScalarArgs& O) {
    KernelScope ks(c, "out_scalars");
    out_scalars_kernel<<<(O.tracks + 127) / 128, 128, 0, c->stream>>>(O);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_finalize(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* mul, const double* width, int n_fade,
                 int16_t* pcm, const float* noise, unsigned long long seed, double* nonfinite) {
    FinalArgs A;
    A.in = in; A.out = out; A.n = g->n; A.stride = g->stride; A.tracks = g->tracks; A.mul = mul; A.width = width;
    A.n_fade = n_fade; A.fade_step = n_fade > 1 ? 1.0 / (double)(n_fade - 1) : 0.0;
    A.pcm = pcm; A.noise = noise; A.seed = seed; A.nonfinite = nonfinite; A.track_base = g->track_base;
    A.frame_base = c->slice ? c->slice->global_off : 0;
    A.track_ids = c->track_ids_dev;
    dim3 grid((unsigned)((g->n + kFinFrames - 1) / kFinFrames), (unsigned)g->tracks);
    KernelScope ks(c, pcm ? "finalize_dither_int16" : "finalize");
#define MM_FIN(C_, PCM_, NZ_) finalize_kernel<C_, PCM_, NZ_><<<grid, kFinThreads, 0, c->stream>>>(A)
    if (g->channels == 2) {
        if (!pcm) MM_FIN(2, false, false); else if (noise) MM_FIN(2, true, true); else MM_FIN(2, true, false);
    } else {
        if (!pcm) MM_FIN(1, false, false); else if (noise) MM_FIN(1, true, true); else MM_FIN(1, true, false);
    }
#undef MM_FIN
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_quantize(mm_ctx* c, const QuantArgs& Q) {
    dim3 grid((unsigned)((Q.n + kPwThreads - 1) / kPwThreads), (unsigned)Q.tracks);
    KernelScope ks(c, "quantize_int16");
    quantize_kernel<<<grid, kPwThreads, 0, c->stream>>>(Q);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_layout_pcm16(mm_ctx* c, const mm_geom* g, const int16_t* interleaved, float* planar) {
    dim3 grid((unsigned)((g->n + kPwThreads - 1) / kPwThreads), (unsigned)g->tracks);
    KernelScope ks(c, "deinterleave_pcm16");
    deinterleave_pcm16_kernel<<<grid, kPwThreads, 0, c->stream>>>(interleaved, planar, g->n, g->stride, g->channels);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_signal_metrics(mm_ctx* c, const mm_geom* g, const float* in, double* out3) {
    MM_CUDA(cudaMemsetAsync(out3, 0, (size_t)g->tracks * 3 * sizeof(double), c->stream));
    const unsigned gx = (unsigned)std::min<long long>((g->n + kPwThreads - 1) / kPwThreads, 1024);
    dim3 grid(gx, (unsigned)(g->tracks * g->channels));
    KernelScope ks(c, "trace_signal_metrics");
    signal_metrics_kernel<<<grid, kPwThreads, 0, c->stream>>>(in, g->n, g->stride, g->channels, out3);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_white_noise(mm_ctx* c, const WhiteArgs& W) {
    dim3 grid((unsigned)((W.n + kPwThreads - 1) / kPwThreads), (unsigned)W.tracks);
    KernelScope ks(c, "dither_white_noise");
    white_noise_kernel<<<grid, kPwThreads, 0, c->stream>>>(W);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int run_layout(mm_ctx* c, const mm_geom* g, const float* interleaved, float* planar, int dir) {
    dim3 grid((unsigned)((g->n + kPwThreads - 1) / kPwThreads), (unsigned)g->tracks);
    KernelScope ks(c, dir ? "interleave" : "deinterleave");
    if (dir) interleave_kernel<<<grid, kPwThreads, 0, c->stream>>>(planar, const_cast<float*>(interleaved), g->n, g->stride, g->channels);
    else deinterleave_kernel<<<grid, kPwThreads, 0, c->stream>>>(interleaved, planar, g->n, g->stride, g->channels);
    MM_CUDA(cudaGetLastError());
    return 0;
}

void fill_dyn(DynParams* d, double knee_db, const double* band_ratios, double max_upward_boost_db) {
    static const double cfg[4][4] = {{-7.2, 1.0, -7.2, 1.5}, {-18.5, 2.2, -18.5, 1.8}, {-17.0, 1.55, -17.0, 1.65}, {-15.0, 1.35, -15.0, 1.2}};
    memset(d, 0, sizeof(*d));
    knee_db = std::max(0.0, knee_db);
    for (int i = 0; i < 4; ++i) {
        DynBand& b = d->band[i];
        const double lim_db = cfg[i][0], thr_db = cfg[i][2], gain = cfg[i][3];
        const double ratio = band_ratios ? band_ratios[i] : cfg[i][1];
        b.thr_db = thr_db;
        b.thr = std::pow(10.0, thr_db / 20.0);
        b.ratio = ratio;
        b.max_boost_db = std::max(0.1, max_upward_boost_db);
        b.lower = b.thr * std::pow(10.0, -knee_db / 20.0);
        b.upper = b.thr * std::pow(10.0, knee_db / 20.0);
        b.slope = (b.upper > b.lower) ? (b.thr + (b.upper - b.thr) / ratio - b.lower) / (b.upper - b.lower) : 1.0;
        b.lim = (float)std::pow(10.0, lim_db / 20.0);
        b.gain = (float)gain;
        // lines of the concave downward knee (band_chain): identity by default
        b.s_mid = 1.f; b.c_mid = 0.f; b.s_hi = 1.f; b.c_hi = 0.f;
        if (ratio <= 0.0 || ratio == 1.0) b.mode = 0;
        else if (ratio < 1.0) b.mode = 3;
        else {
            b.s_hi = (float)(1.0 / ratio);
            b.c_hi = (float)(b.thr * (1.0 - 1.0 / ratio));
            if (knee_db < 0.5) { b.mode = 1; b.s_mid = b.s_hi; b.c_mid = b.c_hi; }
            else { b.mode = 2; b.s_mid = (float)b.slope; b.c_mid = (float)(b.lower * (1.0 - b.slope)); }
        }
    }
    const double thr = std::pow(10.0, -2.5 / 20.0), ceil_ = std::pow(10.0, -0.3 / 20.0);
    const double k = (ceil_ - thr) / (1.0 - thr);
    d->max_k = (float)k;
    d->max_c = (float)(thr * (1.0 - k));
    d->max_top = (float)std::min(ceil_, std::pow(10.0, -1.5 / 20.0));   // min(ceil) then clip(+-TRUE_PEAK_LIMIT_DB)
    d->par_mix = nullptr;
    fill_parallel(d, 8.0, -20.0);
}

void fill_parallel(DynParams* d, double ratio, double threshold_db) {
    const double thr = std::pow(10.0, threshold_db / 20.0);
    const double lower = thr * std::pow(10.0, -6.0 / 20.0), upper = thr * std::pow(10.0, 6.0 / 20.0);
    const double slope = (thr + (upper - thr) / ratio - lower) / (upper - lower);
    d->par_slope = (float)slope;
    d->par_cmid = (float)(lower * (1.0 - slope));
    d->par_shi = (float)(1.0 / ratio);
    d->par_chi = (float)(thr * (1.0 - 1.0 / ratio));
}

// ---------------------------------------------------------------------------------------------------
// stages
// ---------------------------------------------------------------------------------------------------
// apply_target_curve, IIR path (backend/app/pipeline.py:238-273, designs :170-184)
int st_target_curve(mm_ctx* c, const mm_geom* g, const float* in, float* out, const Pro& pro) {
    const double nyq = g->sr / 2.0;
    const FilterPlan* hp = plan_butter(c, 2, kHigh, std::min(40.0 / nyq, 0.99), 0);
    const FilterPlan* lp = plan_butter(c, 2, kLow, std::min(18000.0 / nyq, 0.99), 0);
    const double fp = std::min(3000.0 / nyq, 0.99), fm = std::min(300.0 / nyq, 0.99);
    const FilterPlan* pres = plan_butter(c, 1, kBand, fp * 0.7, fp * 1.3, PREC_F32);   // enter through weights of 0.04 / 0.03
    const FilterPlan* mud = plan_butter(c, 1, kBand, fm * 0.7, fm * 1.3, PREC_F32);
    if (!hp || !lp || !pres || !mud) return 1;
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    const double gp = std::pow(10.0, 0.35 / 20.0), gm = std::pow(10.0, -0.25 / 20.0);
    Pro none;
    Epi store;
    {
        const FilterPlan* p[1] = {hp};
        const float* i1[1] = {in};
        float* o1[1] = {B.E[0]};
        MM_TRY(sweep_fwd(c, g, 1, 1, p, i1, o1, pro, 9));
        const float* i2[1] = {B.E[0]};
        float* o2[1] = {B.T[0]};
        MM_TRY(sweep_bwd(c, g, 1, p, i2, o2, 1, store, 9));
    }
    {
        const FilterPlan* p[1] = {lp};
        const float* i1[1] = {B.T[0]};
        float* o1[1] = {B.E[0]};
        MM_TRY(sweep_fwd(c, g, 1, 1, p, i1, o1, none, 9));
        const float* i2[1] = {B.E[0]};
        float* o2[1] = {B.T[0]};
        MM_TRY(sweep_bwd(c, g, 1, p, i2, o2, 1, store, 9));
    }
    {
        const FilterPlan* p[2] = {pres, mud};
        const float* i1[1] = {B.T[0]};
        float* o1[2] = {B.E[0], B.E[1]};
        MM_TRY(sweep_fwd(c, g, 2, 1, p, i1, o1, none, 9));
        const float* i2[2] = {B.E[0], B.E[1]};
        float* o2[1] = {out};
        Epi e;
        e.mode = EPI_COMBINE;
        e.aux0 = B.T[0];
        e.w[0] = gp - 1.0;
        e.w[1] = gm - 1.0;
        MM_TRY(sweep_bwd(c, g, 2, p, i2, o2, 1, e, 9));
    }
    return 0;
}

// apply_dynamics = apply_multiband_dynamics (numpy branch) + apply_maximizer + hard limiter
// (backend/app/pipeline.py:610-641, :414-481, :333-364)
int st_dynamics(mm_ctx* c, const mm_geom* g, const float* in, float* out, double knee_db, const double* crossovers_hz,
                const double* band_ratios, double max_upward_boost_db, const double* par_mix_rows, float* peak, int bands_only,
                int compressor) {
    double cross[3] = {214.0, 3500.0, 10000.0};
    if (crossovers_hz) {
        double t[3];
        for (int i = 0; i < 3; ++i) t[i] = std::min(std::max(crossovers_hz[i], 20.0), 20000.0);
        if (!(t[0] >= t[1] || t[1] >= t[2])) { cross[0] = t[0]; cross[1] = t[1]; cross[2] = t[2]; }
    }
    const double nyq = g->sr / 2.0;
    double f[3];
    for (int i = 0; i < 3; ++i) f[i] = std::min(cross[i] / nyq, 0.99);
    const FilterPlan* lp1 = plan_butter(c, 2, kLow, f[0], 0);
    const FilterPlan* hp1 = plan_butter(c, 2, kHigh, f[0], 0);
    const FilterPlan* lp2 = plan_butter(c, 2, kLow, f[1], 0);
    const FilterPlan* hp2 = plan_butter(c, 2, kHigh, f[1], 0);
    const FilterPlan* lp3 = plan_butter(c, 2, kLow, f[2], 0);
    const FilterPlan* hp3 = plan_butter(c, 2, kHigh, f[2], 0);
    if (!lp1 || !hp1 || !lp2 || !hp2 || !lp3 || !hp3) return 1;
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    Pro none;
    Epi store;
    {
        const FilterPlan* p[4] = {lp1, hp1, hp2, hp3};
        const float* i1[1] = {in};
        float* o1[4] = {B.E[0], B.E[1], B.E[2], B.E[3]};
        MM_TRY(sweep_fwd(c, g, 4, 1, p, i1, o1, none, 9));
        const float* i2[4] = {B.E[0], B.E[1], B.E[2], B.E[3]};
        float* o2[4] = {B.T[1], B.T[2], B.T[3], B.T[4]};
        MM_TRY(sweep_bwd(c, g, 4, p, i2, o2, 4, store, 9));
    }
    {
        const FilterPlan* p[2] = {lp2, lp3};
        const float* i1[2] = {B.T[2], B.T[3]};
        float* o1[2] = {B.E[0], B.E[1]};
        MM_TRY(sweep_fwd(c, g, 2, 2, p, i1, o1, none, 9));
        DynParams d;
        fill_dyn(&d, knee_db, band_ratios, max_upward_boost_db);
        d.par_mix = par_mix_rows;
        if (bands_only) {                            // min(|s|, |s| + inf, inf) = |s|: the maximizer / limiter step is the identity
            d.max_k = 1.0f; d.max_c = INFINITY; d.max_top = INFINITY;
        }
        if (compressor == MM_COMPRESSOR_ENVELOPE) {
            // envelope-compressor mode (pipeline.py:373-411): the two middle bands are materialised (2 R / 2 W), then one pass over
            // the four bands runs the followers, limiters, sum, maximizer and limiter (4 R / 1 W) -- bandcomp.cu
            if (peak) { set_error("apply_dynamics (envelope compressor): output-peak tracking is not fused into this mode"); return 1; }
            const float* i2[2] = {B.E[0], B.E[1]};
            float* o2[2] = {B.T[2], B.T[3]};
            MM_TRY(sweep_bwd(c, g, 2, p, i2, o2, 2, store, 9));
            const float* bands[4] = {B.T[1], B.T[2], B.T[3], B.T[4]};
            return launch_band_compress(c, g, bands, out, d);
        }
        bool general = par_mix_rows != nullptr;
        for (int i = 0; i < 4; ++i) general |= d.band[i].mode == 3;
        Epi e;
        e.mode = general ? EPI_DYNAMICS_GEN : EPI_DYNAMICS;
        e.aux0 = B.T[1];
        e.aux1 = B.T[4];
        e.dyn = &d;
        e.peak = peak;
        const float* i2[2] = {B.E[0], B.E[1]};
        float* o2[1] = {out};
        MM_TRY(sweep_bwd(c, g, 2, p, i2, o2, 1, e, 9));
    }
    return 0;
}

int st_split_bands(mm_ctx* c, const mm_geom* g, const float* in, const double* cross, float** bands) {
    const double nyq = g->sr / 2.0;
    double f[3];
    for (int i = 0; i < 3; ++i) f[i] = std::min(cross[i] / nyq, 0.99);
    const FilterPlan* lp1 = plan_butter(c, 2, kLow, f[0], 0);
    const FilterPlan* hp1 = plan_butter(c, 2, kHigh, f[0], 0);
    const FilterPlan* lp2 = plan_butter(c, 2, kLow, f[1], 0);
    const FilterPlan* hp2 = plan_butter(c, 2, kHigh, f[1], 0);
    const FilterPlan* lp3 = plan_butter(c, 2, kLow, f[2], 0);
    const FilterPlan* hp3 = plan_butter(c, 2, kHigh, f[2], 0);
    if (!lp1 || !hp1 || !lp2 || !hp2 || !lp3 || !hp3) return 1;
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    Pro none;
    Epi store;
    {
        const FilterPlan* p[4] = {lp1, hp1, hp2, hp3};
        const float* i1[1] = {in};
        float* o1[4] = {B.E[0], B.E[1], B.E[2], B.E[3]};
        MM_TRY(sweep_fwd(c, g, 4, 1, p, i1, o1, none, 9));
        const float* i2[4] = {B.E[0], B.E[1], B.E[2], B.E[3]};
        float* o2[4] = {B.T[1], B.T[2], B.T[3], B.T[4]};
        MM_TRY(sweep_bwd(c, g, 4, p, i2, o2, 4, store, 9));
    }
    {
        const FilterPlan* p[2] = {lp2, lp3};
        const float* i1[2] = {B.T[2], B.T[3]};
        float* o1[2] = {B.E[0], B.E[1]};
        MM_TRY(sweep_fwd(c, g, 2, 2, p, i1, o1, none, 9));
        const float* i2[2] = {B.E[0], B.E[1]};
        float* o2[2] = {B.T[2], B.T[3]};
        MM_TRY(sweep_bwd(c, g, 2, p, i2, o2, 2, store, 9));
    }
    bands[0] = B.T[1]; bands[1] = B.T[2]; bands[2] = B.T[3]; bands[3] = B.T[4];
    return 0;
}

// pyloudnorm.Meter(sr).integrated_loudness (+ the gain normalize_lufs derives from it,
// backend/app/pipeline.py:644-664)
int st_lufs(mm_ctx* c, const mm_geom* g, const float* in, const Pro& pro, double* lufs_dev, const double* target_dev,
            double* gain_row, double* gain_db) {
    if (pro.mode != PRO_NONE) { set_error("the loudness kernel takes its input as it is (no fused prologue)"); return 1; }
    const LufsPlan* lp;
    const mm_slice* sl = c->slice;
    if (sl) MM_TRY(get_lufs_plan(c, sl->global_n, g->sr, &lp, sl->global_off, g->n));
    else MM_TRY(get_lufs_plan(c, g->n, g->sr, &lp));
    const int rows = g->tracks * g->channels;
    unsigned long long* segsum;
    MM_TRY(arena(c, SL_SEGSUM, (size_t)rows * (size_t)std::max(lp->nseg, 1), &segsum));
    if (lp->valid) {
        const KwPlan* kw = get_kw_plan(c, g->sr);
        if (!kw) return 1;
        MM_CUDA(cudaMemsetAsync(segsum, 0, (size_t)rows * lp->nseg * sizeof(unsigned long long), c->stream));
        const int S = lp->min_span2 >= 2048 ? 64 : 32;            // samples per lane and warp scan (64 unless the hops are too short)
        const ScanTables& tb = S == 64 ? kw->tabs64 : kw->tabs;
        const int wt = 32 * S;                                      // samples per warp-tile
        auto kern = S == 64 ? lufs_kernel<64> : lufs_kernel<32>;
        const int smem = S == 64 ? LufsCfg<64>::kSmem : LufsCfg<32>::kSmem;
        int bps = 0;
        MM_TRY(kernel_setup(c, (const void*)kern, kT, smem, true, &bps));
        const int capacity = std::max(1, std::max(1, bps) * c->num_sms / std::max(1, c->grid_div));
        LufsArgs A;
        memset(&A, 0, sizeof(A));
        for (int j = 0; j < S; ++j) {
            A.g[j][0] = make_float2((float)tb.g[(size_t)j * 4 + 0], (float)tb.g[(size_t)j * 4 + 1]);
            A.g[j][1] = make_float2((float)tb.g[(size_t)j * 4 + 2], (float)tb.g[(size_t)j * 4 + 3]);
        }
        for (int k = 0; k < 4; ++k)
            for (int h = 0; h < 2; ++h) {
                for (int d = 0; d < 5; ++d)
                    A.Pw2[d][k][h] = make_float2((float)tb.Pw[(size_t)d * 16 + (2 * h) * 4 + k], (float)tb.Pw[(size_t)d * 16 + (2 * h + 1) * 4 + k]);
                A.Q1[k][h] = make_float2((float)tb.Qpow[(size_t)16 + (2 * h) * 4 + k], (float)tb.Qpow[(size_t)16 + (2 * h + 1) * 4 + k]);
            }
        {
            // shelf in DF2T coordinates: z = T s with T = O_d^-1 O_b (observability matrices of the DF2T and the balanced realization)
            const Ba sh = k_weighting_stage(0, (double)g->sr);
            const double* Ab = kw->sec[0].A;
            const double* Cb = kw->sec[0].C;
            const double a1 = sh.a[1], a2 = sh.a[2];
            const double ob[2][2] = {{Cb[0], Cb[1]}, {Cb[0] * Ab[0] + Cb[1] * Ab[2], Cb[0] * Ab[1] + Cb[1] * Ab[3]}};     // [C; C A]
            for (int j = 0; j < 2; ++j) {
                A.shT[0][j] = (float)ob[0][j];                       // O_d^-1 = [[1, 0], [a1, 1]]
                A.shT[1][j] = (float)(a1 * ob[0][j] + ob[1][j]);
            }
            for (int i = 0; i < 3; ++i) A.sh_b[i] = (float)sh.b[i];
            A.sh_na[0] = (float)-a1;
            A.sh_na[1] = (float)-a2;
            // high-pass: Chamberlin state-variable form in its own coordinates (lp, bp) -- the scan's states 2, 3 are used as they are
            A.hp_f = (float)kw->hp_f;
            A.hp_nq = (float)-kw->hp_q;
            A.hp_g2 = kw->hp_g * kw->hp_g;
        }
        A.plane = S == 64 ? kw->plane64 : kw->dev + Tab<4>::Plane;
        A.in = in; A.n = g->n; A.stride = g->stride; A.rows = rows; A.channels = g->channels;
        A.ntiles = (int)((kLead + g->n + wt - 1) / wt);             // warp-tiles (the plan counts 4096-sample tiles)
        A.pro_mode = pro.mode; A.pro_sub = pro.sub; A.pro_mul = pro.mul;
        A.bnd = lp->bnd; A.nhop = lp->nseg; A.tile_seg = lp->tile_seg; A.segsum = segsum;
        A.goff = sl ? sl->global_off : 0;
        A.own_lo = sl ? sl->own_lo : 0;
        A.own_hi = sl ? sl->own_hi : g->n;
        A.whalo = tb.Wq;                                           // warp-tiles of 32 S samples
        // every warp is an independent worker: capacity and segments are counted in warps / warp-tiles
        choose_segments(rows, A.ntiles, A.whalo, capacity * kNW, &A.nseg, &A.seglen);
        const long long items = (long long)rows * A.nseg;
        const unsigned grid = (unsigned)std::min<long long>((items + kNW - 1) / kNW, capacity);
        {
            KernelScope ks(c, "lufs_kweight_blocks");
            ks.samples = (double)rows * (double)g->n;
            kern<<<grid, kT, smem, c->stream>>>(A);
        }
        MM_CUDA(cudaGetLastError());
    }
    if (slice_exchanges(c) && lp->valid)         // block sums of the other ranks' frames (exact: 64-bit fixed point)
        MM_TRY(slice_allreduce(c, segsum, (int64_t)rows * lp->nseg, 1, 0, "the loudness block sums"));
    GateArgs G;
    memset(&G, 0, sizeof(G));
    G.segsum = segsum; G.nseg = std::max(lp->nseg, 1); G.nblocks = lp->nblocks; G.channels = g->channels; G.tracks = g->tracks;
    G.blk_lo = lp->blk_lo; G.blk_hi = lp->blk_hi; G.scale = lp->scale; G.valid = lp->valid;
    G.lufs = lufs_dev; G.target = target_dev; G.gain_row = gain_row; G.gain_db = gain_db;
    {
        KernelScope ks(c, "lufs_gate");
        if (lp->nblocks >= kGateLongMinBlocks) gate_long_kernel<<<g->tracks * kGateCluster, kGateLongThreads, 0, c->stream>>>(G);
        else gate_kernel<<<g->tracks, 256, 0, c->stream>>>(G);
    }
    MM_CUDA(cudaGetLastError());
    return 0;
}

// apply_final_spectral_balance (backend/app/pipeline.py:576-607); `pro` scales the input first
// (normalize_lufs's gain folds in here at zero traffic)
int st_final_balance(mm_ctx* c, const mm_geom* g, const float* in, float* out, const Pro& pro, float* peak) {
    const double nyq = g->sr / 2.0;
    const double f3 = std::min(3000.0 / nyq, 0.99), f8 = std::min(8000.0 / nyq, 0.99);
    const FilterPlan* p3 = plan_butter(c, 1, kBand, f3 * 0.8, f3 * 1.2, PREC_F32);      // all four enter through weights <= 0.015
    const FilterPlan* p16 = plan_butter(c, 2, kHigh, std::min(16000.0 / nyq, 0.99), 0, PREC_F32);
    const FilterPlan* plo = plan_butter(c, 2, kLow, std::min(180.0 / nyq, 0.99), 0, PREC_F32);
    const FilterPlan* p8 = plan_butter(c, 1, kBand, f8 * 0.8, f8 * 1.2, PREC_F32);
    if (!p3 || !p16 || !plo || !p8) return 1;
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    const FilterPlan* p[4] = {p3, p16, plo, p8};
    const float* i1[1] = {in};
    float* o1[4] = {B.E[0], B.E[1], B.E[2], B.E[3]};
    MM_TRY(sweep_fwd(c, g, 4, 1, p, i1, o1, pro, 9));
    Epi e;
    e.mode = EPI_COMBINE;
    e.aux0 = in;
    e.auxp = pro;
    e.aux_pro = pro.mode != PRO_NONE;
    e.w[0] = (std::pow(10.0, -0.5 / 20.0) - 1.0) * 0.25;
    e.w[1] = (std::pow(10.0, -0.3 / 20.0) - 1.0) * 0.25;
    e.w[2] = (std::pow(10.0, 0.3 / 20.0) - 1.0) * 0.25;
    e.w[3] = (std::pow(10.0, 0.2 / 20.0) - 1.0) * 0.25;
    e.trim = std::pow(10.0, 0.5 / 20.0);
    e.peak = peak;
    const float* i2[4] = {B.E[0], B.E[1], B.E[2], B.E[3]};
    float* o2[1] = {out};
    MM_TRY(sweep_bwd(c, g, 4, p, i2, o2, 1, e, 9));
    return 0;
}

// one zero-phase section with "x + w * filtered" recombination: a style-EQ band
// (pipeline.py:1427-1431), the exciter's high-pass (:1303-1315), or plain filtfilt (w_x = 0, w = 1)
int st_filtfilt_combine(mm_ctx* c, const mm_geom* g, const FilterPlan* plan, const float* in, float* out, const Epi& epi_in,
                        const Pro& pro) {
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    const FilterPlan* p[1] = {plan};
    const float* i1[1] = {in};
    float* o1[1] = {B.E[0]};
    MM_TRY(sweep_fwd(c, g, 1, 1, p, i1, o1, pro, plan->pad));
    const float* i2[1] = {B.E[0]};
    float* o2[1] = {out};
    MM_TRY(sweep_bwd(c, g, 1, p, i2, o2, 1, epi_in, plan->pad));
    return 0;
}

// apply_style_eq (backend/app/pipeline.py:1401-1434)
int st_style_eq(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* gain_db, float* peak, int* fired,
                int reset_peak) {
    const double nyq = g->sr / 2.0;
    const double lo[5] = {30.0, 90.0, 700.0, 2800.0, 10000.0};
    const double hi[5] = {90.0, 280.0, 2800.0, 9000.0, std::min(g->sr * 0.46, 18000.0)};
    int last = -1;
    for (int b = 0; b < 5; ++b) {
        if (std::fabs(gain_db[b]) < 0.05) continue;
        const double l = std::min(lo[b] / nyq, 0.98), h = std::min(hi[b] / nyq, 0.98);
        if (l >= h) continue;
        last = b;
    }
    const float* cur = in;
    int n_fired = 0;
    for (int b = 0; b < 5; ++b) {
        if (std::fabs(gain_db[b]) < 0.05) continue;
        const double l = std::min(lo[b] / nyq, 0.98), h = std::min(hi[b] / nyq, 0.98);
        if (l >= h) continue;
        const double wb = std::pow(10.0, gain_db[b] / 20.0) - 1.0;
        const FilterPlan* p = plan_butter(c, 1, kBand, l, h, std::fabs(wb) <= 0.3 ? PREC_F32 : PREC_AUTO);
        if (!p) return 1;
        Epi e;
        e.mode = EPI_COMBINE;
        e.aux0 = cur;
        e.w[0] = wb;
        e.peak = (b == last) ? peak : nullptr;
        if (e.peak && reset_peak) MM_CUDA(cudaMemsetAsync(peak, 0, (size_t)g->tracks * sizeof(float), c->stream));
        Pro none;
        MM_TRY(st_filtfilt_combine(c, g, p, cur, out, e, none));
        cur = out;
        ++n_fired;
    }
    if (fired) *fired = n_fired;
    if (n_fired == 0 && in != out) {
        MM_CUDA(cudaMemcpyAsync(out, in, (size_t)g->tracks * g->channels * g->stride * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

// apply_harmonic_exciter, oversample == 1 (backend/app/pipeline.py:1267-1326)
int st_exciter(mm_ctx* c, const mm_geom* g, const float* in, float* out, double exciter_db, int mode, float* peak) {
    const double nyq = g->sr / 2.0;
    const FilterPlan* p = plan_butter(c, 2, kHigh, std::min(6000.0 / nyq, 0.97), 0, PREC_F32);   // side chain, scaled by (10^(dB/20) - 1) / 4
    if (!p) return 1;
    Epi e;
    e.mode = EPI_EXCITER;
    e.aux0 = in;
    e.exc_gain = std::pow(10.0, exciter_db / 20.0) - 1.0;
    e.exc_mode = (mode >= 0 && mode <= 4) ? mode : 0;
    e.exc_k = (e.exc_mode == 0) ? 2.5 : 2.0;
    e.peak = peak;
    Pro none;
    return st_filtfilt_combine(c, g, p, in, out, e, none);
}

}  // namespace mm
