// C ABI (include/mm_b200.h): context management, stage entry points, whole chains, host-buffer
// drop-in.  Every entry point names the reference function it replaces in the header.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include "context.h"
#include "pw_args.h"
#include "stages_internal.h"

namespace mm {
const char* get_error();

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- per-track stats assembly ---------------------------------------------------------------------
struct StatsArgs {
    mm_track_stats* out;
    int tracks, channels;
    const double *lufs_in, *lufs_mid, *lufs_out, *gain_db, *peak_in, *peak_out, *mean_row, *nonfinite;
};
__global__ void stats_kernel(const StatsArgs P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.tracks) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    mm_track_stats s;
    s.lufs_in = P.lufs_in ? P.lufs_in[t] : nan;
    s.lufs_mid = P.lufs_mid ? P.lufs_mid[t] : nan;
    s.lufs_out = P.lufs_out ? P.lufs_out[t] : nan;
    s.gain_db = P.gain_db ? P.gain_db[t] : nan;
    s.peak_in = P.peak_in ? P.peak_in[t] : nan;
    s.peak_out = P.peak_out ? P.peak_out[t] : nan;
    s.mean[0] = P.mean_row ? P.mean_row[t * P.channels] : nan;
    s.mean[1] = P.mean_row ? P.mean_row[t * P.channels + (P.channels > 1)] : nan;
    s.nonfinite = P.nonfinite ? P.nonfinite[t] : 0.0;
    P.out[t] = s;
}

static void set_identity_dyn(DynParams* d) { memset(d, 0, sizeof(*d)); }

static int pw_base(PwArgs* A, const float* in, float* out, int mode) {
    memset(A, 0, sizeof(*A));
    A->in = in;
    A->out = out;
    A->mode = mode;
    set_identity_dyn(&A->dyn);
    return 0;
}

static inline size_t batch_floats(const mm_geom* g) { return (size_t)g->tracks * g->channels * (size_t)g->stride; }

static int fade_len(const mm_geom* g, double fade_ms) {
    // apply_output_edge_fade_in (backend/app/pipeline.py:152-167)
    if (fade_ms <= 0 || g->sr <= 0) return 0;
    long long nf = (long long)std::nearbyint((double)g->sr * (fade_ms / 1000.0));
    nf = std::max<long long>(2, std::min<long long>(nf, (long long)((double)g->sr * 0.1)));
    return (int)std::min<long long>(nf, g->n);
}

static int upload_doubles(mm_ctx* c, int slot, const double* host, size_t count, double** dev) {
    MM_TRY(arena(c, slot, count, dev));
    // small parameter vectors: synchronous copy from a host temporary is fine (KBs)
    MM_CUDA(cudaMemcpyAsync(*dev, host, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    MM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// whole chains
// ---------------------------------------------------------------------------------------------------
struct StyleSig {
    double eq[5], exciter_db, width, par_mix;
    bool operator==(const StyleSig& o) const { return memcmp(this, &o, sizeof(StyleSig)) == 0; }
};

// ---- lanes --------------------------------------------------------------------------------------------------------------
// One chain is ~21 dependent kernels; each has a prologue (scan tables into shared memory), a tail (the last warps of a persistent
// grid) and a dependent successor, and a small batch cannot fill 148 SMs through all of them.  Measured on B200 (64 x 180 s tracks,
// tools/lanes_probe.py): one-track launches 169 k audio-s/s on one stream, 228 k on two, 251 k on four; four-track launches 243 k on
// one stream, 288 k on two -- ABOVE the 64-track batch on one stream (279 k).  So a call is spread over LANES: sub-batches (device
// entry) or chunks (host entries) go round-robin to child contexts with their own streams and workspaces, all at full grid size.
// Every track's arithmetic is its own and its dither stream is keyed by its index in the call; what a split can change is only where
// a launch cuts its rows into segments (halo-rebuilt start states: 1e-18), i.e. the last float32 bits of a few samples of a
// full-length track, exactly as between batches of different sizes (tools/lanes_soak.py; short tracks: bit for bit).
static int lanes_wanted(const mm_ctx* c, int dflt) {
    if (c->lanes_cfg > 0) return c->lanes_cfg;
    if (const char* e = getenv("MM_LANES")) { const int v = atoi(e); if (v > 0) return std::min(v, 8); }
    return dflt;
}
static int get_lane(mm_ctx* c, int i, mm_ctx** out) {
    if (i == 0) { *out = c; return 0; }
    while ((int)c->lanes.size() < i) {
        mm_ctx* l = new mm_ctx();
        l->device = c->device;
        l->grid_div = c->grid_div;
        if (cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) != cudaSuccess) { delete l; set_error("lane: cudaStreamCreate failed"); return 1; }
        l->own_stream = true;
        if (cudaEventCreateWithFlags(&l->ev_join, cudaEventDisableTiming) != cudaSuccess) { cudaStreamDestroy(l->stream); delete l; set_error("lane: cudaEventCreate failed"); return 1; }
        c->lanes.push_back(l);
    }
    *out = c->lanes[i - 1];
    return 0;
}
// lane streams start after everything queued on the parent's stream so far
static int lanes_fork(mm_ctx* c, int L) {
    if (L <= 1) return 0;
    if (!c->ev_fork) MM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    MM_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    for (int l = 1; l < L; ++l) {
        mm_ctx* cl;
        MM_TRY(get_lane(c, l, &cl));
        plan_gc(cl);
        cl->track_ids_host = c->track_ids_host;
        MM_CUDA(cudaStreamWaitEvent(cl->stream, c->ev_fork, 0));
    }
    return 0;
}
// ... and the parent's stream continues after everything the lanes have queued
static int lanes_join(mm_ctx* c, int L) {
    for (int l = 1; l < L && l <= (int)c->lanes.size(); ++l) {
        mm_ctx* cl = c->lanes[l - 1];
        cl->track_ids_host = nullptr;
        cl->track_ids_dev = nullptr;
        MM_CUDA(cudaEventRecord(cl->ev_join, cl->stream));
        MM_CUDA(cudaStreamWaitEvent(c->stream, cl->ev_join, 0));
    }
    return 0;
}

static int master_impl(mm_ctx* c, const mm_geom* g, int chain, const mm_style* styles, const float* in, float* out,
                       int16_t* pcm, const float* noise, uint64_t seed, mm_track_stats* stats_dev, uint32_t flags);

// device-resident batch: contiguous runs of tracks go to the lanes (default 2: halves of the batch)
static int master_lanes(mm_ctx* c, const mm_geom* g, int chain, const mm_style* styles, const float* in, float* out,
                        int16_t* pcm, const float* noise, uint64_t seed, mm_track_stats* stats_dev, uint32_t flags) {
    MM_TRY(check_geom(g));
    int L = std::min(lanes_wanted(c, 2), (int)g->tracks);
    if (c->timing || c->slice || !styles || !in || !out) L = 1;      // per-kernel events time kernels ALONE; a time slice is one track
    if (L <= 1) return master_impl(c, g, chain, styles, in, out, pcm, noise, seed, stats_dev, flags);
    MM_TRY(lanes_fork(c, L));
    int rc = 0;
    const size_t row_floats = (size_t)g->channels * (size_t)g->stride, frame_samples = (size_t)g->n * g->channels;
    for (int l = 0; l < L && rc == 0; ++l) {
        const int t0 = (int)((long long)g->tracks * l / L), t1 = (int)((long long)g->tracks * (l + 1) / L);
        mm_ctx* cl;
        if ((rc = get_lane(c, l, &cl)) != 0) break;
        mm_geom gl = *g;
        gl.tracks = t1 - t0;
        gl.track_base = g->track_base + t0;
        rc = master_impl(cl, &gl, chain, styles + t0, in + t0 * row_floats, out + t0 * row_floats, pcm ? pcm + t0 * frame_samples : nullptr,
                         noise ? noise + t0 * frame_samples : nullptr, seed, stats_dev ? stats_dev + t0 : nullptr, flags);
    }
    const int rj = lanes_join(c, L);
    return rc ? rc : rj;
}

static int master_impl(mm_ctx* c, const mm_geom* g, int chain, const mm_style* styles, const float* in, float* out,
                       int16_t* pcm, const float* noise, uint64_t seed, mm_track_stats* stats_dev, uint32_t flags) {
    MM_TRY(check_geom(g));
    if (chain != MM_CHAIN_V1 && chain != MM_CHAIN_V2) { set_error("unknown chain id %d", chain); return 1; }
    if (!styles || !in || !out) { set_error("mm_dev_master: styles, in and out are required"); return 1; }
    const int T = g->tracks, C = g->channels, rows = T * C;
    const bool v1 = chain == MM_CHAIN_V1;

    // per-track parameter vectors
    std::vector<double> h_target(T), h_width(T), h_parmix(rows);
    std::vector<StyleSig> sig(T);
    for (int t = 0; t < T; ++t) {
        const mm_style& s = styles[t];
        h_target[t] = s.target_lufs;
        memset(&sig[t], 0, sizeof(StyleSig));
        for (int b = 0; b < 5; ++b) sig[t].eq[b] = s.eq_gain_db[b];
        // gates: v1 pipeline.py:1889 (exciter_db > 0.05), :1894 (|w - 1| > 0.01), :1856 (mix > 0.01);
        //        v2 chain.py:120-121 (enabled flags |db| >= 0.05, |w - 1| >= 0.01)
        const bool exc = v1 ? (s.exciter_db > 0.05) : (std::fabs(s.exciter_db) >= 0.05);
        const bool img = C == 2 && (v1 ? (std::fabs(s.imager_width - 1.0) > 0.01) : (std::fabs(s.imager_width - 1.0) >= 0.01));
        const bool par = v1 && s.parallel_mix > 0.01;
        sig[t].exciter_db = exc ? s.exciter_db : 0.0;
        sig[t].width = img ? s.imager_width : 1.0;
        sig[t].par_mix = par ? std::min(std::max(s.parallel_mix, 0.0), 1.0) : 0.0;
        h_width[t] = sig[t].width;
        for (int ch = 0; ch < C; ++ch) h_parmix[t * C + ch] = sig[t].par_mix;
    }
    bool any_par = false, any_img = false;
    for (int t = 0; t < T; ++t) { any_par |= sig[t].par_mix >= 0.01; any_img |= sig[t].width != 1.0; }

    double *d_target, *d_width, *d_parmix = nullptr;
    MM_TRY(upload_doubles(c, SL_TARGET, h_target.data(), T, &d_target));
    MM_TRY(upload_doubles(c, SL_WIDTH, h_width.data(), T, &d_width));
    if (any_par) MM_TRY(upload_doubles(c, SL_PARMIX, h_parmix.data(), rows, &d_parmix));

    double *d_sub, *d_mul, *d_gain, *d_mulout, *d_lufs_mid, *d_gaindb, *d_peakin, *d_peakout, *d_mean, *d_nonfinite;
    double *d_lufs_in = nullptr, *d_lufs_out = nullptr;
    float* d_peakbits;
    MM_TRY(arena(c, SL_SUB, (size_t)rows, &d_sub));
    MM_TRY(arena(c, SL_MUL, (size_t)rows, &d_mul));
    MM_TRY(arena(c, SL_GAIN, (size_t)rows, &d_gain));
    MM_TRY(arena(c, SL_MUL_OUT, (size_t)rows, &d_mulout));
    MM_TRY(arena(c, SL_LUFS, (size_t)T, &d_lufs_mid));
    MM_TRY(arena(c, SL_GAINDB, (size_t)T, &d_gaindb));
    MM_TRY(arena(c, SL_PEAKIN, (size_t)T, &d_peakin));
    MM_TRY(arena(c, SL_MISC, (size_t)T, &d_peakout));
    MM_TRY(arena(c, SL_MEAN, (size_t)rows, &d_mean));
    MM_TRY(arena(c, SL_NONFINITE, (size_t)T, &d_nonfinite));
    MM_TRY(arena(c, SL_PEAKBITS, (size_t)T, &d_peakbits));
    MM_CUDA(cudaMemsetAsync(d_nonfinite, 0, (size_t)T * sizeof(double), c->stream));
    c->track_ids_dev = nullptr;
    if (c->track_ids_host) {
        int* d_ids;
        MM_TRY(arena(c, SL_TRACKIDS, (size_t)T, &d_ids));
        // pageable source: staged before the call returns, so no host synchronisation is needed (one would put a bubble between
        // the chunks of the host pipeline)
        MM_CUDA(cudaMemcpyAsync(d_ids, c->track_ids_host + g->track_base, (size_t)T * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        c->track_ids_dev = d_ids;
    }

    // 1. remove_dc_offset + remove_intersample_peaks(0.5): one statistics read; the affine map rides
    //    as the prologue of the first sweeps (pipeline.py:134-149, :1833-1837; chain.py:112-113)
    RowStats* st;
    MM_TRY(run_row_stats(c, g, in, &st));
    const mm_slice* sl = c->slice;
    auto reduce = [&](void* ptr, int64_t count, int dtype, int op, const char* what) -> int {
        return slice_allreduce(c, ptr, count, dtype, op, what);
    };
    MM_TRY(exchange_row_stats(c, st, rows));     // time slices: sums / minima / maxima over every rank's own frames
    MM_TRY(run_in_scalars(c, g, st, 1, 1, 0.5, d_sub, d_mul, d_peakin, d_mean));
    Pro pin;
    pin.mode = PRO_SUBMUL_F32;
    pin.sub = d_sub;
    pin.mul = d_mul;
    if (flags & MM_FLAG_MEASURE_IN) {
        MM_TRY(arena(c, SL_LUFS2, (size_t)T, &d_lufs_in));
        Pro none;
        MM_TRY(st_lufs(c, g, in, none, d_lufs_in, nullptr, nullptr, nullptr));
    }
    // 2. apply_target_curve (pipeline.py:1841 / chain.py:114)
    MM_TRY(st_target_curve(c, g, in, out, pin));
    // 3. v1 only: apply_deesser (pipeline.py:1845)
    if (v1) MM_TRY(st_deesser(c, g, out, out, -6.0, 3.0, 5000.0, 9000.0, 4.0, 85.0));
    // 4. apply_dynamics (+ style-driven parallel compression in v1, pipeline.py:1850-1858)
    {
        const double v2x[3] = {214.0, 2230.0, 10000.0};   // chain.py:116
        MM_TRY(st_dynamics(c, g, out, out, 6.0, v1 ? nullptr : v2x, nullptr, 12.0, any_par ? d_parmix : nullptr, nullptr, 0,
                           (flags & MM_FLAG_ENVELOPE_COMPRESSOR) ? MM_COMPRESSOR_ENVELOPE : MM_COMPRESSOR_SOFT_KNEE));
    }
    // 5. normalize_lufs: measure, derive the gain; the multiply rides as the next prologue
    Pro none;
    MM_TRY(st_lufs(c, g, out, none, d_lufs_mid, d_target, d_gain, d_gaindb));
    Pro pgain;
    pgain.mode = PRO_MUL_F64;
    pgain.mul = d_gain;
    // 6. apply_final_spectral_balance; the output peak is tracked by the last epilogue that touches
    //    a track's samples
    MM_CUDA(cudaMemsetAsync(d_peakbits, 0, (size_t)T * sizeof(float), c->stream));
    MM_TRY(st_final_balance(c, g, out, out, pgain, d_peakbits));
    // 7./8. apply_style_eq, apply_harmonic_exciter (pipeline.py:1401-1434, :1267-1326).  The five band designs and the exciter's
    //       high-pass are the same for every style; only the recombination weight differs.  So a mixed batch costs ONE forward
    //       and ONE backward sweep per stage (band 0..4, exciter) over the rows whose style fires that stage -- a device row
    //       list -- with the weight read per row, instead of one pair of launches per stage and style group (round 1: ~60
    //       launches for the eight presets, now <= 12).  Each row's last firing stage tracks the track's output peak.
    {
        const double nyq = g->sr / 2.0;
        const double lo[5] = {30.0, 90.0, 700.0, 2800.0, 10000.0};
        const double hi[5] = {90.0, 280.0, 2800.0, 9000.0, std::min(g->sr * 0.46, 18000.0)};
        constexpr int kStages = 6;                                 // 5 style-EQ bands + exciter
        std::vector<double> h_w((size_t)kStages * rows, 0.0);      // per stage and row: band weight / exciter gain (0 = does not fire)
        std::vector<unsigned char> h_last((size_t)kStages * rows, 0);
        std::vector<int> last_stage(T, -1);
        bool any_fires = false;
        for (int t = 0; t < T; ++t) {
            for (int b = 0; b < 5; ++b) {
                const double l = std::min(lo[b] / nyq, 0.98), h = std::min(hi[b] / nyq, 0.98);
                if (std::fabs(sig[t].eq[b]) < 0.05 || l >= h) continue;                    // pipeline.py:1421-1426
                const double wb = std::pow(10.0, sig[t].eq[b] / 20.0) - 1.0;
                for (int ch = 0; ch < C; ++ch) h_w[(size_t)b * rows + t * C + ch] = wb;
                last_stage[t] = b;
            }
            if (sig[t].exciter_db != 0.0) {
                for (int ch = 0; ch < C; ++ch) h_w[(size_t)5 * rows + t * C + ch] = std::pow(10.0, sig[t].exciter_db / 20.0) - 1.0;
                last_stage[t] = 5;
            }
            if (last_stage[t] >= 0) {
                any_fires = true;
                for (int ch = 0; ch < C; ++ch) h_last[(size_t)last_stage[t] * rows + t * C + ch] = 1;
            }
        }
        if (any_fires) {
            // launch lists: per stage, the firing rows split by the precision class of their weight (style_eq: float32 pass 2 behind
            // weights <= 0.3, the automatic policy otherwise -- the choice a single-track run makes, so batch == single bit for bit)
            struct Launch { int stage, cls, off, cnt; bool uniform; double w; };
            std::vector<Launch> launches;
            std::vector<int> rowlist;
            for (int st_ = 0; st_ < kStages; ++st_)
                for (int cls = 0; cls < (st_ < 5 ? 2 : 1); ++cls) {
                    Launch L{st_, cls, (int)rowlist.size(), 0, true, 0.0};
                    for (int r = 0; r < rows; ++r) {
                        const double w = h_w[(size_t)st_ * rows + r];
                        if (w == 0.0) continue;
                        if (st_ < 5 && (std::fabs(w) <= 0.3 ? 0 : 1) != cls) continue;
                        if (L.cnt == 0) L.w = w; else if (w != L.w) L.uniform = false;
                        rowlist.push_back(r);
                        ++L.cnt;
                    }
                    if (L.cnt) launches.push_back(L);
                }
            int* d_rows = nullptr;
            double* d_w = nullptr;
            unsigned char* d_last = nullptr;
            MM_TRY(arena(c, SL_ROWMAP, rowlist.size(), &d_rows));
            MM_TRY(arena(c, SL_ROWW, h_w.size(), &d_w));
            MM_TRY(arena(c, SL_ROWFLAG, h_last.size(), &d_last));
            MM_CUDA(cudaMemcpyAsync(d_rows, rowlist.data(), rowlist.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            MM_CUDA(cudaMemcpyAsync(d_w, h_w.data(), h_w.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            MM_CUDA(cudaMemcpyAsync(d_last, h_last.data(), h_last.size(), cudaMemcpyHostToDevice, c->stream));
            // tracks with a firing stage re-track their output peak from scratch in their last stage
            for (int t0 = 0; t0 < T;) {
                int t1 = t0 + 1;
                const bool on = last_stage[t0] >= 0;
                while (t1 < T && (last_stage[t1] >= 0) == on) ++t1;
                if (on) MM_CUDA(cudaMemsetAsync(d_peakbits + t0, 0, (size_t)(t1 - t0) * sizeof(float), c->stream));
                t0 = t1;
            }
            MM_CUDA(cudaStreamSynchronize(c->stream));           // the host vectors above are temporaries
            for (const Launch& L : launches) {
                const bool all_rows = L.cnt == rows;             // rows come out in increasing order: the list is the identity
                c->row_map = all_rows ? nullptr : d_rows + L.off;
                c->row_map_rows = all_rows ? 0 : L.cnt;
                Epi e;
                e.aux0 = out;
                e.peak = d_peakbits;
                e.peak_row = d_last + (size_t)L.stage * rows;
                const FilterPlan* p;
                if (L.stage < 5) {
                    const int bnd = L.stage;
                    p = plan_butter(c, 1, kBand, std::min(lo[bnd] / nyq, 0.98), std::min(hi[bnd] / nyq, 0.98), L.cls == 0 ? PREC_F32 : PREC_AUTO);
                    e.mode = EPI_COMBINE;
                    e.w[0] = L.w;
                    if (!L.uniform) e.w_row = d_w + (size_t)L.stage * rows;
                } else {
                    p = plan_butter(c, 2, kHigh, std::min(6000.0 / nyq, 0.97), 0, PREC_F32);      // as st_exciter
                    e.mode = EPI_EXCITER;
                    e.exc_gain = L.w;
                    e.exc_mode = 0;
                    e.exc_k = 2.5;
                    if (!L.uniform) e.exc_row = d_w + (size_t)L.stage * rows;
                }
                int rc = p ? 0 : 1;
                Pro none2;
                if (rc == 0) rc = st_filtfilt_combine(c, g, p, out, out, e, none2);
                c->row_map = nullptr;
                c->row_map_rows = 0;
                if (rc != 0) return rc;
            }
        }
    }
    // 9. apply_stereo_imager: folded into the final pass; tracks with an active imager need their
    //    post-imager peak first (read-only pass over those runs)
    if (any_img) {
        // one launch over the whole batch: CTAs of tracks without an active imager return immediately (round 1 launched once per
        // run of consecutive imager tracks -- 32 small launches for the eight presets cycling over 128 tracks)
        MM_TRY(reset_imager_peaks(c, d_peakbits, d_width, T));
        mm_geom sub = *g;
        PwArgs A;
        pw_base(&A, out + (sl ? sl->own_lo : 0), nullptr, PW_PEAK);
        if (sl) sub.n = sl->own_hi - sl->own_lo;
        A.width = d_width;
        A.peak = d_peakbits;
        A.skip_unity = 1;
        MM_TRY(run_pointwise(c, &sub, A, "peak_after_imager"));
    }
    // 10. remove_intersample_peaks(0.5) + clip/nan_to_num + 6 ms fade-in (+ TPDF dither to int16)
    {
        MM_TRY(reduce(d_peakbits, T, 2, 2, "the output peak"));
        OutScalarArgs O;
        O.peak_bits = d_peakbits; O.tracks = T; O.channels = C;
        O.limit = (float)std::pow(10.0, -0.5 / 20.0);
        O.mul = d_mulout; O.peak_track = d_peakout;
        MM_TRY(run_out_scalars(c, O));
        const bool fade = (v1 || !(flags & MM_FLAG_NO_JOB_FADE)) && !(sl && sl->global_off != 0);   // the file's first frames only
        MM_TRY(run_finalize(c, g, out, out, d_mulout, any_img ? d_width : nullptr, fade ? fade_len(g, 6.0) : 0, pcm, noise, seed,
                            d_nonfinite));
    }
    if (flags & MM_FLAG_MEASURE_OUT) {
        MM_TRY(arena(c, SL_LUFS3, (size_t)T, &d_lufs_out));
        MM_TRY(st_lufs(c, g, out, none, d_lufs_out, nullptr, nullptr, nullptr));
    }
    if (stats_dev) {
        StatsArgs S;
        S.out = stats_dev; S.tracks = T; S.channels = C;
        S.lufs_in = d_lufs_in; S.lufs_mid = d_lufs_mid; S.lufs_out = d_lufs_out; S.gain_db = d_gaindb;
        S.peak_in = d_peakin; S.peak_out = d_peakout; S.mean_row = d_mean; S.nonfinite = d_nonfinite;
        KernelScope ks(c, "track_stats");
        stats_kernel<<<(T + 127) / 128, 128, 0, c->stream>>>(S);
        MM_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace mm

using namespace mm;

// ---- pageable host buffers: multi-threaded staging through pinned memory -------------------------------------------------
// A pageable cudaMemcpy of a 3-minute track (63.5 MB) runs at 3-6 GB/s and one memcpy thread at ~10 GB/s, against ~50 GB/s for
// a pinned DMA: for ONE job (the reference's product case, jobs_store.py:19-20) the two staging copies were 12 of its 20 ms.
// Here `T` threads copy blocks into / out of a pinned slot while the DMA engine moves the blocks already (still) in order.
namespace mm {
static int host_threads() {
    static const int t = [] {
        int v = 0;
        if (const char* e = getenv("MM_HOST_THREADS")) v = atoi(e);
        if (v <= 0) { const unsigned hw = std::thread::hardware_concurrency(); v = (int)std::min<unsigned>(8u, std::max<unsigned>(1u, hw / 2)); }
        return std::min(v, 32);
    }();
    return t;
}
constexpr size_t kHostBlk = (size_t)4 << 20;

// plain parallel memcpy (both sides host memory); small copies stay on the calling thread
void host_par_memcpy(void* dst, const void* src, size_t bytes) {
    const int T = host_threads();
    if (T <= 1 || bytes < 2 * kHostBlk) { memcpy(dst, src, bytes); return; }
    const size_t nblk = (bytes + kHostBlk - 1) / kHostBlk;
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < nblk;)
            memcpy((char*)dst + i * kHostBlk, (const char*)src + i * kHostBlk, std::min(kHostBlk, bytes - i * kHostBlk));
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T && (size_t)t < nblk; ++t) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

static bool host_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
static int host_pin_slot(mm_ctx* c, int i, size_t need, char** out) {
    Slot& hp = c->host_pin[i];
    if (hp.cap < need) {
        if (hp.p) cudaFreeHost(hp.p);
        hp.p = nullptr; hp.cap = 0;
        MM_CUDA(cudaHostAlloc(&hp.p, need, cudaHostAllocDefault));
        hp.cap = need;
    }
    *out = reinterpret_cast<char*>(hp.p);
    return 0;
}
constexpr size_t kHostRound = (size_t)256 << 20;      // pinned staging per round (a 180 s stereo float32 track is 63.5 MB)

// host (pageable) -> device on the context stream; returns once the host buffer has been read (the DMA may still be in flight)
static int pin_in_wait(mm_ctx* c) {        // an earlier call's DMA may still be reading host_pin[0]
    if (c->pin_in_done) MM_CUDA(cudaEventSynchronize(c->pin_in_done));
    return 0;
}
static int staged_copy_in(mm_ctx* c, char* dev, const char* src, size_t bytes) {
    MM_TRY(pin_in_wait(c));
    if (!c->pin_in_done) MM_CUDA(cudaEventCreateWithFlags(&c->pin_in_done, cudaEventDisableTiming));
    for (size_t r0 = 0; r0 < bytes; r0 += kHostRound) {
        const size_t rb = std::min(kHostRound, bytes - r0);
        if (r0) MM_CUDA(cudaStreamSynchronize(c->stream));         // the slot is read by the previous round's DMA
        char* pin;
        MM_TRY(host_pin_slot(c, 0, rb, &pin));
        if (rb <= kHostBlk) {                                      // small: no helper threads
            memcpy(pin, src + r0, rb);
            MM_CUDA(cudaMemcpyAsync(dev + r0, pin, rb, cudaMemcpyHostToDevice, c->stream));
            continue;
        }
        const size_t nblk = (rb + kHostBlk - 1) / kHostBlk;
        std::vector<std::atomic<int>> done(nblk);
        for (auto& d : done) d.store(0, std::memory_order_relaxed);
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t i; (i = next.fetch_add(1)) < nblk;) {
                memcpy(pin + i * kHostBlk, src + r0 + i * kHostBlk, std::min(kHostBlk, rb - i * kHostBlk));
                done[i].store(1, std::memory_order_release);
            }
        };
        const int T = host_threads();
        std::vector<std::thread> th;
        for (int t = 0; t < T && (size_t)t < nblk; ++t) th.emplace_back(work);
        cudaError_t e = cudaSuccess;
        // the calling thread issues a DMA per run of finished blocks, in order (16 MB runs keep the per-copy overhead small)
        for (size_t i = 0; i < nblk;) {
            while (!done[i].load(std::memory_order_acquire)) std::this_thread::yield();
            size_t j = i + 1;
            while (j < nblk && j - i < 4 && done[j].load(std::memory_order_acquire)) ++j;
            const size_t off = i * kHostBlk, len = std::min(j * kHostBlk, rb) - off;
            if (e == cudaSuccess) e = cudaMemcpyAsync(dev + r0 + off, pin + off, len, cudaMemcpyHostToDevice, c->stream);
            i = j;
        }
        for (auto& x : th) x.join();
        if (e != cudaSuccess) { set_error("host->device copy failed: %s", cudaGetErrorString(e)); return 1; }
    }
    MM_CUDA(cudaEventRecord(c->pin_in_done, c->stream));
    return 0;
}

// device -> host (pageable) after everything queued on the context stream; returns when the host buffer is complete
static int staged_copy_out(mm_ctx* c, char* dst, const char* dev, size_t bytes) {
    for (size_t r0 = 0; r0 < bytes; r0 += kHostRound) {
        const size_t rb = std::min(kHostRound, bytes - r0);
        char* pin;
        MM_TRY(host_pin_slot(c, 2, rb, &pin));
        constexpr size_t kRun = 4 * kHostBlk;                      // one event per 16 MB
        const size_t nrun = (rb + kRun - 1) / kRun;
        std::vector<cudaEvent_t> ev(nrun);
        cudaError_t e = cudaSuccess;
        for (size_t i = 0; i < nrun; ++i) {
            const size_t off = i * kRun, len = std::min(kRun, rb - off);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); else ev[i] = nullptr;
            if (e == cudaSuccess) e = cudaMemcpyAsync(pin + off, dev + r0 + off, len, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaEventRecord(ev[i], c->stream);
        }
        std::atomic<int> bad{e != cudaSuccess};
        if (e == cudaSuccess) {
            const size_t nblk = (rb + kHostBlk - 1) / kHostBlk;
            std::atomic<size_t> next{0};
            auto work = [&] {
                for (size_t i; (i = next.fetch_add(1)) < nblk;) {
                    if (cudaEventSynchronize(ev[i * kHostBlk / kRun]) != cudaSuccess) { bad.store(1); return; }
                    memcpy(dst + r0 + i * kHostBlk, pin + i * kHostBlk, std::min(kHostBlk, rb - i * kHostBlk));
                }
            };
            const int T = host_threads();
            std::vector<std::thread> th;
            for (int t = 1; t < T && (size_t)t < nblk; ++t) th.emplace_back([&, dv = c->device] { cudaSetDevice(dv); work(); });
            work();
            for (auto& x : th) x.join();
        }
        cudaStreamSynchronize(c->stream);
        for (cudaEvent_t v : ev) if (v) cudaEventDestroy(v);
        if (bad.load()) { set_error("device->host copy failed: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError())); return 1; }
    }
    return 0;
}
}  // namespace mm

#ifndef MM_D2H_STREAMS_DEFAULT
#define MM_D2H_STREAMS_DEFAULT 1
#endif
#define MM_API_BEGIN(ctx)                                   \
    if (!(ctx)) { mm::set_error("null context"); return 1; } \
    mm::DeviceGuard _guard((ctx)->device);                   \
    mm::plan_gc(ctx);

extern "C" {

int mm_abi_version(void) { return MM_ABI_VERSION; }
const char* mm_last_error(void) { return mm::get_error(); }

int64_t mm_row_stride(int64_t n) {
    if (n < 0) n = 0;
    return ((int64_t)kLead + n + 32 + 31) & ~(int64_t)31;
}

int mm_ctx_create(int device, void* stream, mm_ctx** out) {
    if (!out) { set_error("mm_ctx_create: out is null"); return 1; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return 1;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (%d devices)", device, count); return 1; }
    MM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library holds sm_100a code only", device, prop.major, prop.minor); return 1; }
    mm_ctx* c = new mm_ctx();
    c->device = device;
    if (const char* e = getenv("MM_GRID_DIV")) c->grid_div = std::max(1, atoi(e));      // experiments: several contexts sharing the device
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; set_error("cudaStreamCreate failed"); return 1; }
        c->own_stream = true;
    }
    *out = c;
    return 0;
}

void mm_ctx_destroy(mm_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    for (mm_ctx* l : c->lanes) mm_ctx_destroy(l);
    c->lanes.clear();
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    bigfft_release(c);
    for (int i = 0; i < SL_COUNT; ++i) if (c->slots[i].p) cudaFree(c->slots[i].p);
    for (auto& kv : c->plans) if (kv.second.dev) cudaFree(kv.second.dev);
    for (auto& kv : c->kw_plans) { if (kv.second.dev) cudaFree(kv.second.dev); if (kv.second.plane64) cudaFree(kv.second.plane64); }
    for (auto& kv : c->lp_taps) if (kv.second) cudaFree(kv.second);
    for (auto& kv : c->lufs_plans) {
        cudaFree(kv.second.bnd); cudaFree(kv.second.tile_seg); cudaFree(kv.second.blk_lo); cudaFree(kv.second.blk_hi);
    }
    for (auto& k : c->ktimes) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
    for (Slot& hp : c->host_pin) if (hp.p) cudaFreeHost(hp.p);
    if (c->pin_in_done) cudaEventDestroy(c->pin_in_done);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->d2h_stream2) cudaStreamDestroy(c->d2h_stream2);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mm_ctx_release_workspace(mm_ctx* c) {
    MM_API_BEGIN(c);
    MM_CUDA(cudaStreamSynchronize(c->stream));
    for (mm_ctx* l : c->lanes) MM_TRY(mm_ctx_release_workspace(l));
    bigfft_release(c);
    for (int i = 0; i < SL_COUNT; ++i)
        if (c->slots[i].p) { cudaFree(c->slots[i].p); c->slots[i].p = nullptr; c->slots[i].cap = 0; }
    c->workspace_bytes = 0;
    for (Slot& hp : c->host_pin) { if (hp.p) cudaFreeHost(hp.p); hp.p = nullptr; hp.cap = 0; }
    for (auto& kv : c->lufs_plans) { cudaFree(kv.second.bnd); cudaFree(kv.second.tile_seg); cudaFree(kv.second.blk_lo); cudaFree(kv.second.blk_hi); }
    c->lufs_plans.clear();
    return 0;
}

int mm_ctx_sync(mm_ctx* c) {
    MM_API_BEGIN(c);
    MM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int64_t mm_ctx_launch_count(mm_ctx* c) {
    if (!c) return 0;
    int64_t n = c->launches;
    for (const mm_ctx* l : c->lanes) n += l->launches;
    return n;
}

int mm_ctx_set_lanes(mm_ctx* c, int lanes) {
    MM_API_BEGIN(c);
    if (lanes < 0 || lanes > 8) { set_error("mm_ctx_set_lanes: 0 (automatic) .. 8"); return 1; }
    c->lanes_cfg = lanes;
    return 0;
}

int mm_ctx_timing(mm_ctx* c, int enable) {
    MM_API_BEGIN(c);
    MM_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& k : c->ktimes) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
    c->ktimes.clear();
    c->kacc.clear();
    c->timing = enable != 0;
    // per-kernel events time kernels alone: calls stay on the context's own stream while timing is on, and the lanes' scratch is
    // handed back so that lane 0 can grow to the whole batch
    if (enable) for (mm_ctx* l : c->lanes) MM_TRY(mm_ctx_release_workspace(l));
    return 0;
}

int mm_ctx_kernel_times(mm_ctx* c, mm_ktime* out, int cap, int* count) {
    MM_API_BEGIN(c);
    MM_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& k : c->ktimes) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, k.a, k.b) == cudaSuccess) {
            auto& acc = c->kacc[k.name];
            acc.ms += (double)ms;
            acc.launches += 1;
            acc.samples += k.samples;
        }
        cudaEventDestroy(k.a);
        cudaEventDestroy(k.b);
    }
    c->ktimes.clear();
    int i = 0;
    for (auto& kv : c->kacc) {
        if (i < cap && out) {
            memset(&out[i], 0, sizeof(mm_ktime));
            strncpy(out[i].name, kv.first.c_str(), sizeof(out[i].name) - 1);
            out[i].ms = kv.second.ms;
            out[i].launches = kv.second.launches;
            out[i].samples = kv.second.samples;
        }
        ++i;
    }
    if (count) *count = i;
    return 0;
}

int64_t mm_ctx_workspace_bytes(mm_ctx* c) {
    if (!c) return 0;
    int64_t n = c->workspace_bytes;
    for (const mm_ctx* l : c->lanes) n += l->workspace_bytes;
    return n;
}

int64_t mm_master_workspace_bytes(const mm_geom* g, int chain) {
    if (!g) return 0;
    const int64_t buf = (int64_t)g->tracks * g->channels * g->stride * 4;
    (void)chain;
    return buf * 9 + (1 << 20);   // E0..E3, T0..T4 (stages.cu get_bufs) + scalars
}

// ---- pinned host memory for the host-buffer entry points ----------------------------------------
int mm_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes <= 0) { set_error("mm_host_alloc: bad arguments"); return 1; }
    MM_CUDA(cudaMallocHost(out, (size_t)bytes));
    return 0;
}
int mm_host_free(void* p) {
    if (p) MM_CUDA(cudaFreeHost(p));
    return 0;
}

// ---- host <-> device transfers of the Python mirror (Engine.upload / download): pinned buffers go as one DMA, pageable ones
// (a numpy array: what run_mastering_pipeline and every stage function receive and return) are staged by several threads ----
int mm_ctx_copy_in(mm_ctx* c, void* dev_dst, const void* host_src, int64_t bytes) {
    MM_API_BEGIN(c);
    if (!dev_dst || !host_src || bytes < 0) { set_error("mm_ctx_copy_in: bad arguments"); return 1; }
    if (bytes == 0) return 0;
    if (host_is_pinned(host_src)) { MM_CUDA(cudaMemcpyAsync(dev_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream)); return 0; }
    return staged_copy_in(c, reinterpret_cast<char*>(dev_dst), reinterpret_cast<const char*>(host_src), (size_t)bytes);
}
int mm_ctx_copy_out(mm_ctx* c, void* host_dst, const void* dev_src, int64_t bytes) {
    MM_API_BEGIN(c);
    if (!host_dst || !dev_src || bytes < 0) { set_error("mm_ctx_copy_out: bad arguments"); return 1; }
    if (bytes == 0) return 0;
    if (host_is_pinned(host_dst)) {
        MM_CUDA(cudaMemcpyAsync(host_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
        MM_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    }
    return staged_copy_out(c, reinterpret_cast<char*>(host_dst), reinterpret_cast<const char*>(dev_src), (size_t)bytes);
}

// ---- layout helpers ---------------------------------------------------------------------------
int mm_dev_deinterleave(mm_ctx* c, const mm_geom* g, const float* il, float* pl) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return run_layout(c, g, il, pl, 0);
}
int mm_dev_interleave(mm_ctx* c, const mm_geom* g, const float* pl, float* il) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return run_layout(c, g, il, const_cast<float*>(pl), 1);
}

// ---- stage functions ----------------------------------------------------------------------------
int mm_dev_remove_dc_offset(mm_ctx* c, const mm_geom* g, const float* in, float* out) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    const int rows = g->tracks * g->channels;
    double *sub, *mul;
    MM_TRY(arena(c, SL_SUB, (size_t)rows, &sub));
    MM_TRY(arena(c, SL_MUL, (size_t)rows, &mul));
    RowStats* st;
    MM_TRY(run_row_stats(c, g, in, &st));
    MM_TRY(run_in_scalars(c, g, st, 1, 0, 0.0, sub, mul, nullptr, nullptr));
    PwArgs A;
    pw_base(&A, in, out, PW_AFFINE);
    A.sub = sub;
    A.mul = mul;
    return run_pointwise(c, g, A, "remove_dc_offset");
}

int mm_dev_remove_intersample_peaks(mm_ctx* c, const mm_geom* g, const float* in, float* out, double headroom_db) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    const int rows = g->tracks * g->channels;
    double *sub, *mul;
    MM_TRY(arena(c, SL_SUB, (size_t)rows, &sub));
    MM_TRY(arena(c, SL_MUL, (size_t)rows, &mul));
    RowStats* st;
    MM_TRY(run_row_stats(c, g, in, &st));
    MM_TRY(run_in_scalars(c, g, st, 0, 1, headroom_db, sub, mul, nullptr, nullptr));
    PwArgs A;
    pw_base(&A, in, out, PW_AFFINE);
    A.sub = sub;
    A.mul = mul;
    A.clip = 1;
    return run_pointwise(c, g, A, "remove_intersample_peaks");
}

int mm_dev_fade_in(mm_ctx* c, const mm_geom* g, const float* in, float* out, double fade_ms) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    PwArgs A;
    pw_base(&A, in, out, PW_FADE);
    A.n_fade = fade_len(g, fade_ms);
    A.fade_step = A.n_fade > 1 ? 1.0 / (double)(A.n_fade - 1) : 0.0;
    return run_pointwise(c, g, A, "fade_in");
}

int mm_dev_blend(mm_ctx* c, const mm_geom* g, const float* dry, float* out, const float* processed, double amount) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    PwArgs A;
    pw_base(&A, dry, out, PW_BLEND);
    A.in2 = processed;
    A.blend = (float)std::min(std::max(amount, 0.0), 1.0);
    return run_pointwise(c, g, A, "module_blend");
}

int mm_dev_apply_target_curve(mm_ctx* c, const mm_geom* g, const float* in, float* out, int eq_ms) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (eq_ms && g->channels == 2) {
        // pipeline.py:248-255: EQ on mid/side, then decode and clip
        Bufs B;
        MM_TRY(get_bufs(c, g, &B));
        float* ms = B.T[1];
        PwArgs A;
        pw_base(&A, in, ms, PW_MS_ENCODE);
        MM_TRY(run_pointwise(c, g, A, "ms_encode"));
        Pro none;
        MM_TRY(st_target_curve(c, g, ms, ms, none));
        PwArgs D;
        pw_base(&D, ms, out, PW_MS_DECODE);
        return run_pointwise(c, g, D, "ms_decode_clip");
    }
    Pro none;
    return st_target_curve(c, g, in, out, none);
}

int mm_dev_apply_target_curve_linear_phase(mm_ctx* c, const mm_geom* g, const float* in, float* out, int eq_ms) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    if (eq_ms && g->channels == 2) {       // pipeline.py:248-255 with phase_mode="linear_phase" on mid and side
        PwArgs A;
        pw_base(&A, in, B.T[1], PW_MS_ENCODE);
        MM_TRY(run_pointwise(c, g, A, "ms_encode"));
        MM_TRY(st_target_curve_linear_phase(c, g, B.T[1], B.T[2]));
        PwArgs D;
        pw_base(&D, B.T[2], out, PW_MS_DECODE);
        return run_pointwise(c, g, D, "ms_decode_clip");
    }
    if (in != out) return st_target_curve_linear_phase(c, g, in, out);
    MM_TRY(st_target_curve_linear_phase(c, g, in, B.T[1]));
    MM_CUDA(cudaMemcpyAsync(out, B.T[1], batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int mm_design_linear_phase_ir(int sr, int n_fft, float* ir) {
    if (!ir || !linear_phase_target_ir(sr, n_fft, ir)) { set_error("mm_design_linear_phase_ir: bad arguments"); return 1; }
    return 0;
}

int mm_dev_apply_deesser(mm_ctx* c, const mm_geom* g, const float* in, float* out, double threshold_db, double ratio,
                         double freq_lo, double freq_hi, double attack_ms, double release_ms) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_deesser(c, g, in, out, threshold_db, ratio, freq_lo, freq_hi, attack_ms, release_ms);
}

int mm_dev_apply_dynamics(mm_ctx* c, const mm_geom* g, const float* in, float* out, double knee_db, const double* crossovers_hz,
                          const double* band_ratios, double max_upward_boost_db) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_dynamics(c, g, in, out, knee_db, crossovers_hz, band_ratios, max_upward_boost_db, nullptr, nullptr);
}

int mm_dev_apply_maximizer(mm_ctx* c, const mm_geom* g, const float* in, float* out);

int mm_dev_apply_dynamics_mode(mm_ctx* c, const mm_geom* g, const float* in, float* out, double knee_db, const double* crossovers_hz,
                               const double* band_ratios, double max_upward_boost_db, int bands_only, int compressor) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (compressor != MM_COMPRESSOR_SOFT_KNEE && compressor != MM_COMPRESSOR_ENVELOPE) { set_error("unknown compressor mode %d", compressor); return 2; }
    return st_dynamics(c, g, in, out, knee_db, crossovers_hz, band_ratios, max_upward_boost_db, nullptr, nullptr, bands_only ? 1 : 0, compressor);
}

int mm_dev_apply_multiband_dynamics(mm_ctx* c, const mm_geom* g, const float* in, float* out, double knee_db, const double* crossovers_hz,
                                    const double* band_ratios, double max_upward_boost_db) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_dynamics(c, g, in, out, knee_db, crossovers_hz, band_ratios, max_upward_boost_db, nullptr, nullptr, 1);
}

int mm_dev_apply_maximizer_lookahead(mm_ctx* c, const mm_geom* g, const float* in, float* out, double lookahead_ms) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (in == out) { set_error("mm_dev_apply_maximizer_lookahead: not in place"); return 2; }
    const long long delay_n = (long long)((double)g->sr * (lookahead_ms / 1000.0));
    if (delay_n <= 0 || delay_n >= g->n) return mm_dev_apply_maximizer(c, g, in, out);       // pipeline.py:554-555
    const int cf = (int)std::min<long long>(delay_n, std::max(2, (int)((double)g->sr * 0.002)));
    return st_maximizer_lookahead(c, g, in, out, delay_n, cf);
}

int mm_dev_apply_maximizer(mm_ctx* c, const mm_geom* g, const float* in, float* out) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    PwArgs A;
    pw_base(&A, in, out, PW_MAXIMIZER);
    fill_dyn(&A.dyn, 6.0, nullptr, 12.0);
    A.dyn.max_top = (float)std::pow(10.0, -0.3 / 20.0);   // maximizer alone: the ceiling only, no limiter behind it
    return run_pointwise(c, g, A, "apply_maximizer");
}

int mm_dev_apply_parallel_compression(mm_ctx* c, const mm_geom* g, const float* in, float* out, double mix, double ratio,
                                      double threshold_db) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    mix = std::min(std::max(mix, 0.0), 1.0);
    if (mix < 0.01) {   // pipeline.py:1787-1788: unchanged
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    PwArgs A;
    pw_base(&A, in, out, PW_PARALLEL);
    fill_dyn(&A.dyn, 6.0, nullptr, 12.0);
    fill_parallel(&A.dyn, ratio, threshold_db);
    A.par_mix = mix;
    return run_pointwise(c, g, A, "apply_parallel_compression");
}

int mm_dev_measure_lufs(mm_ctx* c, const mm_geom* g, const float* in, double* lufs_dev) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    Pro none;
    return st_lufs(c, g, in, none, lufs_dev, nullptr, nullptr, nullptr);
}

int mm_dev_normalize_lufs(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* target_lufs_host) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    const int rows = g->tracks * g->channels;
    double *target, *gain, *lufs, *gdb;
    MM_TRY(upload_doubles(c, SL_TARGET, target_lufs_host, (size_t)g->tracks, &target));
    MM_TRY(arena(c, SL_GAIN, (size_t)rows, &gain));
    MM_TRY(arena(c, SL_LUFS, (size_t)g->tracks, &lufs));
    MM_TRY(arena(c, SL_GAINDB, (size_t)g->tracks, &gdb));
    Pro none;
    MM_TRY(st_lufs(c, g, in, none, lufs, target, gain, gdb));
    PwArgs A;
    pw_base(&A, in, out, PW_GAIN_F64);
    A.mul = gain;
    return run_pointwise(c, g, A, "normalize_lufs_gain");
}

int mm_dev_apply_final_spectral_balance(mm_ctx* c, const mm_geom* g, const float* in, float* out) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    Pro none;
    return st_final_balance(c, g, in, out, none, nullptr);
}

int mm_dev_apply_style_eq(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* eq_gain_db) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_style_eq(c, g, in, out, eq_gain_db, nullptr, nullptr, 0);
}

int mm_dev_apply_harmonic_exciter(mm_ctx* c, const mm_geom* g, const float* in, float* out, double exciter_db, int mode) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (std::fabs(exciter_db) < 0.05) {   // pipeline.py:1282-1283
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    return st_exciter(c, g, in, out, exciter_db, mode, nullptr);
}

int mm_dev_apply_stereo_imager(mm_ctx* c, const mm_geom* g, const float* in, float* out, double width) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (g->channels != 2) {   // pipeline.py:1355-1356
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    std::vector<double> w((size_t)g->tracks, width);
    double* dw;
    MM_TRY(upload_doubles(c, SL_WIDTH, w.data(), w.size(), &dw));
    PwArgs A;
    pw_base(&A, in, out, PW_IMAGER);
    A.width = dw;
    A.force_imager = 1;
    return run_pointwise(c, g, A, "apply_stereo_imager");
}

int mm_dev_apply_rumble_filter(mm_ctx* c, const mm_geom* g, const float* in, float* out, double cutoff_hz) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    // pipeline.py:1449-1469: butter(2, clip(cutoff, 20, 200) / nyq <= 0.99, 'high') through filtfilt
    const double nyq = g->sr / 2.0;
    const double fc = std::min(std::max(cutoff_hz, 20.0), 200.0);
    const FilterPlan* p = plan_butter(c, 2, kHigh, std::min(fc / nyq, 0.99), 0);
    if (!p) return 1;
    Epi e;
    Pro none;
    return st_filtfilt_combine(c, g, p, in, out, e, none);
}

int mm_dev_apply_dynamic_eq(mm_ctx* c, const mm_geom* g, const float* in, float* out, int nbands, const double* params) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (nbands < 0 || nbands > 64 || (nbands > 0 && !params)) { set_error("mm_dev_apply_dynamic_eq: 0..64 bands with 7 parameters each"); return 2; }
    return st_dynamic_eq(c, g, in, out, nbands, params, 0, nullptr);
}

int mm_dev_apply_dynamic_eq2(mm_ctx* c, const mm_geom* g, const float* in, float* out, int nbands, const double* params, uint32_t flags,
                             int32_t* classes) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (nbands < 0 || nbands > 64 || (nbands > 0 && !params)) { set_error("mm_dev_apply_dynamic_eq2: 0..64 bands with 7 parameters each"); return 2; }
    return st_dynamic_eq(c, g, in, out, nbands, params, flags, classes);
}

int mm_dev_apply_transient_designer(mm_ctx* c, const mm_geom* g, const float* in, float* out, double attack_gain, double sustain_gain) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    // pipeline.py:1749-1752: gains clipped to [0.1, 3]; both within 0.02 of 1 -> the input object is returned
    attack_gain = std::min(std::max(attack_gain, 0.1), 3.0);
    sustain_gain = std::min(std::max(sustain_gain, 0.1), 3.0);
    if (std::fabs(attack_gain - 1.0) < 0.02 && std::fabs(sustain_gain - 1.0) < 0.02) {
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    return st_transient_designer(c, g, in, out, attack_gain, sustain_gain);
}

int mm_dev_apply_maximizer_transient_aware(mm_ctx* c, const mm_geom* g, const float* in, float* out, double sensitivity) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_maximizer_transient_aware(c, g, in, out, sensitivity);
}

int mm_dev_apply_high_freq_trim(mm_ctx* c, const mm_geom* g, const float* in, float* out, double crossover_hz, double high_gain) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (std::fabs(high_gain - 1.0) < 0.001) {     // pipeline.py:1719-1720
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    // low + g (x - low) = g x + (1 - g) low, then clip (pipeline.py:1724-1730)
    const FilterPlan* p = plan_butter(c, 2, kLow, std::min(crossover_hz / (g->sr / 2.0), 0.98), 0);
    if (!p) return 1;
    Epi e;
    e.mode = EPI_COMBINE;
    e.aux0 = in;
    e.wc = high_gain;
    e.w[0] = 1.0 - high_gain;
    e.clip = 1;
    Pro none;
    return st_filtfilt_combine(c, g, p, in, out, e, none);
}

int mm_dev_apply_reverb(mm_ctx* c, const mm_geom* g, const float* in, float* out, int reverb_type, double decay_sec, double mix,
                        int use_ms, double mix_mid, double mix_side) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_reverb(c, g, in, out, reverb_type, decay_sec, mix, use_ms, mix_mid, mix_side);
}

int mm_dev_spectral_envelope(mm_ctx* c, const mm_geom* g, const float* in, float* env_dev) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_spectral_envelope(c, g, in, env_dev);
}

int mm_dev_fft_resample(mm_ctx* c, const mm_geom* gin, const float* in, const mm_geom* gout, float* out) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(gin));
    MM_TRY(check_geom(gout));
    if (in == out) { set_error("mm_dev_fft_resample: not in place"); return 2; }
    return st_fft_resample(c, gin, in, gout, out);
}

int mm_dev_apply_spectral_denoise(mm_ctx* c, const mm_geom* g, const float* in, float* out, double strength, double noise_percentile) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (!(noise_percentile >= 0.0 && noise_percentile <= 100.0)) { set_error("Percentiles must be in the range [0, 100]"); return 2; }
    strength = std::min(1.0, std::max(0.0, strength));
    if (strength < 0.01) {                       // bypass (pipeline.py:1489-1490)
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    if (in != out) return st_spectral_denoise(c, g, in, out, strength, noise_percentile);
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    MM_TRY(st_spectral_denoise(c, g, in, B.T[1], strength, noise_percentile));
    MM_CUDA(cudaMemcpyAsync(out, B.T[1], batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int mm_dev_fir_same(mm_ctx* c, const mm_geom* g, const float* in, float* out, const float* taps_host, int ntaps, int clip) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (!taps_host) { set_error("mm_dev_fir_same: taps is null"); return 1; }
    float* taps;
    MM_TRY(arena(c, SL_ROWMAP, (size_t)ntaps, &taps));
    MM_CUDA(cudaMemcpyAsync(taps, taps_host, (size_t)ntaps * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    MM_CUDA(cudaStreamSynchronize(c->stream));
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    if (in != out) return st_fir_same(c, g, in, out, taps, ntaps, clip);
    MM_TRY(st_fir_same(c, g, in, B.T[1], taps, ntaps, clip));
    MM_CUDA(cudaMemcpyAsync(out, B.T[1], batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int mm_dev_apply_stereo_imager_4band(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* band_widths,
                                     const double* crossovers_hz) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (!band_widths) { set_error("mm_dev_apply_stereo_imager_4band: band_widths is null"); return 1; }
    if (g->channels != 2) {   // pipeline.py:1355-1356
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    return st_imager4(c, g, in, out, band_widths, crossovers_hz);
}

int mm_dev_apply_stereoize(mm_ctx* c, const mm_geom* g, const float* in, float* out, double width, double delay_ms, double mix) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (g->channels != 2) {   // pipeline.py:1355-1356
        if (in != out) MM_CUDA(cudaMemcpyAsync(out, in, batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    return st_haas_imager(c, g, in, out, width, delay_ms, mix);
}

int mm_dev_iir(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* b, const double* a, int ncoef,
               int zero_phase) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if ((ncoef != 3 && ncoef != 5) || !b || !a || a[0] == 0.0) { set_error("mm_dev_iir: ncoef must be 3 or 5 and a[0] != 0"); return 1; }
    Ba ba;
    memset(&ba, 0, sizeof(ba));
    ba.m = ncoef - 1;
    for (int i = 0; i < ncoef; ++i) { ba.b[i] = b[i] / a[0]; ba.a[i] = a[i] / a[0]; }
    const FilterPlan* p = get_plan(c, ba);
    if (!p) return 1;
    Epi e;
    Pro none;
    if (zero_phase) return st_filtfilt_combine(c, g, p, in, out, e, none);
    const FilterPlan* pl[1] = {p};
    const float* i1[1] = {in};
    float* o1[1] = {out};
    if (in == out) {   // the causal sweep reads its odd... no extension here, but tiles must not race: go through E0
        Bufs B;
        MM_TRY(get_bufs(c, g, &B));
        o1[0] = B.E[0];
        MM_TRY(sweep_fwd(c, g, 1, 1, pl, i1, o1, none, 0));
        MM_CUDA(cudaMemcpyAsync(out, B.E[0], batch_floats(g) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    return sweep_fwd(c, g, 1, 1, pl, i1, o1, none, 0);
}

// ---- export -------------------------------------------------------------------------------------
int mm_dev_finalize_clip(mm_ctx* c, const mm_geom* g, const float* in, float* out) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_finalize_clip(c, g, in, out);
}

int mm_dev_last_above(mm_ctx* c, const mm_geom* g, const float* in, double threshold_lin, int64_t* idx_dev) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_last_above(c, g, in, threshold_lin, reinterpret_cast<long long*>(idx_dev));
}

int mm_dev_quantize_pcm24(mm_ctx* c, const mm_geom* g, const float* in, int32_t* out_interleaved) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_quantize_pcm24(c, g, in, out_interleaved);
}

int mm_dev_quantize_int16(mm_ctx* c, const mm_geom* g, const float* in, int16_t* pcm, const float* noise, uint64_t seed) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    QuantArgs Q;
    Q.in = in; Q.n = g->n; Q.stride = g->stride; Q.tracks = g->tracks; Q.channels = g->channels;
    Q.pcm = pcm; Q.noise = noise; Q.seed = seed; Q.track_base = g->track_base;
    Q.noise_planar = nullptr; Q.noise_scale = 1.f;
    return run_quantize(c, Q);
}

// _dither_noise_ns_e / _dither_noise_ns_itu (pipeline.py:835-877) + the quantiser: white = 2 rand - 1 (float32), shaped by a
// causal IIR from zero state -- ns_e: y[n] = x[n] - x[n-1] + 0.99 y[n-1]; ns_itu: lfilter([1,-2,1], [1,-1.96,0.9604]) --
// scaled by 0.9 in float32, added to samples * 32767 in float64.  The shaping filter is one forward sweep (float64 state).
int mm_dev_quantize_int16_shaped(mm_ctx* c, const mm_geom* g, const float* in, int16_t* pcm, const float* uniform, uint64_t seed,
                                 int shape) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    if (shape != 1 && shape != 2) { set_error("mm_dev_quantize_int16_shaped: shape must be 1 (ns_e) or 2 (ns_itu)"); return 1; }
    if (g->n < (shape == 1 ? 4 : 8)) {      // pipeline.py:840-841 / :861-862: TPDF for very short buffers
        set_error("mm_dev_quantize_int16_shaped: buffer too short for noise shaping (the reference falls back to TPDF: use mm_dev_quantize_int16)");
        return 1;
    }
    Bufs B;
    MM_TRY(get_bufs(c, g, &B));
    WhiteArgs W;
    W.uniform = uniform; W.out = B.T[0]; W.n = g->n; W.stride = g->stride; W.tracks = g->tracks; W.channels = g->channels;
    W.seed = seed; W.track_base = g->track_base;
    MM_TRY(run_white_noise(c, W));
    Ba ba;
    ba.m = 2;
    if (shape == 1) { ba.b[0] = 1.0; ba.b[1] = -1.0; ba.b[2] = 0.0; ba.a[0] = 1.0; ba.a[1] = -0.99; ba.a[2] = 0.0; }
    else { ba.b[0] = 1.0; ba.b[1] = -2.0; ba.b[2] = 1.0; ba.a[0] = 1.0; ba.a[1] = -1.96; ba.a[2] = 0.9604; }
    const FilterPlan* p = get_plan_mode(c, ba, kDf2tF64);
    if (!p) return 1;
    const FilterPlan* pl[1] = {p};
    const float* i1[1] = {B.T[0]};
    float* o1[1] = {B.T[1]};
    Pro none;
    MM_TRY(sweep_fwd(c, g, 1, 1, pl, i1, o1, none, 0));
    QuantArgs Q;
    Q.in = in; Q.n = g->n; Q.stride = g->stride; Q.tracks = g->tracks; Q.channels = g->channels;
    Q.pcm = pcm; Q.noise = nullptr; Q.seed = seed; Q.track_base = g->track_base;
    Q.noise_planar = B.T[1]; Q.noise_scale = 0.9f;
    return run_quantize(c, Q);
}

// ---- analyzers ----------------------------------------------------------------------------------
int mm_dev_true_peak(mm_ctx* c, const mm_geom* g, const float* in, double* tp) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_true_peak(c, g, in, tp);
}
int mm_dev_true_peak_correlation(mm_ctx* c, const mm_geom* g, const float* in, double* tp, double* corr, double* peak) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_true_peak_corr(c, g, in, tp, corr, peak);
}
int mm_dev_spectrum_bars(mm_ctx* c, const mm_geom* g, const float* in, int view, double* bars) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_spectrum_bars(c, g, in, view, bars);
}
int mm_dev_stereo_correlation(mm_ctx* c, const mm_geom* g, const float* in, double* corr, double* peak) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return st_correlation(c, g, in, corr, peak);
}

// mastering_trace.signal_metrics on the device (trace hook): out[tracks][3] = peak over finite samples,
// non-finite count, infinity count
int mm_dev_signal_metrics(mm_ctx* c, const mm_geom* g, const float* in, double* out3) {
    MM_API_BEGIN(c);
    MM_TRY(check_geom(g));
    return run_signal_metrics(c, g, in, out3);
}

// ---- chains ---------------------------------------------------------------------------------------
int mm_dev_master(mm_ctx* c, const mm_geom* g, int chain, const mm_style* styles, const float* in, float* out,
                  int16_t* pcm, const float* noise, uint64_t seed, mm_track_stats* stats_dev, uint32_t flags) {
    MM_API_BEGIN(c);
    return master_lanes(c, g, chain, styles, in, out, pcm, noise, seed, stats_dev, flags);
}

int64_t mm_slice_margin(int32_t sr) {
    // every sweep's start-up (odd extension / zero state at a cut instead of the true neighbourhood) dies within a
    // few time constants of its slowest pole; the chain strings ~16 low-cut-off sweeps per direction together
    // (edm, 30-90 Hz band: ~3 tiles each at 96 kHz).  128 tiles at 96 kHz, scaled with the rate.
    const long long tiles = std::max<long long>(64, (128LL * std::max(sr, 1) + 95999) / 96000);
    return tiles * kL;
}

int mm_dev_master_slice(mm_ctx* c, const mm_geom* g, int chain, const mm_style* style, const float* in, float* out,
                        int16_t* pcm, const float* noise, uint64_t seed, mm_track_stats* stats_dev, uint32_t flags,
                        const mm_slice* slice) {
    MM_API_BEGIN(c);
    if (!slice) return master_lanes(c, g, chain, style, in, out, pcm, noise, seed, stats_dev, flags);
    MM_TRY(check_geom(g));
    if (g->tracks != 1) { set_error("mm_dev_master_slice: one track (file) per call"); return 1; }
    if (slice->own_lo < 0 || slice->own_hi > g->n || slice->own_lo >= slice->own_hi || (slice->own_lo & 3) || slice->global_off < 0 || (slice->global_off & 1) ||
        slice->global_off + g->n > slice->global_n) {
        set_error("mm_dev_master_slice: bad slice (own frames must lie inside the slice, own_lo a multiple of 4, slice inside the file)");
        return 1;
    }
    c->slice = slice;
    const int rc = master_impl(c, g, chain, style, in, out, pcm, noise, seed, stats_dev, flags);
    c->slice = nullptr;
    return rc;
}

// Host-buffer entry points.  The call is cut into CHUNKS of equally-shaped tracks that flow through a three-stage pipeline --
// host->device copy (copy stream), mastering chain (context stream), device->host copy (second copy stream) -- so
// that on a PCIe-attached GPU the call costs about max(copy in, compute, copy out) instead of their sum.  Staging
// buffers are double-buffered per direction; the chain's own workspace is sized for the largest chunk.  Results do not
// depend on the chunking: the dither counter is keyed by the track's index in the whole call (or its explicit id).
// A chunk carries its own geometry, and every track its own host pointers: the equal-shape entry points hand in one contiguous
// buffer (consecutive tracks merge into one cudaMemcpyAsync), mm_master_host_jobs a list of uploads of different shapes.
struct HostTrack {
    const void* in;           // float32 or int16 interleaved frames of this track
    float* out_f32;           // may be null
    int16_t* out_pcm;         // may be null
    const float* noise;       // may be null
};
struct HostChunk { int t0, tn; int64_t n; int32_t channels, sr; };      // tracks [t0, t0 + tn) of the plan

static int master_host_plan(mm_ctx* c, int chain, const std::vector<HostChunk>& chunks, const std::vector<HostTrack>& trk, bool pcm_in,
                            const mm_style* styles, uint64_t seed, mm_track_stats* stats_host, uint32_t flags, bool stage_in = false) {
    const int nchunks = (int)chunks.size(), tracks = (int)trk.size();
    if (nchunks == 0) return 0;
    const size_t in_elem = pcm_in ? sizeof(int16_t) : sizeof(float);
    if (!c->h2d_stream) MM_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    if (!c->d2h_stream) MM_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    // copy-out on one or two streams (two copies in flight keep the link's read side busier while the chain loads HBM)
    static const int n_d2h = [] { const char* e = getenv("MM_D2H_STREAMS"); const int v = e ? atoi(e) : 0; return v == 2 ? 2 : (v == 1 ? 1 : MM_D2H_STREAMS_DEFAULT); }();
    if (n_d2h > 1 && !c->d2h_stream2) MM_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream2, cudaStreamNonBlocking));
    auto d2h_of = [&](int k) { return (n_d2h > 1 && (k & 1)) ? c->d2h_stream2 : c->d2h_stream; };
    bool any_noise = false, any_f32 = false, any_pcm = false;
    for (const HostTrack& t : trk) { any_noise |= t.noise != nullptr; any_f32 |= t.out_f32 != nullptr; any_pcm |= t.out_pcm != nullptr; }
    size_t cframes = 0, pfloats = 0;                               // staging sizes: the largest chunk
    for (const HostChunk& ch : chunks) {
        mm_geom gc;
        gc.n = ch.n; gc.stride = mm_row_stride(ch.n); gc.tracks = ch.tn; gc.channels = ch.channels; gc.sr = ch.sr; gc.track_base = 0;
        MM_TRY(check_geom(&gc));
        cframes = std::max(cframes, (size_t)ch.tn * (size_t)ch.n * ch.channels);
        pfloats = std::max(pfloats, batch_floats(&gc));
    }
    // lanes: chunk k's chain runs on lane k % L (own stream, own planar buffer and scratch), so that consecutive chunks' chains
    // overlap; the device staging rings are 2 L deep so that a lane never waits for its previous chunk's copy-out
    const int L = c->timing ? 1 : std::max(1, std::min(lanes_wanted(c, 4), nchunks));
    const int R = std::min(2 * L, nchunks);
    std::vector<mm_ctx*> lane((size_t)L);
    std::vector<float*> pls((size_t)L, nullptr);
    for (int l = 0; l < L; ++l) MM_TRY(get_lane(c, l, &lane[l]));
    float *il_ring = nullptr, *ol_ring = nullptr, *nz_ring = nullptr;
    int16_t* pcm_ring = nullptr;
    mm_track_stats* st = nullptr;
    MM_TRY(arena(c, SL_STAGE_IL, cframes * R, &il_ring));
    for (int l = 0; l < L; ++l) MM_TRY(arena(lane[l], SL_STAGE_PL, pfloats, &pls[l]));
    if (any_pcm) MM_TRY(arena(c, SL_STAGE_PCM, cframes * R, &pcm_ring));
    if (any_noise) MM_TRY(arena(c, SL_STAGE_NOISE, cframes * R, &nz_ring));
    if (any_f32) MM_TRY(arena(c, SL_STAGE_OL, cframes * R, &ol_ring));
    if (stats_host) MM_TRY(arena(c, SL_STATS, (size_t)tracks, &st));
    auto il_of = [&](int k) { return il_ring + (size_t)(k % R) * cframes; };
    auto ol_of = [&](int k) { return ol_ring ? ol_ring + (size_t)(k % R) * cframes : nullptr; };
    auto nz_of = [&](int k) { return nz_ring ? nz_ring + (size_t)(k % R) * cframes : nullptr; };
    auto pcm_of = [&](int k) { return pcm_ring ? pcm_ring + (size_t)(k % R) * cframes : nullptr; };
    // pageable inputs: a cudaMemcpyAsync from pageable memory runs at a few GB/s and holds the calling thread; instead the thread
    // copies chunk k + 1 into a pinned ring slot (while the device works on chunk k) and the slot crosses PCIe at link speed
    char *hpin[2] = {nullptr, nullptr}, *hpout[2] = {nullptr, nullptr};       // host_pin 0, 1: inputs; 2, 3: outputs
    const size_t out_f32_off = any_pcm ? cframes * sizeof(int16_t) : 0;         // an output slot holds the chunk's PCM_16, then its float32
    const size_t out_bytes = out_f32_off + (any_f32 ? cframes * sizeof(float) : 0);
    if (stage_in) {
        MM_TRY(pin_in_wait(c));
        for (int i = 0; i < 4; ++i) {
            if (nchunks == 1 && (i & 1)) continue;
            Slot& hp = c->host_pin[i];
            const size_t need = i < 2 ? cframes * in_elem : out_bytes;
            if (hp.cap < need) {
                if (hp.p) cudaFreeHost(hp.p);
                hp.p = nullptr; hp.cap = 0;
                MM_CUDA(cudaHostAlloc(&hp.p, need, cudaHostAllocDefault));
                hp.cap = need;
            }
            (i < 2 ? hpin[i] : hpout[i - 2]) = reinterpret_cast<char*>(hp.p);
        }
    }

    // MM_HOST_TIMELINE=1: the pipeline's events carry timestamps and the call prints when each chunk's copy-in, layout, chain and
    // copy-out ended (ms after the call's start) -- the tool that shows which stage a chunk waited for
    static const bool timeline = [] { const char* e = getenv("MM_HOST_TIMELINE"); return e && atoi(e) != 0; }();
    const unsigned ev_flags = timeline ? cudaEventDefault : cudaEventDisableTiming;
    std::vector<cudaEvent_t> ev_in(nchunks), ev_deint(nchunks), ev_done(nchunks), ev_out(nchunks);
    for (int k = 0; k < nchunks; ++k) {
        MM_CUDA(cudaEventCreateWithFlags(&ev_in[k], ev_flags));
        MM_CUDA(cudaEventCreateWithFlags(&ev_deint[k], ev_flags));
        MM_CUDA(cudaEventCreateWithFlags(&ev_done[k], ev_flags));
        MM_CUDA(cudaEventCreateWithFlags(&ev_out[k], ev_flags));
    }
    int rc = 0;
    // everything queued earlier on the context stream (the caller's own work, buffer growth) precedes the copies
    cudaEvent_t ev_start;
    MM_CUDA(cudaEventCreateWithFlags(&ev_start, ev_flags));
    MM_CUDA(cudaEventRecord(ev_start, c->stream));
    MM_CUDA(cudaStreamWaitEvent(c->h2d_stream, ev_start, 0));
    MM_CUDA(cudaStreamWaitEvent(c->d2h_stream, ev_start, 0));
    if (n_d2h > 1) MM_CUDA(cudaStreamWaitEvent(c->d2h_stream2, ev_start, 0));
    MM_TRY(lanes_fork(c, L));
    // one cudaMemcpyAsync per run of tracks whose host buffers lie back to back (the whole chunk, for the equal-shape entry points)
    auto copy_runs = [&](const HostChunk& ch, size_t elem, auto host_of, auto issue) -> int {
        const size_t per_track = (size_t)ch.n * ch.channels;
        for (int a = 0; a < ch.tn;) {
            const char* h0 = reinterpret_cast<const char*>(host_of(trk[ch.t0 + a]));
            if (!h0) { ++a; continue; }
            int b = a + 1;
            while (b < ch.tn && reinterpret_cast<const char*>(host_of(trk[ch.t0 + b])) == h0 + (size_t)(b - a) * per_track * elem) ++b;
            if (issue((size_t)a * per_track, h0, (size_t)(b - a) * per_track * elem) != cudaSuccess) { set_error("mm_master_host: copy failed"); return 1; }
            a = b;
        }
        return 0;
    };
    auto copy_in = [&](int k) -> int {
        const HostChunk& ch = chunks[k];
        if (k >= R) MM_CUDA(cudaStreamWaitEvent(c->h2d_stream, ev_deint[k - R], 0));    // staging slot k % R is free again
        char* dst = reinterpret_cast<char*>(il_of(k));
        if (stage_in) {
            if (k >= 2) MM_CUDA(cudaEventSynchronize(ev_in[k - 2]));                    // the ring slot has crossed the link
            const size_t per_track = (size_t)ch.n * ch.channels * in_elem;
            for (int t = 0; t < ch.tn; ++t) host_par_memcpy(hpin[k & 1] + (size_t)t * per_track, trk[ch.t0 + t].in, per_track);
            MM_CUDA(cudaMemcpyAsync(dst, hpin[k & 1], (size_t)ch.tn * per_track, cudaMemcpyHostToDevice, c->h2d_stream));
        } else
        MM_TRY(copy_runs(ch, in_elem, [](const HostTrack& t) { return t.in; }, [&](size_t off, const char* h, size_t bytes) {
            return cudaMemcpyAsync(dst + off * in_elem, h, bytes, cudaMemcpyHostToDevice, c->h2d_stream); }));
        if (any_noise) {
            if (k >= R) MM_CUDA(cudaStreamWaitEvent(c->h2d_stream, ev_done[k - R], 0));
            float* nd = nz_of(k);
            MM_TRY(copy_runs(ch, sizeof(float), [](const HostTrack& t) { return (const void*)t.noise; }, [&](size_t off, const char* h, size_t bytes) {
                return cudaMemcpyAsync(nd + off, h, bytes, cudaMemcpyHostToDevice, c->h2d_stream); }));
        }
        MM_CUDA(cudaEventRecord(ev_in[k], c->h2d_stream));
        return 0;
    };
    auto drain_out = [&](int k) -> int {                                        // pinned ring slot of chunk k -> the caller's buffers
        const HostChunk& ch = chunks[k];
        if (cudaEventSynchronize(ev_out[k]) != cudaSuccess) { set_error("mm_master_host: copy-out failed"); return 1; }
        const size_t per_track = (size_t)ch.n * ch.channels;
        for (int t = 0; t < ch.tn; ++t) {
            const HostTrack& h = trk[ch.t0 + t];
            if (h.out_pcm) host_par_memcpy(h.out_pcm, hpout[k & 1] + (size_t)t * per_track * sizeof(int16_t), per_track * sizeof(int16_t));
            if (h.out_f32) host_par_memcpy(h.out_f32, hpout[k & 1] + out_f32_off + (size_t)t * per_track * sizeof(float), per_track * sizeof(float));
        }
        return 0;
    };
    rc = copy_in(0);
    for (int k = 0; k < nchunks && rc == 0; ++k) {
        const HostChunk& ch = chunks[k];
        if (k + 1 < nchunks && (rc = copy_in(k + 1)) != 0) break;
        mm_geom gk;
        gk.n = ch.n; gk.stride = mm_row_stride(ch.n); gk.channels = ch.channels; gk.sr = ch.sr;
        gk.tracks = ch.tn;
        gk.track_base = ch.t0;                               // index of the chunk's first track in the call (dither counter)
        bool c_noise = false, c_f32 = false, c_pcm = false;
        for (int t = 0; t < ch.tn; ++t) { const HostTrack& h = trk[ch.t0 + t]; c_noise |= h.noise != nullptr; c_f32 |= h.out_f32 != nullptr; c_pcm |= h.out_pcm != nullptr; }
        mm_ctx* cl = lane[k % L];
        float* pl = pls[k % L];
        if ((rc = cudaStreamWaitEvent(cl->stream, ev_in[k], 0) != cudaSuccess)) { set_error("cudaStreamWaitEvent failed"); break; }
        if (pcm_in) { if ((rc = run_layout_pcm16(cl, &gk, reinterpret_cast<const int16_t*>(il_of(k)), pl)) != 0) break; }
        else if ((rc = mm_dev_deinterleave(cl, &gk, il_of(k), pl)) != 0) break;
        cudaEventRecord(ev_deint[k], cl->stream);
        if (k >= R) cudaStreamWaitEvent(cl->stream, ev_out[k - R], 0);     // pcm / ol staging slot k % R has left the device
        if ((rc = master_impl(cl, &gk, chain, styles + ch.t0, pl, pl, c_pcm ? pcm_of(k) : nullptr, c_noise ? nz_of(k) : nullptr, seed,
                              st ? st + ch.t0 : nullptr, flags)) != 0) break;
        if (c_f32 && (rc = mm_dev_interleave(cl, &gk, pl, ol_of(k))) != 0) break;
        cudaEventRecord(ev_done[k], cl->stream);
        cudaStream_t d2h = d2h_of(k);
        cudaStreamWaitEvent(d2h, ev_done[k], 0);
        if (stage_in) {
            // pageable outputs: the chunk lands in a pinned ring slot; the thread hands it to the caller's buffers one chunk later
            // (a device->pageable cudaMemcpyAsync would hold the thread until chunk k is done, with chunk k + 1's chain not yet queued)
            if (k >= 2 && (rc = drain_out(k - 2)) != 0) break;
            const size_t cnt = (size_t)ch.tn * ch.n * ch.channels;
            if (c_pcm && cudaMemcpyAsync(hpout[k & 1], pcm_of(k), cnt * sizeof(int16_t), cudaMemcpyDeviceToHost, d2h) != cudaSuccess) rc = 1;
            if (c_f32 && cudaMemcpyAsync(hpout[k & 1] + out_f32_off, ol_of(k), cnt * sizeof(float), cudaMemcpyDeviceToHost, d2h) != cudaSuccess) rc = 1;
            if (rc) { set_error("mm_master_host: copy failed"); break; }
        } else {
        if (c_f32) {
            const float* src = ol_of(k);
            rc = copy_runs(ch, sizeof(float), [](const HostTrack& t) { return (const void*)t.out_f32; }, [&](size_t off, const char* h, size_t bytes) {
                return cudaMemcpyAsync(const_cast<char*>(h), src + off, bytes, cudaMemcpyDeviceToHost, d2h); });
        }
        if (rc == 0 && c_pcm) {
            const int16_t* src = pcm_of(k);
            rc = copy_runs(ch, sizeof(int16_t), [](const HostTrack& t) { return (const void*)t.out_pcm; }, [&](size_t off, const char* h, size_t bytes) {
                return cudaMemcpyAsync(const_cast<char*>(h), src + off, bytes, cudaMemcpyDeviceToHost, d2h); });
        }
        }
        cudaEventRecord(ev_out[k], d2h);
    }
    if (stage_in)
        for (int k = std::max(0, nchunks - 2); k < nchunks && rc == 0; ++k) rc = drain_out(k);
    {
        const int rj = lanes_join(c, L);           // the parent's stream (the stats copy, the caller's next call) follows every lane
        if (rc == 0) rc = rj;
    }
    if (rc == 0 && stats_host) {
        if (cudaMemcpyAsync(stats_host, st, (size_t)tracks * sizeof(mm_track_stats), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) {
            set_error("copy of the track stats failed");
            rc = 1;
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(c->h2d_stream), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(c->d2h_stream);
    if (c->d2h_stream2 && e3 == cudaSuccess) e3 = cudaStreamSynchronize(c->d2h_stream2);
    if (timeline && rc == 0 && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess) {
        fprintf(stderr, "[mm timeline] %d chunks on %d lanes (ms after the start: copy-in end | layout end | chain end | copy-out end)\n", nchunks, L);
        for (int k = 0; k < nchunks; ++k) {
            float a = 0, b = 0, d = 0, o = 0;
            cudaEventElapsedTime(&a, ev_start, ev_in[k]); cudaEventElapsedTime(&b, ev_start, ev_deint[k]);
            cudaEventElapsedTime(&d, ev_start, ev_done[k]); cudaEventElapsedTime(&o, ev_start, ev_out[k]);
            fprintf(stderr, "[mm timeline] chunk %3d lane %d tracks %2d: %8.3f | %8.3f | %8.3f | %8.3f\n", k, k % L, chunks[k].tn, a, b, d, o);
        }
    }
    for (int k = 0; k < nchunks; ++k) { cudaEventDestroy(ev_in[k]); cudaEventDestroy(ev_deint[k]); cudaEventDestroy(ev_done[k]); cudaEventDestroy(ev_out[k]); }
    cudaEventDestroy(ev_start);
    if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
        set_error("mm_master_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
        rc = 1;
    }
    return rc;
}

// tracks per chunk: ~256 MB of float32 input.  Measured (64 x 180 s, float32 / PCM_16 in): 2 tracks 140 / 173 k audio-s/s,
// 4: 138 / 189 k, 8: 131 / 188 k, 16: 118 / 168 k -- short pipeline fill against small grids
static int host_chunk_tracks(size_t per_track) {
    int tc = 0;
    if (const char* e = getenv("MM_HOST_CHUNK")) tc = atoi(e);
    if (tc <= 0) tc = (int)std::max<size_t>(1, ((size_t)256 << 20) / std::max<size_t>(per_track * sizeof(float), 1));
    return tc;
}

static int master_host_impl(mm_ctx* c, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr, const mm_style* styles,
                            const float* audio_in, const int16_t* pcm16_in, float* audio_out, int16_t* pcm16_out,
                            const float* noise_host, uint64_t seed, mm_track_stats* stats_host, uint32_t flags) {
    mm_geom g;
    g.n = n; g.stride = mm_row_stride(n); g.tracks = tracks; g.channels = channels; g.sr = sr; g.track_base = 0;
    MM_TRY(check_geom(&g));
    if (!audio_in && !pcm16_in) { set_error("mm_master_host: the input buffer is null"); return 1; }
    const size_t per_track = (size_t)n * channels;                  // interleaved samples of one track
    const int tc = std::min(host_chunk_tracks(per_track), (int)tracks);
    // chunk plan: full chunks of tc tracks, with short chunks at both ends (1, 2, ... tracks) so that the pipeline's fill (the
    // first copy-in before any compute) and drain (the last chunk's chain + copy-out after the last copy-in) cost one track
    // instead of one full chunk.  MM_HOST_RAMP=0 switches the ramps off.
    std::vector<HostChunk> chunks;
    {
        std::vector<int> head, tail;
        int rem = tracks;
        const char* er = getenv("MM_HOST_RAMP");
        const bool ramp = !(er && atoi(er) == 0);
        for (int s = 1; ramp && s < tc && rem >= 2 * s + tc; s *= 2) { head.push_back(s); tail.push_back(s); rem -= 2 * s; }
        std::vector<int> sizes(head);
        while (rem > 0) { const int s = std::min(tc, rem); sizes.push_back(s); rem -= s; }
        for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
        int t0 = 0;
        for (int s : sizes) { chunks.push_back(HostChunk{t0, s, n, channels, sr}); t0 += s; }
    }
    std::vector<HostTrack> trk((size_t)tracks);
    for (int t = 0; t < tracks; ++t) {
        const size_t off = (size_t)t * per_track;
        trk[t].in = pcm16_in ? (const void*)(pcm16_in + off) : (const void*)(audio_in + off);
        trk[t].out_f32 = audio_out ? audio_out + off : nullptr;
        trk[t].out_pcm = pcm16_out ? pcm16_out + off : nullptr;
        trk[t].noise = noise_host ? noise_host + off : nullptr;
    }
    return master_host_plan(c, chain, chunks, trk, pcm16_in != nullptr, styles, seed, stats_host, flags);
}

int mm_master_host(mm_ctx* c, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr, const mm_style* styles,
                   const float* audio_in, float* audio_out, int16_t* pcm16_out, const float* noise_host, uint64_t seed,
                   mm_track_stats* stats_host, uint32_t flags) {
    MM_API_BEGIN(c);
    return master_host_impl(c, chain, tracks, n, channels, sr, styles, audio_in, nullptr, audio_out, pcm16_out, noise_host, seed,
                            stats_host, flags);
}

int mm_master_host_ids(mm_ctx* c, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr, const mm_style* styles,
                       const float* audio_in, const int16_t* pcm16_in, float* audio_out, int16_t* pcm16_out, uint64_t seed,
                       mm_track_stats* stats_host, uint32_t flags, const int32_t* track_ids) {
    MM_API_BEGIN(c);
    c->track_ids_host = track_ids;
    const int rc = master_host_impl(c, chain, tracks, n, channels, sr, styles, audio_in, pcm16_in, audio_out, pcm16_out, nullptr, seed,
                                    stats_host, flags);
    c->track_ids_host = nullptr;
    c->track_ids_dev = nullptr;
    return rc;
}

// Job-level variant (what _run_mastering_job does with a PCM_16 WAV upload, routers/mastering.py:350-441): the data chunk's
// int16 frames go to the device as they are (half the bytes of float32 over PCIe) and are widened there exactly as
// libsndfile does for dtype="float32" (x / 32768, pipeline.py:814-817); the result comes back as PCM_16 again.
int mm_master_host_pcm16(mm_ctx* c, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr, const mm_style* styles,
                         const int16_t* pcm16_in, float* audio_out, int16_t* pcm16_out, uint64_t seed, mm_track_stats* stats_host,
                         uint32_t flags) {
    MM_API_BEGIN(c);
    return master_host_impl(c, chain, tracks, n, channels, sr, styles, nullptr, pcm16_in, audio_out, pcm16_out, nullptr, seed,
                            stats_host, flags);
}

// ---- filter design ----------------------------------------------------------------------------
// A list of uploads of DIFFERENT shapes in one call (what /api/v2/batch receives, routers/mastering.py:855-1037): the jobs are
// ordered by (rate, channels, frames), runs of equal shape become chunks of up to ~256 MB, and the chunks of all shapes flow through
// ONE copy-in / chain / copy-out pipeline -- an upload's copy-in overlaps the previous upload's chain even when every upload has its
// own length (round 1: one blocking call per shape).  Host buffers are the caller's (pinned or pageable), one per job.
int mm_master_host_jobs(mm_ctx* c, int chain, int32_t njobs, mm_host_job* jobs, uint64_t seed, uint32_t flags) {
    MM_API_BEGIN(c);
    if (njobs < 0 || (njobs > 0 && !jobs)) { set_error("mm_master_host_jobs: bad job list"); return 1; }
    if (njobs == 0) return 0;
    bool pcm_in = jobs[0].pcm16_in != nullptr;
    for (int j = 0; j < njobs; ++j) {
        const mm_host_job& b = jobs[j];
        if ((b.audio_in != nullptr) == (b.pcm16_in != nullptr)) { set_error("mm_master_host_jobs: job %d needs exactly one of audio_in / pcm16_in", j); return 1; }
        if ((b.pcm16_in != nullptr) != pcm_in) { set_error("mm_master_host_jobs: float32 and PCM_16 inputs cannot be mixed in one call"); return 1; }
        if (b.n <= 0 || (b.channels != 1 && b.channels != 2) || b.sr <= 0) { set_error("mm_master_host_jobs: job %d has a bad shape", j); return 1; }
    }
    std::vector<int> order((size_t)njobs);
    for (int j = 0; j < njobs; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const mm_host_job &x = jobs[a], &y = jobs[b];
        if (x.sr != y.sr) return x.sr < y.sr;
        if (x.channels != y.channels) return x.channels < y.channels;
        return x.n < y.n;
    });
    std::vector<HostChunk> chunks;
    std::vector<HostTrack> trk((size_t)njobs);
    std::vector<mm_style> styles((size_t)njobs);
    std::vector<int32_t> ids((size_t)njobs);
    std::vector<mm_track_stats> stats((size_t)njobs);
    for (int p = 0; p < njobs; ++p) {
        const mm_host_job& b = jobs[order[p]];
        trk[p].in = pcm_in ? (const void*)b.pcm16_in : (const void*)b.audio_in;
        trk[p].out_f32 = b.audio_out; trk[p].out_pcm = b.pcm16_out; trk[p].noise = nullptr;
        styles[p] = b.style;
        ids[p] = b.dither_id;
        const int tc = host_chunk_tracks((size_t)b.n * b.channels);
        if (!chunks.empty() && chunks.back().n == b.n && chunks.back().channels == b.channels && chunks.back().sr == b.sr && chunks.back().tn < tc)
            chunks.back().tn += 1;
        else
            chunks.push_back(HostChunk{p, 1, b.n, b.channels, b.sr});
    }
    bool pageable = false;
    for (int p = 0; p < njobs && !pageable; ++p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, trk[p].in) != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = at.type == cudaMemoryTypeUnregistered;
    }
    c->track_ids_host = ids.data();
    const int rc = master_host_plan(c, chain, chunks, trk, pcm_in, styles.data(), seed, stats.data(), flags, pageable);
    c->track_ids_host = nullptr;
    c->track_ids_dev = nullptr;
    if (rc == 0)
        for (int p = 0; p < njobs; ++p) jobs[order[p]].stats = stats[p];
    return rc;
}

int mm_design_butter(int order, int btype, const double* wn, double* b, double* a) {
    Ba f;
    memset(&f, 0, sizeof(f));
    if (!wn || !b || !a || btype < 0 || btype > 2) { set_error("mm_design_butter: bad arguments"); return -1; }
    if (!butter(order, (BType)btype, wn, &f)) { set_error("mm_design_butter: unsupported order or critical frequencies"); return -1; }
    for (int i = 0; i <= f.m; ++i) { b[i] = f.b[i]; a[i] = f.a[i]; }
    return f.m + 1;
}

int mm_design_iirpeak(double w0, double q, double* b, double* a, int* kind, double* rmax) {
    Ba f;
    memset(&f, 0, sizeof(f));
    if (!b || !a || !iirpeak(w0, q, &f)) { set_error("mm_design_iirpeak: bad arguments"); return 1; }
    for (int i = 0; i < 3; ++i) { b[i] = f.b[i]; a[i] = f.a[i]; }
    const int k = dyneq_band_kind(f, rmax);
    if (kind) *kind = k;
    return 0;
}

int mm_design_lfilter_zi(const double* b, const double* a, int ncoef, double* zi) {
    if (!b || !a || !zi || ncoef < 2 || ncoef > kMaxOrder + 1 || a[0] == 0.0) { set_error("mm_design_lfilter_zi: bad arguments"); return 1; }
    Ba f;
    memset(&f, 0, sizeof(f));
    f.m = ncoef - 1;
    for (int i = 0; i < ncoef; ++i) { f.b[i] = b[i] / a[0]; f.a[i] = a[i] / a[0]; }
    if (!lfilter_zi(f, zi)) { set_error("mm_design_lfilter_zi: singular system"); return 1; }
    return 0;
}

int mm_design_k_weighting(int stage, double rate, double* b, double* a) {
    if (!b || !a || stage < 0 || stage > 1 || !(rate > 0)) { set_error("mm_design_k_weighting: bad arguments"); return 1; }
    Ba f = k_weighting_stage(stage, rate);
    for (int i = 0; i < 3; ++i) { b[i] = f.b[i]; a[i] = f.a[i]; }
    return 0;
}

// The K-weighting high-pass as the state-variable filter the loudness kernel runs (host-side verification, tests/):
// fqg = (f, q, g); abcd = A (2x2, row-major), B (2), C (2), D of  s[n] = A s[n-1] + B x[n],  y[n] = C s[n-1] + D x[n],  s = (lp, bp).
int mm_design_svf_highpass(const double* b, const double* a, double* fqg, double* abcd) {
    if (!b || !a || !fqg || !abcd || a[0] == 0.0) { set_error("mm_design_svf_highpass: bad arguments"); return 1; }
    Ba f;
    f.m = 2;
    for (int i = 0; i < 3; ++i) { f.b[i] = b[i] / a[0]; f.a[i] = a[i] / a[0]; }
    StateSpace ss;
    if (!svf_highpass_realization(f, &ss, &fqg[0], &fqg[1], &fqg[2])) {
        set_error("mm_design_svf_highpass: not a high-pass with a double zero at z = 1 (b = g [1, -2, 1], 1 + a1 + a2 > 0)");
        return 1;
    }
    for (int i = 0; i < 4; ++i) abcd[i] = (double)ss.A[i];
    for (int i = 0; i < 2; ++i) { abcd[4 + i] = (double)ss.B[i]; abcd[6 + i] = (double)ss.C[i]; }
    abcd[8] = (double)ss.D;
    return 0;
}

// Scan tables of a section, for host-side verification of the tile decomposition (tests/):
// returns the look-back window W; g[kS*m], Pw[5*m*m], Plane[32*m*m], Qpow[(kNW+1)*m*m],
// Mpow[cap_w*m*m] (first min(W, cap_w) powers), Apow[(kS+1)*m*m], zi[m].
int mm_design_scan_tables(const double* b, const double* a, int ncoef, double* g_, double* Pw, double* Plane, double* Qpow,
                          double* Mpow, int cap_w, double* Apow, double* zi, int* S, int* T) {
    if (!b || !a || (ncoef != 3 && ncoef != 5)) { set_error("mm_design_scan_tables: ncoef must be 3 or 5"); return -1; }
    Ba f;
    memset(&f, 0, sizeof(f));
    f.m = ncoef - 1;
    for (int i = 0; i < ncoef; ++i) { f.b[i] = b[i] / a[0]; f.a[i] = a[i] / a[0]; }
    ScanTables t;
    if (!build_scan_tables(f, kS, kT, &t)) { set_error("mm_design_scan_tables: pole too close to the unit circle"); return -1; }
    const int mm2 = f.m * f.m;
    if (g_) std::copy(t.g.begin(), t.g.end(), g_);
    if (Pw) std::copy(t.Pw.begin(), t.Pw.end(), Pw);
    if (Plane) std::copy(t.Plane.begin(), t.Plane.end(), Plane);
    if (Qpow) std::copy(t.Qpow.begin(), t.Qpow.end(), Qpow);
    if (Mpow) std::copy(t.Mpow.begin(), t.Mpow.begin() + (size_t)std::min(t.W, cap_w) * mm2, Mpow);
    if (Apow) std::copy(t.Apow.begin(), t.Apow.end(), Apow);
    if (zi) for (int i = 0; i < f.m; ++i) zi[i] = t.zi[i];
    if (S) *S = kS;
    if (T) *T = kT;
    return t.W;
}

// Same tables for a chosen realization (0 = float64 DF2T, 1 = balanced coordinates of the float32 pass 2) plus
// the realization itself: ss = [A (m*m), B (m), C (m), D, ||A||_2]  (design.h).
int mm_design_scan_tables2(const double* b, const double* a, int ncoef, int mode, double* g_, double* Pw, double* Plane,
                           double* Qpow, double* Mpow, int cap_w, double* Apow, double* zi, double* ss) {
    if (!b || !a || (ncoef != 3 && ncoef != 5) || mode < 0 || mode > 1) { set_error("mm_design_scan_tables2: bad arguments"); return -1; }
    Ba f;
    f.m = ncoef - 1;
    for (int i = 0; i < ncoef; ++i) { f.b[i] = b[i] / a[0]; f.a[i] = a[i] / a[0]; }
    ScanTables t;
    const bool ok = mode == kBalancedF32 ? build_scan_tables_balanced(f, kS, kT, &t) : build_scan_tables(f, kS, kT, &t);
    if (!ok) { set_error("mm_design_scan_tables2: section cannot be tabulated in this realization"); return -1; }
    const int m = f.m, mm2 = m * m;
    if (g_) std::copy(t.g.begin(), t.g.end(), g_);
    if (Pw) std::copy(t.Pw.begin(), t.Pw.end(), Pw);
    if (Plane) std::copy(t.Plane.begin(), t.Plane.end(), Plane);
    if (Qpow) std::copy(t.Qpow.begin(), t.Qpow.end(), Qpow);
    if (Mpow) std::copy(t.Mpow.begin(), t.Mpow.begin() + (size_t)std::min(t.W, cap_w) * mm2, Mpow);
    if (Apow) std::copy(t.Apow.begin(), t.Apow.end(), Apow);
    if (zi) for (int i = 0; i < m; ++i) zi[i] = t.zi[i];
    if (ss) {
        for (int i = 0; i < mm2; ++i) ss[i] = t.A[i];
        for (int i = 0; i < m; ++i) { ss[mm2 + i] = t.B[i]; ss[mm2 + m + i] = t.C[i]; }
        ss[mm2 + 2 * m] = t.D;
        ss[mm2 + 2 * m + 1] = mode == kBalancedF32 ? t.norm2 : balanced_norm(f);
    }
    return t.W;
}

}  // extern "C"
