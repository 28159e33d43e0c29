/* mm_b200.h -- C ABI of the B200-native Magic Master mastering hot path.
 *
 * The reference (denisok-ai/audio-mastering-web) has no FFI of its own: its hot path is the
 * Python module backend/app/pipeline.py, called in-process by the FastAPI routes.  This header
 * is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); every entry point
 * names the reference function it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on failure; mm_last_error() (thread-local)
 *    holds the message.  There is NO CPU fallback: without a CUDA device mm_ctx_create fails.
 *  - "dev" entry points take DEVICE pointers to planar batches:
 *        row r = track * channels + channel,   row pointer = base + r * stride,
 *        sample i of a row lives at float offset MM_LEAD + i,
 *        stride = mm_row_stride(n)  (multiple of 32 floats, >= MM_LEAD + n + 32).
 *    The lead-in/lead-out floats of a row are scratch (their content is ignored and may be
 *    overwritten).  All work is enqueued on the context's stream; nothing synchronises unless
 *    stated.  out may alias in for every stage function.
 *  - "host" entry points take HOST pointers in the reference's own layout: float32, C-order,
 *    (n, channels) interleaved -- what soundfile.read(always_2d=True) returns -- and include
 *    the host<->device copies.
 */
#ifndef MM_B200_H
#define MM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_LEAD 32
#define MM_ABI_VERSION 1

typedef struct mm_ctx mm_ctx;

/* Geometry of a device-resident batch: every track has n frames, `channels` channels, rate sr. */
typedef struct mm_geom {
    int64_t n;        /* frames per track */
    int64_t stride;   /* floats per row, = mm_row_stride(n) */
    int32_t tracks;
    int32_t channels; /* 1 or 2 */
    int32_t sr;       /* sample rate, Hz */
    int32_t track_base; /* index of track 0 inside a larger logical batch (dither counter only); normally 0 */
} mm_geom;

/* STYLE_CONFIGS row (backend/app/pipeline.py:69-86) + target; one per track. */
typedef struct mm_style {
    double target_lufs;
    double eq_gain_db[5];   /* sub, bass, mids, presence, air */
    double exciter_db;
    double imager_width;
    double parallel_mix;    /* v1 only (pipeline.py:1856) */
} mm_style;

/* Per-track results gathered by the chain (device array of these, or host copy). */
typedef struct mm_track_stats {
    double lufs_in;        /* measure_lufs(input)  (routers/mastering.py:491); NaN if too short */
    double lufs_mid;       /* loudness seen by normalize_lufs (pipeline.py:648) */
    double lufs_out;       /* measure_lufs(output) (routers/mastering.py:587) */
    double gain_db;        /* gain normalize_lufs applied, clipped to +-20 dB */
    double peak_in;        /* max|x - mean| seen by the input peak guard */
    double peak_out;       /* max|x| seen by the output peak guard */
    double mean[2];        /* DC offset removed per channel */
    double nonfinite;      /* count of non-finite samples in the final buffer (trace hook) */
} mm_track_stats;

/* ---- context ------------------------------------------------------------------------------ */
int         mm_abi_version(void);
const char* mm_last_error(void);
int64_t     mm_row_stride(int64_t n);
/* stream: a cudaStream_t (as void*) owned by the caller, or NULL for a private stream. */
int  mm_ctx_create(int device, void* stream, mm_ctx** out);
void mm_ctx_destroy(mm_ctx* ctx);
int  mm_ctx_sync(mm_ctx* ctx);
/* Free the context's device scratch (arena slots, FFT plan spectra); they grow on demand and otherwise stay at their high-water
 * mark -- a service calls this after an unusually large job (a 2-hour file, a 64-track denoise).  Filter tables are kept. */
int  mm_ctx_release_workspace(mm_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t mm_ctx_launch_count(mm_ctx* ctx);
/* Device-time accounting: when enabled, every kernel is bracketed by CUDA events on the
 * context's stream; mm_ctx_kernel_times fills up to `cap` entries (name, total ms, launches, channel-samples the launches
 * visited where the launcher knows it -- row lists and track runs of a mixed batch -- else 0). */
int  mm_ctx_timing(mm_ctx* ctx, int enable);
typedef struct mm_ktime { char name[48]; double ms; int64_t launches; double samples; } mm_ktime;
int  mm_ctx_kernel_times(mm_ctx* ctx, mm_ktime* out, int cap, int* count);

/* ---- layout helpers -------------------------------------------------------------------------*/
/* interleaved (n, ch) device buffer <-> planar rows */
int mm_dev_deinterleave(mm_ctx*, const mm_geom*, const float* interleaved /*[tracks][n][ch]*/, float* planar);
int mm_dev_interleave(mm_ctx*, const mm_geom*, const float* planar, float* interleaved);

/* ---- stage functions (device batches); each mirrors one pipeline.py function ----------------*/
/* remove_dc_offset                     backend/app/pipeline.py:134-138 */
int mm_dev_remove_dc_offset(mm_ctx*, const mm_geom*, const float* in, float* out);
/* remove_intersample_peaks             backend/app/pipeline.py:141-149 */
int mm_dev_remove_intersample_peaks(mm_ctx*, const mm_geom*, const float* in, float* out, double headroom_db);
/* apply_output_edge_fade_in            backend/app/pipeline.py:152-167 */
int mm_dev_fade_in(mm_ctx*, const mm_geom*, const float* in, float* out, double fade_ms);
/* BaseModule.process blend                backend/app/modules/base.py:44-46: out = dry*(1-amount) + processed*amount */
int mm_dev_blend(mm_ctx*, const mm_geom*, const float* dry, float* out, const float* processed, double amount);
/* apply_target_curve (IIR, minimum)    backend/app/pipeline.py:238-273; eq_ms -> :248-255 */
int mm_dev_apply_target_curve(mm_ctx*, const mm_geom*, const float* in, float* out, int eq_ms);
/* apply_target_curve(phase_mode="linear_phase") = apply_target_curve_linear_phase, backend/app/pipeline.py:187-235 */
int mm_dev_apply_target_curve_linear_phase(mm_ctx*, const mm_geom*, const float* in, float* out, int eq_ms);
/* apply_deesser                        backend/app/pipeline.py:1200-1264 */
int mm_dev_apply_deesser(mm_ctx*, const mm_geom*, const float* in, float* out,
                         double threshold_db, double ratio, double freq_lo, double freq_hi,
                         double attack_ms, double release_ms);
/* apply_dynamics (numpy compressor branch) backend/app/pipeline.py:610-641, :414-481 */
int mm_dev_apply_dynamics(mm_ctx*, const mm_geom*, const float* in, float* out, double knee_db,
                          const double* crossovers_hz /*3 or NULL*/, const double* band_ratios /*4 or NULL*/,
                          double max_upward_boost_db);
/* apply_multiband_dynamics alone (backend/app/pipeline.py:414-481, numpy branch): split, per-band soft knee + limiter + gain,
 * sum -- apply_dynamics without its maximizer / limiter tail.  Arguments as mm_dev_apply_dynamics */
int mm_dev_apply_multiband_dynamics(mm_ctx*, const mm_geom*, const float* in, float* out, double knee_db,
                                    const double* crossovers_hz, const double* band_ratios, double max_upward_boost_db);
/* apply_maximizer_lookahead (pipeline.py:548-573); not in place */
int mm_dev_apply_maximizer_lookahead(mm_ctx*, const mm_geom*, const float* in, float* out, double lookahead_ms);
/* apply_maximizer                      backend/app/pipeline.py:484-492 */
int mm_dev_apply_maximizer(mm_ctx*, const mm_geom*, const float* in, float* out);
/* apply_parallel_compression           backend/app/pipeline.py:1771-1797 */
int mm_dev_apply_parallel_compression(mm_ctx*, const mm_geom*, const float* in, float* out,
                                      double mix, double ratio, double threshold_db);
/* measure_lufs                         backend/app/pipeline.py:658-664 (pyloudnorm BS.1770-4)
 * lufs_dev: device array [tracks] of double. */
int mm_dev_measure_lufs(mm_ctx*, const mm_geom*, const float* in, double* lufs_dev);
/* normalize_lufs                       backend/app/pipeline.py:644-655; target per track (host array) */
int mm_dev_normalize_lufs(mm_ctx*, const mm_geom*, const float* in, float* out, const double* target_lufs_host);
/* apply_final_spectral_balance         backend/app/pipeline.py:576-607 */
int mm_dev_apply_final_spectral_balance(mm_ctx*, const mm_geom*, const float* in, float* out);
/* apply_style_eq                       backend/app/pipeline.py:1401-1434; one gain set for the batch */
int mm_dev_apply_style_eq(mm_ctx*, const mm_geom*, const float* in, float* out, const double* eq_gain_db /*5*/);
/* apply_harmonic_exciter (warm, oversample 1) backend/app/pipeline.py:1267-1326 */
int mm_dev_apply_harmonic_exciter(mm_ctx*, const mm_geom*, const float* in, float* out, double exciter_db, int mode);
/* apply_stereo_imager (width mode)     backend/app/pipeline.py:1339-1398 */
int mm_dev_apply_stereo_imager(mm_ctx*, const mm_geom*, const float* in, float* out, double width);
/* apply_rumble_filter                  backend/app/pipeline.py:1449-1469 */
int mm_dev_apply_rumble_filter(mm_ctx*, const mm_geom*, const float* in, float* out, double cutoff_hz);
/* second-wave stages built on the same two primitives */
/* apply_transient_designer             backend/app/pipeline.py:1736-1768 */
int mm_dev_apply_transient_designer(mm_ctx*, const mm_geom*, const float* in, float* out, double attack_gain, double sustain_gain);
/* apply_maximizer_transient_aware      backend/app/pipeline.py:521-545 */
int mm_dev_apply_maximizer_transient_aware(mm_ctx*, const mm_geom*, const float* in, float* out, double sensitivity);
/* apply_high_freq_trim                 backend/app/pipeline.py:1705-1733 */
int mm_dev_apply_high_freq_trim(mm_ctx*, const mm_geom*, const float* in, float* out, double crossover_hz, double high_gain);
/* apply_stereo_imager with stereoize_delay_ms > 0 (single-band width + Haas cross-delay), :1339-1398; in != out;
 * width = NaN skips the mid/side step (the pair already went through the 4-band mode) */
int mm_dev_apply_stereoize(mm_ctx*, const mm_geom*, const float* in, float* out, double width, double delay_ms, double mix);
/* apply_dynamics / apply_multiband_dynamics (bands_only != 0) with an explicit compressor mode (MM_COMPRESSOR_*) */
int mm_dev_apply_dynamics_mode(mm_ctx*, const mm_geom*, const float* in, float* out, double knee_db, const double* crossovers_hz /*3 or NULL*/,
                               const double* band_ratios /*4 or NULL*/, double max_upward_boost_db, int bands_only, int compressor);
/* apply_stereo_imager with band_widths (4-band mode: _split_bands + per-band width + merge), :1360-1386;
 * crossovers_hz NULL -> MULTIBAND_CROSSOVERS_HZ (214, 3500, 10000) */
int mm_dev_apply_stereo_imager_4band(mm_ctx*, const mm_geom*, const float* in, float* out, const double* band_widths /*4*/,
                                     const double* crossovers_hz /*3 or NULL*/);
/* apply_reverb (Schroeder comb + allpass), backend/app/pipeline.py:1055-1176.  reverb_type: 0 plate, 1 room, 2 hall,
 * 3 theater, 4 cathedral; decay_sec <= 0 -> the preset's; use_ms != 0 (stereo only): separate mixes on mid and side. */
int mm_dev_apply_reverb(mm_ctx*, const mm_geom*, const float* in, float* out, int reverb_type, double decay_sec, double mix,
                        int use_ms, double mix_mid, double mix_side);
/* scipy.signal.resample(x, gout->n) of every row of the input batch (whole-signal FFT resampling, any pair of lengths up
 * to 2^27 combined points): resample_audio (backend/app/pipeline.py:920-936), the oversampled exciter (:1294-1320), a
 * reference track at another rate (:1581-1584).  gout has the same tracks / channels; gin->n != gout->n; not in place */
int mm_dev_fft_resample(mm_ctx*, const mm_geom* gin, const float* in, const mm_geom* gout, float* out);
/* apply_dynamic_eq (backend/app/pipeline.py:1628-1700): params[nbands][7] = {w0, bw, threshold_db, ratio, attack_ms,
 * release_ms, max_cut_db}, w0 / bw as the reference clips them (:1657-1658) and hands them to scipy.signal.iirpeak(w0, bw).
 * The reference passes a bandwidth where scipy expects Q, so most of its DEFAULT bands (DYNAMIC_EQ_MASTERING_BANDS,
 * :1616-1625) are unstable or degenerate sections; what `_safe_filtfilt` + nan_to_num (:36-52, :1677) make of them is
 * reproduced per band class (csrc/deesser.cu st_dynamic_eq): an unstable band whose forward pass must overflow is zeroed by
 * the reference and therefore the identity; a degenerate b0 [1,0,-1] / [1,~0,~-1] section returns b0 x (lfilter fallback) or
 * b0^2 (x - last sample of the odd extension) (filtfilt with poles at +-1). */
int mm_dev_apply_dynamic_eq(mm_ctx*, const mm_geom*, const float* in, float* out, int nbands, const double* params);
/* Same, reporting each band's class into classes[nbands] (host, may be NULL).  flags & MM_DYNEQ_STRICT: an unstable band
 * whose overflow is not certain (class MM_DYNEQ_SKIPPED: passed through by default) is refused with return code 3. */
#define MM_DYNEQ_STRICT 1u
enum { MM_DYNEQ_STABLE = 0, MM_DYNEQ_OVERFLOW = 1, MM_DYNEQ_LFILTER = 2, MM_DYNEQ_MARGINAL = 3, MM_DYNEQ_SKIPPED = 4 };
int mm_dev_apply_dynamic_eq2(mm_ctx*, const mm_geom*, const float* in, float* out, int nbands, const double* params,
                             uint32_t flags, int32_t* classes);
/* apply_spectral_denoise (backend/app/pipeline.py:1472-1524): 2048/512 STFT (scipy.signal.stft conventions), per-bin
 * percentile noise floor over the frames capped by 0.85 x the median, Wiener gain clipped to [0.25, 1], inverse STFT, clip.
 * n >= 2048 (the reference's scipy call raises below that); strength < 0.01 is a bypass */
int mm_dev_apply_spectral_denoise(mm_ctx*, const mm_geom*, const float* in, float* out, double strength, double noise_percentile);
/* compute_spectral_envelope (backend/app/pipeline.py:1527-1551): per track, RMS over 8192-sample Hann frames (hop 2048) of
 * |rfft| of the channel mean -> env_dev[tracks][4097] float32 (device). */
int mm_dev_spectral_envelope(mm_ctx*, const mm_geom*, const float* in, float* env_dev);
/* scipy.signal.fftconvolve(x, taps, mode="same") on every row as a direct FIR (float32 products, float64 carries);
 * ntaps a multiple of 64; taps on the HOST; clip != 0 clips to +-1 (apply_reference_match, :1600-1606) */
int mm_dev_fir_same(mm_ctx*, const mm_geom*, const float* in, float* out, const float* taps_host, int ntaps, int clip);
/* generic zero-phase / causal IIR on every row: scipy filtfilt / lfilter semantics of
 * _safe_filtfilt (backend/app/pipeline.py:36-52). nb == na in {3, 5}; zero_phase 0 -> lfilter. */
int mm_dev_iir(mm_ctx*, const mm_geom*, const float* in, float* out,
               const double* b, const double* a, int ncoef, int zero_phase);

/* ---- export ---------------------------------------------------------------------------------*/
/* clip to +-1, then nan_to_num(nan=0, posinf=1, neginf=-1): the closing lines of run_mastering_pipeline (backend/app/pipeline.py:
 * 1904-1906) and MasteringChain.process (backend/app/chain.py:93-94); in == out allowed */
int mm_dev_finalize_clip(mm_ctx*, const mm_geom*, const float* in, float* out);
/* _auto_blank_end (backend/app/pipeline.py:900-918): idx_dev[tracks] <- the last frame whose peak over the channels (after
 * the +-1 clip of export_audio) exceeds threshold_lin, -1 if none; the caller keeps min(n, idx + 1 + int(sr * min_silence)) */
int mm_dev_last_above(mm_ctx*, const mm_geom*, const float* in, double threshold_lin, int64_t* idx_dev);
/* export_audio's FLAC branch (pipeline.py:981-985) hands float32 to libsndfile as PCM_24: clip, lrintf(x * 0x7FFFFF);
 * interleaved int32 [tracks][n][ch] for a host-side encoder (parity unpinned: libsndfile is not available to compare) */
int mm_dev_quantize_pcm24(mm_ctx*, const mm_geom*, const float* in, int32_t* out_interleaved);
/* _write_wav_16bit_dithered quantiser  backend/app/pipeline.py:880-898
 * planar float rows -> interleaved int16 [tracks][n][ch].
 * noise: NULL -> TPDF from counter-based Philox4x32-10 keyed by (seed, track);
 *        else device float32 [tracks][n][ch] interleaved (bit-exact mode, _dither_noise_tpdf :830). */
int mm_dev_quantize_int16(mm_ctx*, const mm_geom*, const float* in, int16_t* out_interleaved,
                          const float* noise_interleaved, uint64_t seed);
/* Noise-shaped dither variants of the same export (dither_type "ns_e" = 1, "ns_itu" = 2; _dither_noise_ns_e / _ns_itu,
 * backend/app/pipeline.py:835-877): uniform_interleaved holds np.random.rand(n, ch).astype(float32) (bit-exact mode) or is
 * NULL (Philox).  Uses the context's workspace. */
int mm_dev_quantize_int16_shaped(mm_ctx*, const mm_geom*, const float* in, int16_t* out_interleaved,
                                 const float* uniform_interleaved, uint64_t dither_seed, int shape);

/* ---- analyzers ------------------------------------------------------------------------------*/
/* _true_peak_dbfs                      backend/app/routers/tools.py:44-54; out: device double[tracks] (dBFS) */
int mm_dev_true_peak(mm_ctx*, const mm_geom*, const float* in, double* tp_dbfs_dev);
/* compute_spectrum_bars                backend/app/pipeline.py:700-739
 * view: 0 = mean of channels, 1 = mid (L+R)/2, 2 = side (L-R)/2; out: device double[tracks][64] */
int mm_dev_spectrum_bars(mm_ctx*, const mm_geom*, const float* in, int view, double* bars_dev);
/* measure_stereo_correlation           backend/app/pipeline.py:766-791; out: device double[tracks]
 * (NaN encodes the reference's None), plus sample peak per track (double[tracks], may be NULL) */
int mm_dev_stereo_correlation(mm_ctx*, const mm_geom*, const float* in, double* corr_dev, double* sample_peak_dev);
/* mm_dev_true_peak and mm_dev_stereo_correlation in ONE pass over the samples (the FIR is FP32 bound: the correlation sums ride
 * on the samples its register windows hold); tp / corr / peak as in the two separate calls */
int mm_dev_true_peak_correlation(mm_ctx*, const mm_geom*, const float* in, double* tp_db, double* corr, double* peak);

/* mastering_trace.signal_metrics (backend/app/mastering_trace.py:115-149) as a device reduction: out3_dev[tracks][3] =
 * max |x| over the finite samples, count of non-finite samples, count of infinities. */
int mm_dev_signal_metrics(mm_ctx*, const mm_geom*, const float* in, double* out3_dev);

/* ---- whole chains (fused sweep plan; what bench.py times) ------------------------------------*/
#define MM_CHAIN_V1 1   /* run_mastering_pipeline            backend/app/pipeline.py:1800-1909 */
#define MM_CHAIN_V2 2   /* MasteringChain.default_chain(...).process + job fade-in
                           backend/app/chain.py:66-134, backend/app/routers/mastering.py:583 */
#define MM_FLAG_MEASURE_IN   1u  /* also measure_lufs(input)  */
#define MM_FLAG_MEASURE_OUT  2u  /* also measure_lufs(output) */
#define MM_FLAG_NO_JOB_FADE  4u  /* v2: return chain.process output without the job's fade-in */
#define MM_FLAG_ENVELOPE_COMPRESSOR 8u  /* multiband dynamics with the envelope (pedalboard-style) compressor, see below */
/* Which compressor apply_multiband_dynamics runs per band (backend/app/pipeline.py:442-474).  The reference takes the
 * envelope branch (_compress_band_pedalboard, :373-411: JUCE ballistics follower, attack / release 10/80, 10/80, 12/130,
 * 18/180 ms per band, gain (env / thr)^(1/ratio - 1)) whenever `pedalboard` imports and the memoryless soft-knee branch
 * otherwise.  SOFT_KNEE is pinned against the unmodified reference; ENVELOPE is PARITY UNPINNED (pedalboard is absent from the
 * build image and from the reference's tests): it restates the published JUCE arithmetic and is checked against its own CPU
 * restatement only (csrc/bandcomp.cu, oracle/chain.py). */
enum { MM_COMPRESSOR_SOFT_KNEE = 0, MM_COMPRESSOR_ENVELOPE = 1 };
/* styles: host array, one per track.  out_f32 planar (may alias in); out_i16 interleaved int16 or
 * NULL; noise as in mm_dev_quantize_int16; stats_dev: device mm_track_stats[tracks] or NULL. */
int mm_dev_master(mm_ctx*, const mm_geom*, int chain, const mm_style* styles_host,
                  const float* in, float* out_f32, int16_t* out_i16,
                  const float* noise_interleaved, uint64_t dither_seed,
                  mm_track_stats* stats_dev, uint32_t flags);

/* ---- one long file split in time over several GPUs (BASELINE config 5) --------------------------
 * Each rank holds a SLICE of one track: its own frames [own_lo, own_hi) (slice-local indices) plus margins on the
 * cut sides that are wide enough for every recurrence on the chain to forget its start (mm_slice_margin).  The
 * slice is mastered like a track of its own; the only coupling between ranks is the chain's global scalars --
 * channel means and peak (pipeline.py:134-149), the BS.1770 block sums (:644-664) and the output peak (:1899) --
 * which are reduced over the ranks' OWN frames through `allreduce` (NCCL in production: mm_b200/longform.py).
 * geom.tracks must be 1; geom.n is the slice length including margins.
 * allreduce(user, dev_ptr, count, dtype, op): in-place over the ranks, ordered after the work already queued on
 * the context's stream; dtype 0 = float64, 1 = int64, 2 = float32; op 0 = sum, 1 = min, 2 = max; returns 0 on
 * success.  NULL = single rank. */
typedef int (*mm_allreduce_fn)(void* user, void* dev_ptr, int64_t count, int dtype, int op);
typedef struct mm_slice {
    int64_t global_n;     /* frames of the whole file */
    int64_t global_off;   /* index in the file of the slice's frame 0 */
    int64_t own_lo, own_hi; /* frames of the slice this rank owns (multiples of 4 except at the file's end) */
    mm_allreduce_fn allreduce;
    void* user;
    void* nccl_comm;      /* ncclComm_t from mm_nccl_comm_create, or NULL: when set, the exchanges are ncclAllReduce calls enqueued
                             from C on the context's stream and `allreduce` is not used */
} mm_slice;
/* NCCL communicator for the exchange step, created from C (libnccl.so.2 resolved with dlopen at first use).  Rank 0 makes a
 * 128-byte id with mm_nccl_unique_id and hands it to the other ranks (any transport: torch.distributed broadcast in
 * mm_b200/longform.py); every rank then calls mm_nccl_comm_create.  mm_nccl_version: NCCL_VERSION_CODE, 0 if unavailable. */
int mm_nccl_unique_id(void* id128);
int mm_nccl_comm_create(mm_ctx*, const void* id128, int world, int rank, void** comm_out);
int mm_nccl_comm_destroy(void* comm);
int mm_nccl_version(void);
/* Margin (frames, multiple of 4096) a cut side needs at this sample rate. */
int64_t mm_slice_margin(int32_t sr);
int mm_dev_master_slice(mm_ctx*, const mm_geom*, int chain, const mm_style* style_host,
                        const float* in, float* out_f32, int16_t* out_i16,
                        const float* noise_interleaved, uint64_t dither_seed,
                        mm_track_stats* stats_dev, uint32_t flags, const mm_slice* slice);

/* Host-buffer drop-in for one call of run_mastering_pipeline / chain.process (+ optional
 * export): audio_in/out are (n, channels) interleaved float32 HOST arrays; pcm16_out may be NULL.
 * Copies in, masters `tracks` equally-shaped tracks stored back to back, copies out, syncs. */
int mm_master_host(mm_ctx*, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr,
                   const mm_style* styles_host, const float* audio_in, float* audio_out,
                   int16_t* pcm16_out, const float* noise_host, uint64_t dither_seed,
                   mm_track_stats* stats_host, uint32_t flags);

/* Job-level variant: PCM_16 frames in (the WAV data chunk of an upload, widened on the device as libsndfile does for
 * dtype="float32": x / 32768, backend/app/pipeline.py:814-817), PCM_16 (and/or float32) out.  Half the host->device
 * bytes of the float32 entry point; same pipelining, same results as mm_master_host on the widened input. */
int mm_master_host_pcm16(mm_ctx*, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr,
                         const mm_style* styles_host, const int16_t* pcm16_in, float* audio_out,
                         int16_t* pcm16_out, uint64_t dither_seed, mm_track_stats* stats_host, uint32_t flags);

/* One group of a batch of uploads (backend/app/routers/mastering.py:855-1037, /api/v2/batch): `tracks` equally-shaped tracks stored
 * back to back, float32 (audio_in) or PCM_16 (pcm16_in) -- exactly one of the two non-NULL.  track_ids[t] (NULL: t) keys track t's
 * dither stream, so that uploads of different shapes, which mm_b200.pipeline.master_wav_jobs masters group by group, each keep
 * the stream of their index in the caller's list: a track's result does not depend on what else was uploaded with it. */
int mm_master_host_ids(mm_ctx*, int chain, int32_t tracks, int64_t n, int32_t channels, int32_t sr, const mm_style* styles_host,
                       const float* audio_in, const int16_t* pcm16_in, float* audio_out, int16_t* pcm16_out, uint64_t dither_seed,
                       mm_track_stats* stats_host, uint32_t flags, const int32_t* track_ids);

/* A list of uploads of DIFFERENT shapes in one call (backend/app/routers/mastering.py:855-1037, /api/v2/batch: up to ten arbitrary
 * uploads, mastered one after the other by the reference).  Every job names its own frames / channels / rate, its own host buffers
 * (exactly one of audio_in / pcm16_in; all jobs of a call the same kind; any subset of the outputs), its style and the index of
 * its dither stream.  The library orders the jobs by shape, merges runs of equal shape into chunks and sends all chunks through ONE
 * copy-in / chain / copy-out pipeline, so that an upload's transfer overlaps its neighbour's chain even when no two uploads are
 * alike.  A job's result equals what the same upload gives alone with the same dither_id.  stats is written per job. */
typedef struct mm_host_job {
    int64_t n;
    int32_t channels;
    int32_t sr;
    const float* audio_in;      /* (n, channels) interleaved float32, or NULL */
    const int16_t* pcm16_in;    /* (n, channels) interleaved PCM_16 (widened on the device: x / 32768), or NULL */
    float* audio_out;           /* may be NULL */
    int16_t* pcm16_out;         /* may be NULL */
    mm_style style;
    int32_t dither_id;
    int32_t reserved;
    mm_track_stats stats;       /* out */
} mm_host_job;
int mm_master_host_jobs(mm_ctx*, int chain, int32_t njobs, mm_host_job* jobs, uint64_t dither_seed, uint32_t flags);

/* Pinned host memory for the host-buffer entry point (cudaMallocHost / cudaFreeHost). */
int mm_host_alloc(void** out, int64_t bytes);
int mm_host_free(void* p);

/* Host <-> device transfers on the context stream for the Python mirror's numpy arguments and results (replaces: the implicit
 * host arrays of backend/app/pipeline.py's function surface -- every stage takes and returns a numpy array, pipeline.py:134 ...).
 * A pinned host buffer moves as one DMA; a pageable one is staged through the context's pinned slots by several host threads
 * (MM_HOST_THREADS, default min(8, cores / 2)) with the DMA of finished blocks overlapping the staging of the next ones.
 * mm_ctx_copy_in returns once the host buffer has been read (the DMA may be in flight: stream order protects the device side);
 * mm_ctx_copy_out returns when the host buffer is complete. */
int mm_ctx_copy_in(mm_ctx*, void* dev_dst, const void* host_src, int64_t bytes);
int mm_ctx_copy_out(mm_ctx*, void* host_dst, const void* dev_src, int64_t bytes);

/* Lanes: how many child contexts (own stream and workspace, created on demand, owned by this context) a call may spread its
 * work over -- mm_dev_master splits a batch into that many runs of tracks (automatic: 2), the host entries send consecutive chunks
 * to consecutive lanes (automatic: 4).  One chain is ~21 dependent kernels and a second stream's kernels fill the tails and launch
 * gaps of the first: 64 one-track chains 169 -> 251 k audio-s/s, the 64-track batch 279 -> ~290 k.  Every lane count is bit-reproducible; between lane counts full-length tracks can
 * differ in the last float32 bits (a launch's size decides where its rows are cut into segments), as between batches of different sizes.
 * 0 = automatic (or MM_LANES), 1 = everything on the context's own stream.  No reference counterpart (an execution policy). */
int mm_ctx_set_lanes(mm_ctx*, int lanes);

/* Bytes of device workspace the context currently holds (for sizing sub-batches). */
int64_t mm_ctx_workspace_bytes(mm_ctx*);
/* Workspace the chain needs for a geometry (rows * stride * 4 * k + carries). */
int64_t mm_master_workspace_bytes(const mm_geom*, int chain);

/* ---- filter design (host, float64) -- scipy.signal.butter(order, Wn, btype, output="ba") and
 * lfilter_zi as used at backend/app/pipeline.py:175-183, :345-353, :590-599, :1231, :1303, :1427 */
#define MM_LOWPASS 0
#define MM_HIGHPASS 1
#define MM_BANDPASS 2
/* wn: 1 value (low/high) or 2 (band), normalised to Nyquist. b,a receive ncoef = order+1
 * (low/high) or 2*order+1 (band) values. Returns ncoef, or <0 on error. */
int mm_design_butter(int order, int btype, const double* wn, double* b, double* a);
/* scipy.signal.iirpeak(w0, Q) as scipy 1.18 evaluates it (b[3], a[3]) and the class apply_dynamic_eq's band falls into
 * before the signal is looked at: MM_DYNEQ_STABLE / MM_DYNEQ_LFILTER / MM_DYNEQ_MARGINAL, or -1 for an unstable section
 * (rmax: its largest pole radius).  kind / rmax may be NULL. */
int mm_design_iirpeak(double w0, double q, double* b, double* a, int* kind, double* rmax);
/* _build_linear_phase_ir (backend/app/pipeline.py:187-217): ir[n_fft] float32 */
int mm_design_linear_phase_ir(int sr, int n_fft, float* ir);
int mm_design_lfilter_zi(const double* b, const double* a, int ncoef, double* zi);
/* pyloudnorm K-weighting stage (0 = high shelf, 1 = high pass) for a sample rate; b[3], a[3]
 * (call sites backend/app/pipeline.py:646-648). */
int mm_design_k_weighting(int stage, double rate, double* b, double* a);
/* The K-weighting high-pass (b = g [1, -2, 1]) as the Chamberlin state-variable filter the loudness kernel runs in float32:
 * lp += f bp; hp = x - lp - q bp; bp += f hp; y = g hp.  fqg[3] = (f, q, g); abcd[9] = A (2x2 row-major), B[2], C[2], D of
 * s[n] = A s[n-1] + B x[n], y[n] = C s[n-1] + D x[n] with s = (lp, bp).  Host-side verification (tests/); no reference
 * counterpart: pyloudnorm runs scipy.signal.lfilter on (b, a) (call sites backend/app/pipeline.py:646-648). */
int mm_design_svf_highpass(const double* b, const double* a, double* fqg, double* abcd);
/* Tables of the chunked linear-recurrence scan for one section (host-side verification of the
 * tile decomposition in tests/): returns the look-back window W (tiles), <0 on error.
 * g[S*m], Pw[5*m*m], Plane[32*m*m], Qpow[(T/32+1)*m*m], Mpow[cap_w*m*m], Apow[(S+1)*m*m], zi[m];
 * S = samples per thread, T = threads per tile. Any output pointer may be NULL. */
int mm_design_scan_tables(const double* b, const double* a, int ncoef, double* g, double* Pw, double* Plane,
                          double* Qpow, double* Mpow, int cap_w, double* Apow, double* zi, int* S, int* T);

/* The same tables in a chosen realization: mode 0 = float64 DF2T (as above), 1 = the internally balanced
 * realization whose per-chunk recurrence the kernels evaluate in float32 (EQ-weighted sections, sections with
 * cut-offs above ~1 kHz, the loudness meter).  ss receives [A (m*m), B (m), C (m), D, ||A_balanced||_2]:
 * s[n] = A s[n-1] + B x[n], y[n] = C s[n-1] + D x[n]; zi is expressed in the same coordinates. */
int mm_design_scan_tables2(const double* b, const double* a, int ncoef, int mode, double* g, double* Pw, double* Plane,
                           double* Qpow, double* Mpow, int cap_w, double* Apow, double* zi, double* ss);

#ifdef __cplusplus
}
#endif
#endif /* MM_B200_H */
