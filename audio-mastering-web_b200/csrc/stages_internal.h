// Internal (C++) stage interface shared by stages.cu, analyzers.cu and capi.cu.
#pragma once
#include "common.cuh"
#include "context.h"
#include "pw_args.h"

namespace mm {


struct Pro {                     // prologue applied while loading x-domain samples
    int mode = PRO_NONE;
    const double* sub = nullptr;
    const double* mul = nullptr;
};

struct Epi {                     // epilogue of a backward sweep
    int mode = EPI_STORE;
    const float* aux0 = nullptr;
    const float* aux1 = nullptr;
    Pro auxp;                    // prologue applied to aux0 when aux_pro != 0
    int aux_pro = 0;
    double w[4] = {0, 0, 0, 0};
    double wc = 1.0, trim = 1.0;
    const DynParams* dyn = nullptr;
    double exc_gain = 0, exc_k = 2.5;
    int exc_mode = 0;
    float* peak = nullptr;
    int clip = 0;                // clip the recombined output to +-1 (apply_high_freq_trim)
    // mixed-preset batches: per-row overrides (device arrays indexed by batch row; see SweepArgs)
    const double* w_row = nullptr;
    const double* exc_row = nullptr;
    const unsigned char* peak_row = nullptr;
};

struct Bufs { float* E[4]; float* T[5]; };

int check_geom(const mm_geom* g);
int get_bufs(mm_ctx* c, const mm_geom* g, Bufs* B);
const FilterPlan* plan_butter(mm_ctx* c, int order, BType bt, double w0, double w1, int prec = PREC_AUTO);

int sweep_fwd(mm_ctx* c, const mm_geom* g, int nf, int nin, const FilterPlan* const* plans, const float* const* in,
              float* const* out, const Pro& pro, int pad);
int sweep_bwd(mm_ctx* c, const mm_geom* g, int nf, const FilterPlan* const* plans, const float* const* in,
              float* const* out, int nout, const Epi& epi, int pad);

int run_row_stats(mm_ctx* c, const mm_geom* g, const float* in, RowStats** st_out);
int exchange_row_stats(mm_ctx* c, RowStats* st, int rows);
int slice_allreduce(mm_ctx* c, void* ptr, int64_t count, int dtype, int op, const char* what);
// nccl_shim.cu
int nccl_allreduce(mm_ctx* c, void* comm, void* ptr, int64_t count, int dtype, int op);
int run_in_scalars(mm_ctx* c, const mm_geom* g, const RowStats* st, int use_dc, int use_guard, double headroom_db,
                   double* sub, double* mul, double* peak_track, double* mean_row);
int run_pointwise(mm_ctx* c, const mm_geom* g, PwArgs& A, const char* name);
int run_out_scalars(mm_ctx* c, const OutScalarArgs& O);
int reset_imager_peaks(mm_ctx* c, float* peak, const double* width, int tracks);
struct FinalArgs;
int run_finalize(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* mul, const double* width, int n_fade,
                 int16_t* pcm, const float* noise, unsigned long long seed, double* nonfinite);
int run_quantize(mm_ctx* c, const QuantArgs& Q);
int run_signal_metrics(mm_ctx* c, const mm_geom* g, const float* in, double* out3);
int run_white_noise(mm_ctx* c, const WhiteArgs& W);
// dir 0: interleaved -> planar, 1: planar -> interleaved
int run_layout(mm_ctx* c, const mm_geom* g, const float* interleaved, float* planar, int dir);
int run_layout_pcm16(mm_ctx* c, const mm_geom* g, const int16_t* interleaved, float* planar);
void fill_dyn(DynParams* d, double knee_db, const double* band_ratios, double max_upward_boost_db);
void fill_parallel(DynParams* d, double ratio, double threshold_db);

int st_target_curve(mm_ctx* c, const mm_geom* g, const float* in, float* out, const Pro& pro);
int st_dynamics(mm_ctx* c, const mm_geom* g, const float* in, float* out, double knee_db, const double* crossovers_hz,
                const double* band_ratios, double max_upward_boost_db, const double* par_mix_rows, float* peak,
                int bands_only = 0,      // bands_only: apply_multiband_dynamics alone (no maximizer, no limiter)
                int compressor = 0);     // MM_COMPRESSOR_SOFT_KNEE (numpy branch) or MM_COMPRESSOR_ENVELOPE (pedalboard-style branch)
// bandcomp.cu: followers + limiters + sum + maximizer + limiter over the four float32 bands
int launch_band_compress(mm_ctx* c, const mm_geom* g, const float* const* bands, float* out, const DynParams& d);
int st_lufs(mm_ctx* c, const mm_geom* g, const float* in, const Pro& pro, double* lufs_dev, const double* target_dev,
            double* gain_row, double* gain_db);
int st_final_balance(mm_ctx* c, const mm_geom* g, const float* in, float* out, const Pro& pro, float* peak);
int st_filtfilt_combine(mm_ctx* c, const mm_geom* g, const FilterPlan* plan, const float* in, float* out, const Epi& epi,
                        const Pro& pro);
int st_style_eq(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* gain_db, float* peak, int* fired,
                int reset_peak);
int st_exciter(mm_ctx* c, const mm_geom* g, const float* in, float* out, double exciter_db, int mode, float* peak);

// _split_bands (pipeline.py:333-364): the four zero-phase bands of every row, left in workspace buffers (bands[0..3])
int st_split_bands(mm_ctx* c, const mm_geom* g, const float* in, const double* cross_hz, float** bands);
// reverb.cu: type 0 plate, 1 room, 2 hall, 3 theater, 4 cathedral (in == out allowed)
int st_reverb(mm_ctx* c, const mm_geom* g, const float* in, float* out, int type, double decay_sec, double mix, int use_ms,
              double mix_mid, double mix_side);
// spectral.cu: compute_spectral_envelope (pipeline.py:1527-1551): env_dev[tracks][4097] float32
int st_spectral_envelope(mm_ctx* c, const mm_geom* g, const float* in, float* env_dev);
// bigfft.cu: scipy.signal.resample of every row (gi->n -> go->n frames; same tracks / channels); plan cache per context
int st_fft_resample(mm_ctx* c, const mm_geom* gi, const float* in, const mm_geom* go, float* out);
void bigfft_release(mm_ctx* c);
// bigfft.cu: the same "same"-mode FIR as st_fir_same, evaluated as one circular FFT convolution per track (both channels in one
// complex transform); usable while n + K - 1 <= 2^27
bool fft_convolve_fits(const mm_geom* g, int K);
int st_fft_convolve_same(mm_ctx* c, const mm_geom* g, const float* in, float* out, const float* taps_dev, int K, int clip);
// export.cu: _auto_blank_end's scan (idx_dev[tracks]: last frame above the threshold, -1 if none) and the PCM_24 conversion
int st_finalize_clip(mm_ctx* c, const mm_geom* g, const float* in, float* out);
int st_last_above(mm_ctx* c, const mm_geom* g, const float* in, double threshold, long long* idx_dev);
int st_quantize_pcm24(mm_ctx* c, const mm_geom* g, const float* in, int32_t* out);
// denoise.cu: apply_spectral_denoise (pipeline.py:1472-1524); not in place
int st_spectral_denoise(mm_ctx* c, const mm_geom* g, const float* in, float* out, double strength, double noise_percentile);
// followers.cu: out[i] = sum_k taps[k] x[i + (K-1)/2 - k] (fftconvolve mode="same"), K a multiple of 64, taps on the device
int st_fir_same(mm_ctx* c, const mm_geom* g, const float* in, float* out, const float* taps_dev, int K, int clip);
// followers.cu
int st_target_curve_linear_phase(mm_ctx* c, const mm_geom* g, const float* in, float* out);
int st_imager4(mm_ctx* c, const mm_geom* g, const float* in, float* out, const double* widths, const double* crossovers_hz);
int st_transient_designer(mm_ctx* c, const mm_geom* g, const float* in, float* out, double attack_gain, double sustain_gain);
int st_maximizer_transient_aware(mm_ctx* c, const mm_geom* g, const float* in, float* out, double sensitivity);
int st_haas_imager(mm_ctx* c, const mm_geom* g, const float* in, float* out, double width, double delay_ms, double mix);

// analyzers.cu / deesser.cu
int st_deesser(mm_ctx* c, const mm_geom* g, const float* in, float* out, double threshold_db, double ratio, double freq_lo,
               double freq_hi, double attack_ms, double release_ms);
// export.cu: apply_maximizer_lookahead (pipeline.py:548-573); not in place
int st_maximizer_lookahead(mm_ctx* c, const mm_geom* g, const float* in, float* out, long long delay_n, int cf);
// deesser.cu: apply_dynamic_eq over nbands x {w0, bw, threshold_db, ratio, attack_ms, release_ms, max_cut_db}; in == out allowed
// class of a dynamic-EQ section before the signal is looked at: MM_DYNEQ_STABLE / _LFILTER / _MARGINAL, or -1 (unstable)
int dyneq_band_kind(const Ba& ba, double* rmax_out);
int st_dynamic_eq(mm_ctx* c, const mm_geom* g, const float* in, float* out, int nbands, const double* params, unsigned flags = 0,
                  int* classes = nullptr);
int st_true_peak(mm_ctx* c, const mm_geom* g, const float* in, double* tp_dev);
// true peak + stereo correlation + sample peak in one pass over the samples (stereo); mono falls back to the two kernels
int st_true_peak_corr(mm_ctx* c, const mm_geom* g, const float* in, double* tp_dev, double* corr_dev, double* peak_dev);
int st_spectrum_bars(mm_ctx* c, const mm_geom* g, const float* in, int view, double* bars_dev);
int st_correlation(mm_ctx* c, const mm_geom* g, const float* in, double* corr_dev, double* peak_dev);

}  // namespace mm
