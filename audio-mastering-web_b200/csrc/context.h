// Host-side context: stream, device arena, filter-plan cache, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/mm_b200.h"
#include "design.h"

namespace mm {

void set_error(const char* fmt, ...);

#define MM_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
#define MM_TRY(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r != 0) return _r; \
    } while (0)

// How a section's pass 2 (the per-sample recurrence inside a 32-sample chunk) is evaluated.  The chunk-start
// states always come from the float64 scan.
//   PREC_F64  : float64 DF2T, bit-for-bit the arithmetic of scipy's lfilter up to summation order
//   PREC_F32  : float32 on the balanced realization -- for sections whose output reaches the signal only through a
//               small recombination weight (EQ bands, exciter side chain) and for the loudness meter
//   PREC_AUTO : float32 when state errors die within a chunk (||A_balanced||_2 < 0.97: cut-offs above ~1 kHz),
//               float64 otherwise -- for sections the full signal passes through
// MM_PASS2=f64 forces float64 everywhere, MM_PASS2=f32 float32 wherever a float32 kernel exists (experiments).
enum Prec { PREC_F64 = 0, PREC_F32 = 1, PREC_AUTO = 2 };

struct FilterPlan {
    Ba ba;
    ScanTables tabs;
    double* dev = nullptr;      // device copy laid out as common.cuh Tab<M>
    int pad = 0;                // filtfilt padlen = 3 * max(len(a), len(b))
    uint64_t last_use = 0;
};

struct KwPlan {                 // K-weighting cascade (shelf -> high-pass) of one sample rate as a 4-state system
    ScanTables tabs;            // tables of the cascade in [balanced shelf; state-variable high-pass (lp, bp)] coordinates, 32-sample chunks
    ScanTables tabs64;          // the same for 64-sample chunks (the loudness kernel's default)
    ScanTables sec[2];          // the two sections' realizations (A, B, C, D used): balanced shelf, state-variable high-pass
    double hp_f = 0, hp_q = 0, hp_g = 0;   // the high-pass as a Chamberlin state-variable filter (design.h)
    double* dev = nullptr;      // Tab<4> of `tabs`
    double* plane64 = nullptr;  // tabs64.Plane ([32][16]) on the device
};

struct LufsPlan {               // per (n, sr): gating blocks expressed over merged segments
    int nseg = 0, nblocks = 0, valid = 0, ntiles = 0;
    long long min_span2 = 0;    // min over s of bnd[s + 2] - bnd[s] (interior): a warp-tile no longer than this holds <= 2 hop ends
    double scale = 0;
    long long* bnd = nullptr;
    int* tile_seg = nullptr;
    int* blk_lo = nullptr;
    int* blk_hi = nullptr;
    uint64_t last_use = 0;
};

struct Slot { void* p = nullptr; size_t cap = 0; };

enum SlotId {
    SL_E0 = 0, SL_E1, SL_E2, SL_E3, SL_T0, SL_T1, SL_T2, SL_T3, SL_T4,
    SL_ROWSTATS, SL_SUB, SL_MUL, SL_GAIN, SL_MUL_OUT, SL_SEGSUM, SL_LUFS, SL_TARGET, SL_GAINDB,
    SL_PEAKBITS, SL_WIDTH, SL_PARMIX, SL_PEAKIN, SL_MEAN, SL_NONFINITE, SL_LUFS2, SL_LUFS3,
    SL_STAGE_IL, SL_STAGE_PCM, SL_STAGE_NOISE, SL_STAGE_PL, SL_STATS, SL_ENV0, SL_ENV1, SL_MISC,
    SL_STAGE_IL1, SL_STAGE_PCM1, SL_STAGE_NOISE1, SL_STAGE_OL, SL_STAGE_OL1, SL_XCHG, SL_ROWMAP, SL_REV0, SL_REV1,
    SL_DN_MAG, SL_DN_NOISE, SL_BIGFFT, SL_BIGFFT_H, SL_ENV_WIN, SL_ENV_TW, SL_TRACKIDS, SL_ROWW, SL_ROWFLAG,
    SL_COUNT
};

struct KTime { std::string name; cudaEvent_t a, b; double samples; };
struct KAcc { double ms = 0; int64_t launches = 0; double samples = 0; };

}  // namespace mm

struct mm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr, d2h_stream2 = nullptr;   // d2h_stream2: odd chunks' copy-out (MM_D2H_STREAMS=2)
    const int* row_map = nullptr;       // set around the per-style stages of a mixed batch: device list of the rows to visit
    int row_map_rows = 0;
    const int32_t* track_ids_host = nullptr;   // set for the duration of mm_master_host_ids: dither-stream index of every track of the call
    const int* track_ids_dev = nullptr;        // ... and the current chunk's slice of it on the device
    const mm_slice* slice = nullptr;    // set for the duration of mm_dev_master_slice: the batch is a time slice of one file   // copy streams of the host-buffer entry point (lazy)
    mm::Slot slots[mm::SL_COUNT];
    cudaEvent_t pin_in_done = nullptr;  // recorded after the last DMA out of host_pin[0] by mm_ctx_copy_in (the slot is rewritten only after it)
    mm::Slot host_pin[4];               // pinned staging ring of mm_master_host_jobs (pageable uploads and results pass through it): 0, 1 in; 2, 3 out
    std::map<std::string, mm::FilterPlan> plans;
    std::map<std::string, mm::LufsPlan> lufs_plans;
    std::map<int, mm::KwPlan> kw_plans;
    int64_t launches = 0;
    bool timing = false;
    std::vector<mm::KTime> ktimes;
    std::map<std::string, mm::KAcc> kacc;
    int64_t workspace_bytes = 0;
    std::map<const void*, int> occupancy;   // kernel -> resident CTAs per SM on THIS device (attributes set when the entry is made)
    int num_sms = 0;
    // LANES: child contexts (own stream, own workspace and plan caches) a call may spread its sub-batches / chunks over so that
    // the tails, launch gaps and dependent-kernel bubbles of one chain are filled by another's kernels.  lanes[i] is lane i + 1;
    // lane 0 is this context.  Owned by the parent, created lazily, never shared between threads (contexts are per thread).
    std::vector<mm_ctx*> lanes;
    int lanes_cfg = 0;                      // 0: automatic (MM_LANES or the entry point's default); >= 1: fixed (mm_ctx_set_lanes)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // parent: "everything queued so far"; lane: "this lane's part of the call is queued"
    int grid_div = 1;                       // persistent grids are sized for 1 / grid_div of the device (lanes of mm_master_host_jobs share it)
    std::map<int, float*> lp_taps;          // linear-phase target-curve IR per sample rate (device)
    uint64_t tick = 0;                      // use counter of the plan caches (least-recently-used eviction at API entry)
    void* bigfft = nullptr;             // bigfft.cu's plan cache (FFT tables, chirp-filter spectra); owned by the context: contexts are per thread
};

namespace mm {

int arena_get(mm_ctx* c, int slot, size_t bytes, void** out);
template <class T> inline int arena(mm_ctx* c, int slot, size_t count, T** out) {
    void* p = nullptr;
    int r = arena_get(c, slot, count * sizeof(T), &p);
    *out = reinterpret_cast<T*>(p);
    return r;
}
const FilterPlan* get_plan(mm_ctx* c, const Ba& ba, int prec = PREC_F64);
const FilterPlan* get_plan_mode(mm_ctx* c, const Ba& ba, int mode);   // mode: design.h Realization
// n, sr: the (whole) signal the gating blocks are laid over; local tiles cover [goff, goff + n_local) of it
int get_lufs_plan(mm_ctx* c, long long n, int sr, const LufsPlan** out, long long goff = 0, long long n_local = -1);
const KwPlan* get_kw_plan(mm_ctx* c, int sr);

// Per-context (= per device, per thread) launch configuration of a kernel that needs more than 48 KB of dynamic shared memory:
// sets cudaFuncAttributeMaxDynamicSharedMemorySize (a per-device attribute) the first time this context sees the kernel and
// caches the resident CTAs per SM.  Nothing here is process-global: contexts on different devices do not share state.
int kernel_setup(mm_ctx* c, const void* kern, int threads, size_t smem, bool max_carveout, int* blocks_per_sm);
// drop least-recently-used filter / loudness plans beyond the caps (called at API entry, when no plan pointer is held)
void plan_gc(mm_ctx* c);

struct KernelScope {            // brackets a launch with events when timing is on
    mm_ctx* c; const char* name; cudaEvent_t a = nullptr, b = nullptr;
    double samples = 0;         // channel-samples this launch visits (per-kernel GB/s of launches over row lists / track runs)
    KernelScope(mm_ctx* ctx, const char* nm);
    ~KernelScope();
};

}  // namespace mm
