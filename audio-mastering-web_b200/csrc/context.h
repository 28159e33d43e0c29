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

struct FilterPlan {
    Ba ba;
    ScanTables tabs;
    double* dev = nullptr;      // device copy laid out as common.cuh Tab<M>
    int pad = 0;                // filtfilt padlen = 3 * max(len(a), len(b))
};

struct LufsPlan {               // per (n, sr): gating blocks expressed over merged segments
    int nseg = 0, nblocks = 0, valid = 0, ntiles = 0;
    double scale = 0;
    long long* bnd = nullptr;
    int* tile_seg = nullptr;
    int* blk_lo = nullptr;
    int* blk_hi = nullptr;
};

struct Slot { void* p = nullptr; size_t cap = 0; };

enum SlotId {
    SL_E0 = 0, SL_E1, SL_E2, SL_E3, SL_T0, SL_T1, SL_T2, SL_T3, SL_T4,
    SL_ROWSTATS, SL_SUB, SL_MUL, SL_GAIN, SL_MUL_OUT, SL_SEGSUM, SL_LUFS, SL_TARGET, SL_GAINDB,
    SL_PEAKBITS, SL_WIDTH, SL_PARMIX, SL_PEAKIN, SL_MEAN, SL_NONFINITE, SL_LUFS2, SL_LUFS3,
    SL_STAGE_IL, SL_STAGE_PCM, SL_STAGE_NOISE, SL_STAGE_PL, SL_STATS, SL_ENV0, SL_ENV1, SL_MISC,
    SL_COUNT
};

struct KTime { std::string name; cudaEvent_t a, b; };

}  // namespace mm

struct mm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    mm::Slot slots[mm::SL_COUNT];
    std::map<std::string, mm::FilterPlan> plans;
    std::map<std::string, mm::LufsPlan> lufs_plans;
    int64_t launches = 0;
    bool timing = false;
    std::vector<mm::KTime> ktimes;
    std::map<std::string, std::pair<double, int64_t>> kacc;
    int64_t workspace_bytes = 0;
};

namespace mm {

int arena_get(mm_ctx* c, int slot, size_t bytes, void** out);
template <class T> inline int arena(mm_ctx* c, int slot, size_t count, T** out) {
    void* p = nullptr;
    int r = arena_get(c, slot, count * sizeof(T), &p);
    *out = reinterpret_cast<T*>(p);
    return r;
}
const FilterPlan* get_plan(mm_ctx* c, const Ba& ba);
int get_lufs_plan(mm_ctx* c, long long n, int sr, const LufsPlan** out);

struct KernelScope {            // brackets a launch with events when timing is on
    mm_ctx* c; const char* name; cudaEvent_t a = nullptr, b = nullptr;
    KernelScope(mm_ctx* ctx, const char* nm);
    ~KernelScope();
};

}  // namespace mm
