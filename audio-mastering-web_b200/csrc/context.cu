#include "context.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int arena_get(mm_ctx* c, int slot, size_t bytes, void** out) {
    Slot& s = c->slots[slot];
    if (bytes == 0) bytes = 16;
    if (s.cap < bytes) {
        if (s.p) {
            MM_CUDA(cudaStreamSynchronize(c->stream));
            MM_CUDA(cudaFree(s.p));
            c->workspace_bytes -= (int64_t)s.cap;
            s.p = nullptr;
            s.cap = 0;
        }
        size_t want = (bytes + 255) & ~(size_t)255;
        MM_CUDA(cudaMalloc(&s.p, want));
        s.cap = want;
        c->workspace_bytes += (int64_t)want;
    }
    *out = s.p;
    return 0;
}

int kernel_setup(mm_ctx* c, const void* kern, int threads, size_t smem, bool max_carveout, int* blocks_per_sm) {
    auto it = c->occupancy.find(kern);
    if (it == c->occupancy.end()) {
        if (c->num_sms == 0) {
            cudaDeviceProp prop;
            MM_CUDA(cudaGetDeviceProperties(&prop, c->device));
            c->num_sms = prop.multiProcessorCount;
        }
        if (smem > 48 * 1024) MM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (max_carveout) MM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int bps = 0;
        MM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem));
        if (bps < 1) { set_error("kernel does not fit on an SM (%zu bytes of shared memory, %d threads)", smem, threads); return 1; }
        it = c->occupancy.emplace(kern, bps).first;
    }
    if (blocks_per_sm) *blocks_per_sm = it->second;
    return 0;
}

template <int M> static void pack_tables(const ScanTables& t, std::vector<double>& h) {
    typedef Tab<M> TB;
    h.assign((size_t)TB::Mpow + (size_t)t.W * TB::MM, 0.0);
    std::copy(t.Pw.begin(), t.Pw.end(), h.begin() + TB::Pw);
    std::copy(t.Plane.begin(), t.Plane.end(), h.begin() + TB::Plane);
    std::copy(t.Qpow.begin(), t.Qpow.end(), h.begin() + TB::Qpow);
    for (int i = 0; i < M; ++i) h[TB::Zi + i] = t.zi[i];
    std::copy(t.Apow.begin(), t.Apow.end(), h.begin() + TB::Apow);
    std::copy(t.Mpow.begin(), t.Mpow.end(), h.begin() + TB::Mpow);
}

static int pass2_policy() {            // -1 auto (default), 0 force float64, 1 force float32
    static int pol = -2;
    if (pol == -2) {
        const char* e = getenv("MM_PASS2");
        pol = -1;
        if (e && !strcmp(e, "f64")) pol = 0;
        if (e && !strcmp(e, "f32")) pol = 1;
    }
    return pol;
}

const FilterPlan* get_plan(mm_ctx* c, const Ba& ba, int prec) {
    int mode = kDf2tF64;
    if (ba.m == 2) {
        const int pol = pass2_policy();
        if (pol == 1) mode = kBalancedF32;
        else if (pol == -1) {
            if (prec == PREC_F32) mode = kBalancedF32;
            else if (prec == PREC_AUTO) mode = balanced_norm(ba) < 0.97 ? kBalancedF32 : kDf2tF64;
        }
    }
    return get_plan_mode(c, ba, mode);
}

const FilterPlan* get_plan_mode(mm_ctx* c, const Ba& ba, int mode) {
    if (ba.m != 2 && ba.m != 4) { set_error("unsupported section order %d (2 or 4)", ba.m); return nullptr; }
    if (ba.m != 2) mode = kDf2tF64;
    std::string key((const char*)&ba, sizeof(Ba));
    key.push_back((char)('0' + mode));
    auto it = c->plans.find(key);
    if (it != c->plans.end()) { it->second.last_use = ++c->tick; return &it->second; }
    FilterPlan p;
    p.ba = ba;
    bool ok = false;
    if (mode == kBalancedF32) ok = build_scan_tables_balanced(ba, kS, kT, &p.tabs);
    if (!ok) {
        if (mode == kBalancedF32) {         // not balanceable (non-minimal section): float64 DF2T instead
            return get_plan_mode(c, ba, kDf2tF64);
        }
        ok = build_scan_tables(ba, kS, kT, &p.tabs);
    }
    if (!ok) {
        set_error("scan tables: pole too close to the unit circle for the %d-sample tile", kL);
        return nullptr;
    }
    p.pad = 3 * (ba.m + 1);
    std::vector<double> h;
    if (ba.m == 2) pack_tables<2>(p.tabs, h); else pack_tables<4>(p.tabs, h);
    if (cudaMalloc(&p.dev, h.size() * sizeof(double)) != cudaSuccess) { set_error("cudaMalloc(filter tables) failed"); return nullptr; }
    if (cudaMemcpyAsync(p.dev, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error("upload of filter tables failed");
        return nullptr;
    }
    p.last_use = ++c->tick;
    auto res = c->plans.emplace(key, p);
    return &res.first->second;
}

const KwPlan* get_kw_plan(mm_ctx* c, int sr) {
    auto it = c->kw_plans.find(sr);
    if (it != c->kw_plans.end()) return &it->second;
    KwPlan p;
    StateSpace s0, s1, cas;
    const Ba f0 = k_weighting_stage(0, (double)sr), f1 = k_weighting_stage(1, (double)sr);
    if (!balanced_realization(f0, &s0, nullptr) || !svf_highpass_realization(f1, &s1, &p.hp_f, &p.hp_q, &p.hp_g)) {
        set_error("K-weighting at %d Hz: realization of the sections failed", sr);
        return nullptr;
    }
    cascade_realization(s0, s1, &cas);
    if (!build_scan_tables_ss(cas, kS, kT, &p.tabs) || !build_scan_tables_ss(cas, 2 * kS, kT, &p.tabs64) ||
        !build_scan_tables_ss(s0, kS, kT, &p.sec[0]) || !build_scan_tables_ss(s1, kS, kT, &p.sec[1])) {
        set_error("K-weighting at %d Hz: pole too close to the unit circle for the %d-sample tile", sr, kL);
        return nullptr;
    }
    std::vector<double> h;
    pack_tables<4>(p.tabs, h);
    if (cudaMalloc(&p.dev, h.size() * sizeof(double)) != cudaSuccess) { set_error("cudaMalloc(K-weighting tables) failed"); return nullptr; }
    if (cudaMemcpyAsync(p.dev, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error("upload of K-weighting tables failed");
        return nullptr;
    }
    if (cudaMalloc(&p.plane64, p.tabs64.Plane.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpyAsync(p.plane64, p.tabs64.Plane.data(), p.tabs64.Plane.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error("upload of K-weighting tables failed");
        return nullptr;
    }
    auto res = c->kw_plans.emplace(sr, p);
    return &res.first->second;
}

template <class T> static int upload_vec(mm_ctx* c, const std::vector<T>& v, T** dev) {
    MM_CUDA(cudaMalloc(dev, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) MM_CUDA(cudaMemcpyAsync(*dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    MM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// pyloudnorm block bounds (meter.py integrated_loudness): T_g = 0.4, step = 0.25,
//   numBlocks = int(round((T - T_g) / (T_g * step)) + 1),  l_j = int(T_g*(j*step)*rate),
//   u_j = int(T_g*(j*step + 1)*rate)   -- evaluated in the same float64 order here.
int get_lufs_plan(mm_ctx* c, long long n, int sr, const LufsPlan** out, long long goff, long long n_local) {
    char keyb[96];
    if (n_local < 0) n_local = n;
    snprintf(keyb, sizeof(keyb), "%lld:%d:%lld:%lld", n, sr, goff, n_local);
    auto it = c->lufs_plans.find(keyb);
    if (it != c->lufs_plans.end()) { it->second.last_use = ++c->tick; *out = &it->second; return 0; }
    LufsPlan p;
    const double rate = (double)sr, T_g = 0.4, step = 1.0 - 0.75;
    p.valid = !((double)n < T_g * rate);
    p.scale = 1.0 / (T_g * rate);
    const long long q_last = kLead + n_local - 1;
    p.ntiles = (int)((q_last + kL) / kL);
    std::vector<long long> lo, hi, bnd;
    if (p.valid) {
        const double T = (double)n / rate;
        const long long nb = (long long)std::nearbyint((T - T_g) / (T_g * step)) + 1;   // np.round = half-to-even
        for (long long j = 0; j < nb; ++j) {
            const long long l = (long long)(T_g * ((double)j * step) * rate);
            long long u = (long long)(T_g * ((double)j * step + 1.0) * rate);
            lo.push_back(l);
            hi.push_back(std::min(u, n));      // numpy slicing clamps at the array end
        }
        bnd = lo;
        bnd.insert(bnd.end(), hi.begin(), hi.end());
        std::sort(bnd.begin(), bnd.end());
        bnd.erase(std::unique(bnd.begin(), bnd.end()), bnd.end());
    }
    if (bnd.size() < 2) { bnd.clear(); bnd.push_back(0); bnd.push_back(0); p.valid = p.valid && false; }
    p.nseg = (int)bnd.size() - 1;
    p.nblocks = (int)lo.size();
    // the loudness kernel splits a warp-tile (2048 or 1024 samples) over at most three hops (lufs_kernel.cuh)
    p.min_span2 = 1LL << 40;
    for (size_t s2 = 0; p.valid && s2 + 3 < bnd.size(); ++s2) p.min_span2 = std::min(p.min_span2, bnd[s2 + 2] - bnd[s2]);
    if (p.valid && p.min_span2 < 1024) {
        set_error("loudness: %d Hz is too low a sample rate for the block partition of this kernel (100 ms hops of at least 512 samples)", sr);
        return 1;
    }
    std::vector<int> blo(lo.size()), bhi(lo.size());
    for (size_t j = 0; j < lo.size(); ++j) {
        blo[j] = (int)(std::lower_bound(bnd.begin(), bnd.end(), lo[j]) - bnd.begin());
        bhi[j] = (int)(std::lower_bound(bnd.begin(), bnd.end(), hi[j]) - bnd.begin());
    }
    std::vector<int> tseg(p.ntiles);
    for (int t = 0; t < p.ntiles; ++t) {
        long long i0 = std::max<long long>((long long)t * kL - kLead, 0) + goff;
        int s = (int)(std::upper_bound(bnd.begin(), bnd.end(), i0) - bnd.begin()) - 1;
        tseg[t] = std::min(std::max(s, 0), p.nseg);
    }
    MM_TRY(upload_vec(c, bnd, &p.bnd));
    MM_TRY(upload_vec(c, tseg, &p.tile_seg));
    MM_TRY(upload_vec(c, blo, &p.blk_lo));
    MM_TRY(upload_vec(c, bhi, &p.blk_hi));
    p.last_use = ++c->tick;
    auto res = c->lufs_plans.emplace(keyb, p);
    *out = &res.first->second;
    return 0;
}

// Filter plans are keyed by coefficients the caller controls (crossovers, cut-offs, dynamic-EQ bands), loudness plans by the
// exact frame count of an upload: a service would grow both without bound.  Beyond the caps the least-recently-used half is
// dropped -- only here, at API entry, when no stage holds a plan pointer -- after the stream has drained (kernels may still be
// reading the device tables).
constexpr size_t kMaxFilterPlans = 192, kMaxLufsPlans = 24;
template <class Map, class Free> static void evict_lru(mm_ctx* c, Map& m, size_t cap, Free free_one, bool* synced) {
    if (m.size() <= cap) return;
    std::vector<uint64_t> uses;
    for (auto& kv : m) uses.push_back(kv.second.last_use);
    std::nth_element(uses.begin(), uses.begin() + uses.size() / 2, uses.end());
    const uint64_t cut = uses[uses.size() / 2];
    if (!*synced) { cudaStreamSynchronize(c->stream); *synced = true; }
    for (auto it = m.begin(); it != m.end();) {
        if (it->second.last_use < cut) { free_one(it->second); it = m.erase(it); }
        else ++it;
    }
}
void plan_gc(mm_ctx* c) {
    bool synced = false;
    evict_lru(c, c->plans, kMaxFilterPlans, [](FilterPlan& p) { if (p.dev) cudaFree(p.dev); }, &synced);
    evict_lru(c, c->lufs_plans, kMaxLufsPlans, [](LufsPlan& p) { cudaFree(p.bnd); cudaFree(p.tile_seg); cudaFree(p.blk_lo); cudaFree(p.blk_hi); }, &synced);
}

KernelScope::KernelScope(mm_ctx* ctx, const char* nm) : c(ctx), name(nm) {
    c->launches += 1;
    if (c->timing) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
    }
}
KernelScope::~KernelScope() {
    if (c->timing && a) {
        cudaEventRecord(b, c->stream);
        KTime k;
        k.name = name; k.a = a; k.b = b; k.samples = samples;
        c->ktimes.push_back(k);
    }
}

}  // namespace mm
