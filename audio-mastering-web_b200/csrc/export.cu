// Export-side helpers of export_audio (backend/app/pipeline.py:900-918, :965-991) that are not the 16-bit quantiser:
//   last_above_kernel : _auto_blank_end -- the last frame whose peak over the channels exceeds the threshold
//   pcm24_kernel      : the FLAC branch's sample conversion (libsndfile PCM_24 from float: lrintf(x * 0x7FFFFF)), interleaved
//                       int32 for a host-side encoder.  libsndfile is absent here and in the oracle: PARITY UNPINNED against
//                       the library itself; the formula is the one its f2flac24_array applies with normalisation on.
#include <algorithm>
#include <cmath>

#include "context.h"
#include "stages_internal.h"

namespace mm {

__global__ void __launch_bounds__(256) last_above_kernel(const float* in, long long n, long long stride, int channels, float thr,
                                                        long long* idx) {
    const int track = blockIdx.y;
    const float* r0 = in + (size_t)(track * channels) * (size_t)stride + kLead;
    const float* r1 = channels > 1 ? r0 + stride : r0;
    long long best = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = fminf(fabsf(r0[i]), 1.0f), b = fminf(fabsf(r1[i]), 1.0f);       // export_audio clips to +-1 first
        if (fmaxf(a, b) > thr) best = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(idx + track, best);
}

__global__ void __launch_bounds__(256) pcm24_kernel(const float* in, long long n, long long stride, int channels, int32_t* out) {
    const int track = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < channels; ++c) {
        float x = in[(size_t)(track * channels + c) * (size_t)stride + kLead + i];
        x = x != x ? 0.0f : fminf(fmaxf(x, -1.0f), 1.0f);
        out[((size_t)track * (size_t)n + (size_t)i) * channels + c] = __float2int_rn(__fmul_rn(x, 8388607.0f));
    }
}

// apply_maximizer_lookahead (pipeline.py:548-573): the first delay_n frames pass unlimited, the rest is the soft-knee maximizer
// of the signal delay_n frames EARLIER (that is what the reference's splice of `limited[delay_n:]` returns), with a cf-frame
// cross-fade before the seam against limited = maximizer(0) = 0.
__global__ void __launch_bounds__(256) lookahead_kernel(const float* in, float* out, long long n, long long stride, long long delay_n,
                                                       int cf, float max_k, float max_c, float max_top) {
    const size_t ro = (size_t)blockIdx.y * (size_t)stride + kLead;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r;
    if (i >= delay_n) {
        const float s = in[ro + i - delay_n];
        const float ax = fabsf(s);
        r = copysignf(fminf(fminf(ax, fmaf(ax, max_k, max_c)), max_top), s);
    } else if (i >= delay_n - cf) {
        const double a = (double)(i - (delay_n - cf) + 1) / (double)cf;
        r = __fadd_rn(__fmul_rn((float)(1.0 - a), in[ro + i]), __fmul_rn((float)a, 0.0f));
    } else {
        r = in[ro + i];
    }
    out[ro + i] = r;
}

int st_maximizer_lookahead(mm_ctx* c, const mm_geom* g, const float* in, float* out, long long delay_n, int cf) {
    DynParams d;
    fill_dyn(&d, 6.0, nullptr, 12.0);
    d.max_top = (float)std::pow(10.0, -0.3 / 20.0);       // the maximizer alone: its ceiling, no limiter behind it
    KernelScope ks(c, "maximizer_lookahead");
    lookahead_kernel<<<dim3((unsigned)((g->n + 255) / 256), (unsigned)(g->tracks * g->channels)), 256, 0, c->stream>>>(
        in, out, g->n, g->stride, delay_n, cf, d.max_k, d.max_c, d.max_top);
    MM_CUDA(cudaGetLastError());
    return 0;
}

// np.clip(x, -1, 1) followed by nan_to_num(nan=0, posinf=1, neginf=-1): the last two lines of run_mastering_pipeline (pipeline.py:
// 1904-1906) and of MasteringChain.process (chain.py:93-94)
__global__ void __launch_bounds__(256) finalize_clip_kernel(const float* in, float* out, long long n, long long stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)blockIdx.y * (size_t)stride + kLead + i;
    const float x = in[o];
    out[o] = (x != x) ? 0.0f : fminf(fmaxf(x, -1.0f), 1.0f);
}

int st_finalize_clip(mm_ctx* c, const mm_geom* g, const float* in, float* out) {
    KernelScope ks(c, "finalize_clip");
    finalize_clip_kernel<<<dim3((unsigned)((g->n + 255) / 256), (unsigned)(g->tracks * g->channels)), 256, 0, c->stream>>>(in, out, g->n, g->stride);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int st_last_above(mm_ctx* c, const mm_geom* g, const float* in, double threshold, long long* idx_dev) {
    MM_CUDA(cudaMemsetAsync(idx_dev, 0xff, (size_t)g->tracks * sizeof(long long), c->stream));      // -1
    const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((g->n + 255) / 256, 148LL * 8 / std::max(1, std::min(g->tracks, 148 * 8)) + 1));
    KernelScope ks(c, "auto_blank_last_above");
    last_above_kernel<<<dim3(bx, (unsigned)g->tracks), 256, 0, c->stream>>>(in, g->n, g->stride, g->channels, (float)threshold, idx_dev);
    MM_CUDA(cudaGetLastError());
    return 0;
}

int st_quantize_pcm24(mm_ctx* c, const mm_geom* g, const float* in, int32_t* out) {
    KernelScope ks(c, "quantize_pcm24");
    pcm24_kernel<<<dim3((unsigned)((g->n + 255) / 256), (unsigned)g->tracks), 256, 0, c->stream>>>(in, g->n, g->stride, g->channels, out);
    MM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mm
