// Reference mastering (backend/app/pipeline.py:1527-1612): compute_spectral_envelope + apply_reference_match.
//
//   envelope : mono = mean over channels (float32); frames of 8192 samples, hop 2048, Hann window; RMS over the frames of
//              |rfft|  per bin (4097 bins).  Each 8192-point real FFT is one 4096-point complex FFT of the
//              even/odd-packed frame (16 points per thread in registers, radix 16.16.16) plus the real-FFT untangling step; a CTA walks a run of
//              frames and keeps the 4097 float64 power sums in shared memory, one atomic flush per CTA.
//   matching : ratio curve -> Savitzky-Golay smoothing -> 8192-tap FIR are DESIGN steps on 4097 numbers (host, numpy/scipy
//              in mm_b200/pipeline.py, as the reference does); the FIR itself runs through fir_same_kernel (followers.cu).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "context.h"
#include "fft.cuh"
#include "stages_internal.h"

namespace mm {

constexpr int kEnvN = 8192;             // n_fft
constexpr int kEnvM = kEnvN / 2;        // complex FFT size
constexpr int kEnvHop = kEnvN / 4;
constexpr int kEnvThreads = 256;
constexpr int kEnvBins = kEnvN / 2 + 1;

struct EnvArgs2 {
    const float* in;
    long long n, stride;
    int channels, frames, frames_per_cta;
    double* accum;          // [tracks][kEnvBins]
    const float* window;    // [kEnvN] np.hanning(8192).astype(float32)
    const float2* tw;       // [kEnvM] e^{-2 pi i k / 4096}
};

__global__ void envelope_tables_kernel(float* window, float2* tw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    // np.hanning(M).astype(float32): 0.5 - 0.5 cos(2 pi j / (M - 1)) in float64, cast
    if (j < kEnvN) window[j] = (float)(0.5 - 0.5 * cospi(2.0 * (double)j / (double)(kEnvN - 1)));
    if (j < kEnvM) {
        double s, c;
        sincospi(-2.0 * (double)j / (double)kEnvM, &s, &c);
        tw[j] = make_float2((float)c, (float)s);
    }
}

__device__ __forceinline__ int epad(int i) { return i + (i >> 4); }
constexpr int kEnvBuf = kEnvM + kEnvM / 16;

// One frame per CTA pass: the 4096-point packed transform is 16 points per thread in registers -- radix 16 . 16 . 16 with two
// exchanges through padded shared memory -- then the real-FFT untangling straight into the float64 power sums.
__global__ void __launch_bounds__(kEnvThreads) envelope_kernel(const EnvArgs2 P) {
    extern __shared__ __align__(16) unsigned char esm[];
    float2* buf = reinterpret_cast<float2*>(esm);
    float2* tw = buf + kEnvBuf;
    double* acc = reinterpret_cast<double*>(tw + kEnvM);
    const int track = blockIdx.y, j = threadIdx.x;
    const int f0 = blockIdx.x * P.frames_per_cta, f1 = min(P.frames, f0 + P.frames_per_cta);
    for (int k = j; k < kEnvBins; k += kEnvThreads) acc[k] = 0.0;
    for (int k = j; k < kEnvM; k += kEnvThreads) tw[k] = P.tw[k];
    const float* r0 = P.in + (size_t)(track * P.channels) * (size_t)P.stride + kLead;
    const float* r1 = r0 + (P.channels > 1 ? (size_t)P.stride : 0);
    const int kk = j & 15;
    for (int f = f0; f < f1; ++f) {
        const long long start = (long long)f * kEnvHop;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int m = j + 256 * q;                       // packed sample m = (x[2m], x[2m+1])
            const float2 a = *reinterpret_cast<const float2*>(r0 + start + 2 * m);
            const float2 w = *reinterpret_cast<const float2*>(P.window + 2 * m);
            float2 x = a;
            if (P.channels > 1) {
                // np.mean(audio, axis=1).astype(float32): float32 sum of the two channels, divided by 2
                const float2 b2 = *reinterpret_cast<const float2*>(r1 + start + 2 * m);
                x = make_float2(__fmul_rn(__fadd_rn(a.x, b2.x), 0.5f), __fmul_rn(__fadd_rn(a.y, b2.y), 0.5f));
            }
            v[q] = make_float2(__fmul_rn(x.x, w.x), __fmul_rn(x.y, w.y));
        }
        __syncthreads();                                     // the previous frame's untangling is done with buf
        dft_reg<16>(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[17 * j + k] = v[dft_pos<16>(k)];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float2 x = buf[epad(j + 256 * q)];
            if (q) x = cmulf(x, tw[16 * kk * q]);            // W_256^{k q}
            v[q] = x;
        }
        __syncthreads();
        dft_reg<16>(v);
        const int o = (j - kk) * 16 + kk;
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[epad(o + 16 * k)] = v[dft_pos<16>(k)];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float2 x = buf[epad(j + 256 * q)];
            if (q) x = cmulf(x, tw[j * q]);                  // W_4096^{j q}
            v[q] = x;
        }
        __syncthreads();
        dft_reg<16>(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[epad(j + 256 * k)] = v[dft_pos<16>(k)];      // Z[j + 256 k], natural order
        __syncthreads();
        // X[k] = (Z[k] + conj Z[M-k]) / 2 - (i / 2) e^{-2 pi i k / N} (Z[k] - conj Z[M-k]),  k = 0 .. M  (Z[M] = Z[0])
        for (int k = j; k <= kEnvM; k += kEnvThreads) {
            const float2 zk = buf[epad(k & (kEnvM - 1))], zm = buf[epad((kEnvM - k) & (kEnvM - 1))];
            const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);        // even part
            const float dr = 0.5f * (zk.x - zm.x), di = 0.5f * (zk.y + zm.y);        // (Z[k] - conj Z[M-k]) / 2
            float s, c;
            sincospif(-2.0f * (float)k / (float)kEnvN, &s, &c);
            // -i (dr + i di) = di - i dr ; times (c + i s)
            const float orr = di * c + dr * s, oi = di * s - dr * c;
            const float xr = er + orr, xi = ei + oi;
            acc[k] += (double)(xr * xr + xi * xi);
        }
    }
    __syncthreads();
    double* dstp = P.accum + (size_t)track * kEnvBins;
    for (int k = j; k < kEnvBins; k += kEnvThreads)
        if (acc[k] != 0.0) atomicAdd(dstp + k, acc[k]);
}

__global__ void envelope_final_kernel(const double* accum, int tracks, int frames, float* env) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tracks * kEnvBins) return;
    env[i] = frames > 0 ? (float)sqrt(accum[i] / (double)frames) : 1.0f;            // count == 0 -> ones (pipeline.py:1549-1550)
}

int st_spectral_envelope(mm_ctx* c, const mm_geom* g, const float* in, float* env_dev) {
    const int frames = g->n >= kEnvN ? (int)((g->n - kEnvN) / kEnvHop + 1) : 0;
    double* accum;
    MM_TRY(arena(c, SL_XCHG, (size_t)g->tracks * kEnvBins, &accum));
    MM_CUDA(cudaMemsetAsync(accum, 0, (size_t)g->tracks * kEnvBins * sizeof(double), c->stream));
    if (frames > 0) {
        const size_t smem = (size_t)(kEnvBuf + kEnvM) * sizeof(float2) + kEnvBins * sizeof(double) + 16;
        float* window;
        float2* twd;
        MM_TRY(arena(c, SL_ENV_WIN, (size_t)kEnvN, &window));
        MM_TRY(arena(c, SL_ENV_TW, (size_t)kEnvM, &twd));
        envelope_tables_kernel<<<kEnvN / 256, 256, 0, c->stream>>>(window, twd);
        MM_CUDA(cudaGetLastError());
        MM_CUDA(cudaFuncSetAttribute(envelope_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      // per device; cheap
        EnvArgs2 A;
        A.in = in; A.n = g->n; A.stride = g->stride; A.channels = g->channels; A.frames = frames;
        // enough CTAs to fill the machine, few enough flushes: ~8 frames per CTA unless the batch is small
        A.frames_per_cta = std::max(1, std::min(16, (int)((long long)frames * g->tracks / (148 * 4) + 1)));
        A.accum = accum; A.window = window; A.tw = twd;
        dim3 grid((unsigned)((frames + A.frames_per_cta - 1) / A.frames_per_cta), (unsigned)g->tracks);
        KernelScope ks(c, "spectral_envelope_rfft8192");
        envelope_kernel<<<grid, kEnvThreads, smem, c->stream>>>(A);
        MM_CUDA(cudaGetLastError());
    }
    envelope_final_kernel<<<(g->tracks * kEnvBins + 255) / 256, 256, 0, c->stream>>>(accum, g->tracks, frames, env_dev);
    MM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mm
