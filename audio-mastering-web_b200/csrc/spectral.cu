// Reference mastering (backend/app/pipeline.py:1527-1612): compute_spectral_envelope + apply_reference_match.
//
//   envelope : mono = mean over channels (float32); frames of 8192 samples, hop 2048, Hann window; RMS over the frames of
//              |rfft|  per bin (4097 bins).  Each 8192-point real FFT is one 4096-point complex Stockham FFT of the
//              even/odd-packed frame (shared memory, radix 4) plus the real-FFT untangling step; a CTA walks a run of
//              frames and keeps the 4097 float64 power sums in shared memory, one atomic flush per CTA.
//   matching : ratio curve -> Savitzky-Golay smoothing -> 8192-tap FIR are DESIGN steps on 4097 numbers (host, numpy/scipy
//              in mm_b200/pipeline.py, as the reference does); the FIR itself runs through fir_same_kernel (followers.cu).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "context.h"
#include "stages_internal.h"

namespace mm {

constexpr int kEnvN = 8192;             // n_fft
constexpr int kEnvM = kEnvN / 2;        // complex FFT size
constexpr int kEnvHop = kEnvN / 4;
constexpr int kEnvThreads = 256;
constexpr int kEnvBins = kEnvN / 2 + 1;

// radix-4 Stockham, decimation in frequency, 4096 = 4^6 points, forward (e^{-i...}); returns the buffer holding the result
__device__ float2* fft4096(float2* A, float2* B) {
    int Ns = 1;
    float2* src = A;
    float2* dst = B;
#pragma unroll 1
    for (int pass = 0; pass < 6; ++pass) {
        for (int j = threadIdx.x; j < kEnvM / 4; j += kEnvThreads) {
            const int k = j & (Ns - 1);
            float2 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = src[j + r * (kEnvM / 4)];
            const float ang = -2.0f * (float)k / (float)(4 * Ns);   // in units of pi
#pragma unroll
            for (int r = 1; r < 4; ++r) {
                float s, cth;
                sincospif(ang * (float)r, &s, &cth);
                const float2 t = v[r];
                v[r] = make_float2(t.x * cth - t.y * s, t.x * s + t.y * cth);
            }
            const float2 a02 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
            const float2 s02 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
            const float2 a13 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
            const float2 s13 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
            const int j0 = ((j - k) << 2) + k;
            dst[j0] = make_float2(a02.x + a13.x, a02.y + a13.y);
            dst[j0 + Ns] = make_float2(s02.x + s13.y, s02.y - s13.x);
            dst[j0 + 2 * Ns] = make_float2(a02.x - a13.x, a02.y - a13.y);
            dst[j0 + 3 * Ns] = make_float2(s02.x - s13.y, s02.y + s13.x);
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
        Ns <<= 2;
    }
    return src;
}

struct EnvArgs2 {
    const float* in;
    long long n, stride;
    int channels, frames, frames_per_cta;
    double* accum;          // [tracks][kEnvBins]
};

__global__ void __launch_bounds__(kEnvThreads) envelope_kernel(const EnvArgs2 P) {
    extern __shared__ __align__(16) unsigned char esm[];
    float2* A = reinterpret_cast<float2*>(esm);
    float2* B = A + kEnvM;
    double* acc = reinterpret_cast<double*>(B + kEnvM);
    const int track = blockIdx.y;
    const int f0 = blockIdx.x * P.frames_per_cta, f1 = min(P.frames, f0 + P.frames_per_cta);
    for (int k = threadIdx.x; k < kEnvBins; k += kEnvThreads) acc[k] = 0.0;
    const float* r0 = P.in + (size_t)(track * P.channels) * (size_t)P.stride + kLead;
    const float* r1 = r0 + (P.channels > 1 ? (size_t)P.stride : 0);
    for (int f = f0; f < f1; ++f) {
        const long long start = (long long)f * kEnvHop;
        __syncthreads();
        for (int k = threadIdx.x; k < kEnvM; k += kEnvThreads) {
            float xs[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * k + e;
                // np.mean(audio, axis=1).astype(float32): float32 sum of the two channels, divided by 2
                const float m = P.channels > 1 ? __fmul_rn(__fadd_rn(r0[start + j], r1[start + j]), 0.5f) : r0[start + j];
                // np.hanning(M).astype(float32): 0.5 - 0.5 cos(2 pi j / (M - 1)) in float64, cast
                const float w = (float)(0.5 - 0.5 * cospi(2.0 * (double)j / (double)(kEnvN - 1)));
                xs[e] = __fmul_rn(m, w);
            }
            A[k] = make_float2(xs[0], xs[1]);
        }
        __syncthreads();
        const float2* Z = fft4096(A, B);
        // X[k] = (Z[k] + conj Z[M-k]) / 2 - (i / 2) e^{-2 pi i k / N} (Z[k] - conj Z[M-k]),  k = 0 .. M  (Z[M] = Z[0])
        for (int k = threadIdx.x; k <= kEnvM; k += kEnvThreads) {
            const float2 zk = Z[k & (kEnvM - 1)], zm = Z[(kEnvM - k) & (kEnvM - 1)];
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
    for (int k = threadIdx.x; k < kEnvBins; k += kEnvThreads)
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
        const size_t smem = 2 * kEnvM * sizeof(float2) + kEnvBins * sizeof(double) + 16;
        static bool attr = false;
        if (!attr) {
            MM_CUDA(cudaFuncSetAttribute(envelope_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        EnvArgs2 A;
        A.in = in; A.n = g->n; A.stride = g->stride; A.channels = g->channels; A.frames = frames;
        // enough CTAs to fill the machine, few enough flushes: ~8 frames per CTA unless the batch is small
        A.frames_per_cta = std::max(1, std::min(16, (int)((long long)frames * g->tracks / (148 * 4) + 1)));
        A.accum = accum;
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
