// Shared-memory radix-4 Stockham FFT (M = 4^p complex points, float32) with a twiddle table, and the even/odd packing
// identities that turn a 2M-point real FFT / inverse real FFT into one M-point complex FFT.
#pragma once
#include <cuda_runtime.h>

namespace mm {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// tw[k] = e^{-2 pi i k / M}, k < M, filled from float64 sincospi
template <int M, int THREADS> __device__ __forceinline__ void fft_fill_twiddles(float2* tw) {
    for (int k = threadIdx.x; k < M; k += THREADS) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)M, &s, &c);
        tw[k] = make_float2((float)c, (float)s);
    }
}

// Forward transform (e^{-i...}) of the M points in A; B is scratch.  The caller synchronises after filling A; the result
// buffer (A or B) is returned and is complete (a __syncthreads() has been passed) on return.
template <int M, int THREADS> __device__ __forceinline__ float2* fft_r4_smem(float2* A, float2* B, const float2* tw) {
    float2* src = A;
    float2* dst = B;
#pragma unroll 1
    for (int Ns = 1; Ns < M; Ns <<= 2) {
        const int tstep = M / (4 * Ns);
        for (int j = threadIdx.x; j < M / 4; j += THREADS) {
            const int k = j & (Ns - 1);
            float2 v0 = src[j], v1 = src[j + M / 4], v2 = src[j + M / 2], v3 = src[j + 3 * M / 4];
            if (Ns > 1) {
                v1 = cmulf(v1, tw[k * tstep]);
                v2 = cmulf(v2, tw[2 * k * tstep]);
                v3 = cmulf(v3, tw[3 * k * tstep]);
            }
            const float2 a02 = make_float2(v0.x + v2.x, v0.y + v2.y), s02 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 a13 = make_float2(v1.x + v3.x, v1.y + v3.y), s13 = make_float2(v1.x - v3.x, v1.y - v3.y);
            const int j0 = ((j - k) << 2) + k;
            dst[j0] = make_float2(a02.x + a13.x, a02.y + a13.y);
            dst[j0 + Ns] = make_float2(s02.x + s13.y, s02.y - s13.x);
            dst[j0 + 2 * Ns] = make_float2(a02.x - a13.x, a02.y - a13.y);
            dst[j0 + 3 * Ns] = make_float2(s02.x - s13.y, s02.y + s13.x);
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    return src;
}

}  // namespace mm
