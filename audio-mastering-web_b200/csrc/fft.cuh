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

// ---- in-register DFTs (forward sign); dft_pos<N>(k) is where output k ends up --------------------------------------
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 a02 = make_float2(a.x + c.x, a.y + c.y), s02 = make_float2(a.x - c.x, a.y - c.y);
    const float2 a13 = make_float2(b.x + d.x, b.y + d.y), s13 = make_float2(b.x - d.x, b.y - d.y);
    a = make_float2(a02.x + a13.x, a02.y + a13.y);
    b = make_float2(s02.x + s13.y, s02.y - s13.x);
    c = make_float2(a02.x - a13.x, a02.y - a13.y);
    d = make_float2(s02.x - s13.y, s02.y + s13.x);
}
__device__ __forceinline__ float2 w16(int m) {          // e^{-2 pi i m / 16}, m folded at compile time
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    switch (m) {
        case 0: return make_float2(1.0f, 0.0f);
        case 1: return make_float2(c1, -s1);
        case 2: return make_float2(h, -h);
        case 3: return make_float2(s1, -c1);
        case 4: return make_float2(0.0f, -1.0f);
        case 6: return make_float2(-h, -h);
        default: return make_float2(-c1, s1);           // m == 9
    }
}
template <int N> __device__ __forceinline__ void dft_reg(float2* v);
template <> __device__ __forceinline__ void dft_reg<16>(float2* v) {
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) dft4(v[n0], v[4 + n0], v[8 + n0], v[12 + n0]);
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
        for (int n0 = 1; n0 < 4; ++n0) v[4 * k1 + n0] = cmulf(v[4 * k1 + n0], w16(n0 * k1));
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}
template <> __device__ __forceinline__ void dft_reg<8>(float2* v) {
#pragma unroll
    for (int n0 = 0; n0 < 2; ++n0) dft4(v[n0], v[2 + n0], v[4 + n0], v[6 + n0]);
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1) v[2 * k1 + 1] = cmulf(v[2 * k1 + 1], w16(2 * k1));
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        const float2 a = v[2 * k1], b = v[2 * k1 + 1];
        v[2 * k1] = make_float2(a.x + b.x, a.y + b.y);
        v[2 * k1 + 1] = make_float2(a.x - b.x, a.y - b.y);
    }
}
template <int N> __device__ __forceinline__ constexpr int dft_pos(int k) { return N == 16 ? 4 * (k & 3) + (k >> 2) : 2 * (k & 3) + (k >> 2); }

}  // namespace mm
