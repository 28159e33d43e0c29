// HBM bandwidth by read:write mix on one B200 -- the denominators behind the per-kernel roofline fractions
// (DESIGN.md): a streaming kernel that reads R float4 streams and writes W float4 streams of `n` bytes each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu && tools/membench
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int W> __global__ void mix_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 acc = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 v = __ldcs(in + (size_t)r * nvec + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
#pragma unroll
        for (int w = 0; w < W; ++w) __stcs(out + (size_t)w * nvec + i, acc);
        if (W == 0 && acc.x == 123.456f) out[0] = acc;
    }
}

template <int R, int W> void run(const float4* in, float4* out, size_t nvec, int sms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int bps : {4, 8, 16}) {
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            mix_kernel<R, W><<<sms * bps, 256>>>(in, out, nvec);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms < best) best = ms;
        }
    }
    const double bytes = (double)(R + W) * nvec * 16;
    printf("R%d:W%d  %8.1f GB/s total   (read %7.1f, write %7.1f)   %.3f ms\n", R, W, bytes / best / 1e6, R * nvec * 16.0 / best / 1e6,
           W * nvec * 16.0 / best / 1e6, best);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t nvec = (size_t)1 << 28 >> 2;      // 1 GiB per stream
    float4 *in, *out;
    cudaMalloc(&in, 6 * nvec * 16); cudaMalloc(&out, 6 * nvec * 16);
    cudaMemset(in, 0, 6 * nvec * 16); cudaMemset(out, 0, 6 * nvec * 16);
    printf("%s, %d SMs; 1 GiB per stream\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<1, 0>(in, out, nvec, sms); run<4, 0>(in, out, nvec, sms);
    run<0, 1>(in, out, nvec, sms); run<0, 4>(in, out, nvec, sms);
    run<1, 1>(in, out, nvec, sms); run<2, 2>(in, out, nvec, sms);
    run<1, 2>(in, out, nvec, sms); run<1, 4>(in, out, nvec, sms);
    run<2, 1>(in, out, nvec, sms); run<4, 1>(in, out, nvec, sms); run<5, 1>(in, out, nvec, sms);
    return 0;
}
