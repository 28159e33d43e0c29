// Pipe-rate microbenchmarks behind the sweep kernel's cost model: DFMA latency/throughput,
// F2F conversion throughput, integer float->double widening, on one B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void dfma_kernel(double* out, double a, double b, int iters) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// float -> double (F2F.F64.F32) + dependent double -> float (F2F.F32.F64): 2 conversions per step
template <int ILP> __global__ void f2f_kernel(float* out, int iters) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3f + i + 1.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v[i]));
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(v[i]) : "d"(d));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ double f2d_bits(float x) {
    const unsigned b = __float_as_uint(x);
    const unsigned m = b << 1;
    const unsigned hi = (m != 0u ? (m >> 4) + 0x38000000u : 0u) | (b & 0x80000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
}
// integer widening + one DADD consumer per step
template <int ILP, bool BITS> __global__ void widen_kernel(double* out, int iters) {
    float v[ILP];
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x * 1e-3f + i + 1.0f; acc[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            double d;
            if (BITS) d = f2d_bits(v[i]);
            else asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v[i]));
            acc[i] += d;
            v[i] = __int_as_float(__float_as_int(v[i]) + 1);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    printf("%s, %d SMs, max clock %.0f MHz (cycle figures assume the max clock)\n", p.name, sms, clk / 1e6);
    double* d; cudaMalloc(&d, 1 << 26);
    float* df = (float*)d;
    const int iters = 8192;
    const int cfg[5][2] = {{1, 32}, {1, 128}, {2, 256}, {4, 256}, {8, 256}};   // blocks/SM, threads
    for (int k = 0; k < 5; ++k) {
        const int grid = sms * cfg[k][0], thr = cfg[k][1], warps = cfg[k][0] * thr / 32;
        auto cyc = [&](float ms, int ops_per_iter) { return ms * 1e-3 * clk / ((double)iters * ops_per_iter * warps); };
        const float a1 = timeit([&] { dfma_kernel<1><<<grid, thr>>>(d, 1.0000001, 1e-9, iters); });
        const float a4 = timeit([&] { dfma_kernel<4><<<grid, thr>>>(d, 1.0000001, 1e-9, iters); });
        const float a8 = timeit([&] { dfma_kernel<8><<<grid, thr>>>(d, 1.0000001, 1e-9, iters); });
        const float c1 = timeit([&] { f2f_kernel<1><<<grid, thr>>>(df, iters); });
        const float c8 = timeit([&] { f2f_kernel<8><<<grid, thr>>>(df, iters); });
        const float w8 = timeit([&] { widen_kernel<8, false><<<grid, thr>>>(d, iters); });
        const float b8 = timeit([&] { widen_kernel<8, true><<<grid, thr>>>(d, iters); });
        printf("warps/SM %2d | DFMA chain %.2f cyc/step ; SM-cycles per warp-instr: DFMA ilp4 %.3f ilp8 %.3f | F2F pair chain %.2f cyc ; "
               "F2F ilp8 %.3f | cvt+DADD %.3f, bits+DADD %.3f per step\n",
               warps, a1 * 1e-3 * clk / iters, cyc(a4, 4), cyc(a8, 8), c1 * 1e-3 * clk / iters, cyc(c8, 16), cyc(w8, 8), cyc(b8, 8));
    }
    return 0;
}
