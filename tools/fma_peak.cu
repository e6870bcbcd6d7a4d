// Micro-benchmark: fp32 FMA issue rate on B200, scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_peak tools/fma_peak.cu && /tmp/fma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_ffma(float* out, float a, float b, int iters) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
    float2 v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chain variants (latency)
__global__ void k_chain(float* out, float a, float b, int iters, int packed) {
    float2 v = make_float2(threadIdx.x, 1.f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    if (packed) for (int it = 0; it < iters; ++it) v = __ffma2_rn(v, a2, b2);
    else for (int it = 0; it < iters; ++it) v.x = fmaf(v.x, a, b);
    out[blockIdx.x * blockDim.x + threadIdx.x] = v.x + v.y;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
    const int iters = 4096;
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        int threads = 128 * warps_per_smsp;   // 4 SMSPs x warps
        int blocks = sms;
        if (threads > 1024) { threads = 512; blocks = sms * 2; }
        float m1 = timeit([&] { k_ffma<16><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
        float m2 = timeit([&] { k_ffma2<16><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
        double fma1 = (double)blocks * threads * 16.0 * iters / (m1 * 1e-3);
        double fma2 = (double)blocks * threads * 32.0 * iters / (m2 * 1e-3);
        printf("warps/SMSP %d: FFMA %.2f TFMA/s (%.1f fma/clk/SM @1.965GHz)   FFMA2 %.2f TFMA/s (%.1f fma/clk/SM)\n",
               warps_per_smsp, fma1 / 1e12, fma1 / sms / 1.965e9, fma2 / 1e12, fma2 / sms / 1.965e9);
    }
    float c1 = timeit([&] { k_chain<<<sms, 32>>>(out, 1.0001f, 0.5f, 1 << 16, 0); });
    float c2 = timeit([&] { k_chain<<<sms, 32>>>(out, 1.0001f, 0.5f, 1 << 16, 1); });
    printf("dependent chain: FFMA %.2f ns/op (%.2f cyc @1.965)  FFMA2 %.2f ns/op (%.2f cyc)\n", c1 * 1e6 / 65536,
           c1 * 1e6 / 65536 * 1.965, c2 * 1e6 / 65536, c2 * 1e6 / 65536 * 1.965);
    return 0;
}
