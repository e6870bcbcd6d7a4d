// Microbenchmark: cycles per step of the attack/release recursion walked by one lane (dependent FFMA,FFMA->FMNMX chain),
// from registers and from shared memory, with 1..8 such warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_regs(float* out, float at, float rt, int steps, long long* cyc) {
    float y = 1.f;
    const float omat = 1.f - at, omrt = 1.f - rt;
    float ua[16], ur[16];
    for (int i = 0; i < 16; ++i) { ua[i] = at * (0.01f * i + threadIdx.x); ur[i] = rt * (0.01f * i + threadIdx.x); }
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int s = 0; s < steps; s += 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y = fminf(fmaf(omat, y, ua[i]), fmaf(omrt, y, ur[i]));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = y; cyc[blockIdx.x] = t1 - t0; }
}
__global__ void chain_smem(float* out, float at, float rt, int steps, long long* cyc) {
    __shared__ float4 wa[512], wr[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { wa[i] = make_float4(at * i, at, at * 2, at * 3); wr[i] = make_float4(rt * i, rt, rt * 2, rt * 3); }
    __syncthreads();
    float y = 1.f;
    const float omat = 1.f - at, omrt = 1.f - rt;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int s = 0; s < steps; s += 2048) {
            float4 ca[4], cr[4], na[4], nr[4];
            for (int c = 0; c < 4; ++c) { ca[c] = wa[c]; cr[c] = wr[c]; }
#pragma unroll 1
            for (int h = 0; h < 128; ++h) {
                const int hn = h + 1 < 128 ? h + 1 : h;
#pragma unroll
                for (int c = 0; c < 4; ++c) { na[c] = wa[hn * 4 + c]; nr[c] = wr[hn * 4 + c]; }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float* a = reinterpret_cast<float*>(&ca[c]);
                    const float* b = reinterpret_cast<const float*>(&cr[c]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { y = fminf(fmaf(omat, y, a[k]), fmaf(omrt, y, b[k])); a[k] = y; }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) { wa[h * 4 + c] = ca[c]; ca[c] = na[c]; cr[c] = nr[c]; }
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[blockIdx.x] = y + wa[5].x; cyc[blockIdx.x] = t1 - t0; }
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 4096 * 8);
    const int steps = 1 << 20;
    for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
        for (int variant = 0; variant < 2; ++variant) {
            const int grid = 148 * per_sm;
            for (int rep = 0; rep < 2; ++rep) {
                if (variant == 0) chain_regs<<<grid, 64>>>(out, 0.3f, 0.01f, steps, cyc);
                else chain_smem<<<grid, 64>>>(out, 0.3f, 0.01f, steps, cyc);
                cudaDeviceSynchronize();
            }
            long long h[4096];
            cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%s  CTAs/SM=%d : %.2f cycles/step (max over CTAs)\n", variant ? "smem" : "regs", per_sm, (double)mx / steps);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
