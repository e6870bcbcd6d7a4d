// Library-level entry points of the C ABI (version, errors, device info).
#include "common.cuh"
#include "../../include/grafx_b200.h"

int g_gfx_last_cuda_error = 0;
unsigned long long g_gfx_launch_count = 0;

namespace gfx {
const DeviceInfo& device_info() {
    static DeviceInfo info[64];
    static bool have[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!have[dev]) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
            info[dev].sm_count = prop.multiProcessorCount;
            info[dev].max_smem_optin = (int)prop.sharedMemPerBlockOptin;
        } else {
            info[dev].sm_count = 148;
            info[dev].max_smem_optin = 227 * 1024;
        }
        have[dev] = true;
    }
    return info[dev];
}
}  // namespace gfx

namespace gfx {
// fp32 FMA issue-rate probe: 16 independent FFMA chains per thread (tools/fma_peak.cu in library form)
__global__ void __launch_bounds__(512) fma_probe_kernel(float* out, float a, float b, int iters) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;  // (keeps the chains alive; practically never true)
}
}  // namespace gfx

extern "C" {

long long gfx_fma_probe_f32(float* out, int iters, void* stream) {
    if (!out || iters <= 0) return GFX_ERR_INVALID;
    const int blocks = gfx::device_info().sm_count * 4;
    gfx::fma_probe_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(out, 1.0001f, 0.5f, iters);
    ++g_gfx_launch_count;
    if (cudaGetLastError() != cudaSuccess) return GFX_ERR_CUDA;
    return (long long)blocks * 512LL * 16LL * (long long)iters;  // FMAs executed by the launch
}

int gfx_version(void) { return 0 * 10000 + 1 * 100 + 0; }

int gfx_last_cuda_error(void) { return g_gfx_last_cuda_error; }

unsigned long long gfx_kernel_launch_count(void) { return g_gfx_launch_count; }

const char* gfx_error_string(int code) {
    switch (code) {
        case GFX_OK: return "ok";
        case GFX_ERR_INVALID: return "invalid argument";
        case GFX_ERR_WORKSPACE: return "workspace missing or too small";
        case GFX_ERR_CUDA: return "CUDA runtime error (see gfx_last_cuda_error)";
        case GFX_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}

int gfx_device_sm_count(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    return gfx::device_info().sm_count;
}

}  // extern "C"
