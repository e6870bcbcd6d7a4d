// Stereo <-> mid/side (reference: processors/core/midside.py:4-17) as one streaming pass.
//   lr_to_ms: m = (l + r) * mult, s = (l - r) * mult   (mult = 0.5 upstream)
//   ms_to_lr: l = m + s,          r = m - s            (mult = 1)
// Both are the same butterfly; x, y are [batch, 2, L].
#include "common.cuh"

namespace gfx {

__global__ void __launch_bounds__(256) midside_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      long long L, long long total_vec, float mult, int vec) {
    // one work unit = 4 consecutive samples of one batch item (both channels)
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const long long per_item = L / 4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
            const long long b = i / per_item, q = i - b * per_item;
            const float4* pa = reinterpret_cast<const float4*>(x + (b * 2) * L) + q;
            const float4* pb = reinterpret_cast<const float4*>(x + (b * 2 + 1) * L) + q;
            const float4 a = ldg_stream(pa), c = ldg_stream(pb);
            float4 u, v;
            u.x = (a.x + c.x) * mult; u.y = (a.y + c.y) * mult; u.z = (a.z + c.z) * mult; u.w = (a.w + c.w) * mult;
            v.x = (a.x - c.x) * mult; v.y = (a.y - c.y) * mult; v.z = (a.z - c.z) * mult; v.w = (a.w - c.w) * mult;
            stg_stream(reinterpret_cast<float4*>(y + (b * 2) * L) + q, u);
            stg_stream(reinterpret_cast<float4*>(y + (b * 2 + 1) * L) + q, v);
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
            const long long b = i / L, n = i - b * L;
            const float a = x[(b * 2) * L + n], c = x[(b * 2 + 1) * L + n];
            y[(b * 2) * L + n] = (a + c) * mult;
            y[(b * 2 + 1) * L + n] = (a - c) * mult;
        }
    }
}

}  // namespace gfx

extern "C" int gfx_midside_f32(const float* x, float* y, int batch, long long L, float mult, void* stream) {
    if (!x || !y || batch <= 0 || L <= 0) return GFX_ERR_INVALID;
    const int vec = (((uintptr_t)x | (uintptr_t)y) % 16 == 0) && (L % 4 == 0);
    const long long total = vec ? (long long)batch * (L / 4) : (long long)batch * L;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)gfx::device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    gfx::midside_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, L, total, mult, vec);
    GFX_CUDA_CHECK(cudaGetLastError());
    return GFX_OK;
}
