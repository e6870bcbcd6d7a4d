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
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

// ------------------------------------------------------------------------------------------------
// Dry/wet mix (reference: processors/container.py:62-67):  y = w * wet + (1 - w) * dry, w per item.
namespace gfx {
__global__ void __launch_bounds__(256) drywet_kernel(const float* __restrict__ dry, const float* __restrict__ wet,
                                                     const float* __restrict__ w, float* __restrict__ y,
                                                     long long inner, long long total, int vec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (vec) {
            const long long b = (i * 4) / inner;
            const float a = w[b], c = 1.f - a;
            const float4 d = ldg_stream(reinterpret_cast<const float4*>(dry) + i);
            const float4 e = ldg_stream(reinterpret_cast<const float4*>(wet) + i);
            float4 o;
            o.x = a * e.x + c * d.x; o.y = a * e.y + c * d.y; o.z = a * e.z + c * d.z; o.w = a * e.w + c * d.w;
            stg_stream(reinterpret_cast<float4*>(y) + i, o);
        } else {
            const float a = w[i / inner];
            y[i] = a * wet[i] + (1.f - a) * dry[i];
        }
    }
}

// Node-axis aggregation of the render loop (reference: render/core.py:101-112 aggregate_tensor,
// "sum" and "scatter").  src is a [batch, n_src, inner] view (strided), dst a [batch, n_dst, inner]
// view; dst[b, j] = sum over sources i with index[i] == j  (index == NULL: every source -> j = 0).
__global__ void __launch_bounds__(256) node_sum_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                       const int* __restrict__ index, int batch, int n_src, int n_dst,
                                                       long long inner, long long src_bs, long long src_ns,
                                                       long long dst_bs, long long dst_ns, int vec) {
    const long long per = vec ? inner / 4 : inner;
    const long long total = (long long)batch * n_dst * per;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long q = i % per;
        const long long bj = i / per;
        const int j = (int)(bj % n_dst);
        const long long b = bj / n_dst;
        const float* sb = src + b * src_bs;
        float* out = dst + b * dst_bs + (long long)j * dst_ns;
        if (vec) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int s = 0;
            if (!index) {
                // plain bus sum: eight independent 16-byte loads in flight per thread (same summation order)
                for (; s + 8 <= n_src; s += 8) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = ldg_stream(reinterpret_cast<const float4*>(sb + (long long)(s + u) * src_ns) + q);
#pragma unroll
                    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
                }
            }
            for (; s < n_src; ++s) {
                if (index && index[s] != j) continue;
                const float4 v = ldg_stream(reinterpret_cast<const float4*>(sb + (long long)s * src_ns) + q);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            stg_stream(reinterpret_cast<float4*>(out) + q, acc);
        } else {
            float acc = 0.f;
            for (int s = 0; s < n_src; ++s) {
                if (index && index[s] != j) continue;
                acc += sb[(long long)s * src_ns + q];
            }
            out[q] = acc;
        }
    }
}
// Strided block copy of a [batch, nodes, inner] view (render/core.py:6-33 create_signal_buffer writes the
// sources into the buffer; here the buffer is node-major, so a batched input [B, V0, C, L] is transposed on
// the way in).  Four independent 16-byte loads per thread keep enough bytes in flight for HBM.
__global__ void __launch_bounds__(256) node_copy_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int nodes, long long inner, long long src_bs,
                                                        long long src_ns, long long dst_bs, long long dst_ns, int vec,
                                                        long long total) {
    const long long per = vec ? inner / 4 : inner;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto locate = [&](long long e, const float*& s, float*& d, long long& q) {
        q = e % per;
        const long long bj = e / per;
        const long long j = bj % nodes, b = bj / nodes;
        s = src + b * src_bs + j * src_ns;
        d = dst + b * dst_bs + j * dst_ns;
    };
    if (vec) {
        constexpr int NU = 4;  // independent 16-byte loads in flight per thread (8 measured: no change, 5.5 TB/s either way)
        for (; i + (NU - 1) * stride < total; i += NU * stride) {
            const float* s[NU]; float* d[NU]; long long q[NU]; float4 v[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) locate(i + u * stride, s[u], d[u], q[u]);
#pragma unroll
            for (int u = 0; u < NU; ++u) v[u] = ldg_stream(reinterpret_cast<const float4*>(s[u]) + q[u]);
#pragma unroll
            for (int u = 0; u < NU; ++u) stg_stream(reinterpret_cast<float4*>(d[u]) + q[u], v[u]);
        }
        for (; i < total; i += stride) {
            const float* s; float* d; long long q;
            locate(i, s, d, q);
            stg_stream(reinterpret_cast<float4*>(d) + q, ldg_stream(reinterpret_cast<const float4*>(s) + q));
        }
    } else {
        for (; i < total; i += stride) {
            const float* s; float* d; long long q;
            locate(i, s, d, q);
            d[q] = s[q];
        }
    }
}
// Lagged inner products of row signals, the coefficient gradients of one biquad section (adjoint of
// IIRFilter._process_lfilter, processors/core/iir.py:154-196; upstream gets them from torchaudio's lfilter autograd):
//   out0[row][j] = sum_n u[n] * s0[n - j],  j = 0, 1, 2   (s[n] = 0 for n < 0), and the same for s1 -> out1 if given,
// with u optionally stored time-reversed (u[n] read at L-1-n: the adjoint recursion runs on the reversed gradient).
// One CTA per row; four samples per thread and step (16-byte loads); products of four samples are summed in fp32, the
// running sums in double; fixed summation order (deterministic).
constexpr int LAG_NT = 512;

struct Lag3 {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    __device__ __forceinline__ void add4(const float4& u, const float4& s, float sm1, float sm2) {
        a0 += (double)(fmaf(u.x, s.x, u.y * s.y) + fmaf(u.z, s.z, u.w * s.w));
        a1 += (double)(fmaf(u.x, sm1, u.y * s.x) + fmaf(u.z, s.y, u.w * s.z));
        a2 += (double)(fmaf(u.x, sm2, u.y * sm1) + fmaf(u.z, s.x, u.w * s.y));
    }
    __device__ __forceinline__ void add1(float u, float s, float sm1, float sm2) {
        a0 += (double)u * (double)s; a1 += (double)u * (double)sm1; a2 += (double)u * (double)sm2;
    }
};

__global__ void __launch_bounds__(LAG_NT) lag_dots_kernel(const float* __restrict__ u, const float* __restrict__ s0,
                                                          const float* __restrict__ s1, float* __restrict__ out0,
                                                          float* __restrict__ out1, long long L, int u_reversed, int vec) {
    const float* ur = u + (size_t)blockIdx.x * L;
    const float* r0 = s0 + (size_t)blockIdx.x * L;
    const float* r1 = s1 ? s1 + (size_t)blockIdx.x * L : nullptr;
    Lag3 A, B;
    if (vec) {
        const long long n4 = L / 4;
        for (long long i = threadIdx.x; i < n4; i += LAG_NT) {
            float4 uq;
            if (u_reversed) {
                const float4 t = ldg_stream(reinterpret_cast<const float4*>(ur) + (n4 - 1 - i));
                uq = make_float4(t.w, t.z, t.y, t.x);
            } else {
                uq = ldg_stream(reinterpret_cast<const float4*>(ur) + i);
            }
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a = __ldg(reinterpret_cast<const float4*>(r0) + i);
            const float4 ap = i > 0 ? __ldg(reinterpret_cast<const float4*>(r0) + i - 1) : z;
            A.add4(uq, a, ap.w, ap.z);
            if (r1) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(r1) + i);
                const float4 bp = i > 0 ? __ldg(reinterpret_cast<const float4*>(r1) + i - 1) : z;
                B.add4(uq, b, bp.w, bp.z);
            }
        }
    } else {
        for (long long n = threadIdx.x; n < L; n += LAG_NT) {
            const float uv = u_reversed ? ur[L - 1 - n] : ur[n];
            A.add1(uv, r0[n], n >= 1 ? r0[n - 1] : 0.f, n >= 2 ? r0[n - 2] : 0.f);
            if (r1) B.add1(uv, r1[n], n >= 1 ? r1[n - 1] : 0.f, n >= 2 ? r1[n - 2] : 0.f);
        }
    }
    __shared__ double red[6][LAG_NT / 32];
    double v[6] = {A.a0, A.a1, A.a2, B.a0, B.a1, B.a2};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double sum = 0.0;
        for (int w = 0; w < LAG_NT / 32; ++w) sum += red[threadIdx.x][w];
        if (threadIdx.x < 3) out0[(size_t)blockIdx.x * 3 + threadIdx.x] = (float)sum;
        else if (out1) out1[(size_t)blockIdx.x * 3 + threadIdx.x - 3] = (float)sum;
    }
}
}  // namespace gfx

extern "C" int gfx_drywet_f32(const float* dry, const float* wet, const float* weight, float* y, int batch,
                              long long inner, void* stream) {
    if (!dry || !wet || !weight || !y || batch <= 0 || inner <= 0) return GFX_ERR_INVALID;
    const int vec = (((uintptr_t)dry | (uintptr_t)wet | (uintptr_t)y) % 16 == 0) && (inner % 4 == 0);
    const long long total = vec ? (long long)batch * inner / 4 : (long long)batch * inner;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)gfx::device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    gfx::drywet_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dry, wet, weight, y, inner, total, vec);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

extern "C" int gfx_node_sum_f32(const float* src, float* dst, const int* index, int batch, int n_src, int n_dst,
                                long long inner, long long src_batch_stride, long long src_node_stride,
                                long long dst_batch_stride, long long dst_node_stride, void* stream) {
    if (!src || !dst || batch <= 0 || n_src <= 0 || n_dst <= 0 || inner <= 0) return GFX_ERR_INVALID;
    const int vec = (((uintptr_t)src | (uintptr_t)dst) % 16 == 0) && (inner % 4 == 0) && (src_batch_stride % 4 == 0) &&
                    (src_node_stride % 4 == 0) && (dst_batch_stride % 4 == 0) && (dst_node_stride % 4 == 0);
    const long long total = (long long)batch * n_dst * (vec ? inner / 4 : inner);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)gfx::device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    gfx::node_sum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        src, dst, index, batch, n_src, n_dst, inner, src_batch_stride, src_node_stride, dst_batch_stride,
        dst_node_stride, vec);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

extern "C" int gfx_node_copy_f32(const float* src, float* dst, int batch, int nodes, long long inner,
                                 long long src_batch_stride, long long src_node_stride, long long dst_batch_stride,
                                 long long dst_node_stride, void* stream) {
    if (!src || !dst || batch <= 0 || nodes <= 0 || inner <= 0) return GFX_ERR_INVALID;
    const int vec = (((uintptr_t)src | (uintptr_t)dst) % 16 == 0) && (inner % 4 == 0) && (src_batch_stride % 4 == 0) &&
                    (src_node_stride % 4 == 0) && (dst_batch_stride % 4 == 0) && (dst_node_stride % 4 == 0);
    const long long total = (long long)batch * nodes * (vec ? inner / 4 : inner);
    long long blocks = (total + 1023) / 1024;
    const long long cap = (long long)gfx::device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    gfx::node_copy_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        src, dst, nodes, inner, src_batch_stride, src_node_stride, dst_batch_stride, dst_node_stride, vec, total);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

extern "C" int gfx_lag_dots_f32(const float* u, const float* s0, const float* s1, float* out0, float* out1, int rows,
                                long long L, int u_reversed, void* stream) {
    if (!u || !s0 || !out0 || rows <= 0 || L <= 0 || ((s1 == nullptr) != (out1 == nullptr))) return GFX_ERR_INVALID;
    const int vec = (L % 4 == 0) && (((uintptr_t)u | (uintptr_t)s0 | (uintptr_t)s1) % 16 == 0);
    gfx::lag_dots_kernel<<<(unsigned)rows, gfx::LAG_NT, 0, (cudaStream_t)stream>>>(u, s0, s1, out0, out1, L, u_reversed, vec);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}
