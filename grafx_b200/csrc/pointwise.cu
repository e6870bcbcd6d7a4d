// Memoryless (sample-wise) processors as one streaming pass: x read once, y written once.
//
// Replaces (reference, /root/reference/src/grafx/processors):
//   stereo.py:31-38     StereoGain.forward          y[b,c] = x[b,c] * exp(log_gain[b,c])
//   stereo.py:71-84     SideGainImager.forward      side channel of the mid/side pair scaled by exp(log_gain[b])
//   nonlinear.py:64-89  TanhDistortion.forward      (tanh((x - dc) g_pre + bias) - tanh(bias)) g_post
//   nonlinear.py:159-205 PiecewiseTanhDistortion    tanh in the middle, scaled/offset tanh lobes beyond +-thresholds
//   nonlinear.py:268-285 PowerDistortion            sum_k tanh(w_k) [tanh] (x g_pre)^k, k = 0..order-1
//   nonlinear.py:349-384 ChebyshevDistortion        sum_k tanh(w_k) [tanh] T_k(x g_pre)
//   container.py:203-216 ParallelMix accumulation   y (+)= w[b] * x
// plus the per-row mean (`remove_dc`, nonlinear.py:65-66) and mean square (rms_difference, core/utils.py:7-11, of
// GainStagingRegularization) as a deterministic one-CTA-per-row reduction.
// Each thread handles 4 consecutive samples (float4) of one row; per-row parameters are read once per thread
// block iteration (rows are long: L >> block).  Upstream these are 5-20 full-length elementwise launches each.
#include "common.cuh"

namespace gfx {

constexpr int PW_MAX_ORDER = 32;

struct PwParams {
    int op, flags, order, channels;
    const float* x;
    float* y;
    const float* p0;  // meaning per op, see gfx_pointwise_f32
    const float* p1;
    const float* p2;
    const float* p3;
    const float* dc;  // [rows] mean to subtract first, or null
    long long L;
    int rows;
};

// flags
constexpr int PW_INVERSE_POST = 1;  // post gain = 1 / pre gain
constexpr int PW_USE_TANH = 2;      // tanh on every basis function (power / chebyshev)
constexpr int PW_ACCUMULATE = 4;    // scale-add: add to y

struct RowConst {
    float pre, post, bias, tb;  // gains, bias, tanh(bias)
    float kp, kn, gp, gn, ap, an, bp, bn;  // piecewise tanh
    float dc;
};

__device__ __forceinline__ float pw_apply(int op, int flags, int order, float v, const RowConst& c, const float* w) {
    switch (op) {
        case 2: {  // tanh distortion
            return (tanhf(fmaf(v - c.dc, c.pre, c.bias)) - c.tb) * c.post;
        }
        case 3: {  // piecewise tanh
            const float u = (v - c.dc) * c.pre;
            float o;
            if (u > c.kp) o = fmaf(c.ap, tanhf(c.gp * (u - c.kp)), c.bp);
            else if (u < -c.kn) o = fmaf(c.an, tanhf(c.gn * (u + c.kn)), c.bn);
            else o = tanhf(u);
            return o * c.post;
        }
        case 4: {  // power series
            const float u = (v - c.dc) * c.pre;
            float pw = 1.f, acc = 0.f;
            for (int k = 0; k < order; ++k) {
                acc = fmaf(w[k], (flags & PW_USE_TANH) ? tanhf(pw) : pw, acc);
                pw *= u;
            }
            return acc;
        }
        case 5: {  // chebyshev series
            const float u = (v - c.dc) * c.pre;
            float t0 = 1.f, t1 = u, acc;
            const bool th = flags & PW_USE_TANH;
            acc = w[0] * (th ? tanhf(t0) : t0);
            if (order > 1) acc = fmaf(w[1], th ? tanhf(t1) : t1, acc);
            for (int k = 2; k < order; ++k) {
                const float t2 = 2.f * u * t1 - t0;
                acc = fmaf(w[k], th ? tanhf(t2) : t2, acc);
                t0 = t1; t1 = t2;
            }
            return acc;
        }
        default: return v;
    }
}

__global__ void __launch_bounds__(256) pointwise_kernel(const PwParams p) {
    // grid: x = chunks of a row, y = rows (b * channels + c); side-gain: y = batch items
    const int row = blockIdx.y;
    const int b = p.op == 1 ? row : row / p.channels;
    __shared__ float w[PW_MAX_ORDER];
    RowConst c;
    c.dc = p.dc ? p.dc[row] : 0.f;
    c.pre = 1.f; c.post = 1.f; c.bias = 0.f; c.tb = 0.f;
    c.kp = c.kn = c.gp = c.gn = c.ap = c.an = c.bp = c.bn = 0.f;
    if (p.op == 0) {
        c.pre = expf(p.p0[row]);  // log_gain [B, C]
    } else if (p.op == 1) {
        c.pre = expf(p.p0[b]);    // log_gain [B]
    } else if (p.op == 6) {
        c.pre = p.p0[b];          // weight [B]
    } else {
        const float* lpre = p.op == 2 ? p.p0 : (p.op == 3 ? p.p2 : p.p1);
        const float* lpost = p.op == 2 ? p.p1 : (p.op == 3 ? p.p3 : nullptr);
        if (lpre) c.pre = expf(lpre[b]);
        if (p.flags & PW_INVERSE_POST) c.post = 1.f / c.pre;
        else if (lpost) c.post = expf(lpost[b]);
        if (p.op == 2 && p.p2) { c.bias = p.p2[b]; c.tb = tanhf(c.bias); }
        if (p.op == 3) {
            // hardness = exp(log_hardness) = (gp, gn); threshold = sigmoid(z_threshold) = (kn, kp)  (nonlinear.py:190-193)
            c.gp = expf(p.p0[2 * b]); c.gn = expf(p.p0[2 * b + 1]);
            c.kn = 1.f / (1.f + expf(-p.p1[2 * b])); c.kp = 1.f / (1.f + expf(-p.p1[2 * b + 1]));
            c.bp = tanhf(c.kp); c.bn = -tanhf(c.kn);
            c.ap = (1.f - c.bp) / c.gp; c.an = (1.f + c.bn) / c.gn;
        }
        if (p.op == 4 || p.op == 5) {
            for (int k = threadIdx.x; k < p.order; k += blockDim.x) w[k] = tanhf(p.p0[(size_t)b * p.order + k]);
            __syncthreads();
        }
    }
    const long long L = p.L;
    const bool vec = (((uintptr_t)p.x | (uintptr_t)p.y) % 16 == 0) && (L % 4 == 0);
    if (p.op == 1) {
        // side gain on the (left, right) pair of item b
        const float* xl = p.x + (size_t)(2 * b) * L;
        const float* xr = xl + L;
        float* yl = p.y + (size_t)(2 * b) * L;
        float* yr = yl + L;
        const float g = c.pre;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
            const float l = xl[i], r = xr[i];
            const float mid = l + r, side = g * (l - r);
            yl[i] = (mid + side) / 2.f;
            yr[i] = (mid - side) / 2.f;
        }
        return;
    }
    const float* xr = p.x + (size_t)row * L;
    float* yr = p.y + (size_t)row * L;
    if (vec) {
        const long long n4 = L / 4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            float4 v = ldg_stream(reinterpret_cast<const float4*>(xr) + i);
            if (p.op == 0) {
                v.x *= c.pre; v.y *= c.pre; v.z *= c.pre; v.w *= c.pre;
            } else if (p.op == 6) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.flags & PW_ACCUMULATE) a = reinterpret_cast<const float4*>(yr)[i];
                v.x = fmaf(c.pre, v.x, a.x); v.y = fmaf(c.pre, v.y, a.y); v.z = fmaf(c.pre, v.z, a.z); v.w = fmaf(c.pre, v.w, a.w);
            } else {
                v.x = pw_apply(p.op, p.flags, p.order, v.x, c, w);
                v.y = pw_apply(p.op, p.flags, p.order, v.y, c, w);
                v.z = pw_apply(p.op, p.flags, p.order, v.z, c, w);
                v.w = pw_apply(p.op, p.flags, p.order, v.w, c, w);
            }
            reinterpret_cast<float4*>(yr)[i] = v;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
            float v = xr[i];
            if (p.op == 0) v *= c.pre;
            else if (p.op == 6) v = fmaf(c.pre, v, (p.flags & PW_ACCUMULATE) ? yr[i] : 0.f);
            else v = pw_apply(p.op, p.flags, p.order, v, c, w);
            yr[i] = v;
        }
    }
}

// mean (or mean square) over time of every row: one CTA per row, fixed summation order (deterministic)
template <bool SQUARE>
__global__ void __launch_bounds__(256) row_mean_kernel(const float* __restrict__ x, float* __restrict__ mean, long long L) {
    const float* xr = x + (size_t)blockIdx.x * L;
    double acc = 0.0;  // (double partials: the torch reference reduces pairwise in fp32; this stays within 1e-7 of it)
    if ((((uintptr_t)xr) % 16 == 0) && (L % 4 == 0)) {
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        for (long long i = threadIdx.x; i < L / 4; i += 256) {
            const float4 v = ldg_stream(x4 + i);
            if (SQUARE) acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
            else acc += (double)(v.x + v.y) + (double)(v.z + v.w);
        }
    } else {
        for (long long i = threadIdx.x; i < L; i += 256) acc += SQUARE ? (double)(xr[i] * xr[i]) : (double)xr[i];
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[blockIdx.x] = (float)(red[0] / (double)L);
}

}  // namespace gfx

extern "C" {

int gfx_row_mean_f32(const float* x, float* mean, int rows, long long L, void* stream) {
    if (!x || !mean || rows <= 0 || L <= 0) return GFX_ERR_INVALID;
    gfx::row_mean_kernel<false><<<rows, 256, 0, (cudaStream_t)stream>>>(x, mean, L);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

int gfx_row_mean_square_f32(const float* x, float* mean, int rows, long long L, void* stream) {
    if (!x || !mean || rows <= 0 || L <= 0) return GFX_ERR_INVALID;
    gfx::row_mean_kernel<true><<<rows, 256, 0, (cudaStream_t)stream>>>(x, mean, L);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

int gfx_pointwise_f32(int op, const float* x, float* y, int batch, int channels, long long L, const float* p0,
                      const float* p1, const float* p2, const float* p3, const float* dc, int order, int flags,
                      void* stream) {
    using namespace gfx;
    if (!x || !y || batch <= 0 || channels <= 0 || L <= 0) return GFX_ERR_INVALID;
    if (op < 0 || op > 6) return GFX_ERR_INVALID;
    if ((op == 0 || op == 1 || op == 6 || op == 4 || op == 5) && !p0) return GFX_ERR_INVALID;
    if (op == 3 && (!p0 || !p1)) return GFX_ERR_INVALID;
    if (op == 1 && channels != 2) return GFX_ERR_INVALID;
    if ((op == 4 || op == 5) && (order < 1 || order > PW_MAX_ORDER)) return GFX_ERR_UNSUPPORTED;
    PwParams p;
    p.op = op; p.flags = flags; p.order = order; p.channels = channels;
    p.x = x; p.y = y; p.p0 = p0; p.p1 = p1; p.p2 = p2; p.p3 = p3; p.dc = dc; p.L = L;
    p.rows = batch * channels;
    const int grid_y = op == 1 ? batch : p.rows;
    if (grid_y > 65535) return GFX_ERR_UNSUPPORTED;
    long long per_row = (L / 4 + 255) / 256;
    const long long want = (long long)device_info().sm_count * 8 / grid_y + 1;
    if (per_row > want) per_row = want;
    if (per_row < 1) per_row = 1;
    pointwise_kernel<<<dim3((unsigned)per_row, (unsigned)grid_y), 256, 0, (cudaStream_t)stream>>>(p);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

}  // extern "C"
