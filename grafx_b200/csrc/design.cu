// Parameter activations -> biquad coefficients, one thread per (row, section), one launch.
//
// Replaces the dozens of tiny elementwise launches of (reference, /root/reference/src/grafx/processors):
//   filter.py:144-154  BiquadFilter (stability-constrained direct coefficients)
//   filter.py:303-338  StateVariableFilter
//   filter.py:373-383, 592-604  w0 = pi*sigmoid, 1/q = exp, A = exp, alpha = sin(w0)/(2q)
//   filter.py:416-556  LowPass / HighPass / BandPass / BandReject / AllPass
//   filter.py:645-656, 687-705, 736-754  Peaking / LowShelf / HighShelf
//   eq.py:300-314      ParametricEqualizer band layout (low shelf, K-2 peaks, high shelf)
// Formulas follow the reference CODE (e.g. the low-pass numerator is (cos w0 - 1)/2 as shipped).
// Compiled with -fmad=false so every product/sum rounds like the separate torch ops do.
#include "common.cuh"

namespace gfx {

enum DesignFamily {
    DF_PEQ = 0,        // p0 = w0, p1 = q_inv, p2 = log_gain; flags & 1: shelving layout
    DF_PEAKING = 1,
    DF_LOWSHELF = 2,
    DF_HIGHSHELF = 3,
    DF_LOWPASS = 4,    // p0 = w0, p1 = q_inv
    DF_HIGHPASS = 5,
    DF_BANDPASS = 6,
    DF_BANDREJECT = 7,
    DF_ALLPASS = 8,
    DF_STABLE = 9,     // p0 = Bs [n,K,3], p1 = A1_pre, p2 = A2_pre, p3 = A0 (flags & 2: multiply by A0)
    DF_SVF = 10,       // p0 = twoR, p1 = G, p2 = c_hp, p3 = c_bp, p4 = c_lp
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ void eq_band(int kind, float c, float alpha, float A, float* b, float* a) {
    if (kind == DF_PEAKING) {
        const float aA = alpha * A, adA = alpha / A;
        b[0] = 1.f + aA; b[1] = -2.f * c; b[2] = 1.f - aA;
        a[0] = 1.f + adA; a[1] = -2.f * c; a[2] = 1.f - adA;
        return;
    }
    const float ap = A + 1.f, am = A - 1.f;
    const float apc = ap * c, amc = am * c;
    const float root = 2.f * sqrtf(A) * alpha;
    if (kind == DF_LOWSHELF) {
        b[0] = A * (ap - amc + root); b[1] = 2.f * A * (am - apc); b[2] = A * (ap - amc - root);
        a[0] = ap + amc + root;       a[1] = -2.f * (am + apc);    a[2] = ap + amc - root;
    } else {
        b[0] = A * (ap + amc + root); b[1] = -2.f * A * (am + apc); b[2] = A * (ap + amc - root);
        a[0] = ap - amc + root;       a[1] = 2.f * (am - apc);      a[2] = ap - amc - root;
    }
}

__global__ void __launch_bounds__(128) biquad_design_kernel(int family, const float* __restrict__ p0,
                                                            const float* __restrict__ p1, const float* __restrict__ p2,
                                                            const float* __restrict__ p3, const float* __restrict__ p4,
                                                            float* __restrict__ Bs, float* __restrict__ As, int n, int K,
                                                            int flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // section index = row * K + k
    if (i >= n * K) return;
    const int k = i % K;
    float b[3], a[3];
    if (family <= DF_ALLPASS) {
        const float w0 = 3.14159265358979323846f * sigmoidf_(p0[i]);
        const float qinv = expf(p1[i]);
        const float c = cosf(w0);
        const float alpha = sinf(w0) * qinv * 0.5f;
        if (family <= DF_HIGHSHELF) {
            const float A = expf(p2[i]);
            int kind = family;
            if (family == DF_PEQ) {
                kind = DF_PEAKING;
                if (flags & 1) kind = (k == 0) ? DF_LOWSHELF : ((k == K - 1) ? DF_HIGHSHELF : DF_PEAKING);
            }
            eq_band(kind, c, alpha, A, b, a);
        } else {
            a[0] = 1.f + alpha; a[1] = -2.f * c; a[2] = 1.f - alpha;
            switch (family) {
                case DF_LOWPASS: { const float h = (c - 1.f) / 2.f; b[0] = h; b[1] = c - 1.f; b[2] = h; break; }
                case DF_HIGHPASS: { const float h = (1.f + c) / 2.f; b[0] = h; b[1] = -(1.f + c); b[2] = h; break; }
                case DF_BANDPASS: b[0] = alpha; b[1] = 0.f; b[2] = -alpha; break;
                case DF_BANDREJECT: b[0] = 1.f; b[1] = -2.f * c; b[2] = 1.f; break;
                default: b[0] = a[2]; b[1] = a[1]; b[2] = a[0]; break;  // all-pass
            }
        }
    } else if (family == DF_STABLE) {
        const float a1 = 2.f * tanhf(p1[i]);
        const float mag = fabsf(a1);
        const float a2 = ((2.f - mag) * tanhf(p2[i]) + mag) / 2.f;
        a[0] = 1.f; a[1] = a1; a[2] = a2;
        if (flags & 2) { const float s = p3[i]; a[0] *= s; a[1] *= s; a[2] *= s; }
        b[0] = p0[3 * i] + 1.f; b[1] = p0[3 * i + 1]; b[2] = p0[3 * i + 2];
    } else {  // DF_SVF
        const float g = tanf(1.5707963267948966f * sigmoidf_(p1[i]));
        const float v = p0[i];
        const float sp = v > 20.f ? v : log1pf(expf(v));
        const float r2 = 1.4426950408889634f * sp + 1e-2f;
        const float g2 = g * g;
        const float hp = p2[i], bp = p3[i], lp = p4[i];
        b[0] = hp + bp * g + lp * g2; b[1] = -hp * 2.f + lp * 2.f * g2; b[2] = hp - bp * g + lp * g2;
        a[0] = 1.f + g2 + r2 * g;     a[1] = 2.f * g2 - 2.f;            a[2] = 1.f + g2 - r2 * g;
    }
    Bs[3 * i] = b[0]; Bs[3 * i + 1] = b[1]; Bs[3 * i + 2] = b[2];
    As[3 * i] = a[0]; As[3 * i + 1] = a[1]; As[3 * i + 2] = a[2];
}

}  // namespace gfx

extern "C" int gfx_biquad_design_f32(int family, const float* p0, const float* p1, const float* p2, const float* p3,
                                     const float* p4, float* Bs, float* As, int n_rows, int K, int flags,
                                     void* stream) {
    using namespace gfx;
    if (family < 0 || family > DF_SVF || !p0 || !p1 || !Bs || !As || n_rows <= 0 || K <= 0) return GFX_ERR_INVALID;
    if (family <= DF_HIGHSHELF && !p2) return GFX_ERR_INVALID;
    if (family == DF_STABLE && (!p2 || ((flags & 2) && !p3))) return GFX_ERR_INVALID;
    if (family == DF_SVF && (!p2 || !p3 || !p4)) return GFX_ERR_INVALID;
    const long long total = (long long)n_rows * K;
    biquad_design_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(family, p0, p1, p2, p3, p4,
                                                                                            Bs, As, n_rows, K, flags);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}
