// Causal / zero-phase FIR convolution by overlap-save with an in-shared-memory FFT.
//
// Replaces (reference, /root/reference/src/grafx/processors/core/convolution.py:119-134):
//   convolve(): F.pad x2 -> torch.fft.rfft x2 (cuFFT, length Lx+Lh-1, Bluestein for the odd
//   lengths every BASELINE config produces) -> complex multiply -> irfft -> slice; >= 6 passes over
//   padded data in HBM.  Users: FIRFilter (filter.py:65-77), the fsm IIR backend (core/iir.py:147-152),
//   TruncatedOnePoleIIRFilter (core/envelope.py:47-49; handled by a recursion in csrc/dynamics.cu),
//   STFTMaskedNoiseReverb (reverb.py:215-228), ZeroPhaseFIREqualizer (eq.py:70-72, 208-214).
// Semantics: y[n] = sum_k h[k] x[n + shift - k], n < L; shift = 0 (causal) or N//2 (zerophase);
// rows of x and h broadcast over the channel axis like tensor broadcasting does upstream.
//
// On chip: a real block of F = 2n samples is packed into n complex points (even/odd), transformed
// by an in-place radix-4 decimation-in-frequency FFT in shared memory (split re/im arrays, output
// left in base-4 digit-reversed order), untangled / multiplied by the filter spectrum / re-tangled
// pairwise in that order, and brought back by the matching decimation-in-time inverse -- no
// bit-reversal pass, no complex intermediate in HBM.  n is 1024, 4096 or 16384.
//   short filters (N <= 16384): one spectrum per filter row, blocks of F-N+1 new outputs;
//   long filters (reverb IRs):  uniformly partitioned overlap-save, partition = n taps, X and H
//       partition spectra kept in a workspace sized to stay L2-resident, frequency-domain
//       accumulation, one inverse FFT per block.
// Twiddles e^{-i pi k / n} come from a table built once per n (gfx_fft_plan_init) in double.
#include "common.cuh"

namespace gfx {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

template <int LOG4>
__device__ __forceinline__ int drev4(int p) {
    unsigned r = __brev((unsigned)p) >> (32 - 2 * LOG4);
    return (int)(((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u));
}

// ------------------------------------------------------------------ FFT passes (in place, smem)
template <int N, int NT, int M>
__device__ __forceinline__ void dif_pass(float* re, float* im, const float2* __restrict__ tw) {
    constexpr int ST = M / 4;
    constexpr int TWS = 2 * (N / M);  // omega_M^j = tw[j * TWS],  tw[k] = exp(-i pi k / N)
    for (int b = threadIdx.x; b < N / 4; b += NT) {
        const int j = b & (ST - 1);
        const int i0 = ((b - j) << 2) + j;
        float ar[4], ai[4];
        if (ST == 1) {
            const float4 vr = *reinterpret_cast<const float4*>(re + i0);
            const float4 vi = *reinterpret_cast<const float4*>(im + i0);
            ar[0] = vr.x; ar[1] = vr.y; ar[2] = vr.z; ar[3] = vr.w;
            ai[0] = vi.x; ai[1] = vi.y; ai[2] = vi.z; ai[3] = vi.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) { ar[q] = re[i0 + q * ST]; ai[q] = im[i0 + q * ST]; }
        }
        // 4-point DFT, omega_4 = -i
        const float s02r = ar[0] + ar[2], s02i = ai[0] + ai[2], d02r = ar[0] - ar[2], d02i = ai[0] - ai[2];
        const float s13r = ar[1] + ar[3], s13i = ai[1] + ai[3], d13r = ar[1] - ar[3], d13i = ai[1] - ai[3];
        float2 u0 = make_float2(s02r + s13r, s02i + s13i);
        float2 u2 = make_float2(s02r - s13r, s02i - s13i);
        float2 u1 = make_float2(d02r + d13i, d02i - d13r);  // d02 - i d13
        float2 u3 = make_float2(d02r - d13i, d02i + d13r);  // d02 + i d13
        if (ST == 1) {
            *reinterpret_cast<float4*>(re + i0) = make_float4(u0.x, u1.x, u2.x, u3.x);
            *reinterpret_cast<float4*>(im + i0) = make_float4(u0.y, u1.y, u2.y, u3.y);
        } else {
            const float2 w1 = __ldg(tw + j * TWS);
            const float2 w2 = cmul(w1, w1);
            const float2 w3 = cmul(w2, w1);
            u1 = cmul(u1, w1); u2 = cmul(u2, w2); u3 = cmul(u3, w3);
            re[i0] = u0.x; im[i0] = u0.y;
            re[i0 + ST] = u1.x; im[i0 + ST] = u1.y;
            re[i0 + 2 * ST] = u2.x; im[i0 + 2 * ST] = u2.y;
            re[i0 + 3 * ST] = u3.x; im[i0 + 3 * ST] = u3.y;
        }
    }
}

template <int N, int NT, int M>
__device__ __forceinline__ void dit_pass(float* re, float* im, const float2* __restrict__ tw) {
    constexpr int ST = M / 4;
    constexpr int TWS = 2 * (N / M);
    for (int b = threadIdx.x; b < N / 4; b += NT) {
        const int j = b & (ST - 1);
        const int i0 = ((b - j) << 2) + j;
        float2 u0, u1, u2, u3;
        if (ST == 1) {
            const float4 vr = *reinterpret_cast<const float4*>(re + i0);
            const float4 vi = *reinterpret_cast<const float4*>(im + i0);
            u0 = make_float2(vr.x, vi.x); u1 = make_float2(vr.y, vi.y);
            u2 = make_float2(vr.z, vi.z); u3 = make_float2(vr.w, vi.w);
        } else {
            u0 = make_float2(re[i0], im[i0]);
            u1 = make_float2(re[i0 + ST], im[i0 + ST]);
            u2 = make_float2(re[i0 + 2 * ST], im[i0 + 2 * ST]);
            u3 = make_float2(re[i0 + 3 * ST], im[i0 + 3 * ST]);
            const float2 w1 = __ldg(tw + j * TWS);
            const float2 w2 = cmul(w1, w1);
            const float2 w3 = cmul(w2, w1);
            u1 = cmulc(u1, w1); u2 = cmulc(u2, w2); u3 = cmulc(u3, w3);
        }
        // inverse 4-point DFT (omega_4^-1 = +i), unnormalised
        const float s02r = u0.x + u2.x, s02i = u0.y + u2.y, d02r = u0.x - u2.x, d02i = u0.y - u2.y;
        const float s13r = u1.x + u3.x, s13i = u1.y + u3.y, d13r = u1.x - u3.x, d13i = u1.y - u3.y;
        const float a0r = s02r + s13r, a0i = s02i + s13i;
        const float a2r = s02r - s13r, a2i = s02i - s13i;
        const float a1r = d02r - d13i, a1i = d02i + d13r;  // d02 + i d13
        const float a3r = d02r + d13i, a3i = d02i - d13r;  // d02 - i d13
        if (ST == 1) {
            *reinterpret_cast<float4*>(re + i0) = make_float4(a0r, a1r, a2r, a3r);
            *reinterpret_cast<float4*>(im + i0) = make_float4(a0i, a1i, a2i, a3i);
        } else {
            re[i0] = a0r; im[i0] = a0i;
            re[i0 + ST] = a1r; im[i0 + ST] = a1i;
            re[i0 + 2 * ST] = a2r; im[i0 + 2 * ST] = a2i;
            re[i0 + 3 * ST] = a3r; im[i0 + 3 * ST] = a3i;
        }
    }
}

template <int N, int NT, int M>
struct FwdPasses {
    static __device__ __forceinline__ void run(float* re, float* im, const float2* tw) {
        dif_pass<N, NT, M>(re, im, tw);
        __syncthreads();
        if constexpr (M > 4) FwdPasses<N, NT, M / 4>::run(re, im, tw);
    }
};
template <int N, int NT, int M>
struct InvPasses {
    static __device__ __forceinline__ void run(float* re, float* im, const float2* tw) {
        dit_pass<N, NT, M>(re, im, tw);
        __syncthreads();
        if constexpr (M < N) InvPasses<N, NT, M * 4>::run(re, im, tw);
    }
};

template <int N> struct Log4;
template <> struct Log4<1024> { static constexpr int v = 5; };
template <> struct Log4<4096> { static constexpr int v = 6; };
template <> struct Log4<16384> { static constexpr int v = 7; };

// packed complex FFT Z (digit-reversed order, in smem) -> true half spectrum A (same order):
//   slot 0 holds (A[0], A[N]) (both real); slot pos(k) holds A[k], 0 < k < N.  `scale` multiplies.
template <int N, int NT>
__device__ __forceinline__ void untangle_inplace(float* re, float* im, const float2* __restrict__ tw, float scale) {
    constexpr int L4 = Log4<N>::v;
    for (int q = threadIdx.x; q < N / 2; q += NT) {
        const int p = ((q >> 1) << 2) | (q & 1);
        const int k = drev4<L4>(p);
        if (k == 0) {
            const float a = re[0], b = im[0];
            re[0] = (a + b) * scale; im[0] = (a - b) * scale;
            im[2] = -im[2] * scale; re[2] = re[2] * scale;  // k = N/2 lives at position 2: conj
            continue;
        }
        const int pm = drev4<L4>(N - k);
        const float2 zk = make_float2(re[p], im[p]), zm = make_float2(re[pm], im[pm]);
        const float2 w = __ldg(tw + k);
        const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
        const float dr = zk.x - zm.x, di = zk.y + zm.y;
        // O = -(i/2) w D
        const float2 wd = cmul(w, make_float2(dr, di));
        const float orr = 0.5f * wd.y, oi = -0.5f * wd.x;
        re[p] = (er + orr) * scale; im[p] = (ei + oi) * scale;
        re[pm] = (er - orr) * scale; im[pm] = -(ei - oi) * scale;
    }
}

// true spectrum Y (digit-reversed order, slot 0 = (Y[0], Y[N])) -> packed complex Z' for the inverse
template <int N, int NT>
__device__ __forceinline__ void retangle_inplace(float* re, float* im, const float2* __restrict__ tw) {
    constexpr int L4 = Log4<N>::v;
    for (int q = threadIdx.x; q < N / 2; q += NT) {
        const int p = ((q >> 1) << 2) | (q & 1);
        const int k = drev4<L4>(p);
        if (k == 0) {
            const float y0 = re[0], yn = im[0];
            re[0] = 0.5f * (y0 + yn); im[0] = 0.5f * (y0 - yn);
            im[2] = -im[2];
            continue;
        }
        const int pm = drev4<L4>(N - k);
        const float2 yk = make_float2(re[p], im[p]), ym = make_float2(re[pm], im[pm]);
        const float2 w = __ldg(tw + k);
        const float er = 0.5f * (yk.x + ym.x), ei = 0.5f * (yk.y - ym.y);
        const float dr = yk.x - ym.x, di = yk.y + ym.y;
        // O' = (i/2) conj(w) D'
        const float2 wd = cmulc(make_float2(dr, di), w);
        const float orr = -0.5f * wd.y, oi = 0.5f * wd.x;
        re[p] = er + orr; im[p] = ei + oi;
        re[pm] = er - orr; im[pm] = -(ei - oi);
    }
}

// fused: untangle X, multiply by the (already 1/N-scaled) filter spectrum H, retangle
template <int N, int NT>
__device__ __forceinline__ void pointwise_filter(float* re, float* im, const float2* __restrict__ tw,
                                                 const float2* __restrict__ H) {
    constexpr int L4 = Log4<N>::v;
    for (int q = threadIdx.x; q < N / 2; q += NT) {
        const int p = ((q >> 1) << 2) | (q & 1);
        const int k = drev4<L4>(p);
        if (k == 0) {
            const float a = re[0], b = im[0];
            const float2 h0 = H[0];
            const float y0 = (a + b) * h0.x, yn = (a - b) * h0.y;
            re[0] = 0.5f * (y0 + yn); im[0] = 0.5f * (y0 - yn);
            // k = N/2 at position 2: X = conj(Z), Y = X H, Z' = conj(Y) = Z conj(H)
            const float2 z2 = make_float2(re[2], im[2]);
            const float2 r2 = cmulc(z2, H[2]);
            re[2] = r2.x; im[2] = r2.y;
            continue;
        }
        const int pm = drev4<L4>(N - k);
        const float2 zk = make_float2(re[p], im[p]), zm = make_float2(re[pm], im[pm]);
        const float2 w = __ldg(tw + k);
        float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
        float dr = zk.x - zm.x, di = zk.y + zm.y;
        float2 wd = cmul(w, make_float2(dr, di));
        float orr = 0.5f * wd.y, oi = -0.5f * wd.x;
        const float2 xk = make_float2(er + orr, ei + oi);
        const float2 xm = make_float2(er - orr, -(ei - oi));
        const float2 yk = cmul(xk, H[p]);
        const float2 ym = cmul(xm, H[pm]);
        er = 0.5f * (yk.x + ym.x); ei = 0.5f * (yk.y - ym.y);
        dr = yk.x - ym.x; di = yk.y + ym.y;
        wd = cmulc(make_float2(dr, di), w);
        orr = -0.5f * wd.y; oi = 0.5f * wd.x;
        re[p] = er + orr; im[p] = ei + oi;
        re[pm] = er - orr; im[pm] = -(ei - oi);
    }
}

// ------------------------------------------------------------------ segment load / store
// loads F = 2N real samples src[s0 .. s0+F) (zero outside [0, len)) as z[j] = x[2j] + i x[2j+1]
template <int N, int NT>
__device__ __forceinline__ void load_packed(float* re, float* im, const float* __restrict__ src, long long s0,
                                            long long len, bool vec_ok) {
    if (vec_ok && (s0 & 3) == 0) {
        for (int t = threadIdx.x; t < N / 2; t += NT) {
            const long long pos = s0 + 4LL * t;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos >= 0 && pos + 4 <= len) {
                v = __ldg(reinterpret_cast<const float4*>(src + pos));
            } else if (pos + 4 > 0 && pos < len) {
                float e[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) e[c] = (pos + c >= 0 && pos + c < len) ? src[pos + c] : 0.f;
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            *reinterpret_cast<float2*>(re + 2 * t) = make_float2(v.x, v.z);
            *reinterpret_cast<float2*>(im + 2 * t) = make_float2(v.y, v.w);
        }
    } else {
        for (int i = threadIdx.x; i < 2 * N; i += NT) {
            const long long pos = s0 + i;
            const float v = (pos >= 0 && pos < len) ? src[pos] : 0.f;
            if (i & 1) im[i >> 1] = v; else re[i >> 1] = v;
        }
    }
}

// writes segment samples [i_lo, i_lo + count) to dst[d0 .. d0+count), clipped to [0, len)
template <int N, int NT>
__device__ __forceinline__ void store_packed(const float* re, const float* im, float* __restrict__ dst, int i_lo,
                                             int count, long long d0, long long len, bool vec_ok) {
    if (vec_ok && (i_lo & 3) == 0 && (d0 & 3) == 0 && (count & 3) == 0) {
        for (int t = threadIdx.x; t < count / 4; t += NT) {
            const int i = i_lo + 4 * t;
            const long long pos = d0 + 4LL * t;
            if (pos >= 0 && pos + 4 <= len) {
                const float2 r = *reinterpret_cast<const float2*>(re + (i >> 1));
                const float2 m = *reinterpret_cast<const float2*>(im + (i >> 1));
                stg_stream(reinterpret_cast<float4*>(dst + pos), make_float4(r.x, m.x, r.y, m.y));
            } else {
                for (int c = 0; c < 4; ++c) {
                    const long long pc = pos + c;
                    if (pc >= 0 && pc < len) dst[pc] = ((i + c) & 1) ? im[(i + c) >> 1] : re[(i + c) >> 1];
                }
            }
        }
    } else {
        for (int t = threadIdx.x; t < count; t += NT) {
            const int i = i_lo + t;
            const long long pos = d0 + t;
            if (pos >= 0 && pos < len) dst[pos] = (i & 1) ? im[i >> 1] : re[i >> 1];
        }
    }
}

struct RowMap {  // output row -> (x row, h row) with channel broadcasting
    int c_out, cx, ch;
    __device__ __forceinline__ void map(int r, int& xr, int& hr) const {
        const int b = r / c_out, c = r - b * c_out;
        xr = b * cx + (cx == 1 ? 0 : c);
        hr = b * ch + (ch == 1 ? 0 : c);
    }
};

// ------------------------------------------------------------------ kernels
// spectra of filter partitions: Hs[(hrow * P + part) * N + pos], scaled by 1/N
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_spectrum_kernel(const float* __restrict__ h, float2* __restrict__ Hs,
                                                          int Nh, int part_len, int P,
                                                          const float2* __restrict__ tw, int vec_ok) {
    extern __shared__ __align__(16) float smem_f[];
    float* re = smem_f;
    float* im = smem_f + N;
    const int hrow = blockIdx.x / P, part = blockIdx.x - hrow * P;
    const float* src = h + (size_t)hrow * Nh + (size_t)part * part_len;
    long long len = (long long)Nh - (long long)part * part_len;
    if (len > part_len) len = part_len;
    const bool v = vec_ok && ((((size_t)hrow * Nh + (size_t)part * part_len) & 3) == 0);
    load_packed<N, NT>(re, im, src, 0, len, v);
    __syncthreads();
    FwdPasses<N, NT, N>::run(re, im, tw);
    untangle_inplace<N, NT>(re, im, tw, 1.f / (float)N);
    __syncthreads();
    float2* out = Hs + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += NT) out[i] = make_float2(re[i], im[i]);
}

// single-partition overlap-save: block j produces full-convolution samples [j*hop, (j+1)*hop)
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_ols_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                     const float2* __restrict__ Hs, RowMap rm, long long L, int pre,
                                                     int hop, int shift, int nblk, const float2* __restrict__ tw,
                                                     int vec_ok) {
    extern __shared__ __align__(16) float smem_f[];
    float* re = smem_f;
    float* im = smem_f + N;
    const int row = blockIdx.x / nblk, j = blockIdx.x - row * nblk;
    int xr, hr;
    rm.map(row, xr, hr);
    const long long m0 = (long long)j * hop;
    load_packed<N, NT>(re, im, x + (size_t)xr * L, m0 - pre, L, vec_ok);
    __syncthreads();
    FwdPasses<N, NT, N>::run(re, im, tw);
    pointwise_filter<N, NT>(re, im, tw, Hs + (size_t)hr * N);
    __syncthreads();
    InvPasses<N, NT, 4>::run(re, im, tw);
    store_packed<N, NT>(re, im, y + (size_t)row * L, pre, hop, m0 - shift, L, vec_ok);
}

// UPOLS step 1: spectra of input blocks.  Xs[(rloc * nblk + j) * N + pos] = rfft of x[(j-1)B, (j+1)B), B = N
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_xspec_kernel(const float* __restrict__ x, float2* __restrict__ Xs,
                                                       int xrow0, long long L, int nblk,
                                                       const float2* __restrict__ tw, int vec_ok) {
    extern __shared__ __align__(16) float smem_f[];
    float* re = smem_f;
    float* im = smem_f + N;
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    load_packed<N, NT>(re, im, x + (size_t)(xrow0 + rloc) * L, ((long long)j - 1) * N, L, vec_ok);
    __syncthreads();
    FwdPasses<N, NT, N>::run(re, im, tw);
    untangle_inplace<N, NT>(re, im, tw, 1.f);
    __syncthreads();
    float2* out = Xs + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += NT) out[i] = make_float2(re[i], im[i]);
}

// UPOLS step 2: Y_j = sum_p X_{j-p} H_p, inverse FFT, keep the second half of the block
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_upols_kernel(const float2* __restrict__ Xs, const float2* __restrict__ Hs,
                                                       float* __restrict__ y, RowMap rm, int row0, int xrow0,
                                                       int hrow0, long long L, int P, int nblk, int shift,
                                                       const float2* __restrict__ tw, int vec_ok) {
    extern __shared__ __align__(16) float smem_f[];
    float* re = smem_f;
    float* im = smem_f + N;
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    const int row = row0 + rloc;
    int xr, hr;
    rm.map(row, xr, hr);
    const float2* Xrow = Xs + (size_t)(xr - xrow0) * nblk * N;
    const float2* Hrow = Hs + (size_t)(hr - hrow0) * P * N;
    const int pmax = j < P - 1 ? j : P - 1;
    for (int i = threadIdx.x; i < N; i += NT) {
        float ar = 0.f, ai = 0.f;
        if (i == 0) {
            for (int p = 0; p <= pmax; ++p) {
                const float2 xv = __ldg(Xrow + (size_t)(j - p) * N), hv = __ldg(Hrow + (size_t)p * N);
                ar = fmaf(xv.x, hv.x, ar);
                ai = fmaf(xv.y, hv.y, ai);
            }
        } else {
            for (int p = 0; p <= pmax; ++p) {
                const float2 xv = __ldg(Xrow + (size_t)(j - p) * N + i), hv = __ldg(Hrow + (size_t)p * N + i);
                ar = fmaf(xv.x, hv.x, ar); ar = fmaf(-xv.y, hv.y, ar);
                ai = fmaf(xv.x, hv.y, ai); ai = fmaf(xv.y, hv.x, ai);
            }
        }
        re[i] = ar; im[i] = ai;
    }
    __syncthreads();
    retangle_inplace<N, NT>(re, im, tw);
    __syncthreads();
    InvPasses<N, NT, 4>::run(re, im, tw);
    store_packed<N, NT>(re, im, y + (size_t)row * L, N, N, (long long)j * N - shift, L, vec_ok);
}

__global__ void fft_plan_kernel(float2* tw, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 2 * n) {
        const double a = (double)k / (double)n;
        tw[k] = make_float2((float)cospi(a), (float)(-sinpi(a)));
    }
}

static int pick_fft_size(int Nh) {
    if (Nh <= 512) return 1024;
    if (Nh <= 2048) return 4096;
    return 16384;
}

template <typename K>
static int set_smem(K kern, size_t smem) {
    GFX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return GFX_OK;
}

struct FirArgs {
    const float* x; const float* h; float* y;
    int batch, cx, ch; long long L; int Nh; int shift;
    const float2* tw; unsigned char* ws; size_t ws_bytes; cudaStream_t stream;
};

template <int N, int NT>
static int run_ols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    const int rows = a.batch * c_out, hrows = a.batch * a.ch;
    const size_t need = (size_t)hrows * N * sizeof(float2);
    if (!a.ws || a.ws_bytes < need) return GFX_ERR_WORKSPACE;
    float2* Hs = (float2*)a.ws;
    const size_t smem = (size_t)2 * N * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (set_smem(fir_spectrum_kernel<N, NT>, smem) || set_smem(fir_ols_kernel<N, NT>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    const int hvec = ((uintptr_t)a.h % 16 == 0);
    fir_spectrum_kernel<N, NT><<<hrows, NT, smem, a.stream>>>(a.h, Hs, a.Nh, a.Nh, 1, a.tw, hvec);
    GFX_CUDA_CHECK(cudaGetLastError());
    const int pre = (a.Nh - 1 + 3) & ~3;
    const int hop = (2 * N - pre) & ~3;
    const long long total = a.L + a.shift;
    const long long nblk = (total + hop - 1) / hop;
    if (nblk * rows > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
    const int vec = (((uintptr_t)a.x | (uintptr_t)a.y) % 16 == 0) && (a.L % 4 == 0);
    RowMap rm{c_out, a.cx, a.ch};
    fir_ols_kernel<N, NT><<<(unsigned)(nblk * rows), NT, smem, a.stream>>>(a.x, a.y, Hs, rm, a.L, pre, hop, a.shift,
                                                                           (int)nblk, a.tw, vec);
    GFX_CUDA_CHECK(cudaGetLastError());
    return GFX_OK;
}

static void upols_geometry(int batch, int cx, int ch, long long L, int Nh, int shift, int N, int& P, long long& nblk,
                           size_t& per_item_bytes) {
    P = (Nh + N - 1) / N;
    nblk = (L + shift + N - 1) / N;
    per_item_bytes = ((size_t)ch * P + (size_t)cx * nblk) * N * sizeof(float2);
}

template <int N, int NT>
static int run_upols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    int P; long long nblk; size_t per_item;
    upols_geometry(a.batch, a.cx, a.ch, a.L, a.Nh, a.shift, N, P, nblk, per_item);
    if (!a.ws || a.ws_bytes < per_item) return GFX_ERR_WORKSPACE;
    long long chunk = (long long)(a.ws_bytes / per_item);
    if (chunk > a.batch) chunk = a.batch;
    const size_t smem = (size_t)2 * N * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (set_smem(fir_spectrum_kernel<N, NT>, smem) || set_smem(fir_xspec_kernel<N, NT>, smem) ||
            set_smem(fir_upols_kernel<N, NT>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    const int hvec = ((uintptr_t)a.h % 16 == 0) && (a.Nh % 4 == 0);
    const int vec = (((uintptr_t)a.x | (uintptr_t)a.y) % 16 == 0) && (a.L % 4 == 0);
    RowMap rm{c_out, a.cx, a.ch};
    for (long long b0 = 0; b0 < a.batch; b0 += chunk) {
        const int nb = (int)((a.batch - b0 < chunk) ? a.batch - b0 : chunk);
        float2* Hs = (float2*)a.ws;
        float2* Xs = Hs + (size_t)nb * a.ch * P * N;
        const int hrow0 = (int)b0 * a.ch, xrow0 = (int)b0 * a.cx, row0 = (int)b0 * c_out;
        if ((long long)nb * a.cx * nblk > 0x7fffffffLL || (long long)nb * c_out * nblk > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
        fir_spectrum_kernel<N, NT><<<nb * a.ch * P, NT, smem, a.stream>>>(a.h + (size_t)hrow0 * a.Nh, Hs, a.Nh, N, P,
                                                                          a.tw, hvec);
        GFX_CUDA_CHECK(cudaGetLastError());
        fir_xspec_kernel<N, NT><<<(unsigned)(nb * a.cx * nblk), NT, smem, a.stream>>>(a.x, Xs, xrow0, a.L, (int)nblk,
                                                                                      a.tw, vec);
        GFX_CUDA_CHECK(cudaGetLastError());
        fir_upols_kernel<N, NT><<<(unsigned)(nb * c_out * nblk), NT, smem, a.stream>>>(
            Xs, Hs, a.y, rm, row0, xrow0, hrow0, a.L, P, (int)nblk, a.shift, a.tw, vec);
        GFX_CUDA_CHECK(cudaGetLastError());
    }
    return GFX_OK;
}

}  // namespace gfx

extern "C" {

int gfx_fir_fft_size(int filter_len) { return filter_len <= 0 ? GFX_ERR_INVALID : gfx::pick_fft_size(filter_len); }

size_t gfx_fft_plan_bytes(int n) { return (size_t)2 * n * sizeof(float2); }

int gfx_fft_plan_init(void* plan, int n, void* stream) {
    if (!plan || (n != 1024 && n != 4096 && n != 16384)) return GFX_ERR_INVALID;
    gfx::fft_plan_kernel<<<(2 * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float2*)plan, n);
    GFX_CUDA_CHECK(cudaGetLastError());
    return GFX_OK;
}

size_t gfx_fir_conv_workspace_bytes(int batch, int cx, int ch, long long L, int filter_len, int zerophase) {
    if (batch <= 0 || cx <= 0 || ch <= 0 || L <= 0 || filter_len <= 0) return 0;
    const int n = gfx::pick_fft_size(filter_len);
    if (filter_len <= 16384) return (size_t)batch * ch * n * sizeof(float2);
    int P; long long nblk; size_t per_item;
    gfx::upols_geometry(batch, cx, ch, L, filter_len, zerophase ? filter_len / 2 : 0, n, P, nblk, per_item);
    // spectra of ~48 MB worth of batch items per sweep stay L2-resident on B200 (126 MB L2)
    size_t items = (size_t)(48u << 20) / per_item;
    if (items < 1) items = 1;
    if (items > (size_t)batch) items = batch;
    return items * per_item;
}

int gfx_fir_conv_f32(const float* x, const float* h, float* y, int batch, int cx, int ch, long long L,
                     int filter_len, int zerophase, const void* plan, void* workspace, size_t workspace_bytes,
                     void* stream) {
    using namespace gfx;
    if (!x || !h || !y || !plan) return GFX_ERR_INVALID;
    if (batch <= 0 || cx <= 0 || ch <= 0 || L <= 0 || filter_len <= 0) return GFX_ERR_INVALID;
    if (cx != ch && cx != 1 && ch != 1) return GFX_ERR_INVALID;
    FirArgs a{x, h, y, batch, cx, ch, L, filter_len, zerophase ? filter_len / 2 : 0, (const float2*)plan,
              (unsigned char*)workspace, workspace_bytes, (cudaStream_t)stream};
    const int n = pick_fft_size(filter_len);
    if (filter_len <= 16384) {
        if (n == 1024) return run_ols<1024, 256>(a);
        if (n == 4096) return run_ols<4096, 256>(a);
        return run_ols<16384, 1024>(a);
    }
    return run_upols<16384, 1024>(a);
}

}  // extern "C"
