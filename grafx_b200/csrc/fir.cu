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

// ------------------------------------------------------------------ complex arithmetic on packed pairs
// A complex number is one 64-bit register pair (re, im); add/sub are one FADD2, a twiddle multiply is
// FMUL2 + FFMA2 with the twiddle supplied as two pairs:  forward (wr, wi) & (-wi, wr);  conjugate
// (wr, -wi) & (wi, wr).  The table entry of angle t is { fwd.xy, fwd.zw, inv.xy, inv.zw } (32 bytes).
struct Tw { pk2 a, b; };
__device__ __forceinline__ pk2 cmul_tw(pk2 v, const Tw& w) {
    float re, im;
    pk_split(v, re, im);
    return pk_fma(pk_dup(re), w.a, pk_mul(pk_dup(im), w.b));
}
__device__ __forceinline__ float4 ld_tw4(const float4* p) {
    // volatile: keeps the twiddle requests where they are written (first), so that their latency
    // overlaps the shared-memory loads and the butterfly that follow
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ Tw as_tw(const float4& e) {
    Tw w;
    w.a = pk_make(e.x, e.y);
    w.b = pk_make(e.z, e.w);
    return w;
}
__device__ __forceinline__ Tw const_tw(float c, float s, bool inv) {
    // omega = exp(-i theta) = (c, -s) forward; conj for the inverse
    Tw w;
    const float wi = inv ? s : -s;
    w.a = pk_make(c, wi);
    w.b = pk_make(-wi, c);
    return w;
}

template <bool INV>
__device__ __forceinline__ void r4(pk2& a0, pk2& a1, pk2& a2, pk2& a3) {
    // 4-point DFT (INV: inverse, unnormalised); omega_4 = -i
    const pk2 s02 = pk_add(a0, a2), d02 = pk_sub(a0, a2), s13 = pk_add(a1, a3), d13 = pk_sub(a1, a3);
    const pk2 r = pk_swap(d13);                     // (d13.im, d13.re)
    const pk2 pm = pk_make(1.f, -1.f), mp = pk_make(-1.f, 1.f);
    a0 = pk_add(s02, s13);
    a2 = pk_sub(s02, s13);
    a1 = pk_fma(r, INV ? mp : pm, d02);             // d02 -/+ i d13
    a3 = pk_fma(r, INV ? pm : mp, d02);
}

// 16-point DFT in registers.  In: a[q] natural order.  Out: register a[4*i + j] holds X[4*j + i].
template <bool INV>
__device__ __forceinline__ void r16(pk2 (&a)[16]) {
#pragma unroll
    for (int q0 = 0; q0 < 4; ++q0) r4<INV>(a[q0], a[q0 + 4], a[q0 + 8], a[q0 + 12]);
    // a[q0 + 4 r0] *= omega_16^(q0 r0)
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
    const Tw w1 = const_tw(C1, S1, INV), w2 = const_tw(H, H, INV), w3 = const_tw(S1, C1, INV);
    const Tw w4 = const_tw(0.f, 1.f, INV), w6 = const_tw(-H, H, INV), w9 = const_tw(-C1, -S1, INV);
    a[1 + 4] = cmul_tw(a[1 + 4], w1); a[2 + 4] = cmul_tw(a[2 + 4], w2); a[3 + 4] = cmul_tw(a[3 + 4], w3);
    a[1 + 8] = cmul_tw(a[1 + 8], w2); a[2 + 8] = cmul_tw(a[2 + 8], w4); a[3 + 8] = cmul_tw(a[3 + 8], w6);
    a[1 + 12] = cmul_tw(a[1 + 12], w3); a[2 + 12] = cmul_tw(a[2 + 12], w6); a[3 + 12] = cmul_tw(a[3 + 12], w9);
#pragma unroll
    for (int r0 = 0; r0 < 4; ++r0) r4<INV>(a[4 * r0], a[4 * r0 + 1], a[4 * r0 + 2], a[4 * r0 + 3]);
}

// shared-memory layout: complex point i lives at 64-bit slot i + (i >> 4) (one pad slot per 16 points):
// every pass below is then bank-conflict free for 64-bit accesses.
__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int fft_smem_slots(int n) { return n + n / 16 + 16; }

// ---- radix plans (forward order).  N = 1024: 4,16,16   N = 4096: 16,16,16   N = 16384: 4,16,16,16
__host__ __device__ constexpr int plan_radix(int n, int s) {
    return n == 4096 ? (s < 3 ? 16 : 1) : (s == 0 ? 4 : (s < (n == 1024 ? 3 : 4) ? 16 : 1));
}
__host__ __device__ constexpr int plan_m(int n, int s) {  // sub-transform length entering pass s
    int m = n;
    for (int i = 0; i < s; ++i) m /= plan_radix(n, i);
    return m;
}
__host__ __device__ constexpr int plan_entries(int n, int s) {  // twiddles of pass s: (R-1) * ST, none when ST == 1
    const int r = plan_radix(n, s);
    const int st = r > 1 ? plan_m(n, s) / r : 1;
    return (r > 1 && st > 1) ? (r - 1) * st : 0;
}
__host__ __device__ constexpr int plan_offset(int n, int s) {
    int o = 0;
    for (int i = 0; i < s; ++i) o += plan_entries(n, i);
    return o;
}
__host__ __device__ constexpr int plan_total(int n) { return plan_offset(n, 4); }
// plan memory (float4 units): [ half-angle table: n float2 = n/2 float4 | forward pass tables | inverse pass tables ]
//   pass table entry (r-1) * ST + j = twiddle omega_M^(j r) as { wr, wi, -wi, wr } (forward) / { wr, -wi, wi, wr }
//   (conjugate): consecutive threads (j) read consecutive 16-byte entries.

template <int N, int NT, int S, bool INV>
__device__ __forceinline__ void fft_pass(pk2* z, const float4* __restrict__ plan) {
    constexpr int R = plan_radix(N, S);
    constexpr int M = plan_m(N, S);
    constexpr int ST = M / R;
    const float4* tp = plan + N / 2 + plan_offset(N, S) + (INV ? plan_total(N) : 0);
    // slot of element q of a butterfly: pidx(i0 + q ST) = pidx(i0) + q * PST  (ST is a multiple of 16, or 1)
    constexpr int PST = ST >= 16 ? ST + ST / 16 : 1;
    static_assert(ST == 1 || ST % 16 == 0, "stride must keep the padding pattern linear");
    for (int b = threadIdx.x; b < N / R; b += NT) {
        const int j = b & (ST - 1);
        const int i0 = (b - j) * R + j;
        pk2* zb = z + pidx(i0);
        if constexpr (R == 4) {
            float4 t1, t2, t3;
            if constexpr (ST > 1) { t1 = ld_tw4(tp + j); t2 = ld_tw4(tp + ST + j); t3 = ld_tw4(tp + 2 * ST + j); }
            pk2 a0 = zb[0], a1 = zb[PST], a2 = zb[2 * PST], a3 = zb[3 * PST];
            if constexpr (INV && ST > 1) {
                a1 = cmul_tw(a1, as_tw(t1)); a2 = cmul_tw(a2, as_tw(t2)); a3 = cmul_tw(a3, as_tw(t3));
            }
            r4<INV>(a0, a1, a2, a3);
            if constexpr (!INV && ST > 1) {
                a1 = cmul_tw(a1, as_tw(t1)); a2 = cmul_tw(a2, as_tw(t2)); a3 = cmul_tw(a3, as_tw(t3));
            }
            zb[0] = a0; zb[PST] = a1; zb[2 * PST] = a2; zb[3 * PST] = a3;
        } else {
            float4 tw[15];
            if constexpr (ST > 1) {
#pragma unroll
                for (int r = 1; r < 16; ++r) tw[r - 1] = ld_tw4(tp + (r - 1) * ST + j);
            }
            pk2 a[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) a[q] = zb[q * PST];
            if constexpr (INV && ST > 1) {
#pragma unroll
                for (int q = 1; q < 16; ++q) a[q] = cmul_tw(a[q], as_tw(tw[q - 1]));
            }
            r16<INV>(a);
            // register a[4*i + jj] holds output index 4*jj + i
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int r = 4 * jj + i;
                    pk2 v = a[4 * i + jj];
                    if constexpr (!INV && ST > 1) {
                        if (r > 0) v = cmul_tw(v, as_tw(tw[r - 1]));
                    }
                    zb[r * PST] = v;
                }
            }
        }
    }
}

template <int N, int NT>
__device__ __forceinline__ void fft_forward(pk2* z, const float4* __restrict__ plan) {
    fft_pass<N, NT, 0, false>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 1, false>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 2, false>(z, plan);
    __syncthreads();
    if constexpr (plan_radix(N, 3) > 1) {
        fft_pass<N, NT, 3, false>(z, plan);
        __syncthreads();
    }
}
template <int N, int NT>
__device__ __forceinline__ void fft_inverse(pk2* z, const float4* __restrict__ plan) {
    if constexpr (plan_radix(N, 3) > 1) {
        fft_pass<N, NT, 3, true>(z, plan);
        __syncthreads();
    }
    fft_pass<N, NT, 2, true>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 1, true>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 0, true>(z, plan);
    __syncthreads();
}

// digit reversal of the plans above: position (mixed-radix digits, first radix most significant) <-> bin
template <int N> __device__ __forceinline__ int bin_of_pos(int p);
template <int N> __device__ __forceinline__ int pos_of_bin(int k);
template <> __device__ __forceinline__ int bin_of_pos<1024>(int p) { return (p >> 8) | (((p >> 4) & 15) << 2) | ((p & 15) << 6); }
template <> __device__ __forceinline__ int pos_of_bin<1024>(int k) { return ((k & 3) << 8) | (((k >> 2) & 15) << 4) | (k >> 6); }
template <> __device__ __forceinline__ int bin_of_pos<4096>(int p) { return (p >> 8) | (((p >> 4) & 15) << 4) | ((p & 15) << 8); }
template <> __device__ __forceinline__ int pos_of_bin<4096>(int k) { return ((k & 15) << 8) | (((k >> 4) & 15) << 4) | (k >> 8); }
template <> __device__ __forceinline__ int bin_of_pos<16384>(int p) {
    return (p >> 12) | (((p >> 8) & 15) << 2) | (((p >> 4) & 15) << 6) | ((p & 15) << 10);
}
template <> __device__ __forceinline__ int pos_of_bin<16384>(int k) {
    return ((k & 3) << 12) | (((k >> 2) & 15) << 8) | (((k >> 6) & 15) << 4) | (k >> 10);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 ld_c(const pk2* z, int i) {
    float re, im;
    pk_split(z[pidx(i)], re, im);
    return make_float2(re, im);
}
__device__ __forceinline__ void st_c(pk2* z, int i, float re, float im) { z[pidx(i)] = pk_make(re, im); }
__device__ __forceinline__ float2 half_tw(const float4* __restrict__ plan, int k) {  // exp(-i pi k / N), k < N
    return __ldg(reinterpret_cast<const float2*>(plan) + k);
}

// The last radix is 16 in every plan: bins k < N/2 are the positions whose last digit is < 8; bin N/2 is
// position 8, bin 0 position 0.
#define GFX_FOR_PAIRS(N, NT)                                                     \
    for (int q = threadIdx.x; q < (N) / 2; q += (NT))                            \
        if (const int p = ((q >> 3) << 4) | (q & 7); true)                       \
            if (const int k = bin_of_pos<N>(p); true)

// packed complex FFT Z (digit-reversed order, in smem) -> true half spectrum A (same order):
//   slot 0 holds (A[0], A[N]) (both real); slot pos(k) holds A[k], 0 < k < N.  `scale` multiplies.
template <int N, int NT>
__device__ __forceinline__ void untangle_inplace(pk2* z, const float4* __restrict__ tw4, float scale) {
    GFX_FOR_PAIRS(N, NT) {
        if (k == 0) {
            const float2 z0 = ld_c(z, 0), zh = ld_c(z, 8);
            st_c(z, 0, (z0.x + z0.y) * scale, (z0.x - z0.y) * scale);
            st_c(z, 8, zh.x * scale, -zh.y * scale);  // k = N/2: conj
            continue;
        }
        const int pm = pos_of_bin<N>(N - k);
        const float2 zk = ld_c(z, p), zm = ld_c(z, pm);
        const float2 w = half_tw(tw4, k);
        const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
        const float dr = zk.x - zm.x, di = zk.y + zm.y;
        const float2 wd = cmul(w, make_float2(dr, di));  // O = -(i/2) w D
        const float orr = 0.5f * wd.y, oi = -0.5f * wd.x;
        st_c(z, p, (er + orr) * scale, (ei + oi) * scale);
        st_c(z, pm, (er - orr) * scale, -(ei - oi) * scale);
    }
}

// true spectrum Y (digit-reversed order, slot 0 = (Y[0], Y[N])) -> packed complex Z' for the inverse
template <int N, int NT>
__device__ __forceinline__ void retangle_inplace(pk2* z, const float4* __restrict__ tw4) {
    GFX_FOR_PAIRS(N, NT) {
        if (k == 0) {
            const float2 y0 = ld_c(z, 0), yh = ld_c(z, 8);
            st_c(z, 0, 0.5f * (y0.x + y0.y), 0.5f * (y0.x - y0.y));
            st_c(z, 8, yh.x, -yh.y);
            continue;
        }
        const int pm = pos_of_bin<N>(N - k);
        const float2 yk = ld_c(z, p), ym = ld_c(z, pm);
        const float2 w = half_tw(tw4, k);
        const float er = 0.5f * (yk.x + ym.x), ei = 0.5f * (yk.y - ym.y);
        const float dr = yk.x - ym.x, di = yk.y + ym.y;
        const float2 wd = cmulc(make_float2(dr, di), w);  // O' = (i/2) conj(w) D'
        const float orr = -0.5f * wd.y, oi = 0.5f * wd.x;
        st_c(z, p, er + orr, ei + oi);
        st_c(z, pm, er - orr, -(ei - oi));
    }
}

// fused: untangle X, multiply by the (already 1/N-scaled) filter spectrum H, retangle
template <int N, int NT>
__device__ __forceinline__ void pointwise_filter(pk2* z, const float4* __restrict__ tw4, const float2* __restrict__ H) {
    GFX_FOR_PAIRS(N, NT) {
        if (k == 0) {
            const float2 z0 = ld_c(z, 0), zh = ld_c(z, 8);
            const float2 h0 = H[0];
            const float y0 = (z0.x + z0.y) * h0.x, yn = (z0.x - z0.y) * h0.y;
            st_c(z, 0, 0.5f * (y0 + yn), 0.5f * (y0 - yn));
            const float2 r2 = cmulc(zh, H[8]);  // X = conj(Z), Y = X H, Z' = conj(Y) = Z conj(H)
            st_c(z, 8, r2.x, r2.y);
            continue;
        }
        const int pm = pos_of_bin<N>(N - k);
        const float2 zk = ld_c(z, p), zm = ld_c(z, pm);
        const float2 w = half_tw(tw4, k);
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
        st_c(z, p, er + orr, ei + oi);
        st_c(z, pm, er - orr, -(ei - oi));
    }
}

// ------------------------------------------------------------------ segment load / store
// loads F = 2N real samples src[s0 .. s0+F) (zero outside [0, len)) as z[j] = x[2j] + i x[2j+1]
template <int N, int NT>
__device__ __forceinline__ void load_packed(pk2* z, const float* __restrict__ src, long long s0, long long len,
                                            bool vec_ok) {
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
            z[pidx(2 * t)] = pk_make(v.x, v.y);
            z[pidx(2 * t + 1)] = pk_make(v.z, v.w);
        }
    } else {
        for (int j = threadIdx.x; j < N; j += NT) {
            const long long pos = s0 + 2LL * j;
            const float a = (pos >= 0 && pos < len) ? src[pos] : 0.f;
            const float b = (pos + 1 >= 0 && pos + 1 < len) ? src[pos + 1] : 0.f;
            z[pidx(j)] = pk_make(a, b);
        }
    }
}

// writes segment samples [i_lo, i_lo + count) to dst[d0 .. d0+count), clipped to [0, len)
template <int N, int NT>
__device__ __forceinline__ void store_packed(const pk2* z, float* __restrict__ dst, int i_lo, int count, long long d0,
                                             long long len, bool vec_ok) {
    if (vec_ok && (i_lo & 3) == 0 && (d0 & 3) == 0 && (count & 3) == 0) {
        for (int t = threadIdx.x; t < count / 4; t += NT) {
            const int i = i_lo + 4 * t;
            const long long pos = d0 + 4LL * t;
            const float2 a = ld_c(z, i >> 1), b = ld_c(z, (i >> 1) + 1);
            if (pos >= 0 && pos + 4 <= len) {
                stg_stream(reinterpret_cast<float4*>(dst + pos), make_float4(a.x, a.y, b.x, b.y));
            } else {
                const float e[4] = {a.x, a.y, b.x, b.y};
                for (int c = 0; c < 4; ++c)
                    if (pos + c >= 0 && pos + c < len) dst[pos + c] = e[c];
            }
        }
    } else {
        for (int t = threadIdx.x; t < count; t += NT) {
            const int i = i_lo + t;
            const long long pos = d0 + t;
            if (pos >= 0 && pos < len) {
                const float2 a = ld_c(z, i >> 1);
                dst[pos] = (i & 1) ? a.y : a.x;
            }
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
                                                          const float4* __restrict__ tw, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int hrow = blockIdx.x / P, part = blockIdx.x - hrow * P;
    const float* src = h + (size_t)hrow * Nh + (size_t)part * part_len;
    long long len = (long long)Nh - (long long)part * part_len;
    if (len > part_len) len = part_len;
    const bool v = vec_ok && ((((size_t)hrow * Nh + (size_t)part * part_len) & 3) == 0);
    load_packed<N, NT>(zbuf, src, 0, len, v);
    __syncthreads();
    fft_forward<N, NT>(zbuf, tw);
    untangle_inplace<N, NT>(zbuf, tw, 1.f / (float)N);
    __syncthreads();
    float2* out = Hs + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += NT) out[i] = ld_c(zbuf, i);
}

// single-partition overlap-save: block j produces full-convolution samples [j*hop, (j+1)*hop)
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_ols_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                     const float2* __restrict__ Hs, RowMap rm, long long L, int pre,
                                                     int hop, int shift, int nblk, const float4* __restrict__ tw,
                                                     int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int row = blockIdx.x / nblk, j = blockIdx.x - row * nblk;
    int xr, hr;
    rm.map(row, xr, hr);
    const long long m0 = (long long)j * hop;
    load_packed<N, NT>(zbuf, x + (size_t)xr * L, m0 - pre, L, vec_ok);
    __syncthreads();
    fft_forward<N, NT>(zbuf, tw);
    pointwise_filter<N, NT>(zbuf, tw, Hs + (size_t)hr * N);
    __syncthreads();
    fft_inverse<N, NT>(zbuf, tw);
    store_packed<N, NT>(zbuf, y + (size_t)row * L, pre, hop, m0 - shift, L, vec_ok);
}

// UPOLS step 1: spectra of input blocks.  Xs[(rloc * nblk + j) * N + pos] = rfft of x[(j-1)B, (j+1)B), B = N
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_xspec_kernel(const float* __restrict__ x, float2* __restrict__ Xs,
                                                       int xrow0, long long L, int nblk,
                                                       const float4* __restrict__ tw, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    load_packed<N, NT>(zbuf, x + (size_t)(xrow0 + rloc) * L, ((long long)j - 1) * N, L, vec_ok);
    __syncthreads();
    fft_forward<N, NT>(zbuf, tw);
    untangle_inplace<N, NT>(zbuf, tw, 1.f);
    __syncthreads();
    float2* out = Xs + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += NT) out[i] = ld_c(zbuf, i);
}

// UPOLS step 2: Y_j = sum_p X_{j-p} H_p, inverse FFT, keep the second half of the block
template <int N, int NT>
__global__ void __launch_bounds__(NT) fir_upols_kernel(const float2* __restrict__ Xs, const float2* __restrict__ Hs,
                                                       float* __restrict__ y, RowMap rm, int row0, int xrow0,
                                                       int hrow0, long long L, int P, int nblk, int shift,
                                                       const float4* __restrict__ tw, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
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
        st_c(zbuf, i, ar, ai);
    }
    __syncthreads();
    retangle_inplace<N, NT>(zbuf, tw);
    __syncthreads();
    fft_inverse<N, NT>(zbuf, tw);
    store_packed<N, NT>(zbuf, y + (size_t)row * L, N, N, (long long)j * N - shift, L, vec_ok);
}

__global__ void fft_plan_half_kernel(float2* ht, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const double a = (double)k / (double)n;
        ht[k] = make_float2((float)cospi(a), (float)(-sinpi(a)));
    }
}
__global__ void fft_plan_pass_kernel(float4* fwd, float4* inv, int M, int R) {
    const int ST = M / R;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (R - 1) * ST) {
        const int r = e / ST + 1, j = e - (r - 1) * ST;
        const double a = 2.0 * (double)((long long)j * r) / (double)M;  // angle = 2 pi j r / M
        const float wr = (float)cospi(a), wi = (float)(-sinpi(a));
        fwd[e] = make_float4(wr, wi, -wi, wr);
        inv[e] = make_float4(wr, -wi, wi, wr);
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
    const float4* tw; unsigned char* ws; size_t ws_bytes; cudaStream_t stream;
};

template <int N, int NT>
static int run_ols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    const int rows = a.batch * c_out, hrows = a.batch * a.ch;
    const size_t need = (size_t)hrows * N * sizeof(float2);
    if (!a.ws || a.ws_bytes < need) return GFX_ERR_WORKSPACE;
    float2* Hs = (float2*)a.ws;
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
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
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
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

size_t gfx_fft_plan_bytes(int n) {
    if (n != 1024 && n != 4096 && n != 16384) return 0;
    return ((size_t)n / 2 + 2 * (size_t)gfx::plan_total(n)) * sizeof(float4);
}

int gfx_fft_plan_init(void* plan, int n, void* stream) {
    if (!plan || (n != 1024 && n != 4096 && n != 16384)) return GFX_ERR_INVALID;
    float4* base = (float4*)plan;
    gfx::fft_plan_half_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float2*)plan, n);
    GFX_CUDA_CHECK(cudaGetLastError());
    for (int s = 0; s < 4; ++s) {
        const int entries = gfx::plan_entries(n, s);
        if (entries == 0) continue;
        float4* fwd = base + n / 2 + gfx::plan_offset(n, s);
        gfx::fft_plan_pass_kernel<<<(entries + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            fwd, fwd + gfx::plan_total(n), gfx::plan_m(n, s), gfx::plan_radix(n, s));
        GFX_CUDA_CHECK(cudaGetLastError());
    }
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
    FirArgs a{x, h, y, batch, cx, ch, L, filter_len, zerophase ? filter_len / 2 : 0, (const float4*)plan,
              (unsigned char*)workspace, workspace_bytes, (cudaStream_t)stream};
    const int n = pick_fft_size(filter_len);
    if (filter_len <= 16384) {
        if (n == 1024) return run_ols<1024, 128>(a);
        if (n == 4096) return run_ols<4096, 256>(a);
        return run_ols<16384, 1024>(a);
    }
    return run_upols<16384, 1024>(a);
}

}  // extern "C"
