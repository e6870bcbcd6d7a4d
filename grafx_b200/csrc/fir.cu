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
// On chip: a real block of F = 2n samples is packed into n complex points (even/odd) and transformed
// by an in-place mixed-radix decimation-in-frequency FFT in shared memory (packed (re,im) 64-bit
// slots, radix 16 / 16 / {-,2,4} / 16, output left digit-reversed); the matching decimation-in-time
// inverse brings it back -- no bit-reversal pass, no complex intermediate of the FFT itself in HBM.
// Half spectra live in HBM/L2 as "pair slots": float4 q = { A[k], A[n-k] } (the two bins the real-FFT
// untangling couples), so that untangle / multiply / retangle are one fused, coalesced step.
//   short filters (N <= 16384): one spectrum per filter row, blocks of F-N+1 new outputs, one kernel;
//   long filters (reverb IRs):  uniformly partitioned overlap-save, partition = n taps:
//       X-block and H-partition spectra (FFT kernels) -> per-bin multiply-accumulate over the
//       partitions (streaming kernel, filter and input history in registers: each spectrum is read
//       once) -> inverse FFT kernel.
// Twiddles: per pass only w^j, w^2j, w^3j (and w^4j, w^8j, w^12j for radix 16) come from a small
// L1-resident table built in double (gfx_fft_plan_init); the other radix-16 twiddles are products.
#include "common.cuh"
#include "fftcore.cuh"

namespace gfx {

// ------------------------------------------------------------------ radix plans
// n = 1024: 16,4,16   n = 4096: 16,16,16   n = 8192: 16,16,2,16   n = 16384: 16,16,4,16
// (every pass has stride 1 or a multiple of 16, the last radix is always 16)
__host__ __device__ constexpr bool plan_ok(int n) { return n == 1024 || n == 4096 || n == 8192 || n == 16384; }
__host__ __device__ constexpr int plan_stages(int n) { return (n == 1024 || n == 4096) ? 3 : 4; }
__host__ __device__ constexpr int plan_radix(int n, int s) {
    if (s < 0 || s >= plan_stages(n)) return 1;
    if (n == 1024) return s == 1 ? 4 : 16;
    if (n == 4096) return 16;
    if (n == 8192) return s == 2 ? 2 : 16;
    return s == 2 ? 4 : 16;
}
__host__ __device__ constexpr int plan_m(int n, int s) {  // sub-transform length entering pass s
    int m = n;
    for (int i = 0; i < s; ++i) m /= plan_radix(n, i);
    return m;
}
__host__ __device__ constexpr int tw_per_butterfly(int r) { return r == 16 ? 6 : (r == 4 ? 3 : (r == 2 ? 1 : 0)); }
__host__ __device__ constexpr int tw_exponent(int r, int e) {  // exponent of table entry e
    return r == 16 ? (e < 3 ? e + 1 : 4 * (e - 2)) : e + 1;
}
__host__ __device__ constexpr int plan_entries(int n, int s) {
    const int r = plan_radix(n, s);
    const int st = r > 1 ? plan_m(n, s) / r : 1;
    return (r > 1 && st > 1) ? tw_per_butterfly(r) * st : 0;
}
__host__ __device__ constexpr int plan_offset(int n, int s) {
    int o = 0;
    for (int i = 0; i < s; ++i) o += plan_entries(n, i);
    return o;
}
__host__ __device__ constexpr int plan_total(int n) { return plan_offset(n, plan_stages(n)); }
__host__ __device__ constexpr int ilog2c(int v) {
    int l = 0;
    while (v > 1) { v >>= 1; ++l; }
    return l;
}
// plan memory (float2 units): [ pair half-twiddles: n/2 | pass tables: plan_total(n) ]
//   pair table entry q = exp(-i pi k(q) / n), k(q) = the bin (< n/2) of pair slot q
//   pass table entry e * ST + j = omega_M^(j * tw_exponent(R, e)) = (cos, -sin): consecutive threads (j) read
//   consecutive 8-byte entries; the inverse uses the conjugate of the same table.

// digit reversal: position (mixed-radix digits, first radix most significant) <-> bin
template <int N>
__host__ __device__ __forceinline__ int pos_of_bin(int k) {
    int p = 0, kshift = 0, pshift = ilog2c(N);
#pragma unroll
    for (int s = 0; s < plan_stages(N); ++s) {
        const int lr = ilog2c(plan_radix(N, s));
        pshift -= lr;
        p |= ((k >> kshift) & ((1 << lr) - 1)) << pshift;
        kshift += lr;
    }
    return p;
}
template <int N>
__host__ __device__ __forceinline__ int bin_of_pos(int p) {
    int k = 0, kshift = 0, pshift = ilog2c(N);
#pragma unroll
    for (int s = 0; s < plan_stages(N); ++s) {
        const int lr = ilog2c(plan_radix(N, s));
        pshift -= lr;
        k |= ((p >> pshift) & ((1 << lr) - 1)) << kshift;
        kshift += lr;
    }
    return k;
}
// pair slot q < N/2 -> position of its bin k < N/2 (last digit < 8 because the last radix is 16)
__host__ __device__ __forceinline__ int pair_pos(int q) { return ((q >> 3) << 4) | (q & 7); }

// shared-memory layout: complex point i lives at 64-bit slot i + (i >> 4) (one pad slot per 16 points):
// every pass below is then bank-conflict free for 64-bit accesses.
__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int fft_smem_slots(int n) { return n + n / 16 + 16; }

template <int N, int NT, int S, bool INV>
__device__ __forceinline__ void fft_pass(pk2* z, const float2* __restrict__ plan) {
    constexpr int R = plan_radix(N, S);
    constexpr int M = plan_m(N, S);
    constexpr int ST = M / R;
    const float2* tp = plan + N / 2 + plan_offset(N, S);
    // slot of element q of a butterfly: pidx(i0 + q ST) = pidx(i0) + q * PST  (ST is a multiple of 16, or 1)
    constexpr int PST = ST >= 16 ? ST + ST / 16 : 1;
    static_assert(ST == 1 || ST % 16 == 0, "stride must keep the padding pattern linear");
#pragma unroll (R == 2 ? 4 : 1)
    for (int b = threadIdx.x; b < N / R; b += NT) {
        const int j = b & (ST - 1);
        const int i0 = (b - j) * R + j;
        pk2* zb = z + pidx(i0);
        if constexpr (R == 2) {
            float2 t1 = make_float2(1.f, 0.f);
            if constexpr (ST > 1) t1 = ld_tw(tp + j);
            pk2 a0 = zb[0], a1 = zb[PST];
            if constexpr (INV && ST > 1) a1 = tw_apply<true>(a1, t1);
            const pk2 s = pk_add(a0, a1);
            pk2 d = pk_sub(a0, a1);
            if constexpr (!INV && ST > 1) d = tw_apply<false>(d, t1);
            zb[0] = s; zb[PST] = d;
        } else if constexpr (R == 4) {
            float2 t1, t2, t3;
            if constexpr (ST > 1) { t1 = ld_tw(tp + j); t2 = ld_tw(tp + ST + j); t3 = ld_tw(tp + 2 * ST + j); }
            pk2 a0 = zb[0], a1 = zb[PST], a2 = zb[2 * PST], a3 = zb[3 * PST];
            if constexpr (INV && ST > 1) {
                a1 = tw_apply<true>(a1, t1); a2 = tw_apply<true>(a2, t2); a3 = tw_apply<true>(a3, t3);
            }
            r4<INV>(a0, a1, a2, a3);
            if constexpr (!INV && ST > 1) {
                a1 = tw_apply<false>(a1, t1); a2 = tw_apply<false>(a2, t2); a3 = tw_apply<false>(a3, t3);
            }
            zb[0] = a0; zb[PST] = a1; zb[2 * PST] = a2; zb[3 * PST] = a3;
        } else {
            // twiddle of index r = r0 + 4 r1:  omega^(j r0) * omega^(4 j r1);  lo[r0], hi[r1] from the table
            float2 lo[4], hi[4];
            if constexpr (ST > 1) {
#pragma unroll
                for (int e = 0; e < 3; ++e) { lo[e + 1] = ld_tw(tp + e * ST + j); hi[e + 1] = ld_tw(tp + (e + 3) * ST + j); }
            }
            pk2 a[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) a[q] = zb[q * PST];
            if constexpr (INV && ST > 1) {
#pragma unroll
                for (int q = 1; q < 16; ++q) {
                    const int r0 = q & 3, r1 = q >> 2;
                    const float2 w = r1 == 0 ? lo[r0] : (r0 == 0 ? hi[r1] : cmul(lo[r0], hi[r1]));
                    a[q] = tw_apply<true>(a[q], w);
                }
            }
            r16<INV>(a);
            // register a[4*i + jj] holds output index r = 4*jj + i
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int r = 4 * jj + i;
                    pk2 v = a[4 * i + jj];
                    if constexpr (!INV && ST > 1) {
                        if (r > 0) {
                            const float2 w = jj == 0 ? lo[i] : (i == 0 ? hi[jj] : cmul(lo[i], hi[jj]));
                            v = tw_apply<false>(v, w);
                        }
                    }
                    zb[r * PST] = v;
                }
            }
        }
    }
}

template <int N, int NT>
__device__ __forceinline__ void fft_forward(pk2* z, const float2* __restrict__ plan) {
    fft_pass<N, NT, 0, false>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 1, false>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 2, false>(z, plan);
    __syncthreads();
    if constexpr (plan_stages(N) > 3) {
        fft_pass<N, NT, 3, false>(z, plan);
        __syncthreads();
    }
}
template <int N, int NT>
__device__ __forceinline__ void fft_inverse(pk2* z, const float2* __restrict__ plan) {
    if constexpr (plan_stages(N) > 3) {
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

__device__ __forceinline__ float2 ld_c(const pk2* z, int i) {
    float re, im;
    pk_split(z[pidx(i)], re, im);
    return make_float2(re, im);
}
__device__ __forceinline__ void st_c(pk2* z, int i, float re, float im) { z[pidx(i)] = pk_make(re, im); }

// plan memory: [ pair half-twiddles: n/2 float2 | pass tables: plan_total(n) float2 | partner positions: n/2 u16 ]
//   partner[q] = position of bin n - k(q): the slot the real-FFT algebra couples with pair_pos(q)
template <int N>
__device__ __forceinline__ const unsigned short* plan_partner(const float2* plan) {
    return reinterpret_cast<const unsigned short*>(plan + N / 2 + plan_total(N));
}

// The three pair loops below run in batches of PB slots: all global / table loads of a batch are issued
// before the first use (the loops are latency-, not throughput-limited otherwise).
constexpr int PB = 4;

// smem FFT output -> pair slots in global memory (scaled)
template <int N, int NT>
__device__ __forceinline__ void untangle_store(const pk2* z, const float2* __restrict__ plan, float4* __restrict__ out,
                                               float scale) {
    const unsigned short* partner = plan_partner<N>(plan);
    static_assert((N / 2) % (NT * PB) == 0, "pair loop batches");
#pragma unroll 1
    for (int q0 = threadIdx.x; q0 < N / 2; q0 += NT * PB) {
        float2 w[PB];
        int pm[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            w[u] = __ldg(plan + q0 + u * NT);
            pm[u] = __ldg(partner + q0 + u * NT);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int q = q0 + u * NT;
            float4 o;
            if (q == 0) {
                const float2 z0 = ld_c(z, 0), zh = ld_c(z, 8);
                o = make_float4((z0.x + z0.y) * scale, (z0.x - z0.y) * scale, zh.x * scale, -zh.y * scale);
            } else {
                const PairA a = untangle_pair(ld_c(z, pair_pos(q)), ld_c(z, pm[u]), w[u]);
                o = make_float4(a.k.x * scale, a.k.y * scale, a.m.x * scale, a.m.y * scale);
            }
            out[q] = o;
        }
    }
}

// pair slots in global memory -> packed spectrum in smem, ready for the inverse FFT
template <int N, int NT>
__device__ __forceinline__ void retangle_load(pk2* z, const float2* __restrict__ plan, const float4* __restrict__ Y) {
    const unsigned short* partner = plan_partner<N>(plan);
    constexpr int YB = 8;
    static_assert((N / 2) % (NT * YB) == 0, "pair loop batches");
#pragma unroll 1
    for (int q0 = threadIdx.x; q0 < N / 2; q0 += NT * YB) {
        float4 y[YB];
        float2 w[YB];
        int pm[YB];
#pragma unroll
        for (int u = 0; u < YB; ++u) y[u] = ldg_stream(Y + q0 + u * NT);
#pragma unroll
        for (int u = 0; u < YB; ++u) {
            w[u] = __ldg(plan + q0 + u * NT);
            pm[u] = __ldg(partner + q0 + u * NT);
        }
#pragma unroll
        for (int u = 0; u < YB; ++u) {
            const int q = q0 + u * NT;
            if (q == 0) {
                st_c(z, 0, 0.5f * (y[u].x + y[u].y), 0.5f * (y[u].x - y[u].y));
                st_c(z, 8, y[u].z, -y[u].w);
            } else {
                const PairA a = retangle_pair(make_float2(y[u].x, y[u].y), make_float2(y[u].z, y[u].w), w[u]);
                st_c(z, pair_pos(q), a.k.x, a.k.y);
                st_c(z, pm[u], a.m.x, a.m.y);
            }
        }
    }
}

// fused in place: untangle X, multiply by the (already 1/N-scaled) filter spectrum H (pair slots), retangle
template <int N, int NT>
__device__ __forceinline__ void pointwise_filter(pk2* z, const float2* __restrict__ plan, const float4* __restrict__ H) {
    const unsigned short* partner = plan_partner<N>(plan);
#pragma unroll 1
    for (int q0 = threadIdx.x; q0 < N / 2; q0 += NT * PB) {
        float4 h[PB];
        float2 w[PB];
        int pm[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            h[u] = __ldg(H + q0 + u * NT);
            w[u] = __ldg(plan + q0 + u * NT);
            pm[u] = __ldg(partner + q0 + u * NT);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int q = q0 + u * NT;
            if (q == 0) {
                const float2 z0 = ld_c(z, 0), zh = ld_c(z, 8);
                const float y0 = (z0.x + z0.y) * h[u].x, yn = (z0.x - z0.y) * h[u].y;
                st_c(z, 0, 0.5f * (y0 + yn), 0.5f * (y0 - yn));
                // X_{N/2} = conj(Z), Y = X H, Z' = conj(Y) = Z conj(H)
                const float2 r2 = cmulc(zh, make_float2(h[u].z, h[u].w));
                st_c(z, 8, r2.x, r2.y);
            } else {
                const int p = pair_pos(q);
                const PairA x = untangle_pair(ld_c(z, p), ld_c(z, pm[u]), w[u]);
                const PairA y = retangle_pair(cmul(x.k, make_float2(h[u].x, h[u].y)),
                                              cmul(x.m, make_float2(h[u].z, h[u].w)), w[u]);
                st_c(z, p, y.k.x, y.k.y);
                st_c(z, pm[u], y.m.x, y.m.y);
            }
        }
    }
}

// ------------------------------------------------------------------ segment load / store
// loads F = 2N real samples src[s0 .. s0+F) (zero outside [0, len)) as z[j] = x[2j] + i x[2j+1];
// with src2 != nullptr the sample is src + sgn2 * src2 (mid/side -> left/right of a reverb IR).
template <int N, int NT>
__device__ __forceinline__ void load_packed(pk2* z, const float* __restrict__ src, const float* __restrict__ src2,
                                            float sgn2, long long s0, long long len, bool vec_ok) {
    if (vec_ok && (s0 & 3) == 0 && src2 == nullptr && s0 >= 0 && s0 + 2 * N <= len) {
        // interior segment: asynchronous copies straight into the packed layout, the whole segment in flight
        // at once (two 8-byte halves per 16 bytes: the padded slots are only 8-byte aligned); the slot of
        // t = tid + i NT is linear in i (NT is a multiple of 8), so the unrolled loop has no index math
        // one complex point (8 bytes) per copy, consecutive lanes -> consecutive points: a warp's copy is
        // contiguous in global memory and (up to one pad slot) in shared memory: no bank conflicts
        static_assert(NT % 16 == 0 && N % NT == 0, "load tiling");
        const float* g = src + s0 + 2 * (int)threadIdx.x;
        const uint32_t d = smem_u32(z) + 8u * (uint32_t)pidx((int)threadIdx.x);
#pragma unroll
        for (int i = 0; i < N / NT; ++i)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d + 8u * (uint32_t)(i * (NT + NT / 16))),
                         "l"(g + i * 2 * NT));
        cp_async_commit();
        cp_async_wait<0>();
    } else if (vec_ok && (s0 & 3) == 0 && src2 == nullptr) {
        // segment crossing the start / end of the row: same copies with zero fill
#pragma unroll 4
        for (int t = threadIdx.x; t < N / 2; t += NT) {
            const long long pos = s0 + 4LL * t;
            const long long rem = (pos >= 0) ? (len - pos) : 0;  // samples available from pos on (a 4-group never straddles 0)
            const int b0 = rem >= 2 ? 8 : (rem > 0 ? (int)rem * 4 : 0);
            const int b1 = rem >= 4 ? 8 : (rem > 2 ? (int)(rem - 2) * 4 : 0);
            const float* g = (rem > 0) ? src + pos : src;
            cp_async8(&z[pidx(2 * t)], g, b0);
            cp_async8(&z[pidx(2 * t + 1)], b1 > 0 ? g + 2 : src, b1);
        }
        cp_async_commit();
        cp_async_wait<0>();
    } else if (vec_ok && (s0 & 3) == 0) {
#pragma unroll 4
        for (int t = threadIdx.x; t < N / 2; t += NT) {
            const long long pos = s0 + 4LL * t;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos >= 0 && pos + 4 <= len) {
                v = ldg_stream(reinterpret_cast<const float4*>(src + pos));
                if (src2) {
                    const float4 u = ldg_stream(reinterpret_cast<const float4*>(src2 + pos));
                    v.x = fmaf(sgn2, u.x, v.x); v.y = fmaf(sgn2, u.y, v.y); v.z = fmaf(sgn2, u.z, v.z); v.w = fmaf(sgn2, u.w, v.w);
                }
            } else if (pos + 4 > 0 && pos < len) {
                float e[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool in = (pos + c >= 0 && pos + c < len);
                    e[c] = in ? src[pos + c] : 0.f;
                    if (in && src2) e[c] = fmaf(sgn2, src2[pos + c], e[c]);
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            z[pidx(2 * t)] = pk_make(v.x, v.y);
            z[pidx(2 * t + 1)] = pk_make(v.z, v.w);
        }
    } else {
        for (int j = threadIdx.x; j < N; j += NT) {
            const long long pos = s0 + 2LL * j;
            float a = 0.f, b = 0.f;
            if (pos >= 0 && pos < len) { a = src[pos]; if (src2) a = fmaf(sgn2, src2[pos], a); }
            if (pos + 1 >= 0 && pos + 1 < len) { b = src[pos + 1]; if (src2) b = fmaf(sgn2, src2[pos + 1], b); }
            z[pidx(j)] = pk_make(a, b);
        }
    }
}

// writes segment samples [i_lo, i_lo + count) to dst[d0 .. d0+count), clipped to [0, len)
template <int N, int NT>
__device__ __forceinline__ void store_packed(const pk2* z, float* __restrict__ dst, int i_lo, int count, long long d0,
                                             long long len, bool vec_ok) {
    if (vec_ok && (i_lo & 3) == 0 && (d0 & 3) == 0 && (count & 3) == 0 && d0 >= 0 && d0 + count <= len) {
        // interior block: no clipping, 32-bit index math
        float* o = dst + d0;
        const int h_lo = i_lo >> 1;
#pragma unroll 4
        for (int t = threadIdx.x; t < count / 4; t += NT) {
            const int c = h_lo + 2 * t;  // complex index of the first of the two packed points
            float a0, a1, b0, b1;
            pk_split(z[pidx(c)], a0, a1);
            pk_split(z[pidx(c + 1)], b0, b1);
            stg_stream(reinterpret_cast<float4*>(o + 4 * t), make_float4(a0, a1, b0, b1));
        }
    } else if (vec_ok && (i_lo & 3) == 0 && (d0 & 3) == 0 && (count & 3) == 0) {
#pragma unroll 4
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

// filter rows, optionally with the unit-energy normalisation of a reverb IR folded into the spectra
// (reference: normalize_impulse after the optional ms_to_lr, reverb.py:215-228, core/utils.py:14-18)
struct FilterSrc {
    const float* h;       // [hrows, Nh]
    const float* energy;  // [batch, 2] sum_t of the squared raw mid / side rows, or null (plain filter)
    int to_lr;            // with energy: rows are left/right = mid +- side (mean-over-channels energy = e0 + e1)
};

constexpr int fir_min_blocks(int n) { return n <= 4096 ? 3 : (n <= 8192 ? 3 : 1); }

// ------------------------------------------------------------------ kernels
// spectra of filter partitions: Hs[(hrow * P + part) * N/2 + q] pair slots, scaled by 1/N (and the energy norm)
template <int N, int NT>
__global__ void __launch_bounds__(NT, fir_min_blocks(N)) fir_spectrum_kernel(FilterSrc fs, float4* __restrict__ Hs,
                                                                             int hrow0, int Nh, int part_len, int P,
                                                                             const float2* __restrict__ plan, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int hloc = blockIdx.x / P, part = blockIdx.x - hloc * P;
    const int hrow = hrow0 + hloc;
    const size_t off = (size_t)hrow * Nh + (size_t)part * part_len;
    long long len = (long long)Nh - (long long)part * part_len;
    if (len > part_len) len = part_len;
    float scale = 1.f / (float)N;
    if (fs.energy) {
        // normalize_impulse of a reverb IR: one scale per batch item from the raw mid/side energies
        const float e0 = fs.energy[hrow & ~1], e1 = fs.energy[hrow | 1];
        scale *= fs.to_lr ? rsqrtf(e0 + e1 + 1e-12f) : rsqrtf(0.5f * (e0 + e1) + 1e-12f);
    }
    const bool v = vec_ok && ((off & 3) == 0);
    load_packed<N, NT>(zbuf, fs.h + off, nullptr, 0.f, 0, len, v);
    __syncthreads();
    fft_forward<N, NT>(zbuf, plan);
    untangle_store<N, NT>(zbuf, plan, Hs + (size_t)blockIdx.x * (N / 2), scale);
}

// single-partition overlap-save: block j produces full-convolution samples [j*hop, (j+1)*hop)
template <int N, int NT>
__global__ void __launch_bounds__(NT, fir_min_blocks(N)) fir_ols_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                        const float4* __restrict__ Hs, RowMap rm,
                                                                        long long L, int pre, int hop, int shift, int nblk,
                                                                        const float2* __restrict__ plan, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int row = blockIdx.x / nblk, j = blockIdx.x - row * nblk;
    int xr, hr;
    rm.map(row, xr, hr);
    const long long m0 = (long long)j * hop;
    load_packed<N, NT>(zbuf, x + (size_t)xr * L, nullptr, 0.f, m0 - pre, L, vec_ok);
    __syncthreads();
    fft_forward<N, NT>(zbuf, plan);
    pointwise_filter<N, NT>(zbuf, plan, Hs + (size_t)hr * (N / 2));
    __syncthreads();
    fft_inverse<N, NT>(zbuf, plan);
    store_packed<N, NT>(zbuf, y + (size_t)row * L, pre, hop, m0 - shift, L, vec_ok);
}

// UPOLS step 1: spectra of input blocks.  Xs[(rloc * nblk + j) * N/2 + q] = rfft of x[(j-1)B, (j+1)B), B = N
template <int N, int NT>
__global__ void __launch_bounds__(NT, fir_min_blocks(N)) fir_xspec_kernel(const float* __restrict__ x, float4* __restrict__ Xs,
                                                                          int xrow0, long long L, int nblk,
                                                                          const float2* __restrict__ plan, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    load_packed<N, NT>(zbuf, x + (size_t)(xrow0 + rloc) * L, nullptr, 0.f, ((long long)j - 1) * N, L, vec_ok);
    __syncthreads();
    fft_forward<N, NT>(zbuf, plan);
    untangle_store<N, NT>(zbuf, plan, Xs + (size_t)blockIdx.x * (N / 2), 1.f);
}

// UPOLS step 2: Y_j = sum_p X_{j-p} H_p on pair slots, partitions [p0, p0 + PC).  One thread owns one pair slot of
// one output row and walks the blocks in order: the PC filter values and the last PC input values stay in
// registers, so every spectrum is read exactly once.  accumulate != 0 adds to Ys (filters with more than
// MAC_MAX_PC partitions are processed in groups).
constexpr int MAC_MAX_PC = 12;
constexpr int MAC_NT = 256;

__device__ __forceinline__ void cmac4(float4& acc, const float4& x, const float4& h) {
    acc.x = fmaf(x.x, h.x, acc.x); acc.x = fmaf(-x.y, h.y, acc.x);
    acc.y = fmaf(x.x, h.y, acc.y); acc.y = fmaf(x.y, h.x, acc.y);
    acc.z = fmaf(x.z, h.z, acc.z); acc.z = fmaf(-x.w, h.w, acc.z);
    acc.w = fmaf(x.z, h.w, acc.w); acc.w = fmaf(x.w, h.z, acc.w);
}

template <int PC>
__global__ void __launch_bounds__(MAC_NT, 2) fir_mac_kernel(const float4* __restrict__ Xs, const float4* __restrict__ Hs,
                                                            float4* __restrict__ Ys, RowMap rm, int xrow0, int hrow0,
                                                            int row0, int P, int p0, int nblk, int half, int accumulate) {
    const int q = blockIdx.x * MAC_NT + threadIdx.x;
    const int rloc = blockIdx.y;
    int xr, hr;
    rm.map(row0 + rloc, xr, hr);
    const float4* X = Xs + (size_t)(xr - xrow0) * nblk * half + q;
    const float4* H = Hs + ((size_t)(hr - hrow0) * P + p0) * half + q;
    float4* Y = Ys + (size_t)rloc * nblk * half + q;
    float4 h[PC], ring[PC];
#pragma unroll
    for (int p = 0; p < PC; ++p) {
        h[p] = __ldg(H + (size_t)p * half);
        ring[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (q == 0) {
        // slot 0 carries (A_0, A_N), two REAL bins, in .xy: not a complex product -- warp 0 adds them below
#pragma unroll
        for (int p = 0; p < PC; ++p) { h[p].x = 0.f; h[p].y = 0.f; }
    }
    constexpr int G = PC < 4 ? PC : (PC >= 10 ? 2 : 4);  // loads issued together
    const int last = nblk - 1 - p0;     // last valid input block index for this partition group
#pragma unroll 1
    for (int j0 = p0; j0 < nblk; j0 += PC) {
#pragma unroll
        for (int g0 = 0; g0 < PC; g0 += G) {
            float4 xn[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                int jb = j0 - p0 + g0 + g;  // input block feeding output block j0 + g0 + g
                jb = jb < last ? jb : last;
                xn[g] = ldg_stream(X + (size_t)jb * half);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (g0 + g < PC) {
                    const int jj = g0 + g;
                    const int j = j0 + jj;
                    ring[jj] = xn[g];
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < nblk) {
                        if (accumulate) acc = Y[(size_t)j * half];
#pragma unroll
                        for (int p = 0; p < PC; ++p) cmac4(acc, ring[(jj - p + PC) % PC], h[p]);
                        Y[(size_t)j * half] = acc;
                    }
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // DC / Nyquist of every block (the .xy lanes of slot 0): lanes over output blocks
        __syncwarp();
        const float4* X0 = X - q;
        const float4* H0 = H - q;
        float4* Y0 = Y - q;
        for (int j = p0 + (int)threadIdx.x; j < nblk; j += 32) {
            float a0 = 0.f, an = 0.f;
            for (int p = 0; p < PC && p0 + p <= j; ++p) {
                const float4 xv = __ldg(X0 + (size_t)(j - p0 - p) * half);
                const float4 hv = __ldg(H0 + (size_t)p * half);
                a0 = fmaf(xv.x, hv.x, a0);
                an = fmaf(xv.y, hv.y, an);
            }
            // lane 0 (q == 0) wrote slot 0 of block j in the main loop; the __syncwarp above orders that store
            Y0[(size_t)j * half].x += a0;
            Y0[(size_t)j * half].y += an;
        }
    }
}

// UPOLS step 3: inverse FFT of Y_j, keep the second half of the block
template <int N, int NT>
__global__ void __launch_bounds__(NT, fir_min_blocks(N)) fir_inv_kernel(const float4* __restrict__ Ys, float* __restrict__ y,
                                                                        int row0, long long L, int nblk, int shift,
                                                                        const float2* __restrict__ plan, int vec_ok) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    retangle_load<N, NT>(zbuf, plan, Ys + (size_t)blockIdx.x * (N / 2));
    __syncthreads();
    fft_inverse<N, NT>(zbuf, plan);
    store_packed<N, NT>(zbuf, y + (size_t)(row0 + rloc) * L, N, N, (long long)j * N - shift, L, vec_ok);
}

// ------------------------------------------------------------------ plan construction
template <int N>
__global__ void fft_plan_pairs_kernel(float2* ht) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < N / 2) {
        const int k = bin_of_pos<N>(pair_pos(q));
        const double a = (double)k / (double)N;
        ht[q] = make_float2((float)cospi(a), (float)(-sinpi(a)));
    }
}
template <int N>
__global__ void fft_plan_partner_kernel(unsigned short* partner) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < N / 2) partner[q] = q == 0 ? 8 : (unsigned short)pos_of_bin<N>(N - bin_of_pos<N>(pair_pos(q)));
}
__global__ void fft_plan_pass_kernel(float2* tab, int M, int R) {
    const int ST = M / R;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < tw_per_butterfly(R) * ST) {
        const int ei = e / ST, j = e - ei * ST;
        const double a = 2.0 * (double)((long long)j * tw_exponent(R, ei)) / (double)M;  // angle = 2 pi j r / M
        tab[e] = make_float2((float)cospi(a), (float)(-sinpi(a)));
    }
}

// ------------------------------------------------------------------ host side
static int g_long_n = 8192;  // partition size of the long-filter path (tunable: gfx_fir_set_tuning)
static int g_mid_n = 8192;   // FFT size for 2048 < taps <= g_mid_n / 2 (longer single-partition filters: 16384)

static int pick_fft_size(int Nh) {
    if (Nh <= 512) return 1024;
    if (Nh <= 2048) return 4096;
    if (Nh <= 16384) return Nh <= g_mid_n / 2 ? g_mid_n : 16384;
    return g_long_n;
}

template <typename K>
static int set_smem(K kern, size_t smem) {
    GFX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return GFX_OK;
}

struct FirArgs {
    const float* x; FilterSrc fs; float* y;
    int batch, cx, ch; long long L; int Nh; int shift;
    const float2* plan; unsigned char* ws; size_t ws_bytes; cudaStream_t stream;
};

template <int N, int NT>
static int run_ols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    const int rows = a.batch * c_out, hrows = a.batch * a.ch;
    const size_t need = (size_t)hrows * (N / 2) * sizeof(float4);
    if (!a.ws || a.ws_bytes < need) return GFX_ERR_WORKSPACE;
    float4* Hs = (float4*)a.ws;
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
    static bool configured = false;
    if (!configured) {
        if (set_smem(fir_spectrum_kernel<N, NT>, smem) || set_smem(fir_ols_kernel<N, NT>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    const int hvec = ((uintptr_t)a.fs.h % 16 == 0);
    fir_spectrum_kernel<N, NT><<<hrows, NT, smem, a.stream>>>(a.fs, Hs, 0, a.Nh, a.Nh, 1, a.plan, hvec);
    GFX_LAUNCH_CHECK();
    const int pre = (a.Nh - 1 + 3) & ~3;
    const int hop = (2 * N - pre) & ~3;
    const long long total = a.L + a.shift;
    const long long nblk = (total + hop - 1) / hop;
    if (nblk * rows > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
    const int vec = (((uintptr_t)a.x | (uintptr_t)a.y) % 16 == 0) && (a.L % 4 == 0);
    RowMap rm{c_out, a.cx, a.ch};
    fir_ols_kernel<N, NT><<<(unsigned)(nblk * rows), NT, smem, a.stream>>>(a.x, a.y, Hs, rm, a.L, pre, hop, a.shift,
                                                                           (int)nblk, a.plan, vec);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

static void upols_geometry(int cx, int ch, long long L, int Nh, int shift, int N, int& P, long long& nblk,
                           size_t& per_item_bytes) {
    const int c_out = cx > ch ? cx : ch;
    P = (Nh + N - 1) / N;
    nblk = (L + shift + N - 1) / N;
    per_item_bytes = ((size_t)ch * P + (size_t)(cx + c_out) * nblk) * (N / 2) * sizeof(float4);
}

template <int PC>
static void launch_mac(dim3 grid, cudaStream_t st, const float4* Xs, const float4* Hs, float4* Ys, RowMap rm, int xrow0,
                       int hrow0, int row0, int P, int p0, int nblk, int half, int acc) {
    fir_mac_kernel<PC><<<grid, MAC_NT, 0, st>>>(Xs, Hs, Ys, rm, xrow0, hrow0, row0, P, p0, nblk, half, acc);
}

template <int N, int NT>
static int run_upols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    int P; long long nblk; size_t per_item;
    upols_geometry(a.cx, a.ch, a.L, a.Nh, a.shift, N, P, nblk, per_item);
    if (!a.ws || a.ws_bytes < per_item) return GFX_ERR_WORKSPACE;
    long long chunk = (long long)(a.ws_bytes / per_item);
    if (chunk > a.batch) chunk = a.batch;
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
    static bool configured = false;
    if (!configured) {
        if (set_smem(fir_spectrum_kernel<N, NT>, smem) || set_smem(fir_xspec_kernel<N, NT>, smem) ||
            set_smem(fir_inv_kernel<N, NT>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    constexpr int half = N / 2;
    const int hvec = ((uintptr_t)a.fs.h % 16 == 0);
    const int vec = (((uintptr_t)a.x | (uintptr_t)a.y) % 16 == 0) && (a.L % 4 == 0);
    RowMap rm{c_out, a.cx, a.ch};
    for (long long b0 = 0; b0 < a.batch; b0 += chunk) {
        const int nb = (int)((a.batch - b0 < chunk) ? a.batch - b0 : chunk);
        float4* Hs = (float4*)a.ws;
        float4* Xs = Hs + (size_t)nb * a.ch * P * half;
        float4* Ys = Xs + (size_t)nb * a.cx * nblk * half;
        const int hrow0 = (int)b0 * a.ch, xrow0 = (int)b0 * a.cx, row0 = (int)b0 * c_out;
        if ((long long)nb * a.cx * nblk > 0x7fffffffLL || (long long)nb * c_out * nblk > 0x7fffffffLL ||
            (long long)nb * c_out > 65535) return GFX_ERR_UNSUPPORTED;
        fir_spectrum_kernel<N, NT><<<nb * a.ch * P, NT, smem, a.stream>>>(a.fs, Hs, hrow0, a.Nh, N, P, a.plan, hvec);
        GFX_LAUNCH_CHECK();
        fir_xspec_kernel<N, NT><<<(unsigned)(nb * a.cx * nblk), NT, smem, a.stream>>>(a.x, Xs, xrow0, a.L, (int)nblk,
                                                                                      a.plan, vec);
        GFX_LAUNCH_CHECK();
        const dim3 grid(half / MAC_NT, nb * c_out);
        for (int p0 = 0; p0 < P; p0 += MAC_MAX_PC) {
            const int pc = P - p0 < MAC_MAX_PC ? P - p0 : MAC_MAX_PC;
            const int acc = p0 > 0;
#define GFX_MAC_CASE(PCV) case PCV: launch_mac<PCV>(grid, a.stream, Xs, Hs, Ys, rm, xrow0, hrow0, row0, P, p0, (int)nblk, half, acc); break;
            switch (pc) {
                GFX_MAC_CASE(1) GFX_MAC_CASE(2) GFX_MAC_CASE(3) GFX_MAC_CASE(4) GFX_MAC_CASE(5) GFX_MAC_CASE(6)
                GFX_MAC_CASE(7) GFX_MAC_CASE(8) GFX_MAC_CASE(9) GFX_MAC_CASE(10) GFX_MAC_CASE(11) GFX_MAC_CASE(12)
            }
#undef GFX_MAC_CASE
            GFX_LAUNCH_CHECK();
        }
        fir_inv_kernel<N, NT><<<(unsigned)(nb * c_out * nblk), NT, smem, a.stream>>>(Ys, a.y, row0, a.L, (int)nblk,
                                                                                     a.shift, a.plan, vec);
        GFX_LAUNCH_CHECK();
    }
    return GFX_OK;
}

static int fir_dispatch(const FirArgs& a) {
    const int n = pick_fft_size(a.Nh);
    if (a.Nh <= 16384) {
        if (n == 1024) return run_ols<1024, 128>(a);
        if (n == 4096) return run_ols<4096, 256>(a);
        if (n == 8192) return run_ols<8192, 256>(a);
        return run_ols<16384, 512>(a);
    }
    if (n == 4096) return run_upols<4096, 256>(a);
    if (n == 8192) return run_upols<8192, 256>(a);
    return run_upols<16384, 512>(a);
}

}  // namespace gfx

extern "C" {

int gfx_fir_set_tuning(int long_n, int mid_n) {
    if (long_n) {
        if (long_n != 4096 && long_n != 8192 && long_n != 16384) return GFX_ERR_INVALID;
        gfx::g_long_n = long_n;
    }
    if (mid_n) {
        if (mid_n != 8192 && mid_n != 16384) return GFX_ERR_INVALID;
        gfx::g_mid_n = mid_n;
    }
    return GFX_OK;
}

int gfx_fir_fft_size(int filter_len) { return filter_len <= 0 ? GFX_ERR_INVALID : gfx::pick_fft_size(filter_len); }

size_t gfx_fft_plan_bytes(int n) {
    if (!gfx::plan_ok(n)) return 0;
    return ((size_t)n / 2 + (size_t)gfx::plan_total(n)) * sizeof(float2) + ((size_t)n / 2) * sizeof(unsigned short);
}

int gfx_fft_plan_init(void* plan, int n, void* stream) {
    using namespace gfx;
    if (!plan || !plan_ok(n)) return GFX_ERR_INVALID;
    float2* base = (float2*)plan;
    cudaStream_t st = (cudaStream_t)stream;
    const int gb = (n / 2 + 255) / 256;
    if (n == 1024) fft_plan_pairs_kernel<1024><<<gb, 256, 0, st>>>(base);
    else if (n == 4096) fft_plan_pairs_kernel<4096><<<gb, 256, 0, st>>>(base);
    else if (n == 8192) fft_plan_pairs_kernel<8192><<<gb, 256, 0, st>>>(base);
    else fft_plan_pairs_kernel<16384><<<gb, 256, 0, st>>>(base);
    GFX_LAUNCH_CHECK();
    unsigned short* partner = (unsigned short*)(base + n / 2 + plan_total(n));
    if (n == 1024) fft_plan_partner_kernel<1024><<<gb, 256, 0, st>>>(partner);
    else if (n == 4096) fft_plan_partner_kernel<4096><<<gb, 256, 0, st>>>(partner);
    else if (n == 8192) fft_plan_partner_kernel<8192><<<gb, 256, 0, st>>>(partner);
    else fft_plan_partner_kernel<16384><<<gb, 256, 0, st>>>(partner);
    GFX_LAUNCH_CHECK();
    for (int s = 0; s < plan_stages(n); ++s) {
        const int entries = plan_entries(n, s);
        if (entries == 0) continue;
        fft_plan_pass_kernel<<<(entries + 255) / 256, 256, 0, st>>>(base + n / 2 + plan_offset(n, s), plan_m(n, s),
                                                                   plan_radix(n, s));
        GFX_LAUNCH_CHECK();
    }
    return GFX_OK;
}

size_t gfx_fir_conv_workspace_bytes(int batch, int cx, int ch, long long L, int filter_len, int zerophase) {
    if (batch <= 0 || cx <= 0 || ch <= 0 || L <= 0 || filter_len <= 0) return 0;
    const int n = gfx::pick_fft_size(filter_len);
    if (filter_len <= 16384) return (size_t)batch * ch * (n / 2) * sizeof(float4);
    int P; long long nblk; size_t per_item;
    gfx::upols_geometry(cx, ch, L, filter_len, zerophase ? filter_len / 2 : 0, n, P, nblk, per_item);
    // spectra of up to ~1.5 GB worth of batch items per sweep (every kernel of a sweep then has several full waves)
    size_t items = ((size_t)1536 << 20) / per_item;
    if (items < 1) items = 1;
    if (items > (size_t)batch) items = batch;
    return items * per_item;
}

static int fir_conv_common(const float* x, gfx::FilterSrc fs, float* y, int batch, int cx, int ch, long long L,
                           int filter_len, int zerophase, const void* plan, void* workspace, size_t workspace_bytes,
                           void* stream) {
    using namespace gfx;
    if (!x || !fs.h || !y || !plan) return GFX_ERR_INVALID;
    if (batch <= 0 || cx <= 0 || ch <= 0 || L <= 0 || filter_len <= 0) return GFX_ERR_INVALID;
    if (cx != ch && cx != 1 && ch != 1) return GFX_ERR_INVALID;
    FirArgs a{x, fs, y, batch, cx, ch, L, filter_len, zerophase ? filter_len / 2 : 0, (const float2*)plan,
              (unsigned char*)workspace, workspace_bytes, (cudaStream_t)stream};
    return fir_dispatch(a);
}

int gfx_fir_conv_f32(const float* x, const float* h, float* y, int batch, int cx, int ch, long long L,
                     int filter_len, int zerophase, const void* plan, void* workspace, size_t workspace_bytes,
                     void* stream) {
    return fir_conv_common(x, gfx::FilterSrc{h, nullptr, 0}, y, batch, cx, ch, L, filter_len, zerophase, plan,
                           workspace, workspace_bytes, stream);
}

int gfx_fir_conv_midside_ir_f32(const float* x, const float* ir_raw, const float* energy, float* y, int batch, int cx,
                                long long L, int ir_len, int ms_to_lr, const void* plan, void* workspace,
                                size_t workspace_bytes, void* stream) {
    if (!energy) return GFX_ERR_INVALID;
    return fir_conv_common(x, gfx::FilterSrc{ir_raw, energy, ms_to_lr ? 1 : 0}, y, batch, cx, 2, L, ir_len, 0, plan,
                           workspace, workspace_bytes, stream);
}

}  // extern "C"
