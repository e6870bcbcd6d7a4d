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
// The first forward pass reads its inputs straight from global memory, the last inverse pass writes straight to
// global memory, and the stride-1 pass is fused in registers with the real-FFT pair algebra (and, for short
// filters, with the spectral multiply): a block costs 2 passes' worth of shared-memory traffic less per direction.
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
#ifndef GFX_TW15
#define GFX_TW15 0  // 1: all 15 radix-16 twiddles from the table; 0: 6 from the table, 9 products
#endif
__host__ __device__ constexpr int tw_per_butterfly(int r) { return r == 16 ? (GFX_TW15 ? 15 : 6) : (r == 4 ? 3 : (r == 2 ? 1 : 0)); }
__host__ __device__ constexpr int tw_exponent(int r, int e) {  // exponent of table entry e
    return (r == 16 && !GFX_TW15) ? (e < 3 ? e + 1 : 4 * (e - 2)) : e + 1;
}
#ifndef GFX_PASS_UNROLL
#define GFX_PASS_UNROLL 1
#endif
constexpr int kPassUnroll = GFX_PASS_UNROLL;
// GUNR (fft_pass): butterflies of a global-memory pass issued together.  2 in the one-transform kernels (spectra, inverse:
// both butterflies' 32 loads / stores of a thread in flight, reverb-shape convolution 1.66 -> 1.59 ms on B200); 1 in the
// fused forward + inverse kernel (fir_ols_kernel: 128 registers, the second set spills, 0.367 -> 0.388 ms).  // butterflies of a shared-memory radix-16 pass processed together
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

// Thread organisation: NT = N/32 threads per transform; every radix-16 pass has N/16 = 2 NT butterflies
// (two per thread).  The first pass takes its inputs straight from global memory, the last inverse pass
// writes its outputs straight to global memory, and the last forward / first inverse pass (stride 1: a
// butterfly = 16 consecutive positions) is fused, in registers, with the real-FFT pair algebra:
//   position 16 b + r holds bin  klow(b) + (N/16) r;   bin N - k sits in butterfly b' = b(N/16 - klow) at 15 - r,
// so a thread that owns the two butterflies {bA, bB = partner(bA)} (plan "pair table", thread 0 owns the two
// self-paired ones) holds both members of 16 pairs.  Pair slot numbering (shared by every spectrum in HBM):
//   slot(t, half, rk) = half N/4 + rk N/32 + t    (half 0: bin klow(bA) + (N/16) rk, rk < 8;  half 1: same for bB)
// -> for fixed (half, rk) consecutive threads touch consecutive float4: coalesced.  Slot 0 = DC/Nyquist.
#define OUT16(r) (4 * ((r) & 3) + ((r) >> 2))  // register of r16<>'s output index r

// plan memory: [ pair half-twiddles by slot: n/2 float2 | pass tables: plan_total(n) float2 | pair table: n/32 ushort2 ]
template <int N>
__device__ __forceinline__ const ushort2* plan_pairtab(const float2* plan) {
    return reinterpret_cast<const ushort2*>(plan + N / 2 + plan_total(N));
}

// ---- sources / sinks of the fused first / last passes
__device__ __forceinline__ pk2 ldg_pk(const float* p) {
    pk2 r;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_pk(float* p, pk2 v) {
    asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(v));
}

// reads block-relative real samples (2c, 2c+1) of a row segment starting at s0 (zero outside [0, len)).
// FAST: every (2c, 2c+1) pair is 8-byte aligned and wholly inside or outside the row (the host checks
// pointer alignment and the parity of every offset involved) -> one predicated 64-bit load, no branches.
template <bool FAST>
struct SegSrc {
    const float* base;  // row + s0
    int lo, hi;         // valid block-relative sample range [lo, hi)
    __device__ __forceinline__ SegSrc(const float* row, long long s0, long long len, int span) {
        base = row + s0;
        const long long l = s0 < 0 ? -s0 : 0, h = len - s0;
        lo = l > span ? span : (int)l;
        hi = h > span ? span : (h < 0 ? 0 : (int)h);
    }
    __device__ __forceinline__ pk2 get(int c) const {
        const int i = 2 * c;
        if constexpr (FAST) {
            pk2 r = 0ull;
            if (i >= lo && i < hi) r = ldg_pk(base + i);
            return r;
        } else {
            const float a = (i >= lo && i < hi) ? __ldg(base + i) : 0.f;
            const float b = (i + 1 >= lo && i + 1 < hi) ? __ldg(base + i + 1) : 0.f;
            return pk_make(a, b);
        }
    }
};
// writes block-relative real samples (2c, 2c+1), restricted to the window [w_lo, w_hi) of the block, to
// dst[pos0 + (i - w_lo)], clipped to [0, len)
template <bool FAST>
struct SegDst {
    float* base;  // row + pos0 - w_lo  (indexable by the block-relative sample index)
    int lo, hi;   // writable block-relative sample range
    __device__ __forceinline__ SegDst(float* row, long long pos0, long long len, int w_lo, int w_hi) {
        const long long off = pos0 - w_lo;  // global position of block-relative sample 0
        base = row + off;
        long long l = -off, h = len - off;
        if (l < w_lo) l = w_lo;
        if (h > w_hi) h = w_hi;
        if (h < l) h = l;
        lo = (int)l; hi = (int)h;
    }
    __device__ __forceinline__ void put(int c, pk2 v) const {
        const int i = 2 * c;
        if constexpr (FAST) {
            if (i >= lo && i < hi) stg_pk(base + i, v);
        } else {
            float a, b;
            pk_split(v, a, b);
            if (i >= lo && i < hi) base[i] = a;
            if (i + 1 >= lo && i + 1 < hi) base[i + 1] = b;
        }
    }
};
struct NoSrc { __device__ __forceinline__ pk2 get(int) const { return 0ull; } };
struct NoDst { __device__ __forceinline__ void put(int, pk2) const {} };

// One FFT pass over shared memory (see the plan tables).  FROM_GLOBAL (forward pass 0 only): inputs come from
// `src` (complex index c = b + (N/16) q); TO_GLOBAL (inverse pass 0 only): outputs go to `dst`.
template <int N, int NT, int S, bool INV, bool FROM_GLOBAL = false, bool TO_GLOBAL = false, typename Src = NoSrc,
          typename Dst = NoDst, int GUNR = 1>
__device__ __forceinline__ void fft_pass(pk2* z, const float2* __restrict__ plan, const Src& src = Src(),
                                         const Dst& dst = Dst()) {
    constexpr int R = plan_radix(N, S);
    constexpr int M = plan_m(N, S);
    constexpr int ST = M / R;
    const float2* tp = plan + N / 2 + plan_offset(N, S);
    // slot of element q of a butterfly: pidx(i0 + q ST) = pidx(i0) + q * PST  (ST is a multiple of 16, or 1)
    constexpr int PST = ST >= 16 ? ST + ST / 16 : 1;
    static_assert(ST == 1 || ST % 16 == 0, "stride must keep the padding pattern linear");
    static_assert(!(FROM_GLOBAL || TO_GLOBAL) || (S == 0 && R == 16), "fused passes are the radix-16 pass 0");
#pragma unroll(R == 2 ? 4 : ((FROM_GLOBAL || TO_GLOBAL) ? GUNR : kPassUnroll))
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
            float2 lo[4], hi[4], tw[16];
            if constexpr (ST > 1) {
                if constexpr (GFX_TW15) {
#pragma unroll
                    for (int e = 1; e < 16; ++e) tw[e] = ld_tw(tp + (e - 1) * ST + j);
                } else {
#pragma unroll
                    for (int e = 0; e < 3; ++e) { lo[e + 1] = ld_tw(tp + e * ST + j); hi[e + 1] = ld_tw(tp + (e + 3) * ST + j); }
                }
            }
            pk2 a[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if constexpr (FROM_GLOBAL) a[q] = src.get(b + ST * q);
                else a[q] = zb[q * PST];
            }
            if constexpr (INV && ST > 1) {
#pragma unroll
                for (int q = 1; q < 16; ++q) {
                    const int r0 = q & 3, r1 = q >> 2;
                    const float2 w = GFX_TW15 ? tw[q] : (r1 == 0 ? lo[r0] : (r0 == 0 ? hi[r1] : cmul(lo[r0], hi[r1])));
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
                            const float2 w = GFX_TW15 ? tw[r] : (jj == 0 ? lo[i] : (i == 0 ? hi[jj] : cmul(lo[i], hi[jj])));
                            v = tw_apply<false>(v, w);
                        }
                    }
                    if constexpr (TO_GLOBAL) dst.put(b + ST * r, v);
                    else zb[r * PST] = v;
                }
            }
        }
    }
}

// forward passes 0 .. last-1 (pass 0 reads global memory); ends with a barrier
template <int N, int NT, typename Src, int GUNR = 1>
__device__ __forceinline__ void fft_forward_front(pk2* z, const float2* __restrict__ plan, const Src& src) {
    fft_pass<N, NT, 0, false, true, false, Src, NoDst, GUNR>(z, plan, src);
    __syncthreads();
    fft_pass<N, NT, 1, false>(z, plan);
    __syncthreads();
    if constexpr (plan_stages(N) > 3) {
        fft_pass<N, NT, 2, false>(z, plan);
        __syncthreads();
    }
}
// inverse passes last-1 .. 0 (pass 0 writes global memory); starts with a barrier
template <int N, int NT, typename Dst, int GUNR = 1>
__device__ __forceinline__ void fft_inverse_back(pk2* z, const float2* __restrict__ plan, const Dst& dst) {
    __syncthreads();
    if constexpr (plan_stages(N) > 3) {
        fft_pass<N, NT, 2, true>(z, plan);
        __syncthreads();
    }
    fft_pass<N, NT, 1, true>(z, plan);
    __syncthreads();
    fft_pass<N, NT, 0, true, false, true, NoSrc, Dst, GUNR>(z, plan, NoSrc(), dst);
}

// ---- the fused last forward / first inverse pass
// forward: butterfly b from smem -> r16 -> registers (output index r in a[OUT16(r)])
__device__ __forceinline__ void last_forward(const pk2* z, int b, pk2 (&a)[16]) {
    const pk2* zb = z + 17 * b;
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = zb[q];
    r16<false>(a);
}
// inverse: registers (index r in a[OUT16(r)]) -> r16 inverse -> smem butterfly b
__device__ __forceinline__ void first_inverse(pk2* z, int b, const pk2 (&a)[16]) {
    pk2 in[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) in[q] = a[OUT16(q)];
    r16<true>(in);
    pk2* zb = z + 17 * b;
#pragma unroll
    for (int r = 0; r < 16; ++r) zb[r] = in[OUT16(r)];
}

// pair operators: what happens to (Z_k, Z_{N-k}) once both are in registers
struct OpSpectrum {  // untangle -> global pair slot (scaled)
    float4* out;
    float scale;
    typedef int Data;
    __device__ __forceinline__ Data fetch(int) const { return 0; }
    __device__ __forceinline__ void pair(pk2& zk, pk2& zm, float2 w, Data, int slot) const {
        float a, b, c, d;
        pk_split(zk, a, b);
        pk_split(zm, c, d);
        const PairA x = untangle_pair(make_float2(a, b), make_float2(c, d), w);
        out[slot] = make_float4(x.k.x * scale, x.k.y * scale, x.m.x * scale, x.m.y * scale);
    }
    __device__ __forceinline__ void dc(pk2& z0, pk2& zh, Data, int slot) const {
        float a, b, c, d;
        pk_split(z0, a, b);
        pk_split(zh, c, d);
        out[slot] = make_float4((a + b) * scale, (a - b) * scale, c * scale, -d * scale);
    }
};
struct OpFilter {  // untangle, multiply by the (1/N-scaled) filter spectrum, retangle -- in place
    const float4* H;
    typedef float4 Data;
    __device__ __forceinline__ Data fetch(int slot) const { return __ldg(H + slot); }
    __device__ __forceinline__ void pair(pk2& zk, pk2& zm, float2 w, Data h, int) const {
        float a, b, c, d;
        pk_split(zk, a, b);
        pk_split(zm, c, d);
        const PairA x = untangle_pair(make_float2(a, b), make_float2(c, d), w);
        const PairA y = retangle_pair(cmul(x.k, make_float2(h.x, h.y)), cmul(x.m, make_float2(h.z, h.w)), w);
        zk = pk_make(y.k.x, y.k.y);
        zm = pk_make(y.m.x, y.m.y);
    }
    __device__ __forceinline__ void dc(pk2& z0, pk2& zh, Data h, int) const {
        float a, b, c, d;
        pk_split(z0, a, b);
        pk_split(zh, c, d);
        const float y0 = (a + b) * h.x, yn = (a - b) * h.y;
        z0 = pk_make(0.5f * (y0 + yn), 0.5f * (y0 - yn));
        // X_{N/2} = conj(Z), Y = X H, Z' = conj(Y) = Z conj(H)
        const float2 r2 = cmulc(make_float2(c, d), make_float2(h.z, h.w));
        zh = pk_make(r2.x, r2.y);
    }
};
template <bool COHERENT = false>
struct OpLoad {  // global pair slot -> retangle -> registers (COHERENT: the slot was written earlier in this launch)
    const float4* Y;
    typedef float4 Data;
    __device__ __forceinline__ Data fetch(int slot) const { return COHERENT ? __ldcg(Y + slot) : ldg_stream(Y + slot); }
    __device__ __forceinline__ void pair(pk2& zk, pk2& zm, float2 w, Data y, int) const {
        const PairA a = retangle_pair(make_float2(y.x, y.y), make_float2(y.z, y.w), w);
        zk = pk_make(a.k.x, a.k.y);
        zm = pk_make(a.m.x, a.m.y);
    }
    __device__ __forceinline__ void dc(pk2& z0, pk2& zh, Data y, int) const {
        z0 = pk_make(0.5f * (y.x + y.y), 0.5f * (y.x - y.y));
        zh = pk_make(y.z, -y.w);
    }
};

// Thread 0 owns the two self-paired butterflies: bA (klow 0) pairs index rk with 16 - rk (rk = 0: DC with the
// Nyquist bin at index 8), bB (klow N/32) pairs rk with 15 - rk.  Swapping register halves between A and B puts
// them into the general cross pattern (A'[rk] <-> B'[15 - rk], B'[rk] <-> A'[15 - rk]):
//   A' = { A[0..7], B[8..15] },   B' = { B[0..7], A[9..15], A[8] }.
// The swap is its own inverse up to the rotation of B'[8..15]; both directions below (selects, no branches).
__device__ __forceinline__ void self_pair_swap(pk2 (&A)[16], pk2 (&B)[16], bool t0, bool back) {
    if (!back) {
        pk2 a8 = A[OUT16(8)];
#pragma unroll
        for (int j = 8; j < 15; ++j) {
            const pk2 an = A[OUT16(j + 1)], bj = B[OUT16(j)];
            B[OUT16(j)] = t0 ? an : bj;
            A[OUT16(j)] = t0 ? bj : A[OUT16(j)];
        }
        const pk2 b15 = B[OUT16(15)];
        B[OUT16(15)] = t0 ? a8 : b15;
        A[OUT16(15)] = t0 ? b15 : A[OUT16(15)];
    } else {
        // A[8] = B'[15], A[j+1] = B'[j] (j = 8..14), B[j] = A'[j] (j = 8..15)
        const pk2 na8 = B[OUT16(15)];
        pk2 nb[8], na[8];
#pragma unroll
        for (int j = 8; j < 16; ++j) nb[j - 8] = A[OUT16(j)];
#pragma unroll
        for (int j = 9; j < 16; ++j) na[j - 8] = B[OUT16(j - 1)];
        na[0] = na8;
#pragma unroll
        for (int j = 8; j < 16; ++j) {
            A[OUT16(j)] = t0 ? na[j - 8] : A[OUT16(j)];
            B[OUT16(j)] = t0 ? nb[j - 8] : B[OUT16(j)];
        }
    }
}

// all 16 pairs of a thread; A / B = the r16 output registers of its butterflies bA / bB
// IN: A, B hold forward outputs (false for the load-only operator); OUT: A, B are consumed afterwards
template <int N, bool IN, bool OUT, typename Op>
__device__ __forceinline__ void pair_phase(pk2 (&A)[16], pk2 (&B)[16], int t, const float2* __restrict__ plan, const Op& op) {
    constexpr int NTT = N / 32, Q = N / 4;
    const bool t0 = t == 0;
    if constexpr (IN) self_pair_swap(A, B, t0, false);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float2 w[4];
            typename Op::Data d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int slot = h * Q + (4 * g + u) * NTT + t;
                w[u] = __ldg(plan + slot);
                d[u] = op.fetch(slot);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rk = 4 * g + u;
                const int slot = h * Q + rk * NTT + t;
                // bin klow(b) + (N/16) rk of one butterfly pairs with index 15 - rk of the other
                if (h == 0 && rk == 0) {
                    if (t0) op.dc(A[OUT16(0)], B[OUT16(15)], d[u], slot);  // slot 0: DC / Nyquist and the bin N/2
                    else op.pair(A[OUT16(0)], B[OUT16(15)], w[u], d[u], slot);
                } else if (h == 0) {
                    op.pair(A[OUT16(rk)], B[OUT16(15 - rk)], w[u], d[u], slot);
                } else {
                    op.pair(B[OUT16(rk)], A[OUT16(15 - rk)], w[u], d[u], slot);
                }
            }
        }
    }
    if constexpr (OUT) self_pair_swap(A, B, t0, true);
}

struct RowMap {  // output row -> (x row, h row) with channel broadcasting; h_rep consecutive batch items share a filter
    int c_out, cx, ch, h_rep;
    __device__ __forceinline__ void map(int r, int& xr, int& hr) const {
        const int b = r / c_out, c = r - b * c_out;
        xr = b * cx + (cx == 1 ? 0 : c);
        hr = (b / h_rep) * ch + (ch == 1 ? 0 : c);
    }
};

// filter rows, optionally with the unit-energy normalisation of a reverb IR folded into the spectra
// (reference: normalize_impulse after the optional ms_to_lr, reverb.py:215-228, core/utils.py:14-18)
struct FilterSrc {
    const float* h;       // [hrows, Nh]
    const float* energy;  // [batch, 2] sum_t of the squared raw mid / side rows, or null (plain filter)
    int to_lr;            // with energy: rows are left/right = mid +- side (mean-over-channels energy = e0 + e1)
    int act;              // 1: FIRFilter (filter.py:65-77): taps = tanh(h), energy [batch] = mean_c sum_t tanh(h)^2; ch_act rows per item
    int ch_act;
};
// tanh on the fly (tanh(0) = 0: the zero padding of a segment is unaffected)
template <typename Src>
struct TanhSrc {
    Src s;
    __device__ __forceinline__ pk2 get(int c) const {
        float a, b;
        pk_split(s.get(c), a, b);
        return pk_make(tanhf(a), tanhf(b));
    }
};
// scale of a filter row: 1/N, times the unit-energy normalisation where the source asks for it
__device__ __forceinline__ float filter_scale(const FilterSrc& fs, int hrow, int N) {
    float scale = 1.f / (float)N;
    if (fs.act) {
        scale *= rsqrtf(fs.energy[hrow / fs.ch_act] + 1e-12f);
    } else if (fs.energy) {
        // normalize_impulse of a reverb IR: one scale per batch item from the raw mid/side energies
        const float e0 = fs.energy[hrow & ~1], e1 = fs.energy[hrow | 1];
        scale *= fs.to_lr ? rsqrtf(e0 + e1 + 1e-12f) : rsqrtf(0.5f * (e0 + e1) + 1e-12f);
    }
    return scale;
}
// energy [batch] = mean over the ch rows of sum_t tanh(h)^2 (normalize_impulse of the activated taps, core/utils.py:14-18)
__global__ void __launch_bounds__(256) fir_tanh_energy_kernel(const float* __restrict__ h, float* __restrict__ energy, int ch, int Nh) {
    __shared__ float red[8];
    const float* row = h + (size_t)blockIdx.x * ch * Nh;
    float acc = 0.f;
    for (int i = threadIdx.x; i < ch * Nh; i += 256) {
        const float t = tanhf(row[i]);
        acc = fmaf(t, t, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        energy[blockIdx.x] = t / (float)ch;
    }
}

// NT = N/32 threads; registers: two butterflies (64) + pair operands live in the fused phase
__host__ __device__ constexpr int fir_nt(int n) { return n / 32; }
#ifndef GFX_FIR_MB8192
#define GFX_FIR_MB8192 2
#endif
__host__ __device__ constexpr int fir_min_blocks(int n) { return n == 1024 ? 16 : (n == 4096 ? 4 : (n == 8192 ? GFX_FIR_MB8192 : 1)); }
// the one-transform kernels of the long-filter path (x-block spectra, inverse) at 4096 points: CTAs per SM
#ifndef GFX_XSPEC_MB4096
#define GFX_XSPEC_MB4096 4
#endif
#ifndef GFX_INV_MB4096
#define GFX_INV_MB4096 4
#endif
__host__ __device__ constexpr int xspec_min_blocks(int n) { return n == 4096 ? GFX_XSPEC_MB4096 : fir_min_blocks(n); }
__host__ __device__ constexpr int inv_min_blocks(int n) { return n == 4096 ? GFX_INV_MB4096 : fir_min_blocks(n); }

// ------------------------------------------------------------------ kernels
// forward transform of one real segment into pair slots: shared tail of the spectrum kernels
template <int N, typename Src>
__device__ __forceinline__ void segment_spectrum(pk2* zbuf, const float2* __restrict__ plan, const Src& src,
                                                 float4* __restrict__ out, float scale) {
    constexpr int NT = fir_nt(N);
    fft_forward_front<N, NT, Src, 2>(zbuf, plan, src);
    const int t = threadIdx.x;
    const ushort2 pr = plan_pairtab<N>(plan)[t];
    pk2 A[16], B[16];
    last_forward(zbuf, pr.x, A);
    last_forward(zbuf, pr.y, B);
    pair_phase<N, true, false>(A, B, t, plan, OpSpectrum{out, scale});
}

// spectra of filter partitions: Hs[(hrow * P + part) * N/2 + slot], scaled by 1/N (and the energy norm)
template <int N, bool FAST>
__global__ void __launch_bounds__(fir_nt(N), fir_min_blocks(N)) fir_spectrum_kernel(FilterSrc fs, float4* __restrict__ Hs,
                                                                                    int hrow0, int Nh, int part_len, int P,
                                                                                    const float2* __restrict__ plan) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int hloc = blockIdx.x / P, part = blockIdx.x - hloc * P;
    const int hrow = hrow0 + hloc;
    const long long s0 = (long long)part * part_len;
    long long end = s0 + part_len;
    if (end > Nh) end = Nh;
    const float scale = filter_scale(fs, hrow, N);
    const SegSrc<FAST> src(fs.h + (size_t)hrow * Nh, s0, end, 2 * N);
    if (fs.act) segment_spectrum<N>(zbuf, plan, TanhSrc<SegSrc<FAST>>{src}, Hs + (size_t)blockIdx.x * (N / 2), scale);
    else segment_spectrum<N>(zbuf, plan, src, Hs + (size_t)blockIdx.x * (N / 2), scale);
}

// UPOLS step 1: spectra of input blocks.  Xs[(rloc * nblk + j) * N/2 + slot] = rfft of x[(j-1)B, (j+1)B), B = N
template <int N, bool FAST>
__global__ void __launch_bounds__(fir_nt(N), xspec_min_blocks(N)) fir_xspec_kernel(const float* __restrict__ x,
                                                                                 float4* __restrict__ Xs, int xrow0,
                                                                                 long long L, int nblk,
                                                                                 const float2* __restrict__ plan) {
    extern __shared__ __align__(16) pk2 zbuf[];
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    const size_t xr = (size_t)(xrow0 + rloc);
    const SegSrc<FAST> src(x + xr * L, ((long long)j - 1) * N, L, 2 * N);
    segment_spectrum<N>(zbuf, plan, src, Xs + (size_t)blockIdx.x * (N / 2), 1.f);
}

// single-partition overlap-save: block j produces full-convolution samples [j*hop, (j+1)*hop)
template <int N, bool FAST>
__global__ void __launch_bounds__(fir_nt(N), fir_min_blocks(N)) fir_ols_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                               const float4* __restrict__ Hs, RowMap rm,
                                                                               long long L, int pre, int hop, int shift,
                                                                               int nblk, const float2* __restrict__ plan) {
    constexpr int NT = fir_nt(N);
    extern __shared__ __align__(16) pk2 zbuf[];
    const int row = blockIdx.x / nblk, j = blockIdx.x - row * nblk;
    int xr, hr;
    rm.map(row, xr, hr);
    const long long m0 = (long long)j * hop;
    const SegSrc<FAST> src(x + (size_t)xr * L, m0 - pre, L, 2 * N);
    fft_forward_front<N, NT>(zbuf, plan, src);
    const int t = threadIdx.x;
    const ushort2 pr = plan_pairtab<N>(plan)[t];
    {
        pk2 A[16], B[16];
        last_forward(zbuf, pr.x, A);
        last_forward(zbuf, pr.y, B);
        pair_phase<N, true, true>(A, B, t, plan, OpFilter{Hs + (size_t)hr * (N / 2)});
        first_inverse(zbuf, pr.x, A);  // (each thread rewrites exactly the positions it read: no barrier in between)
        first_inverse(zbuf, pr.y, B);
    }
    const SegDst<FAST> dst(y + (size_t)row * L, m0 - shift, L, pre, pre + hop);
    fft_inverse_back<N, NT>(zbuf, plan, dst);
}

// UPOLS step 2: Y_j = sum_p X_{j-p} H_p on pair slots, partitions [p0, p0 + PC).  One thread owns one pair slot of
// one output row and walks the blocks in order: the PC filter values and the last PC input values stay in
// registers, so every spectrum is read exactly once.  accumulate != 0 adds to Ys (filters with more than
// MAC_MAX_PC partitions are processed in groups).
constexpr int MAC_MAX_PC = 12;
// CTA shape of the MAC kernel: (threads, min CTAs per SM, loads issued together).  The kernel is HBM-bound, so what
// matters is bytes in flight per SM and no spills.  Measured on B200, config 3 step: (256, 2, 4) 2.446 ms (128
// registers, 104 B of spills) | (256, 1, 12) 2.603 | (128, 3, 6) 2.450 | (128, 2, 12) 2.550 | (64, 5, 12) 2.386 ms (160
// registers, no spills, 61 KB of loads in flight per SM).
#ifndef GFX_MAC_NT
#define GFX_MAC_NT 64
#endif
#ifndef GFX_MAC_MINB
#define GFX_MAC_MINB 5
#endif
#ifndef GFX_MAC_G
#define GFX_MAC_G 12
#endif
constexpr int MAC_NT = GFX_MAC_NT;

__device__ __forceinline__ void cmac4(float4& acc, const float4& x, const float4& h) {
    acc.x = fmaf(x.x, h.x, acc.x); acc.x = fmaf(-x.y, h.y, acc.x);
    acc.y = fmaf(x.x, h.y, acc.y); acc.y = fmaf(x.y, h.x, acc.y);
    acc.z = fmaf(x.z, h.z, acc.z); acc.z = fmaf(-x.w, h.w, acc.z);
    acc.w = fmaf(x.z, h.w, acc.w); acc.w = fmaf(x.w, h.z, acc.w);
}

template <int PC>
__global__ void __launch_bounds__(MAC_NT, GFX_MAC_MINB) fir_mac_kernel(const float4* __restrict__ Xs, const float4* __restrict__ Hs,
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
    constexpr int G = PC < GFX_MAC_G ? PC : ((PC % GFX_MAC_G == 0) ? GFX_MAC_G : (PC < 4 ? PC : 4));  // loads issued together
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

// UPOLS step 2, streaming form.  fir_mac_kernel above is latency-bound (profiles/r01_final_fir_mac_12.txt: 2.8 warps per
// scheduler, 40 % of the stalls wait for the 12 loads a thread issues and then consumes): with the 4096-point partitioning
// (24 partitions of a 96000-tap response; FFT kernels at 4 CTAs per SM) the arithmetic per byte doubles and a load-then-
// compute thread takes twice as long.  Here (a) one thread owns ONE complex bin (half a pair slot, 8 bytes), which halves
// the registers per partition: up to MAC2_MAX_PC = 24 partitions (filter 48 + input ring 48 registers) in one pass over
// the spectra; (b) the input blocks arrive through a cp.async ring in shared memory, MAC2_STAGES blocks ahead of the
// arithmetic (every thread copies and later reads its own 8 bytes: no barrier), so no register holds a load in flight;
// (c) the loop body covers U = 4 output blocks with compile-time operand indices and then moves the ring (small code);
// the bodies in which the ring is still filling up are peeled and skip the empty entries.
// A CTA can also walk `bpc` consecutive batch items that share the filter (render_grafx's 4-D sources) with the filter partitions loaded once.
// Element e = 2 slot + (0: bin k | 1: bin n-k); element 0 is (A_0, A_N), two real bins.
constexpr int MAC2_MAX_PC = 24;
#ifndef GFX_MAC2_NT
#define GFX_MAC2_NT 64
#endif
#ifndef GFX_MAC2_MINB
#define GFX_MAC2_MINB 8
#endif
#ifndef GFX_MAC2_STAGES
#define GFX_MAC2_STAGES 8
#endif
constexpr int MAC2_NT = GFX_MAC2_NT, MAC2_D = GFX_MAC2_STAGES;

__device__ __forceinline__ void cmac2(float2& acc, const float2& x, const float2& h) {
    acc.x = fmaf(x.x, h.x, acc.x); acc.x = fmaf(-x.y, h.y, acc.x);
    acc.y = fmaf(x.x, h.y, acc.y); acc.y = fmaf(x.y, h.x, acc.y);
}
__device__ __forceinline__ void stg_stream2(float2* p, const float2& v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y));
}

// state of one row walk.  The copy ring is two sets of U stages: while a body works on one set, the U blocks of the
// next body land in the other (offsets inside a set are compile-time constants; the sets swap by two pointer moves)
struct Mac2Walk {
    const float2* x;     // input block 0 of the row (this thread's bin)
    float2* yp;          // output block written next
    float2* s_read;      // stage set the body reads
    float2* s_write;     // stage set the copies of the next body land in
    int next;            // index of the first input block of the NEXT body
    int last;            // last input block index (copies past the end re-read it; their products are never stored)
    int to_do;           // output blocks left
};

// U output blocks per loop body.  ring[k] = input block (newest - k) when the body starts; inside the body every operand
// index is a compile-time constant; afterwards the ring moves U places (PC - U register moves per U x PC products).
// The body is U x PC complex products = 4 U PC FFMA: 6 KB of code at U = 4, PC = 24 -- the fully unrolled PC x PC tile
// (93 KB) ran out of the 32 KB instruction cache (profiles/r02_mac2_unrolled.txt: 45 % of the stalls no_inst).
// VALID = ring entries that hold data (the ring fills up over the first PC blocks of a row: products with the
// still-empty entries are left out of the peeled bodies -- 36 % of the multiplies at 24 partitions x 32 blocks).
// GUARD: the body may run past the end of the row (peeled bodies and the tail); the steady loop has no per-block tests.
template <int PC, int U, int VALID, bool GUARD>
__device__ __forceinline__ void mac2_body(Mac2Walk& w, float2 (&h)[PC], float2 (&ring)[PC], size_t half2, int accumulate) {
    // the U blocks of the next body enter the other stage set ...
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int jb = min(w.next + u, w.last);
        cp_async_small<8>(w.s_write + u * MAC2_NT, w.x + (size_t)jb * half2);
    }
    cp_async_commit();
    w.next += U;
    // ... while this body's blocks (requested one body ago) are complete
    cp_async_wait<1>();
    float2 xn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xn[u] = w.s_read[u * MAC2_NT];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float2 acc = make_float2(0.f, 0.f);
        if (accumulate && (!GUARD || u < w.to_do)) acc = w.yp[(size_t)u * half2];
#pragma unroll
        for (int p = 0; p < PC; ++p) {
            if (p > u && p - u - 1 >= VALID) continue;
            cmac2(acc, p <= u ? xn[u - p] : ring[p - u - 1], h[p]);
        }
        if (!GUARD || u < w.to_do) stg_stream2(w.yp + (size_t)u * half2, acc);
    }
#pragma unroll
    for (int k = PC - 1; k >= U; --k) ring[k] = ring[k - U];
#pragma unroll
    for (int k = 0; k < U && k < PC; ++k) ring[k] = xn[U - 1 - k];
    w.yp += (size_t)U * half2;
    w.to_do -= U;
    float2* t = w.s_read; w.s_read = w.s_write; w.s_write = t;
}
// the bodies of a row: ring filling up (VALID = 0, U, 2U, ...), then the steady loop, then a guarded tail
template <int PC, int U, int VALID>
__device__ __forceinline__ void mac2_row(Mac2Walk& w, float2 (&h)[PC], float2 (&ring)[PC], size_t half2, int accumulate) {
    if constexpr (VALID >= PC) {
#pragma unroll 1
        while (w.to_do >= U) mac2_body<PC, U, PC, false>(w, h, ring, half2, accumulate);
        if (w.to_do > 0) mac2_body<PC, U, PC, true>(w, h, ring, half2, accumulate);
    } else {
        if (w.to_do > 0) {
            mac2_body<PC, U, VALID, true>(w, h, ring, half2, accumulate);
            mac2_row<PC, U, (VALID + U < PC ? VALID + U : PC)>(w, h, ring, half2, accumulate);
        }
    }
}

// grid: (2 * half / MAC2_NT, groups * c_out); group g covers batch items [b0 + g bpc, b0 + (g + 1) bpc) of the sweep, all
// on the filter of its first item (the host guarantees bpc divides h_rep and b0 is a multiple of it)
template <int PC>
__global__ void __launch_bounds__(MAC2_NT, GFX_MAC2_MINB) fir_mac2_kernel(const float2* __restrict__ Xs, const float2* __restrict__ Hs,
                                                                          float2* __restrict__ Ys, RowMap rm, int xrow0, int hrow0,
                                                                          int b0, int nb, int bpc, int P, int p0, int nblk,
                                                                          int half2, int accumulate) {
    __shared__ float2 stage_all[MAC2_D * MAC2_NT];
    float2* stage = stage_all + threadIdx.x;
    const int e = blockIdx.x * MAC2_NT + threadIdx.x;
    const int g = blockIdx.y / rm.c_out, c = blockIdx.y - g * rm.c_out;
    const int bfirst = b0 + g * bpc;
    int xr, hr;
    rm.map(bfirst * rm.c_out + c, xr, hr);
    const float2* H = Hs + ((size_t)(hr - hrow0) * P + p0) * half2 + e;
    float2 h[PC], ring[PC];
#pragma unroll
    for (int p = 0; p < PC; ++p) h[p] = (p0 + p < P) ? __ldg(H + (size_t)p * half2) : make_float2(0.f, 0.f);
    if (e == 0) {
        // (A_0, A_N): two REAL bins, not a complex product -- lanes of warp 0 add them below
#pragma unroll
        for (int p = 0; p < PC; ++p) h[p] = make_float2(0.f, 0.f);
    }
    int blast = bfirst + bpc;
    if (blast > b0 + nb) blast = b0 + nb;
#pragma unroll 1
    for (int b = bfirst; b < blast; ++b) {
        rm.map(b * rm.c_out + c, xr, hr);
        const float2* X = Xs + (size_t)(xr - xrow0) * nblk * half2 + e;
        float2* Y = Ys + (size_t)((b - b0) * rm.c_out + c) * nblk * half2 + e;
#pragma unroll
        for (int p = 0; p < PC; ++p) ring[p] = make_float2(0.f, 0.f);
        // prologue of the copy ring: the U blocks of the first body
#ifdef GFX_MAC2_U
        constexpr int U = GFX_MAC2_U;
#else
        constexpr int U = 4;  // (measured on B200 at 24 partitions: U = 4 1.655 ms per reverb-shape convolution, U = 8 1.737 ms)
#endif
        static_assert(2 * U <= MAC2_D, "two stage sets of U blocks");
        Mac2Walk w;
        w.x = X; w.yp = Y + (size_t)p0 * half2; w.s_read = stage; w.s_write = stage + U * MAC2_NT;
        w.last = nblk - 1 - p0; w.to_do = nblk - p0; w.next = U;
#pragma unroll
        for (int u = 0; u < U; ++u) cp_async_small<8>(w.s_read + u * MAC2_NT, X + (size_t)min(u, w.last) * half2);
        cp_async_commit();
        mac2_row<PC, U, 0>(w, h, ring, (size_t)half2, accumulate);
        cp_async_wait<0>();
        if (blockIdx.x == 0 && threadIdx.x < 32) {
            // DC / Nyquist of every block (element 0): lanes over output blocks
            __syncwarp();
            const float2* X0 = X - e;
            const float2* H0 = H - e;
            float2* Y0 = Y - e;
            for (int j = p0 + (int)threadIdx.x; j < nblk; j += 32) {
                float a0 = 0.f, an = 0.f;
                for (int p = 0; p < PC && p0 + p <= j && p0 + p < P; ++p) {
                    const float2 xv = __ldg(X0 + (size_t)(j - p0 - p) * half2);
                    const float2 hv = __ldg(H0 + (size_t)p * half2);
                    a0 = fmaf(xv.x, hv.x, a0);
                    an = fmaf(xv.y, hv.y, an);
                }
                // lane 0 (e == 0) wrote element 0 of block j in the main loop; the __syncwarp above orders that store
                float2 v = Y0[(size_t)j * half2];
                v.x += a0;
                v.y += an;
                Y0[(size_t)j * half2] = v;
            }
            __syncwarp();
        }
    }
}

// UPOLS step 2, packed form (the default where it applies: all partitions in one pass, P <= 24, and rows of at most
// MAC3_MAXBLK = 32 blocks -- 131072 samples at 4096-tap partitions, every BASELINE shape).  fir_mac2_kernel above issues
// 115 instructions per (bin, block) of which 64.5 are scalar FFMA (profiles/r02_final_mac2_24.txt: issue 66 %, FMA pipe
// 42 %): it is bound by instruction issue, not by the pipe.  Here
//   (a) the complex product is two PACKED accumulations with the input as loaded, x = (re, im):
//         s1 += x * (h.re, h.re) = (x.re h.re, x.im h.re),   s2 += x * (h.im, h.im) = (x.re h.im, x.im h.im),
//         y = (s1.lo - s2.hi, s1.hi + s2.lo)   -- 2 FFMA2 per complex product instead of 4 FFMA, no swaps or negations;
//       the filter sits in registers as duplicated pairs (4 registers per partition);
//   (b) the whole input history of the row (<= 32 blocks x 8 bytes per thread, 16 KB per CTA) is copied to shared memory by
//       cp.async up front, in commit groups of U = 4 blocks, and read back with compile-time offsets: no register ring,
//       no ring moves, no stage juggling; body k waits for group k only (cp.async.wait_group 7 - k);
//   (c) the (at most 8) bodies of U output blocks are unrolled with compile-time block indices, the products with blocks
//       before the start of the row simply do not exist (the peeled ramp-up of fir_mac2_kernel, for free);
//   (d) element 0 -- (A_0, A_N), two REAL bins whose products are component-wise -- takes h = ((h.x, h.y), 0): the same
//       arithmetic then yields (x.x h.x, x.y h.y), so there is no separate DC / Nyquist pass.
constexpr int MAC3_NT = 64, MAC3_MAXBLK = 32, MAC3_U = 4;
#ifndef GFX_MAC3_MINB
#define GFX_MAC3_MINB 8  // (measured on B200, 96000 taps x 512 stereo rows: 6 CTAs per SM 1.486 ms, 7 or 8 (128 registers, no spills) 1.387 ms)
#endif

// FULL: the row has exactly MAC3_MAXBLK blocks (no per-block tests: 131072 samples at 4096-tap partitions)
template <int PC, int K, bool FULL>
__device__ __forceinline__ void mac3_body(const pk2* __restrict__ xs, const pk2 (&hrr)[PC], const pk2 (&hii)[PC],
                                          float2*& yp, unsigned long long bstride, int nblk) {
    constexpr int U = MAC3_U, J0 = K * U;
    if (!FULL && J0 >= nblk) return;             // (uniform over the CTA)
    cp_async_wait<MAC3_MAXBLK / U - 1 - K>();     // groups 0 .. K have landed (each thread reads its own copies only)
    pk2 s1[U], s2[U];
    bool started[U];
#pragma unroll
    for (int u = 0; u < U; ++u) started[u] = false;
#pragma unroll
    for (int d = U - 1; d >= -(PC - 1); --d) {   // input block J0 + d feeds output block J0 + u through partition u - d
        if (J0 + d < 0) continue;
        const pk2 x = xs[(J0 + d) * MAC3_NT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = u - d;
            if (p < 0 || p >= PC) continue;
            if (!started[u]) {                    // (resolved at compile time: the first product initialises the sums)
                s1[u] = pk_mul(x, hrr[p]);
                s2[u] = pk_mul(x, hii[p]);
                started[u] = true;
            } else {
                pk_fma_acc(s1[u], x, hrr[p]);
                pk_fma_acc(s2[u], x, hii[p]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (FULL || J0 + u < nblk) {
            float a, b, c, d;
            pk_split(s1[u], a, b);
            pk_split(s2[u], c, d);
            stg_stream2(yp, make_float2(a - d, b + c));
        }
        yp = reinterpret_cast<float2*>(reinterpret_cast<unsigned long long>(yp) + bstride);
    }
}

// a pointer ptxas must treat as an opaque 64-bit value: bumping it is one 64-bit add (2 instructions) instead of the
// index + scale + base form (4 instructions) it otherwise rebuilds for every block
template <typename T>
__device__ __forceinline__ T* opaque_ptr(T* p) {
    asm volatile("" : "+l"(p));
    return p;
}
template <typename T>
__device__ __forceinline__ T* bump_bytes(T* p, unsigned long long bytes) {
    return reinterpret_cast<T*>(reinterpret_cast<unsigned long long>(p) + bytes);
}

// grid: (2 * half / MAC3_NT, c_out, nb): one CTA = 64 consecutive elements of one output row, all its blocks
template <int PC, bool FULL>
__global__ void __launch_bounds__(MAC3_NT, GFX_MAC3_MINB) fir_mac3_kernel(const float2* __restrict__ Xs, const float2* __restrict__ Hs,
                                                                          float2* __restrict__ Ys, RowMap rm, int xrow0, int hrow0,
                                                                          int b0, int P, int nblk, int half2) {
    __shared__ pk2 xsm[MAC3_MAXBLK * MAC3_NT];
    pk2* xs = xsm + threadIdx.x;
    const int e = blockIdx.x * MAC3_NT + threadIdx.x;
    const int c = blockIdx.y, bl = blockIdx.z, b = b0 + bl;
    const int xr = b * rm.cx + (rm.cx == 1 ? 0 : c);
    const int hr = (rm.h_rep == 1 ? b : b / rm.h_rep) * rm.ch + (rm.ch == 1 ? 0 : c);
    const unsigned long long bstride = (unsigned long long)half2 * sizeof(float2);
    const float2* xp = opaque_ptr(Xs + (size_t)(xr - xrow0) * nblk * half2 + e);
    const float2* hp = opaque_ptr(Hs + (size_t)(hr - hrow0) * P * half2 + e);
    float2* yp = opaque_ptr(Ys + (size_t)(bl * rm.c_out + c) * nblk * half2 + e);
#pragma unroll
    for (int g = 0; g < MAC3_MAXBLK / MAC3_U; ++g) {
#pragma unroll
        for (int u = 0; u < MAC3_U; ++u) {
            const int j = g * MAC3_U + u;
            if (FULL || j < nblk) cp_async_small<8>(xs + j * MAC3_NT, xp);
            xp = bump_bytes(xp, bstride);
        }
        cp_async_commit();
    }
    pk2 hrr[PC], hii[PC];
    const bool e0 = e == 0;
#pragma unroll
    for (int p = 0; p < PC; ++p) {
        const float2 h = p < P ? __ldg(hp) : make_float2(0.f, 0.f);
        hp = bump_bytes(hp, bstride);
        hrr[p] = pk_make(h.x, e0 ? h.y : h.x);
        hii[p] = pk_dup(e0 ? 0.f : h.y);
    }
    mac3_body<PC, 0, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 1, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 2, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 3, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 4, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 5, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 6, FULL>(xs, hrr, hii, yp, bstride, nblk);
    mac3_body<PC, 7, FULL>(xs, hrr, hii, yp, bstride, nblk);
    cp_async_wait<0>();
}

// UPOLS step 3: inverse FFT of Y_j, keep the second half of the block
template <int N, bool FAST>
__global__ void __launch_bounds__(fir_nt(N), inv_min_blocks(N)) fir_inv_kernel(const float4* __restrict__ Ys, float* __restrict__ y,
                                                                               int row0, long long L, int nblk, int shift,
                                                                               const float2* __restrict__ plan) {
    constexpr int NT = fir_nt(N);
    extern __shared__ __align__(16) pk2 zbuf[];
    const int rloc = blockIdx.x / nblk, j = blockIdx.x - rloc * nblk;
    const int t = threadIdx.x;
    const ushort2 pr = plan_pairtab<N>(plan)[t];
    {
        pk2 A[16], B[16];
        pair_phase<N, false, true>(A, B, t, plan, OpLoad<false>{Ys + (size_t)blockIdx.x * (N / 2)});
        first_inverse(zbuf, pr.x, A);
        first_inverse(zbuf, pr.y, B);
    }
    const size_t row = (size_t)(row0 + rloc);
    const SegDst<FAST> dst(y + row * L, (long long)j * N - shift, L, N, 2 * N);
    fft_inverse_back<N, NT, SegDst<FAST>, 2>(zbuf, plan, dst);
}

// ------------------------------------------------------------------ long filters: one persistent pipelined launch
// The four steps above as WORK ITEMS of one launch (one CTA = one item at a time, atomic ticket), so that the
// spectra never leave L2: item kinds  H (filter-partition spectrum), X (input-block spectrum), M (per-slot
// multiply-accumulate of one output row over one sixteenth of the slots), O (inverse transform of one output block).
// Tickets are laid out as a software pipeline over the batch: step s issues  H/X of item s+2D,  M of item s+D,
// O of item s;  an item that needs data waits on per-item completion counters, and since everything it can wait
// for has a smaller ticket (already running or done) the launch cannot deadlock.  Spectra live in a ring of R item
// slots (R >= 2D + 3, ~5.5 MB per item at the BASELINE reverb shape): the live set is ~50 MB and stays in the 126 MB L2.
struct UpolsJob {
    const float* x; FilterSrc fs; float* y;
    float4* Hs; float4* Xs; float4* Ys;   // rings [R][ch*P][half], [R][cx*nblk][half], [R][c_out*nblk][half]
    unsigned int* ticket;
    int* cntH; int* cntX; int* cntM; int* cntMH; int* cntO;  // [batch] completion counters (H, MH by filter item)
    int batch, cx, ch, c_out, h_rep;
    long long L;
    int Nh, P, nblk, shift, R, D;
    const float2* plan;
};

__device__ __forceinline__ void item_wait(const int* cnt, int target) {
    if (threadIdx.x == 0) {
        const volatile int* v = cnt;
        while (*v < target) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void item_done(int* cnt) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cnt, 1);
    }
}

// M item: output row (b, c), slots [slab NT, (slab+1) NT): all partitions (P <= MAC_MAX_PC, the rest of the register
// ring holds zeros).  Same loop as fir_mac_kernel, on L2-coherent loads.
template <int NT>
__device__ __forceinline__ void mac_item(const float4* __restrict__ Xrow, const float4* __restrict__ Hrow,
                                         float4* __restrict__ Yrow, int P, int nblk, int half, int slab) {
    constexpr int PC = MAC_MAX_PC;
    const int q = slab * NT + threadIdx.x;
    const float4* X = Xrow + q;
    const float4* H = Hrow + q;
    float4* Y = Yrow + q;
    float4 h[PC], ring[PC];
#pragma unroll
    for (int p = 0; p < PC; ++p) {
        h[p] = p < P ? __ldcg(H + (size_t)p * half) : make_float4(0.f, 0.f, 0.f, 0.f);
        ring[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (q == 0) {
#pragma unroll
        for (int p = 0; p < PC; ++p) { h[p].x = 0.f; h[p].y = 0.f; }
    }
    constexpr int G = 4;
    const int last = nblk - 1;
#pragma unroll 1
    for (int j0 = 0; j0 < nblk; j0 += PC) {
#pragma unroll
        for (int g0 = 0; g0 < PC; g0 += G) {
            float4 xn[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                int jb = j0 + g0 + g;
                jb = jb < last ? jb : last;
                xn[g] = __ldcg(X + (size_t)jb * half);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int jj = g0 + g;
                const int j = j0 + jj;
                ring[jj] = xn[g];
                if (j < nblk) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < PC; ++p) cmac4(acc, ring[(jj - p + PC) % PC], h[p]);
                    Y[(size_t)j * half] = acc;
                }
            }
        }
    }
    if (slab == 0 && threadIdx.x < 32) {
        // DC / Nyquist of every block (the .xy lanes of slot 0): lanes over output blocks
        __syncwarp();
        for (int j = (int)threadIdx.x; j < nblk; j += 32) {
            float a0 = 0.f, an = 0.f;
            for (int p = 0; p < P && p <= j; ++p) {
                const float4 xv = __ldcg(Xrow + (size_t)(j - p) * half);
                const float4 hv = __ldcg(Hrow + (size_t)p * half);
                a0 = fmaf(xv.x, hv.x, a0);
                an = fmaf(xv.y, hv.y, an);
            }
            // slot 0 of block j was stored by lane 0 in the loop above (ordered by the __syncwarp); go through L2
            float4 v = __ldcg(Yrow + (size_t)j * half);
            v.x += a0;
            v.y += an;
            __stcg(Yrow + (size_t)j * half, v);
        }
    }
}

template <int N, bool FAST>
__global__ void __launch_bounds__(fir_nt(N), fir_min_blocks(N)) fir_upols_pipeline_kernel(const UpolsJob jb) {
    constexpr int NT = fir_nt(N), half = N / 2, NSLAB = half / NT;
    extern __shared__ __align__(16) pk2 zbuf[];
    __shared__ unsigned int sh_ticket;
    const int nH = jb.ch * jb.P, nX = jb.cx * jb.nblk, nM = jb.c_out * NSLAB, nO = jb.c_out * jb.nblk;
    const int per_step = nH + nX + nM + nO;
    const long long total = (long long)(jb.batch + 2 * jb.D) * per_step;
    const int t = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (t == 0) sh_ticket = atomicAdd(jb.ticket, 1u);
        __syncthreads();
        const unsigned int tk = sh_ticket;
        if ((long long)tk >= total) break;
        const int step = (int)(tk / (unsigned)per_step) - 2 * jb.D;
        int u = (int)(tk % (unsigned)per_step);
        if (u < nH) {
            // ---- H item: partition `part` of filter row (hb, c)
            const int b = step + 2 * jb.D;
            if (b >= jb.batch || (b % jb.h_rep) != 0) continue;
            const int hb = b / jb.h_rep;
            if (hb >= jb.R) item_wait(jb.cntMH + (hb - jb.R), jb.h_rep * nM);  // ring slot free (its readers are done)
            const int c = u / jb.P, part = u - c * jb.P;
            const int hrow = hb * jb.ch + c;
            const long long s0 = (long long)part * N;
            long long end = s0 + N;
            if (end > jb.Nh) end = jb.Nh;
            const float scale = filter_scale(jb.fs, hrow, N);
            // (the filter rows are read-only inputs: the generic, alignment-agnostic source is used when FAST is off)
            const SegSrc<FAST> src(jb.fs.h + (size_t)hrow * jb.Nh, s0, end, 2 * N);
            segment_spectrum<N>(zbuf, jb.plan, src, jb.Hs + ((size_t)(hb % jb.R) * nH + u) * half, scale);
            item_done(jb.cntH + hb);
        } else if (u < nH + nX) {
            // ---- X item: block j of input row (b, c)
            u -= nH;
            const int b = step + 2 * jb.D;
            if (b >= jb.batch) continue;
            if (b >= jb.R) item_wait(jb.cntM + (b - jb.R), nM);
            const int c = u / jb.nblk, j = u - c * jb.nblk;
            const SegSrc<FAST> src(jb.x + (size_t)(b * jb.cx + c) * jb.L, ((long long)j - 1) * N, jb.L, 2 * N);
            segment_spectrum<N>(zbuf, jb.plan, src, jb.Xs + ((size_t)(b % jb.R) * nX + u) * half, 1.f);
            item_done(jb.cntX + b);
        } else if (u < nH + nX + nM) {
            // ---- M item
            u -= nH + nX;
            const int b = step + jb.D;
            if (b < 0 || b >= jb.batch) continue;
            const int hb = b / jb.h_rep;
            item_wait(jb.cntH + hb, nH);
            item_wait(jb.cntX + b, nX);
            if (b >= jb.R) item_wait(jb.cntO + (b - jb.R), nO);
            const int c = u / NSLAB, slab = u - c * NSLAB;
            const float4* Xrow = jb.Xs + ((size_t)(b % jb.R) * nX + (size_t)(jb.cx == 1 ? 0 : c) * jb.nblk) * half;
            const float4* Hrow = jb.Hs + ((size_t)(hb % jb.R) * nH + (size_t)(jb.ch == 1 ? 0 : c) * jb.P) * half;
            float4* Yrow = jb.Ys + ((size_t)(b % jb.R) * nO + (size_t)c * jb.nblk) * half;
            mac_item<NT>(Xrow, Hrow, Yrow, jb.P, jb.nblk, half, slab);
            __syncthreads();
            if (t == 0) {
                __threadfence();
                atomicAdd(jb.cntM + b, 1);
                atomicAdd(jb.cntMH + hb, 1);
            }
        } else {
            // ---- O item: output block j of row (b, c)
            u -= nH + nX + nM;
            const int b = step;
            if (b < 0) continue;
            item_wait(jb.cntM + b, nM);
            const int c = u / jb.nblk, j = u - c * jb.nblk;
            const ushort2 pr = plan_pairtab<N>(jb.plan)[t];
            {
                pk2 A[16], B[16];
                pair_phase<N, false, true>(A, B, t, jb.plan, OpLoad<true>{jb.Ys + ((size_t)(b % jb.R) * nO + u) * half});
                first_inverse(zbuf, pr.x, A);
                first_inverse(zbuf, pr.y, B);
            }
            const SegDst<FAST> dst(jb.y + (size_t)(b * jb.c_out + c) * jb.L, (long long)j * N - jb.shift, jb.L, N, 2 * N);
            fft_inverse_back<N, NT>(zbuf, jb.plan, dst);
            item_done(jb.cntO + b);
        }
    }
}

// ------------------------------------------------------------------ plan construction
// pair table: thread 0 owns the two self-paired butterflies (klow 0 and N/32); the others take the remaining
// butterflies in increasing order together with their partners (neighbouring threads -> neighbouring
// butterflies: few bank conflicts in the fused pass)
template <int N>
__global__ void fft_plan_pairtab_kernel(ushort2* tab) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    constexpr int NB = N / 16;
    bool used[NB];
    for (int b = 0; b < NB; ++b) used[b] = false;
    const int bh = pos_of_bin<N>(NB / 2) >> 4;
    tab[0] = make_ushort2(0, (unsigned short)bh);
    used[0] = used[bh] = true;
    int t = 1;
    for (int b = 1; b < NB; ++b) {
        if (used[b]) continue;
        const int bp = pos_of_bin<N>(NB - bin_of_pos<N>(16 * b)) >> 4;
        tab[t++] = make_ushort2((unsigned short)b, (unsigned short)bp);
        used[b] = used[bp] = true;
    }
}
// half-twiddle of every pair slot: exp(-i pi k / N), k = klow(butterfly) + (N/16) rk
template <int N>
__global__ void fft_plan_pairs_kernel(float2* ht, const ushort2* __restrict__ tab) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < N / 2) {
        constexpr int NTT = N / 32;
        const int t = slot % NTT, rk = (slot / NTT) & 7, half = slot / (N / 4);
        const int b = half ? tab[t].y : tab[t].x;
        const int k = bin_of_pos<N>(16 * b) + (N / 16) * rk;
        const double a = (double)k / (double)N;
        ht[slot] = make_float2((float)cospi(a), (float)(-sinpi(a)));
    }
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
static int g_long_n = 4096;  // partition size of the long-filter path (tunable: gfx_fir_set_tuning).  Measured on B200 at the
                             // BASELINE reverb shape (profiles/r02_*): 4096-point FFT kernels run 4 CTAs per SM (955 us for the three
                             // FFT kernels vs 1350 us at 8192 points, 2 CTAs per SM); fir_mac2_kernel keeps the 24 partitions in one pass
static int g_sweep_mb = 6144; // spectra workspace per sweep of the long-filter path (gfx_fir_set_sweep_mb); fewer, larger launches
                              // win: 768 / 1536 / 3072 MiB = 1.473 / 1.390 / 1.342 ms at the BASELINE reverb shape (profiles/r02_sweep_size_mac3.txt)
static int g_mac_form = 2;   // 0: fir_mac_kernel (<= 12 partitions per pass) when it applies; 1: fir_mac2_kernel;
                             // 2: fir_mac3_kernel (packed) when it applies (P <= 24, <= 32 blocks per row), else fir_mac2_kernel
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
    int batch, cx, ch; long long L; int Nh; int shift; int h_rep;
    const float2* plan; unsigned char* ws; size_t ws_bytes; cudaStream_t stream;
    bool fast_x, fast_h;  // every (even, odd) sample pair of x / y resp. h is 8-byte aligned and never straddles a row end
};

template <int N>
static int run_ols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    const int rows = a.batch * c_out, hrows = (a.batch / a.h_rep) * a.ch;
    const size_t need = (size_t)hrows * (N / 2) * sizeof(float4);
    if (!a.ws || a.ws_bytes < need) return GFX_ERR_WORKSPACE;
    float4* Hs = (float4*)a.ws;
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[device_slot()];
    if (!configured) {
        if (set_smem(fir_spectrum_kernel<N, true>, smem) || set_smem(fir_spectrum_kernel<N, false>, smem) ||
            set_smem(fir_ols_kernel<N, true>, smem) || set_smem(fir_ols_kernel<N, false>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    if (a.fast_h) fir_spectrum_kernel<N, true><<<hrows, fir_nt(N), smem, a.stream>>>(a.fs, Hs, 0, a.Nh, a.Nh, 1, a.plan);
    else fir_spectrum_kernel<N, false><<<hrows, fir_nt(N), smem, a.stream>>>(a.fs, Hs, 0, a.Nh, a.Nh, 1, a.plan);
    GFX_LAUNCH_CHECK();
    const int pre = (a.Nh - 1 + 3) & ~3;
    const int hop = (2 * N - pre) & ~3;
    const long long total = a.L + a.shift;
    const long long nblk = (total + hop - 1) / hop;
    if (nblk * rows > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
    RowMap rm{c_out, a.cx, a.ch, a.h_rep};
    const unsigned grid = (unsigned)(nblk * rows);
    if (a.fast_x) fir_ols_kernel<N, true><<<grid, fir_nt(N), smem, a.stream>>>(a.x, a.y, Hs, rm, a.L, pre, hop, a.shift, (int)nblk, a.plan);
    else fir_ols_kernel<N, false><<<grid, fir_nt(N), smem, a.stream>>>(a.x, a.y, Hs, rm, a.L, pre, hop, a.shift, (int)nblk, a.plan);
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

template <int N>
static int run_upols(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    int P; long long nblk; size_t per_item;
    upols_geometry(a.cx, a.ch, a.L, a.Nh, a.shift, N, P, nblk, per_item);
    if (!a.ws || a.ws_bytes < per_item) return GFX_ERR_WORKSPACE;
    long long chunk = (long long)(a.ws_bytes / per_item);
    if (chunk > a.batch) chunk = a.batch;
    if (a.h_rep > 1 && chunk >= a.h_rep) chunk -= chunk % a.h_rep;  // sweeps of whole filter runs (fir_mac2_kernel walks a run per CTA)
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[device_slot()];
    if (!configured) {
        if (set_smem(fir_spectrum_kernel<N, true>, smem) || set_smem(fir_spectrum_kernel<N, false>, smem) ||
            set_smem(fir_xspec_kernel<N, true>, smem) || set_smem(fir_xspec_kernel<N, false>, smem) ||
            set_smem(fir_inv_kernel<N, true>, smem) || set_smem(fir_inv_kernel<N, false>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    constexpr int half = N / 2;
    RowMap rm{c_out, a.cx, a.ch, a.h_rep};
    for (long long b0 = 0; b0 < a.batch; b0 += chunk) {
        const int nb = (int)((a.batch - b0 < chunk) ? a.batch - b0 : chunk);
        const int hb0 = (int)(b0 / a.h_rep), nh = (int)((b0 + nb - 1) / a.h_rep) + 1 - hb0;  // filters of this sweep
        float4* Hs = (float4*)a.ws;
        float4* Xs = Hs + (size_t)nh * a.ch * P * half;
        float4* Ys = Xs + (size_t)nb * a.cx * nblk * half;
        const int hrow0 = hb0 * a.ch, xrow0 = (int)b0 * a.cx, row0 = (int)b0 * c_out;
        if ((long long)nb * a.cx * nblk > 0x7fffffffLL || (long long)nb * c_out * nblk > 0x7fffffffLL ||
            (long long)nb * c_out > 65535) return GFX_ERR_UNSUPPORTED;
        const unsigned gh = (unsigned)(nh * a.ch * P), gx = (unsigned)(nb * a.cx * nblk);
        if (a.fast_h) fir_spectrum_kernel<N, true><<<gh, fir_nt(N), smem, a.stream>>>(a.fs, Hs, hrow0, a.Nh, N, P, a.plan);
        else fir_spectrum_kernel<N, false><<<gh, fir_nt(N), smem, a.stream>>>(a.fs, Hs, hrow0, a.Nh, N, P, a.plan);
        GFX_LAUNCH_CHECK();
        if (a.fast_x) fir_xspec_kernel<N, true><<<gx, fir_nt(N), smem, a.stream>>>(a.x, Xs, xrow0, a.L, (int)nblk, a.plan);
        else fir_xspec_kernel<N, false><<<gx, fir_nt(N), smem, a.stream>>>(a.x, Xs, xrow0, a.L, (int)nblk, a.plan);
        GFX_LAUNCH_CHECK();
        if (g_mac_form == 2 && P <= MAC2_MAX_PC && nblk <= MAC3_MAXBLK) {
            // packed form: all partitions in one pass, the row's input blocks resident in shared memory
            const dim3 grid(2 * half / MAC3_NT, c_out, nb);
            const int pc = P <= 4 ? 4 : (P <= 8 ? 8 : (P <= 12 ? 12 : (P <= 16 ? 16 : (P <= 20 ? 20 : 24))));
#define GFX_MAC3_CASE(PCV) case PCV: \
    if (nblk == MAC3_MAXBLK) fir_mac3_kernel<PCV, true><<<grid, MAC3_NT, 0, a.stream>>>((const float2*)Xs, (const float2*)Hs, (float2*)Ys, rm, xrow0, hrow0, (int)b0, P, (int)nblk, 2 * half); \
    else fir_mac3_kernel<PCV, false><<<grid, MAC3_NT, 0, a.stream>>>((const float2*)Xs, (const float2*)Hs, (float2*)Ys, rm, xrow0, hrow0, (int)b0, P, (int)nblk, 2 * half); \
    break;
            switch (pc) {
                GFX_MAC3_CASE(4) GFX_MAC3_CASE(8) GFX_MAC3_CASE(12) GFX_MAC3_CASE(16) GFX_MAC3_CASE(20) GFX_MAC3_CASE(24)
            }
#undef GFX_MAC3_CASE
            GFX_LAUNCH_CHECK();
        } else if (g_mac_form == 0 && P <= MAC_MAX_PC) {
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
        } else {
            // (walking the batch items that share a filter in one CTA -- bpc = h_rep -- was measured: slower; the
            //  shared filter spectra are L2 hits anyway)
            const int bpc = 1;
            const int groups = (nb + bpc - 1) / bpc;
            {
                // one complex bin per thread: up to 24 partitions per pass, inputs through a cp.async ring
                const dim3 grid(2 * half / MAC2_NT, groups * c_out);
                for (int p0 = 0; p0 < P; p0 += MAC2_MAX_PC) {
                    const int rem = P - p0 < MAC2_MAX_PC ? P - p0 : MAC2_MAX_PC;
                    const int pc = rem <= 4 ? 4 : (rem <= 8 ? 8 : (rem <= 12 ? 12 : (rem <= 16 ? 16 : (rem <= 20 ? 20 : 24))));
                    const int acc = p0 > 0;
#define GFX_MAC2_CASE(PCV) case PCV: fir_mac2_kernel<PCV><<<grid, MAC2_NT, 0, a.stream>>>((const float2*)Xs, (const float2*)Hs, (float2*)Ys, rm, xrow0, hrow0, (int)b0, nb, bpc, P, p0, (int)nblk, 2 * half, acc); break;
                    switch (pc) {
                        GFX_MAC2_CASE(4) GFX_MAC2_CASE(8) GFX_MAC2_CASE(12) GFX_MAC2_CASE(16) GFX_MAC2_CASE(20) GFX_MAC2_CASE(24)
                    }
#undef GFX_MAC2_CASE
                    GFX_LAUNCH_CHECK();
                }
            }
        }
        const unsigned gy = (unsigned)(nb * c_out * nblk);
        if (a.fast_x) fir_inv_kernel<N, true><<<gy, fir_nt(N), smem, a.stream>>>(Ys, a.y, row0, a.L, (int)nblk, a.shift, a.plan);
        else fir_inv_kernel<N, false><<<gy, fir_nt(N), smem, a.stream>>>(Ys, a.y, row0, a.L, (int)nblk, a.shift, a.plan);
        GFX_LAUNCH_CHECK();
    }
    return GFX_OK;
}

static int g_upols_d = 4;  // pipeline look-ahead (batch items); ring slots = 2 D + 4
static int upols_r() { return 2 * g_upols_d + 4; }
static size_t upols_ctr_bytes(int batch) { return (256 + (size_t)5 * batch * sizeof(int) + 255) / 256 * 256; }

template <int N>
static int run_upols_pipeline(const FirArgs& a) {
    const int c_out = a.cx > a.ch ? a.cx : a.ch;
    int P; long long nblk; size_t per_item;
    upols_geometry(a.cx, a.ch, a.L, a.Nh, a.shift, N, P, nblk, per_item);
    const size_t ctr = upols_ctr_bytes(a.batch);
    if (!a.ws || a.ws_bytes < ctr + (size_t)upols_r() * per_item) return GFX_ERR_WORKSPACE;
    const size_t smem = (size_t)fft_smem_slots(N) * sizeof(pk2);
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[device_slot()];
    if (!configured) {
        if (set_smem(fir_upols_pipeline_kernel<N, true>, smem) || set_smem(fir_upols_pipeline_kernel<N, false>, smem)) return GFX_ERR_CUDA;
        configured = true;
    }
    constexpr int half = N / 2;
    UpolsJob j;
    j.x = a.x; j.fs = a.fs; j.y = a.y;
    unsigned char* w = a.ws;
    j.ticket = (unsigned int*)w;
    int* cnt = (int*)(w + 256);
    j.cntH = cnt; j.cntX = cnt + a.batch; j.cntM = cnt + 2 * (size_t)a.batch; j.cntMH = cnt + 3 * (size_t)a.batch;
    j.cntO = cnt + 4 * (size_t)a.batch;
    j.Hs = (float4*)(w + ctr);
    j.Xs = j.Hs + (size_t)upols_r() * a.ch * P * half;
    j.Ys = j.Xs + (size_t)upols_r() * a.cx * nblk * half;
    j.batch = a.batch; j.cx = a.cx; j.ch = a.ch; j.c_out = c_out; j.h_rep = a.h_rep;
    j.L = a.L; j.Nh = a.Nh; j.P = P; j.nblk = (int)nblk; j.shift = a.shift; j.R = upols_r(); j.D = g_upols_d;
    j.plan = a.plan;
    GFX_CUDA_CHECK(cudaMemsetAsync(w, 0, ctr, a.stream));
    const bool fast = a.fast_x && a.fast_h;
    int occ = 0;
    if (fast) GFX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fir_upols_pipeline_kernel<N, true>, fir_nt(N), smem));
    else GFX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fir_upols_pipeline_kernel<N, false>, fir_nt(N), smem));
    if (occ < 1) return GFX_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)(device_info().sm_count * occ);
    if (fast) fir_upols_pipeline_kernel<N, true><<<grid, fir_nt(N), smem, a.stream>>>(j);
    else fir_upols_pipeline_kernel<N, false><<<grid, fir_nt(N), smem, a.stream>>>(j);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

static int g_long_mode = 0;  // 0: four kernels per sweep (fastest on B200: 1.94 vs 2.16 ms at the BASELINE reverb shape);
                             // 1: one persistent pipelined launch, spectra L2-resident (2.4x less DRAM traffic), P <= 12

static int fir_dispatch(const FirArgs& a) {
    const int n = pick_fft_size(a.Nh);
    if (a.Nh <= 16384) {
        if (n == 1024) return run_ols<1024>(a);
        if (n == 4096) return run_ols<4096>(a);
        if (n == 8192) return run_ols<8192>(a);
        return run_ols<16384>(a);
    }
    const int P = (a.Nh + n - 1) / n;
    if (g_long_mode == 1 && P <= MAC_MAX_PC && n == 8192 && !a.fs.act) {
        int Pg; long long nblk; size_t per_item;
        upols_geometry(a.cx, a.ch, a.L, a.Nh, a.shift, n, Pg, nblk, per_item);
        if (a.ws_bytes >= upols_ctr_bytes(a.batch) + (size_t)upols_r() * per_item && (long long)a.batch * 5 < 0x7fffffffLL)
            return run_upols_pipeline<8192>(a);
    }
    if (n == 4096) return run_upols<4096>(a);
    if (n == 8192) return run_upols<8192>(a);
    return run_upols<16384>(a);
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

int gfx_fir_set_long_mode(int mode, int lookahead) {
    if (mode != 0 && mode != 1) return GFX_ERR_INVALID;
    if (lookahead < 0 || lookahead > 16) return GFX_ERR_INVALID;
    gfx::g_long_mode = mode;
    if (lookahead) gfx::g_upols_d = lookahead;
    return GFX_OK;
}

int gfx_fir_set_sweep_mb(int mb) {
    if (mb < 8 || mb > 16384) return GFX_ERR_INVALID;
    gfx::g_sweep_mb = mb;
    return GFX_OK;
}

int gfx_fir_set_mac_form(int form) {
    if (form < 0 || form > 2) return GFX_ERR_INVALID;
    gfx::g_mac_form = form;
    return GFX_OK;
}

int gfx_fir_fft_size(int filter_len) { return filter_len <= 0 ? GFX_ERR_INVALID : gfx::pick_fft_size(filter_len); }

size_t gfx_fft_plan_bytes(int n) {
    if (!gfx::plan_ok(n)) return 0;
    return ((size_t)n / 2 + (size_t)gfx::plan_total(n)) * sizeof(float2) + ((size_t)n / 32) * sizeof(ushort2);
}

int gfx_fft_plan_init(void* plan, int n, void* stream) {
    using namespace gfx;
    if (!plan || !plan_ok(n)) return GFX_ERR_INVALID;
    float2* base = (float2*)plan;
    cudaStream_t st = (cudaStream_t)stream;
    const int gb = (n / 2 + 255) / 256;
    ushort2* tab = (ushort2*)(base + n / 2 + plan_total(n));
    if (n == 1024) { fft_plan_pairtab_kernel<1024><<<1, 32, 0, st>>>(tab); fft_plan_pairs_kernel<1024><<<gb, 256, 0, st>>>(base, tab); }
    else if (n == 4096) { fft_plan_pairtab_kernel<4096><<<1, 32, 0, st>>>(tab); fft_plan_pairs_kernel<4096><<<gb, 256, 0, st>>>(base, tab); }
    else if (n == 8192) { fft_plan_pairtab_kernel<8192><<<1, 32, 0, st>>>(tab); fft_plan_pairs_kernel<8192><<<gb, 256, 0, st>>>(base, tab); }
    else { fft_plan_pairtab_kernel<16384><<<1, 32, 0, st>>>(tab); fft_plan_pairs_kernel<16384><<<gb, 256, 0, st>>>(base, tab); }
    ++g_gfx_launch_count;
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
    if (gfx::g_long_mode == 1 && P <= gfx::MAC_MAX_PC && n == 8192)
        return gfx::upols_ctr_bytes(batch) + (size_t)gfx::upols_r() * per_item;  // pipelined launch: a ring of item slots
    // spectra of up to g_sweep_mb worth of batch items per sweep (every kernel of a sweep then has many full waves)
    size_t items = ((size_t)gfx::g_sweep_mb << 20) / per_item;
    if (items < 1) items = 1;
    if (items > (size_t)batch) items = batch;
    return items * per_item;
}

static int fir_conv_common(const float* x, gfx::FilterSrc fs, float* y, int batch, int cx, int ch, long long L,
                           int filter_len, int zerophase, int filter_repeat, const void* plan, void* workspace,
                           size_t workspace_bytes, void* stream) {
    using namespace gfx;
    if (!x || !fs.h || !y || !plan) return GFX_ERR_INVALID;
    if (batch <= 0 || cx <= 0 || ch <= 0 || L <= 0 || filter_len <= 0) return GFX_ERR_INVALID;
    if (filter_repeat <= 0 || batch % filter_repeat != 0) return GFX_ERR_INVALID;
    if (cx != ch && cx != 1 && ch != 1) return GFX_ERR_INVALID;
    FirArgs a{x, fs, y, batch, cx, ch, L, filter_len, zerophase ? filter_len / 2 : 0, filter_repeat, (const float2*)plan,
              (unsigned char*)workspace, workspace_bytes, (cudaStream_t)stream, false, false};
    a.fast_x = (((uintptr_t)x | (uintptr_t)y) % 8 == 0) && (L % 2 == 0) && (a.shift % 2 == 0);
    a.fast_h = ((uintptr_t)fs.h % 8 == 0) && (filter_len % 2 == 0);
    return fir_dispatch(a);
}

int gfx_fir_conv_f32(const float* x, const float* h, float* y, int batch, int cx, int ch, long long L,
                     int filter_len, int zerophase, int filter_repeat, const void* plan, void* workspace,
                     size_t workspace_bytes, void* stream) {
    return fir_conv_common(x, gfx::FilterSrc{h, nullptr, 0, 0, 1}, y, batch, cx, ch, L, filter_len, zerophase, filter_repeat,
                           plan, workspace, workspace_bytes, stream);
}

int gfx_fir_filter_f32(const float* x, const float* fir_raw, float* y, int batch, int cx, int ch, long long L,
                       int filter_len, const void* plan, void* workspace, size_t workspace_bytes, void* stream) {
    if (!fir_raw || !workspace || batch <= 0 || ch <= 0 || filter_len <= 0) return GFX_ERR_INVALID;
    const size_t ebytes = ((size_t)batch * sizeof(float) + 255) / 256 * 256;
    if (workspace_bytes < ebytes) return GFX_ERR_WORKSPACE;
    float* energy = (float*)workspace;
    gfx::fir_tanh_energy_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(fir_raw, energy, ch, filter_len);
    GFX_LAUNCH_CHECK();
    return fir_conv_common(x, gfx::FilterSrc{fir_raw, energy, 0, 1, ch}, y, batch, cx, ch, L, filter_len, 0, 1, plan,
                           (unsigned char*)workspace + ebytes, workspace_bytes - ebytes, stream);
}

int gfx_fir_conv_midside_ir_f32(const float* x, const float* ir_raw, const float* energy, float* y, int batch, int cx,
                                long long L, int ir_len, int ms_to_lr, int filter_repeat, const void* plan,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (!energy) return GFX_ERR_INVALID;
    return fir_conv_common(x, gfx::FilterSrc{ir_raw, energy, ms_to_lr ? 1 : 0, 0, 1}, y, batch, cx, 2, L, ir_len, 0,
                           filter_repeat, plan, workspace, workspace_bytes, stream);
}

}  // extern "C"
