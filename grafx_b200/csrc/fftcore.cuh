// Register-level FFT building blocks shared by csrc/fir.cu and csrc/reverb.cu: packed complex
// arithmetic (one complex = one 64-bit register pair), radix-4 / radix-16 butterflies and the pair
// algebra of the real-FFT untangling.  Header-only, private to csrc/.
#pragma once
#include "common.cuh"

namespace gfx {

// ------------------------------------------------------------------ complex helpers
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// v * w (forward) or v * conj(w) (inverse), v packed
template <bool INV>
__device__ __forceinline__ pk2 tw_apply(pk2 v, float2 w) {
    float re, im;
    pk_split(v, re, im);
    if constexpr (!INV) return pk_make(fmaf(re, w.x, -(im * w.y)), fmaf(re, w.y, im * w.x));
    else return pk_make(fmaf(re, w.x, im * w.y), fmaf(im, w.x, -(re * w.y)));
}
template <bool INV>
__device__ __forceinline__ pk2 mul_minus_i(pk2 v) {  // forward: v * (-i); inverse: v * (+i)
    float re, im;
    pk_split(v, re, im);
    return INV ? pk_make(-im, re) : pk_make(im, -re);
}
__device__ __forceinline__ float2 ld_tw(const float2* p) { return __ldg(p); }

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
    // a[q0 + 4 r0] *= omega_16^(q0 r0)   (forward constants; tw_apply<INV> conjugates)
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
    const float2 w1 = make_float2(C1, -S1), w2 = make_float2(H, -H), w3 = make_float2(S1, -C1);
    const float2 w6 = make_float2(-H, -H), w9 = make_float2(-C1, S1);
    a[1 + 4] = tw_apply<INV>(a[1 + 4], w1); a[2 + 4] = tw_apply<INV>(a[2 + 4], w2); a[3 + 4] = tw_apply<INV>(a[3 + 4], w3);
    a[1 + 8] = tw_apply<INV>(a[1 + 8], w2); a[2 + 8] = mul_minus_i<INV>(a[2 + 8]);  a[3 + 8] = tw_apply<INV>(a[3 + 8], w6);
    a[1 + 12] = tw_apply<INV>(a[1 + 12], w3); a[2 + 12] = tw_apply<INV>(a[2 + 12], w6); a[3 + 12] = tw_apply<INV>(a[3 + 12], w9);
#pragma unroll
    for (int r0 = 0; r0 < 4; ++r0) r4<INV>(a[4 * r0], a[4 * r0 + 1], a[4 * r0 + 2], a[4 * r0 + 3]);
}

// ------------------------------------------------------------------ real-FFT untangling on pair slots
// Z = FFT of z[j] = x[2j] + i x[2j+1] (digit-reversed in smem).  For the pair (k, N-k), w = exp(-i pi k / N):
//   E = (Z_k + conj Z_{N-k}) / 2,  O = -(i/2) w (Z_k - conj Z_{N-k}),  A_k = E + O,  A_{N-k} = conj(E - O).
// Slot 0 is special: { (A_0, A_N) both real, A_{N/2} }.
struct PairA { float2 k, m; };
__device__ __forceinline__ PairA untangle_pair(float2 zk, float2 zm, float2 w) {
    const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
    const float dr = zk.x - zm.x, di = zk.y + zm.y;
    const float2 wd = cmul(w, make_float2(dr, di));
    const float orr = 0.5f * wd.y, oi = -0.5f * wd.x;
    PairA r;
    r.k = make_float2(er + orr, ei + oi);
    r.m = make_float2(er - orr, -(ei - oi));
    return r;
}
// inverse of the above: from Y_k, Y_{N-k} to the packed Z'_k, Z'_{N-k}
__device__ __forceinline__ PairA retangle_pair(float2 yk, float2 ym, float2 w) {
    const float er = 0.5f * (yk.x + ym.x), ei = 0.5f * (yk.y - ym.y);
    const float dr = yk.x - ym.x, di = yk.y + ym.y;
    const float2 wd = cmulc(make_float2(dr, di), w);  // O' = (i/2) conj(w) D'
    const float orr = -0.5f * wd.y, oi = 0.5f * wd.x;
    PairA r;
    r.k = make_float2(er + orr, ei + oi);
    r.m = make_float2(er - orr, -(ei - oi));
    return r;
}

}  // namespace gfx
