// Impulse-response synthesis of STFTMaskedNoiseReverb: masked noise STFT -> inverse STFT ->
// (mid/side -> left/right) -> unit-energy normalisation, without materialising the masked STFT.
//
// Replaces (reference, /root/reference/src/grafx/processors/reverb.py):
//   :189-200 compute_stft_mask   mask = exp((H0 - softplus(Hd) * m [+ G[m]]) / 8)
//   :161-187 compute_ir          torch.istft(noise_stft * mask, n_fft=384, hop=192, hann, length=ir_len)
//   :215-228 _process_*          ms_to_lr (pseudo_midside) and normalize_impulse (core/utils.py:14-18)
// torch.istft(center=True) is: per frame irfft * window, overlap-add at m*hop, divide by the
// overlap-added squared window, drop n_fft/2 leading samples, keep ir_len.
//
// One CTA owns 16 consecutive frames of one (batch item, mid|side) row and emits the 15 hops they
// fully determine (one frame of overlap between neighbouring CTAs is recomputed: +6.7 % work, no
// carried state, 34 x rows independent CTAs).  A frame is inverted by 16 threads: the 384-point
// inverse real DFT is a 192-point complex inverse FFT of the re-tangled half spectrum
// (fftcore.cuh: retangle_pair), done as 192 = 16 x 12:
//   thread t: inverse DFT-12 (3 x radix-4, 4 x radix-3) of Z[t + 16 m], twiddle e^{2 pi i r t / 192},
//   exchange through shared memory, thread r < 12: inverse DFT-16 -> z[r + 12 s] = x[2n] + i x[2n+1].
// Row energies are reduced per CTA into `partial` and summed in a fixed order (deterministic).
// Supported geometry: n_fft = 384, hop = 192 (the reference default).
#include "common.cuh"
#include "fftcore.cuh"

namespace gfx {

constexpr int RV_NFFT = 384, RV_HOP = 192, RV_BINS = 193, RV_N = 192;
constexpr int RV_FT = 16;         // frames per CTA
constexpr int RV_HT = RV_FT - 1;  // hops emitted per CTA
constexpr int RV_NT = 256;
constexpr int RV_ZS = 209;        // 64-bit slots per frame work buffer (odd: conflict-free across frames)

struct ReverbParams {
    const float2* noise;      // [noise_batch?][2][bins][frames]
    long long noise_bstride;  // elements between batch items (0: shared noise)
    const float* h0;          // [B][2][bins]
    const float* hd;          // [B][2][bins]
    const float* genv;        // [B][2][frames] or null
    const float* window;      // [384]
    float* ir;                // [B][2][ir_len]  (mid/side, un-normalised)
    float* partial;           // [B*2][tiles]  sum of ir^2 per CTA
    int frames, ir_len, tiles, vec_ok, to_lr;
};

__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// inverse DFT-3 (unnormalised): y_r = sum_m a_m e^{+2 pi i m r / 3}
__device__ __forceinline__ void idft3(pk2& a0, pk2& a1, pk2& a2) {
    const pk2 s = pk_add(a1, a2), d = pk_sub(a1, a2);
    const pk2 t = pk_fma(s, pk_dup(-0.5f), a0);
    const pk2 r = pk_mul(pk_swap(d), pk_make(-0.86602540378443865f, 0.86602540378443865f));  // i (sqrt3/2) d
    a0 = pk_add(a0, s);
    a1 = pk_add(t, r);
    a2 = pk_sub(t, r);
}

// inverse DFT-12 in registers.  In: a[m] natural order.  Out: register a[r1 + 3 r2] holds A[r2 + 4 r1].
__device__ __forceinline__ void idft12(pk2 (&a)[12]) {
#pragma unroll
    for (int m1 = 0; m1 < 3; ++m1) r4<true>(a[m1], a[m1 + 3], a[m1 + 6], a[m1 + 9]);
    // a[m1 + 3 r2] *= e^{+2 pi i m1 r2 / 12}   (tw_apply<true> multiplies by the conjugate of its argument)
    const float C = 0.86602540378443865f;
    a[1 + 3] = tw_apply<true>(a[1 + 3], make_float2(C, -0.5f));      // e^{i pi/6}
    a[1 + 6] = tw_apply<true>(a[1 + 6], make_float2(0.5f, -C));      // e^{i pi/3}
    a[1 + 9] = mul_minus_i<true>(a[1 + 9]);                          // i
    a[2 + 3] = tw_apply<true>(a[2 + 3], make_float2(0.5f, -C));      // e^{i pi/3}
    a[2 + 6] = tw_apply<true>(a[2 + 6], make_float2(-0.5f, -C));     // e^{2 i pi/3}
    a[2 + 9] = pk_mul(a[2 + 9], pk_dup(-1.f));                       // -1
#pragma unroll
    for (int r2 = 0; r2 < 4; ++r2) idft3(a[3 * r2], a[3 * r2 + 1], a[3 * r2 + 2]);
}

// smem carve-up (per CTA = one batch item, both mid and side rows, 16 frames each)
constexpr int RV_ZB_BYTES = 2 * RV_FT * RV_ZS * 8;          // pk2 zb[2][16][RV_ZS]
constexpr int RV_FR_BYTES = 2 * RV_FT * RV_NFFT * 4;        // float fr[2][16][384]
constexpr int RV_TW_BYTES = (RV_N + RV_N / 2) * 8;          // float2 w192[192], hw[96]
constexpr int RV_FL_FLOATS = RV_NFFT + 4 * (RV_BINS + 3) + 2 * RV_FT + 2 * 16 + 8;
constexpr int RV_SMEM_BYTES = RV_ZB_BYTES + RV_FR_BYTES + RV_TW_BYTES + RV_FL_FLOATS * 4;

__global__ void __launch_bounds__(2 * RV_NT, 2) reverb_ir_kernel(const ReverbParams p) {
    extern __shared__ __align__(16) unsigned char rv_smem[];
    pk2* zb_all = reinterpret_cast<pk2*>(rv_smem);                                    // [2][16][RV_ZS]
    float* fr_all = reinterpret_cast<float*>(rv_smem + RV_ZB_BYTES);                  // [2][16][384] windowed frames
    float2* w192 = reinterpret_cast<float2*>(rv_smem + RV_ZB_BYTES + RV_FR_BYTES);    // [192] e^{+2 pi i j / 192}
    float2* hw = w192 + RV_N;                                                         // [96]  e^{-i pi k / 192}
    float* win = reinterpret_cast<float*>(hw + RV_N / 2);                             // [384]
    float* a0_all = win + RV_NFFT;                                                    // [2][196] H0
    float* a1_all = a0_all + 2 * (RV_BINS + 3);                                       // [2][196] softplus(Hd)
    float* xn_all = a1_all + 2 * (RV_BINS + 3);                                       // [2][16] Nyquist bin per frame
    float* red = xn_all + 2 * RV_FT;                                                  // [2][16]

    const int ch = threadIdx.x >> 8, tid = threadIdx.x & (RV_NT - 1);  // channel (mid | side), thread within it
    const int tile = blockIdx.x % p.tiles, b = blockIdx.x / p.tiles;
    const int row = 2 * b + ch;
    pk2* zb = zb_all + ch * (RV_FT * RV_ZS);
    float* fr = fr_all + ch * (RV_FT * RV_NFFT);
    float* a0 = a0_all + ch * (RV_BINS + 3);
    float* a1 = a1_all + ch * (RV_BINS + 3);
    float* xn = xn_all + ch * RV_FT;
    const float2* noise = p.noise + (size_t)b * p.noise_bstride + (size_t)ch * RV_BINS * p.frames;
    const float* genv = p.genv ? p.genv + (size_t)row * p.frames : nullptr;
    const int m0 = tile * RV_HT;  // first frame of the CTA; hops m0+1 .. m0+15

    for (int t = threadIdx.x; t < RV_NFFT; t += 2 * RV_NT) {
        win[t] = p.window[t];
        if (t < RV_N) {
            float sn, c;
            sincospif((float)t * (2.f / (float)RV_N), &sn, &c);
            w192[t] = make_float2(c, sn);
        } else if (t < RV_N + RV_N / 2) {
            float sn, c;
            sincospif((float)(t - RV_N) * (1.f / (float)RV_N), &sn, &c);
            hw[t - RV_N] = make_float2(c, -sn);
        }
    }
    // mask exponent in base 2: exp((H0 - softplus(Hd) m + G[m]) / 8) = ex2((H0 - softplus(Hd) m + G[m]) log2(e) / 8)
    constexpr float kScale = 1.4426950408889634f * 0.125f;
    for (int k = tid; k < RV_BINS; k += RV_NT) {
        a0[k] = p.h0[(size_t)row * RV_BINS + k] * kScale;
        const float d = p.hd[(size_t)row * RV_BINS + k];
        a1[k] = (d > 20.f ? d : log1pf(expf(d))) * kScale;  // torch softplus
    }
    __syncthreads();

    // 1 + 2. masked spectra of the CTA's frames, re-tangled into the 192-point complex spectrum on the way: lane = frame
    //    (global reads coalesced over the frame axis), 16 bin lanes; a thread takes the bins k and 192 - k of a pair,
    //    masks both (mask = exp((H0 - softplus(Hd) m [+ G[m]]) / 8), reverb.py:189-200), re-tangles them in registers and
    //    writes each slot once (the spectrum used to be written, re-read pair by pair and written again).  Bins 0 and
    //    192 are real and share slot 0; bin 96 pairs with itself.
    {
        const int f = tid & (RV_FT - 1), kq = tid >> 4;
        const int m = m0 + f;
        const bool fv = m < p.frames;
        const float fm = (float)m;
        const float ge = (genv && fv) ? genv[m] * kScale : 0.f;
        const float2* np = noise + (fv ? m : 0);
        pk2* zf = zb + f * RV_ZS;
        auto masked = [&](int k) {
            const float2 nz = __ldg(np + (size_t)k * p.frames);
            const float mk = fv ? ex2_approx((a0[k] - a1[k] * fm) + ge) : 0.f;
            return make_float2(nz.x * mk, nz.y * mk);
        };
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int k = kq + 16 * i;
            if (k == 0) {
                const float2 x0 = masked(0), xN = masked(RV_BINS - 1), xh = masked(RV_N / 2);
                zf[0] = pk_make(0.5f * (x0.x + xN.x), 0.5f * (x0.x - xN.x));
                zf[RV_N / 2] = pk_make(xh.x, -xh.y);
            } else {
                const PairA a = retangle_pair(masked(k), masked(RV_N - k), hw[k]);
                zf[k] = pk_make(a.k.x, a.k.y);
                zf[RV_N - k] = pk_make(a.m.x, a.m.y);
            }
        }
    }
    __syncthreads();

    const int g = tid >> 4, tau = tid & 15;  // frame slot, thread within the frame group
    pk2* z = zb + g * RV_ZS;
    // 3. thread tau: inverse DFT-12 over Z[tau + 16 m], twiddle, to B[r][tau] at slot r * 17 + tau
    {
        pk2 a[12];
#pragma unroll
        for (int m = 0; m < 12; ++m) a[m] = z[tau + 16 * m];
        __syncwarp();  // every Z has been read before B overwrites the buffer
        idft12(a);
#pragma unroll
        for (int r1 = 0; r1 < 3; ++r1) {
#pragma unroll
            for (int r2 = 0; r2 < 4; ++r2) {
                const int r = r2 + 4 * r1;
                pk2 v = a[r1 + 3 * r2];
                if (r > 0) v = tw_apply<false>(v, w192[r * tau]);
                z[r * 17 + tau] = v;
            }
        }
    }
    __syncwarp();
    // 4. thread r < 12: inverse DFT-16 over B[r][.] -> z[r + 12 s], scaled and windowed into the frame
    if (tau < 12) {
        pk2 bq[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) bq[t] = z[tau * 17 + t];
        r16<true>(bq);
        float* frame = fr + g * RV_NFFT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sidx = 4 * j + i;
                const int n = tau + 12 * sidx;
                float re, im;
                pk_split(bq[4 * i + j], re, im);
                const float2 w2 = *reinterpret_cast<const float2*>(win + 2 * n);
                *reinterpret_cast<float2*>(frame + 2 * n) =
                    make_float2(re * (1.f / (float)RV_N) * w2.x, im * (1.f / (float)RV_N) * w2.y);
            }
        }
    }
    __syncthreads();

    // 5. overlap-add, envelope division, trim -- both channels by every thread, so that the rows can be
    //    written as mid/side or, for the pseudo_midside mode, as left/right = mid +- side (reverb.py:225-228)
    float em = 0.f, es = 0.f;
    float* row0 = p.ir + (size_t)(2 * b) * p.ir_len;
    float* row1 = row0 + p.ir_len;
    const float* frm = fr_all;
    const float* frs = fr_all + RV_FT * RV_NFFT;
    if (p.vec_ok) {
        for (int i = threadIdx.x; i < RV_HT * (RV_HOP / 4); i += 2 * RV_NT) {
            const int hh = i / (RV_HOP / 4), n = (i - hh * (RV_HOP / 4)) * 4;
            const int h = m0 + 1 + hh;                 // hop index = frame index of its first half
            const int t = (h - 1) * RV_HOP + n;        // output sample (tp - hop); frame h-1 exists whenever t < ir_len
            if (t >= p.ir_len) continue;
            const int o1 = (hh + 1) * RV_NFFT + n, o2 = hh * RV_NFFT + n + RV_HOP;
            const float4 am = *reinterpret_cast<const float4*>(frm + o1);       // zero beyond the last frame
            const float4 cm = *reinterpret_cast<const float4*>(frm + o2);
            const float4 as = *reinterpret_cast<const float4*>(frs + o1);
            const float4 cs = *reinterpret_cast<const float4*>(frs + o2);
            const float4 w1 = *reinterpret_cast<const float4*>(win + n);
            const float4 w2 = *reinterpret_cast<const float4*>(win + n + RV_HOP);
            const float e1 = h < p.frames ? 1.f : 0.f;
            const float ex = fmaf(e1 * w1.x, w1.x, w2.x * w2.x), ey = fmaf(e1 * w1.y, w1.y, w2.y * w2.y);
            const float ez = fmaf(e1 * w1.z, w1.z, w2.z * w2.z), ew = fmaf(e1 * w1.w, w1.w, w2.w * w2.w);
            // envelope division as a multiplication by the SFU reciprocal (error ~1 ulp), shared by both channels
            const float rx = __frcp_rn(ex), ry = __frcp_rn(ey), rz = __frcp_rn(ez), rw = __frcp_rn(ew);
            const float4 vm = make_float4((am.x + cm.x) * rx, (am.y + cm.y) * ry, (am.z + cm.z) * rz, (am.w + cm.w) * rw);
            const float4 vs = make_float4((as.x + cs.x) * rx, (as.y + cs.y) * ry, (as.z + cs.z) * rz, (as.w + cs.w) * rw);
            em = fmaf(vm.x, vm.x, em); em = fmaf(vm.y, vm.y, em); em = fmaf(vm.z, vm.z, em); em = fmaf(vm.w, vm.w, em);
            es = fmaf(vs.x, vs.x, es); es = fmaf(vs.y, vs.y, es); es = fmaf(vs.z, vs.z, es); es = fmaf(vs.w, vs.w, es);
            if (p.to_lr) {
                *reinterpret_cast<float4*>(row0 + t) = make_float4(vm.x + vs.x, vm.y + vs.y, vm.z + vs.z, vm.w + vs.w);
                *reinterpret_cast<float4*>(row1 + t) = make_float4(vm.x - vs.x, vm.y - vs.y, vm.z - vs.z, vm.w - vs.w);
            } else {
                *reinterpret_cast<float4*>(row0 + t) = vm;
                *reinterpret_cast<float4*>(row1 + t) = vs;
            }
        }
    } else {
        for (int i = threadIdx.x; i < RV_HT * RV_HOP; i += 2 * RV_NT) {
            const int hh = i / RV_HOP, n = i - hh * RV_HOP;
            const int h = m0 + 1 + hh;
            const int t = (h - 1) * RV_HOP + n;
            if (t >= p.ir_len) continue;
            const float e1 = h < p.frames ? 1.f : 0.f;
            const float env = fmaf(e1 * win[n], win[n], win[n + RV_HOP] * win[n + RV_HOP]);
            const float vm = (frm[(hh + 1) * RV_NFFT + n] + frm[hh * RV_NFFT + n + RV_HOP]) / env;
            const float vs = (frs[(hh + 1) * RV_NFFT + n] + frs[hh * RV_NFFT + n + RV_HOP]) / env;
            em = fmaf(vm, vm, em);
            es = fmaf(vs, vs, es);
            row0[t] = p.to_lr ? vm + vs : vm;
            row1[t] = p.to_lr ? vm - vs : vs;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        em += __shfl_xor_sync(0xffffffffu, em, o);
        es += __shfl_xor_sync(0xffffffffu, es, o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = em; red[16 + (threadIdx.x >> 5)] = es; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float sum = 0.f;
        for (int w = 0; w < 2 * RV_NT / 32; ++w) sum += red[16 * threadIdx.x + w];
        p.partial[(size_t)(2 * b + threadIdx.x) * p.tiles + tile] = sum;
    }
}

__global__ void reverb_energy_kernel(const float* __restrict__ partial, float* __restrict__ energy, int rows, int tiles) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        float s = 0.f;
        for (int t = 0; t < tiles; ++t) s += partial[(size_t)r * tiles + t];
        energy[r] = s;
    }
}

// unit-energy normalisation in place (normalize_impulse, core/utils.py:14-18): ir [B][2][T], rows already in
// their final channel layout; energy = sum_t of the squared RAW mid / side rows
__global__ void __launch_bounds__(256) reverb_finalize_kernel(float* ir, const float* __restrict__ energy, int batch,
                                                              int T, int to_lr) {
    const long long total = (long long)batch * 2 * T;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long b = i / (2LL * T);
        const float e0 = energy[2 * b], e1 = energy[2 * b + 1];
        // left/right: mean_c sum_t (m +- s)^2 = sum m^2 + sum s^2
        const float sc = to_lr ? 1.f / sqrtf(e0 + e1 + 1e-12f) : 1.f / sqrtf(0.5f * (e0 + e1) + 1e-12f);
        ir[i] *= sc;
    }
}

// ---- FilteredNoiseShapingReverb: ir[b,c,t] = sum_k noise[c,k,t0+t] * gain[b,c,k] * (exp(t decay[b,c,k]) - fg[b,c,k] exp(t fade[b,c,k]))
// (reference: reverb.py:364-380, a [B, C, K, T] broadcast + sum upstream: 2.9 GB of intermediates at B = 512, T = 60000).
// One thread = 4 consecutive taps of one (item, channel) row; the K band envelopes are evaluated on the fly
// (ex2.approx); the row energy is reduced per CTA into `partial` (summed in fixed order afterwards).
constexpr int NS_NT = 256, NS_MAX_BANDS = 32;

__global__ void __launch_bounds__(NS_NT) noise_shaping_ir_kernel(const float* __restrict__ noise, long long noise_len,
                                                                 long long t0, const float* __restrict__ decay,
                                                                 const float* __restrict__ gain,
                                                                 const float* __restrict__ fade,
                                                                 const float* __restrict__ fade_gain,
                                                                 float* __restrict__ ir, float* __restrict__ partial,
                                                                 int channels, int bands, int ir_len, int chunks) {
    __shared__ float sd[NS_MAX_BANDS], sg[NS_MAX_BANDS], sf[NS_MAX_BANDS], sfg[NS_MAX_BANDS], red[NS_NT / 32];
    const int row = blockIdx.x / chunks, chunk = blockIdx.x - row * chunks;
    const int c = row % channels;
    constexpr float LOG2E = 1.4426950408889634f;
    if ((int)threadIdx.x < bands) {
        const size_t i = (size_t)row * bands + threadIdx.x;
        sd[threadIdx.x] = decay[i] * LOG2E;
        sg[threadIdx.x] = gain[i];
        sf[threadIdx.x] = fade ? fade[i] * LOG2E : 0.f;
        sfg[threadIdx.x] = fade ? fade_gain[i] : 0.f;
    }
    __syncthreads();
    const int t = (chunk * NS_NT + (int)threadIdx.x) * 4;
    float esum = 0.f;
    if (t < ir_len) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* nrow = noise + (size_t)c * bands * noise_len + t0 + t;
        const int n = ir_len - t < 4 ? ir_len - t : 4;
        for (int k = 0; k < bands; ++k) {
            const float* nk = nrow + (size_t)k * noise_len;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e < n) {
                    const float tt = (float)(t + e);
                    float env = ex2_approx(tt * sd[k]);
                    if (fade) env -= sfg[k] * ex2_approx(tt * sf[k]);
                    acc[e] = fmaf(__ldg(nk + e) * sg[k], env, acc[e]);
                }
            }
        }
        float* out = ir + (size_t)row * ir_len + t;
        for (int e = 0; e < n; ++e) { out[e] = acc[e]; esum = fmaf(acc[e], acc[e], esum); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = esum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        for (int w = 0; w < NS_NT / 32; ++w) sum += red[w];
        partial[(size_t)row * chunks + chunk] = sum;
    }
}

// ------------------------------------------------------------------------------------------------
// General STFT geometry (n_fft a power of two, 32..4096, any hop <= n_fft): the same pipeline in two plain kernels.
// The tuned kernel above covers the reference default (384 / 192); this one keeps the other constructor arguments of
// STFTMaskedNoiseReverb(n_fft=..., hop_length=...) on the device path.
//   1. one CTA per (row, frame): masked half spectrum -> Hermitian-extended n_fft-point inverse FFT (radix 2, in
//      shared memory, twiddles from sincospi) -> real part / n_fft * window -> frames [rows][frames][n_fft] (workspace);
//   2. overlap-add of the <= ceil(n_fft / hop) frames that cover an output sample (ascending frame order), division by
//      the overlap-added squared window, trim, mid/side or left/right rows, per-CTA partial energies.
constexpr int RG_NT = 256;
constexpr int RG_MAX_NFFT = 4096;

__global__ void __launch_bounds__(RG_NT) reverb_frames_generic_kernel(const ReverbParams p, float* __restrict__ frames_out,
                                                                      int n_fft, int log2n) {
    extern __shared__ __align__(16) unsigned char rg_smem[];
    float2* z = reinterpret_cast<float2*>(rg_smem);  // [n_fft]
    const int m = blockIdx.x % p.frames;
    const int row = blockIdx.x / p.frames;  // 2 * b + channel
    const int b = row >> 1, ch = row & 1;
    const int bins = n_fft / 2 + 1;
    const float2* noise = p.noise + (size_t)b * p.noise_bstride + (size_t)ch * bins * p.frames + m;
    const float ge = p.genv ? p.genv[(size_t)row * p.frames + m] : 0.f;
    const float fm = (float)m;
    // mask = exp((H0 - softplus(Hd) m [+ G[m]]) / 8)  (reverb.py:189-200); bit-reversed placement for the DIT passes
    for (int k = threadIdx.x; k < bins; k += RG_NT) {
        const float d = p.hd[(size_t)row * bins + k];
        const float sp = d > 20.f ? d : log1pf(expf(d));
        const float mk = expf((p.h0[(size_t)row * bins + k] - sp * fm + ge) * 0.125f);
        const float2 nz = noise[(size_t)k * p.frames];
        const float2 v = make_float2(nz.x * mk, nz.y * mk);
        z[__brev((unsigned)k) >> (32 - log2n)] = v;
        if (k > 0 && k < n_fft / 2) z[__brev((unsigned)(n_fft - k)) >> (32 - log2n)] = make_float2(v.x, -v.y);
    }
    __syncthreads();
    for (int s = 1; s <= log2n; ++s) {
        const int half = 1 << (s - 1);
        for (int i = threadIdx.x; i < n_fft / 2; i += RG_NT) {
            const int pos = i & (half - 1);
            const int a = ((i >> (s - 1)) << s) + pos, c = a + half;
            float sn, cs;
            sincospif((float)pos / (float)half, &sn, &cs);  // e^{+i pi pos / half} = e^{+2 pi i pos / 2^s}
            const float2 x = z[c];
            const float2 t = make_float2(x.x * cs - x.y * sn, x.x * sn + x.y * cs);
            const float2 u = z[a];
            z[a] = make_float2(u.x + t.x, u.y + t.y);
            z[c] = make_float2(u.x - t.x, u.y - t.y);
        }
        __syncthreads();
    }
    float* out = frames_out + ((size_t)row * p.frames + m) * n_fft;
    const float scale = 1.f / (float)n_fft;
    for (int n = threadIdx.x; n < n_fft; n += RG_NT) out[n] = z[n].x * scale * p.window[n];
}

__global__ void __launch_bounds__(RG_NT) reverb_ola_generic_kernel(const ReverbParams p, const float* __restrict__ frames_in,
                                                                   int n_fft, int hop) {
    const int tile = blockIdx.x % p.tiles, b = blockIdx.x / p.tiles;
    const int t = tile * RG_NT + threadIdx.x;
    float vm = 0.f, vs = 0.f;
    if (t < p.ir_len) {
        const int tp = t + n_fft / 2;  // position in the centred overlap-add buffer
        int m_hi = tp / hop;
        if (m_hi > p.frames - 1) m_hi = p.frames - 1;
        int m_lo = tp - n_fft + 1 <= 0 ? 0 : (tp - n_fft + hop) / hop;  // ceil((tp - n_fft + 1) / hop)
        const float* fm = frames_in + (size_t)(2 * b) * p.frames * n_fft;
        const float* fs = fm + (size_t)p.frames * n_fft;
        float env = 0.f;
        for (int m = m_lo; m <= m_hi; ++m) {
            const int n = tp - m * hop;
            const float w = p.window[n];
            env = fmaf(w, w, env);
            vm += fm[(size_t)m * n_fft + n];
            vs += fs[(size_t)m * n_fft + n];
        }
        vm /= env;
        vs /= env;
        float* row0 = p.ir + (size_t)(2 * b) * p.ir_len;
        float* row1 = row0 + p.ir_len;
        row0[t] = p.to_lr ? vm + vs : vm;
        row1[t] = p.to_lr ? vm - vs : vs;
    }
    float em = vm * vm, es = vs * vs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        em += __shfl_xor_sync(0xffffffffu, em, o);
        es += __shfl_xor_sync(0xffffffffu, es, o);
    }
    __shared__ float red[2][RG_NT / 32];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = em; red[1][threadIdx.x >> 5] = es; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float sum = 0.f;
        for (int w = 0; w < RG_NT / 32; ++w) sum += red[threadIdx.x][w];
        p.partial[(size_t)(2 * b + threadIdx.x) * p.tiles + tile] = sum;
    }
}

static bool reverb_fast_geometry(int n_fft, int hop) { return n_fft == RV_NFFT && hop == RV_HOP; }
static bool reverb_generic_geometry(int n_fft, int hop) {
    return n_fft >= 32 && n_fft <= RG_MAX_NFFT && (n_fft & (n_fft - 1)) == 0 && hop >= 1 && hop <= n_fft;
}
static int reverb_generic_tiles(int ir_len) { return (ir_len + RG_NT - 1) / RG_NT; }

static int reverb_tiles(int ir_len) {
    const int last_hop = (RV_HOP - 1 + ir_len) / RV_HOP;  // hops 1 .. last_hop carry output samples
    return (last_hop + RV_HT - 1) / RV_HT;
}

}  // namespace gfx

extern "C" {

size_t gfx_reverb_ir_workspace_bytes(int batch, int n_fft, int hop, int ir_len) {
    if (batch <= 0 || ir_len <= 0) return 0;
    if (gfx::reverb_fast_geometry(n_fft, hop)) return ((size_t)batch * 2 * (gfx::reverb_tiles(ir_len) + 1)) * sizeof(float);
    if (!gfx::reverb_generic_geometry(n_fft, hop)) return 0;
    // partial energies, then the windowed frames [batch * 2][frames][n_fft]
    const size_t partial = ((size_t)batch * 2 * gfx::reverb_generic_tiles(ir_len) * sizeof(float) + 255) / 256 * 256;
    return partial + (size_t)batch * 2 * (size_t)(1 + ir_len / hop) * n_fft * sizeof(float);
}

size_t gfx_noise_shaping_ir_workspace_bytes(int batch, int channels, int ir_len) {
    if (batch <= 0 || channels <= 0 || ir_len <= 0) return 0;
    const int chunks = (ir_len + gfx::NS_NT * 4 - 1) / (gfx::NS_NT * 4);
    return (size_t)batch * channels * chunks * sizeof(float);
}

int gfx_noise_shaping_ir_f32(const float* noise, long long noise_len, long long noise_offset, const float* decay,
                             const float* gain, const float* fade, const float* fade_gain, float* ir, float* energy,
                             void* workspace, size_t workspace_bytes, int batch, int channels, int bands, int ir_len,
                             void* stream) {
    using namespace gfx;
    if (!noise || !decay || !gain || !ir || !energy) return GFX_ERR_INVALID;
    if (batch <= 0 || channels <= 0 || bands <= 0 || ir_len <= 0 || noise_offset < 0 || noise_offset + ir_len > noise_len) return GFX_ERR_INVALID;
    if ((fade == nullptr) != (fade_gain == nullptr)) return GFX_ERR_INVALID;
    if (bands > NS_MAX_BANDS) return GFX_ERR_UNSUPPORTED;
    const int chunks = (ir_len + NS_NT * 4 - 1) / (NS_NT * 4);
    const long long rows = (long long)batch * channels;
    if (!workspace || workspace_bytes < (size_t)rows * chunks * sizeof(float)) return GFX_ERR_WORKSPACE;
    if (rows * chunks > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    noise_shaping_ir_kernel<<<(unsigned)(rows * chunks), NS_NT, 0, st>>>(noise, noise_len, noise_offset, decay, gain, fade,
                                                                       fade_gain, ir, (float*)workspace, channels, bands,
                                                                       ir_len, chunks);
    GFX_LAUNCH_CHECK();
    reverb_energy_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>((const float*)workspace, energy, (int)rows, chunks);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

// mode: 0 raw mid/side + energies, 1 mid/side normalised, 2 left/right normalised, 3 raw left/right + energies
int gfx_reverb_ir_f32(const float* noise_stft, long long noise_batch_stride, const float* init_log_magnitude,
                      const float* delta_log_magnitude, const float* gain_env_log_magnitude,
                      const float* window, float* ir, float* energy, void* workspace, size_t workspace_bytes,
                      int batch, int n_fft, int hop, int ir_len, int mode, void* stream) {
    using namespace gfx;
    if (!noise_stft || !init_log_magnitude || !delta_log_magnitude || !window || !ir || !energy) return GFX_ERR_INVALID;
    if (batch <= 0 || ir_len <= 0 || mode < 0 || mode > 3) return GFX_ERR_INVALID;
    const bool fast = reverb_fast_geometry(n_fft, hop);
    if (!fast && !reverb_generic_geometry(n_fft, hop)) return GFX_ERR_UNSUPPORTED;
    const int tiles = fast ? reverb_tiles(ir_len) : reverb_generic_tiles(ir_len);
    if (!workspace || workspace_bytes < gfx_reverb_ir_workspace_bytes(batch, n_fft, hop, ir_len)) return GFX_ERR_WORKSPACE;
    if ((long long)batch * tiles > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
    ReverbParams p;
    p.noise = (const float2*)noise_stft;
    p.noise_bstride = noise_batch_stride;
    p.h0 = init_log_magnitude; p.hd = delta_log_magnitude; p.genv = gain_env_log_magnitude;
    p.window = window; p.ir = ir; p.partial = (float*)workspace;
    p.frames = 1 + ir_len / hop;
    p.ir_len = ir_len;
    p.tiles = tiles;
    p.vec_ok = ((uintptr_t)ir % 16 == 0) && (ir_len % 4 == 0);
    p.to_lr = mode >= 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (fast) {
        static bool configured_dev[64] = {false};
        bool& configured = configured_dev[device_slot()];
        if (!configured) {
            GFX_CUDA_CHECK(cudaFuncSetAttribute(reverb_ir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RV_SMEM_BYTES));
            configured = true;
        }
        reverb_ir_kernel<<<(unsigned)(batch * tiles), 2 * RV_NT, RV_SMEM_BYTES, st>>>(p);
        GFX_LAUNCH_CHECK();
    } else {
        if ((long long)batch * 2 * p.frames > 0x7fffffffLL) return GFX_ERR_UNSUPPORTED;
        const size_t partial = ((size_t)batch * 2 * tiles * sizeof(float) + 255) / 256 * 256;
        float* frames_ws = (float*)((unsigned char*)workspace + partial);
        int log2n = 0;
        while ((1 << log2n) < n_fft) ++log2n;
        reverb_frames_generic_kernel<<<(unsigned)(batch * 2 * p.frames), RG_NT, (size_t)n_fft * sizeof(float2), st>>>(
            p, frames_ws, n_fft, log2n);
        GFX_LAUNCH_CHECK();
        reverb_ola_generic_kernel<<<(unsigned)(batch * tiles), RG_NT, 0, st>>>(p, frames_ws, n_fft, hop);
        GFX_LAUNCH_CHECK();
    }
    reverb_energy_kernel<<<(batch * 2 + 127) / 128, 128, 0, st>>>(p.partial, energy, batch * 2, tiles);
    GFX_LAUNCH_CHECK();
    if (mode == 1 || mode == 2) {
        const long long total = (long long)batch * 2 * ir_len;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)device_info().sm_count * 8;
        if (blocks > cap) blocks = cap;
        reverb_finalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(ir, energy, batch, ir_len, mode == 2);
        GFX_LAUNCH_CHECK();
    }
    return GFX_OK;
}

}  // extern "C"
