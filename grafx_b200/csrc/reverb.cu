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
// One CTA owns one (batch item, mid|side) row and walks its frames in tiles of 16.  A frame is
// inverted by 16 threads: the 384-point real inverse DFT is split as n = 6q + r into six 64-point
// inverse DFTs of  G_r[k0] = sum_j X[k0 + 64 j] e^{2 pi i (k0 + 64 j) r / 384};  residues are paired
// (G_2s + i G_2s+1) so that three complex radix-4 FFT-64 (shared memory, digit-reversed input)
// yield the 384 real samples.  Supported geometry: n_fft = 384, hop = 192 (the reference default).
#include "common.cuh"

namespace gfx {

constexpr int RV_NFFT = 384, RV_HOP = 192, RV_BINS = 193, RV_FT = 16;  // frames per tile
constexpr int RV_NT = 256;

__device__ __forceinline__ int drev4_3(int p) { return ((p & 3) << 4) | (p & 12) | ((p >> 4) & 3); }

struct ReverbParams {
    const float2* noise;      // [noise_batch?][2][bins][frames]
    long long noise_bstride;  // elements between batch items (0: shared noise)
    const float* h0;          // [B][2][bins]
    const float* hd;          // [B][2][bins]
    const float* genv;        // [B][2][frames] or null
    const float* window;      // [384]
    float* ir;                // [B][2][ir_len]  (mid/side, un-normalised)
    float* energy;            // [B][2]  sum_t ir^2
    int frames, ir_len;
};

__global__ void __launch_bounds__(RV_NT) reverb_ir_kernel(const ReverbParams p) {
    extern __shared__ __align__(16) unsigned char rv_smem[];
    float2* tw = reinterpret_cast<float2*>(rv_smem);                        // [384] e^{+2 pi i t / 384}
    float2 (*X)[RV_BINS + 1] = reinterpret_cast<float2 (*)[RV_BINS + 1]>(tw + RV_NFFT);   // [16][194] masked spectra
    float (*fr)[3][64] = reinterpret_cast<float (*)[3][64]>(X + RV_FT);     // [16][3][64]
    float (*fi)[3][64] = fr + RV_FT;
    float (*frame)[RV_NFFT] = reinterpret_cast<float (*)[RV_NFFT]>(fi + RV_FT);  // [16][384] windowed frames
    float* a0 = reinterpret_cast<float*>(frame + RV_FT);                    // [193] H0/8 ... actually H0
    float* a1 = a0 + RV_BINS + 3;                                           // [193] softplus(Hd)
    float* win = a1 + RV_BINS + 3;                                          // [384]
    float* carry = win + RV_NFFT;                                           // [192]
    float* red = carry + RV_HOP;                                            // [8]

    const int tid = threadIdx.x;
    const int row = blockIdx.x;  // b * 2 + ch
    const int b = row >> 1, ch = row & 1;
    const float2* noise = p.noise + (size_t)b * p.noise_bstride + (size_t)ch * RV_BINS * p.frames;
    const float* genv = p.genv ? p.genv + (size_t)row * p.frames : nullptr;

    for (int t = tid; t < RV_NFFT; t += RV_NT) {
        float s, c;
        sincospif(2.f * (float)t / (float)RV_NFFT, &s, &c);
        tw[t] = make_float2(c, s);
        win[t] = p.window[t];
    }
    for (int k = tid; k < RV_BINS; k += RV_NT) {
        a0[k] = p.h0[(size_t)row * RV_BINS + k];
        const float d = p.hd[(size_t)row * RV_BINS + k];
        a1[k] = d > 20.f ? d : log1pf(expf(d));  // torch softplus
    }
    for (int n = tid; n < RV_HOP; n += RV_NT) carry[n] = 0.f;
    __syncthreads();

    const int g = tid >> 4, tau = tid & 15;  // frame slot within the tile, thread within the frame group
    float esum = 0.f;
    const int last_tp = RV_HOP + p.ir_len - 1;           // last needed overlap-add position
    const int ntiles = last_tp / (RV_FT * RV_HOP) + 1;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int m0 = tile * RV_FT;
        // 1. masked spectra of the tile's frames (coalesced over the frame axis)
        for (int i = tid; i < RV_BINS * RV_FT; i += RV_NT) {
            const int k = i / RV_FT, f = i - k * RV_FT;
            const int m = m0 + f;
            float2 v = make_float2(0.f, 0.f);
            if (m < p.frames) {
                const float2 nz = __ldg(noise + (size_t)k * p.frames + m);
                float lg = a0[k] - a1[k] * (float)m;
                if (genv) lg += genv[m];
                const float mk = expf(lg * 0.125f);
                v = make_float2(nz.x * mk, (k == 0 || k == RV_BINS - 1) ? 0.f : nz.y * mk);
            }
            X[f][k] = v;
        }
        __syncthreads();
        // 2. residue spectra G_r[k0], paired and scattered in digit-reversed order
        if (m0 + g < p.frames) {
            for (int k0 = tau; k0 < 64; k0 += 16) {
                float2 xf[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int k = k0 + 64 * j;
                    if (k <= 192) xf[j] = X[g][k];
                    else { const float2 c = X[g][384 - k]; xf[j] = make_float2(c.x, -c.y); }
                }
                float2 G[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    float gr = 0.f, gi = 0.f;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const float2 w = tw[((k0 + 64 * j) * r) % RV_NFFT];
                        gr = fmaf(xf[j].x, w.x, gr); gr = fmaf(-xf[j].y, w.y, gr);
                        gi = fmaf(xf[j].x, w.y, gi); gi = fmaf(xf[j].y, w.x, gi);
                    }
                    G[r] = make_float2(gr, gi);
                }
                const int pos = drev4_3(k0);
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    fr[g][s][pos] = G[2 * s].x - G[2 * s + 1].y;
                    fi[g][s][pos] = G[2 * s].y + G[2 * s + 1].x;
                }
            }
        }
        __syncwarp();
        // 3. three inverse FFT-64 (decimation in time, radix 4), 16 threads = 16 butterflies per pass.
        // NB every lane of the warp reaches each __syncwarp (the two frame groups of a warp may differ
        // in validity at the end of the frame axis).
        const bool fvalid = (m0 + g < p.frames);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
            if (fvalid) {
                const int ST = 1 << (2 * pass);              // 1, 4, 16
                const int TWS = RV_NFFT / (4 * ST);          // 384/M : 96, 24, 6
                const int j = tau & (ST - 1);
                const int i0 = ((tau - j) << 2) + j;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    float* re = fr[g][s];
                    float* im = fi[g][s];
                    float2 u0 = make_float2(re[i0], im[i0]);
                    float2 u1 = make_float2(re[i0 + ST], im[i0 + ST]);
                    float2 u2 = make_float2(re[i0 + 2 * ST], im[i0 + 2 * ST]);
                    float2 u3 = make_float2(re[i0 + 3 * ST], im[i0 + 3 * ST]);
                    if (pass > 0) {
                        const float2 w1 = tw[j * TWS], w2 = tw[2 * j * TWS], w3 = tw[3 * j * TWS];
                        u1 = make_float2(u1.x * w1.x - u1.y * w1.y, u1.x * w1.y + u1.y * w1.x);
                        u2 = make_float2(u2.x * w2.x - u2.y * w2.y, u2.x * w2.y + u2.y * w2.x);
                        u3 = make_float2(u3.x * w3.x - u3.y * w3.y, u3.x * w3.y + u3.y * w3.x);
                    }
                    const float s02r = u0.x + u2.x, s02i = u0.y + u2.y, d02r = u0.x - u2.x, d02i = u0.y - u2.y;
                    const float s13r = u1.x + u3.x, s13i = u1.y + u3.y, d13r = u1.x - u3.x, d13i = u1.y - u3.y;
                    re[i0] = s02r + s13r; im[i0] = s02i + s13i;
                    re[i0 + ST] = d02r - d13i; im[i0 + ST] = d02i + d13r;
                    re[i0 + 2 * ST] = s02r - s13r; im[i0 + 2 * ST] = s02i - s13i;
                    re[i0 + 3 * ST] = d02r + d13i; im[i0 + 3 * ST] = d02i - d13r;
                }
            }
            __syncwarp();
        }
        if (fvalid) {
            // 4. interleave residues, scale, window
            for (int n = tau; n < RV_NFFT; n += 16) {
                const int q = n / 6, r = n - 6 * q;
                const float v = (r & 1) ? fi[g][r >> 1][q] : fr[g][r >> 1][q];
                frame[g][n] = v * (1.f / (float)RV_NFFT) * win[n];
            }
        }
        __syncthreads();
        // 5. overlap-add, envelope division, trim
        for (int i = tid; i < RV_FT * RV_HOP; i += RV_NT) {
            const int f = i / RV_HOP, n = i - f * RV_HOP;
            const int m = m0 + f;
            const int tp = m * RV_HOP + n;
            const int t = tp - RV_HOP;
            if (t < 0 || t >= p.ir_len) continue;
            float val = 0.f, env = 0.f;
            if (m < p.frames) { val = frame[f][n]; env = win[n] * win[n]; }
            if (m >= 1 && m - 1 < p.frames) {
                val += (f > 0) ? frame[f - 1][n + RV_HOP] : carry[n];
                env += win[n + RV_HOP] * win[n + RV_HOP];
            }
            val = val / env;
            p.ir[(size_t)row * p.ir_len + t] = val;
            esum = fmaf(val, val, esum);
        }
        __syncthreads();
        for (int n = tid; n < RV_HOP; n += RV_NT)
            carry[n] = (m0 + RV_FT - 1 < p.frames) ? frame[RV_FT - 1][n + RV_HOP] : 0.f;
        __syncthreads();
    }
    // energy of the row
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    if ((tid & 31) == 0) red[tid >> 5] = esum;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < RV_NT / 32; ++w) s += red[w];
        p.energy[row] = s;
    }
}

// ms_to_lr (optional) + unit-energy normalisation, in place: ir [B][2][T]
__global__ void __launch_bounds__(256) reverb_finalize_kernel(float* ir, const float* __restrict__ energy, int batch,
                                                              int T, int to_lr) {
    const long long total = (long long)batch * T;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long b = i / T, t = i - b * T;
        const float e0 = energy[2 * b], e1 = energy[2 * b + 1];
        float* pm = ir + (b * 2) * T + t;
        float* ps = ir + (b * 2 + 1) * T + t;
        const float m = *pm, s = *ps;
        if (to_lr) {
            // mean_c sum_t (m +- s)^2 = sum m^2 + sum s^2
            const float sc = 1.f / sqrtf(e0 + e1 + 1e-12f);
            *pm = (m + s) * sc;
            *ps = (m - s) * sc;
        } else {
            const float sc = 1.f / sqrtf(0.5f * (e0 + e1) + 1e-12f);
            *pm = m * sc;
            *ps = s * sc;
        }
    }
}

}  // namespace gfx

extern "C" int gfx_reverb_ir_f32(const float* noise_stft, long long noise_batch_stride, const float* init_log_magnitude,
                                 const float* delta_log_magnitude, const float* gain_env_log_magnitude,
                                 const float* window, float* ir, float* energy_ws, int batch, int n_fft, int hop,
                                 int ir_len, int ms_to_lr, void* stream) {
    using namespace gfx;
    if (!noise_stft || !init_log_magnitude || !delta_log_magnitude || !window || !ir || !energy_ws) return GFX_ERR_INVALID;
    if (batch <= 0 || ir_len <= 0) return GFX_ERR_INVALID;
    if (n_fft != RV_NFFT || hop != RV_HOP) return GFX_ERR_UNSUPPORTED;
    ReverbParams p;
    p.noise = (const float2*)noise_stft;
    p.noise_bstride = noise_batch_stride;
    p.h0 = init_log_magnitude; p.hd = delta_log_magnitude; p.genv = gain_env_log_magnitude;
    p.window = window; p.ir = ir; p.energy = energy_ws;
    p.frames = 1 + ir_len / hop;
    p.ir_len = ir_len;
    const size_t smem = sizeof(float2) * (RV_NFFT + RV_FT * (RV_BINS + 1)) +
                        sizeof(float) * (2 * RV_FT * 3 * 64 + RV_FT * RV_NFFT + 2 * (RV_BINS + 3) + RV_NFFT + RV_HOP + 8);
    static bool configured = false;
    if (!configured) {
        GFX_CUDA_CHECK(cudaFuncSetAttribute(reverb_ir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    reverb_ir_kernel<<<batch * 2, RV_NT, smem, (cudaStream_t)stream>>>(p);
    GFX_CUDA_CHECK(cudaGetLastError());
    const long long total = (long long)batch * ir_len;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    reverb_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ir, energy_ws, batch, ir_len, ms_to_lr);
    GFX_CUDA_CHECK(cudaGetLastError());
    return GFX_OK;
}
