// Fused dynamic-range processors: Compressor / NoiseGate (and serial chains of them) in ONE pass
// over HBM: energy -> envelope smoother -> log -> knee -> gain smoother -> gain * x.
//
// Replaces (reference, /root/reference/src/grafx/processors):
//   dynamics.py:361-419,443-489  Compressor.forward + knees
//   dynamics.py:598-651,675-721  NoiseGate.forward + knees
//   core/envelope.py:34-60       TruncatedOnePoleIIRFilter  (16384-tap FFT convolution + relu upstream)
//   core/envelope.py:84-101      Ballistics -> torchcomp.compressor_core (sequential attack/release)
//   container.py:116-140         SerialChain of dynamics processors (stages fused here)
//
// Layout: like csrc/biquad.cu -- a row (= one batch item, all C channels) is cut into tiles of
// NT*32 samples, staged in shared memory with the XOR swizzle; thread t owns 32 consecutive
// samples.  Tiles of a row are chained through a few words of smoother state (common.cuh).
//
// One-pole smoother ("iir"): upstream convolves with h[n] = (1-a) a^n, n < N, then relu.  That
// truncated response obeys the exact recursion
//      T[n] = a T[n-1] + (1-a) (u[n] - a^N u[n-N]),
// which is evaluated with a chunked scan (zero-state pass, shuffle scan with a^(32*2^j), re-run
// from the true state).  The lagged term only matters when a^N is representable (a within
// ~100/N of 1): then u[n-N] is re-derived from x (first stage energy) or read back from a
// caller-provided history buffer (other smoothers).  Constants a^k are formed in double.
// Ballistics (y = (1-c) y + c u, c = at if u < y else rt, y[-1] = 1) is not associative: one
// thread walks the tile; parallelism comes from rows (small tiles, many resident CTAs).
#include "common.cuh"

namespace gfx {

constexpr int DYN_MAX_STAGES = 4;

struct SmootherDesc {
    int kind;        // 0 none, 1 truncated one-pole, 2 ballistics
    const float* z;  // [rows, 1] (iir) or [rows, 2] (ballistics)
    float* hist;     // [rows, L] input history for the truncation tail (may be null)
};
struct StageDesc {
    int kind;        // 0 compressor, 1 noise gate
    int knee;        // 0 hard, 1 quadratic, 2 exponential, 3 quadratic as shipped in ApproxNoiseGate (gate only)
    int log_domain;  // gain smoother runs on the log-gain
    const float* log_threshold;
    const float* log_ratio;
    const float* log_knee;
    SmootherDesc pre, post;
};
struct DynParams {
    const float* x;
    float* y;
    int batch, C;
    long long L;
    int tiles;
    unsigned int n_items;
    unsigned int* ticket;
    int* flags;
    float* state;  // [batch][2*n_stages]
    float* tables; // [batch][DYN_ROW_FLOATS] per-row constants (dynamics_tables_kernel)
    int n_stages;
    int iir_len;
    int aligned;
    // envelope mode (gfx_envelope_f32: the stand-alone smoothers / envelope followers, core/envelope.py:10-101,
    // dynamics.py:745-790): one stage, no knee, y [batch, 1, L] = the smoothed detector signal (or its log)
    int* lookback;  // [batch] ballistics launches: samples of warm-up that make a chunk independent of its past (>= 0: the row
                    // is done by dynamics_spec_kernel), or -1 (the row-owner kernel walks the row); null: walk every row
    int spec_ch;    // samples per thread of dynamics_spec_kernel
    int param_rep;  // runs of this many consecutive rows share one parameter row (render_grafx's 4-D sources; 1 otherwise):
                    // only dynamics_tables_kernel reads the parameter tensors, it writes every row's table
    int envelope;   // 0: dynamics processors; 1: envelope output
    int detect;     // envelope mode: 0 mean_c x^2, 1 mean_c |x|, 2 x itself (C == 1)
    int env_log;    // envelope mode: y = log(envelope + 1e-5)
    StageDesc st[DYN_MAX_STAGES];
};

__device__ __forceinline__ float softplus_torch(float v) {
    // torch.nn.functional.softplus (beta=1, threshold=20)
    return v > 20.f ? v : log1pf(expf(v));
}

// Per-row constants, formed once per call by dynamics_tables_kernel (double precision where it matters)
// and staged in shared memory with each tile: the sample loop never touches expf/logf/double.
struct SmootherConst {  // 48 floats
    float alpha, oma;   // one-pole coefficient, 1 - alpha
    float aN;           // alpha^iir_len (0 when it underflows fp32 or when the signal is shorter than iir_len)
    float aW;           // alpha^(32*32): one warp of chunks
    float apow[5];      // alpha^(32 * 2^j)
    float at, rt;       // ballistics attack / release coefficients
    float pad[5];
    float pl[32];       // alpha^(32 * lane)
};
// Knee constants.  Per-sample math uses the SFU intrinsics __logf/__expf (abs error ~1e-6 on the log-energy /
// relative ~2e-6 on the gain, far inside the 1e-4 parity budget).
struct KneeConst {      // 16 floats
    float T, W, lo, hi;       // threshold, knee half-width (quadratic) or width (exponential), T-W, T+W
    float slope;              // log-gain slope outside the knee: 1/R - 1 (compressor) | R - 1 (gate)
    float mid_scale;          // (1/R - 1)/(4W)  |  (1 - R)/(4W)
    float exp_scale;          // exponential knee: (1/R - 1)/W | -exp(lr)/W
    float slope2, mid2, exp2s; // the three scales above times log2(e): the gain is then ex2() of the result
    float Wl2;                // W * log2(e)
    float pad[5];
};
constexpr int DYN_SC_FLOATS = sizeof(SmootherConst) / 4, DYN_KC_FLOATS = sizeof(KneeConst) / 4;
constexpr int DYN_ROW_FLOATS = DYN_MAX_STAGES * (2 * DYN_SC_FLOATS + DYN_KC_FLOATS);  // per-row table

template <int NT, bool RO>
struct DynCtx {
    static constexpr int S = 32;
    static constexpr int NW = NT / 32;
    float4* xs4;     // [C][NT*8] swizzled input tile
    float4* work4;   // [NT*8] scratch tile (ballistics)
    float* wt;       // [2][NW]
    float* s_state;  // [2*DYN_MAX_STAGES]
    const SmootherConst* sc;  // [2*DYN_MAX_STAGES] constants of this row (smem)
    const KneeConst* kc;      // [DYN_MAX_STAGES]
    int tid, lane, warp;
    int walker;      // thread that walks the ballistics recursion
    int t_idx, row;
    long long t0, remain;
    int sync_parity;
    bool have_state;
};

// The ballistics variant (NT == 64) gives a whole ROW to one CTA, which walks its tiles in order: the recursion of a
// row is one long dependent chain anyway, so nothing is gained by spreading its tiles over CTAs and a hand-over through
// global memory (flag, fence, L1 invalidate: ~18 us per tile as measured) is saved.  The smoother states then live
// in shared memory, double-buffered by tile parity.  The scan variant (NT == DYN_SCAN_NT or DYN_SCAN_NT_SMALL) chains tiles across CTAs.
// Threads per CTA of the scan variant (80 registers: 3 CTAs of 256 threads or 6 of 128 per SM).  Measured on B200:
// config 4 (two fused stages, 1024 rows x 65536) 256 threads 0.211 ms, 128 threads 0.193 ms -- the same warps per SM,
// but a block barrier joins 4 warps instead of 8; config 5's compressor (512 rows x 2 x 131072) 0.204 ms vs 0.301 ms --
// with fewer rows than resident CTAs the extra CTAs only queue up behind the tile chain of a row.  So: the small CTA
// when every resident CTA can own a different row, the large one otherwise.
constexpr int DYN_SCAN_NT = 256, DYN_SCAN_NT_SMALL = 128, DYN_SCAN_CTAS_SMALL = 6;
static int g_dyn_scan_nt = 0;  // 0: by row count (below); 128 | 256: forced (gfx_dynamics_set_tuning)
// RO: the CTA owns whole rows (ballistics); otherwise one (tile, row) item at a time (scan variant)

template <int NT, bool RO>
__device__ __forceinline__ float* state_slots(const DynCtx<NT, RO>& cx, int t_idx) {
    return cx.s_state + (RO ? (t_idx & 1) * 2 * DYN_MAX_STAGES : 0);
}

template <int NT, bool RO>
__device__ __forceinline__ void ensure_state(DynCtx<NT, RO>& cx, const DynParams& p) {
    if (cx.have_state) return;
    cx.have_state = true;
    const int ns2 = 2 * p.n_stages;
    if (cx.warp == 0) {
        if (cx.t_idx > 0) {
            if constexpr (!RO) {
                if (cx.lane == 0) chain_wait(p.flags + cx.row, cx.t_idx);
                __syncwarp();
                if (cx.lane < ns2) cx.s_state[cx.lane] = __ldcg(p.state + (size_t)cx.row * ns2 + cx.lane);
            }
        } else if (cx.lane < ns2) {
            const StageDesc& sd = p.st[cx.lane >> 1];
            const int kind = (cx.lane & 1) ? sd.post.kind : sd.pre.kind;
            state_slots<NT, RO>(cx, 0)[cx.lane] = kind == 2 ? 1.f : 0.f;  // ballistics starts from zi = 1
        }
    }
    __syncthreads();
}

// the state a smoother leaves for the next tile of the row
template <int NT, bool RO>
__device__ __forceinline__ void leave_state(DynCtx<NT, RO>& cx, const DynParams& p, int slot, float value) {
    if (cx.t_idx + 1 >= p.tiles) return;
    if constexpr (RO) state_slots<NT, RO>(cx, cx.t_idx + 1)[slot] = value;
    else p.state[(size_t)cx.row * 2 * p.n_stages + slot] = value;
}

// ---- truncated one-pole smoother on the register chunk u[32]; `slot` = index into the row state
template <int NT, bool RO, typename LagFn>
__device__ __forceinline__ void smooth_iir(DynCtx<NT, RO>& cx, const DynParams& p, float (&u)[32],
                                           const SmootherDesc& sm, int slot, LagFn lag_input) {
    constexpr int S = 32, NW = NT / 32;
    const SmootherConst& c = cx.sc[slot];
    const float alpha = c.alpha, oma = c.oma, aN = c.aN;
    if (aN != 0.f) {
        const long long chunk0 = cx.t0 + (long long)cx.tid * S;
        if (sm.hist != nullptr) {
            float* hr = sm.hist + (size_t)cx.row * (size_t)p.L;
#pragma unroll
            for (int i = 0; i < S; ++i)
                if (chunk0 + i < p.L) hr[chunk0 + i] = u[i];
            __threadfence();
            // earlier tiles of the row must have finished writing their history; lagged
            // samples of this very tile come from other threads of this CTA
            ensure_state(cx, p);
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const long long pos = chunk0 + i - p.iir_len;
            const float lagv = (pos >= 0 && chunk0 + i < p.L) ? lag_input(pos) : 0.f;
            u[i] = fmaf(-aN, lagv, u[i]);
        }
    }
    // zero-state pass (on the un-scaled input: the factor 1 - alpha is applied once, at the end)
    float w = 0.f;
#pragma unroll
    for (int i = 0; i < S; ++i) w = fmaf(alpha, w, u[i]);
    float z = w;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float pv = __shfl_up_sync(0xffffffffu, z, 1 << j);
        if (cx.lane >= (1 << j)) z = fmaf(c.apow[j], pv, z);
    }
    const float aW = c.aW;
    float* wt = cx.wt + (cx.sync_parity & 1) * NW;
    cx.sync_parity++;
    if (cx.lane == 31) wt[cx.warp] = z;
    ensure_state(cx, p);  // (first stateful op of the tile also syncs here)
    __syncthreads();
    float s = state_slots<NT, RO>(cx, cx.t_idx)[slot];
    for (int q = 0; q < cx.warp; ++q) s = fmaf(aW, s, wt[q]);
    float ex = __shfl_up_sync(0xffffffffu, z, 1);
    if (cx.lane == 0) ex = 0.f;
    float y = fmaf(c.pl[cx.lane], s, ex);  // (un-scaled) T[-1] of this chunk
    // true pass, scale, relu
#pragma unroll
    for (int i = 0; i < S; ++i) {
        y = fmaf(alpha, y, u[i]);
        u[i] = fmaxf(y * oma, 0.f);
    }
    if (cx.tid == NT - 1) leave_state<NT, RO>(cx, p, slot, y);
}

// The walk itself: NT rows of 32 samples.  Every instruction of the walking lane costs a full warp slot of its
// sub-partition (2 cycles per pipe), so nothing but the chain (FFMA, FFMA, FMNMX per sample) and the 128-bit
// shared-memory accesses is issued: rows are taken eight at a time so that the XOR swizzle (unit c of row r sits at
// physical unit c ^ (r & 7)) and all offsets are compile-time, and two register sets alternate (the next row loads
// while this one is walked) without copies.
template <int NT, bool USE_MIN>
__device__ __forceinline__ float ballistics_walk(float4* wa, const float4* wr, float y, float at, float rt) {
    const float omat = 1.f - at, omrt = 1.f - rt;
    float4 a0[4], r0[4], a1[4], r1[4];  // two register sets of half a row (16 samples) each
#pragma unroll
    for (int c = 0; c < 4; ++c) { a0[c] = wa[c]; r0[c] = wr[c]; }  // row 0: swizzle 0, logical = physical
    static_assert(NT % 8 == 0, "rows are walked in groups of eight");
#pragma unroll 1
    for (int r8 = 0; r8 < NT / 8; ++r8) {
        float4* pa = wa + r8 * 64;        // 8 rows x 8 units
        const float4* pr = wr + r8 * 64;
        const bool more = r8 + 1 < NT / 8;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4(&ca)[4] = h ? a1 : a0;
                float4(&cr)[4] = h ? r1 : r0;
                float4(&na)[4] = h ? a0 : a1;
                float4(&nr)[4] = h ? r0 : r1;
                // next half row: the second half of this row, the first half of the next one, or of the next group
                if (h == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { na[c] = pa[rr * 8 + ((4 + c) ^ rr)]; nr[c] = pr[rr * 8 + ((4 + c) ^ rr)]; }
                } else if (rr < 7) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { na[c] = pa[(rr + 1) * 8 + (c ^ (rr + 1))]; nr[c] = pr[(rr + 1) * 8 + (c ^ (rr + 1))]; }
                } else if (more) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { na[c] = pa[64 + c]; nr[c] = pr[64 + c]; }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float* a = reinterpret_cast<float*>(&ca[c]);
                    const float* b = reinterpret_cast<const float*>(&cr[c]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float ya = fmaf(omat, y, a[k]), yr = fmaf(omrt, y, b[k]);
                        y = USE_MIN ? fminf(ya, yr) : fmaxf(ya, yr);
                        a[k] = y;
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) pa[rr * 8 + ((4 * h + c) ^ rr)] = ca[c];
            }
        }
    }
    return y;
}

// ---- attack/release ballistics on the register chunk (sequential over the tile)
//   y' = (u < y) ? ya : yr,  ya = (1-at) y + at u,  yr = (1-rt) y + rt u.   ya - yr = (at - rt)(u - y), so the select
//   is min(ya, yr) when at >= rt and max(ya, yr) otherwise: the chain is FFMA || FFMA -> FMNMX, no predicate.
// Only one lane walks the tile, but every instruction it issues still occupies its scheduler's pipe for a whole warp
// slot (2 cycles; tools/chain_bench.cu: one SM sub-partition sustains ~2 such chains at full speed, 11-14 cycles
// per step, and halves beyond that), so (a) the products at*u and rt*u are formed by all threads beforehand: the
// walk issues 3 instructions per sample; (b) the walking lane is lane 0 of the warp chosen from the hardware
// warp slot so that the walkers of the CTAs resident on an SM spread over all four sub-partitions.
template <int NT, bool RO>
__device__ __forceinline__ void smooth_ballistics(DynCtx<NT, RO>& cx, const DynParams& p, float (&u)[32],
                                                  const SmootherDesc& sm, int slot) {
    const float at = cx.sc[slot].at, rt = cx.sc[slot].rt;
    float4* wa = cx.work4;            // at * u, overwritten by y
    float4* wr = cx.work4 + NT * 8;   // rt * u
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int idx = swz_unit(cx.tid, c);
        wa[idx] = make_float4(at * u[4 * c], at * u[4 * c + 1], at * u[4 * c + 2], at * u[4 * c + 3]);
        wr[idx] = make_float4(rt * u[4 * c], rt * u[4 * c + 1], rt * u[4 * c + 2], rt * u[4 * c + 3]);
    }
    ensure_state(cx, p);
    __syncthreads();
    if (cx.tid == cx.walker) {
        const float y0 = state_slots<NT, RO>(cx, cx.t_idx)[slot];
        const float yend = (at >= rt) ? ballistics_walk<NT, true>(wa, wr, y0, at, rt) : ballistics_walk<NT, false>(wa, wr, y0, at, rt);
        (void)yend;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = wa[swz_unit(cx.tid, c)];
        u[4 * c] = v.x; u[4 * c + 1] = v.y; u[4 * c + 2] = v.z; u[4 * c + 3] = v.w;
    }
    if (cx.tid == NT - 1) leave_state<NT, RO>(cx, p, slot, u[31]);  // the state after the last sample of a FULL tile
}

// bare SFU ops (no denormal / range fix-up code around them: the arguments here are >= 1e-5 resp. bounded)
__device__ __forceinline__ float lg2_fast(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float ex2_fast(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
constexpr float DYN_LN2 = 0.69314718055994531f, DYN_LOG2E = 1.4426950408889634f;

// u[i] (smoothed energy) -> LINEAR gain exp(G_out - G) of the knee, for stages without a gain smoother: the
// log runs in base 2 (one MUFU), the knee scales carry log2(e), the exp is one ex2.
__device__ __forceinline__ void knee_gain(float (&u)[32], const KneeConst& k, int mode) {
    const float T = k.T, W = k.W, slope = k.slope2, mid = k.mid2, es = k.exp2s, Wl = k.Wl2;
    switch (mode) {
        case 0:
        case 3:
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = ex2_fast(fminf(0.f, fmaf(lg2_fast(u[i] + 1e-5f), DYN_LN2, -T) * slope));
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = fmaf(lg2_fast(u[i] + 1e-5f), DYN_LN2, -T), e = d + W;
                const float m = mid * e * e;
                u[i] = ex2_fast(d > W ? d * slope : (d < -W ? 0.f : m));
            }
            break;
        case 4:
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = fmaf(lg2_fast(u[i] + 1e-5f), DYN_LN2, -T), e = d - W;
                const float m = mid * e * e;
                u[i] = ex2_fast(d < -W ? d * slope : (d > W ? 0.f : m));
            }
            break;
        default: {  // exponential knees: scale * softplus(+-W d), softplus(v) = ln2 * lg2(1 + ex2(v log2e))
            const float sg = mode == 2 ? Wl : -Wl;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float d = fmaf(lg2_fast(u[i] + 1e-5f), DYN_LN2, -T);
                const float v = sg * d;  // = (+-W d) log2(e)
                const float sp = v > 28.f ? v * DYN_LN2 : DYN_LN2 * lg2_fast(1.f + ex2_fast(v));
                u[i] = ex2_fast(es * sp);
            }
            break;
        }
    }
}

__device__ __forceinline__ float softplus_fast(float v) {
    // torch softplus (threshold 20); log1p(exp(v)) with the SFU exp/log
    return v > 20.f ? v : __logf(1.f + __expf(v));
}

// u[i] (smoothed energy) -> log-gain G_out - G of the knee (dynamics.py:443-489 compressor, :675-721 gate).
// `mode` = kind * 3 + knee is uniform over the launch: the switch sits outside the sample loop, the loop
// bodies are branch-free (selects).
__device__ __forceinline__ void knee_log_gain(float (&u)[32], const KneeConst& k, int mode) {
    const float T = k.T, W = k.W, lo = k.lo, hi = k.hi, slope = k.slope, mid = k.mid_scale, es = k.exp_scale;
    switch (mode) {
        case 0:  // compressor, hard:  min(G, T + d/R) - G = min(0, d (1/R - 1))
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = fminf(0.f, (__logf(u[i] + 1e-5f) - T) * slope);
            break;
        case 1:  // compressor, quadratic: below G | above T + d/R | knee G + (1/R - 1)(d + W)^2 / (4W)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float G = __logf(u[i] + 1e-5f), d = G - T, e = d + W;
                const float m = mid * e * e;
                u[i] = G > hi ? d * slope : (G < lo ? 0.f : m);
            }
            break;
        case 2:  // compressor, exponential
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = es * softplus_fast(W * (__logf(u[i] + 1e-5f) - T));
            break;
        case 3:  // gate, hard:  min(G, R d + T) - G = min(0, (R - 1) d)
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = fminf(0.f, (__logf(u[i] + 1e-5f) - T) * slope);
            break;
        case 4:  // gate, quadratic: below R d + T | above G | knee G + (1 - R)(d - W)^2 / (4W)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float G = __logf(u[i] + 1e-5f), d = G - T, e = d - W;
                const float m = mid * e * e;
                u[i] = G < lo ? d * slope : (G > hi ? 0.f : m);
            }
            break;
        default:  // gate, exponential
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = es * softplus_fast(-W * (__logf(u[i] + 1e-5f) - T));
            break;
    }
}

// one thread per (row, stage): every constant of the sample loops
// Warm-up length of an attack / release follower.  Both branches of y' = (1 - c) y + c u are affine in y with slope
// 1 - c in (0, 1), and the branch taken only makes the result their min or max: the step is a contraction with factor
// 1 - min(at, rt) WHATEVER the input, so two runs from different states differ by at most (1 - cmin)^W times their
// initial distance after W samples: 2^-24 of it (below fp32 resolution) once W >= ln(2^-24) / ln(1 - cmin).
__device__ __forceinline__ int ballistics_warmup(float at, float rt) {
    const float cmin = fminf(at, rt);
    if (!(cmin > 1e-6f)) return 0x3fffffff;
    const float w = ceilf(-16.635532f / log1pf(-fminf(cmin, 0.999f)));
    return w > 1e9f ? 0x3fffffff : (int)w + 1;
}

__global__ void dynamics_tables_kernel(const DynParams p, float* __restrict__ tables) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.batch * p.n_stages) return;
    const int row = idx / p.n_stages, d = idx - row * p.n_stages;
    const int prow = row / p.param_rep;  // row of the parameter tensors
    if (d == 0 && p.lookback) {
        // total warm-up of the chain: the followers of the stages settle one after the other
        long long lb = 0;
        for (int dd = 0; dd < p.n_stages; ++dd) {
            for (int which = 0; which < 2; ++which) {
                const SmootherDesc& sm = which ? p.st[dd].post : p.st[dd].pre;
                if (sm.kind == 2)
                    lb += ballistics_warmup(1.f / (1.f + expf(-sm.z[(size_t)prow * 2 + 0])), 1.f / (1.f + expf(-sm.z[(size_t)prow * 2 + 1])));
            }
        }
        // worth it while the redundant warm-up stays below ~16 chunks (a single slow row is still far ahead of a walk)
        p.lookback[row] = (p.spec_ch > 0 && lb <= 16LL * p.spec_ch) ? (int)lb : -1;
    }
    const StageDesc& sd = p.st[d];
    float* base = tables + (size_t)row * DYN_ROW_FLOATS;
    for (int which = 0; which < 2; ++which) {
        const SmootherDesc& sm = which ? sd.post : sd.pre;
        SmootherConst c;
        for (int i = 0; i < DYN_SC_FLOATS; ++i) reinterpret_cast<float*>(&c)[i] = 0.f;
        if (sm.kind == 1) {
            // TruncatedOnePoleIIRFilter (core/envelope.py:44-49): alpha = min(sigmoid(z), 1 - 1e-5)
            float alpha = 1.f / (1.f + expf(-sm.z[prow]));
            alpha = fminf(alpha, 1.f - 1e-5f);
            const double ad = (double)alpha;
            c.alpha = alpha;
            c.oma = 1.f - alpha;
            if ((long long)p.iir_len < p.L) {
                const double v = exp((double)p.iir_len * log(ad));
                c.aN = v < 1e-37 ? 0.f : (float)v;
            }
            double sq = ad;
            for (int k = 0; k < 5; ++k) sq *= sq;  // alpha^32
            double run = 1.0;
            for (int l = 0; l < 32; ++l) { c.pl[l] = (float)run; run *= sq; }
            for (int j = 0; j < 5; ++j) { c.apow[j] = (float)sq; sq *= sq; }
            c.aW = (float)sq;
        } else if (sm.kind == 2) {
            // Ballistics (core/envelope.py:84-101)
            c.at = 1.f / (1.f + expf(-sm.z[(size_t)prow * 2 + 0]));
            c.rt = 1.f / (1.f + expf(-sm.z[(size_t)prow * 2 + 1]));
        }
        float* dst = base + (2 * d + which) * DYN_SC_FLOATS;
        for (int i = 0; i < DYN_SC_FLOATS; ++i) dst[i] = reinterpret_cast<const float*>(&c)[i];
    }
    KneeConst k;
    for (int i = 0; i < DYN_KC_FLOATS; ++i) reinterpret_cast<float*>(&k)[i] = 0.f;
    const float T = sd.log_threshold[prow] - 6.f, lr = sd.log_ratio[prow], lk = sd.log_knee ? sd.log_knee[prow] : 0.f;
    // ApproxNoiseGate.compute_gain (dynamics.py:186-204): R = exp(log_ratio) (no +1) and the knee term is
    // (1 - R)(d - W)^2 / (2 (2W + 1e-3)); the three regions are those of the quadratic gate knee
    const bool approx_gate = sd.knee == 3;
    const float elr = expf(lr), ratio = approx_gate ? elr : 1.f + elr, inv_ratio = 1.f / ratio;
    k.T = T;
    k.W = (sd.knee == 1 || approx_gate) ? expf(lk) * 0.5f : (sd.knee == 2 ? expf(lk) : 0.f);
    k.lo = T - k.W; k.hi = T + k.W;
    k.slope = sd.kind == 0 ? (inv_ratio - 1.f) : (ratio - 1.f);
    k.mid_scale = (sd.kind == 0 ? (inv_ratio - 1.f) : (1.f - ratio)) / (approx_gate ? 4.f * k.W + 2e-3f : 4.f * k.W);
    k.exp_scale = (sd.kind == 0 ? (inv_ratio - 1.f) : -elr) / k.W;
    k.slope2 = k.slope * DYN_LOG2E; k.mid2 = k.mid_scale * DYN_LOG2E; k.exp2s = k.exp_scale * DYN_LOG2E;
    k.Wl2 = k.W * DYN_LOG2E;
    float* dst = base + 2 * DYN_MAX_STAGES * DYN_SC_FLOATS + d * DYN_KC_FLOATS;
    for (int i = 0; i < DYN_KC_FLOATS; ++i) dst[i] = reinterpret_cast<const float*>(&k)[i];
}

__host__ __device__ constexpr int dyn_min_ctas(int NT, bool RO) {
    return RO ? 8 : (NT >= 256 ? 3 : (768 / NT > 24 ? 24 : 768 / NT));  // scan variant: 24 warps per SM (80 registers)
}
// ENV: envelope mode (gfx_envelope_f32) -- a separate instantiation so that the processors' kernel carries none of it
template <int NT, bool RO, bool ENV>
__global__ void __launch_bounds__(NT, dyn_min_ctas(NT, RO)) dynamics_kernel(const DynParams p) {
    constexpr int S = 32, TILE = NT * S, NW = NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DynCtx<NT, RO> cx;
    cx.xs4 = reinterpret_cast<float4*>(smem_raw);
    cx.work4 = cx.xs4 + (size_t)p.C * NT * 8;
    float* consts = reinterpret_cast<float*>(cx.work4 + (RO ? (size_t)NT * 16 : 0));  // [DYN_ROW_FLOATS]
    cx.sc = reinterpret_cast<const SmootherConst*>(consts);
    cx.kc = reinterpret_cast<const KneeConst*>(consts + 2 * DYN_MAX_STAGES * DYN_SC_FLOATS);
    cx.wt = consts + DYN_ROW_FLOATS;
    cx.s_state = cx.wt + 2 * NW;
    float* xs = reinterpret_cast<float*>(cx.xs4);
    __shared__ unsigned int sh_item;
    cx.tid = threadIdx.x; cx.lane = cx.tid & 31; cx.warp = cx.tid >> 5;
    {
        // two-warp CTAs usually occupy hardware warp slots (2j, 2j+1) -> sub-partitions (0,1) or (2,3); taking warp
        // (j >> 1) & 1 as the walker puts the j-th resident CTA's walker on sub-partition {0, 2, 1, 3}[j & 3].
        // One thread decides for the CTA (the slots of its warps need not be adjacent).
        __shared__ int sh_walker;
        if (cx.tid == 0) {
            unsigned wid;
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
            sh_walker = RO ? (int)(((wid >> 2) & 1u) * 32u) : 0;
        }
        __syncthreads();
        cx.walker = sh_walker;
    }
    const int C = p.C;
    const float inv_c = 1.f / (float)C;

    for (;;) {
        __syncthreads();
        if (cx.tid == 0) sh_item = take_ticket(p.ticket, p.n_items + gridDim.x - 1);
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= p.n_items) break;
        // scan variant: item = (tile, row), one tile; ballistics variant: item = row, all its tiles in order
        const int t_first = RO ? 0 : (int)(item / (unsigned)p.batch);
        const int t_last = RO ? p.tiles - 1 : t_first;
        cx.row = RO ? (int)item : (int)(item - (unsigned)t_first * (unsigned)p.batch);
        if (RO && p.lookback != nullptr && p.lookback[cx.row] >= 0) continue;  // done by dynamics_spec_kernel
        for (int tt = t_first; tt <= t_last; ++tt) {
        if (tt > t_first) __syncthreads();  // the previous tile has left shared memory
        cx.t_idx = tt;
        cx.t0 = (long long)cx.t_idx * TILE;
        cx.remain = p.L - cx.t0;
        cx.sync_parity = 0;
        cx.have_state = false;
        const float* xrow = p.x + (size_t)cx.row * C * (size_t)p.L;
        float* yrow = p.y + (size_t)cx.row * C * (size_t)p.L;

        // ---- stage the row constants and all channels of the tile
        {
            const float4* trow = reinterpret_cast<const float4*>(p.tables + (size_t)cx.row * DYN_ROW_FLOATS);
            const int n4 = p.n_stages * 2 * DYN_SC_FLOATS / 4;
            for (int i = cx.tid; i < n4; i += NT) cp_async16(consts + 4 * i, trow + i, 16);
            constexpr int k4 = DYN_MAX_STAGES * 2 * DYN_SC_FLOATS / 4;
            for (int i = cx.tid; i < p.n_stages * DYN_KC_FLOATS / 4; i += NT) cp_async16(consts + 4 * (k4 + i), trow + k4 + i, 16);
        }
        const bool full_tile = p.aligned && cx.remain >= TILE;
        if (full_tile) {
            // interior tile: unit g = tid + j NT sits at swz_unit(g >> 3, g & 7) = swz_unit(tid >> 3, tid & 7) + j NT
            // (NT is a multiple of 64), so the copies are a constant stride apart
            const uint32_t d0 = smem_u32(cx.xs4 + swz_unit(cx.tid >> 3, cx.tid & 7));
            const uint32_t d1 = smem_u32(cx.xs4 + ((cx.tid >> 3) << 3 | (((cx.tid & 7) ^ ((cx.tid >> 3) & 7)) ^ 4)));  // odd j at NT = 32
            for (int c = 0; c < C; ++c) {
                const float4* src = reinterpret_cast<const float4*>(xrow + (size_t)c * p.L + cx.t0) + cx.tid;
                const uint32_t dc = d0 + (uint32_t)c * (NT * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // (NT = 32: the row parity, hence the swizzle, alternates with j: unit ^ 4 on odd j)
                    const uint32_t dj = (NT % 64 == 0 || (j & 1) == 0) ? dc + (uint32_t)j * (NT * 16)
                                                                        : d1 + (uint32_t)c * (NT * 128) + (uint32_t)j * (NT * 16);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dj), "l"(src + j * NT));
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
        } else if (p.aligned) {
            for (int c = 0; c < C; ++c) {
                const float* xr = xrow + (size_t)c * p.L;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = cx.tid + j * NT;
                    const long long pos = (long long)g * 4;
                    const long long nb = (cx.remain - pos) * 4;
                    const int src_bytes = nb >= 16 ? 16 : (nb > 0 ? (int)nb : 0);
                    const float* src = src_bytes > 0 ? xr + cx.t0 + pos : xr;
                    cp_async16(&cx.xs4[(size_t)c * NT * 8 + swz_unit(g >> 3, g & 7)], src, src_bytes);
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            cp_async_commit();
            for (int c = 0; c < C; ++c) {
                const float* xr = xrow + (size_t)c * p.L;
                for (int i = cx.tid; i < TILE; i += NT) {
                    const float val = (i < cx.remain) ? xr[cx.t0 + i] : 0.f;
                    xs[(size_t)c * TILE + (size_t)swz_unit(i >> 5, (i & 31) >> 2) * 4 + (i & 3)] = val;
                }
            }
            cp_async_wait<0>();
        }
        __syncthreads();

        float u[S];  // working chunk: energy -> smoothed energy -> log-gain -> gain
        for (int d = 0; d < p.n_stages; ++d) {
            const StageDesc& sd = p.st[d];
            // energy of the signal entering this stage (earlier stages already applied their gain in place)
#pragma unroll
            for (int i = 0; i < S; ++i) u[i] = 0.f;
            if (ENV && p.detect != 0) {
                for (int c = 0; c < C; ++c) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 v = cx.xs4[(size_t)c * NT * 8 + swz_unit(cx.tid, q)];
                        if (p.detect == 1) {
                            u[4 * q] += fabsf(v.x); u[4 * q + 1] += fabsf(v.y); u[4 * q + 2] += fabsf(v.z); u[4 * q + 3] += fabsf(v.w);
                        } else {
                            u[4 * q] = v.x; u[4 * q + 1] = v.y; u[4 * q + 2] = v.z; u[4 * q + 3] = v.w;
                        }
                    }
                }
            } else
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = cx.xs4[(size_t)c * NT * 8 + swz_unit(cx.tid, q)];
                    u[4 * q] = fmaf(v.x, v.x, u[4 * q]);
                    u[4 * q + 1] = fmaf(v.y, v.y, u[4 * q + 1]);
                    u[4 * q + 2] = fmaf(v.z, v.z, u[4 * q + 2]);
                    u[4 * q + 3] = fmaf(v.w, v.w, u[4 * q + 3]);
                }
            }
            if (C > 1) {
#pragma unroll
                for (int i = 0; i < S; ++i) u[i] *= inv_c;
            }

            if (sd.pre.kind == 1) {
                if (d == 0) {
                    // lagged energy re-derived from x itself (read-only input): no history buffer
                    auto lag = [&](long long pos) {
                        float e = 0.f;
                        const int det = ENV ? p.detect : 0;
                        for (int c = 0; c < C; ++c) {
                            const float v = xrow[(size_t)c * p.L + pos];
                            e = det == 0 ? fmaf(v, v, e) : (det == 1 ? e + fabsf(v) : v);
                        }
                        return C > 1 ? e * inv_c : e;
                    };
                    SmootherDesc sm = sd.pre;
                    sm.hist = nullptr;
                    smooth_iir<NT, RO>(cx, p, u, sm, 2 * d, lag);
                } else {
                    const float* hr = sd.pre.hist + (size_t)cx.row * (size_t)p.L;
                    auto lag = [&](long long pos) { return __ldcg(hr + pos); };
                    smooth_iir<NT, RO>(cx, p, u, sd.pre, 2 * d, lag);
                }
            } else if (sd.pre.kind == 2) {
                if constexpr (RO) smooth_ballistics<NT, RO>(cx, p, u, sd.pre, 2 * d);  // (ballistics launches use NT = 64)
            }

            if constexpr (ENV) {
                // stand-alone smoother / envelope follower: the (log of the) smoothed detector signal is the output
                if (p.env_log) {
#pragma unroll
                    for (int i = 0; i < S; ++i) u[i] = logf(u[i] + 1e-5f);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    cx.xs4[swz_unit(cx.tid, q)] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
            } else {
            const int knee_mode = sd.kind * 3 + (sd.knee == 3 ? 1 : sd.knee);  // 3: quadratic regions, own constants
            if (sd.post.kind == 0) {
                knee_gain(u, cx.kc[d], knee_mode);
            } else {
                knee_log_gain(u, cx.kc[d], knee_mode);
                if (!sd.log_domain) {
#pragma unroll
                    for (int i = 0; i < S; ++i) u[i] = __expf(u[i]);
                }
                if (sd.post.kind == 1) {
                    const float* hr = sd.post.hist ? sd.post.hist + (size_t)cx.row * (size_t)p.L : nullptr;
                    auto lag = [&](long long pos) { return __ldcg(hr + pos); };
                    smooth_iir<NT, RO>(cx, p, u, sd.post, 2 * d + 1, lag);
                } else {
                    if constexpr (RO) smooth_ballistics<NT, RO>(cx, p, u, sd.post, 2 * d + 1);
                }
                if (sd.log_domain) {
#pragma unroll
                    for (int i = 0; i < S; ++i) u[i] = __expf(u[i]);
                }
            }
            // apply the gain in place in shared memory (each thread owns its chunk: no barrier between stages)
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4* pv = &cx.xs4[(size_t)c * NT * 8 + swz_unit(cx.tid, q)];
                    float4 v = *pv;
                    v.x *= u[4 * q]; v.y *= u[4 * q + 1]; v.z *= u[4 * q + 2]; v.w *= u[4 * q + 3];
                    *pv = v;
                }
            }
            }  // !ENV
        }
        if constexpr (!RO) {
            if (cx.tid == NT - 1 && cx.t_idx + 1 < p.tiles && cx.have_state) chain_publish(p.flags + cx.row, cx.t_idx + 1);
        }

        // ---- store coalesced
        __syncthreads();
        const int C_out = ENV ? 1 : C;
        if (ENV) yrow = p.y + (size_t)cx.row * (size_t)p.L;
        for (int c = 0; c < C_out; ++c) {
            float* yr = yrow + (size_t)c * p.L;
            if (full_tile) {
                float4* dst = reinterpret_cast<float4*>(yr + cx.t0) + cx.tid;
                const float4* sv = cx.xs4 + (size_t)c * NT * 8 + swz_unit(cx.tid >> 3, cx.tid & 7);
                const float4* sv1 = cx.xs4 + (size_t)c * NT * 8 + ((cx.tid >> 3) << 3 | (((cx.tid & 7) ^ ((cx.tid >> 3) & 7)) ^ 4));
#pragma unroll
                for (int j = 0; j < 8; ++j) stg_stream(dst + j * NT, (NT % 64 == 0 || (j & 1) == 0) ? sv[j * NT] : sv1[j * NT]);
            } else if (p.aligned) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = cx.tid + j * NT;
                    const long long pos = (long long)g * 4;
                    if (pos + 4 <= cx.remain) {
                        stg_stream(reinterpret_cast<float4*>(yr + cx.t0 + pos),
                                   cx.xs4[(size_t)c * NT * 8 + swz_unit(g >> 3, g & 7)]);
                    } else if (pos < cx.remain) {
                        const float* sv = xs + (size_t)c * TILE + (size_t)swz_unit(g >> 3, g & 7) * 4;
                        for (int e = 0; e < 4 && pos + e < cx.remain; ++e) yr[cx.t0 + pos + e] = sv[e];
                    }
                }
            } else {
                for (int i = cx.tid; i < TILE && i < cx.remain; i += NT)
                    yr[cx.t0 + i] = xs[(size_t)c * TILE + (size_t)swz_unit(i >> 5, (i & 31) >> 2) * 4 + (i & 3)];
            }
        }
        }  // tiles of the item
    }
}

// ------------------------------------------------------------------------------------------------
// Attack / release ballistics without the serial walk.  The recursion is not associative, but it is a contraction (see
// ballistics_warmup): a thread that starts `lookback` samples before its chunk from ANY state holds the true state of
// every follower (to fp32 resolution) when it reaches the chunk.  So one thread = one chunk of spec_ch samples + its
// warm-up, the whole chain of stages evaluated sample by sample in registers, no shared memory, no barrier, no hand-over
// between tiles or CTAs: rows x L / spec_ch independent threads instead of `rows` walkers.  The warm-up is redundant
// work (lookback / spec_ch of it, ~0.5-1x for coefficients sigmoid(N(0,1)) at 256-sample chunks); rows whose followers
// are too slow for that (lookback > 16 chunks, i.e. time constants beyond a few hundred samples) keep the row-owner
// walk (dynamics_kernel<64, true>), launched right after and skipping the rows done here.
// Stages with a one-pole ("iir") smoother are not handled here (launch-level decision on the host).
__device__ __forceinline__ float ball_step(float y, float u, float at, float rt, float omat, float omrt, bool use_min) {
    const float ya = fmaf(omat, y, at * u), yr = fmaf(omrt, y, rt * u);
    return use_min ? fminf(ya, yr) : fmaxf(ya, yr);
}
// scalar forms of knee_gain / knee_log_gain (same arithmetic)
__device__ __forceinline__ float knee_gain1(float e, const KneeConst& k, int mode) {
    const float d = fmaf(lg2_fast(e + 1e-5f), DYN_LN2, -k.T);
    switch (mode) {
        case 0:
        case 3: return ex2_fast(fminf(0.f, d * k.slope2));
        case 1: { const float q = d + k.W; return ex2_fast(d > k.W ? d * k.slope2 : (d < -k.W ? 0.f : k.mid2 * q * q)); }
        case 4: { const float q = d - k.W; return ex2_fast(d < -k.W ? d * k.slope2 : (d > k.W ? 0.f : k.mid2 * q * q)); }
        default: {
            const float v = (mode == 2 ? k.Wl2 : -k.Wl2) * d;
            const float sp = v > 28.f ? v * DYN_LN2 : DYN_LN2 * lg2_fast(1.f + ex2_fast(v));
            return ex2_fast(k.exp2s * sp);
        }
    }
}
__device__ __forceinline__ float knee_log_gain1(float e, const KneeConst& k, int mode) {
    const float G = __logf(e + 1e-5f), d = G - k.T;
    switch (mode) {
        case 0:
        case 3: return fminf(0.f, d * k.slope);
        case 1: { const float q = d + k.W; return G > k.hi ? d * k.slope : (G < k.lo ? 0.f : k.mid_scale * q * q); }
        case 4: { const float q = d - k.W; return G < k.lo ? d * k.slope : (G > k.hi ? 0.f : k.mid_scale * q * q); }
        case 2: return k.exp_scale * softplus_fast(k.W * d);
        default: return k.exp_scale * softplus_fast(-k.W * d);
    }
}

struct SpecStage {  // per-thread registers of one stage
    float at0, rt0, at1, rt1;   // pre / post follower coefficients
    bool min0, min1;
    float s0, s1;               // follower states
};

#ifndef GFX_SPEC_NT
#define GFX_SPEC_NT 128
#endif
__host__ __device__ constexpr int SPEC_NT(int) { return GFX_SPEC_NT; }
template <int C, int NS>
__global__ void __launch_bounds__(SPEC_NT(C)) dynamics_spec_kernel(const DynParams p) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = (int)((p.L + p.spec_ch - 1) / p.spec_ch);
    const int row = (int)(gid / chunks);
    if (row >= p.batch) return;
    const int lb = p.lookback[row];
    if (lb < 0) return;
    const long long p0 = (long long)(gid - (long long)row * chunks) * p.spec_ch;
    long long n0 = p0 - lb;
    if (n0 < 0) n0 = 0;
    long long n1 = p0 + p.spec_ch;
    if (n1 > p.L) n1 = p.L;
    const float* tab = p.tables + (size_t)row * DYN_ROW_FLOATS;
    const SmootherConst* sc = reinterpret_cast<const SmootherConst*>(tab);
    const KneeConst* kc = reinterpret_cast<const KneeConst*>(tab + 2 * DYN_MAX_STAGES * DYN_SC_FLOATS);
    SpecStage st[NS];
    KneeConst kk[NS];
    int mode[NS];
#pragma unroll
    for (int d = 0; d < NS; ++d) {
        st[d].at0 = sc[2 * d].at; st[d].rt0 = sc[2 * d].rt; st[d].at1 = sc[2 * d + 1].at; st[d].rt1 = sc[2 * d + 1].rt;
        st[d].min0 = st[d].at0 >= st[d].rt0; st[d].min1 = st[d].at1 >= st[d].rt1;
        st[d].s0 = 1.f; st[d].s1 = 1.f;  // ballistics starts from zi = 1 (exact when the thread starts at sample 0)
        kk[d] = kc[d];
        mode[d] = p.st[d].kind * 3 + (p.st[d].knee == 3 ? 1 : p.st[d].knee);
    }
    const float inv_c = 1.f / (float)C;
    auto sample = [&](float (&v)[C]) {
#pragma unroll
        for (int d = 0; d < NS; ++d) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) e = fmaf(v[c], v[c], e);
            if (C > 1) e *= inv_c;
            const StageDesc& sd = p.st[d];
            if (sd.pre.kind == 2) {
                st[d].s0 = ball_step(st[d].s0, e, st[d].at0, st[d].rt0, 1.f - st[d].at0, 1.f - st[d].rt0, st[d].min0);
                e = st[d].s0;
            }
            float g;
            if (sd.post.kind == 0) {
                g = knee_gain1(e, kk[d], mode[d]);
            } else {
                float lg = knee_log_gain1(e, kk[d], mode[d]);
                if (!sd.log_domain) lg = __expf(lg);
                st[d].s1 = ball_step(st[d].s1, lg, st[d].at1, st[d].rt1, 1.f - st[d].at1, 1.f - st[d].rt1, st[d].min1);
                g = sd.log_domain ? __expf(st[d].s1) : st[d].s1;
            }
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] *= g;
        }
    };

    // Every thread streams its own samples with 16-byte accesses (a thread's lines stay in L1 between its accesses).
    // Measured on B200 (config 4b): this form 0.45 ms; warp-cooperative coalesced tiles in shared memory 0.94 ms, the same
    // with cp.async double buffering 0.64 ms -- the per-sample shared-memory round trip costs more than the sectors save.
    const float* x0 = p.x + (size_t)row * C * (size_t)p.L;
    float* y0 = p.y + (size_t)row * C * (size_t)p.L;
    long long n = n0;
    if (p.aligned) {
        // one loop over 16-byte groups for warm-up and chunk (the warm-up may start up to 3 samples early: harmless),
        // the next group loaded while this one is walked; only groups inside the chunk are stored
        n = n0 & ~3LL;
        float4 q[C], qn[C];
#pragma unroll
        for (int c = 0; c < C; ++c) q[c] = __ldg(reinterpret_cast<const float4*>(x0 + (size_t)c * p.L + n));
#pragma unroll 1
        for (; n < n1; n += 4) {
            const long long nn = n + 4 < n1 ? n + 4 : n;
#pragma unroll
            for (int c = 0; c < C; ++c) qn[c] = __ldg(reinterpret_cast<const float4*>(x0 + (size_t)c * p.L + nn));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[C];
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = reinterpret_cast<const float*>(&q[c])[i];
                sample(v);
#pragma unroll
                for (int c = 0; c < C; ++c) reinterpret_cast<float*>(&q[c])[i] = v[c];
            }
            if (n >= p0) {
#pragma unroll
                for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(y0 + (size_t)c * p.L + n) = q[c];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) q[c] = qn[c];
        }
        return;
    }
    for (; n < n1; ++n) {
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x0 + (size_t)c * p.L + n);
        sample(v);
        if (n >= p0) {
#pragma unroll
            for (int c = 0; c < C; ++c) y0[(size_t)c * p.L + n] = v[c];
        }
    }
}

static size_t dyn_smem_bytes(int NT, int C, bool RO) {
    // two scratch tiles (at*u, rt*u) are only used by the ballistics (row owner) variant
    return (size_t)(C + (RO ? 2 : 0)) * NT * 128 +
           (size_t)(DYN_ROW_FLOATS + 2 * (NT / 32) + 4 * DYN_MAX_STAGES) * sizeof(float) + 64;
}

static size_t dyn_tables_offset(int batch, int n_stages) {
    const size_t flags = ((size_t)batch * sizeof(int) + 255) / 256 * 256;
    return (256 + flags + (size_t)batch * 2 * n_stages * sizeof(float) + 255) / 256 * 256;
}
static size_t dyn_lookback_offset(int batch, int n_stages) {
    return (dyn_tables_offset(batch, n_stages) + (size_t)batch * DYN_ROW_FLOATS * sizeof(float) + 255) / 256 * 256;
}
static size_t dyn_workspace_bytes(int batch, int n_stages) {
    return dyn_lookback_offset(batch, n_stages) + ((size_t)batch * sizeof(int) + 255) / 256 * 256;
}
static int g_dyn_spec_threads_per_sm = 1536;  // chunk size: as large as still gives this many threads per SM
static int g_dyn_ballistics_spec = 1;  // 1: warm-up chunks (dynamics_spec_kernel) where the followers allow it; 0: always walk

template <int C>
static int launch_spec_c(const DynParams& p, cudaStream_t stream) {
    const long long chunks = (p.L + p.spec_ch - 1) / p.spec_ch;
    const long long threads = (long long)p.batch * chunks;
    constexpr int NT = SPEC_NT(C);
    const unsigned grid = (unsigned)((threads + NT - 1) / NT);
    switch (p.n_stages) {
        case 1: dynamics_spec_kernel<C, 1><<<grid, NT, 0, stream>>>(p); break;
        case 2: dynamics_spec_kernel<C, 2><<<grid, NT, 0, stream>>>(p); break;
        case 3: dynamics_spec_kernel<C, 3><<<grid, NT, 0, stream>>>(p); break;
        default: dynamics_spec_kernel<C, 4><<<grid, NT, 0, stream>>>(p); break;
    }
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

template <int NT, bool RO, bool ENV = false>
static int launch_dynamics(DynParams& p, cudaStream_t stream) {
    const size_t smem = dyn_smem_bytes(NT, p.C, RO);
    if (smem > (size_t)device_info().max_smem_optin) return GFX_ERR_UNSUPPORTED;
    auto kern = dynamics_kernel<NT, RO, ENV>;
    static size_t configured_dev[64] = {0};
    size_t& configured = configured_dev[device_slot()];
    if (smem > configured) {
        GFX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int occ = 0;
    GFX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return GFX_ERR_UNSUPPORTED;
    long long grid = (long long)device_info().sm_count * occ;
    if (grid > (long long)p.n_items) grid = p.n_items;
    dynamics_tables_kernel<<<(p.batch * p.n_stages + 127) / 128, 128, 0, stream>>>(p, p.tables);
    GFX_LAUNCH_CHECK();
    if (RO && !ENV && p.lookback != nullptr) {
        const int rc = p.C == 1 ? launch_spec_c<1>(p, stream) : launch_spec_c<2>(p, stream);
        if (rc != GFX_OK) return rc;
    }
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
}

static int scan_nt_for(int batch) {
    if (g_dyn_scan_nt) return g_dyn_scan_nt;
    return batch >= device_info().sm_count * DYN_SCAN_CTAS_SMALL ? DYN_SCAN_NT_SMALL : DYN_SCAN_NT;
}
template <bool ENV>
static int launch_scan(int NT, DynParams& p, cudaStream_t st) {
    // (32- and 64-thread CTAs were measured and are 1.5-3x slower: profiles/r02_dynamics_cta_size.txt; not instantiated)
    return NT == 128 ? launch_dynamics<128, false, ENV>(p, st) : launch_dynamics<256, false, ENV>(p, st);
}

}  // namespace gfx

extern "C" {

int gfx_dynamics_set_tuning(int scan_threads) {
    if (scan_threads != 0 && scan_threads != 128 && scan_threads != 256) return GFX_ERR_INVALID;
    gfx::g_dyn_scan_nt = scan_threads;
    return GFX_OK;
}

int gfx_dynamics_set_ballistics_mode(int mode) {
    if (mode >= 64) { gfx::g_dyn_spec_threads_per_sm = mode; return GFX_OK; }  // (tuning: threads per SM the chunk size aims at)
    if (mode != 0 && mode != 1) return GFX_ERR_INVALID;
    gfx::g_dyn_ballistics_spec = mode;
    return GFX_OK;
}

size_t gfx_dynamics_workspace_bytes(int batch, int n_stages) { return gfx::dyn_workspace_bytes(batch, n_stages); }

int gfx_dynamics_f32(const float* x, float* y, int batch, int channels, long long L,
                     const gfx_dynamics_stage* stages, int n_stages, int iir_len, void* workspace,
                     size_t workspace_bytes, void* stream) {
    return gfx_dynamics_rep_f32(x, y, batch, channels, L, stages, n_stages, iir_len, 1, workspace, workspace_bytes, stream);
}

int gfx_dynamics_rep_f32(const float* x, float* y, int batch, int channels, long long L,
                         const gfx_dynamics_stage* stages, int n_stages, int iir_len, int param_repeat, void* workspace,
                         size_t workspace_bytes, void* stream) {
    using namespace gfx;
    if (!x || !y || !stages) return GFX_ERR_INVALID;
    if (batch <= 0 || channels <= 0 || L <= 0 || n_stages <= 0 || iir_len <= 0) return GFX_ERR_INVALID;
    if (param_repeat <= 0 || batch % param_repeat != 0) return GFX_ERR_INVALID;
    if (n_stages > DYN_MAX_STAGES) return GFX_ERR_UNSUPPORTED;
    DynParams p;
    p.param_rep = param_repeat;
    p.envelope = 0; p.detect = 0; p.env_log = 0; p.lookback = nullptr; p.spec_ch = 0;
    bool any_ballistics = false, any_iir = false;
    for (int d = 0; d < n_stages; ++d) {
        const gfx_dynamics_stage& s = stages[d];
        if (s.kind < 0 || s.kind > 1 || s.knee < 0 || s.knee > 3 || (s.knee == 3 && s.kind != 1)) return GFX_ERR_INVALID;
        if (s.energy_smoother < 0 || s.energy_smoother > 2 || s.gain_smoother < 0 || s.gain_smoother > 2) return GFX_ERR_INVALID;
        if (!s.log_threshold || !s.log_ratio) return GFX_ERR_INVALID;
        if (s.knee != 0 && !s.log_knee) return GFX_ERR_INVALID;
        if (s.energy_smoother && !s.z_alpha_pre) return GFX_ERR_INVALID;
        if (s.gain_smoother && !s.z_alpha_post) return GFX_ERR_INVALID;
        // history buffers are needed whenever the truncation tail could matter
        if ((long long)iir_len < L) {
            if (s.energy_smoother == 1 && d > 0 && !s.hist_pre) return GFX_ERR_WORKSPACE;
            if (s.gain_smoother == 1 && !s.hist_post) return GFX_ERR_WORKSPACE;
        }
        StageDesc& o = p.st[d];
        o.kind = s.kind; o.knee = s.knee; o.log_domain = s.gain_smooth_in_log;
        o.log_threshold = s.log_threshold; o.log_ratio = s.log_ratio; o.log_knee = s.log_knee;
        o.pre = SmootherDesc{s.energy_smoother, s.z_alpha_pre, s.hist_pre};
        o.post = SmootherDesc{s.gain_smoother, s.z_alpha_post, s.hist_post};
        any_ballistics |= (s.energy_smoother == 2 || s.gain_smoother == 2);
        any_iir |= (s.energy_smoother == 1 || s.gain_smoother == 1);
    }
    const int NT = any_ballistics ? 64 : scan_nt_for(batch);
    const long long tile = (long long)NT * 32;
    const long long tiles_ll = (L + tile - 1) / tile;
    if ((long long)batch * tiles_ll > 0x7fff0000LL) return GFX_ERR_UNSUPPORTED;
    const size_t need = dyn_workspace_bytes(batch, n_stages);
    if (!workspace || workspace_bytes < need) return GFX_ERR_WORKSPACE;
    p.x = x; p.y = y; p.batch = batch; p.C = channels; p.L = L;
    p.tiles = (int)tiles_ll;
    p.n_items = any_ballistics ? (unsigned)batch : (unsigned)((long long)batch * tiles_ll);  // rows | (tile, row) pairs
    unsigned char* w = (unsigned char*)workspace;
    const size_t flags_bytes = ((size_t)batch * sizeof(int) + 255) / 256 * 256;
    p.ticket = (unsigned int*)w;
    p.flags = (int*)(w + 256);
    p.state = (float*)(w + 256 + flags_bytes);
    p.n_stages = n_stages;
    p.iir_len = iir_len;
    p.aligned = (((uintptr_t)x | (uintptr_t)y) % 16 == 0) && (L % 4 == 0);
    GFX_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 256 + flags_bytes, (cudaStream_t)stream));
    p.tables = (float*)(w + dyn_tables_offset(batch, n_stages));
    cudaStream_t st = (cudaStream_t)stream;
    if (any_ballistics && !any_iir && channels <= 2 && g_dyn_ballistics_spec && L < 0x40000000LL) {
        // chunks with warm-up: as many independent threads as fill the chip (at least 64 samples each)
        int ch = 1024;
        while (ch > 64 && (long long)batch * L / ch < (long long)device_info().sm_count * g_dyn_spec_threads_per_sm) ch >>= 1;
        p.lookback = (int*)(w + dyn_lookback_offset(batch, n_stages));
        p.spec_ch = ch;
    }
    if (any_ballistics) return launch_dynamics<64, true>(p, st);
    return launch_scan<false>(NT, p, st);
}

int gfx_envelope_f32(const float* x, float* y, int batch, int channels, long long L, int smoother, const float* z,
                     int detect, int log_out, int iir_len, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace gfx;
    if (!x || !y || !z) return GFX_ERR_INVALID;
    if (batch <= 0 || channels <= 0 || L <= 0 || iir_len <= 0) return GFX_ERR_INVALID;
    if (smoother < 1 || smoother > 2 || detect < 0 || detect > 2 || (detect == 2 && channels != 1)) return GFX_ERR_INVALID;
    DynParams p;
    p.param_rep = 1;
    p.envelope = 1; p.detect = detect; p.env_log = log_out ? 1 : 0; p.lookback = nullptr; p.spec_ch = 0;
    StageDesc& o = p.st[0];
    o.kind = 0; o.knee = 0; o.log_domain = 0;
    o.log_threshold = z; o.log_ratio = z; o.log_knee = nullptr;  // (the knee constants are formed but never used)
    o.pre = SmootherDesc{smoother, z, nullptr};
    o.post = SmootherDesc{0, nullptr, nullptr};
    const bool ball = smoother == 2;
    const int NT = ball ? 64 : scan_nt_for(batch);
    const long long tile = (long long)NT * 32;
    const long long tiles_ll = (L + tile - 1) / tile;
    if ((long long)batch * tiles_ll > 0x7fff0000LL) return GFX_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < dyn_workspace_bytes(batch, 1)) return GFX_ERR_WORKSPACE;
    p.x = x; p.y = y; p.batch = batch; p.C = channels; p.L = L;
    p.tiles = (int)tiles_ll;
    p.n_items = ball ? (unsigned)batch : (unsigned)((long long)batch * tiles_ll);
    unsigned char* w = (unsigned char*)workspace;
    const size_t flags_bytes = ((size_t)batch * sizeof(int) + 255) / 256 * 256;
    p.ticket = (unsigned int*)w;
    p.flags = (int*)(w + 256);
    p.state = (float*)(w + 256 + flags_bytes);
    p.n_stages = 1;
    p.iir_len = iir_len;
    p.aligned = (((uintptr_t)x | (uintptr_t)y) % 16 == 0) && (L % 4 == 0);
    GFX_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 256 + flags_bytes, (cudaStream_t)stream));
    p.tables = (float*)(w + dyn_tables_offset(batch, 1));
    cudaStream_t st = (cudaStream_t)stream;
    if (ball) return launch_dynamics<64, true, true>(p, st);
    return launch_scan<true>(NT, p, st);
}

}  // extern "C"
