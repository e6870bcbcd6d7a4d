// Exact cascade of K second-order sections (biquads), time-parallel, one pass over HBM.
//
// Replaces (reference, /root/reference/src/grafx):
//   processors/core/iir.py:154-184  IIRFilter._process_lfilter  (K x torchaudio lfilter: a conv1d
//       launch + `iir_cu_kernel`, one thread per (batch,channel) walking all L samples)
//   processors/core/iir.py:186-261  IIRFilter._process_ssm      (same transfer function)
// Semantics per section (coefficients normalised by that section's a0, as torchaudio does):
//   v[n] = b0 u[n] + b1 u[n-1] + b2 u[n-2];   y[n] = v[n] - a1 y[n-1] - a2 y[n-2];   zero initial state.
//
// Algorithm.  A row is cut into tiles of NT*S samples; a tile is staged in shared memory
// (cp.async, XOR-swizzled 128-byte rows, double buffered: the next tile streams in while this one
// is computed) so that thread t owns the S contiguous samples [t*S, (t+1)*S) in registers.
// Per section:
//   1. feed-forward part in place (needs the two samples before the chunk: the neighbour's state);
//   2. zero-state recursion over the chunk -> end state z_t (two values);
//   3. the true state at each chunk start is s_{t+1} = M s_t + z_t with M = A^S, A = [[-a1,-a2],[1,0]]:
//      warp-shuffle scan with the matrices M^(2^j), warps stitched through shared memory, the tile's
//      incoming state taken from the previous tile of the row (ordered chain, see common.cuh);
//   4. the recursion is re-run from the true state -- this pass is arithmetically the sequential
//      DF-I loop, so rounding differs from the reference only through the carried state.
// The normalised coefficients and the powers M^l (l = 0..32) of every (coefficient row, section)
// are produced once per call in double precision by a small prologue kernel (36 x 4 words each)
// and travel to shared memory with the tile.
// HBM traffic: x read once, y written once (8 B/sample fp32) + 2K words of state per tile.
#include <cstdlib>

#include "common.cuh"

namespace gfx {

constexpr int TAB_ENTRIES = 36;  // per (coef row, section): 33 matrices, [33] = (b0,b1,b2,-a1), [34] = (-a2,..)

template <typename T>
struct CascadeParams {
    const T* x;
    T* y;
    const T* tables;  // [coef rows][K][36][4]
    int batch, c_sig, c_filt, c_out, K;
    long long L;
    int rows, tiles;
    unsigned int n_items;
    unsigned int* ticket;
    int* flags;  // [rows]   number of finished tiles of the row
    T* state;    // [rows][2K] (y_k[-1], y_k[-2]) left by the last finished tile
    unsigned long long* xstate;  // packed-kernel variant: [rows][K][2] {value, tag} words
    const double* xtables;       // packed-kernel variant: [coef rows][K][X2_TAB][4] doubles
    int aligned;  // 1: x/y rows are 16-byte aligned (vector path)
    // packed-kernel variant, render_grafx's first render order (gfx_biquad_cascade_ex_f32 with xcopy): the signal is read from
    // the caller's [src_outer, src_inner, c_sig, L] sources (item b of this launch = source item
    // (b % src_outer) * src_inner + b / src_outer: the node-major order of the signal buffer) and every staged tile is
    // also stored to xcopy [rows, L] -- the source slice of the signal buffer -- so no separate copy pass reads it again
    T* xcopy;
    int src_outer, src_inner;
};

__device__ __forceinline__ void mat2_mul(const double a[4], const double b[4], double c[4]) {
    c[0] = a[0] * b[0] + a[1] * b[2];
    c[1] = a[0] * b[1] + a[1] * b[3];
    c[2] = a[2] * b[0] + a[3] * b[2];
    c[3] = a[2] * b[1] + a[3] * b[3];
}

// ---- prologue: one warp per (coefficient row, section)
template <typename T>
__global__ void __launch_bounds__(128) cascade_tables_kernel(const T* __restrict__ Bs, const T* __restrict__ As,
                                                             T* __restrict__ tables, int n_sections) {
    constexpr int S = 128 / (int)sizeof(T);
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_sections) return;
    const T* bp = Bs + (size_t)w * 3;
    const T* ap = As + (size_t)w * 3;
    const T a0 = ap[0];
    const T nb0 = bp[0] / a0, nb1 = bp[1] / a0, nb2 = bp[2] / a0;
    const T na1 = ap[1] / a0, na2 = ap[2] / a0;
    double base[4] = {-(double)na1, -(double)na2, 1.0, 0.0}, tmp[4];
#pragma unroll 1
    for (int sq = 1; sq < S; sq <<= 1) {  // M = A^S
        mat2_mul(base, base, tmp);
        base[0] = tmp[0]; base[1] = tmp[1]; base[2] = tmp[2]; base[3] = tmp[3];
    }
    double res[4] = {1.0, 0.0, 0.0, 1.0};  // lane l -> M^l
#pragma unroll 1
    for (int bit = 0; bit < 5; ++bit) {
        if ((lane >> bit) & 1) {
            mat2_mul(res, base, tmp);
            res[0] = tmp[0]; res[1] = tmp[1]; res[2] = tmp[2]; res[3] = tmp[3];
        }
        mat2_mul(base, base, tmp);
        base[0] = tmp[0]; base[1] = tmp[1]; base[2] = tmp[2]; base[3] = tmp[3];
    }
    T* out = tables + (size_t)w * TAB_ENTRIES * 4;
    out[lane * 4 + 0] = (T)res[0]; out[lane * 4 + 1] = (T)res[1];
    out[lane * 4 + 2] = (T)res[2]; out[lane * 4 + 3] = (T)res[3];
    if (lane == 0) {
        out[32 * 4 + 0] = (T)base[0]; out[32 * 4 + 1] = (T)base[1];
        out[32 * 4 + 2] = (T)base[2]; out[32 * 4 + 3] = (T)base[3];
        out[33 * 4 + 0] = nb0; out[33 * 4 + 1] = nb1; out[33 * 4 + 2] = nb2; out[33 * 4 + 3] = -na1;
        out[34 * 4 + 0] = -na2; out[34 * 4 + 1] = T(0); out[34 * 4 + 2] = T(0); out[34 * 4 + 3] = T(0);
        mat2_mul(base, base, tmp);  // M^64
        out[35 * 4 + 0] = (T)tmp[0]; out[35 * 4 + 1] = (T)tmp[1]; out[35 * 4 + 2] = (T)tmp[2]; out[35 * 4 + 3] = (T)tmp[3];
    }
}

// ---- tables of the packed fp32 kernel, in DOUBLE: the carried state of a low-frequency section (poles within
// ~1e-2 of z = 1: every bass EQ band) is a difference of terms hundreds of times larger than itself, and the
// chunk matrices M^l have near-parallel eigenvectors -- rounded to fp32 they describe a different (even unstable)
// filter.  So: state propagation in double with exact (double) powers of the matrix the fp32 coefficients define;
// the per-sample recursions stay fp32 on those same coefficients (arithmetically the reference's loop).
// Per (coefficient row, section), X2_TAB entries of 4 doubles:
//   [l] = M^l, l = 0..32 (M = A^32);  [33] = N: end-state response to the two inputs before a chunk;
//   [34] = (b0, b1, b2, -a1);  [35] = M^64;  [36] = (-a2, ill, 0, 0)        (coefficients normalised in fp32)
//   [37..56] = the same matrices rounded to fp32 (float4 l = M^l, [33] = N, [34] = M^64).  ill = 1 when some entry of
//   M^l, l <= 32, exceeds X2_ILL_GROWTH in magnitude (poles below ~2 kHz at 48 kHz); well-conditioned sections, where
//   fp32 carries were measured to match the sequential loop, skip the double arithmetic.
constexpr int X2_TAB = 57;        // 37 double entries + 20 (= 40 float4: M^0..32, N, M^64 as fp32, for well-conditioned sections)
constexpr double X2_ILL_GROWTH = 4.0;  // a section whose chunk-matrix powers exceed this is propagated in double

// coef_rep > 1: runs of coef_rep consecutive batch items share one coefficient row of Bs / As (render_grafx's 4-D sources);
// the table of every (item, channel, section) is written all the same, so the sample kernel is unaware of it
__global__ void __launch_bounds__(128) cascade_x2_tables_kernel(const float* __restrict__ Bs, const float* __restrict__ As,
                                                                double* __restrict__ tables, int n_sections, int K, int c_filt,
                                                                int coef_rep) {
    constexpr int S = 32;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_sections) return;
    size_t src = (size_t)w;
    if (coef_rep > 1) {
        const int crow = w / K, k = w - crow * K;
        const int b = crow / c_filt, cf = crow - b * c_filt;
        src = ((size_t)(b / coef_rep) * c_filt + cf) * K + k;
    }
    const float* bp = Bs + src * 3;
    const float* ap = As + src * 3;
    const float a0 = ap[0];
    const float nb0 = bp[0] / a0, nb1 = bp[1] / a0, nb2 = bp[2] / a0;
    const float na1 = ap[1] / a0, na2 = ap[2] / a0;
    const double A[4] = {-(double)na1, -(double)na2, 1.0, 0.0};
    double base[4] = {A[0], A[1], A[2], A[3]}, tmp[4];
#pragma unroll 1
    for (int sq = 1; sq < S; sq <<= 1) {  // M = A^S
        mat2_mul(base, base, tmp);
        base[0] = tmp[0]; base[1] = tmp[1]; base[2] = tmp[2]; base[3] = tmp[3];
    }
    double res[4] = {1.0, 0.0, 0.0, 1.0};  // lane l -> M^l
#pragma unroll 1
    for (int bit = 0; bit < 5; ++bit) {
        if ((lane >> bit) & 1) {
            mat2_mul(res, base, tmp);
            res[0] = tmp[0]; res[1] = tmp[1]; res[2] = tmp[2]; res[3] = tmp[3];
        }
        mat2_mul(base, base, tmp);
        base[0] = tmp[0]; base[1] = tmp[1]; base[2] = tmp[2]; base[3] = tmp[3];
    }
    double* out = tables + (size_t)w * X2_TAB * 4;
    float4* outf = reinterpret_cast<float4*>(out + 37 * 4);
    out[lane * 4 + 0] = res[0]; out[lane * 4 + 1] = res[1]; out[lane * 4 + 2] = res[2]; out[lane * 4 + 3] = res[3];
    outf[lane] = make_float4((float)res[0], (float)res[1], (float)res[2], (float)res[3]);
    double growth = fmax(fmax(fabs(res[0]), fabs(res[1])), fmax(fabs(res[2]), fabs(res[3])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) growth = fmax(growth, __shfl_xor_sync(0xffffffffu, growth, o));
    if (lane == 0) {
        growth = fmax(growth, fmax(fmax(fabs(base[0]), fabs(base[1])), fmax(fabs(base[2]), fabs(base[3]))));
        out[32 * 4 + 0] = base[0]; out[32 * 4 + 1] = base[1]; out[32 * 4 + 2] = base[2]; out[32 * 4 + 3] = base[3];
        outf[32] = make_float4((float)base[0], (float)base[1], (float)base[2], (float)base[3]);
        mat2_mul(base, base, tmp);  // M^64
        out[35 * 4 + 0] = tmp[0]; out[35 * 4 + 1] = tmp[1]; out[35 * 4 + 2] = tmp[2]; out[35 * 4 + 3] = tmp[3];
        outf[34] = make_float4((float)tmp[0], (float)tmp[1], (float)tmp[2], (float)tmp[3]);
        // impulse response of 1/A(z): h[n] = (A^n)[0][0]; the inputs u[-1], u[-2] reach the recursion as
        // d[0] = b1 u[-1] + b2 u[-2], d[1] = b2 u[-1]  ->  end state (y[S-1], y[S-2]) += N (u[-1], u[-2])
        double pw[4] = {1.0, 0.0, 0.0, 1.0}, h29 = 0.0, h30 = 0.0, h31 = 0.0;
        for (int n = 0; n < S; ++n) {
            if (n == S - 3) h29 = pw[0];
            if (n == S - 2) h30 = pw[0];
            if (n == S - 1) h31 = pw[0];
            mat2_mul(pw, A, tmp);
            pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; pw[3] = tmp[3];
        }
        const double b1 = (double)nb1, b2 = (double)nb2;
        out[33 * 4 + 0] = h31 * b1 + h30 * b2; out[33 * 4 + 1] = h31 * b2;
        out[33 * 4 + 2] = h30 * b1 + h29 * b2; out[33 * 4 + 3] = h30 * b2;
        out[34 * 4 + 0] = (double)nb0; out[34 * 4 + 1] = (double)nb1; out[34 * 4 + 2] = (double)nb2; out[34 * 4 + 3] = -(double)na1;
        outf[33] = make_float4((float)out[33 * 4 + 0], (float)out[33 * 4 + 1], (float)out[33 * 4 + 2], (float)out[33 * 4 + 3]);
        out[36 * 4 + 0] = -(double)na2; out[36 * 4 + 1] = growth > X2_ILL_GROWTH ? 1.0 : 0.0; out[36 * 4 + 2] = 0.0; out[36 * 4 + 3] = 0.0;
        // fp32 copies of the five coefficients + the flag, and a zero matrix (the scan identity for lanes below the stride)
        outf[35] = make_float4(nb0, nb1, nb2, -na1);
        outf[36] = make_float4(-na2, growth > X2_ILL_GROWTH ? 1.f : 0.f, 0.f, 0.f);
        outf[37] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <typename T>
struct Mat2 {
    T m00, m01, m10, m11;
};

template <typename T, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) biquad_cascade_kernel(const CascadeParams<T> p) {
    constexpr int S = 128 / (int)sizeof(T);   // samples per thread (32 fp32 / 16 fp64)
    constexpr int PER = 16 / (int)sizeof(T);  // samples per 16-byte unit
    constexpr int TILE = NT * S;
    constexpr int NW = NT / 32;
    constexpr int TAB_UNITS = TAB_ENTRIES * 4 * (int)sizeof(T) / 16;  // 16-byte units per (row, section)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = p.K;
    const size_t tab_bytes = (size_t)K * TAB_UNITS * 16;
    const size_t stage_bytes = (size_t)NT * 128 + tab_bytes + 16;  // tile | tables | 2 history samples
    T* s_in = reinterpret_cast<T*>(smem_raw + 2 * stage_bytes);   // [K][2]
    T* wtot = s_in + (size_t)K * 2;                                 // [2][NW][2]
    __shared__ unsigned int sh_item[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // issues the asynchronous loads of one work item into stage `st`
    auto prefetch = [&](unsigned int item, int st) {
        unsigned char* base = smem_raw + (size_t)st * stage_bytes;
        uint4* tile4 = reinterpret_cast<uint4*>(base);
        const int t_idx = (int)(item / (unsigned)p.rows);
        const int row = (int)(item - (unsigned)t_idx * (unsigned)p.rows);
        const int b = row / p.c_out, c = row - b * p.c_out;
        const T* xr = p.x + ((size_t)b * p.c_sig + (p.c_sig == 1 ? 0 : c)) * (size_t)p.L;
        const size_t crow = (size_t)b * p.c_filt + (p.c_filt == 1 ? 0 : c);
        const long long t0 = (long long)t_idx * TILE;
        const long long remain = p.L - t0;
        if (p.aligned) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = tid + j * NT;
                const long long pos = (long long)g * PER;
                long long nb = (remain - pos) * (long long)sizeof(T);
                const int src_bytes = nb >= 16 ? 16 : (nb > 0 ? (int)nb : 0);
                const T* src = src_bytes > 0 ? xr + t0 + pos : xr;
                cp_async16(&tile4[swz_unit(g >> 3, g & 7)], src, src_bytes);
            }
        } else {
            T* tile = reinterpret_cast<T*>(base);
            for (int i = tid; i < TILE; i += NT) {
                const T val = (i < remain) ? xr[t0 + i] : T(0);
                const int r = i / S, n = i - r * S;
                tile[(size_t)swz_unit(r, n / PER) * PER + (n % PER)] = val;
            }
        }
        const unsigned char* tsrc = reinterpret_cast<const unsigned char*>(p.tables + crow * K * TAB_ENTRIES * 4);
        unsigned char* tdst = base + (size_t)NT * 128;
        for (int u = tid; u < K * TAB_UNITS; u += NT) cp_async16(tdst + (size_t)u * 16, tsrc + (size_t)u * 16, 16);
        if (tid < 2) {
            // the two input samples before the tile (needed by thread 0 only)
            T* hist = reinterpret_cast<T*>(base + (size_t)NT * 128 + tab_bytes);
            if (t_idx > 0) cp_async_small<(int)sizeof(T)>(hist + tid, xr + t0 - 1 - tid);
            else hist[tid] = T(0);
        }
    };

    if (tid == 0) sh_item[0] = take_ticket(p.ticket, 0xffffffffu);
    __syncthreads();
    unsigned int item = sh_item[0];
    if (item < p.n_items) prefetch(item, 0);
    cp_async_commit();
    int st = 0;

    while (item < p.n_items) {
        // next ticket + its loads go out before this tile is touched
        if (tid == 0) sh_item[(st ^ 1) & 1] = take_ticket(p.ticket, 0xffffffffu);
        __syncthreads();  // (also: every thread is done with stage st^1 of the previous iteration)
        const unsigned int next_item = sh_item[(st ^ 1) & 1];
        if (next_item < p.n_items) prefetch(next_item, st ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        unsigned char* base = smem_raw + (size_t)st * stage_bytes;
        uint4* tile4 = reinterpret_cast<uint4*>(base);
        T* tile = reinterpret_cast<T*>(base);
        const Mat2<T>* tab = reinterpret_cast<const Mat2<T>*>(base + (size_t)NT * 128);
        const T* hist = reinterpret_cast<const T*>(base + (size_t)NT * 128 + tab_bytes);

        const int t_idx = (int)(item / (unsigned)p.rows);
        const int row = (int)(item - (unsigned)t_idx * (unsigned)p.rows);
        T* yr = p.y + (size_t)row * (size_t)p.L;
        const long long t0 = (long long)t_idx * TILE;
        const long long remain = p.L - t0;  // > 0

        // ---- own chunk -> registers
        T v[S];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint4 q = tile4[swz_unit(tid, u)];
            const T* qs = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int e = 0; e < PER; ++e) v[u * PER + e] = qs[e];
        }
        // the two input samples before the chunk
        T um1, um2;
        if (tid > 0) {
            const size_t base_idx = (size_t)swz_unit(tid - 1, 7) * PER;
            um1 = tile[base_idx + PER - 1];
            um2 = tile[base_idx + PER - 2];
        } else {
            um1 = hist[0];
            um2 = hist[1];
        }

        for (int k = 0; k < K; ++k) {
            const Mat2<T>* pk = tab + (size_t)k * TAB_ENTRIES;
            const Mat2<T> c0 = pk[33];
            const T b0 = c0.m00, b1 = c0.m01, b2 = c0.m10, na1 = c0.m11, na2 = pk[34].m00;

            // 1. feed-forward part, in place (descending so the taps are still inputs)
#pragma unroll
            for (int n = S - 1; n >= 2; --n) v[n] = b0 * v[n] + b1 * v[n - 1] + b2 * v[n - 2];
            v[1] = b0 * v[1] + b1 * v[0] + b2 * um1;
            v[0] = b0 * v[0] + b1 * um1 + b2 * um2;

            // 2. zero-state recursion -> end state
            T zx = T(0), zy = T(0);
#pragma unroll
            for (int n = 0; n < S; ++n) {
                const T w = fma(na1, zx, fma(na2, zy, v[n]));
                zy = zx;
                zx = w;
            }

            // 3. scan of s' = M s + z across the tile
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int d = 1 << j;
                const T px = __shfl_up_sync(0xffffffffu, zx, d);
                const T py = __shfl_up_sync(0xffffffffu, zy, d);
                const Mat2<T> m = pk[d];
                if (lane >= d) {
                    zx += m.m00 * px + m.m01 * py;
                    zy += m.m10 * px + m.m11 * py;
                }
            }
            T* wt = wtot + (size_t)(k & 1) * NW * 2;
            if (lane == 31) {
                wt[warp * 2 + 0] = zx;
                wt[warp * 2 + 1] = zy;
            }
            if (k == 0) {
                // incoming state of every section: left by the previous tile of this row
                if (warp == 0) {
                    if (t_idx > 0) {
                        if (lane == 0) chain_wait(p.flags + row, t_idx);
                        __syncwarp();
                        for (int i = lane; i < 2 * K; i += 32)
                            s_in[i] = __ldcg(p.state + (size_t)row * 2 * K + i);
                    } else {
                        for (int i = lane; i < 2 * K; i += 32) s_in[i] = T(0);
                    }
                }
            }
            __syncthreads();
            T sx = s_in[2 * k], sy = s_in[2 * k + 1];
            {
                const Mat2<T> mw = pk[32];
                for (int w = 0; w < warp; ++w) {
                    const T tx = mw.m00 * sx + mw.m01 * sy + wt[w * 2 + 0];
                    const T ty = mw.m10 * sx + mw.m11 * sy + wt[w * 2 + 1];
                    sx = tx;
                    sy = ty;
                }
            }
            T ex = __shfl_up_sync(0xffffffffu, zx, 1);
            T ey = __shfl_up_sync(0xffffffffu, zy, 1);
            if (lane == 0) { ex = T(0); ey = T(0); }
            const Mat2<T> ml = pk[lane];
            T y1 = ml.m00 * sx + ml.m01 * sy + ex;  // y[-1] of this chunk
            T y2 = ml.m10 * sx + ml.m11 * sy + ey;  // y[-2]
            um1 = y1;  // next section's input history
            um2 = y2;

            // 4. the recursion again, from the true state
#pragma unroll
            for (int n = 0; n < S; ++n) {
                const T w = fma(na1, y1, fma(na2, y2, v[n]));
                y2 = y1;
                y1 = w;
                v[n] = w;
            }
            if (tid == NT - 1 && t_idx + 1 < p.tiles) {
                p.state[(size_t)row * 2 * K + 2 * k] = y1;
                p.state[(size_t)row * 2 * K + 2 * k + 1] = y2;
            }
        }
        if (tid == NT - 1 && t_idx + 1 < p.tiles) chain_publish(p.flags + row, t_idx + 1);

        // ---- registers -> tile -> global
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint4 q;
            T* qs = reinterpret_cast<T*>(&q);
#pragma unroll
            for (int e = 0; e < PER; ++e) qs[e] = v[u * PER + e];
            tile4[swz_unit(tid, u)] = q;
        }
        __syncthreads();
        if (p.aligned) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = tid + j * NT;
                const long long pos = (long long)g * PER;
                if (pos + PER <= remain) {
                    const uint4 q = tile4[swz_unit(g >> 3, g & 7)];
                    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(
                                     yr + t0 + pos),
                                 "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w)
                                 : "memory");
                } else if (pos < remain) {
                    const size_t sb = (size_t)swz_unit(g >> 3, g & 7) * PER;
                    for (int e = 0; e < PER && pos + e < remain; ++e) yr[t0 + pos + e] = tile[sb + e];
                }
            }
        } else {
            for (int i = tid; i < TILE && i < remain; i += NT) {
                const int r = i / S, n = i - r * S;
                yr[t0 + i] = tile[(size_t)swz_unit(r, n / PER) * PER + (n % PER)];
            }
        }
        item = next_item;
        st ^= 1;
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// fp32 fast path: TWO chunks per thread, packed into the lanes of Blackwell's FFMA2 / FMUL2
// (fma.rn.f32x2, __ffma2_rn): one issue slot does two FMAs, which frees the other issue slots of
// the SMSP for the scan / staging instructions (measured on B200: FFMA and FFMA2 both peak at
// ~122 FMA/clk/SM, FFMA2 with half the instructions -- tools/fma_peak.cu).
// 128 threads; the tile is 256 rows of 32 samples; thread t owns rows t ("A", lane .x) and t+128
// ("B", lane .y), i.e. samples [32t, 32t+32) and [4096+32t, ...).  The scan runs packed over both
// halves; the B half starts from the state the A half ends with.
// Each WARP is an independent worker: it owns a tile of 64 rows x 32 samples (lane l: rows l and
// l+32), takes its own tickets, and never meets a block barrier -- the only synchronisation is
// __syncwarp and the row chain in global memory.  With 4 such warps per SMSP the dependent
// FFMA2 chains, the shuffle scan and the tile loads of different warps overlap freely.
#ifndef GFX_X2_WARPS
#define GFX_X2_WARPS 4
#endif
// warps per CTA = consecutive sub-tiles of one ticket (16 warps per SM in all).  Measured on B200 (tools/cascade_shapes.py):
// 2 warps gain 2 % on 256 x 2 x 131072, K = 5 but lose 10..40 % on few rows, K = 1 and K >= 10 (twice as many global
// hand-offs along a row); 8 warps lose 10 % (longer ripple inside the CTA).
constexpr int X2_WARPS = GFX_X2_WARPS;
constexpr int X2_ROWS = 64;           // 128-byte rows per warp tile
constexpr int X2_TILE = X2_ROWS * 32; // 2048 samples

static __host__ __device__ __forceinline__ size_t x2_warp_smem_bytes(int) {
    // tile | 2 history samples + pad   (the section tables are shared by the warps of the CTA: same row)
    return (size_t)X2_ROWS * 128 + 16;
}
static __host__ __device__ __forceinline__ size_t x2_tab_smem_bytes(int K) { return (size_t)K * X2_TAB * 32; }

// KT > 0: the section loop is fully unrolled for K == KT (no loop-carried register shuffling of the
// 64 sample registers); KT == 0: generic K.
template <int MINB, int KT>
__global__ void __launch_bounds__(32 * X2_WARPS, MINB) biquad_cascade_x2_kernel(const CascadeParams<float> p) {
    constexpr int S = 32, TILE = X2_TILE, ITEM = X2_WARPS * X2_TILE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = KT > 0 ? KT : p.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smem_raw + (size_t)warp * x2_warp_smem_bytes(K);
    float4* tile4 = reinterpret_cast<float4*>(base);
    float* tile = reinterpret_cast<float*>(base);
    float* hist = reinterpret_cast<float*>(base + (size_t)X2_ROWS * 128);
    unsigned char* tab_raw = smem_raw + (size_t)X2_WARPS * x2_warp_smem_bytes(K);   // one table per CTA (one row per item)
    const double* tab = reinterpret_cast<const double*>(tab_raw);
    // coalesced phases: unit g = lane + 32 j lives in row (lane>>3) + 4 j, column (lane&7) ^ (row&7)
    const int cu0 = (lane & 7) ^ (lane >> 3);
    // the 4 warps of a CTA take 4 CONSECUTIVE sub-tiles of one row (one ticket = 8192 samples) and
    // hand the state from warp to warp through shared memory (cheap hop); only the hop from the last
    // warp to the next item of the row goes through global memory.
    double2* totals = reinterpret_cast<double2*>(tab_raw + x2_tab_smem_bytes(K));  // [K][X2_WARPS]
    float2* sin_relay = reinterpret_cast<float2*>(totals + (size_t)K * X2_WARPS);  // [K] (slot sized as double2)
    volatile int* ready = reinterpret_cast<volatile int*>(reinterpret_cast<double2*>(sin_relay) + K);  // [X2_WARPS]
    volatile int* sin_ready = ready + X2_WARPS;
    __shared__ unsigned int sh_item;
    if (threadIdx.x <= X2_WARPS) ready[threadIdx.x] = 0;
    int seq_base = 0;

    for (;;) {
        __syncthreads();  // every warp is done with the previous item
        if (threadIdx.x == 0) sh_item = take_ticket(p.ticket, 0xffffffffu);
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= p.n_items) break;
        const int t_idx = (int)(item / (unsigned)p.rows);
        const int row = (int)(item - (unsigned)t_idx * (unsigned)p.rows);
        const int b = row / p.c_out, c = row - b * p.c_out;
        int bs = b;
        if (p.src_outer > 0) {
            const int node = b / p.src_outer;
            bs = (b - node * p.src_outer) * p.src_inner + node;
        }
        const float* xr = p.x + ((size_t)bs * p.c_sig + (p.c_sig == 1 ? 0 : c)) * (size_t)p.L;
        float* yr = p.y + (size_t)row * (size_t)p.L;
        const size_t crow = (size_t)b * p.c_filt + (p.c_filt == 1 ? 0 : c);
        const long long t0 = (long long)t_idx * ITEM + (long long)warp * TILE;  // this warp's sub-tile
        const long long remain = p.L - t0;                                        // may be <= 0 (all padding)
        const bool full = p.aligned && remain >= TILE;

        // ---- stage the tile, the section tables and the two samples before the tile
        if (full) {
            const float* src = xr + t0 + lane * 4;
            float4* dst = tile4 + (lane >> 3) * 8;
#pragma unroll
            for (int j = 0; j < 16; ++j) cp_async16(dst + j * 32 + (cu0 ^ ((j & 1) << 2)), src + j * 128, 16);
        } else if (p.aligned) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int g = lane + j * 32;
                const long long pos = (long long)g * 4;
                const long long nb = (remain - pos) * 4;
                const int src_bytes = nb >= 16 ? 16 : (nb > 0 ? (int)nb : 0);
                const float* src = src_bytes > 0 ? xr + t0 + pos : xr;
                cp_async16(&tile4[swz_unit(g >> 3, g & 7)], src, src_bytes);
            }
        } else {
            for (int i = lane; i < TILE; i += 32) {
                const float val = (i < remain) ? xr[t0 + i] : 0.f;
                tile[(size_t)swz_unit(i >> 5, (i & 31) >> 2) * 4 + (i & 3)] = val;
            }
        }
        {
            const unsigned char* tsrc = reinterpret_cast<const unsigned char*>(p.xtables + crow * K * X2_TAB * 4);
            for (int u = threadIdx.x; u < K * X2_TAB * 2; u += 32 * X2_WARPS) cp_async16(tab_raw + (size_t)u * 16, tsrc + (size_t)u * 16, 16);
            if (lane < 2) {
                if (t0 > 0 && t0 - 1 - lane < p.L) cp_async_small<4>(hist + lane, xr + t0 - 1 - lane);
                else hist[lane] = 0.f;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();  // the table was staged by all four warps

        if (p.xcopy != nullptr) {
            // the untouched input tile also goes to the signal buffer's source slice (c_sig == c_out: row = input row)
            float* xc = p.xcopy + (size_t)row * (size_t)p.L;
            if (full) {
                float4* dst = reinterpret_cast<float4*>(xc + t0) + lane;
                const float4* src = tile4 + (lane >> 3) * 8;
#pragma unroll
                for (int j = 0; j < 16; ++j) stg_stream(dst + j * 32, src[j * 32 + (cu0 ^ ((j & 1) << 2))]);
            } else {
                for (int i = lane; i < TILE && i < remain; i += 32)
                    xc[t0 + i] = tile[(size_t)swz_unit(i >> 5, (i & 31) >> 2) * 4 + (i & 3)];
            }
        }

        pk2 v[S];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 qa = tile4[swz_unit(lane, u)];
            const float4 qb = tile4[swz_unit(lane + 32, u)];
            v[4 * u + 0] = pk_make(qa.x, qb.x);
            v[4 * u + 1] = pk_make(qa.y, qb.y);
            v[4 * u + 2] = pk_make(qa.z, qb.z);
            v[4 * u + 3] = pk_make(qa.w, qb.w);
        }
        pk2 um1, um2;  // the two input samples before each chunk
        {
            const int rb = swz_unit(lane + 31, 7) * 4;
            float a1v, a2v;
            if (lane > 0) {
                const int ra = swz_unit(lane - 1, 7) * 4;
                a1v = tile[ra + 3]; a2v = tile[ra + 2];
            } else {
                a1v = hist[0]; a2v = hist[1];
            }
            um1 = pk_make(a1v, tile[rb + 3]);
            um2 = pk_make(a2v, tile[rb + 2]);
        }

        // Incoming section states of the item, all requested at once: the previous item of the row was issued a whole
        // round of tickets earlier and has usually finished every section, so one L2 round trip here replaces one per
        // section on warp 0's critical path (a word whose tag is not there yet is polled for in its section, as before).
        unsigned long long pre_state = 0ull;
        if (warp == 0 && t_idx > 0 && lane < 2 * K)
            pre_state = *reinterpret_cast<const volatile unsigned long long*>(p.xstate + (size_t)row * K * 2 + lane);

#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : K); ++k) {
            const double* pk = tab + (size_t)k * X2_TAB * 4;
            // float copies of the coefficients in the table (no double loads + conversions in the section loop)
            const float4 cf = reinterpret_cast<const float4*>(pk + 37 * 4)[35];
            const float2 cg = *reinterpret_cast<const float2*>(reinterpret_cast<const float4*>(pk + 37 * 4) + 36);
            const pk2 B0 = pk_dup(cf.x), B1 = pk_dup(cf.y), B2 = pk_dup(cf.z);
            const pk2 NA1 = pk_dup(cf.w), NA2 = pk_dup(cg.x);

            // 1. feed-forward part in place (descending, so the taps are still inputs) with ZERO input history:
            //    the zero-state response below is then a genuine filter output (bounded by the filter gain), not
            //    the difference of two large transients
#pragma unroll
            for (int n = S - 1; n >= 2; --n) {
                pk_mul_acc(v[n], B0);
                pk_fma_acc(v[n], B1, v[n - 1]);
                pk_fma_acc(v[n], B2, v[n - 2]);
            }
            pk_mul_acc(v[1], B0); pk_fma_acc(v[1], B1, v[0]);
            pk_mul_acc(v[0], B0);

            // 2. zero-state recursion -> end states (both halves at once)
            pk2 z1 = 0ull, z2 = 0ull;
#pragma unroll
            for (int n = 0; n < S; ++n) {
                const pk2 w = pk_fma(NA1, z1, pk_fma(NA2, z2, v[n]));
                z2 = z1;
                z1 = w;
            }

            // 3. carries: c = z + N (u[-1], u[-2]), the scan of s' = M s + c over the 32 lanes, the warps stitched through
            //    shared memory, the item's incoming state from the previous item of the row.  Ill-conditioned sections
            //    (table flag) in double, the others in packed fp32; the hand-over words are the same for both.
            const bool ill = cg.y != 0.f;
            const float4* pf = reinterpret_cast<const float4*>(pk + 37 * 4);
            const int seq = seq_base + k + 1;
            double cAx = 0.0, cAy = 0.0, cBx = 0.0, cBy = 0.0;  // double path: inclusive carries of the lane's two chunks
            pk2 c1 = 0ull, c2 = 0ull;                            // fp32 path: the same, packed (A, B)
            if (ill) {
                float uA1, uB1, uA2, uB2, zAx, zBx, zAy, zBy;
                pk_split(um1, uA1, uB1);
                pk_split(um2, uA2, uB2);
                pk_split(z1, zAx, zBx);
                pk_split(z2, zAy, zBy);
                const double n00 = pk[33 * 4 + 0], n01 = pk[33 * 4 + 1], n10 = pk[33 * 4 + 2], n11 = pk[33 * 4 + 3];
                cAx = fma(n00, (double)uA1, fma(n01, (double)uA2, (double)zAx));
                cAy = fma(n10, (double)uA1, fma(n11, (double)uA2, (double)zAy));
                cBx = fma(n00, (double)uB1, fma(n01, (double)uB2, (double)zBx));
                cBy = fma(n10, (double)uB1, fma(n11, (double)uB2, (double)zBy));
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int d = 1 << j;
                    const double pAx = __shfl_up_sync(0xffffffffu, cAx, d), pAy = __shfl_up_sync(0xffffffffu, cAy, d);
                    const double pBx = __shfl_up_sync(0xffffffffu, cBx, d), pBy = __shfl_up_sync(0xffffffffu, cBy, d);
                    if (lane >= d) {
                        const double m00 = pk[d * 4 + 0], m01 = pk[d * 4 + 1], m10 = pk[d * 4 + 2], m11 = pk[d * 4 + 3];
                        cAx = fma(m00, pAx, fma(m01, pAy, cAx));
                        cAy = fma(m10, pAx, fma(m11, pAy, cAy));
                        cBx = fma(m00, pBx, fma(m01, pBy, cBx));
                        cBy = fma(m10, pBx, fma(m11, pBy, cBy));
                    }
                }
                if (lane == 31) {
                    const double w00 = pk[32 * 4 + 0], w01 = pk[32 * 4 + 1], w10 = pk[32 * 4 + 2], w11 = pk[32 * 4 + 3];  // M^32
                    totals[k * X2_WARPS + warp] = make_double2(fma(w00, cAx, fma(w01, cAy, cBx)), fma(w10, cAx, fma(w11, cAy, cBy)));
                    __threadfence_block();
                    ready[warp] = seq;
                }
            } else {
                const float4 nn = pf[33];
                c1 = pk_fma(pk_dup(nn.x), um1, pk_fma(pk_dup(nn.y), um2, z1));
                c2 = pk_fma(pk_dup(nn.z), um1, pk_fma(pk_dup(nn.w), um2, z2));
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int d = 1 << j;
                    pk2 p1 = pk_shfl_up(c1, d), p2 = pk_shfl_up(c2, d);
                    const float4 m = pf[lane < d ? 37 : d];  // zero matrix: lanes below the stride keep their carry
                    c1 = pk_fma(pk_dup(m.x), p1, pk_fma(pk_dup(m.y), p2, c1));
                    c2 = pk_fma(pk_dup(m.z), p1, pk_fma(pk_dup(m.w), p2, c2));
                }
                if (lane == 31) {
                    float zAx, zBx, zAy, zBy;
                    pk_split(c1, zAx, zBx);
                    pk_split(c2, zAy, zBy);
                    const float4 mw = pf[32];
                    totals[k * X2_WARPS + warp] = make_double2((double)fmaf(mw.x, zAx, fmaf(mw.y, zAy, zBx)),
                                                               (double)fmaf(mw.z, zAx, fmaf(mw.w, zAy, zBy)));
                    __threadfence_block();
                    ready[warp] = seq;
                }
            }
            // incoming state of the item for THIS section: one 64-bit {value, tag} word per component
            // (an aligned 8-byte access is single-copy atomic: no fence, no separate flag), published
            // section by section so that consecutive items of a row run as a pipeline.  The state is the pair of
            // fp32 OUTPUT samples before the item: exact in fp32.
            // Only warp 0 reads the global word (the last warp of this very item overwrites it with the
            // next tag as soon as IT is done with the section); it relays the state through shared memory.
            float sAxf = 0.f, sAyf = 0.f;
            if (t_idx > 0) {
                if (warp == 0) {
                    float sv = 0.f;
                    // (sections 0..15 come from the words requested before the loop)
                    unsigned long long word = __shfl_sync(0xffffffffu, pre_state, (2 * k + (lane & 1)) & 31);
                    if (lane < 2) {
                        if (2 * k + 1 >= 32 || (unsigned)(word >> 32) != (unsigned)t_idx) {
                            const volatile unsigned long long* w = p.xstate + ((size_t)row * K + k) * 2 + lane;
                            do { word = *w; } while ((unsigned)(word >> 32) != (unsigned)t_idx);
                        }
                        sv = __uint_as_float((unsigned)word);
                    }
                    sAxf = __shfl_sync(0xffffffffu, sv, 0);
                    sAyf = __shfl_sync(0xffffffffu, sv, 1);
                    if (lane == 0) {
                        sin_relay[2 * k] = make_float2(sAxf, sAyf);
                        __threadfence_block();
                        *sin_ready = seq;
                    }
                } else {
                    if (lane == 0) { while (*sin_ready < seq) { } }
                    __syncwarp();
                    __threadfence_block();
                    const float2 si = sin_relay[2 * k];
                    sAxf = si.x;
                    sAyf = si.y;
                }
            }
            if (warp > 0) {
                if (lane < warp) { while (ready[lane] < seq) { } }
                __syncwarp();
                __threadfence_block();
            }
            pk2 y1, y2;
            if (ill) {
                double sAx = (double)sAxf, sAy = (double)sAyf;
                const double q00 = pk[35 * 4 + 0], q01 = pk[35 * 4 + 1], q10 = pk[35 * 4 + 2], q11 = pk[35 * 4 + 3];  // M^64
                for (int q = 0; q < warp; ++q) {
                    const double2 tq = totals[k * X2_WARPS + q];
                    const double tx = fma(q00, sAx, fma(q01, sAy, tq.x));
                    const double ty = fma(q10, sAx, fma(q11, sAy, tq.y));
                    sAx = tx;
                    sAy = ty;
                }
                const double totAx = __shfl_sync(0xffffffffu, cAx, 31), totAy = __shfl_sync(0xffffffffu, cAy, 31);
                const double w00 = pk[32 * 4 + 0], w01 = pk[32 * 4 + 1], w10 = pk[32 * 4 + 2], w11 = pk[32 * 4 + 3];  // M^32
                const double sBx = fma(w00, sAx, fma(w01, sAy, totAx));
                const double sBy = fma(w10, sAx, fma(w11, sAy, totAy));
                double eAx = __shfl_up_sync(0xffffffffu, cAx, 1), eAy = __shfl_up_sync(0xffffffffu, cAy, 1);
                double eBx = __shfl_up_sync(0xffffffffu, cBx, 1), eBy = __shfl_up_sync(0xffffffffu, cBy, 1);
                if (lane == 0) { eAx = 0.0; eAy = 0.0; eBx = 0.0; eBy = 0.0; }
                const double l00 = pk[lane * 4 + 0], l01 = pk[lane * 4 + 1], l10 = pk[lane * 4 + 2], l11 = pk[lane * 4 + 3];
                y1 = pk_make((float)fma(l00, sAx, fma(l01, sAy, eAx)), (float)fma(l00, sBx, fma(l01, sBy, eBx)));  // y[-1] of the chunks
                y2 = pk_make((float)fma(l10, sAx, fma(l11, sAy, eAy)), (float)fma(l10, sBx, fma(l11, sBy, eBy)));  // y[-2]
            } else {
                float sAx = sAxf, sAy = sAyf;
                const float4 m64 = pf[34];
                for (int q = 0; q < warp; ++q) {
                    const double2 tq = totals[k * X2_WARPS + q];
                    const float tx = fmaf(m64.x, sAx, fmaf(m64.y, sAy, (float)tq.x));
                    const float ty = fmaf(m64.z, sAx, fmaf(m64.w, sAy, (float)tq.y));
                    sAx = tx;
                    sAy = ty;
                }
                float zAx, zBx, zAy, zBy;
                pk_split(c1, zAx, zBx);
                pk_split(c2, zAy, zBy);
                const float totAx = __shfl_sync(0xffffffffu, zAx, 31), totAy = __shfl_sync(0xffffffffu, zAy, 31);
                const float4 mw = pf[32];
                const float sBx = fmaf(mw.x, sAx, fmaf(mw.y, sAy, totAx));
                const float sBy = fmaf(mw.z, sAx, fmaf(mw.w, sAy, totAy));
                pk2 e1 = pk_shfl_up(c1, 1), e2 = pk_shfl_up(c2, 1);
                if (lane == 0) { e1 = 0ull; e2 = 0ull; }
                const float4 ml = pf[lane];
                const pk2 S1 = pk_make(sAx, sBx), S2 = pk_make(sAy, sBy);
                y1 = pk_fma(pk_dup(ml.x), S1, pk_fma(pk_dup(ml.y), S2, e1));
                y2 = pk_fma(pk_dup(ml.z), S1, pk_fma(pk_dup(ml.w), S2, e2));
            }
            // the feed-forward taps on the two inputs before the chunk, left out in step 1
            pk_fma_acc(v[0], B1, um1); pk_fma_acc(v[0], B2, um2);
            pk_fma_acc(v[1], B2, um1);
            um1 = y1;
            um2 = y2;

            // 4. the recursion again, from the true states (on the array itself: no register rotation)
            pk_fma_acc(v[0], NA2, y2); pk_fma_acc(v[0], NA1, y1);
            pk_fma_acc(v[1], NA2, y1); pk_fma_acc(v[1], NA1, v[0]);
#pragma unroll
            for (int n = 2; n < S; ++n) {
                pk_fma_acc(v[n], NA2, v[n - 2]);
                pk_fma_acc(v[n], NA1, v[n - 1]);
            }
            if (warp == X2_WARPS - 1 && lane == 31 && t_idx + 1 < p.tiles) {
                float lo, hi1, hi2;
                pk_split(v[S - 1], lo, hi1);
                pk_split(v[S - 2], lo, hi2);
                volatile unsigned long long* w = p.xstate + ((size_t)row * K + k) * 2;
                const unsigned long long tag = (unsigned long long)(unsigned)(t_idx + 1) << 32;
                w[0] = tag | __float_as_uint(hi1);
                w[1] = tag | __float_as_uint(hi2);
            }
        }

        // ---- registers -> tile -> global
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float4 qa, qb;
            pk_split(v[4 * u + 0], qa.x, qb.x);
            pk_split(v[4 * u + 1], qa.y, qb.y);
            pk_split(v[4 * u + 2], qa.z, qb.z);
            pk_split(v[4 * u + 3], qa.w, qb.w);
            tile4[swz_unit(lane, u)] = qa;
            tile4[swz_unit(lane + 32, u)] = qb;
        }
        __syncwarp();
        if (full) {
            float4* dst = reinterpret_cast<float4*>(yr + t0) + lane;
            const float4* src = tile4 + (lane >> 3) * 8;
#pragma unroll
            for (int j = 0; j < 16; ++j) stg_stream(dst + j * 32, src[j * 32 + (cu0 ^ ((j & 1) << 2))]);
        } else if (p.aligned) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int g = lane + j * 32;
                const long long pos = (long long)g * 4;
                if (pos + 4 <= remain) {
                    stg_stream(reinterpret_cast<float4*>(yr + t0 + pos), tile4[swz_unit(g >> 3, g & 7)]);
                } else if (pos < remain) {
                    const int sb = swz_unit(g >> 3, g & 7) * 4;
                    for (int e = 0; e < 4 && pos + e < remain; ++e) yr[t0 + pos + e] = tile[sb + e];
                }
            }
        } else {
            for (int i = lane; i < TILE && i < remain; i += 32)
                yr[t0 + i] = tile[(size_t)swz_unit(i >> 5, (i & 31) >> 2) * 4 + (i & 3)];
        }
        seq_base += K;
    }
}

static size_t cascade_x2_smem_bytes(int K) {
    return (size_t)X2_WARPS * x2_warp_smem_bytes(K) + x2_tab_smem_bytes(K) + (size_t)K * (X2_WARPS + 1) * sizeof(double2) +
           (X2_WARPS + 1) * sizeof(int) + 16;
}

template <typename T>
static size_t cascade_smem_bytes(int NT, int K) {
    const size_t stage = (size_t)NT * 128 + (size_t)K * TAB_ENTRIES * 4 * sizeof(T) + 16;
    return 2 * stage + (size_t)K * 2 * sizeof(T) + (size_t)2 * (NT / 32) * 2 * sizeof(T) + 64;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static size_t cascade_workspace_bytes(int rows, int coef_rows, int K, size_t elem) {
    // [ticket | pad to 256] [flags: rows*K ints] [state: rows*2K elems] [tables: coef_rows*K*36*4 elems]
    // (the packed fp32 kernel keeps its tables in double: X2_TAB entries of 4 doubles per section)
    const size_t tab = elem == 4 ? (size_t)coef_rows * K * X2_TAB * 4 * sizeof(double) : (size_t)coef_rows * K * TAB_ENTRIES * 4 * elem;
    return 256 + align256((size_t)rows * K * sizeof(int)) + align256((size_t)rows * 2 * K * 8) + tab;
}

template <typename T>
static int launch_cascade(const T* x, T* y, const T* Bs, const T* As, int batch, int c_sig,
                          int c_filt, int K, long long L, void* ws, size_t ws_bytes,
                          cudaStream_t stream, T* xcopy = nullptr, int src_outer = 0, int src_inner = 0, int coef_repeat = 1) {
    constexpr int NT = 256;
    constexpr int S = 128 / (int)sizeof(T);
    constexpr int MINB = sizeof(T) == 4 ? 3 : 1;
    if (!x || !y || !Bs || !As) return GFX_ERR_INVALID;
    if (batch <= 0 || c_sig <= 0 || c_filt <= 0 || K <= 0 || L <= 0) return GFX_ERR_INVALID;
    if (c_sig != c_filt && c_sig != 1 && c_filt != 1) return GFX_ERR_INVALID;
    if (K > 64) return GFX_ERR_UNSUPPORTED;
    const int c_out = c_sig > c_filt ? c_sig : c_filt;
    if (coef_repeat < 1 || batch % coef_repeat != 0 || (coef_repeat > 1 && sizeof(T) != 4)) return GFX_ERR_INVALID;
    if (xcopy != nullptr || src_outer != 0 || src_inner != 0) {
        // source-reading form: fp32 packed kernel only, one output row per input row, batch = src_outer * src_inner
        if (sizeof(T) != 4 || xcopy == nullptr || c_sig != c_out) return GFX_ERR_INVALID;
        if (src_outer <= 0 || src_inner <= 0 || (long long)src_outer * src_inner != (long long)batch) return GFX_ERR_INVALID;
    }
    const long long rows_ll = (long long)batch * c_out;
    static const bool force_scalar = (getenv("GFX_CASCADE_SCALAR") != nullptr);  // A/B testing only
    (void)force_scalar;
    const bool X2 = sizeof(T) == 4;  // fp32: packed warp-tile kernel with double-precision carries (fits for every K <= 64)
    if (X2 && cascade_x2_smem_bytes(K) > (size_t)device_info().max_smem_optin) return GFX_ERR_UNSUPPORTED;
    const long long tile_len = X2 ? (long long)X2_WARPS * X2_TILE : (long long)NT * S;
    const long long tiles_ll = (L + tile_len - 1) / tile_len;
    if (rows_ll * tiles_ll > 0x7fff0000LL) return GFX_ERR_UNSUPPORTED;
    const int rows = (int)rows_ll, tiles = (int)tiles_ll;
    const int coef_rows = batch * c_filt;
    const size_t need = cascade_workspace_bytes(rows, coef_rows, K, sizeof(T));
    if (!ws || ws_bytes < need) return GFX_ERR_WORKSPACE;

    CascadeParams<T> p;
    p.x = x; p.y = y;
    p.xcopy = xcopy; p.src_outer = src_outer; p.src_inner = src_inner;
    p.batch = batch; p.c_sig = c_sig; p.c_filt = c_filt; p.c_out = c_out; p.K = K;
    p.L = L; p.rows = rows; p.tiles = tiles;
    p.n_items = (unsigned)(rows * (long long)tiles);
    unsigned char* w = (unsigned char*)ws;
    p.ticket = (unsigned int*)w;
    p.flags = (int*)(w + 256);
    const size_t flags_bytes = align256((size_t)rows * K * sizeof(int));
    p.state = (T*)(w + 256 + flags_bytes);
    p.xstate = (unsigned long long*)(w + 256 + flags_bytes);
    const size_t state_bytes = align256((size_t)rows * 2 * K * 8);
    T* tables = (T*)(w + 256 + flags_bytes + state_bytes);
    p.tables = tables;
    p.aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)xcopy) % 16 == 0) && ((L * (long long)sizeof(T)) % 16 == 0);

    GFX_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256 + flags_bytes + state_bytes, stream));
    const int n_sections = coef_rows * K;
    if constexpr (sizeof(T) == 4) {
        p.xtables = (const double*)tables;
        cascade_x2_tables_kernel<<<(n_sections * 32 + 127) / 128, 128, 0, stream>>>(Bs, As, (double*)tables, n_sections, K, c_filt, coef_repeat);
    } else {
        p.xtables = nullptr;
        cascade_tables_kernel<T><<<(n_sections * 32 + 127) / 128, 128, 0, stream>>>(Bs, As, tables, n_sections);
    }
    GFX_LAUNCH_CHECK();

    if constexpr (sizeof(T) == 4) {
        const size_t smem2 = cascade_x2_smem_bytes(K);
        if (X2) {
            // (fully unrolling the section loop per K was measured: no gain -- ptxas keeps its
            //  register-pair copies -- and a bigger instruction footprint; generic K only)
            // (occupancy was measured on B200, profiles/r02_cascade_occupancy.txt: 5 CTAs / 96 registers 215 us, 6 CTAs / 80
            //  registers 314 us vs 201 us here -- the spills cost more than the extra warps hide; the end states of the
            //  zero-state pass as dot products with the tabulated impulse response: 207 us and less accurate)
            auto kern2 = biquad_cascade_x2_kernel<16 / X2_WARPS, 0>;
            static size_t configured2_dev[64] = {0};
            size_t& configured2 = configured2_dev[device_slot()];
            if (smem2 > configured2) {
                GFX_CUDA_CHECK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                configured2 = smem2;
            }
            int occ2 = 0;
            GFX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern2, 32 * X2_WARPS, smem2));
            if (occ2 < 1) return GFX_ERR_UNSUPPORTED;
            long long grid2 = (long long)device_info().sm_count * occ2;
            if (grid2 > (long long)p.n_items) grid2 = p.n_items;
            kern2<<<(unsigned)grid2, 32 * X2_WARPS, smem2, stream>>>(p);
            GFX_LAUNCH_CHECK();
            return GFX_OK;
        }
    }
    if constexpr (sizeof(T) == 4) {
        return GFX_ERR_UNSUPPORTED;  // (not reached: fp32 always takes the packed kernel above)
    } else {
    const size_t smem = cascade_smem_bytes<T>(NT, K);
    if (smem > (size_t)device_info().max_smem_optin) return GFX_ERR_UNSUPPORTED;
    auto kern = biquad_cascade_kernel<T, NT, MINB>;
    static size_t configured_smem[64][2] = {{0, 0}};
    const int slot = sizeof(T) == 4 ? 0 : 1;
    size_t& conf = configured_smem[device_slot()][slot];
    if (smem > conf) {
        GFX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    int occ = 0;
    GFX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return GFX_ERR_UNSUPPORTED;
    long long grid = (long long)device_info().sm_count * occ;
    if (grid > (long long)p.n_items) grid = p.n_items;
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    GFX_LAUNCH_CHECK();
    return GFX_OK;
    }
}

}  // namespace gfx

extern "C" {

size_t gfx_biquad_cascade_workspace_bytes(int batch, int c_sig, int c_filt, int K, int elem_size) {
    const int c_out = c_sig > c_filt ? c_sig : c_filt;
    return gfx::cascade_workspace_bytes(batch * c_out, batch * c_filt, K, (size_t)elem_size);
}

int gfx_biquad_cascade_f32(const float* x, float* y, const float* Bs, const float* As, int batch,
                           int c_sig, int c_filt, int K, long long L, void* workspace,
                           size_t workspace_bytes, void* stream) {
    return gfx::launch_cascade<float>(x, y, Bs, As, batch, c_sig, c_filt, K, L, workspace,
                                      workspace_bytes, (cudaStream_t)stream);
}

int gfx_biquad_cascade_ex_f32(const float* x, float* xcopy, float* y, const float* Bs, const float* As, int batch,
                              int c_sig, int c_filt, int K, long long L, int src_outer, int src_inner, int coef_repeat,
                              void* workspace, size_t workspace_bytes, void* stream) {
    return gfx::launch_cascade<float>(x, y, Bs, As, batch, c_sig, c_filt, K, L, workspace, workspace_bytes,
                                      (cudaStream_t)stream, xcopy, src_outer, src_inner, coef_repeat);
}

int gfx_biquad_cascade_f64(const double* x, double* y, const double* Bs, const double* As, int batch,
                           int c_sig, int c_filt, int K, long long L, void* workspace,
                           size_t workspace_bytes, void* stream) {
    return gfx::launch_cascade<double>(x, y, Bs, As, batch, c_sig, c_filt, K, L, workspace,
                                       workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
