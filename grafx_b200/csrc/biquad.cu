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
    int aligned;  // 1: x/y rows are 16-byte aligned (vector path)
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
        out[35 * 4 + 0] = T(0); out[35 * 4 + 1] = T(0); out[35 * 4 + 2] = T(0); out[35 * 4 + 3] = T(0);
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

template <typename T>
static size_t cascade_smem_bytes(int NT, int K) {
    const size_t stage = (size_t)NT * 128 + (size_t)K * TAB_ENTRIES * 4 * sizeof(T) + 16;
    return 2 * stage + (size_t)K * 2 * sizeof(T) + (size_t)2 * (NT / 32) * 2 * sizeof(T) + 64;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static size_t cascade_workspace_bytes(int rows, int coef_rows, int K, size_t elem) {
    // [ticket | pad to 256] [flags: rows ints] [state: rows*2K elems] [tables: coef_rows*K*36*4 elems]
    return 256 + align256((size_t)rows * sizeof(int)) + align256((size_t)rows * 2 * K * elem) +
           (size_t)coef_rows * K * TAB_ENTRIES * 4 * elem;
}

template <typename T>
static int launch_cascade(const T* x, T* y, const T* Bs, const T* As, int batch, int c_sig,
                          int c_filt, int K, long long L, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
    constexpr int NT = 256;
    constexpr int S = 128 / (int)sizeof(T);
    constexpr int MINB = sizeof(T) == 4 ? 3 : 1;
    if (!x || !y || !Bs || !As) return GFX_ERR_INVALID;
    if (batch <= 0 || c_sig <= 0 || c_filt <= 0 || K <= 0 || L <= 0) return GFX_ERR_INVALID;
    if (c_sig != c_filt && c_sig != 1 && c_filt != 1) return GFX_ERR_INVALID;
    if (K > 64) return GFX_ERR_UNSUPPORTED;
    const int c_out = c_sig > c_filt ? c_sig : c_filt;
    const long long rows_ll = (long long)batch * c_out;
    const long long tiles_ll = (L + (long long)NT * S - 1) / ((long long)NT * S);
    if (rows_ll * tiles_ll > 0x7fff0000LL) return GFX_ERR_UNSUPPORTED;
    const int rows = (int)rows_ll, tiles = (int)tiles_ll;
    const int coef_rows = batch * c_filt;
    const size_t need = cascade_workspace_bytes(rows, coef_rows, K, sizeof(T));
    if (!ws || ws_bytes < need) return GFX_ERR_WORKSPACE;

    CascadeParams<T> p;
    p.x = x; p.y = y;
    p.batch = batch; p.c_sig = c_sig; p.c_filt = c_filt; p.c_out = c_out; p.K = K;
    p.L = L; p.rows = rows; p.tiles = tiles;
    p.n_items = (unsigned)(rows * (long long)tiles);
    unsigned char* w = (unsigned char*)ws;
    p.ticket = (unsigned int*)w;
    p.flags = (int*)(w + 256);
    const size_t flags_bytes = align256((size_t)rows * sizeof(int));
    p.state = (T*)(w + 256 + flags_bytes);
    T* tables = (T*)(w + 256 + flags_bytes + align256((size_t)rows * 2 * K * sizeof(T)));
    p.tables = tables;
    p.aligned = (((uintptr_t)x | (uintptr_t)y) % 16 == 0) && ((L * (long long)sizeof(T)) % 16 == 0);

    GFX_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256 + flags_bytes, stream));
    const int n_sections = coef_rows * K;
    cascade_tables_kernel<T><<<(n_sections * 32 + 127) / 128, 128, 0, stream>>>(Bs, As, tables, n_sections);
    GFX_CUDA_CHECK(cudaGetLastError());

    const size_t smem = cascade_smem_bytes<T>(NT, K);
    if (smem > (size_t)device_info().max_smem_optin) return GFX_ERR_UNSUPPORTED;
    auto kern = biquad_cascade_kernel<T, NT, MINB>;
    static size_t configured_smem[2] = {0, 0};
    const int slot = sizeof(T) == 4 ? 0 : 1;
    if (smem > configured_smem[slot]) {
        GFX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_smem[slot] = smem;
    }
    int occ = 0;
    GFX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return GFX_ERR_UNSUPPORTED;
    long long grid = (long long)device_info().sm_count * occ;
    if (grid > (long long)p.n_items) grid = p.n_items;
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    GFX_CUDA_CHECK(cudaGetLastError());
    return GFX_OK;
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

int gfx_biquad_cascade_f64(const double* x, double* y, const double* Bs, const double* As, int batch,
                           int c_sig, int c_filt, int K, long long L, void* workspace,
                           size_t workspace_bytes, void* stream) {
    return gfx::launch_cascade<double>(x, y, Bs, As, batch, c_sig, c_filt, K, L, workspace,
                                       workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
