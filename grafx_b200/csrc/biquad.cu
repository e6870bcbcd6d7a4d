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
// (cp.async, XOR-swizzled 128-byte rows) so that thread t owns the S contiguous samples
// [t*S, (t+1)*S) in registers.  Per section:
//   1. feed-forward part in place (needs the two samples before the chunk: the neighbour's state);
//   2. zero-state recursion over the chunk -> end state z_t (two values);
//   3. the true state at each chunk start is s_{t+1} = M s_t + z_t with M = A^S, A = [[-a1,-a2],[1,0]]:
//      warp-shuffle scan with the matrices M^(2^j), warps stitched through shared memory, the tile's
//      incoming state taken from the previous tile of the row (ordered chain, see common.cuh);
//   4. the recursion is re-run from the true state -- this pass is arithmetically the sequential
//      DF-I loop, so rounding differs from the reference only through the carried state.
// Powers M^l (l = 0..32) are computed per tile in double precision by one warp per section.
// HBM traffic: x read once, y written once (8 B/sample fp32) + 2K words of state per tile.
#include "common.cuh"
#include "../../include/grafx_b200.h"

namespace gfx {

template <typename T>
struct CascadeParams {
    const T* x;
    T* y;
    const T* Bs;
    const T* As;
    int batch, c_sig, c_filt, c_out, K;
    long long L;
    int rows, tiles;
    unsigned int n_items;
    unsigned int* ticket;
    int* flags;  // [rows]   number of finished tiles of the row
    T* state;    // [rows][2K] (y_k[-1], y_k[-2]) left by the last finished tile
    int aligned;  // 1: x/y rows are 16-byte aligned (vector path)
};

template <typename T>
struct Mat2 {
    T m00, m01, m10, m11;
};

__device__ __forceinline__ void mat2_mul(const double a[4], const double b[4], double c[4]) {
    c[0] = a[0] * b[0] + a[1] * b[2];
    c[1] = a[0] * b[1] + a[1] * b[3];
    c[2] = a[2] * b[0] + a[3] * b[2];
    c[3] = a[2] * b[1] + a[3] * b[3];
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) biquad_cascade_kernel(const CascadeParams<T> p) {
    constexpr int S = 128 / (int)sizeof(T);   // samples per thread (32 fp32 / 16 fp64)
    constexpr int PER = 16 / (int)sizeof(T);  // samples per 16-byte unit
    constexpr int TILE = NT * S;
    constexpr int NW = NT / 32;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* tile4 = reinterpret_cast<uint4*>(smem_raw);            // NT*8 units
    T* tile = reinterpret_cast<T*>(smem_raw);
    T* coef = reinterpret_cast<T*>(smem_raw + (size_t)NT * 128);  // [K][8] (5 used)
    Mat2<T>* ptab = reinterpret_cast<Mat2<T>*>(coef + (size_t)p.K * 8);  // [K][33]
    T* s_in = reinterpret_cast<T*>(ptab + (size_t)p.K * 33);      // [K][2]
    T* wtot = s_in + (size_t)p.K * 2;                             // [2][NW][2]
    __shared__ unsigned int sh_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.K;

    for (;;) {
        __syncthreads();  // previous tile fully stored / sh_item consumed
        if (tid == 0) sh_item = take_ticket(p.ticket, p.n_items + gridDim.x - 1);
        __syncthreads();
        const unsigned int item = sh_item;
        if (item >= p.n_items) break;
        const int t_idx = (int)(item / (unsigned)p.rows);
        const int row = (int)(item - (unsigned)t_idx * (unsigned)p.rows);
        const int b = row / p.c_out, c = row - b * p.c_out;
        const T* xr = p.x + ((size_t)b * p.c_sig + (p.c_sig == 1 ? 0 : c)) * (size_t)p.L;
        T* yr = p.y + (size_t)row * (size_t)p.L;
        const size_t crow = (size_t)b * p.c_filt + (p.c_filt == 1 ? 0 : c);
        const long long t0 = (long long)t_idx * TILE;
        const long long remain = p.L - t0;  // > 0

        // ---- stage the input tile (zero padded past L)
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
            cp_async_commit();
        } else {
            for (int i = tid; i < TILE; i += NT) {
                const T val = (i < remain) ? xr[t0 + i] : T(0);
                const int r = i / S, n = i - r * S;
                tile[(size_t)swz_unit(r, n / PER) * PER + (n % PER)] = val;
            }
        }

        // ---- per-section constants: normalised coefficients and the powers of M = A^S
        for (int k = warp; k < K; k += NW) {
            const T* bp = p.Bs + (crow * K + k) * 3;
            const T* ap = p.As + (crow * K + k) * 3;
            const T a0 = ap[0];
            const T nb0 = bp[0] / a0, nb1 = bp[1] / a0, nb2 = bp[2] / a0;
            const T na1 = ap[1] / a0, na2 = ap[2] / a0;
            if (lane == 0) {
                T* ck = coef + (size_t)k * 8;
                ck[0] = nb0; ck[1] = nb1; ck[2] = nb2; ck[3] = na1; ck[4] = na2;
            }
            double base[4] = {-(double)na1, -(double)na2, 1.0, 0.0}, tmp[4];
            // M = A^S  (S = 2^5 or 2^4)
#pragma unroll 1
            for (int sq = 1; sq < S; sq <<= 1) {
                mat2_mul(base, base, tmp);
                base[0] = tmp[0]; base[1] = tmp[1]; base[2] = tmp[2]; base[3] = tmp[3];
            }
            // lane l -> M^l by binary exponentiation; after 5 squarings base = M^32
            double res[4] = {1.0, 0.0, 0.0, 1.0};
#pragma unroll 1
            for (int bit = 0; bit < 5; ++bit) {
                if ((lane >> bit) & 1) {
                    mat2_mul(res, base, tmp);
                    res[0] = tmp[0]; res[1] = tmp[1]; res[2] = tmp[2]; res[3] = tmp[3];
                }
                mat2_mul(base, base, tmp);
                base[0] = tmp[0]; base[1] = tmp[1]; base[2] = tmp[2]; base[3] = tmp[3];
            }
            Mat2<T>* pk = ptab + (size_t)k * 33;
            pk[lane] = Mat2<T>{(T)res[0], (T)res[1], (T)res[2], (T)res[3]};
            if (lane == 0) pk[32] = Mat2<T>{(T)base[0], (T)base[1], (T)base[2], (T)base[3]};
        }

        if (p.aligned) cp_async_wait<0>();
        __syncthreads();

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
            um1 = (t_idx > 0) ? xr[t0 - 1] : T(0);
            um2 = (t_idx > 0) ? xr[t0 - 2] : T(0);
        }

        for (int k = 0; k < K; ++k) {
            const T* ck = coef + (size_t)k * 8;
            const T b0 = ck[0], b1 = ck[1], b2 = ck[2], na1 = -ck[3], na2 = -ck[4];
            const Mat2<T>* pk = ptab + (size_t)k * 33;

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
    }
}

template <typename T>
static size_t cascade_smem_bytes(int NT, int K) {
    return (size_t)NT * 128 + (size_t)K * 8 * sizeof(T) + (size_t)K * 33 * 4 * sizeof(T) +
           (size_t)K * 2 * sizeof(T) + (size_t)2 * (NT / 32) * 2 * sizeof(T) + 64;
}

static size_t cascade_workspace_bytes(int rows, int K, size_t elem) {
    // [ticket | pad to 256] [flags: rows ints, padded to 256] [state: rows*2K elems]
    size_t flags = ((size_t)rows * sizeof(int) + 255) / 256 * 256;
    return 256 + flags + (size_t)rows * 2 * K * elem;
}

template <typename T>
static int launch_cascade(const T* x, T* y, const T* Bs, const T* As, int batch, int c_sig,
                          int c_filt, int K, long long L, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
    constexpr int NT = 256;
    constexpr int S = 128 / (int)sizeof(T);
    if (!x || !y || !Bs || !As) return GFX_ERR_INVALID;
    if (batch <= 0 || c_sig <= 0 || c_filt <= 0 || K <= 0 || L <= 0) return GFX_ERR_INVALID;
    if (c_sig != c_filt && c_sig != 1 && c_filt != 1) return GFX_ERR_INVALID;
    if (K > 64) return GFX_ERR_UNSUPPORTED;
    const int c_out = c_sig > c_filt ? c_sig : c_filt;
    const long long rows_ll = (long long)batch * c_out;
    const long long tiles_ll = (L + (long long)NT * S - 1) / ((long long)NT * S);
    if (rows_ll * tiles_ll > 0x7fff0000LL) return GFX_ERR_UNSUPPORTED;
    const int rows = (int)rows_ll, tiles = (int)tiles_ll;
    const size_t need = cascade_workspace_bytes(rows, K, sizeof(T));
    if (!ws || ws_bytes < need) return GFX_ERR_WORKSPACE;

    CascadeParams<T> p;
    p.x = x; p.y = y; p.Bs = Bs; p.As = As;
    p.batch = batch; p.c_sig = c_sig; p.c_filt = c_filt; p.c_out = c_out; p.K = K;
    p.L = L; p.rows = rows; p.tiles = tiles;
    p.n_items = (unsigned)(rows * (long long)tiles);
    unsigned char* w = (unsigned char*)ws;
    p.ticket = (unsigned int*)w;
    p.flags = (int*)(w + 256);
    const size_t flags_bytes = ((size_t)rows * sizeof(int) + 255) / 256 * 256;
    p.state = (T*)(w + 256 + flags_bytes);
    p.aligned = (((uintptr_t)x | (uintptr_t)y) % 16 == 0) && ((L * (long long)sizeof(T)) % 16 == 0);

    GFX_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256 + flags_bytes, stream));

    const size_t smem = cascade_smem_bytes<T>(NT, K);
    auto kern = biquad_cascade_kernel<T, NT>;
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
    return gfx::cascade_workspace_bytes(batch * c_out, K, (size_t)elem_size);
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
