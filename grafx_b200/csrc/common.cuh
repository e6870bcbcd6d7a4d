// grafx_b200 -- shared device/host helpers for the sm_100a kernels.
// Everything here is header-only and private to csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/grafx_b200.h"  // error codes + the exported prototypes

extern int g_gfx_last_cuda_error;

#define GFX_CUDA_CHECK(expr)                                   \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) {                               \
            g_gfx_last_cuda_error = (int)_e;                   \
            return GFX_ERR_CUDA;                               \
        }                                                      \
    } while (0)

// after every kernel launch of the library: count it (gfx_kernel_launch_count) and pick up launch errors
extern unsigned long long g_gfx_launch_count;
#define GFX_LAUNCH_CHECK()                    \
    do {                                      \
        ++g_gfx_launch_count;                 \
        GFX_CUDA_CHECK(cudaGetLastError());   \
    } while (0)

namespace gfx {

// ---------------------------------------------------------------- device info (cached)
struct DeviceInfo {
    int sm_count;
    int max_smem_optin;
};
const DeviceInfo& device_info();
// index of the current device for per-device "already configured" state (cudaFuncSetAttribute is per device)
inline int device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= 64) ? 0 : dev;
}

// ---------------------------------------------------------------- cp.async (LDGSTS) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte async copy global->shared with zero fill of the bytes past src_bytes (0..16).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(src_bytes));
}
// 4- or 8-byte async copy (for the odd scalar that must travel with a tile)
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8, "cp.async.ca supports 4, 8, 16 bytes");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES));
}
// 8-byte async copy with zero fill of the bytes past src_bytes (0..8)
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming (evict-first) 128-bit global access for data touched exactly once
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w));
}

// ---------------------------------------------------------------- packed fp32x2 (Blackwell FFMA2 / FADD2 / FMUL2)
// Packed values live in 64-bit registers end to end (inline PTX on .b64 operands): going through
// float2 made ptxas re-pack the halves around every FFMA2 (28 IMAD.MOV per sample, profiles/).
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk_make(float lo, float hi) {
    pk2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ pk2 pk_dup(float a) { return pk_make(a, a); }
__device__ __forceinline__ void pk_split(pk2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) {
    pk2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b) {
    pk2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// in-place forms (destination tied to the addend / multiplicand): keep the sample array in fixed
// registers across the section loop -- otherwise ptxas computes into fresh pairs and copies back
#ifdef GFX_PK_VOLATILE
#define GFX_PK_ASM asm volatile
#else
#define GFX_PK_ASM asm
#endif
__device__ __forceinline__ void pk_fma_acc(pk2& c, pk2 a, pk2 b) {
    GFX_PK_ASM("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ void pk_mul_acc(pk2& c, pk2 a) { GFX_PK_ASM("mul.rn.f32x2 %0, %1, %0;" : "+l"(c) : "l"(a)); }
__device__ __forceinline__ pk2 pk_shfl_up(pk2 v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b) {
    pk2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pk2 pk_sub(pk2 a, pk2 b) {
    pk2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pk2 pk_swap(pk2 v) {
    float lo, hi;
    pk_split(v, lo, hi);
    return pk_make(hi, lo);
}

// ---------------------------------------------------------------- ordered-chain primitives
// Tiles of one row form a dependency chain (tile t needs the filter state left by tile t-1).
// Work items are handed out by an atomic ticket in tile-major order, so the item a CTA waits
// on always has a smaller ticket and is owned by a CTA that is already running: no deadlock,
// no co-residency requirement, perfect load balance for any rows x tiles.
__device__ __forceinline__ unsigned int take_ticket(unsigned int* ctr, unsigned int wrap_at) {
    // wraps back to 0 after exactly (n_items + gridDim.x) increments => self-resetting
    return atomicInc(ctr, wrap_at);
}
__device__ __forceinline__ void chain_wait(const int* flag, int needed) {
    volatile const int* vf = flag;
    while (*vf < needed) { __nanosleep(32); }
    __threadfence();
}
__device__ __forceinline__ void chain_publish(int* flag, int value) {
    // release: the state words written by this thread before the call are visible to whoever sees the flag
    __threadfence();
    *reinterpret_cast<volatile int*>(flag) = value;
}

// ---------------------------------------------------------------- 128-byte-row XOR swizzle
// A tile is stored in shared memory as rows of 128 bytes (one row per thread = that thread's
// contiguous chunk of samples).  16-byte unit c of row r lives at unit (c ^ (r & 7)): the
// coalesced tile load/store (consecutive threads -> consecutive units) and the per-thread
// chunk access (thread r reads units 0..7 of row r) are both bank-conflict free.
__device__ __forceinline__ int swz_unit(int row, int unit) { return (row << 3) | (unit ^ (row & 7)); }

}  // namespace gfx
