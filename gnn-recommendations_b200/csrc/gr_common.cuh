// Shared device/host helpers for libgr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gr_b200.h"

namespace gr {

constexpr int kSmCountFallback = 148;

// --- error plumbing -----------------------------------------------------------------------
void set_last_cuda_error(cudaError_t e);
int sm_count();

#define GR_CUDA_CHECK(expr)                       \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) {                  \
            ::gr::set_last_cuda_error(_e);        \
            return GR_ERR_CUDA;                   \
        }                                         \
    } while (0)

#define GR_LAUNCH_CHECK() GR_CUDA_CHECK(cudaPeekAtLastError())

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// --- loads ----------------------------------------------------------------------------------
// L2 eviction policies.  CSR structure and values are read once per launch: first-to-evict,
// not allocated in L1.  Gathered embedding rows are re-read (hot item rows by many user rows):
// last-to-evict.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_stream_i32(const int *p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_gather_f4(const float4 *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// acc = v * x + acc, one rounding per element (== fmaf), never re-associated.
__device__ __forceinline__ void fma4(float4 &acc, float v, const float4 &x) {
    acc.x = __fmaf_rn(v, x.x, acc.x);
    acc.y = __fmaf_rn(v, x.y, acc.y);
    acc.z = __fmaf_rn(v, x.z, acc.z);
    acc.w = __fmaf_rn(v, x.w, acc.w);
}

__device__ __forceinline__ float apply_scale(float v, float scale, int mode) {
    if (mode == GR_SCALE_MUL) return __fmul_rn(v, scale);
    if (mode == GR_SCALE_DIV) return __fdiv_rn(v, scale);
    return v;
}

// --- counter-based dropout bits ------------------------------------------------------------------
// splitmix64 stream element `idx` of `seed`: four independent 16-bit uniforms per call.  Forward and
// backward kernels derive the same mask from (seed, element index); nothing is stored.
__host__ __device__ __forceinline__ unsigned long long drop_bits(unsigned long long seed, unsigned long long idx) {
    unsigned long long x = seed + (idx + 1ULL) * 0x9E3779B97F4A7C15ULL;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ bool drop_keep(unsigned long long bits, int lane4, unsigned thr) {
    return ((unsigned)(bits >> (16 * lane4)) & 0xFFFFu) >= thr;
}
inline unsigned drop_threshold(float p) { return p > 0.f ? (unsigned)(p * 65536.f + 0.5f) : 0u; }
inline float drop_scale_of(unsigned thr) { return 65536.f / (float)(65536u - thr); }

// --- cp.async (LDGSTS) -----------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, uint64_t pol) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src, uint64_t pol) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem_src), "l"(pol)
                 : "memory");
}
// hint-free variants (the persistent long-row kernel: ptxas keeps the policy descriptor in uniform
// registers, and the looped kernel faulted with "illegal instruction" on LDGSTS with a hint)
__device__ __forceinline__ void cp_async16_plain(void *smem_dst, const void *gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4_plain(void *smem_dst, const void *gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// --- shared launch helpers (layers_bwd.cu) --------------------------------------------------------
int persistent_grid(int work_items, int per_sm);      // min(work_items, SMs * per_sm), >= 1
// out[e] = sum_p partial[p][e] (p ascending, double accumulator); returns non-zero on launch failure
int launch_reduce_partials(const float *partial, int n_part, long long total, float *out, cudaStream_t st);

// --- device radix sort (graph.cu) -------------------------------------------------------------
// LSD radix sort of 64-bit keys (optionally carrying a 32-bit payload), 8 bits per pass,
// stable.  Sorted data ends in keys_a/payload_a when `*in_a_out` is true, else in *_b.
size_t radix_sort_workspace_bytes(int64_t n);
int radix_sort_u64(uint64_t *keys_a, uint64_t *keys_b, uint32_t *pay_a, uint32_t *pay_b, int64_t n, int begin_bit,
                   int end_bit, void *ws, size_t ws_bytes, bool *in_a_out, cudaStream_t stream);

}  // namespace gr
