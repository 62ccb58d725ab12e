// Reduce + broadcast of partial item rows over NVLink peer memory — the exchange step of the
// user-owner ("1.5-D") multi-GPU propagation (no reference counterpart: the reference is single-process;
// parity target is the single-GPU result, 1e-5 for embeddings per BASELINE.json north_star).
//
// Layout: users are owned by ranks (their rows never leave the rank); every rank computes PARTIAL sums of ALL
// item rows from its own users (P_g = A_items,users(g) X_users(g)) and pushes each row from the SpMM epilogue to
// the rank owning the row's item block (routed peer stores of gr_spmm_csr_f32, one staging slot per sender).  For
// each row j of its block the owner then
//     loads slot_0[j] .. slot_{G-1}[j]  (local HBM; a first version PULLED them over NVLink: peer loads ran at
//                                        370 GB/s against 650 GB/s for peer stores on 8 B200s, hence the push)
//     v = ((slot_0[j] + slot_1[j]) + ...) (fixed rank order: deterministic, identical on every run)
//     stores v into row j of every rank's item table  (G-1 peer stores)   [skipped on the last layer]
//     folds v into the running layer sum of its own block   (out = scale_op(addend + v))
// Per layer a rank sends 2 x (G-1)/G x I x 4d bytes instead of the (G-1)/G x N x 4d of the exact all-gather:
// 4.5 GB instead of 11.2 GB at C5 on 8 GPUs.
#include <cstdlib>

#include "gr_common.cuh"

namespace gr {

constexpr int kMaxRanks = 8;

struct ReduceBcastArgs {
    const float4 *src[kMaxRanks];   // partial buffers, row-major [*, ld4]
    float4 *dst[kMaxRanks];         // item tables of all ranks (n_dst may be 0)
    int n_src, n_dst;
    long long ld_src4, ld_dst4;
    long long row0, dst_row0, n_rows;   // source rows [row0, row0 + n_rows) -> table rows [dst_row0, dst_row0 + n_rows)
    int f4;                         // d / 4
    const float4 *addend;           // [n_rows, lda4] running layer sum of the block (optional)
    float4 *out;                    // [n_rows, ldo4]
    float4 *own;                    // optional: the reduced rows, local copy [n_rows, ldw4]
    long long lda4, ldo4, ldw4;
    float scale;
    int scale_mode;
};

// Partial rows are written by other GPUs (peer stores) or by an earlier kernel of this GPU; a system-scope
// relaxed load never hits a stale L1 line.
__device__ __forceinline__ float4 ld_peer_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

template <int UNROLL>
__global__ void __launch_bounds__(256) reduce_bcast_rows_kernel(const ReduceBcastArgs a) {
    const long long total = a.n_rows * a.f4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * UNROLL) {
        float4 v[UNROLL];
        long long idx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) idx[u] = base + (long long)u * stride;
        // issue every peer load of the batch before the first add: G x UNROLL requests in flight per thread
        float4 part[UNROLL][kMaxRanks];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (idx[u] < total) {
                const long long r = idx[u] / a.f4;
                const int f = (int)(idx[u] % a.f4);
#pragma unroll
                for (int g = 0; g < kMaxRanks; ++g)
                    if (g < a.n_src) part[u][g] = ld_peer_f4(a.src[g] + (a.row0 + r) * a.ld_src4 + f);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (idx[u] < total) {
                float4 s = part[u][0];
#pragma unroll
                for (int g = 1; g < kMaxRanks; ++g)
                    if (g < a.n_src) {
                        s.x = __fadd_rn(s.x, part[u][g].x); s.y = __fadd_rn(s.y, part[u][g].y);
                        s.z = __fadd_rn(s.z, part[u][g].z); s.w = __fadd_rn(s.w, part[u][g].w);
                    }
                v[u] = s;
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (idx[u] < total) {
                const long long r = idx[u] / a.f4;
                const int f = (int)(idx[u] % a.f4);
#pragma unroll
                for (int g = 0; g < kMaxRanks; ++g)
                    if (g < a.n_dst) a.dst[g][(a.dst_row0 + r) * a.ld_dst4 + f] = v[u];
                if (a.own) a.own[r * a.ldw4 + f] = v[u];
                if (a.out) {
                    float4 o = v[u];
                    if (a.addend) {
                        const float4 ad = a.addend[r * a.lda4 + f];
                        o.x = __fadd_rn(ad.x, o.x); o.y = __fadd_rn(ad.y, o.y);
                        o.z = __fadd_rn(ad.z, o.z); o.w = __fadd_rn(ad.w, o.w);
                    }
                    o.x = apply_scale(o.x, a.scale, a.scale_mode); o.y = apply_scale(o.y, a.scale, a.scale_mode);
                    o.z = apply_scale(o.z, a.scale, a.scale_mode); o.w = apply_scale(o.w, a.scale, a.scale_mode);
                    a.out[r * a.ldo4 + f] = o;
                }
            }
        }
    }
}

}  // namespace gr

using namespace gr;

extern "C" int gr_reduce_bcast_rows(const float *const *src_host, int32_t n_src, int64_t ld_src,
                                    float *const *dst_host, int32_t n_dst, int64_t ld_dst, int64_t row0,
                                    int64_t dst_row0, int64_t n_rows, int32_t d, const float *addend, int64_t lda, float *out,
                                    int64_t ldo, float *own, int64_t ldw, float scale, int32_t scale_mode,
                                    void *stream) {
    if (!src_host || n_src < 1 || n_src > kMaxRanks || n_dst < 0 || n_dst > kMaxRanks || (n_dst > 0 && !dst_host))
        return GR_ERR_INVALID;
    if (n_rows < 0 || row0 < 0 || dst_row0 < 0 || d <= 0 || (d & 3) || (ld_src & 3) || ld_src < d) return GR_ERR_INVALID;
    if (n_dst > 0 && ((ld_dst & 3) || ld_dst < d)) return GR_ERR_INVALID;
    if (addend && (!out || (lda & 3) || lda < d)) return GR_ERR_INVALID;
    if (out && ((ldo & 3) || ldo < d)) return GR_ERR_INVALID;
    if (own && ((ldw & 3) || ldw < d)) return GR_ERR_INVALID;
    if (scale_mode < GR_SCALE_NONE || scale_mode > GR_SCALE_DIV) return GR_ERR_INVALID;
    if (!aligned16(addend) || !aligned16(out) || !aligned16(own)) return GR_ERR_INVALID;
    if (n_rows == 0) return GR_OK;
    ReduceBcastArgs a = {};
    for (int g = 0; g < n_src; ++g) {
        if (!src_host[g] || !aligned16(src_host[g])) return GR_ERR_INVALID;
        a.src[g] = reinterpret_cast<const float4 *>(src_host[g]);
    }
    for (int g = 0; g < n_dst; ++g) {
        if (!dst_host[g] || !aligned16(dst_host[g])) return GR_ERR_INVALID;
        a.dst[g] = reinterpret_cast<float4 *>(dst_host[g]);
    }
    a.n_src = n_src; a.n_dst = n_dst; a.ld_src4 = ld_src / 4; a.ld_dst4 = ld_dst / 4;
    a.row0 = row0; a.dst_row0 = dst_row0; a.n_rows = n_rows; a.f4 = d / 4;
    a.addend = reinterpret_cast<const float4 *>(addend); a.out = reinterpret_cast<float4 *>(out);
    a.own = reinterpret_cast<float4 *>(own);
    a.lda4 = lda / 4; a.ldo4 = ldo / 4; a.ldw4 = ldw / 4;
    a.scale = scale; a.scale_mode = scale_mode;
    const long long total = n_rows * (d / 4);
    long long blocks = (total + 256 * 2 - 1) / (256 * 2);
    // NVLink-bound, and meant to run BESIDE the user-row SpMM: a persistent grid that filled every thread slot
    // (8 CTAs per SM) kept the SpMM's CTAs out until it retired (measured: no overlap at all).  One CTA per SM
    // keeps ~19 MB of peer stores in flight, far more than the links need.
    static const int per_sm = [] { const char *e = getenv("GR_REDUCE_CTAS_PER_SM"); int v = e ? atoi(e) : 1; return v > 0 ? v : 1; }();
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    reduce_bcast_rows_kernel<2><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

namespace gr {
struct MultiCopyArgs {
    float4 *dst[kMaxRanks];
    const float4 *src[kMaxRanks];
    long long n4[kMaxRanks];     // float4 elements per segment
    int n_seg, parts;            // CTAs per segment
};

// Up to 8 independent copies (local -> local or peer-mapped memory) in ONE small launch: CTA b works on part
// b / n_seg of segment b % n_seg, 4 independent 16-byte loads in flight per thread, plain stores (posted writes
// over NVLink).  ~30 registers and a few dozen CTAs: it runs beside an SpMM without taking its occupancy.
__global__ void __launch_bounds__(256) peer_copy_multi_kernel(const MultiCopyArgs a) {
    const int seg = blockIdx.x % a.n_seg, part = blockIdx.x / a.n_seg;
    const long long n = a.n4[seg];
    const long long per = (n + a.parts - 1) / a.parts;
    const long long begin = per * part, end = min(n, begin + per);
    const float4 *src = a.src[seg];
    float4 *dst = a.dst[seg];
    long long i = begin + threadIdx.x;
    for (; i + 3 * 256 < end; i += 4 * 256) {
        const float4 v0 = __ldg(src + i), v1 = __ldg(src + i + 256), v2 = __ldg(src + i + 512), v3 = __ldg(src + i + 768);
        dst[i] = v0; dst[i + 256] = v1; dst[i + 512] = v2; dst[i + 768] = v3;
    }
    for (; i < end; i += 256) dst[i] = __ldg(src + i);
}
}  // namespace gr

extern "C" int gr_peer_copy_multi(void *const *dst_host, const void *const *src_host, const size_t *bytes_host,
                                  int32_t n_seg, int32_t ctas, void *stream) {
    if (!dst_host || !src_host || !bytes_host || n_seg < 1 || n_seg > kMaxRanks) return GR_ERR_INVALID;
    MultiCopyArgs a = {};
    int n = 0;
    for (int s = 0; s < n_seg; ++s) {
        if (bytes_host[s] == 0) continue;
        if (!dst_host[s] || !src_host[s] || (bytes_host[s] & 15) || !aligned16(dst_host[s]) || !aligned16(src_host[s]))
            return GR_ERR_INVALID;
        a.dst[n] = static_cast<float4 *>(dst_host[s]);
        a.src[n] = static_cast<const float4 *>(src_host[s]);
        a.n4[n] = (long long)(bytes_host[s] / 16);
        ++n;
    }
    if (n == 0) return GR_OK;
    a.n_seg = n;
    if (ctas < n) ctas = n;
    a.parts = ctas / n;
    peer_copy_multi_kernel<<<a.parts * n, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

/* Asynchronous copy between device buffers that may live on different GPUs of the box (peer-mapped symmetric
 * memory): plain cudaMemcpyAsync, i.e. the COPY ENGINES move the bytes over NVLink and no SM is involved — the
 * user-owner propagation sends its partial item blocks and broadcasts the reduced blocks this way, beside the
 * SpMM kernels (an SM kernel doing the same stores took a third of the register file of every SM it ran on and
 * slowed the HBM-bound SpMM by as much as it saved). */
extern "C" int gr_peer_copy_async(void *dst, const void *src, size_t bytes, void *stream) {
    if ((!dst || !src) && bytes) return GR_ERR_INVALID;
    if (bytes == 0) return GR_OK;
    GR_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return GR_OK;
}
