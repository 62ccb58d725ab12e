// CSR SpMM for the propagation step, fp32, order-preserving.
//
//   t[r,:] = sum over the row's entries k (storage order) of vals[k] * x[indices[k],:]
//
// evaluated as one fmaf chain per (row, feature) from a zero accumulator — the order the
// reference's CPU torch.sparse.mm uses (lightgcn.py:88), so the result is bit-identical.
// Parallelism therefore comes from rows and features only, never from splitting a row's
// entries:
//   * short rows: one row per sub-warp (d/4 lanes, one float4 per lane), 8 independent
//     128-bit gathers in flight per lane, (col,val) pairs loaded coalesced once per 8..32
//     entries and broadcast with shuffles, next chunk prefetched;
//   * long ("hot item") rows: one CTA per row; all 8 warps stream the gathered x rows into
//     an 8-stage shared-memory ring with cp.async (up to 112 KB in flight per SM), d/32
//     consumer warps run the chain out of shared memory, one feature per lane.
// Rows arrive sorted by descending length (gr_row_schedule) so the hardware CTA scheduler
// performs longest-processing-time-first list scheduling; the hot rows start first on a
// high-priority side stream and overlap the short-row kernel.
#include <mutex>

#include "gr_common.cuh"

namespace gr {

struct SpmmArgs {
    const int *indptr;
    const int *indices;
    const float *vals;
    const int *row_order;
    int order_begin;  // first position of row_order handled by this launch
    int order_end;    // one past the last position
    const float4 *x;
    long long ldx4;
    float4 *y;
    long long ldy4;
    const float4 *addend;
    long long lda4;
    float4 *out;
    long long ldo4;
    float scale;
    int scale_mode;
};

template <int D>
struct RowCfg {
    static constexpr int F4 = D / 4;
    static constexpr int LPR = F4 < 32 ? F4 : 32;  // lanes per row
    static constexpr int VPL = F4 / LPR;           // float4 per lane
    static constexpr int RPW = 32 / LPR;           // rows per warp
    static constexpr int UNROLL = VPL == 1 ? 8 : 4;
};

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 scale4(const float4 &a, float s, int mode) {
    return make_float4(apply_scale(a.x, s, mode), apply_scale(a.y, s, mode), apply_scale(a.z, s, mode),
                       apply_scale(a.w, s, mode));
}

// ---------------------------------------------------------------------------------------------
// short rows: one row per sub-warp
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kWarpsPerCta * 32) spmm_warp_rows(const SpmmArgs a) {
    using C = RowCfg<D>;
    constexpr int LPR = C::LPR, VPL = C::VPL, RPW = C::RPW, UNROLL = C::UNROLL;
    constexpr unsigned kFull = 0xffffffffu;

    const uint64_t pol_s = policy_evict_first(), pol_g = policy_evict_last();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane % LPR;   // lane within the row's group
    const int grp = lane / LPR;  // which of the warp's rows
    const long long tile = (long long)blockIdx.x * kWarpsPerCta + warp;
    const long long pos = (long long)a.order_begin + tile * RPW + grp;

    int r = -1, start = 0, len = 0;
    if (pos < a.order_end) {
        r = a.row_order ? a.row_order[pos] : (int)pos;
        start = a.indptr[r];
        len = a.indptr[r + 1] - start;
    }
    int maxlen = len;
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, o));

    float4 acc[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int *ci = a.indices + start;
    const float *cv = a.vals + start;
    int c_cur = 0;
    float v_cur = 0.f;
    if (gl < len) {
        c_cur = ld_stream_i32(ci + gl, pol_s);
        v_cur = ld_stream_f32(cv + gl, pol_s);
    }
    for (int base = 0; base < maxlen; base += LPR) {
        int c_nxt = 0;
        float v_nxt = 0.f;
        const int nidx = base + LPR + gl;
        if (nidx < len) {
            c_nxt = ld_stream_i32(ci + nidx, pol_s);
            v_nxt = ld_stream_f32(cv + nidx, pol_s);
        }
        const int rem = len - base;                   // this row's remaining entries (may be <= 0)
        const int cnt = min(LPR, maxlen - base);      // warp-uniform
#pragma unroll
        for (int k = 0; k < LPR; k += UNROLL) {
            if (k >= cnt) break;
            float4 xv[UNROLL][VPL];
            float vv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = __shfl_sync(kFull, c_cur, k + u, LPR);
                vv[u] = __shfl_sync(kFull, v_cur, k + u, LPR);
                if (k + u < rem) {
                    const float4 *src = a.x + (long long)cc * a.ldx4 + gl;
#pragma unroll
                    for (int j = 0; j < VPL; ++j) xv[u][j] = ld_gather_f4(src + j * LPR, pol_g);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (k + u < rem) {
#pragma unroll
                    for (int j = 0; j < VPL; ++j) fma4(acc[j], vv[u], xv[u][j]);
                }
            }
        }
        c_cur = c_nxt;
        v_cur = v_nxt;
    }

    if (r >= 0) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int off = gl + j * LPR;
            if (a.y) st_stream_f4(a.y + (long long)r * a.ldy4 + off, acc[j]);
            if (a.out) {
                float4 o = acc[j];
                if (a.addend) o = add4(__ldg(a.addend + (long long)r * a.lda4 + off), o);
                st_stream_f4(a.out + (long long)r * a.ldo4 + off, scale4(o, a.scale, a.scale_mode));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// long rows: one CTA per row, cp.async ring in shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 32;  // neighbours per ring stage

template <int D>
struct LongCfg {
    static constexpr int STAGES = D <= 128 ? 8 : 6;
    static constexpr int STAGE_FLOATS = kChunk * D;
    static constexpr size_t SMEM = (size_t)STAGES * (STAGE_FLOATS * 4 + kChunk * 4);
};

template <int D>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1) spmm_long_rows(const SpmmArgs a) {
    using L = LongCfg<D>;
    constexpr int STAGES = L::STAGES;
    constexpr int F4 = D / 4;
    constexpr int CONS = D / 32;                      // consumer warps, one feature per lane
    constexpr int NB_PER_WARP = kChunk / kWarpsPerCta;  // 4 neighbours per warp per stage
    constexpr int ITEMS = (NB_PER_WARP * F4) / 32;    // 16-byte copies per lane per stage
    constexpr unsigned kFull = 0xffffffffu;
    static_assert(CONS >= 1 && CONS <= kWarpsPerCta, "feature dim");
    static_assert(ITEMS >= 1, "feature dim");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *xs = reinterpret_cast<float4 *>(smem_raw);                            // [STAGES][32][F4]
    float *vs = reinterpret_cast<float *>(smem_raw + (size_t)STAGES * L::STAGE_FLOATS * 4);  // [STAGES][32]

    const uint64_t pol_s = policy_evict_first(), pol_g = policy_evict_last();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int r = a.row_order[a.order_begin + blockIdx.x];
    const int start = a.indptr[r];
    const int len = a.indptr[r + 1] - start;
    const int nchunks = (len + kChunk - 1) / kChunk;
    const int *ci = a.indices + start;
    const float *cv = a.vals + start;

    auto fetch = [&](int chunk, int &col, float &val) {
        col = -1;
        val = 0.f;
        if (lane < NB_PER_WARP) {
            const int idx = chunk * kChunk + warp * NB_PER_WARP + lane;
            if (idx < len) {
                col = ld_stream_i32(ci + idx, pol_s);
                val = ld_stream_f32(cv + idx, pol_s);
            }
        }
    };
    auto issue = [&](int chunk, int col, float val) {
        if (chunk < nchunks) {
            const int stage = chunk % STAGES;
            if (lane < NB_PER_WARP) vs[stage * kChunk + warp * NB_PER_WARP + lane] = val;
#pragma unroll
            for (int t = 0; t < ITEMS; ++t) {
                const int item = lane + 32 * t;
                const int nb = item / F4;
                const int f = item % F4;
                const int cc = __shfl_sync(kFull, col, nb);
                if (cc >= 0)
                    cp_async16(xs + ((size_t)stage * kChunk + warp * NB_PER_WARP + nb) * F4 + f,
                               a.x + (long long)cc * a.ldx4 + f, pol_g);
            }
        }
        cp_async_commit();
    };

    // prologue: fetch the (col,val) of the first STAGES-1 chunks together, then fill the ring
    int pc[STAGES - 1];
    float pv[STAGES - 1];
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) fetch(s, pc[s], pv[s]);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s, pc[s], pv[s]);
    int ncol;
    float nval;
    fetch(STAGES - 1, ncol, nval);

    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<STAGES - 2>();  // this thread's copies of chunk c have landed
        __syncthreads();              // everyone's have; everyone is done reading chunk c-1
        issue(c + STAGES - 1, ncol, nval);  // refills the stage chunk c-1 used
        fetch(c + STAGES, ncol, nval);
        if (warp < CONS) {
            const int stage = c % STAGES;
            const float *xr = reinterpret_cast<const float *>(xs + (size_t)stage * kChunk * F4) + warp * 32 + lane;
            const float *vr = vs + stage * kChunk;
            const int n = min(kChunk, len - c * kChunk);
            if (n == kChunk) {
#pragma unroll
                for (int k = 0; k < kChunk; ++k) acc = __fmaf_rn(vr[k], xr[k * D], acc);
            } else {
                for (int k = 0; k < n; ++k) acc = __fmaf_rn(vr[k], xr[k * D], acc);
            }
        }
    }
    cp_async_wait<0>();

    if (warp < CONS) {
        const int f = warp * 32 + lane;
        if (a.y) reinterpret_cast<float *>(a.y + (long long)r * a.ldy4)[f] = acc;
        if (a.out) {
            float o = acc;
            if (a.addend) o = __fadd_rn(reinterpret_cast<const float *>(a.addend + (long long)r * a.lda4)[f], o);
            reinterpret_cast<float *>(a.out + (long long)r * a.ldo4)[f] = apply_scale(o, a.scale, a.scale_mode);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool smem_attr_set[4] = {false, false, false, false};
};

static std::mutex g_side_mu;
static SideStream g_side[64];

static int get_side(SideStream **out) {
    int dev = 0;
    GR_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(g_side_mu);
    SideStream &s = g_side[dev];
    if (!s.stream) {
        int lo = 0, hi = 0;
        GR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GR_CUDA_CHECK(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, hi));
        GR_CUDA_CHECK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        GR_CUDA_CHECK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    *out = &s;
    return GR_OK;
}

template <int D>
static int launch(const SpmmArgs &base, int n_long, long long n_rows, int slot, cudaStream_t stream) {
    using C = RowCfg<D>;
    using L = LongCfg<D>;
    SideStream *side = nullptr;
    const bool use_long = base.row_order != nullptr && n_long > 0;
    if (use_long) {
        int rc = get_side(&side);
        if (rc != GR_OK) return rc;
        if (!side->smem_attr_set[slot]) {
            GR_CUDA_CHECK(cudaFuncSetAttribute(spmm_long_rows<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)L::SMEM));
            side->smem_attr_set[slot] = true;
        }
        GR_CUDA_CHECK(cudaEventRecord(side->fork, stream));
        GR_CUDA_CHECK(cudaStreamWaitEvent(side->stream, side->fork, 0));
        SpmmArgs la = base;
        la.order_begin = 0;
        la.order_end = n_long;
        spmm_long_rows<D><<<n_long, kWarpsPerCta * 32, L::SMEM, side->stream>>>(la);
        GR_LAUNCH_CHECK();
        GR_CUDA_CHECK(cudaEventRecord(side->join, side->stream));
    }
    const long long rest = n_rows - (use_long ? n_long : 0);
    if (rest > 0) {
        SpmmArgs wa = base;
        wa.order_begin = use_long ? n_long : 0;
        wa.order_end = (int)n_rows;
        const long long tiles = (rest + C::RPW - 1) / C::RPW;
        const long long ctas = (tiles + kWarpsPerCta - 1) / kWarpsPerCta;
        spmm_warp_rows<D><<<(unsigned)ctas, kWarpsPerCta * 32, 0, stream>>>(wa);
        GR_LAUNCH_CHECK();
    }
    if (use_long) GR_CUDA_CHECK(cudaStreamWaitEvent(stream, side->join, 0));
    return GR_OK;
}

}  // namespace gr

extern "C" int gr_spmm_csr_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                               const int32_t *row_order, int32_t n_long, int64_t n_rows, int32_t d, const float *x,
                               int64_t ldx, float *y, int64_t ldy, const float *addend, int64_t lda, float *out,
                               int64_t ldo, float scale, int32_t scale_mode, void *stream) {
    using namespace gr;
    if (n_rows == 0) return GR_OK;
    if (!indptr || !indices || !vals || !x || n_rows < 0 || n_long < 0 || n_long > n_rows) return GR_ERR_INVALID;
    if (!y && !out) return GR_ERR_INVALID;
    if (n_rows > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if (scale_mode < GR_SCALE_NONE || scale_mode > GR_SCALE_DIV) return GR_ERR_INVALID;
    if ((ldx & 3) || (y && (ldy & 3)) || (addend && (lda & 3)) || (out && (ldo & 3))) return GR_ERR_INVALID;
    if (ldx < d || (y && ldy < d) || (addend && lda < d) || (out && ldo < d)) return GR_ERR_INVALID;
    if (!aligned16(x) || !aligned16(y) || !aligned16(addend) || !aligned16(out)) return GR_ERR_INVALID;
    SpmmArgs a;
    a.indptr = indptr;
    a.indices = indices;
    a.vals = vals;
    a.row_order = row_order;
    a.order_begin = 0;
    a.order_end = (int)n_rows;
    a.x = reinterpret_cast<const float4 *>(x);
    a.ldx4 = ldx / 4;
    a.y = reinterpret_cast<float4 *>(y);
    a.ldy4 = ldy / 4;
    a.addend = reinterpret_cast<const float4 *>(addend);
    a.lda4 = lda / 4;
    a.out = reinterpret_cast<float4 *>(out);
    a.ldo4 = ldo / 4;
    a.scale = scale;
    a.scale_mode = scale_mode;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
        case 32: return launch<32>(a, n_long, n_rows, 0, s);
        case 64: return launch<64>(a, n_long, n_rows, 1, s);
        case 128: return launch<128>(a, n_long, n_rows, 2, s);
        case 256: return launch<256>(a, n_long, n_rows, 3, s);
        default: return GR_ERR_UNSUPPORTED;
    }
}
