// CSR SpMM for the propagation step, fp32, order-preserving.
//
//   t[r,:] = sum over the row's entries k (storage order) of vals[k] * x[indices[k],:]
//
// evaluated as one fmaf chain per (row, feature) from a zero accumulator — the order the
// reference's CPU torch.sparse.mm uses (lightgcn.py:88), so the result is bit-identical.
// Parallelism therefore comes from rows and features only, never from splitting a row's
// entries:
//   * short rows (spmm_stream_rows): groups of consecutive rows (~256-512 entries) per sub-warp of
//     d/4 lanes, walked as one continuous entry stream: (col,val) chunks loaded coalesced and
//     shuffle-broadcast, 8 independent 128-bit gathers in flight per lane across row boundaries;
//   * long ("hot item") rows (spmm_long_rows / spmm_long_rows_bar, >= 1024 entries): persistent CTAs
//     take rows longest-first from a ticket counter; producer warps stream the gathered x rows into
//     a shared-memory ring with cp.async, d/32 consumer warps run the chain out of shared memory,
//     one feature per lane.  They run on a high-priority side stream and overlap the short-row kernel.
//   * rows beyond 131 072 entries are cut into 65 536-entry segments whose partial chains are added
//     in segment order (spmm_combine_parts) — deterministic, identical on 1 and N GPUs.
// Dense-map epilogue (gr_spmm_csr_map_f32, d <= 64): the finished row t is multiplied by a d x d matrix held
// in shared memory and combined with a residual row before it is written,
//   out[r,:] = alpha * (t M) + beta * addend[r,:]      (y[r,:] = alpha * t, optional)
// which is one Group-and-Shuffle layer (model.py:171-195: c = A x; (c W_conn) W_orth[:, perm]; residual) — and,
// with M^T, its backward — in ONE pass: no second launch, no N x d round trip of c through HBM.
// The epilogue of every kernel can also store the finished row into the layer buffers of the other
// ranks (NVLink P2P stores or one NVSwitch multicast store): the all-gather of the row-partitioned
// propagation is fused into the SpMM.
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "gr_common.cuh"

namespace gr {

constexpr int kMaxPeers = 8;

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
    const int *group_ptr;  // streaming kernel: rows [group_ptr[g], group_ptr[g+1]) per sub-warp
    int n_groups;
    int long_thr;          // rows with >= long_thr entries are left to the long-row kernel
    // long-row work items (optional): [4][n_items] = row, offset into the row, length, partial slot
    // (-1: the item is a whole row).  Rows above the split threshold are cut into segments whose
    // partial chains land in part_buf and are added in segment order by spmm_combine_parts.
    const int *items;
    int n_items;
    const int *split_rows;  // [3][n_split] = row, first partial slot, number of partials
    int n_split;
    float *part_buf;        // [n_partials][D]
    // fused all-gather: t[r,:] is also stored to row (peer_row_off + r) of every peer's gathered
    // buffer (NVLink P2P stores from the epilogue; ld = ldy4).  n_peers = 0 disables it.
    float4 *peer_y[kMaxPeers];
    int n_peers;
    int peer_multicast;     // peer_y[0] is a multicast address
    long long peer_row_off;
    // routed mode (route_block > 0): row r goes to ONE peer, g = r / route_block, as row
    // peer_row_off + r % route_block of its buffer (user-owner propagation: partial item rows are pushed to
    // the rank that owns the item block, into the slot of the sending rank)
    int route_block;
    // long-row scheduler state in CALLER memory: [0] ticket counter, [1] finished-CTA counter.  Zero on
    // entry; the last CTA of a launch zeroes both again.  (No library-global device state: two
    // propagations in flight on different streams use different words.)
    unsigned int *sched;
    // dense-map epilogue (map != nullptr): out = map_alpha * (t @ map) + beta * addend, y = map_alpha * t;
    // beta = map_beta * (*map_beta_dev if given).  map is [D][D] row-major; map_transposed applies map^T.
    const float *map;
    int map_transposed;
    float map_alpha, map_beta;
    const float *map_beta_dev;
};

// the d x d epilogue map into shared memory as Ms[i * D + j] = M[i][j] (or M[j][i]): out_j = sum_i t_i Ms[i][j]
template <int D>
__device__ __forceinline__ void load_map_smem(float *Ms, const SpmmArgs &a, int tid, int nthreads) {
    for (int i = tid; i < D * D; i += nthreads) Ms[i] = a.map_transposed ? __ldg(a.map + (i % D) * D + i / D) : __ldg(a.map + i);
}
__device__ __forceinline__ float map_beta_of(const SpmmArgs &a) {
    return a.map_beta * (a.map_beta_dev ? __ldg(a.map_beta_dev) : 1.f);
}

// kernel template parameter PEERS: how the finished row also leaves the GPU
constexpr int kPeersNone = 0;       // not at all
constexpr int kPeersP2P = 1;        // one store per rank to its mapped buffer
constexpr int kPeersMulticast = 2;  // one multimem.st to an NVSwitch multicast address (peer_y[0]): the switch
                                    // replicates it to every rank, NVLink egress is 1x instead of (G-1)x

__device__ __forceinline__ void st_multimem_f4(float4 *p, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_multimem_f1(float *p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

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
// short rows, streaming: one GROUP of consecutive rows (about group_nnz entries) per sub-warp
//
// The sub-warp walks the group's entries as ONE contiguous stream: (col,val) chunks are loaded
// coalesced and prefetched one chunk ahead, gathers are issued UNROLL at a time without regard to
// row boundaries (no partial batches, no per-row pipeline restart, no per-row dependent
// row_order -> indptr -> indices load chain).  Row boundaries only matter to the fmaf chain: when
// the stream crosses indptr[r+1] the finished row is written (epilogue fused) and the accumulator
// restarts from zero — each (row, feature) is still the exact storage-order chain.  indptr[r+2]
// and the addend row are prefetched at the previous boundary.  Rows with >= long_thr entries
// belong to the long-row kernel: the stream jumps over them.
// ---------------------------------------------------------------------------------------------
// MAP variant (dense-map epilogue): a finished row is not mapped at once — one row against the d x d map re-reads
// the whole map from shared memory (16 KB per row: measured 2x slower than the separate rowmap kernel) and is a
// 64-deep dependent fmaf chain in the middle of the gather stream.  The sub-warp STASHES the row in shared memory
// instead and the map is applied to four stashed rows at a time (4 rows x 4 columns per lane: every map element
// read from shared memory feeds four rows, sixteen independent chains per lane), at the converged point of the
// warp's batch loop so that both sub-warps run the dense loop together.
constexpr int kMapRows = 4;     // rows per sub-warp stash

// out = alpha (t M) + beta R for the `count` (<= kMapRows) rows of one sub-warp's stash; lane gl forms columns
// 4 gl .. 4 gl + 3 of all of them.  Deliberately NOT inlined: inlined, its sixteen accumulators raised the register
// pressure of the gather loop until ptxas spilled the gathered rows themselves (STL / LDL in the hot loop, the
// kernel 2.5x slower); as a call the caller's state is saved once per batch of rows instead.
template <int D>
__device__ __noinline__ void map_stash_rows(const float4 *Ms4, const float4 *st, const int *rows, int count, int gl,
                                            unsigned sub, float4 *out, long long ldo4, const float4 *addend,
                                            long long lda4, float alpha, float beta) {
    constexpr int F4 = D / 4;
    __syncwarp(sub);                         // the stashed rows (written by the other lanes) are visible
    float4 o[kMapRows];
#pragma unroll
    for (int q = 0; q < kMapRows; ++q) o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int i4 = 0; i4 < F4; ++i4) {
        const float4 m0 = Ms4[(4 * i4 + 0) * F4 + gl], m1 = Ms4[(4 * i4 + 1) * F4 + gl];
        const float4 m2 = Ms4[(4 * i4 + 2) * F4 + gl], m3 = Ms4[(4 * i4 + 3) * F4 + gl];
#pragma unroll
        for (int q = 0; q < kMapRows; ++q) {
            const float4 tt = st[q * F4 + i4];
            fma4(o[q], tt.x, m0);
            fma4(o[q], tt.y, m1);
            fma4(o[q], tt.z, m2);
            fma4(o[q], tt.w, m3);
        }
    }
#pragma unroll
    for (int q = 0; q < kMapRows; ++q) {
        if (q < count) {
            const int row = rows[q];
            float4 res = scale4(o[q], alpha, GR_SCALE_MUL);
            if (addend) fma4(res, beta, __ldg(addend + (long long)row * lda4 + gl));
            st_stream_f4(out + (long long)row * ldo4 + gl, res);
        }
    }
    __syncwarp(sub);                         // before the stash is written again
}

template <int D, int PEERS, bool MAP = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MAP ? 2 : 3) spmm_stream_rows(const SpmmArgs a) {
    using C = RowCfg<D>;
    constexpr int LPR = C::LPR, VPL = C::VPL, SPW = C::RPW, UNROLL = C::UNROLL;
    constexpr unsigned kFull = 0xffffffffu;
    static_assert(!MAP || (VPL == 1 && PEERS == kPeersNone), "dense-map epilogue: d <= 128, single GPU");
    constexpr int F4 = D / 4;
    // dense-map epilogue: the map, kMapRows stashed rows per sub-warp and their row ids in shared memory
    extern __shared__ __align__(16) unsigned char stream_smem[];
    float4 *Ms4 = reinterpret_cast<float4 *>(stream_smem);                          // [D][F4]
    float4 *stash4 = Ms4 + (MAP ? D * F4 : 0);                                       // [warps * SPW][kMapRows][F4]
    int *stash_row = reinterpret_cast<int *>(stash4 + (MAP ? kWarpsPerCta * SPW * kMapRows * F4 : 0));
    float map_beta = 0.f;
    int scount = 0;                                                                  // rows in this sub-warp's stash
    if constexpr (MAP) {
        load_map_smem<D>(reinterpret_cast<float *>(Ms4), a, threadIdx.x, kWarpsPerCta * 32);
        map_beta = map_beta_of(a);
        __syncthreads();
    }

    const uint64_t pol_s = policy_evict_first(), pol_g = policy_evict_last();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane % LPR;
    const long long g = ((long long)blockIdx.x * kWarpsPerCta + warp) * SPW + lane / LPR;
    const int n_rows = a.order_end;

    int r = 0, r_end = 0, row_stop = 0, stop_next = 0, stream_end = 0, kc = 0, dead_until = 0;
    float4 acc[VPL], add_cur[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) acc[j] = add_cur[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_addend = [&](int row) {
        if constexpr (!MAP) {          // (the dense-map epilogue reads the residual rows when it writes a batch)
            if (a.addend) {
#pragma unroll
                for (int j = 0; j < VPL; ++j) add_cur[j] = __ldg(a.addend + (long long)row * a.lda4 + gl + j * LPR);
            }
        }
    };
    // dense-map epilogue of this sub-warp's stash (a call, not inlined: see map_stash_rows)
    auto map_stash = [&]() {
        if constexpr (MAP) {
            const unsigned sub = (LPR == 32) ? kFull : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
            const int sw = warp * SPW + lane / LPR;
            map_stash_rows<D>(Ms4, stash4 + sw * kMapRows * F4, stash_row + sw * kMapRows, scount, gl, sub, a.out, a.ldo4,
                              a.addend, a.lda4, a.map_alpha, map_beta);
            scount = 0;
        }
    };
    auto flush = [&](int row) {
        if constexpr (MAP) {
            if (scount == kMapRows) map_stash();     // rare: several rows ended inside one batch of entries
            const int sw = warp * SPW + lane / LPR;
            stash4[(sw * kMapRows + scount) * F4 + gl] = acc[0];
            if (gl == 0) stash_row[sw * kMapRows + scount] = row;
            ++scount;
            if (a.y) st_stream_f4(a.y + (long long)row * a.ldy4 + gl, scale4(acc[0], a.map_alpha, GR_SCALE_MUL));
            acc[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int off = gl + j * LPR;
            if (a.y) st_stream_f4(a.y + (long long)row * a.ldy4 + off, acc[j]);
            if constexpr (PEERS == kPeersP2P) {
                if (a.route_block > 0) {
                    const int pg = row / a.route_block;
                    a.peer_y[pg][(a.peer_row_off + (row - pg * a.route_block)) * a.ldy4 + off] = acc[j];
                } else {
#pragma unroll 1
                    for (int p = 0; p < a.n_peers; ++p) a.peer_y[p][(a.peer_row_off + row) * a.ldy4 + off] = acc[j];
                }
            } else if constexpr (PEERS == kPeersMulticast) {
                st_multimem_f4(a.peer_y[0] + (a.peer_row_off + row) * a.ldy4 + off, acc[j]);
            }
            if (a.out) {
                float4 o = acc[j];
                if (a.addend) o = add4(add_cur[j], o);
                st_stream_f4(a.out + (long long)row * a.ldo4 + off, scale4(o, a.scale, a.scale_mode));
            }
            acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        }
    };
    // move to row r+1 (its end offset comes from the prefetched stop_next); jump over long rows
    auto next_row = [&]() {
        int rs = row_stop;
        ++r;
        row_stop = stop_next;
        stop_next = (r + 2 <= n_rows) ? __ldg(a.indptr + r + 2) : row_stop;
        while (r < r_end && row_stop - rs >= a.long_thr) {
            dead_until = row_stop;
            rs = row_stop;
            ++r;
            row_stop = stop_next;
            stop_next = (r + 2 <= n_rows) ? __ldg(a.indptr + r + 2) : row_stop;
        }
        if (r < r_end) load_addend(r);
    };

    if (g < a.n_groups) {
        r = a.group_ptr[g];
        r_end = a.group_ptr[g + 1];
    }
    if (r < r_end) {
        stream_end = a.indptr[r_end];
        int rs = a.indptr[r];
        row_stop = a.indptr[r + 1];
        stop_next = (r + 2 <= n_rows) ? a.indptr[r + 2] : row_stop;
        while (r < r_end && row_stop - rs >= a.long_thr) {
            rs = row_stop;
            ++r;
            row_stop = stop_next;
            stop_next = (r + 2 <= n_rows) ? a.indptr[r + 2] : row_stop;
        }
        kc = rs;
        dead_until = rs;
        if (r < r_end) load_addend(r);
    }
    if (r >= r_end) {
        stream_end = 0;
        kc = 0;
    }

    int c_cur = 0;
    float v_cur = 0.f;
    if (kc + gl < stream_end) {
        c_cur = ld_stream_i32(a.indices + kc + gl, pol_s);
        v_cur = ld_stream_f32(a.vals + kc + gl, pol_s);
    }
    while (__any_sync(kFull, kc < stream_end)) {
        int c_nxt = 0;
        float v_nxt = 0.f;
        {
            const int idx = kc + LPR + gl;
            if (idx < stream_end) {
                c_nxt = ld_stream_i32(a.indices + idx, pol_s);
                v_nxt = ld_stream_f32(a.vals + idx, pol_s);
            }
        }
        // The batch loop is kept rolled and the row-boundary work (flush / next_row) appears once:
        // fully unrolled, this kernel overflowed the instruction cache (ncu: 2.2 no_instruction
        // stalls per issue at C4).
#pragma unroll 1
        for (int kk = 0; kk < LPR; kk += UNROLL) {
            if (!__any_sync(kFull, kc + kk < stream_end)) break;
            const int e0 = kc + kk;
            float4 xv[UNROLL][VPL];
            float vv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int cc = __shfl_sync(kFull, c_cur, kk + u, LPR);
                vv[u] = __shfl_sync(kFull, v_cur, kk + u, LPR);
                const int e = e0 + u;
                if (e < stream_end && e >= dead_until) {
                    const float4 *src = a.x + (long long)cc * a.ldx4 + gl;
#pragma unroll
                    for (int j = 0; j < VPL; ++j) xv[u][j] = ld_gather_f4(src + j * LPR, pol_g);
                }
            }
            // consume entries [u0, nvalid) of the batch, one row segment at a time
            const int nvalid = min(UNROLL, stream_end - e0);
            int u0 = max(0, dead_until - e0);
            while (u0 < nvalid) {
                if (e0 + u0 >= row_stop) {  // row r is complete (or empty): write it, start the next
                    flush(r);
                    next_row();
                    u0 = max(u0, dead_until - e0);  // next_row may have begun a jump over a long row
                    continue;
                }
                const int lim = min(nvalid, row_stop - e0);  // entries [u0, lim) continue row r
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (u >= u0 && u < lim) {
#pragma unroll
                        for (int j = 0; j < VPL; ++j) fma4(acc[j], vv[u], xv[u][j]);
                    }
                }
                u0 = lim;
            }
            if constexpr (MAP) {     // the sub-warps are converged again: map the stashes that are (nearly) full
                if (__any_sync(kFull, scount >= kMapRows - 1)) map_stash();
            }
        }
        if (dead_until > kc + LPR && kc < stream_end) {  // jump over a long row: reload the chunk registers
            kc = dead_until;
            c_cur = 0;
            v_cur = 0.f;
            if (kc + gl < stream_end) {
                c_cur = ld_stream_i32(a.indices + kc + gl, pol_s);
                v_cur = ld_stream_f32(a.vals + kc + gl, pol_s);
            }
        } else {
            kc += LPR;
            c_cur = c_nxt;
            v_cur = v_nxt;
        }
    }
    // the last row with entries and any trailing empty rows of the group
    while (r < r_end) {
        flush(r);
        next_row();
    }
    if constexpr (MAP) {
        if (__any_sync(kFull, scount > 0)) map_stash();
    }
}

// ---------------------------------------------------------------------------------------------
// long rows: one CTA per row, warp-specialised cp.async pipeline through shared memory
//
//   4 producer warps  : warp p owns chunks p, p+4, ...: it copies the chunk's 32 gathered x rows
//                       (cp.async 16 B) and 32 values into a ring stage.  Column indices are read
//                       straight from global memory (one coalesced load per chunk, four turns ahead)
//                       and broadcast by shuffle, so a gather never waits on them.
//                       `cp.async.mbarrier.arrive.noinc` publishes the stage on full[stage] when
//                       the copies land; a stage is reused once the consumers released empty[stage].
//   d/32 consumer warps: one feature per lane.  They wait on full[stage], pull the chunk into
//                       registers, release the stage at once, and run the dependent fmaf chain of
//                       the PREVIOUS chunk between those loads — no CTA-wide barrier in steady
//                       state, so producers run ahead by the ring depth and the chain (4 cycles per
//                       entry) is the only serial resource.  The first version met at one
//                       __syncthreads per chunk with the producers' issue code in between:
//                       13 cycles per entry at d = 64 (ncu: sm__cycles_active.max 371 K for a
//                       27 950-entry row).
// ---------------------------------------------------------------------------------------------
template <int D>
struct LongCfg {
    static constexpr int CONS = D / 32;
    static constexpr int PROD = 4;
    static constexpr int THREADS = (CONS + PROD) * 32;
    static constexpr int CHUNK = 32;     // two register sets of 32 gathered values + 32 coefficients per lane
    static constexpr int STAGE_BYTES = CHUNK * D * 4 + CHUNK * 4;   // gathered rows + values
    static constexpr int STAGES = D <= 64 ? 16 : (D == 128 ? 12 : 6);   // 128-197 KB of gathers in flight
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16;   // + barriers + ticket
};

// ticket / exit counters of the persistent long-row kernel, re-armed by the last CTA of each launch
// (one gr_spmm_csr_f32 in flight per device at a time)

__device__ __forceinline__ unsigned lr_smem(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void lr_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lr_smem(bar)), "r"(count));
}
__device__ __forceinline__ void lr_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lr_smem(bar)) : "memory");
}
// arrives on `bar` once all cp.async issued so far by this thread have landed (does not add to the
// expected count: the barrier is initialised with one arrival per producer thread)
__device__ __forceinline__ void lr_cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(lr_smem(bar)) : "memory");
}
__device__ __forceinline__ void lr_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LR_DONE;\n\t"
        "bra LR_WAIT;\n\t"
        "LR_DONE:\n\t}" ::"r"(lr_smem(bar)),
        "r"(parity)
        : "memory");
}

template <int D, int PEERS>
__global__ void __launch_bounds__(LongCfg<D>::THREADS, 1) spmm_long_rows(const SpmmArgs a) {
    using L = LongCfg<D>;
    constexpr int STAGES = L::STAGES, CH = L::CHUNK, F4 = D / 4, CONS = L::CONS, PROD = L::PROD;
    constexpr int ITEMS = (CH * F4) / 32;    // 16-byte copies per producer lane per chunk
    static_assert(CH == 32 && ITEMS >= 1, "one column index per producer lane");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * L::STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    int &s_ticket = *reinterpret_cast<int *>(empty + STAGES);
    auto stage_x = [&](int s) { return reinterpret_cast<float4 *>(smem_raw + (size_t)s * L::STAGE_BYTES); };
    auto stage_v = [&](int s) { return reinterpret_cast<float *>(smem_raw + (size_t)s * L::STAGE_BYTES + (size_t)CH * D * 4); };

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool is_cons = warp < CONS;
    const int p = warp - CONS;  // producer index, < 0 for consumers

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            lr_mbar_init(&full[s], 32);           // one cp.async-arrive per lane of the owning producer warp
            lr_mbar_init(&empty[s], CONS * 32);   // one release per consumer thread
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // chunks processed so far by this CTA (over all its rows): stage = g % STAGES, phase = g / STAGES
    unsigned g = 0;

    // persistent CTAs: rows are handed out in row_order (longest first) through a ticket counter,
    // so the hottest row starts first and no CTA queues work behind it.
    for (;;) {
    if (threadIdx.x == 0) s_ticket = (int)atomicAdd(a.sched, 1u);
    __syncthreads();
    const int ticket = s_ticket;
    __syncthreads();   // everyone has read the ticket before thread 0 may overwrite it
    int r, start, len, part = -1;
    if (a.items) {
        if (ticket >= a.n_items) break;
        r = a.items[ticket];
        start = a.indptr[r] + a.items[a.n_items + ticket];
        len = a.items[2 * a.n_items + ticket];
        part = a.items[3 * a.n_items + ticket];
    } else {
        if (ticket >= a.order_end - a.order_begin) break;
        r = a.row_order[a.order_begin + ticket];
        start = a.indptr[r];
        len = a.indptr[r + 1] - start;
    }
    const int nchunks = (len + CH - 1) / CH;
    const int *ci = a.indices + start;
    const float *cv = a.vals + start;

    float acc = 0.f;
    if (!is_cons) {
        // ===== producers: warp p owns chunks p, p + PROD, ... =====
        // lane l holds the column index of entry l of the warp's next two chunks (one coalesced
        // 128-byte load each, issued two turns ahead), broadcast by shuffle when the gathers are issued
        auto load_col = [&](int chunk) {
            const int e = chunk * CH + lane;
            return (chunk < nchunks && e < len) ? __ldg(ci + e) : -1;
        };
        // four turns ahead: one turn (~500 cycles) does not cover an L2 round trip
        int col0 = load_col(p), col1 = load_col(p + PROD), col2 = load_col(p + 2 * PROD), col3 = load_col(p + 3 * PROD);
        for (int c = p; c < nchunks; c += PROD) {
            const unsigned gc = g + (unsigned)c;
            const int s = (int)(gc % STAGES);
            if (gc >= (unsigned)STAGES) lr_mbar_wait(&empty[s], ((gc / STAGES) - 1) & 1);
            float4 *xs = stage_x(s);
#pragma unroll
            for (int t = 0; t < ITEMS; ++t) {
                const int item = lane + 32 * t;
                const int e = item / F4, f = item % F4;
                const int cc = __shfl_sync(0xffffffffu, col0, e);
                if (cc >= 0) cp_async16_plain(xs + (size_t)e * F4 + f, a.x + (long long)cc * a.ldx4 + f);
            }
            if (c * CH + lane < len) cp_async4_plain(stage_v(s) + lane, cv + c * CH + lane);
            lr_cp_async_arrive(&full[s]);
            col0 = col1;
            col1 = col2;
            col2 = col3;
            col3 = load_col(c + 4 * PROD);
        }
    } else {
        // ===== consumers =====
        float xa[CH], xb[CH];
        float4 va[CH / 4], vb[CH / 4];
        const int f = warp * 32 + lane;
        // pulls chunk c into registers and releases its stage; a partial (last) chunk is chained in place
        auto fetch = [&](int c, float (&dx)[CH], float4 (&dv)[CH / 4], bool &pending) {
            const unsigned gc = g + (unsigned)c;
            const int s = (int)(gc % STAGES);
            lr_mbar_wait(&full[s], (gc / STAGES) & 1);
            const float *xr = reinterpret_cast<const float *>(stage_x(s)) + f;
            const int n = min(CH, len - c * CH);
            if (n == CH) {
                const float4 *vr = reinterpret_cast<const float4 *>(stage_v(s));
#pragma unroll
                for (int k = 0; k < CH; ++k) dx[k] = xr[k * D];
#pragma unroll
                for (int k4 = 0; k4 < CH / 4; ++k4) dv[k4] = vr[k4];
                pending = true;
            } else {
                pending = false;
            }
            return n;
        };
        auto chain = [&](const float (&sx)[CH], const float4 (&sv)[CH / 4]) {
#pragma unroll
            for (int k4 = 0; k4 < CH / 4; ++k4) {
                acc = __fmaf_rn(sv[k4].x, sx[4 * k4 + 0], acc);
                acc = __fmaf_rn(sv[k4].y, sx[4 * k4 + 1], acc);
                acc = __fmaf_rn(sv[k4].z, sx[4 * k4 + 2], acc);
                acc = __fmaf_rn(sv[k4].w, sx[4 * k4 + 3], acc);
            }
        };
        auto tail = [&](int c, int n) {       // partial last chunk: straight from shared memory, then release
            const int s = (int)((g + (unsigned)c) % STAGES);
            const float *xr = reinterpret_cast<const float *>(stage_x(s)) + f;
            const float *v1 = stage_v(s);
            for (int k = 0; k < n; ++k) acc = __fmaf_rn(v1[k], xr[k * D], acc);
        };
        bool pa = false, pb = false;
        for (int c = 0; c < nchunks; c += 2) {
            // even chunk -> set A, while the chain of set B (chunk c-1) issues
            int n = fetch(c, xa, va, pa);
            if (pb) { chain(xb, vb); pb = false; }
            if (n < CH) tail(c, n);
            lr_mbar_arrive(&empty[(g + (unsigned)c) % STAGES]);
            if (c + 1 < nchunks) {
                n = fetch(c + 1, xb, vb, pb);
                if (pa) { chain(xa, va); pa = false; }
                if (n < CH) tail(c + 1, n);
                lr_mbar_arrive(&empty[(g + (unsigned)c + 1u) % STAGES]);
            }
        }
        if (pa) chain(xa, va);
        if (pb) chain(xb, vb);
    }
    g += (unsigned)nchunks;

    if (is_cons && part >= 0) {
        a.part_buf[(long long)part * D + warp * 32 + lane] = acc;
    } else if (is_cons) {
        const int f = warp * 32 + lane;
        if (a.y) reinterpret_cast<float *>(a.y + (long long)r * a.ldy4)[f] = acc;
        if constexpr (PEERS == kPeersP2P) {
            if (a.route_block > 0) {
                const int pg = r / a.route_block;
                reinterpret_cast<float *>(a.peer_y[pg] + (a.peer_row_off + (r - pg * a.route_block)) * a.ldy4)[f] = acc;
            } else {
#pragma unroll 1
                for (int p = 0; p < a.n_peers; ++p)
                    reinterpret_cast<float *>(a.peer_y[p] + (a.peer_row_off + r) * a.ldy4)[f] = acc;
            }
        } else if constexpr (PEERS == kPeersMulticast) {
            st_multimem_f1(reinterpret_cast<float *>(a.peer_y[0] + (a.peer_row_off + r) * a.ldy4) + f, acc);
        }
        if (a.out) {
            float o = acc;
            if (a.addend) o = __fadd_rn(reinterpret_cast<const float *>(a.addend + (long long)r * a.lda4)[f], o);
            reinterpret_cast<float *>(a.out + (long long)r * a.ldo4)[f] = apply_scale(o, a.scale, a.scale_mode);
        }
    }
    }  // ticket loop
    // the last CTA to leave re-arms the counters, so every launch (and every profiler replay)
    // starts from ticket 0 without host involvement
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.sched + 1, 1u) == gridDim.x - 1) {
            a.sched[0] = 0;
            a.sched[1] = 0;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// long rows, d <= 64: the same roles with ONE CTA-wide barrier per 64-entry chunk (index ring filled
// 16 chunks ahead by cp.async).  A single CTA gathers at most ~20-25 B/cycle from L2 (measured: 13
// cycles per 256-byte entry, hottest row of the Amazon-Book shape), which bounds a hot row long before
// the 4-cycle fmaf chain does; at d = 64 this simpler pipeline reaches that bound and measured faster
// than the mbarrier version (C4: 22.0 vs 17.9 G edges/s), at d >= 128 the mbarrier version wins
// (C5/8: 34.7 vs 37.6 ms per step).
// ---------------------------------------------------------------------------------------------
template <int D>
struct LongCfgBar {
    static constexpr int CONS = D / 32;
    static constexpr int PROD = 4;
    static constexpr int THREADS = (CONS + PROD) * 32;
    static constexpr int CHUNK = 64;
    static constexpr int STAGE_BYTES = CHUNK * D * 4;
    static constexpr int STAGES = D <= 64 ? 8 : (D == 128 ? 6 : 3);
    static constexpr int IDX_RING = 32;   // chunks of (col,val) resident in shared memory
    static constexpr int IDX_AHEAD = 16;  // index chunks are requested this far ahead of their gathers
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * (size_t)IDX_RING * CHUNK * 4;
    static_assert(STAGES - 1 <= IDX_AHEAD && STAGES + IDX_AHEAD + 1 <= IDX_RING, "index ring too small");
};

template <int D, int PEERS, bool MAP = false>
__global__ void __launch_bounds__(LongCfgBar<D>::THREADS, 1) spmm_long_rows_bar(const SpmmArgs a) {
    using L = LongCfgBar<D>;
    constexpr int STAGES = L::STAGES, CH = L::CHUNK, F4 = D / 4, CONS = L::CONS, PROD = L::PROD;
    constexpr int IR = L::IDX_RING, IA = L::IDX_AHEAD;
    constexpr int NB = CH / PROD;            // entries per producer warp per chunk
    constexpr int ITEMS = (NB * F4) / 32;    // 16-byte copies per producer lane per chunk
    static_assert(ITEMS >= 1 && (NB * F4) % 32 == 0, "feature dim");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *xs = reinterpret_cast<float4 *>(smem_raw);                                  // [STAGES][CH][F4]
    int *cs = reinterpret_cast<int *>(smem_raw + (size_t)STAGES * L::STAGE_BYTES);      // [IR][CH]
    float *vs = reinterpret_cast<float *>(cs + IR * CH);                                // [IR][CH]

    const uint64_t pol_s = policy_evict_first(), pol_g = policy_evict_last();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool is_cons = warp < CONS;
    const int p = warp - CONS;  // producer index, < 0 for consumers

    // persistent CTAs: rows are handed out in row_order (longest first) through a ticket counter,
    // so the hottest row starts first and no CTA queues work behind it.
    int &s_ticket = *reinterpret_cast<int *>(smem_raw + L::SMEM);  // one int after the rings
    // dense-map epilogue: one staging row after the ticket word; the map itself is read through L1 (a few hundred
    // long rows per launch — 16 KB of shared memory here would cost the streaming kernel a resident CTA)
    float *srow = reinterpret_cast<float *>(smem_raw + L::SMEM + 16);   // [D]
    float map_beta = 0.f;
    if constexpr (MAP) map_beta = map_beta_of(a);
    for (;;) {
    if (threadIdx.x == 0) s_ticket = (int)atomicAdd(a.sched, 1u);
    __syncthreads();
    const int ticket = s_ticket;
    int r, start, len, part = -1;
    if (a.items) {
        if (ticket >= a.n_items) break;
        r = a.items[ticket];
        start = a.indptr[r] + a.items[a.n_items + ticket];
        len = a.items[2 * a.n_items + ticket];
        part = a.items[3 * a.n_items + ticket];
    } else {
        if (ticket >= a.order_end - a.order_begin) break;
        r = a.row_order[a.order_begin + ticket];
        start = a.indptr[r];
        len = a.indptr[r + 1] - start;
    }
    const int nchunks = (len + CH - 1) / CH;
    const int *ci = a.indices + start;
    const float *cv = a.vals + start;

    auto issue_idx = [&](int chunk) {
        if (p == 0 && chunk < nchunks) {
            const int slot = chunk % IR;
#pragma unroll
            for (int t = 0; t < CH / 32; ++t) {
                const int e = lane + 32 * t;
                const int idx = chunk * CH + e;
                if (idx < len) {
                    cp_async4_plain(cs + slot * CH + e, ci + idx);
                    cp_async4_plain(vs + slot * CH + e, cv + idx);
                }
            }
        }
    };
    auto issue_data = [&](int chunk) {
        if (p >= 0 && chunk < nchunks) {
            const int stage = chunk % STAGES;
            const int slot = chunk % IR;
            const int nvalid = min(CH, len - chunk * CH);
#pragma unroll
            for (int t = 0; t < ITEMS; ++t) {
                const int item = lane + 32 * t;
                const int e = p * NB + item / F4;
                const int f = item % F4;
                if (e < nvalid) {
                    const int cc = cs[slot * CH + e];
                    cp_async16_plain(xs + ((size_t)stage * CH + e) * F4 + f, a.x + (long long)cc * a.ldx4 + f);
                }
            }
        }
    };

    // prologue: the first IA index chunks, then the first STAGES-1 data stages
    for (int k = 0; k < IA; ++k) issue_idx(k);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        issue_data(s);
        issue_idx(s + IA);
        cp_async_commit();
    }

    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<STAGES - 2>();  // this thread's copies for chunk c (and its index chunk) landed
        __syncthreads();              // everyone's did; consumers are done with chunk c-1
        if (!is_cons) {
            issue_data(c + STAGES - 1);  // refills the stage chunk c-1 used
            issue_idx(c + STAGES - 1 + IA);
        }
        cp_async_commit();
        if (is_cons) {
            const int stage = c % STAGES;
            const float *xr = reinterpret_cast<const float *>(xs + (size_t)stage * CH * F4) + warp * 32 + lane;
            const float4 *vr = reinterpret_cast<const float4 *>(vs + (c % IR) * CH);
            const int n = min(CH, len - c * CH);
            if (n == CH) {
#pragma unroll
                for (int k4 = 0; k4 < CH / 4; ++k4) {
                    const float4 v = vr[k4];
                    acc = __fmaf_rn(v.x, xr[(4 * k4 + 0) * D], acc);
                    acc = __fmaf_rn(v.y, xr[(4 * k4 + 1) * D], acc);
                    acc = __fmaf_rn(v.z, xr[(4 * k4 + 2) * D], acc);
                    acc = __fmaf_rn(v.w, xr[(4 * k4 + 3) * D], acc);
                }
            } else {
                const float *v1 = reinterpret_cast<const float *>(vr);
                for (int k = 0; k < n; ++k) acc = __fmaf_rn(v1[k], xr[k * D], acc);
            }
        }
    }
    cp_async_wait<0>();

    if constexpr (MAP) {
        // the row is spread over the consumer warps (one feature per lane): through shared memory, then thread f
        // forms column f of t M.  `part` is uniform over the CTA, so the barrier is not divergent.
        const int f = warp * 32 + lane;
        if (is_cons && part >= 0) a.part_buf[(long long)part * D + f] = acc;
        if (is_cons && part < 0) srow[f] = acc;
        __syncthreads();
        if (is_cons && part < 0) {
            float o = 0.f;
#pragma unroll 8
            for (int i = 0; i < D; ++i)
                o = __fmaf_rn(srow[i], a.map_transposed ? __ldg(a.map + f * D + i) : __ldg(a.map + i * D + f), o);
            if (a.y) reinterpret_cast<float *>(a.y + (long long)r * a.ldy4)[f] = __fmul_rn(acc, a.map_alpha);
            o = __fmul_rn(o, a.map_alpha);
            if (a.addend) o = __fmaf_rn(map_beta, reinterpret_cast<const float *>(a.addend + (long long)r * a.lda4)[f], o);
            reinterpret_cast<float *>(a.out + (long long)r * a.ldo4)[f] = o;
        }
    } else if (is_cons && part >= 0) {
        a.part_buf[(long long)part * D + warp * 32 + lane] = acc;
    } else if (is_cons) {
        const int f = warp * 32 + lane;
        if (a.y) reinterpret_cast<float *>(a.y + (long long)r * a.ldy4)[f] = acc;
        if constexpr (PEERS == kPeersP2P) {
            if (a.route_block > 0) {
                const int pg = r / a.route_block;
                reinterpret_cast<float *>(a.peer_y[pg] + (a.peer_row_off + (r - pg * a.route_block)) * a.ldy4)[f] = acc;
            } else {
#pragma unroll 1
                for (int p = 0; p < a.n_peers; ++p)
                    reinterpret_cast<float *>(a.peer_y[p] + (a.peer_row_off + r) * a.ldy4)[f] = acc;
            }
        } else if constexpr (PEERS == kPeersMulticast) {
            st_multimem_f1(reinterpret_cast<float *>(a.peer_y[0] + (a.peer_row_off + r) * a.ldy4) + f, acc);
        }
        if (a.out) {
            float o = acc;
            if (a.addend) o = __fadd_rn(reinterpret_cast<const float *>(a.addend + (long long)r * a.lda4)[f], o);
            reinterpret_cast<float *>(a.out + (long long)r * a.ldo4)[f] = apply_scale(o, a.scale, a.scale_mode);
        }
    }
    __syncthreads();  // the rings are reused by the next row
    }  // ticket loop
    // the last CTA to leave re-arms the counters, so every launch (and every profiler replay)
    // starts from ticket 0 without host involvement
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.sched + 1, 1u) == gridDim.x - 1) {
            a.sched[0] = 0;
            a.sched[1] = 0;
            __threadfence();
        }
    }
}

// Split rows: total = ((p0 + p1) + p2) + ... in segment order, then the usual epilogue.
__global__ void spmm_combine_parts(const SpmmArgs a, int d) {
    __shared__ float srow[256];
    const int sidx = blockIdx.x;
    const int r = a.split_rows[sidx];
    const int first = a.split_rows[a.n_split + sidx];
    const int n = a.split_rows[2 * a.n_split + sidx];
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        float t = a.part_buf[(long long)first * d + f];
        for (int k = 1; k < n; ++k) t = __fadd_rn(t, a.part_buf[(long long)(first + k) * d + f]);
        if (a.map) {                // dense-map epilogue below (d <= 256)
            srow[f] = t;
            continue;
        }
        if (a.y) reinterpret_cast<float *>(a.y + (long long)r * a.ldy4)[f] = t;
        if (a.peer_multicast) {
            st_multimem_f1(reinterpret_cast<float *>(a.peer_y[0] + (a.peer_row_off + r) * a.ldy4) + f, t);
        } else if (a.route_block > 0 && a.n_peers > 0) {
            const int pg = r / a.route_block;
            reinterpret_cast<float *>(a.peer_y[pg] + (a.peer_row_off + (r - pg * a.route_block)) * a.ldy4)[f] = t;
        } else {
            for (int p = 0; p < a.n_peers; ++p)
                reinterpret_cast<float *>(a.peer_y[p] + (a.peer_row_off + r) * a.ldy4)[f] = t;
        }
        if (a.out) {
            float o = t;
            if (a.addend) o = __fadd_rn(reinterpret_cast<const float *>(a.addend + (long long)r * a.lda4)[f], o);
            reinterpret_cast<float *>(a.out + (long long)r * a.ldo4)[f] = apply_scale(o, a.scale, a.scale_mode);
        }
    }
    if (a.map) {
        __syncthreads();
        const float beta = map_beta_of(a);
        for (int f = threadIdx.x; f < d; f += blockDim.x) {
            float o = 0.f;
            for (int i = 0; i < d; ++i)
                o = __fmaf_rn(srow[i], a.map_transposed ? __ldg(a.map + f * d + i) : __ldg(a.map + i * d + f), o);
            if (a.y) reinterpret_cast<float *>(a.y + (long long)r * a.ldy4)[f] = __fmul_rn(srow[f], a.map_alpha);
            o = __fmul_rn(o, a.map_alpha);
            if (a.addend) o = __fmaf_rn(beta, reinterpret_cast<const float *>(a.addend + (long long)r * a.lda4)[f], o);
            reinterpret_cast<float *>(a.out + (long long)r * a.ldo4)[f] = o;
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
    bool smem_attr_set_map[4] = {false, false, false, false};
};

// One high-priority side stream (+ fork/join events) per (device, caller stream): the long-row kernel of a
// propagation runs beside its streaming kernel, and propagations issued on different caller streams do not
// share a side stream or its events.  Host-side resource cache only (created on first use, never
// reassigned); all device-side scheduler state lives in caller memory (SpmmArgs::sched).
static std::mutex g_side_mu;
static std::map<std::pair<int, cudaStream_t>, SideStream> g_side;

static int get_side(cudaStream_t caller, SideStream **out) {
    int dev = 0;
    GR_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_side_mu);
    SideStream &s = g_side[std::make_pair(dev, caller)];
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
    SideStream *side = nullptr;
    const bool use_long = base.row_order != nullptr && n_long > 0;
    if (use_long) {
        int rc = get_side(stream, &side);
        if (rc != GR_OK) return rc;
        // d <= 64: barrier pipeline; d >= 128: mbarrier pipeline (see the comments at the kernels)
        constexpr bool kBar = D <= 64;
        constexpr int kThreads = kBar ? LongCfgBar<D>::THREADS : LongCfg<D>::THREADS;
        constexpr size_t kSmem = kBar ? LongCfgBar<D>::SMEM + 16 : LongCfg<D>::SMEM;
        auto k_none = kBar ? spmm_long_rows_bar<D, kPeersNone> : spmm_long_rows<D, kPeersNone>;
        auto k_p2p = kBar ? spmm_long_rows_bar<D, kPeersP2P> : spmm_long_rows<D, kPeersP2P>;
        auto k_mc = kBar ? spmm_long_rows_bar<D, kPeersMulticast> : spmm_long_rows<D, kPeersMulticast>;
        if (!side->smem_attr_set[slot]) {
            GR_CUDA_CHECK(cudaFuncSetAttribute(k_none, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
            GR_CUDA_CHECK(cudaFuncSetAttribute(k_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
            GR_CUDA_CHECK(cudaFuncSetAttribute(k_mc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
            side->smem_attr_set[slot] = true;
        }
        // dense-map epilogue (d <= 64: the barrier pipeline): a staging row follows the rings
        constexpr size_t kSmemMap = kSmem + (size_t)D * 4;
        if constexpr (kBar) {
            if (base.map && !side->smem_attr_set_map[slot]) {
                GR_CUDA_CHECK(cudaFuncSetAttribute(spmm_long_rows_bar<D, kPeersNone, true>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMap));
                side->smem_attr_set_map[slot] = true;
            }
        }
        // GR_LONG_SERIAL=1: long rows first on the caller's stream instead of concurrently on the side stream
        static const bool serial_long = [] { const char *e = getenv("GR_LONG_SERIAL"); return e && atoi(e) != 0; }();
        cudaStream_t ls = serial_long ? stream : side->stream;
        if (!serial_long) {
            GR_CUDA_CHECK(cudaEventRecord(side->fork, stream));
            GR_CUDA_CHECK(cudaStreamWaitEvent(side->stream, side->fork, 0));
        }
        SpmmArgs la = base;
        la.order_begin = 0;
        la.order_end = n_long;
        const int n_work = la.items ? la.n_items : n_long;
        int long_ctas = sm_count();
        if (long_ctas > n_work) long_ctas = n_work;
        if (long_ctas < 1) long_ctas = 1;
        if (la.map) {
            if constexpr (kBar) spmm_long_rows_bar<D, kPeersNone, true><<<long_ctas, kThreads, kSmemMap, ls>>>(la);
        } else if (la.n_peers > 0 && la.peer_multicast)
            k_mc<<<long_ctas, kThreads, kSmem, ls>>>(la);
        else if (la.n_peers > 0)
            k_p2p<<<long_ctas, kThreads, kSmem, ls>>>(la);
        else
            k_none<<<long_ctas, kThreads, kSmem, ls>>>(la);
        GR_LAUNCH_CHECK();
        if (la.items && la.n_split > 0) {
            spmm_combine_parts<<<la.n_split, 128, 0, ls>>>(la, D);
            GR_LAUNCH_CHECK();
        }
        if (!serial_long) GR_CUDA_CHECK(cudaEventRecord(side->join, side->stream));
    }
    const long long rest = n_rows - (use_long ? n_long : 0);
    if (base.group_ptr != nullptr) {
        SpmmArgs wa = base;
        wa.order_begin = 0;
        wa.order_end = (int)n_rows;
        if (!use_long) wa.long_thr = 0x7fffffff;
        const long long per_cta = (long long)kWarpsPerCta * C::RPW;
        const long long ctas = (base.n_groups + per_cta - 1) / per_cta;
        if (ctas > 0) {
            if (base.map) {
                if constexpr (D <= 64) {
                    constexpr size_t kMapSmem = (size_t)D * D * 4 + (size_t)kWarpsPerCta * C::RPW * kMapRows * (D * 4 + 4);
                    spmm_stream_rows<D, kPeersNone, true><<<(unsigned)ctas, kWarpsPerCta * 32, kMapSmem, stream>>>(wa);
                }
            } else if (base.n_peers > 0 && base.peer_multicast)
                spmm_stream_rows<D, kPeersMulticast><<<(unsigned)ctas, kWarpsPerCta * 32, 0, stream>>>(wa);
            else if (base.n_peers > 0)
                spmm_stream_rows<D, kPeersP2P><<<(unsigned)ctas, kWarpsPerCta * 32, 0, stream>>>(wa);
            else
                spmm_stream_rows<D, kPeersNone><<<(unsigned)ctas, kWarpsPerCta * 32, 0, stream>>>(wa);
            GR_LAUNCH_CHECK();
        }
    } else if (rest > 0) {
        SpmmArgs wa = base;
        wa.order_begin = use_long ? n_long : 0;
        wa.order_end = (int)n_rows;
        const long long tiles = (rest + C::RPW - 1) / C::RPW;
        const long long ctas = (tiles + kWarpsPerCta - 1) / kWarpsPerCta;
        spmm_warp_rows<D><<<(unsigned)ctas, kWarpsPerCta * 32, 0, stream>>>(wa);
        GR_LAUNCH_CHECK();
    }
    if (use_long) {
        static const bool serial_long = [] { const char *e = getenv("GR_LONG_SERIAL"); return e && atoi(e) != 0; }();
        if (!serial_long) GR_CUDA_CHECK(cudaStreamWaitEvent(stream, side->join, 0));
    }
    return GR_OK;
}

// rows [0, n_rows) of src -> rows [row_off, row_off + n_rows) of every peer buffer (layer-0 exchange)
__global__ void peer_scatter_rows_kernel(const float4 *src, long long lds4, long long n_rows, int f4,
                                         SpmmArgs a) {
    const long long total = n_rows * f4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / f4;
        const int f = (int)(i % f4);
        const float4 v = __ldg(src + r * lds4 + f);
        if (a.peer_multicast) {
            st_multimem_f4(a.peer_y[0] + (a.peer_row_off + r) * a.ldy4 + f, v);
        } else {
            for (int p = 0; p < a.n_peers; ++p) a.peer_y[p][(a.peer_row_off + r) * a.ldy4 + f] = v;
        }
    }
}

}  // namespace gr

extern "C" int gr_peer_scatter_rows(const float *src, int64_t lds, int64_t n_rows, int32_t d,
                                    float *const *peer_dst_host, int32_t n_peers, int32_t peer_multicast,
                                    int64_t ldd, int64_t peer_row_offset, void *stream) {
    using namespace gr;
    if (!src || !peer_dst_host || n_rows < 0 || n_peers < 1 || n_peers > kMaxPeers) return GR_ERR_INVALID;
    if ((d & 3) || (lds & 3) || (ldd & 3) || lds < d || ldd < d || !aligned16(src)) return GR_ERR_INVALID;
    if (n_rows == 0) return GR_OK;
    SpmmArgs a = {};
    a.n_peers = n_peers;
    a.peer_multicast = peer_multicast ? 1 : 0;
    if (a.peer_multicast && n_peers != 1) return GR_ERR_INVALID;
    a.peer_row_off = peer_row_offset;
    a.ldy4 = ldd / 4;
    for (int p = 0; p < n_peers; ++p) {
        if (!peer_dst_host[p] || !aligned16(peer_dst_host[p])) return GR_ERR_INVALID;
        a.peer_y[p] = reinterpret_cast<float4 *>(peer_dst_host[p]);
    }
    const long long total = n_rows * (d / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 2;     // NVLink-bound; leaves the SMs to a concurrent SpMM
    if (blocks > cap) blocks = cap;
    peer_scatter_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(src), lds / 4, n_rows, d / 4, a);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

static int spmm_entry(const int32_t *indptr, const int32_t *indices, const float *vals,
                      const int32_t *row_order, int32_t n_long, const int32_t *long_items,
                      int32_t n_long_items, const int32_t *split_rows, int32_t n_split, float *part_buf,
                      const int32_t *group_ptr,
                      int32_t n_groups, int32_t long_threshold, int64_t n_rows, int32_t d, const float *x,
                      int64_t ldx, float *y, int64_t ldy, const float *addend, int64_t lda, float *out,
                      int64_t ldo, float scale, int32_t scale_mode, float *const *peer_y_host,
                      int32_t n_peers, int32_t peer_multicast, int64_t peer_row_offset, int32_t peer_route_block,
                      const float *map, int32_t map_transposed, float map_alpha, float map_beta,
                      const float *map_beta_dev, uint32_t *sched_ws, void *stream) {
    using namespace gr;
    if (n_rows == 0) return GR_OK;
    if (row_order && n_long > 0 && !sched_ws) return GR_ERR_INVALID;      // long-row scheduler needs its 2 words
    if (n_peers < 0 || n_peers > kMaxPeers || (n_peers > 0 && !peer_y_host)) return GR_ERR_INVALID;
    if (!indptr || !indices || !vals || !x || n_rows < 0 || n_long < 0 || n_long > n_rows) return GR_ERR_INVALID;
    if (!y && !out && n_peers == 0) return GR_ERR_INVALID;
    if (group_ptr && (n_groups < 0 || long_threshold < 1)) return GR_ERR_INVALID;
    if (long_items && (!row_order || n_long_items < n_long || n_split < 0 || (n_split > 0 && (!split_rows || !part_buf))))
        return GR_ERR_INVALID;
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
    a.ldy4 = ldy / 4;  // also the leading dimension of the peers' gathered buffers
    a.addend = reinterpret_cast<const float4 *>(addend);
    a.lda4 = lda / 4;
    a.out = reinterpret_cast<float4 *>(out);
    a.ldo4 = ldo / 4;
    a.scale = scale;
    a.scale_mode = scale_mode;
    a.group_ptr = group_ptr;
    a.n_groups = group_ptr ? n_groups : 0;
    a.long_thr = long_threshold;
    a.items = long_items;
    a.n_items = long_items ? n_long_items : 0;
    a.split_rows = split_rows;
    a.n_split = long_items ? n_split : 0;
    a.part_buf = part_buf;
    a.n_peers = n_peers;
    a.peer_multicast = (n_peers > 0 && peer_multicast) ? 1 : 0;
    if (a.peer_multicast && n_peers != 1) return GR_ERR_INVALID;
    a.peer_row_off = peer_row_offset;
    a.route_block = peer_route_block > 0 ? peer_route_block : 0;
    if (a.route_block > 0 && (a.peer_multicast || n_peers < 1 || (n_rows + a.route_block - 1) / a.route_block > n_peers))
        return GR_ERR_INVALID;
    a.sched = sched_ws;
    a.map = map;
    a.map_transposed = map_transposed ? 1 : 0;
    a.map_alpha = map_alpha;
    a.map_beta = map_beta;
    a.map_beta_dev = map_beta_dev;
    if (map) {   // dense-map epilogue: d <= 64, streaming schedule, single GPU, result in `out`
        if (d > 64 || !group_ptr || n_peers != 0) return GR_ERR_UNSUPPORTED;
        if (!out || !aligned16(map)) return GR_ERR_INVALID;
    }
    for (int p = 0; p < kMaxPeers; ++p) {
        a.peer_y[p] = p < n_peers ? reinterpret_cast<float4 *>(peer_y_host[p]) : nullptr;
        if (p < n_peers && (!peer_y_host[p] || !aligned16(peer_y_host[p]))) return GR_ERR_INVALID;
    }
    if (n_peers > 0 && ((ldy & 3) || ldy < d)) return GR_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
        case 32: return launch<32>(a, n_long, n_rows, 0, s);
        case 64: return launch<64>(a, n_long, n_rows, 1, s);
        case 128: return launch<128>(a, n_long, n_rows, 2, s);
        case 256: return launch<256>(a, n_long, n_rows, 3, s);
        default: return GR_ERR_UNSUPPORTED;
    }
}

extern "C" int gr_spmm_csr_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                               const int32_t *row_order, int32_t n_long, const int32_t *long_items,
                               int32_t n_long_items, const int32_t *split_rows, int32_t n_split, float *part_buf,
                               const int32_t *group_ptr,
                               int32_t n_groups, int32_t long_threshold, int64_t n_rows, int32_t d, const float *x,
                               int64_t ldx, float *y, int64_t ldy, const float *addend, int64_t lda, float *out,
                               int64_t ldo, float scale, int32_t scale_mode, float *const *peer_y_host,
                               int32_t n_peers, int32_t peer_multicast, int64_t peer_row_offset, int32_t peer_route_block,
                               uint32_t *sched_ws, void *stream) {
    return spmm_entry(indptr, indices, vals, row_order, n_long, long_items, n_long_items, split_rows, n_split, part_buf,
                      group_ptr, n_groups, long_threshold, n_rows, d, x, ldx, y, ldy, addend, lda, out, ldo, scale,
                      scale_mode, peer_y_host, n_peers, peer_multicast, peer_row_offset, peer_route_block, nullptr, 0,
                      1.f, 0.f, nullptr, sched_ws, stream);
}

extern "C" int gr_spmm_csr_map_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                                   const int32_t *row_order, int32_t n_long, const int32_t *long_items,
                                   int32_t n_long_items, const int32_t *split_rows, int32_t n_split, float *part_buf,
                                   const int32_t *group_ptr, int32_t n_groups, int32_t long_threshold, int64_t n_rows,
                                   int32_t d, const float *x, int64_t ldx, float *y, int64_t ldy, const float *addend,
                                   int64_t lda, float *out, int64_t ldo, const float *map, int32_t map_transposed,
                                   float alpha, float beta, const float *beta_dev, uint32_t *sched_ws, void *stream) {
    if (!map) return GR_ERR_INVALID;
    return spmm_entry(indptr, indices, vals, row_order, n_long, long_items, n_long_items, split_rows, n_split, part_buf,
                      group_ptr, n_groups, long_threshold, n_rows, d, x, ldx, y, ldy, addend, lda, out, ldo, 1.f,
                      GR_SCALE_NONE, nullptr, 0, 0, 0, 0, map, map_transposed, alpha, beta, beta_dev, sched_ws, stream);
}
