// GAT layer (src/models/baselines/gat.py:76-151, 283) over the CSR PATTERN of the adjacency, forward and
// backward, with hot rows cut into segments.
//
//   gr_gat_node_scores  s[i,h] = <H_h[i], a_self_h>, t[i,h] = <H_h[i], a_neigh_h>                 (gat.py:106-109)
//   gr_gat_aggregate    out_i = sum_j softmax_j(LeakyReLU(s_i + t_j)) c_ij H[j] over the row's neighbours,
//                       c_ij = attention dropout keep/(1-p) (gat.py:138); heads concatenated or averaged, ELU
//   gr_gat_bwd          its backward (what autograd runs under trainer.py:270): dH and (d a_self | d a_neigh)
//
// Work decomposition.  A warp owns one WORK ITEM: a whole row of at most `seg_len` entries, or one
// seg_len-entry segment of a longer row (power-law graphs: the hottest Yelp2018-shape row has 17 560
// entries; as one warp's job it alone took 3 ms per layer).  Segment items leave partial results — the
// online-softmax triple (max, normaliser, unnormalised accumulator) in the forward pass, plain partial sums in
// the backward pass — that a small combine kernel merges per long row in segment order (deterministic).
// The segment table is built once per adjacency pattern by the host (`gr_gat_segments`).
#include <math_constants.h>

#include "gr_common.cuh"

namespace gr {

__device__ __forceinline__ float elu1(float z) { return z > 0.f ? z : expm1f(z); }
__device__ __forceinline__ float elu_grad_from_output(float a) { return a > 0.f ? 1.f : a + 1.f; }

__device__ __forceinline__ float seg_sum(float v, int dh4) {
    for (int o = dh4 >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// <dO_i,h , H_k,h> restricted to this lane's float4; fixed operation order so the row-role and the
// column-role kernels produce the SAME bits for the same edge (their difference x - D must cancel).
__device__ __forceinline__ float dot4_fixed(const float4 &dO, const float4 &hv) {
    return __fmaf_rn(dO.w, hv.w, __fmaf_rn(dO.z, hv.z, __fmaf_rn(dO.y, hv.y, __fmul_rn(dO.x, hv.x))));
}
// alpha_ik = exp(LeakyReLU(s_i + t_k) - m_i) / z_i, same expression in both backward kernels
__device__ __forceinline__ float gat_alpha(float pre, float slope, float m, float zinv) {
    const float e = pre > 0.f ? pre : pre * slope;
    return __fmul_rn(expf(e - m), zinv);
}

struct SegTable {   // device pointers
    int seg_len, n_seg, n_long;
    const int *seg_row, *seg_begin, *seg_end, *long_rows, *long_seg_ptr;
};

// item -> (row, [start, end), segment index or -1).  Returns false when the item has no work.
__device__ __forceinline__ bool gat_item(const SegTable &sg, const int *indptr, int n_rows, int item, int &row,
                                         int &start, int &end, int &seg) {
    if (item < n_rows) {
        row = item;
        start = indptr[row];
        end = indptr[row + 1];
        seg = -1;
        return !(sg.n_seg > 0 && end - start > sg.seg_len);      // long rows are done by their segment items
    }
    seg = item - n_rows;
    if (seg >= sg.n_seg) return false;
    row = sg.seg_row[seg];
    start = sg.seg_begin[seg];
    end = sg.seg_end[seg];
    return true;
}

// =============================================================================================
// node scores
// =============================================================================================
__global__ void __launch_bounds__(256) gat_node_scores_kernel(const float *h, long long ldh, const float *a_self,
                                                              const float *a_neigh, int n, int heads, int dh,
                                                              float *s, float *t) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    for (int hd = 0; hd < heads; ++hd) {
        float ps = 0.f, pt = 0.f;
        for (int f = lane; f < dh; f += 32) {
            const float v = __ldg(h + (long long)i * ldh + hd * dh + f);
            ps = __fmaf_rn(v, __ldg(a_self + hd * dh + f), ps);
            pt = __fmaf_rn(v, __ldg(a_neigh + hd * dh + f), pt);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            ps += __shfl_xor_sync(0xffffffffu, ps, o);
            pt += __shfl_xor_sync(0xffffffffu, pt, o);
        }
        if (lane == 0) {
            s[(long long)i * heads + hd] = ps;
            t[(long long)i * heads + hd] = pt;
        }
    }
}

// =============================================================================================
// forward aggregation
// =============================================================================================
struct GatArgs {
    const int *indptr, *indices;
    const float *h;
    long long ldh;
    const float *s, *t;
    int n_rows, heads, dh;
    float slope;
    int mean_heads;  // 0: concat heads -> width heads*dh; 1: average heads -> width dh
    int elu;
    float *out;
    long long ldo;
    float *m_out, *z_out;  // [n_rows, heads] softmax statistics for the backward pass (may be NULL)
    long long n_cols;
    unsigned drop_thr;     // attention dropout (gat.py:138): weight kept with prob 1-p, scaled by 1/(1-p)
    float drop_scale;
    unsigned long long drop_seed;
    const unsigned long long *drop_seed_dev;   // optional: added to drop_seed at run time (CUDA-graph replays)
    SegTable sg;
    float *p_m, *p_z;      // [n_seg, heads]
    float *p_acc;          // [n_seg, heads*dh]
};

// acc (unnormalised), m, z of a complete row -> out row (heads concatenated or averaged, ELU), statistics
template <int SLOTS>
__device__ __forceinline__ void gat_finalize(const GatArgs &a, int i, int lane, const int *head, const bool *on,
                                             const float *m, const float *z, float4 *acc, float4 *st) {
    const int dh4 = a.dh / 4;
    // out = acc / z   (a row without neighbours gives 0/0 = NaN, like the reference's softmax of -inf)
#pragma unroll
    for (int q = 0; q < SLOTS; ++q)
        if (on[q]) {
            acc[q].x /= z[q]; acc[q].y /= z[q]; acc[q].z /= z[q]; acc[q].w /= z[q];
            const int slot = lane + 32 * q;
            if (a.m_out && (slot % dh4) == 0) {
                a.m_out[(long long)i * a.heads + head[q]] = m[q];
                a.z_out[(long long)i * a.heads + head[q]] = z[q];
            }
        }
    if (!a.mean_heads) {
#pragma unroll
        for (int q = 0; q < SLOTS; ++q)
            if (on[q]) {
                float4 o = acc[q];
                if (a.elu) { o.x = elu1(o.x); o.y = elu1(o.y); o.z = elu1(o.z); o.w = elu1(o.w); }
                *reinterpret_cast<float4 *>(a.out + (long long)i * a.ldo + (lane + 32 * q) * 4) = o;
            }
    } else {
        // average over heads (torch.stack(heads).mean(0): left-to-right sum / heads): slot (head, f4) -> f4
        __syncwarp();
#pragma unroll
        for (int q = 0; q < SLOTS; ++q)
            if (on[q]) st[lane + 32 * q] = acc[q];
        __syncwarp();
        for (int f = lane; f < dh4; f += 32) {
            float4 sum = st[f];
            for (int hd = 1; hd < a.heads; ++hd) {
                const float4 v = st[hd * dh4 + f];
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            const float hh = (float)a.heads;
            float4 o = make_float4(sum.x / hh, sum.y / hh, sum.z / hh, sum.w / hh);
            if (a.elu) { o.x = elu1(o.x); o.y = elu1(o.y); o.z = elu1(o.z); o.w = elu1(o.w); }
            *reinterpret_cast<float4 *>(a.out + (long long)i * a.ldo + f * 4) = o;
        }
    }
}

// One warp per work item, online softmax (running max / running sum, rescaled accumulator), one pass over
// the item's neighbours; lane owns float4 slots lane, lane+32 of the heads*dh wide row.  SLOTS = 1 or 2.
template <int SLOTS>
__global__ void __launch_bounds__(256) gat_aggregate_kernel(const GatArgs a) {
    __shared__ float4 stage[8][64];
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
    int i, start, end, seg;
    if (!gat_item(a.sg, a.indptr, a.n_rows, item, i, start, end, seg)) return;
    const int width4 = a.heads * a.dh / 4;
    const int dh4 = a.dh / 4;
    const int hgroups = (a.heads + 3) >> 2;
    const unsigned long long seed = a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0ULL);
    int head[SLOTS];
    bool on[SLOTS];
    float si[SLOTS], m[SLOTS], z[SLOTS];
    float4 acc[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        on[q] = slot < width4;
        head[q] = on[q] ? slot / dh4 : 0;
        si[q] = __ldg(a.s + (long long)i * a.heads + head[q]);
        m[q] = -CUDART_INF_F;
        z[q] = 0.f;
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int base = start; base < end; base += 32) {
        const int mycol = (base + lane < end) ? __ldg(a.indices + base + lane) : 0;
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; k += 4) {
            float4 hv[4][SLOTS];
            float tv[4][SLOTS];
            int jj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = __shfl_sync(0xffffffffu, mycol, (k + u) & 31);
                jj[u] = j;
                if (k + u < cnt) {
#pragma unroll
                    for (int q = 0; q < SLOTS; ++q)
                        if (on[q]) {
                            hv[u][q] = __ldg(reinterpret_cast<const float4 *>(a.h + (long long)j * a.ldh) + lane + 32 * q);
                            tv[u][q] = __ldg(a.t + (long long)j * a.heads + head[q]);
                        }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k + u < cnt) {
#pragma unroll
                    for (int q = 0; q < SLOTS; ++q)
                        if (on[q]) {
                            float e = si[q] + tv[u][q];
                            e = e > 0.f ? e : e * a.slope;
                            const float mn = fmaxf(m[q], e);
                            const float sc = expf(m[q] - mn);   // exp(-inf) = 0 on the first neighbour
                            float w = expf(e - mn);
                            z[q] = z[q] * sc + w;
                            if (a.drop_thr) {   // the softmax normaliser keeps every edge; only the weight is dropped
                                const unsigned long long bits = drop_bits(
                                    seed, ((unsigned long long)i * a.n_cols + jj[u]) * hgroups + (head[q] >> 2));
                                w = drop_keep(bits, head[q] & 3, a.drop_thr) ? w * a.drop_scale : 0.f;
                            }
                            acc[q].x = acc[q].x * sc + w * hv[u][q].x;
                            acc[q].y = acc[q].y * sc + w * hv[u][q].y;
                            acc[q].z = acc[q].z * sc + w * hv[u][q].z;
                            acc[q].w = acc[q].w * sc + w * hv[u][q].w;
                            m[q] = mn;
                        }
                }
            }
        }
    }
    if (seg >= 0) {      // a segment of a long row: leave the online-softmax triple for the combine kernel
#pragma unroll
        for (int q = 0; q < SLOTS; ++q)
            if (on[q]) {
                const int slot = lane + 32 * q;
                if ((slot % dh4) == 0) {
                    a.p_m[(long long)seg * a.heads + head[q]] = m[q];
                    a.p_z[(long long)seg * a.heads + head[q]] = z[q];
                }
                *reinterpret_cast<float4 *>(a.p_acc + ((long long)seg * width4 + slot) * 4) = acc[q];
            }
        return;
    }
    gat_finalize<SLOTS>(a, i, lane, head, on, m, z, acc, stage[threadIdx.x >> 5]);
}

// One warp per long row: merge its segments' (m, z, acc) in segment order, then finalize.
template <int SLOTS>
__global__ void __launch_bounds__(256) gat_aggregate_combine_kernel(const GatArgs a) {
    __shared__ float4 stage[8][64];
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= a.sg.n_long) return;
    const int i = a.sg.long_rows[r];
    const int width4 = a.heads * a.dh / 4, dh4 = a.dh / 4;
    int head[SLOTS];
    bool on[SLOTS];
    float m[SLOTS], z[SLOTS];
    float4 acc[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        on[q] = slot < width4;
        head[q] = on[q] ? slot / dh4 : 0;
        m[q] = -CUDART_INF_F;
        z[q] = 0.f;
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int sgi = a.sg.long_seg_ptr[r]; sgi < a.sg.long_seg_ptr[r + 1]; ++sgi) {
#pragma unroll
        for (int q = 0; q < SLOTS; ++q)
            if (on[q]) {
                const float ms = __ldg(a.p_m + (long long)sgi * a.heads + head[q]);
                const float zs = __ldg(a.p_z + (long long)sgi * a.heads + head[q]);
                const float4 as = __ldg(reinterpret_cast<const float4 *>(a.p_acc) + (long long)sgi * width4 + lane + 32 * q);
                const float mn = fmaxf(m[q], ms);
                const float sc = expf(m[q] - mn), ss = expf(ms - mn);
                z[q] = z[q] * sc + zs * ss;
                acc[q].x = acc[q].x * sc + as.x * ss; acc[q].y = acc[q].y * sc + as.y * ss;
                acc[q].z = acc[q].z * sc + as.z * ss; acc[q].w = acc[q].w * sc + as.w * ss;
                m[q] = mn;
            }
    }
    gat_finalize<SLOTS>(a, i, lane, head, on, m, z, acc, stage[threadIdx.x >> 5]);
}

// =============================================================================================
// backward
// =============================================================================================
struct GatPrepArgs {
    const float *dout, *out, *s, *m, *z;
    long long lddo, ldo;
    float *dO;      // [n, heads*dh] gradient w.r.t. the per-head aggregates
    float4 *stat;   // [n, heads] (s, m, 1/z, D); D is filled by the row-role kernels
    int n_rows, heads, dh, mean_heads, elu;
};

// One warp per row: undo ELU / head-mean on the incoming gradient; pack the softmax statistics.
template <int SLOTS>
__global__ void __launch_bounds__(256) gat_bwd_prep_kernel(const GatPrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= a.n_rows) return;
    const int width4 = a.heads * a.dh / 4, dh4 = a.dh / 4;
    const float inv_heads = 1.f / (float)a.heads;
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        if (slot >= width4) continue;
        const int head = slot / dh4;
        const int oslot = a.mean_heads ? slot % dh4 : slot;
        float4 d = __ldg(reinterpret_cast<const float4 *>(a.dout + (long long)i * a.lddo) + oslot);
        if (a.elu) {
            const float4 o = __ldg(reinterpret_cast<const float4 *>(a.out + (long long)i * a.ldo) + oslot);
            d.x *= elu_grad_from_output(o.x); d.y *= elu_grad_from_output(o.y);
            d.z *= elu_grad_from_output(o.z); d.w *= elu_grad_from_output(o.w);
        }
        if (a.mean_heads) { d.x *= inv_heads; d.y *= inv_heads; d.z *= inv_heads; d.w *= inv_heads; }
        *reinterpret_cast<float4 *>(a.dO + (long long)i * a.heads * a.dh + slot * 4) = d;
        if ((slot % dh4) == 0) {
            const long long ih = (long long)i * a.heads + head;
            a.stat[ih] = make_float4(__ldg(a.s + ih), __ldg(a.m + ih), 1.f / __ldg(a.z + ih), 0.f);
        }
    }
}

struct GatBwdArgs {
    const int *indptr, *indices;        // row pattern:    j attends to k in row j
    const int *t_indptr, *t_indices;    // column pattern: rows i that attend to j (== row pattern when symmetric)
    const float *h, *dO, *t, *a_self, *a_neigh;
    float4 *stat;         // [n, heads] (s, m, 1/z, D)
    float *ds;            // [n, heads]
    long long ldh;
    float *dH;            // [n, heads*dh]
    float *partial;       // [n_ctas][2 * heads*dh]: per-CTA sums of ds*H and dt*H (-> da_self, da_neigh)
    long long n_cols;
    int n_rows, heads, dh;
    float slope;
    unsigned drop_thr;
    float drop_scale;
    unsigned long long drop_seed;
    const unsigned long long *drop_seed_dev;
    SegTable rsg, csg;    // segments of the row pattern / of the column pattern
    float4 *p_S;          // [rsg.n_seg, heads]  partial (S0, S1, S2, S3)
    float *p_dt;          // [csg.n_seg, heads]
    float *p_dh;          // [csg.n_seg, heads*dh]
};

// Row role: for row j, one pass over (a segment of) its neighbours k, gathering H_k:
//   S0 = sum alpha, S1 = sum alpha l' c x, S2 = sum alpha l', S3 = sum alpha c x       (x_jk = <dO_j, H_k>)
//   D_j = S3 / S0  (the softmax-backward centring term, formed from the SAME alpha / x the terms use, not from
//   the forward output),  ds_j = sum_k alpha_jk l'_jk (c_jk x_jk - D_j) = S1 - D_j S2
template <int SLOTS>
__global__ void __launch_bounds__(256) gat_bwd_row_kernel(const GatBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
    int j, start, end, seg;
    if (!gat_item(a.rsg, a.indptr, a.n_rows, item, j, start, end, seg)) return;
    const int width4 = a.heads * a.dh / 4, dh4 = a.dh / 4;
    const int hgroups = (a.heads + 3) >> 2;
    const unsigned long long seed = a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0ULL);
    int head[SLOTS];
    bool on[SLOTS];
    float4 doj[SLOTS];
    float sj[SLOTS], mj[SLOTS], zinv[SLOTS], S0[SLOTS], S1[SLOTS], S2[SLOTS], S3[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        on[q] = slot < width4;
        head[q] = on[q] ? slot / dh4 : 0;
        doj[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        sj[q] = mj[q] = 0.f; zinv[q] = 1.f;
        S0[q] = S1[q] = S2[q] = S3[q] = 0.f;
        if (on[q]) {
            doj[q] = __ldg(reinterpret_cast<const float4 *>(a.dO + (long long)j * a.heads * a.dh) + slot);
            const float4 st = a.stat[(long long)j * a.heads + head[q]];
            sj[q] = st.x; mj[q] = st.y; zinv[q] = st.z;
        }
    }
    for (int base = start; base < end; base += 32) {
        const int mycol = (base + lane < end) ? __ldg(a.indices + base + lane) : 0;
        const int cnt = min(32, end - base);
        for (int kk = 0; kk < cnt; kk += 4) {
            float4 hv[4][SLOTS];
            float tv[4][SLOTS];
            int kid[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                kid[u] = __shfl_sync(0xffffffffu, mycol, (kk + u) & 31);
                if (kk + u < cnt) {
#pragma unroll
                    for (int q = 0; q < SLOTS; ++q)
                        if (on[q]) {
                            hv[u][q] = __ldg(reinterpret_cast<const float4 *>(a.h + (long long)kid[u] * a.ldh) + lane + 32 * q);
                            tv[u][q] = __ldg(a.t + (long long)kid[u] * a.heads + head[q]);
                        }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kk + u < cnt) {
#pragma unroll
                    for (int q = 0; q < SLOTS; ++q) {
                        float dot = on[q] ? dot4_fixed(doj[q], hv[u][q]) : 0.f;
                        dot = seg_sum(dot, dh4);
                        if (on[q]) {
                            const float pre = sj[q] + tv[u][q];
                            const float alpha = gat_alpha(pre, a.slope, mj[q], zinv[q]);
                            float c = 1.f;
                            if (a.drop_thr) {
                                const unsigned long long bits = drop_bits(
                                    seed, ((unsigned long long)j * a.n_cols + kid[u]) * hgroups + (head[q] >> 2));
                                c = drop_keep(bits, head[q] & 3, a.drop_thr) ? a.drop_scale : 0.f;
                            }
                            const float lp = pre > 0.f ? 1.f : a.slope;
                            const float ax = alpha * (c * dot);
                            S0[q] += alpha;
                            S1[q] += ax * lp;
                            S2[q] += alpha * lp;
                            S3[q] += ax;
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        if (on[q] && (slot % dh4) == 0) {
            if (seg >= 0) {
                a.p_S[(long long)seg * a.heads + head[q]] = make_float4(S0[q], S1[q], S2[q], S3[q]);
            } else {
                const long long jh = (long long)j * a.heads + head[q];
                const float D = S3[q] / S0[q];
                a.stat[jh].w = D;
                a.ds[jh] = S1[q] - D * S2[q];
            }
        }
    }
}

// thread per (long row, head): add the segments' partial sums in segment order
__global__ void gat_bwd_row_combine_kernel(const GatBwdArgs a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.rsg.n_long * a.heads) return;
    const int r = (int)(idx / a.heads), hd = (int)(idx % a.heads);
    float S0 = 0.f, S1 = 0.f, S2 = 0.f, S3 = 0.f;
    for (int sgi = a.rsg.long_seg_ptr[r]; sgi < a.rsg.long_seg_ptr[r + 1]; ++sgi) {
        const float4 p = a.p_S[(long long)sgi * a.heads + hd];
        S0 += p.x; S1 += p.y; S2 += p.z; S3 += p.w;
    }
    const long long jh = (long long)a.rsg.long_rows[r] * a.heads + hd;
    const float D = S3 / S0;
    a.stat[jh].w = D;
    a.ds[jh] = S1 - D * S2;
}

// Epilogue of node j once dt_j and the aggregated dH_j are complete: the s = <H, a_self>, t = <H, a_neigh>
// paths (dH_j += ds_j a_self + dt_j a_neigh) and this warp's running sums of ds_j H_j / dt_j H_j.
template <int SLOTS>
__device__ __forceinline__ void gat_bwd_node_epilogue(const GatBwdArgs &a, int j, int lane, const bool *on,
                                                      const float4 *as, const float4 *an, const float4 *hj,
                                                      const float *ds_j, const float *dt, const float4 *dh_acc,
                                                      float4 *gs, float4 *gn) {
#pragma unroll
    for (int q = 0; q < SLOTS; ++q)
        if (on[q]) {
            float4 o = dh_acc[q];
            o.x += ds_j[q] * as[q].x + dt[q] * an[q].x; o.y += ds_j[q] * as[q].y + dt[q] * an[q].y;
            o.z += ds_j[q] * as[q].z + dt[q] * an[q].z; o.w += ds_j[q] * as[q].w + dt[q] * an[q].w;
            *reinterpret_cast<float4 *>(a.dH + (long long)j * a.heads * a.dh + (lane + 32 * q) * 4) = o;
            gs[q].x += ds_j[q] * hj[q].x; gs[q].y += ds_j[q] * hj[q].y; gs[q].z += ds_j[q] * hj[q].z; gs[q].w += ds_j[q] * hj[q].w;
            gn[q].x += dt[q] * hj[q].x; gn[q].y += dt[q] * hj[q].y; gn[q].z += dt[q] * hj[q].z; gn[q].w += dt[q] * hj[q].w;
        }
}

template <int SLOTS>
__device__ __forceinline__ void gat_bwd_cta_partial(const GatBwdArgs &a, float4 (*red)[2][32 * SLOTS], int part_row,
                                                    const float4 *gs, const float4 *gn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int width4 = a.heads * a.dh / 4;
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        red[warp][0][lane + 32 * q] = gs[q];
        red[warp][1][lane + 32 * q] = gn[q];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * 32 * SLOTS; e += 256) {
        const int which = e / (32 * SLOTS), slot = e % (32 * SLOTS);
        if (slot >= width4) continue;
        float4 s = red[0][which][slot];
        for (int w = 1; w < 8; ++w) {
            const float4 v = red[w][which][slot];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4 *>(a.partial + ((long long)part_row * 2 + which) * width4 * 4 + slot * 4) = s;
    }
}

// Column role: for node j, one pass over (a segment of) the rows i that attend to it, gathering dO_i and
// row i's statistics: dt_j = sum_i alpha_ij l'_ij (c_ij x_ij - D_i),  dH_j = sum_i alpha_ij c_ij dO_i.
// Persistent warps stride over the work items; whole-row items run the node epilogue, segment items leave
// partial sums for gat_bwd_col_combine_kernel.
template <int SLOTS>
__global__ void __launch_bounds__(256) gat_bwd_col_kernel(const GatBwdArgs a) {
    __shared__ float4 red[8][2][32 * SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int width4 = a.heads * a.dh / 4, dh4 = a.dh / 4;
    const int hgroups = (a.heads + 3) >> 2;
    const unsigned long long seed = a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0ULL);
    int head[SLOTS];
    bool on[SLOTS];
    float4 as[SLOTS], an[SLOTS], gs[SLOTS], gn[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        on[q] = slot < width4;
        head[q] = on[q] ? slot / dh4 : 0;
        as[q] = an[q] = gs[q] = gn[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on[q]) {
            as[q] = __ldg(reinterpret_cast<const float4 *>(a.a_self) + slot);
            an[q] = __ldg(reinterpret_cast<const float4 *>(a.a_neigh) + slot);
        }
    }
    const int total_warps = gridDim.x * 8;
    const int n_items = a.n_rows + a.csg.n_seg;
    for (int item = blockIdx.x * 8 + warp; item < n_items; item += total_warps) {
        int j, start, end, seg;
        if (!gat_item(a.csg, a.t_indptr, a.n_rows, item, j, start, end, seg)) continue;
        float4 hj[SLOTS], dh_acc[SLOTS];
        float tj[SLOTS], dt_acc[SLOTS];
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) {
            hj[q] = dh_acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            tj[q] = dt_acc[q] = 0.f;
            if (on[q]) {
                hj[q] = __ldg(reinterpret_cast<const float4 *>(a.h + (long long)j * a.ldh) + lane + 32 * q);
                tj[q] = __ldg(a.t + (long long)j * a.heads + head[q]);
            }
        }
        for (int base = start; base < end; base += 32) {
            const int myrow = (base + lane < end) ? __ldg(a.t_indices + base + lane) : 0;
            const int cnt = min(32, end - base);
            for (int kk = 0; kk < cnt; kk += 4) {
                float4 dv[4][SLOTS];
                float4 sv[4][SLOTS];
                int iid[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    iid[u] = __shfl_sync(0xffffffffu, myrow, (kk + u) & 31);
                    if (kk + u < cnt) {
#pragma unroll
                        for (int q = 0; q < SLOTS; ++q)
                            if (on[q]) {
                                dv[u][q] = __ldg(reinterpret_cast<const float4 *>(a.dO + (long long)iid[u] * a.heads * a.dh) + lane + 32 * q);
                                sv[u][q] = __ldg(a.stat + (long long)iid[u] * a.heads + head[q]);
                            }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (kk + u < cnt) {
#pragma unroll
                        for (int q = 0; q < SLOTS; ++q) {
                            float dot = on[q] ? dot4_fixed(dv[u][q], hj[q]) : 0.f;
                            dot = seg_sum(dot, dh4);
                            if (on[q]) {
                                const float pre = sv[u][q].x + tj[q];
                                const float alpha = gat_alpha(pre, a.slope, sv[u][q].y, sv[u][q].z);
                                float c = 1.f;
                                if (a.drop_thr) {
                                    const unsigned long long bits = drop_bits(
                                        seed, ((unsigned long long)iid[u] * a.n_cols + j) * hgroups + (head[q] >> 2));
                                    c = drop_keep(bits, head[q] & 3, a.drop_thr) ? a.drop_scale : 0.f;
                                }
                                dt_acc[q] += alpha * (c * dot - sv[u][q].w) * (pre > 0.f ? 1.f : a.slope);
                                const float w = alpha * c;
                                dh_acc[q].x += w * dv[u][q].x; dh_acc[q].y += w * dv[u][q].y;
                                dh_acc[q].z += w * dv[u][q].z; dh_acc[q].w += w * dv[u][q].w;
                            }
                        }
                    }
                }
            }
        }
        if (seg >= 0) {
#pragma unroll
            for (int q = 0; q < SLOTS; ++q)
                if (on[q]) {
                    const int slot = lane + 32 * q;
                    if ((slot % dh4) == 0) a.p_dt[(long long)seg * a.heads + head[q]] = dt_acc[q];
                    *reinterpret_cast<float4 *>(a.p_dh + ((long long)seg * width4 + slot) * 4) = dh_acc[q];
                }
        } else {
            float ds_j[SLOTS];
#pragma unroll
            for (int q = 0; q < SLOTS; ++q) ds_j[q] = on[q] ? __ldg(a.ds + (long long)j * a.heads + head[q]) : 0.f;
            gat_bwd_node_epilogue<SLOTS>(a, j, lane, on, as, an, hj, ds_j, dt_acc, dh_acc, gs, gn);
        }
    }
    gat_bwd_cta_partial<SLOTS>(a, red, blockIdx.x, gs, gn);
}

// Persistent warps over the long nodes: add the segments' (dt, dH) in segment order, node epilogue; the CTA's
// attention-vector partial goes to row `part_base + blockIdx.x` of the partial buffer.
template <int SLOTS>
__global__ void __launch_bounds__(256) gat_bwd_col_combine_kernel(const GatBwdArgs a, int part_base) {
    __shared__ float4 red[8][2][32 * SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int width4 = a.heads * a.dh / 4, dh4 = a.dh / 4;
    int head[SLOTS];
    bool on[SLOTS];
    float4 as[SLOTS], an[SLOTS], gs[SLOTS], gn[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
        const int slot = lane + 32 * q;
        on[q] = slot < width4;
        head[q] = on[q] ? slot / dh4 : 0;
        as[q] = an[q] = gs[q] = gn[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on[q]) {
            as[q] = __ldg(reinterpret_cast<const float4 *>(a.a_self) + slot);
            an[q] = __ldg(reinterpret_cast<const float4 *>(a.a_neigh) + slot);
        }
    }
    const int total_warps = gridDim.x * 8;
    for (int r = blockIdx.x * 8 + warp; r < a.csg.n_long; r += total_warps) {
        const int j = a.csg.long_rows[r];
        float4 hj[SLOTS], dh_acc[SLOTS];
        float dt_acc[SLOTS], ds_j[SLOTS];
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) {
            hj[q] = dh_acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            dt_acc[q] = ds_j[q] = 0.f;
            if (on[q]) {
                hj[q] = __ldg(reinterpret_cast<const float4 *>(a.h + (long long)j * a.ldh) + lane + 32 * q);
                ds_j[q] = __ldg(a.ds + (long long)j * a.heads + head[q]);
            }
        }
        for (int sgi = a.csg.long_seg_ptr[r]; sgi < a.csg.long_seg_ptr[r + 1]; ++sgi) {
#pragma unroll
            for (int q = 0; q < SLOTS; ++q)
                if (on[q]) {
                    dt_acc[q] += __ldg(a.p_dt + (long long)sgi * a.heads + head[q]);
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(a.p_dh) + (long long)sgi * width4 + lane + 32 * q);
                    dh_acc[q].x += v.x; dh_acc[q].y += v.y; dh_acc[q].z += v.z; dh_acc[q].w += v.w;
                }
        }
        gat_bwd_node_epilogue<SLOTS>(a, j, lane, on, as, an, hj, ds_j, dt_acc, dh_acc, gs, gn);
    }
    gat_bwd_cta_partial<SLOTS>(a, red, part_base + blockIdx.x, gs, gn);
}

static SegTable seg_table(const gr_gat_segments *s) {
    SegTable t;
    if (!s || s->n_seg <= 0) {
        t.seg_len = 0; t.n_seg = 0; t.n_long = 0;
        t.seg_row = t.seg_begin = t.seg_end = t.long_rows = t.long_seg_ptr = nullptr;
        return t;
    }
    t.seg_len = s->seg_len; t.n_seg = s->n_seg; t.n_long = s->n_long;
    t.seg_row = s->seg_row; t.seg_begin = s->seg_begin; t.seg_end = s->seg_end;
    t.long_rows = s->long_rows; t.long_seg_ptr = s->long_seg_ptr;
    return t;
}
static bool seg_valid(const gr_gat_segments *s) {
    if (!s || s->n_seg <= 0) return true;
    return s->seg_len > 0 && s->n_long > 0 && s->seg_row && s->seg_begin && s->seg_end && s->long_rows && s->long_seg_ptr;
}
static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace gr

using namespace gr;

extern "C" int gr_gat_node_scores(const float *h, int64_t ldh, const float *a_self, const float *a_neigh,
                                  int64_t n_rows, int32_t heads, int32_t dh, float *s, float *t, void *stream) {
    if (!h || !a_self || !a_neigh || !s || !t || n_rows < 0 || heads <= 0 || dh <= 0 || ldh < (int64_t)heads * dh)
        return GR_ERR_INVALID;
    if (n_rows == 0) return GR_OK;
    if (n_rows > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    gat_node_scores_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        h, ldh, a_self, a_neigh, (int)n_rows, heads, dh, s, t);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" size_t gr_gat_aggregate_workspace_bytes(int32_t n_seg, int32_t heads, int32_t dh) {
    if (n_seg <= 0 || heads <= 0 || dh <= 0) return 256;
    return 2 * al256((size_t)n_seg * heads * 4) + al256((size_t)n_seg * heads * dh * 4) + 256;
}

extern "C" int gr_gat_aggregate(const int32_t *indptr, const int32_t *indices, int64_t n_rows, const float *h,
                                int64_t ldh, const float *s, const float *t, int32_t heads, int32_t dh,
                                float slope, int32_t mean_heads, int32_t elu, float drop_p, uint64_t drop_seed,
                                const uint64_t *drop_seed_dev, int64_t n_cols, const gr_gat_segments *segs_host, float *out, int64_t ldo,
                                float *m_out, float *z_out, void *workspace, size_t workspace_bytes, void *stream) {
    if (!indptr || !indices || !h || !s || !t || !out || n_rows < 0 || heads <= 0 || dh <= 0) return GR_ERR_INVALID;
    if (!(drop_p >= 0.f) || drop_p >= 1.f || n_cols < 0) return GR_ERR_INVALID;
    if ((m_out == nullptr) != (z_out == nullptr)) return GR_ERR_INVALID;
    if (!seg_valid(segs_host)) return GR_ERR_INVALID;
    if (n_rows == 0) return GR_OK;
    const int width = heads * dh;
    if ((dh & 3) || width > 256 || (ldh & 3) || (ldo & 3) || ldh < width) return GR_ERR_UNSUPPORTED;
    if (ldo < (mean_heads ? dh : width)) return GR_ERR_INVALID;
    if (!aligned16(h) || !aligned16(out)) return GR_ERR_INVALID;
    if (n_rows > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    GatArgs a;
    a.indptr = indptr; a.indices = indices; a.h = h; a.ldh = ldh; a.s = s; a.t = t;
    a.n_rows = (int)n_rows; a.heads = heads; a.dh = dh; a.slope = slope; a.mean_heads = mean_heads; a.elu = elu;
    a.out = out; a.ldo = ldo; a.m_out = m_out; a.z_out = z_out;
    a.n_cols = n_cols;
    a.drop_thr = drop_threshold(drop_p); a.drop_scale = drop_scale_of(a.drop_thr); a.drop_seed = drop_seed;
    a.drop_seed_dev = reinterpret_cast<const unsigned long long *>(drop_seed_dev);
    a.sg = seg_table(segs_host);
    a.p_m = a.p_z = a.p_acc = nullptr;
    if (a.sg.n_seg > 0) {
        if (!workspace || !aligned16(workspace) ||
            workspace_bytes < gr_gat_aggregate_workspace_bytes(a.sg.n_seg, heads, dh))
            return GR_ERR_WORKSPACE;
        char *ws = static_cast<char *>(workspace);
        a.p_m = reinterpret_cast<float *>(ws); ws += al256((size_t)a.sg.n_seg * heads * 4);
        a.p_z = reinterpret_cast<float *>(ws); ws += al256((size_t)a.sg.n_seg * heads * 4);
        a.p_acc = reinterpret_cast<float *>(ws);
    }
    const long long items = n_rows + a.sg.n_seg;
    const unsigned grid = (unsigned)((items + 7) / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (width / 4 <= 32) gat_aggregate_kernel<1><<<grid, 256, 0, st>>>(a);
    else gat_aggregate_kernel<2><<<grid, 256, 0, st>>>(a);
    GR_LAUNCH_CHECK();
    if (a.sg.n_seg > 0) {
        const unsigned cg = (unsigned)((a.sg.n_long + 7) / 8);
        if (width / 4 <= 32) gat_aggregate_combine_kernel<1><<<cg, 256, 0, st>>>(a);
        else gat_aggregate_combine_kernel<2><<<cg, 256, 0, st>>>(a);
        GR_LAUNCH_CHECK();
    }
    return GR_OK;
}

static int gat_bwd_ctas(long long items) { return persistent_grid((int)((items + 7) / 8), 8); }
static int gat_bwd_comb_ctas(int n_long) { return n_long > 0 ? persistent_grid((n_long + 7) / 8, 8) : 0; }

extern "C" size_t gr_gat_bwd_workspace_bytes(int64_t n_rows, int32_t heads, int32_t dh, int32_t n_seg_row,
                                             int32_t n_seg_col, int32_t n_long_col) {
    if (n_rows <= 0 || heads <= 0 || dh <= 0) return 256;
    const size_t width = (size_t)heads * dh;
    if (n_seg_row < 0) n_seg_row = 0;
    if (n_seg_col < 0) n_seg_col = 0;
    size_t b = al256((size_t)n_rows * width * 4) + al256((size_t)n_rows * heads * 16) + al256((size_t)n_rows * heads * 4);
    b += al256((size_t)n_seg_row * heads * 16) + al256((size_t)n_seg_col * heads * 4) + al256((size_t)n_seg_col * width * 4);
    b += (size_t)(gat_bwd_ctas(n_rows + n_seg_col) + gat_bwd_comb_ctas(n_seg_col > 0 ? n_long_col : 0)) * 2 * width * 4;
    return b + 256;
}

extern "C" int gr_gat_bwd(const int32_t *indptr, const int32_t *indices, const int32_t *t_indptr,
                          const int32_t *t_indices, int64_t n_rows, int64_t n_cols, const float *h, int64_t ldh,
                          const float *s, const float *t, const float *m, const float *z, const float *out,
                          int64_t ldo, const float *dout, int64_t lddo, const float *a_self, const float *a_neigh,
                          int32_t heads, int32_t dh, float slope, int32_t mean_heads, int32_t elu, float drop_p,
                          uint64_t drop_seed, const uint64_t *drop_seed_dev, const gr_gat_segments *row_segs_host,
                          const gr_gat_segments *col_segs_host, float *dH, float *da, void *workspace,
                          size_t workspace_bytes, void *stream) {
    if (!indptr || !indices || !t_indptr || !t_indices || !h || !s || !t || !m || !z || !dout || !a_self ||
        !a_neigh || !dH || !da || n_rows < 0 || n_cols < 0 || heads <= 0 || dh <= 0)
        return GR_ERR_INVALID;
    if (elu && !out) return GR_ERR_INVALID;
    if (!(drop_p >= 0.f) || drop_p >= 1.f) return GR_ERR_INVALID;
    if (!seg_valid(row_segs_host) || !seg_valid(col_segs_host)) return GR_ERR_INVALID;
    if (n_rows != n_cols) return GR_ERR_UNSUPPORTED;     // the column role indexes the same node set
    const int width = heads * dh;
    const int dh4 = dh / 4;
    if ((dh & 3) || width > 256 || (dh4 & (dh4 - 1)) || dh4 > 32) return GR_ERR_UNSUPPORTED;
    const int wout = mean_heads ? dh : width;
    if ((ldh & 3) || ldh < width || (lddo & 3) || lddo < wout) return GR_ERR_INVALID;
    if (elu && ((ldo & 3) || ldo < wout)) return GR_ERR_INVALID;
    if (!aligned16(h) || !aligned16(out) || !aligned16(dout) || !aligned16(a_self) ||
        !aligned16(a_neigh) || !aligned16(dH) || !aligned16(da) || !aligned16(workspace))
        return GR_ERR_INVALID;
    if (n_rows > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_rows == 0) {
        GR_CUDA_CHECK(cudaMemsetAsync(da, 0, (size_t)2 * width * 4, st));
        return GR_OK;
    }
    const SegTable rsg = seg_table(row_segs_host), csg = seg_table(col_segs_host);
    if (workspace_bytes < gr_gat_bwd_workspace_bytes(n_rows, heads, dh, rsg.n_seg, csg.n_seg, csg.n_long))
        return GR_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    float *dO = reinterpret_cast<float *>(ws); ws += al256((size_t)n_rows * width * 4);
    float4 *stat = reinterpret_cast<float4 *>(ws); ws += al256((size_t)n_rows * heads * 16);
    float *ds = reinterpret_cast<float *>(ws); ws += al256((size_t)n_rows * heads * 4);
    float4 *p_S = reinterpret_cast<float4 *>(ws); ws += al256((size_t)rsg.n_seg * heads * 16);
    float *p_dt = reinterpret_cast<float *>(ws); ws += al256((size_t)csg.n_seg * heads * 4);
    float *p_dh = reinterpret_cast<float *>(ws); ws += al256((size_t)csg.n_seg * width * 4);
    float *partial = reinterpret_cast<float *>(ws);

    GatPrepArgs p;
    p.dout = dout; p.out = out; p.s = s; p.m = m; p.z = z;
    p.lddo = lddo; p.ldo = ldo; p.dO = dO; p.stat = stat;
    p.n_rows = (int)n_rows; p.heads = heads; p.dh = dh; p.mean_heads = mean_heads; p.elu = elu;
    const unsigned rows_grid = (unsigned)((n_rows + 7) / 8);
    const bool one = width / 4 <= 32;
    if (one) gat_bwd_prep_kernel<1><<<rows_grid, 256, 0, st>>>(p);
    else gat_bwd_prep_kernel<2><<<rows_grid, 256, 0, st>>>(p);
    GR_LAUNCH_CHECK();

    GatBwdArgs a;
    a.indptr = indptr; a.indices = indices; a.t_indptr = t_indptr; a.t_indices = t_indices;
    a.h = h; a.dO = dO; a.t = t; a.a_self = a_self; a.a_neigh = a_neigh; a.stat = stat; a.ds = ds; a.ldh = ldh;
    a.dH = dH; a.partial = partial; a.n_cols = n_cols; a.n_rows = (int)n_rows; a.heads = heads; a.dh = dh;
    a.slope = slope;
    a.drop_thr = drop_threshold(drop_p); a.drop_scale = drop_scale_of(a.drop_thr); a.drop_seed = drop_seed;
    a.drop_seed_dev = reinterpret_cast<const unsigned long long *>(drop_seed_dev);
    a.rsg = rsg; a.csg = csg; a.p_S = p_S; a.p_dt = p_dt; a.p_dh = p_dh;
    const unsigned row_items_grid = (unsigned)((n_rows + rsg.n_seg + 7) / 8);
    if (one) gat_bwd_row_kernel<1><<<row_items_grid, 256, 0, st>>>(a);
    else gat_bwd_row_kernel<2><<<row_items_grid, 256, 0, st>>>(a);
    GR_LAUNCH_CHECK();
    if (rsg.n_seg > 0) {
        const long long tot = (long long)rsg.n_long * heads;
        gat_bwd_row_combine_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(a);
        GR_LAUNCH_CHECK();
    }
    const int ctas = gat_bwd_ctas(n_rows + csg.n_seg);
    if (one) gat_bwd_col_kernel<1><<<ctas, 256, 0, st>>>(a);
    else gat_bwd_col_kernel<2><<<ctas, 256, 0, st>>>(a);
    GR_LAUNCH_CHECK();
    int cctas = 0;
    if (csg.n_seg > 0) {
        cctas = gat_bwd_comb_ctas(csg.n_long);
        if (one) gat_bwd_col_combine_kernel<1><<<cctas, 256, 0, st>>>(a, ctas);
        else gat_bwd_col_combine_kernel<2><<<cctas, 256, 0, st>>>(a, ctas);
        GR_LAUNCH_CHECK();
    }
    if (launch_reduce_partials(partial, ctas + cctas, 2LL * width, da, st)) {
        set_last_cuda_error(cudaPeekAtLastError());
        return GR_ERR_CUDA;
    }
    return GR_OK;
}
