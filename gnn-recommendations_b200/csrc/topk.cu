// Full-ranking evaluation: fused score GEMM + seen-item mask + per-user top-K, no score matrix.
//
// Replaces the loop body of Evaluator.evaluate / Trainer.validate
// (src/evaluation/evaluator.py:96-106, src/training/trainer.py:327-337):
//     scores = user_emb[batch] @ item_emb.T ; scores[row, seen(user)] = -inf ; topk(scores, K)
//
// Exactness.  Every score is the k-sequential fmaf chain  s = fma(u[k], i[k], s), k = 0..d-1,
// from zero — what the reference's CPU sgemm produces bit-for-bit on these shapes — so ranking
// ties and near-ties resolve exactly as in the reference; lists are ordered by (score desc,
// item id asc), the canonical order the parity tests derive from the reference's scores
// (torch.topk's own tie order is arbitrary).  fp32 has no tensor-core MMA (tcgen05 offers TF32
// at best), so this exact pass runs on the FFMA pipe; a TF32 candidate pre-filter is the planned
// next step.
//
// Structure.  CTA = 64 eval users x one item range; user tile resident in shared memory
// (k-major), item tiles of 128 streamed through a k-major shared tile, 4x8 register block per
// thread (rank-1 updates in k order), scores of the tile dropped to shared memory, masked from
// the user's sorted seen list (a cursor walks it as item tiles ascend), filtered against the
// running K-th score and inserted warp-cooperatively into a sorted shared-memory list.  Items
// ascend, so equal scores keep ascending ids without comparing ids.
#include <math_constants.h>

#include "gr_common.cuh"

namespace gr {

constexpr int TU = 64;    // users per CTA
constexpr int TI = 128;   // items per tile
constexpr int KC = 16;    // k-chunk staged per step
constexpr int KMAX = 64;  // largest supported K

struct TopkArgs {
    const float *user_emb;
    long long ldu;
    const float *item_emb;  // row 0 = item id item_lo
    long long ldi;
    int d;
    const int64_t *eval_users;
    int n_eval;
    long long item_lo, item_hi;    // this call's item id range
    const int64_t *seen_indptr;    // [n_eval + 1] or NULL
    const int32_t *seen_items;     // sorted ids per eval row
    int k;
    int items_per_split;           // multiple of TI
    float *part_scores;            // [n_splits][n_eval][k]
    int *part_ids;
};

__global__ void __launch_bounds__(256) score_topk_kernel(const TopkArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d;
    float *Au = reinterpret_cast<float *>(smem_raw);          // [d][TU]
    float *Bs = Au + (size_t)d * TU;                          // [KC][TI]
    float *S = Bs + KC * TI;                                  // [TU][TI + 1]
    float *ls = S + TU * (TI + 1);                            // [TU][KMAX] list scores
    int *li = reinterpret_cast<int *>(ls + TU * KMAX);        // [TU][KMAX] list ids
    int *cnt_s = li + TU * KMAX;                              // [TU]
    int *cur_s = cnt_s + TU;                                  // [TU] cursor into the seen list
    int *end_s = cur_s + TU;                                  // [TU]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * TU;
    const long long i_begin = a.item_lo + (long long)blockIdx.y * a.items_per_split;
    const long long i_end = min(a.item_hi, i_begin + (long long)a.items_per_split);
    const int K = a.k;

    // ---- user tile -> shared memory, k-major ------------------------------------------------
    for (int idx = tid; idx < TU * (d / 4); idx += 256) {
        const int r = idx / (d / 4), f = idx % (d / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < a.n_eval) {
            const long long u = a.eval_users[row0 + r];
            v = __ldg(reinterpret_cast<const float4 *>(a.user_emb + u * a.ldu) + f);
        }
        Au[(4 * f + 0) * TU + r] = v.x;
        Au[(4 * f + 1) * TU + r] = v.y;
        Au[(4 * f + 2) * TU + r] = v.z;
        Au[(4 * f + 3) * TU + r] = v.w;
    }
    if (tid < TU) {
        cnt_s[tid] = 0;
        int c = 0, e = 0;
        if (a.seen_indptr && row0 + tid < a.n_eval) {
            long long lo = a.seen_indptr[row0 + tid], hi = a.seen_indptr[row0 + tid + 1];
            e = (int)hi;
            // first seen id >= i_begin
            while (lo < hi) {
                const long long m = (lo + hi) >> 1;
                if (a.seen_items[m] < i_begin) lo = m + 1; else hi = m;
            }
            c = (int)lo;
        }
        cur_s[tid] = c;
        end_s[tid] = e;
    }
    __syncthreads();

    for (long long i0 = i_begin; i0 < i_end; i0 += TI) {
        float acc[4][8];
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) acc[m][n] = 0.f;

        for (int k0 = 0; k0 < d; k0 += KC) {
            // item chunk [TI items][KC k] -> Bs[k][item]; 128*16/4 = 512 float4, 2 per thread
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int idx = tid + 256 * t;
                const int it = idx >> 2, f = idx & 3;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i0 + it < i_end)
                    v = __ldg(reinterpret_cast<const float4 *>(a.item_emb + (i0 + it - a.item_lo) * a.ldi + k0) + f);
                Bs[(4 * f + 0) * TI + it] = v.x;
                Bs[(4 * f + 1) * TI + it] = v.y;
                Bs[(4 * f + 2) * TI + it] = v.z;
                Bs[(4 * f + 3) * TI + it] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const float4 av = *reinterpret_cast<const float4 *>(Au + (size_t)(k0 + kk) * TU + ty * 4);
                const float4 b0 = *reinterpret_cast<const float4 *>(Bs + kk * TI + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4 *>(Bs + kk * TI + 64 + tx * 4);
                const float am[4] = {av.x, av.y, av.z, av.w};
                const float bn[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 8; ++n) acc[m][n] = __fmaf_rn(am[m], bn[n], acc[m][n]);
            }
            __syncthreads();
        }
        // ---- scores -> shared tile ------------------------------------------------------------
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int il = (n < 4) ? tx * 4 + n : 64 + tx * 4 + (n - 4);
                S[(ty * 4 + m) * (TI + 1) + il] = acc[m][n];
            }
        __syncthreads();

        // ---- mask + select: warp w owns user rows w*8 .. w*8+7 -------------------------------------
        for (int rr = 0; rr < 8; ++rr) {
            const int r = warp * 8 + rr;
            if (row0 + r >= a.n_eval) break;
            float *Sr = S + r * (TI + 1);
            {   // seen items inside [i0, i0 + TI) -> -inf
                int p = cur_s[r];
                const int e = end_s[r];
                while (true) {
                    const int q = p + lane;
                    bool in = false;
                    if (q < e) {
                        const long long it = a.seen_items[q];
                        in = it < i0 + TI;
                        if (in && it >= i0) Sr[it - i0] = -CUDART_INF_F;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    p += __popc(m);
                    if (m != 0xffffffffu) break;
                }
                __syncwarp();
                if (lane == 0) cur_s[r] = p;
            }
            int cnt = cnt_s[r];
            float kth = (cnt == K) ? ls[r * KMAX + K - 1] : 0.f;
            float *lsr = ls + r * KMAX;
            int *lir = li + r * KMAX;
#pragma unroll 1
            for (int q = 0; q < TI / 32; ++q) {
                const int il = q * 32 + lane;
                const float s = Sr[il];
                const bool valid = (i0 + il) < i_end;
                unsigned m = __ballot_sync(0xffffffffu, valid && (cnt < K || s > kth));
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float sv = __shfl_sync(0xffffffffu, s, src);
                    if (!(cnt < K || sv > kth)) continue;  // threshold moved since the ballot
                    const int iv = (int)(i0 + q * 32 + src);
                    // position = number of list entries >= sv (equal scores already hold smaller ids)
                    const int t0 = lane, t1 = lane + 32;
                    const float o0 = t0 < cnt ? lsr[t0] : 0.f, o1 = t1 < cnt ? lsr[t1] : 0.f;
                    const int oi0 = t0 < cnt ? lir[t0] : 0, oi1 = t1 < cnt ? lir[t1] : 0;
                    const int pos = __popc(__ballot_sync(0xffffffffu, t0 < cnt && o0 >= sv)) +
                                    __popc(__ballot_sync(0xffffffffu, t1 < cnt && o1 >= sv));
                    const int ncnt = min(cnt + 1, K);
                    __syncwarp();
                    // entries [pos, ncnt-1) move one slot right: slot t+1 <- old t
                    if (t0 >= pos && t0 + 1 < ncnt) { lsr[t0 + 1] = o0; lir[t0 + 1] = oi0; }
                    if (t1 >= pos && t1 + 1 < ncnt) { lsr[t1 + 1] = o1; lir[t1 + 1] = oi1; }
                    if (lane == 0 && pos < ncnt) { lsr[pos] = sv; lir[pos] = iv; }
                    __syncwarp();
                    cnt = ncnt;
                    if (cnt == K) kth = lsr[K - 1];
                }
            }
            if (lane == 0) cnt_s[r] = cnt;
        }
        __syncthreads();
    }

    // ---- write this split's lists ---------------------------------------------------------------
    for (int idx = tid; idx < TU * K; idx += 256) {
        const int r = idx / K, t = idx % K;
        if (row0 + r < a.n_eval) {
            const bool have = t < cnt_s[r];
            const size_t o = ((size_t)blockIdx.y * a.n_eval + row0 + r) * K + t;
            a.part_scores[o] = have ? ls[r * KMAX + t] : -CUDART_INF_F;
            a.part_ids[o] = have ? li[r * KMAX + t] : -1;
        }
    }
}

// Merge P partial lists per user (each canonical, ids disjoint across parts) into the top-K.
// One warp per user; rank by counting predecessors under (score desc, id asc).
__global__ void __launch_bounds__(256) topk_merge_kernel(const float *part_scores, const int *part_ids, int n_parts,
                                                         int n_eval, int k, int64_t *out_ids, float *out_scores) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 8 + warp;
    const int total = n_parts * k;
    float *cs = reinterpret_cast<float *>(smem_raw) + (size_t)warp * total;
    int *ci = reinterpret_cast<int *>(smem_raw + (size_t)8 * total * 4) + (size_t)warp * total;
    if (row >= n_eval) return;
    for (int c = lane; c < total; c += 32) {
        const int p = c / k, t = c % k;
        const size_t o = ((size_t)p * n_eval + row) * k + t;
        cs[c] = part_scores[o];
        ci[c] = part_ids[o];
    }
    __syncwarp();
    for (int t = lane; t < k; t += 32) {  // defaults for rows with fewer than k candidates overall
        out_ids[(size_t)row * k + t] = -1;
        out_scores[(size_t)row * k + t] = -CUDART_INF_F;
    }
    __syncwarp();
    for (int c = lane; c < total; c += 32) {
        const int id = ci[c];
        if (id < 0) continue;
        const float s = cs[c];
        int rank = 0;
        for (int o = 0; o < total; ++o) {
            const int oid = ci[o];
            const float os = cs[o];
            rank += (oid >= 0 && (os > s || (os == s && oid < id))) ? 1 : 0;
        }
        if (rank < k) {
            out_ids[(size_t)row * k + rank] = id;
            out_scores[(size_t)row * k + rank] = s;
        }
    }
}

static size_t topk_smem_bytes(int d) {
    return ((size_t)d * TU + KC * TI + TU * (TI + 1) + TU * KMAX) * 4 + (TU * KMAX + 3 * TU) * 4;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_score_topk_workspace_bytes(int64_t n_eval, int32_t k, int32_t n_splits) {
    if (n_eval < 0 || k <= 0 || n_splits <= 0) return 0;
    return (size_t)n_splits * (size_t)n_eval * (size_t)k * 8 + 256;
}

extern "C" int gr_topk_merge(const float *part_scores, const int32_t *part_ids, int32_t n_parts, int64_t n_eval,
                             int32_t k, int64_t *out_ids, float *out_scores, void *stream) {
    if (!part_scores || !part_ids || !out_ids || !out_scores || n_parts <= 0 || n_eval < 0 || k <= 0 || k > KMAX)
        return GR_ERR_INVALID;
    if (n_eval == 0) return GR_OK;
    if (n_eval > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    const size_t smem = (size_t)8 * n_parts * k * 8;
    if (smem > 200 * 1024) return GR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        GR_CUDA_CHECK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<(unsigned)((n_eval + 7) / 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        part_scores, part_ids, n_parts, (int)n_eval, k, out_ids, out_scores);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

// Partial lists only (item-sharded multi-GPU: each rank calls this on its item range, then the
// gathered partials go through gr_topk_merge).  part_* hold n_splits * n_eval * k entries.
extern "C" int gr_score_topk_partial(const float *user_emb, int64_t ldu, const float *item_emb, int64_t ldi,
                                     int32_t d, const int64_t *eval_users, int64_t n_eval, int64_t item_lo,
                                     int64_t item_hi, const int64_t *seen_indptr, const int32_t *seen_items,
                                     int32_t k, int32_t n_splits, float *part_scores, int32_t *part_ids,
                                     void *stream) {
    if (!user_emb || !item_emb || !eval_users || !part_scores || !part_ids) return GR_ERR_INVALID;
    if (n_eval < 0 || item_hi < item_lo || k <= 0 || k > KMAX || n_splits <= 0 || n_splits > 65535) return GR_ERR_INVALID;
    if (d <= 0 || (d % KC) || d > 512 || (ldu & 3) || (ldi & 3) || ldu < d || ldi < d) return GR_ERR_UNSUPPORTED;
    if (!aligned16(user_emb) || !aligned16(item_emb)) return GR_ERR_INVALID;
    if ((seen_indptr == nullptr) != (seen_items == nullptr)) return GR_ERR_INVALID;
    if (n_eval == 0) return GR_OK;
    if (n_eval > 0x7fffffffLL || item_hi > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    TopkArgs a;
    a.user_emb = user_emb;
    a.ldu = ldu;
    a.item_emb = item_emb;
    a.ldi = ldi;
    a.d = d;
    a.eval_users = eval_users;
    a.n_eval = (int)n_eval;
    a.item_lo = item_lo;
    a.item_hi = item_hi;
    a.seen_indptr = seen_indptr;
    a.seen_items = seen_items;
    a.k = k;
    const long long n_it = item_hi - item_lo;
    long long per = (n_it + n_splits - 1) / n_splits;
    per = (per + TI - 1) / TI * TI;
    if (per < TI) per = TI;
    a.items_per_split = (int)per;
    a.part_scores = part_scores;
    a.part_ids = part_ids;
    const size_t smem = topk_smem_bytes(d);
    if (smem > 227 * 1024) return GR_ERR_UNSUPPORTED;
    GR_CUDA_CHECK(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n_eval + TU - 1) / TU), (unsigned)n_splits);
    score_topk_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" int gr_score_topk(const float *user_emb, int64_t ldu, const float *item_emb, int64_t ldi, int32_t d,
                             const int64_t *eval_users, int64_t n_eval, int64_t n_items,
                             const int64_t *seen_indptr, const int32_t *seen_items, int32_t k, int32_t n_splits,
                             int64_t *topk_ids, float *topk_scores, void *workspace, size_t workspace_bytes,
                             void *stream) {
    if (!topk_ids || !topk_scores || !workspace) return GR_ERR_INVALID;
    if (n_splits <= 0) return GR_ERR_INVALID;
    if (workspace_bytes < gr_score_topk_workspace_bytes(n_eval, k, n_splits)) return GR_ERR_WORKSPACE;
    float *ps = static_cast<float *>(workspace);
    int32_t *pi = reinterpret_cast<int32_t *>(ps + (size_t)n_splits * n_eval * k);
    int rc = gr_score_topk_partial(user_emb, ldu, item_emb, ldi, d, eval_users, n_eval, 0, n_items, seen_indptr,
                                   seen_items, k, n_splits, ps, pi, stream);
    if (rc != GR_OK) return rc;
    return gr_topk_merge(ps, pi, n_splits, n_eval, k, topk_ids, topk_scores, stream);
}
