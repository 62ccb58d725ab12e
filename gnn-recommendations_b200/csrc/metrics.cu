// On-device reduction of the top-K lists to ranking metrics.
//
// Replaces the python loops of compute_metrics_from_topk (src/training/metrics.py:355-432): per user
// hits = |relevant ∩ top-k|, recall = hits/|relevant|, precision = hits/k, DCG = sum over hit ranks of
// 1/log2(rank+2) accumulated in rank order, NDCG = DCG/IDCG(min(|relevant|, k)); users without ground
// truth are skipped (metrics.py:390-394).  Everything per-user is float64 with the reference's operation
// order; the discount and ideal-DCG tables come from the host (numpy log2, metrics.py:402-409) so the
// per-user values carry the same bits.  Coverage and Gini (metrics.py:415-430) are integer work: item
// counts by atomics, counts sorted with the library's radix sort, sum (i+1)*c_(i) in int64 — exact.
// The host finishes with the handful of scalar operations on the returned sums.
#include "gr_common.cuh"

namespace gr {

constexpr int MET_MAX_K = 64;     // list length handled by one warp (two ids per lane)
constexpr int MET_MAX_NK = 8;

struct MetricsArgs {
    const int64_t *topk;      // [n_eval][kmax]
    int kmax;
    long long n_eval;
    const int64_t *gt_indptr; // [n_eval + 1]
    const int32_t *gt_items;  // sorted, unique per row
    long long n_items;
    int nk;
    int ks[MET_MAX_NK];       // already min(k, kmax)
    const double *disc;       // [kmax]   1/log2(rank+2)
    const double *idcg;       // [kmax+1] ideal DCG of the first j ranks
    double *row_vals;         // [nk][3][n_eval]  recall, ndcg, precision
    int *row_valid;           // [n_eval]
    int *item_counts;         // [nk][n_items + 1]  (slot n_items counts the -1 padding ids)
};

__global__ void __launch_bounds__(256) metrics_rows_kernel(const MetricsArgs a) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= a.n_eval) return;
    const long long g0 = a.gt_indptr[row], g1 = a.gt_indptr[row + 1];
    const long long n_rel = g1 - g0;
    unsigned long long hit = 0;
    long long ids[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q;
        ids[q] = c < a.kmax ? a.topk[row * a.kmax + c] : -1;
        bool found = false;
        if (c < a.kmax && ids[q] >= 0) {
            long long lo = g0, hi = g1;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if ((long long)a.gt_items[mid] < ids[q]) lo = mid + 1; else hi = mid;
            }
            found = lo < g1 && (long long)a.gt_items[lo] == ids[q];
        }
        hit |= (unsigned long long)__ballot_sync(0xffffffffu, found) << (32 * q);
    }
    if (lane == 0) a.row_valid[row] = n_rel > 0 ? 1 : 0;
    for (int ki = 0; ki < a.nk; ++ki) {
        const int k = a.ks[ki];
        // item counts cover every row (metrics.py:416-421), metrics only rows with ground truth
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = lane + 32 * q;
            if (c < k) {
                const long long slot = (ids[q] < 0 || ids[q] >= a.n_items) ? a.n_items : ids[q];
                atomicAdd(&a.item_counts[(long long)ki * (a.n_items + 1) + slot], 1);
            }
        }
        if (lane == 0) {
            double recall = 0.0, ndcg = 0.0, precision = 0.0;
            if (n_rel > 0) {
                unsigned long long hk = k >= 64 ? hit : (hit & ((1ULL << k) - 1ULL));
                const int hits = __popcll(hk);
                recall = (double)hits / (double)n_rel;
                precision = (double)hits / (double)k;
                double dcg = 0.0;
                while (hk) {                       // ascending rank, as the reference's loop
                    const int rank = __ffsll((long long)hk) - 1;
                    hk &= hk - 1;
                    dcg += a.disc[rank];
                }
                const double ideal = a.idcg[n_rel < k ? n_rel : k];
                ndcg = ideal > 0.0 ? dcg / ideal : 0.0;
            }
            a.row_vals[((long long)ki * 3 + 0) * a.n_eval + row] = recall;
            a.row_vals[((long long)ki * 3 + 1) * a.n_eval + row] = ndcg;
            a.row_vals[((long long)ki * 3 + 2) * a.n_eval + row] = precision;
        }
    }
}

// block b < n_series: fixed-order sum of row_vals[b][0..n_eval); block n_series: number of valid rows
__global__ void __launch_bounds__(1024) metrics_reduce_kernel(const double *row_vals, const int *row_valid, long long n_eval,
                                                              int n_series, double *out) {
    __shared__ double ws[32];
    double acc = 0.0;
    if ((int)blockIdx.x < n_series) {
        const double *src = row_vals + (long long)blockIdx.x * n_eval;
        for (long long i = threadIdx.x; i < n_eval; i += 1024) acc += src[i];
    } else {
        for (long long i = threadIdx.x; i < n_eval; i += 1024) acc += (double)row_valid[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += ws[w];
        out[blockIdx.x] = s;
    }
}

// keys[i] = count of item i (the -1 slot is folded into the last item, as numpy's item_counts[-1]);
// counts[0] += number of distinct recommended ids (incl. -1), counts[1] += total recommendations
__global__ void metrics_count_keys_kernel(const int *item_counts, long long n_items, uint64_t *keys,
                                          unsigned long long *counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long uniq = 0, tot = 0;
    if (i <= n_items) {
        const int c = item_counts[i];
        uniq = c > 0;
        tot = (unsigned long long)c;
        if (i < n_items) keys[i] = (uint64_t)c + (i == n_items - 1 ? (uint64_t)item_counts[n_items] : 0);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        uniq += __shfl_xor_sync(0xffffffffu, uniq, o);
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if ((threadIdx.x & 31) == 0 && (uniq | tot)) {
        atomicAdd(&counts[0], uniq);
        atomicAdd(&counts[1], tot);
    }
}

// counts[2] += sum (i + 1) * sorted[i]
__global__ void metrics_gini_sum_kernel(const uint64_t *sorted, long long n, unsigned long long *counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = i < n ? (unsigned long long)(i + 1) * sorted[i] : 0ULL;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[2], v);
}

static inline size_t met_align(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int met_bits(uint64_t v) {
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

}  // namespace gr

using namespace gr;

extern "C" size_t gr_topk_metrics_workspace_bytes(int64_t n_eval, int64_t n_items, int32_t nk) {
    if (n_eval < 0 || n_items <= 0 || nk <= 0 || nk > MET_MAX_NK) return 0;
    return met_align((size_t)nk * 3 * n_eval * 8) + met_align((size_t)n_eval * 4) +
           met_align((size_t)nk * (n_items + 1) * 4) + 2 * met_align((size_t)n_items * 8) +
           met_align(radix_sort_workspace_bytes(n_items)) + 256;
}

/* out_sums [nk*3 + 1] doubles: per k (recall, ndcg, precision) sums over users with ground truth, then
 * the number of such users.  out_counts [nk*3] int64: per k (distinct ids, total, sum (i+1) c_(i)). */
extern "C" int gr_topk_metrics(const int64_t *topk_ids, int64_t n_eval, int32_t kmax, const int64_t *gt_indptr,
                               const int32_t *gt_items, int64_t n_items, const int32_t *k_values_host, int32_t nk,
                               const double *disc, const double *idcg, double *out_sums, int64_t *out_counts,
                               void *workspace, size_t workspace_bytes, void *stream) {
    if (!topk_ids || !gt_indptr || !gt_items || !k_values_host || !disc || !idcg || !out_sums || !out_counts || !workspace)
        return GR_ERR_INVALID;
    if (n_eval <= 0 || n_items <= 0 || kmax <= 0 || nk <= 0) return GR_ERR_INVALID;
    if (kmax > MET_MAX_K || nk > MET_MAX_NK) return GR_ERR_UNSUPPORTED;
    if (n_items >= (1LL << 31)) return GR_ERR_OVERFLOW;
    if (workspace_bytes < gr_topk_metrics_workspace_bytes(n_eval, n_items, nk)) return GR_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char *w = static_cast<char *>(workspace);
    MetricsArgs a;
    a.topk = topk_ids; a.kmax = kmax; a.n_eval = n_eval; a.gt_indptr = gt_indptr; a.gt_items = gt_items;
    a.n_items = n_items; a.nk = nk; a.disc = disc; a.idcg = idcg;
    for (int i = 0; i < nk; ++i) {
        if (k_values_host[i] <= 0) return GR_ERR_INVALID;
        a.ks[i] = k_values_host[i] < kmax ? k_values_host[i] : kmax;
    }
    a.row_vals = reinterpret_cast<double *>(w); w += met_align((size_t)nk * 3 * n_eval * 8);
    a.row_valid = reinterpret_cast<int *>(w); w += met_align((size_t)n_eval * 4);
    a.item_counts = reinterpret_cast<int *>(w); w += met_align((size_t)nk * (n_items + 1) * 4);
    uint64_t *keys_a = reinterpret_cast<uint64_t *>(w); w += met_align((size_t)n_items * 8);
    uint64_t *keys_b = reinterpret_cast<uint64_t *>(w); w += met_align((size_t)n_items * 8);
    void *sort_ws = w;
    const size_t sort_bytes = met_align(radix_sort_workspace_bytes(n_items));

    GR_CUDA_CHECK(cudaMemsetAsync(a.item_counts, 0, (size_t)nk * (n_items + 1) * 4, s));
    GR_CUDA_CHECK(cudaMemsetAsync(out_counts, 0, (size_t)nk * 3 * 8, s));
    metrics_rows_kernel<<<(unsigned)((n_eval + 7) / 8), 256, 0, s>>>(a);
    GR_LAUNCH_CHECK();
    metrics_reduce_kernel<<<nk * 3 + 1, 1024, 0, s>>>(a.row_vals, a.row_valid, n_eval, nk * 3, out_sums);
    GR_LAUNCH_CHECK();
    const int bits = met_bits((uint64_t)n_eval * (uint64_t)(kmax + 1));   // one per row, plus the folded padding ids
    for (int ki = 0; ki < nk; ++ki) {
        unsigned long long *cnt = reinterpret_cast<unsigned long long *>(out_counts) + ki * 3;
        metrics_count_keys_kernel<<<(unsigned)((n_items + 1 + 255) / 256), 256, 0, s>>>(
            a.item_counts + (long long)ki * (n_items + 1), n_items, keys_a, cnt);
        GR_LAUNCH_CHECK();
        bool in_a = true;
        const int rc = radix_sort_u64(keys_a, keys_b, nullptr, nullptr, n_items, 0, bits, sort_ws, sort_bytes, &in_a, s);
        if (rc != GR_OK) return rc;
        metrics_gini_sum_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, s>>>(in_a ? keys_a : keys_b, n_items, cnt);
        GR_LAUNCH_CHECK();
    }
    return GR_OK;
}
