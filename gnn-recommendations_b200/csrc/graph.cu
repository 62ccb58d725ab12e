// Graph builder: COO -> CSR, (user,item) pairs -> symmetric-normalised CSR, row schedule.
// Integer work, HBM-bound; everything is bit-exact with scipy's canonical CSR
// (src/data/graph_builder.py:16-144 of the reference).
#include <initializer_list>

#include "gr_common.cuh"

namespace gr {

// =============================================================================================
// exclusive scan of uint32 (hierarchical, in place)
// =============================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    // v = this thread's value; returns the exclusive prefix within the block
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_base = warp ? warp_sums[warp - 1] : 0;
    if (total) *total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return warp_base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *data, long long n, uint32_t *sums) {
    const long long base = (long long)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const long long idx = base + (long long)i * kScanThreads + threadIdx.x;
        if (idx < n) s += data[idx];
    }
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *data, long long n, const uint32_t *sums) {
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        s += v[i];
    }
    uint32_t ex = block_exclusive_scan(s, nullptr) + (sums ? sums[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) data[base + i] = ex;
        ex += v[i];
    }
}

static size_t scan_workspace_bytes(long long n) {
    size_t total = 0;
    while (n > kScanTile) {
        n = (n + kScanTile - 1) / kScanTile;
        total += ((size_t)n * 4 + 255) & ~(size_t)255;
    }
    return total + 256;
}

// in-place exclusive scan; ws must hold scan_workspace_bytes(n)
static int exclusive_scan_u32(uint32_t *data, long long n, void *ws, cudaStream_t stream) {
    if (n <= 0) return GR_OK;
    const long long nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 1) {
        scan_apply_kernel<<<1, kScanThreads, 0, stream>>>(data, n, nullptr);
        GR_LAUNCH_CHECK();
        return GR_OK;
    }
    uint32_t *sums = static_cast<uint32_t *>(ws);
    void *next_ws = static_cast<char *>(ws) + (((size_t)nb * 4 + 255) & ~(size_t)255);
    scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(data, n, sums);
    GR_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(sums, nb, next_ws, stream);
    if (rc != GR_OK) return rc;
    scan_apply_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(data, n, sums);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

// =============================================================================================
// LSD radix sort, 8 bits per pass, stable
// =============================================================================================
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per CTA
constexpr int kRsBins = 256;

__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint64_t *keys, long long n, int shift,
                                                             uint32_t *block_hist, int nblocks) {
    __shared__ uint32_t hist[kRsBins];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kRsTile;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const long long idx = base + (long long)i * kRsThreads + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    block_hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const uint64_t *keys_in, const uint32_t *pay_in,
                                                                uint64_t *keys_out, uint32_t *pay_out, long long n,
                                                                int shift, const uint32_t *offsets, int nblocks) {
    __shared__ uint32_t warp_cnt[kRsWarps][kRsBins];
    __shared__ uint32_t bin_base[kRsBins];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kRsWarps * kRsBins; i += kRsThreads) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    // warp w owns the contiguous segment [seg, seg + 512); round i covers 32 consecutive keys,
    // so (warp, round, lane) order == key index order and the ranks below are stable.
    const long long seg = (long long)blockIdx.x * kRsTile + (long long)warp * (kRsItems * 32);
    uint64_t key[kRsItems];
    uint16_t rank[kRsItems];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const long long idx = seg + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? keys_in[idx] : 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        rank[i] = 0;
        if (valid) {
            const int d = (int)((key[i] >> shift) & 0xff);
            const unsigned peers = __match_any_sync(vmask, d);
            const uint32_t old = warp_cnt[warp][d];
            __syncwarp(vmask);
            if ((peers & lt_mask) == 0) warp_cnt[warp][d] = old + __popc(peers);
            __syncwarp(vmask);
            rank[i] = (uint16_t)(old + __popc(peers & lt_mask));
        }
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // one digit per thread
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
        bin_base[d] = offsets[(size_t)d * nblocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const long long idx = seg + i * 32 + lane;
        if (idx < n) {
            const int d = (int)((key[i] >> shift) & 0xff);
            const size_t pos = (size_t)bin_base[d] + warp_cnt[warp][d] + rank[i];
            keys_out[pos] = key[i];
            if (pay_in) pay_out[pos] = pay_in[idx];
        }
    }
}

size_t radix_sort_workspace_bytes(int64_t n) {
    const long long nblocks = (n + kRsTile - 1) / kRsTile;
    const size_t hist = (((size_t)nblocks * kRsBins * 4) + 255) & ~(size_t)255;
    return hist + scan_workspace_bytes(nblocks * kRsBins) + 256;
}

int radix_sort_u64(uint64_t *keys_a, uint64_t *keys_b, uint32_t *pay_a, uint32_t *pay_b, int64_t n, int begin_bit,
                   int end_bit, void *ws, size_t ws_bytes, bool *in_a_out, cudaStream_t stream) {
    *in_a_out = true;
    if (n <= 1 || end_bit <= begin_bit) return GR_OK;
    if (n >= (1LL << 32)) return GR_ERR_OVERFLOW;
    if (ws_bytes < radix_sort_workspace_bytes(n)) return GR_ERR_WORKSPACE;
    const long long nblocks = (n + kRsTile - 1) / kRsTile;
    uint32_t *hist = static_cast<uint32_t *>(ws);
    void *scan_ws = static_cast<char *>(ws) + ((((size_t)nblocks * kRsBins * 4) + 255) & ~(size_t)255);
    bool in_a = true;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        uint64_t *kin = in_a ? keys_a : keys_b, *kout = in_a ? keys_b : keys_a;
        uint32_t *pin = pay_a ? (in_a ? pay_a : pay_b) : nullptr, *pout = pay_a ? (in_a ? pay_b : pay_a) : nullptr;
        rs_hist_kernel<<<(unsigned)nblocks, kRsThreads, 0, stream>>>(kin, n, shift, hist, (int)nblocks);
        GR_LAUNCH_CHECK();
        int rc = exclusive_scan_u32(hist, nblocks * kRsBins, scan_ws, stream);
        if (rc != GR_OK) return rc;
        rs_scatter_kernel<<<(unsigned)nblocks, kRsThreads, 0, stream>>>(kin, pin, kout, pout, n, shift, hist,
                                                                         (int)nblocks);
        GR_LAUNCH_CHECK();
        in_a = !in_a;
    }
    *in_a_out = in_a;
    return GR_OK;
}

static inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// =============================================================================================
// COO (row-major sorted) -> CSR
// =============================================================================================
__global__ void coo_sorted_to_csr_kernel(const int64_t *rows, const int64_t *cols, const float *vals, long long nnz,
                                         long long n_rows, long long n_cols, int *indptr, int *indices,
                                         float *out_vals, int *status) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const long long r = rows[k], c = cols[k];
        if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
            atomicOr(status, 2);
            continue;
        }
        indices[k] = (int)c;
        if (out_vals != vals) out_vals[k] = vals[k];
        long long prev = k ? rows[k - 1] : -1;
        prev = prev < -1 ? -1 : (prev >= n_rows ? n_rows - 1 : prev);
        if (prev > r) atomicOr(status, 1);
        for (long long rr = prev + 1; rr <= r; ++rr) indptr[rr] = (int)k;
        if (k == nnz - 1)
            for (long long rr = r + 1; rr <= n_rows; ++rr) indptr[rr] = (int)nnz;
    }
}

// =============================================================================================
// row schedule
// =============================================================================================
__global__ void schedule_keys_kernel(const int *indptr, long long n_rows, int row_bits, uint64_t *keys) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) {
        const uint32_t len = (uint32_t)(indptr[r + 1] - indptr[r]);
        keys[r] = ((uint64_t)(0x7fffffffu - len) << row_bits) | (uint64_t)r;
    }
}
__global__ void schedule_emit_kernel(const uint64_t *keys, long long n_rows, int row_bits, int threshold,
                                     int *row_order, int *n_long) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) {
        const uint64_t k = keys[i];
        row_order[i] = (int)(k & ((1ull << row_bits) - 1ull));
        const int len = (int)(0x7fffffffu - (uint32_t)(k >> row_bits));
        const int nxt = (i + 1 < n_rows) ? (int)(0x7fffffffu - (uint32_t)(keys[i + 1] >> row_bits)) : -1;
        if (len >= threshold && nxt < threshold) *n_long = (int)(i + 1);
    }
}

// group_ptr[g] = first row whose entries start at or after g * group_nnz (row groups of about
// group_nnz entries for the streaming SpMM kernel); group_ptr[n_groups] = n_rows.
__global__ void row_groups_kernel(const int *indptr, int n_rows, int group_nnz, int n_groups, int *group_ptr) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    if (g == n_groups) {
        group_ptr[g] = n_rows;
        return;
    }
    const long long target = g * (long long)group_nnz;
    int lo = 0, hi = n_rows;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)indptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    group_ptr[g] = lo;
}

// =============================================================================================
// (user,item) pairs -> CSR pattern with multiplicities and degrees
// =============================================================================================
__global__ void pair_keys_kernel(const int64_t *user, const int64_t *item, long long n_pairs, long long n_users,
                                 long long n_items, int self_loop, int col_bits, uint64_t *keys, int *deg,
                                 int *status) {
    const long long n = n_users + n_items;
    const long long total = 2 * n_pairs + (self_loop ? n : 0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        long long r, c;
        if (k < 2 * n_pairs) {
            const long long p = k < n_pairs ? k : k - n_pairs;
            long long u = user[p], i = item[p];
            if (u < 0 || u >= n_users || i < 0 || i >= n_items) {
                atomicOr(status, 2);
                u = 0;
                i = 0;
            }
            r = k < n_pairs ? u : n_users + i;
            c = k < n_pairs ? n_users + i : u;
        } else {
            r = c = k - 2 * n_pairs;
        }
        keys[k] = ((uint64_t)r << col_bits) | (uint64_t)c;
        atomicAdd(&deg[r], 1);
    }
}

// head flags of the sorted keys -> flags[k] = 1 when key k starts a new (row,col)
__global__ void unique_flags_kernel(const uint64_t *keys, long long n, uint32_t *flags) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

// pos = exclusive scan of flags.  Emits indices / multiplicities / indptr of the unique entries.
__global__ void unique_emit_kernel(const uint64_t *keys, const uint32_t *pos, long long n, int col_bits,
                                   long long n_rows, int *indptr, int *indices, float *mult, long long *nnz_out) {
    const uint64_t col_mask = (1ull << col_bits) - 1ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint64_t key = keys[k];
        const bool head = (k == 0) || key != keys[k - 1];
        if (!head) continue;
        const uint32_t p = pos[k];
        // multiplicity = distance to the next head
        long long e = k + 1;
        while (e < n && keys[e] == key) ++e;
        indices[p] = (int)(key & col_mask);
        mult[p] = (float)(e - k);
        const long long r = (long long)(key >> col_bits);
        const long long prev = k ? (long long)(keys[k - 1] >> col_bits) : -1;
        for (long long rr = prev + 1; rr <= r; ++rr) indptr[rr] = (int)p;
        if (e == n) {
            const uint32_t total = p + 1;
            for (long long rr = r + 1; rr <= n_rows; ++rr) indptr[rr] = (int)total;
            *nnz_out = (long long)total;
        }
    }
}

__global__ void max_deg_kernel(const int *deg, long long n, int *max_deg) {
    int m = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) m = max(m, deg[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_deg, m);
}

// vals[k] = fl(fl(lut[deg[r]] * mult[k]) * lut[deg[c]])   (symmetric; graph_builder.py:123-126)
//         = fl(lut[deg[r]] * mult[k])                      (row;       graph_builder.py:128-134)
__global__ void normalize_kernel(const int *indptr, const int *indices, const float *mult, const int *deg,
                                 const float *lut, long long lut_len, long long n_rows, long long nnz, int mode,
                                 float *vals, int *status) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        // row of entry k: last r with indptr[r] <= k
        long long lo = 0, hi = n_rows;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (indptr[mid] <= k) lo = mid; else hi = mid;
        }
        const int c = indices[k];
        const int dr = deg[lo], dc = deg[c];
        if (dr >= lut_len || dc >= lut_len) {
            atomicOr(status, 4);
            continue;
        }
        const float a = mult[k];
        float v;
        if (mode == 0) v = __fmul_rn(__fmul_rn(lut[dr], a), lut[dc]);
        else if (mode == 1) v = __fmul_rn(lut[dr], a);
        else v = a;
        vals[k] = v;
    }
}

static inline unsigned grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace gr

using namespace gr;

extern "C" int gr_coo_sorted_to_csr(const int64_t *rows, const int64_t *cols, const float *vals, int64_t nnz,
                                    int64_t n_rows, int64_t n_cols, int32_t *indptr, int32_t *indices,
                                    float *out_vals, int32_t *status, void *stream) {
    if (nnz < 0 || n_rows < 0 || n_cols < 0 || !indptr || !status) return GR_ERR_INVALID;
    if (nnz > 0x7fffffffLL || n_rows >= 0x7fffffffLL || n_cols > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nnz == 0) {
        GR_CUDA_CHECK(cudaMemsetAsync(indptr, 0, (size_t)(n_rows + 1) * 4, s));
        return GR_OK;
    }
    if (!rows || !cols || !vals || !indices || !out_vals) return GR_ERR_INVALID;
    coo_sorted_to_csr_kernel<<<grid_for(nnz, 256), 256, 0, s>>>(rows, cols, vals, nnz, n_rows, n_cols, indptr,
                                                                indices, out_vals, status);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" size_t gr_row_schedule_workspace_bytes(int64_t n_rows) {
    if (n_rows < 0) return 0;
    return 2 * align256((size_t)n_rows * 8) + radix_sort_workspace_bytes(n_rows) + 256;
}

extern "C" int gr_row_schedule(const int32_t *indptr, int64_t n_rows, int32_t long_threshold, int32_t *row_order,
                               int32_t *n_long_out, void *workspace, size_t workspace_bytes, void *stream) {
    if (!indptr || !row_order || !n_long_out || n_rows < 0 || long_threshold < 1) return GR_ERR_INVALID;
    if (n_rows >= 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if (workspace_bytes < gr_row_schedule_workspace_bytes(n_rows) || !workspace) return GR_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GR_CUDA_CHECK(cudaMemsetAsync(n_long_out, 0, 4, s));
    if (n_rows == 0) return GR_OK;
    char *w = static_cast<char *>(workspace);
    uint64_t *ka = reinterpret_cast<uint64_t *>(w);
    uint64_t *kb = reinterpret_cast<uint64_t *>(w + align256((size_t)n_rows * 8));
    void *sort_ws = w + 2 * align256((size_t)n_rows * 8);
    const unsigned g = (unsigned)((n_rows + 255) / 256);
    const int row_bits = bits_for((uint64_t)(n_rows - 1));
    schedule_keys_kernel<<<g, 256, 0, s>>>(indptr, n_rows, row_bits, ka);
    GR_LAUNCH_CHECK();
    bool in_a = true;
    // low bits: row id; above them: 0x7fffffff - len (31 bits)
    int rc = radix_sort_u64(ka, kb, nullptr, nullptr, n_rows, 0, row_bits + 31, sort_ws,
                            workspace_bytes - 2 * align256((size_t)n_rows * 8), &in_a, s);
    if (rc != GR_OK) return rc;
    schedule_emit_kernel<<<g, 256, 0, s>>>(in_a ? ka : kb, n_rows, row_bits, long_threshold, row_order, n_long_out);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" int gr_row_groups(const int32_t *indptr, int64_t n_rows, int32_t group_nnz, int32_t n_groups,
                             int32_t *group_ptr, void *stream) {
    if (!indptr || !group_ptr || n_rows < 0 || group_nnz < 1 || n_groups < 1) return GR_ERR_INVALID;
    if (n_rows >= 0x7fffffffLL) return GR_ERR_OVERFLOW;
    row_groups_kernel<<<(unsigned)((n_groups + 1 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        indptr, (int)n_rows, group_nnz, n_groups, group_ptr);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" size_t gr_build_csr_workspace_bytes(int64_t n_pairs, int64_t n_users, int64_t n_items, int32_t self_loop) {
    if (n_pairs < 0 || n_users < 0 || n_items < 0) return 0;
    const int64_t total = 2 * n_pairs + (self_loop ? n_users + n_items : 0);
    return 2 * align256((size_t)total * 8) + align256((size_t)total * 4) + radix_sort_workspace_bytes(total) +
           scan_workspace_bytes(total) + 512;
}

extern "C" int gr_build_csr_pattern(const int64_t *user, const int64_t *item, int64_t n_pairs, int64_t n_users,
                                    int64_t n_items, int32_t self_loop, int32_t *indptr, int32_t *indices,
                                    float *mult, int32_t *deg, int64_t *nnz_out, int32_t *max_deg_out,
                                    int32_t *status, void *workspace, size_t workspace_bytes, void *stream) {
    if (n_pairs < 0 || n_users < 0 || n_items < 0 || !indptr || !deg || !nnz_out || !max_deg_out || !status)
        return GR_ERR_INVALID;
    const int64_t n = n_users + n_items;
    const int64_t total = 2 * n_pairs + (self_loop ? n : 0);
    if (n >= 0x7fffffffLL || total > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if (workspace_bytes < gr_build_csr_workspace_bytes(n_pairs, n_users, n_items, self_loop) || !workspace)
        return GR_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GR_CUDA_CHECK(cudaMemsetAsync(deg, 0, (size_t)n * 4, s));
    GR_CUDA_CHECK(cudaMemsetAsync(nnz_out, 0, 8, s));
    GR_CUDA_CHECK(cudaMemsetAsync(max_deg_out, 0, 4, s));
    if (total == 0) {
        GR_CUDA_CHECK(cudaMemsetAsync(indptr, 0, (size_t)(n + 1) * 4, s));
        return GR_OK;
    }
    if (!user || !item || !indices || !mult) return GR_ERR_INVALID;
    char *w = static_cast<char *>(workspace);
    uint64_t *ka = reinterpret_cast<uint64_t *>(w);
    w += align256((size_t)total * 8);
    uint64_t *kb = reinterpret_cast<uint64_t *>(w);
    w += align256((size_t)total * 8);
    uint32_t *flags = reinterpret_cast<uint32_t *>(w);
    w += align256((size_t)total * 4);
    void *sort_ws = w;
    const size_t sort_bytes = radix_sort_workspace_bytes(total);
    void *scan_ws = w + sort_bytes;

    const int col_bits = bits_for((uint64_t)(n > 0 ? n - 1 : 0));
    pair_keys_kernel<<<grid_for(total, 256), 256, 0, s>>>(user, item, n_pairs, n_users, n_items, self_loop, col_bits,
                                                          ka, deg, status);
    GR_LAUNCH_CHECK();
    bool in_a = true;
    int rc = radix_sort_u64(ka, kb, nullptr, nullptr, total, 0, 2 * col_bits, sort_ws, sort_bytes, &in_a, s);
    if (rc != GR_OK) return rc;
    const uint64_t *sorted = in_a ? ka : kb;
    unique_flags_kernel<<<grid_for(total, 256), 256, 0, s>>>(sorted, total, flags);
    GR_LAUNCH_CHECK();
    rc = exclusive_scan_u32(flags, total, scan_ws, s);
    if (rc != GR_OK) return rc;
    unique_emit_kernel<<<grid_for(total, 256), 256, 0, s>>>(sorted, flags, total, col_bits, n, indptr, indices, mult,
                                                            reinterpret_cast<long long *>(nnz_out));
    GR_LAUNCH_CHECK();
    max_deg_kernel<<<grid_for(n, 256), 256, 0, s>>>(deg, n, max_deg_out);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" int gr_csr_normalize(const int32_t *indptr, const int32_t *indices, const float *mult, const int32_t *deg,
                                const float *lut, int64_t lut_len, int64_t n_rows, int64_t nnz, int32_t mode,
                                float *vals, int32_t *status, void *stream) {
    if (nnz < 0 || n_rows < 0 || mode < 0 || mode > 2 || !status) return GR_ERR_INVALID;
    if (nnz == 0) return GR_OK;
    if (!indptr || !indices || !mult || !deg || !vals || (mode != 2 && (!lut || lut_len <= 0))) return GR_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    normalize_kernel<<<grid_for(nnz, 256), 256, 0, s>>>(indptr, indices, mult, deg, lut, lut_len, n_rows, nnz, mode,
                                                        vals, status);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

// =============================================================================================
// Per-user temporal split (src/data/dataset.py:327-357)
// =============================================================================================
namespace gr {

// key = user << ts_bits | (timestamp - ts_min); payload = original row.  status |= 1 on a range violation.
__global__ void split_keys_kernel(const int64_t *user, const int64_t *ts, long long n, long long n_users, long long ts_min,
                                  long long ts_max, int ts_bits, uint64_t *keys, uint32_t *pay, int *status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long u = user[i], t = ts[i];
    if (u < 0 || u >= n_users || t < ts_min || t > ts_max) {
        atomicOr(status, 1);
        keys[i] = ~0ULL;
        pay[i] = (uint32_t)i;
        return;
    }
    keys[i] = ((uint64_t)u << ts_bits) | (uint64_t)(t - ts_min);
    pay[i] = (uint32_t)i;
}

// label of sorted position j: 2 = test (user's last row, needs >= 2 rows), 1 = valid (second-last row,
// needs >= 3 rows), 0 = train (dataset.py:340-352).  One flag array per label for the scans.
__global__ void split_label_kernel(const uint64_t *keys, long long n, int ts_bits, uint32_t *f_train, uint32_t *f_valid,
                                   uint32_t *f_test) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t u = keys[j] >> ts_bits;
    const bool has_prev = j > 0 && (keys[j - 1] >> ts_bits) == u;
    const bool next_same = j + 1 < n && (keys[j + 1] >> ts_bits) == u;
    const bool next2_same = j + 2 < n && (keys[j + 2] >> ts_bits) == u;
    const bool is_last = !next_same;
    const bool is_second_last = next_same && !next2_same;
    const bool test = is_last && has_prev;
    const bool valid = is_second_last && has_prev;
    f_test[j] = test;
    f_valid[j] = valid;
    f_train[j] = !(test || valid);
}

__global__ void split_emit_kernel(const uint64_t *keys, const uint32_t *pay, const int64_t *item, long long n, int ts_bits,
                                  const uint32_t *p_train, const uint32_t *p_valid, const uint32_t *p_test,
                                  int64_t *train_u, int64_t *train_i, int64_t *valid_u, int64_t *valid_i,
                                  int64_t *test_u, int64_t *test_i, int64_t *counts) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t u = keys[j] >> ts_bits;
    const bool has_prev = j > 0 && (keys[j - 1] >> ts_bits) == u;
    const bool next_same = j + 1 < n && (keys[j + 1] >> ts_bits) == u;
    const bool next2_same = j + 2 < n && (keys[j + 2] >> ts_bits) == u;
    const bool test = !next_same && has_prev;
    const bool valid = next_same && !next2_same && has_prev;
    const int64_t it = item[pay[j]];
    if (test) {
        test_u[p_test[j]] = (int64_t)u;
        test_i[p_test[j]] = it;
    } else if (valid) {
        valid_u[p_valid[j]] = (int64_t)u;
        valid_i[p_valid[j]] = it;
    } else {
        train_u[p_train[j]] = (int64_t)u;
        train_i[p_train[j]] = it;
    }
    if (j == n - 1) {   // exclusive prefix + own flag = totals
        counts[0] = (int64_t)p_train[j] + (!(test || valid));
        counts[1] = (int64_t)p_valid[j] + valid;
        counts[2] = (int64_t)p_test[j] + test;
    }
}

}  // namespace gr

extern "C" size_t gr_temporal_split_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return 2 * align256((size_t)n * 8) + 2 * align256((size_t)n * 4) + 3 * align256((size_t)n * 4) +
           align256(radix_sort_workspace_bytes(n)) + align256(scan_workspace_bytes(n)) + 512;
}

extern "C" int gr_temporal_split(const int64_t *user, const int64_t *item, const int64_t *timestamp, int64_t n,
                                 int64_t n_users, int64_t ts_min, int64_t ts_max, int64_t *train_u, int64_t *train_i,
                                 int64_t *valid_u, int64_t *valid_i, int64_t *test_u, int64_t *test_i, int64_t *counts,
                                 int32_t *status, void *workspace, size_t workspace_bytes, void *stream) {
    if (!user || !item || !timestamp || !train_u || !train_i || !valid_u || !valid_i || !test_u || !test_i || !counts ||
        !status || !workspace)
        return GR_ERR_INVALID;
    if (n < 0 || n_users <= 0 || ts_max < ts_min) return GR_ERR_INVALID;
    if (n >= (1LL << 32)) return GR_ERR_OVERFLOW;
    if (workspace_bytes < gr_temporal_split_workspace_bytes(n)) return GR_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GR_CUDA_CHECK(cudaMemsetAsync(status, 0, 4, s));
    GR_CUDA_CHECK(cudaMemsetAsync(counts, 0, 3 * 8, s));
    if (n == 0) return GR_OK;
    const int ts_bits = bits_for((uint64_t)(ts_max - ts_min));
    const int user_bits = bits_for((uint64_t)(n_users - 1));
    if (ts_bits + user_bits > 63) return GR_ERR_OVERFLOW;     // ~0 stays reserved for rejected rows
    char *w = static_cast<char *>(workspace);
    uint64_t *ka = reinterpret_cast<uint64_t *>(w); w += align256((size_t)n * 8);
    uint64_t *kb = reinterpret_cast<uint64_t *>(w); w += align256((size_t)n * 8);
    uint32_t *pa = reinterpret_cast<uint32_t *>(w); w += align256((size_t)n * 4);
    uint32_t *pb = reinterpret_cast<uint32_t *>(w); w += align256((size_t)n * 4);
    uint32_t *f0 = reinterpret_cast<uint32_t *>(w); w += align256((size_t)n * 4);
    uint32_t *f1 = reinterpret_cast<uint32_t *>(w); w += align256((size_t)n * 4);
    uint32_t *f2 = reinterpret_cast<uint32_t *>(w); w += align256((size_t)n * 4);
    void *sort_ws = w; w += align256(radix_sort_workspace_bytes(n));
    void *scan_ws = w;
    const unsigned grid = (unsigned)((n + 255) / 256);
    split_keys_kernel<<<grid, 256, 0, s>>>(user, timestamp, n, n_users, ts_min, ts_max, ts_bits, ka, pa, status);
    GR_LAUNCH_CHECK();
    bool in_a = true;
    int rc = radix_sort_u64(ka, kb, pa, pb, n, 0, ts_bits + user_bits, sort_ws, align256(radix_sort_workspace_bytes(n)),
                            &in_a, s);
    if (rc != GR_OK) return rc;
    const uint64_t *keys = in_a ? ka : kb;
    const uint32_t *pay = in_a ? pa : pb;
    split_label_kernel<<<grid, 256, 0, s>>>(keys, n, ts_bits, f0, f1, f2);
    GR_LAUNCH_CHECK();
    for (uint32_t *f : {f0, f1, f2}) {
        rc = exclusive_scan_u32(f, n, scan_ws, s);
        if (rc != GR_OK) return rc;
    }
    split_emit_kernel<<<grid, 256, 0, s>>>(keys, pay, item, n, ts_bits, f0, f1, f2, train_u, train_i, valid_u, valid_i,
                                           test_u, test_i, counts);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

// =============================================================================================
// Partitioned graph build: the rows rank, rank + G, ... of Â straight from the pairs
// =============================================================================================
namespace gr {

// Every rank walks all pairs (global degrees need them all) but keeps only the directed entries whose
// row it owns.  key = local row (r / G) << col_bits | global column; appended through a counter — the
// order is irrelevant, the sort that follows fixes it.
__global__ void pair_keys_local_kernel(const int64_t *user, const int64_t *item, long long n_pairs, long long n_users,
                                       long long n_items, int self_loop, int col_bits, int world, int rank,
                                       uint64_t *keys, unsigned long long *counter, long long cap, int *deg,
                                       int *status) {
    const long long n = n_users + n_items;
    const long long total = 2 * n_pairs + (self_loop ? n : 0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        long long r, c;
        if (k < 2 * n_pairs) {
            const long long p = k < n_pairs ? k : k - n_pairs;
            long long u = user[p], i = item[p];
            if (u < 0 || u >= n_users || i < 0 || i >= n_items) {
                atomicOr(status, 2);
                u = 0;
                i = 0;
            }
            r = k < n_pairs ? u : n_users + i;
            c = k < n_pairs ? n_users + i : u;
        } else {
            r = c = k - 2 * n_pairs;
        }
        atomicAdd(&deg[r], 1);
        if (r % world == rank) {
            const unsigned long long pos = atomicAdd(counter, 1ULL);
            if ((long long)pos < cap) keys[pos] = ((uint64_t)(r / world) << col_bits) | (uint64_t)c;
            else atomicOr(status, 8);
        }
    }
}

// as unique_emit_kernel, but rows are local and the column ids leave in the owner-major padded numbering
// of the exchange buffers ((c % G) * H + c / G); the ORDER inside a row stays ascending global column,
// i.e. the single-GPU chain order.
__global__ void unique_emit_local_kernel(const uint64_t *keys, const uint32_t *pos, long long n, int col_bits,
                                         long long n_rows, int world, long long block_rows, int *indptr, int *indices,
                                         float *mult, long long *nnz_out) {
    const uint64_t col_mask = (1ull << col_bits) - 1ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint64_t key = keys[k];
        const bool head = (k == 0) || key != keys[k - 1];
        if (!head) continue;
        const uint32_t p = pos[k];
        long long e = k + 1;
        while (e < n && keys[e] == key) ++e;
        const long long c = (long long)(key & col_mask);
        indices[p] = (int)((c % world) * block_rows + c / world);
        mult[p] = (float)(e - k);
        const long long r = (long long)(key >> col_bits);
        const long long prev = k ? (long long)(keys[k - 1] >> col_bits) : -1;
        for (long long rr = prev + 1; rr <= r; ++rr) indptr[rr] = (int)p;
        if (e == n) {
            const uint32_t total = p + 1;
            for (long long rr = r + 1; rr <= n_rows; ++rr) indptr[rr] = (int)total;
            *nnz_out = (long long)total;
        }
    }
}

__global__ void normalize_local_kernel(const int *indptr, const int *indices, const float *mult, const int *deg,
                                       const float *lut, long long lut_len, long long n_rows, long long nnz, int mode,
                                       int world, int rank, long long block_rows, float *vals, int *status) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        long long lo = 0, hi = n_rows;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (indptr[mid] <= k) lo = mid; else hi = mid;
        }
        const long long pc = indices[k];
        const long long gr_ = lo * world + rank;                               // global row
        const long long gc = (pc % block_rows) * world + pc / block_rows;      // global column
        const int dr = deg[gr_], dc = deg[gc];
        if (dr >= lut_len || dc >= lut_len) {
            atomicOr(status, 4);
            continue;
        }
        const float a = mult[k];
        float v;
        if (mode == 0) v = __fmul_rn(__fmul_rn(lut[dr], a), lut[dc]);
        else if (mode == 1) v = __fmul_rn(lut[dr], a);
        else v = a;
        vals[k] = v;
    }
}

}  // namespace gr

extern "C" size_t gr_build_local_csr_workspace_bytes(int64_t n_local_entries) {
    if (n_local_entries < 0) return 0;
    const int64_t t = n_local_entries > 0 ? n_local_entries : 1;
    return 2 * align256((size_t)t * 8) + align256((size_t)t * 4) + align256(radix_sort_workspace_bytes(t)) +
           align256(scan_workspace_bytes(t)) + 512;
}

extern "C" int gr_build_local_csr_pattern(const int64_t *user, const int64_t *item, int64_t n_pairs, int64_t n_users,
                                          int64_t n_items, int32_t self_loop, int32_t world, int32_t rank,
                                          int64_t block_rows, int64_t n_local_entries, int32_t *indptr,
                                          int32_t *indices, float *mult, int32_t *deg, int64_t *nnz_out,
                                          int32_t *max_deg_out, int32_t *status, void *workspace,
                                          size_t workspace_bytes, void *stream) {
    if (n_pairs < 0 || n_users < 0 || n_items < 0 || world <= 0 || rank < 0 || rank >= world || block_rows <= 0 ||
        n_local_entries < 0 || !indptr || !deg || !nnz_out || !max_deg_out || !status || !workspace)
        return GR_ERR_INVALID;
    const int64_t n = n_users + n_items;
    const int64_t total = 2 * n_pairs + (self_loop ? n : 0);
    if (n >= 0x7fffffffLL || n_local_entries > 0x7fffffffLL || block_rows * world >= 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if (workspace_bytes < gr_build_local_csr_workspace_bytes(n_local_entries)) return GR_ERR_WORKSPACE;
    const int64_t n_local = n > rank ? (n - rank + world - 1) / world : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GR_CUDA_CHECK(cudaMemsetAsync(deg, 0, (size_t)n * 4, s));
    GR_CUDA_CHECK(cudaMemsetAsync(nnz_out, 0, 8, s));
    GR_CUDA_CHECK(cudaMemsetAsync(max_deg_out, 0, 4, s));
    GR_CUDA_CHECK(cudaMemsetAsync(indptr, 0, (size_t)(n_local + 1) * 4, s));
    if (total == 0) return GR_OK;
    if (!user || !item || (n_local_entries > 0 && (!indices || !mult))) return GR_ERR_INVALID;
    const int64_t t = n_local_entries > 0 ? n_local_entries : 1;
    char *w = static_cast<char *>(workspace);
    uint64_t *ka = reinterpret_cast<uint64_t *>(w); w += align256((size_t)t * 8);
    uint64_t *kb = reinterpret_cast<uint64_t *>(w); w += align256((size_t)t * 8);
    uint32_t *flags = reinterpret_cast<uint32_t *>(w); w += align256((size_t)t * 4);
    void *sort_ws = w; w += align256(radix_sort_workspace_bytes(t));
    void *scan_ws = w; w += align256(scan_workspace_bytes(t));
    unsigned long long *counter = reinterpret_cast<unsigned long long *>(w);
    GR_CUDA_CHECK(cudaMemsetAsync(counter, 0, 8, s));

    const int col_bits = bits_for((uint64_t)(n > 0 ? n - 1 : 0));
    const int row_bits = bits_for((uint64_t)(n_local > 0 ? n_local - 1 : 0));
    pair_keys_local_kernel<<<grid_for(total, 256), 256, 0, s>>>(user, item, n_pairs, n_users, n_items, self_loop,
                                                                col_bits, world, rank, ka, counter, n_local_entries, deg,
                                                                status);
    GR_LAUNCH_CHECK();
    max_deg_kernel<<<grid_for(n, 256), 256, 0, s>>>(deg, n, max_deg_out);
    GR_LAUNCH_CHECK();
    if (n_local_entries == 0) return GR_OK;
    bool in_a = true;
    int rc = radix_sort_u64(ka, kb, nullptr, nullptr, n_local_entries, 0, col_bits + row_bits, sort_ws,
                            align256(radix_sort_workspace_bytes(t)), &in_a, s);
    if (rc != GR_OK) return rc;
    const uint64_t *sorted = in_a ? ka : kb;
    unique_flags_kernel<<<grid_for(n_local_entries, 256), 256, 0, s>>>(sorted, n_local_entries, flags);
    GR_LAUNCH_CHECK();
    rc = exclusive_scan_u32(flags, n_local_entries, scan_ws, s);
    if (rc != GR_OK) return rc;
    unique_emit_local_kernel<<<grid_for(n_local_entries, 256), 256, 0, s>>>(
        sorted, flags, n_local_entries, col_bits, n_local, world, block_rows, indptr, indices, mult,
        reinterpret_cast<long long *>(nnz_out));
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" int gr_csr_normalize_local(const int32_t *indptr, const int32_t *indices, const float *mult,
                                      const int32_t *deg, const float *lut, int64_t lut_len, int64_t n_local_rows,
                                      int64_t nnz, int32_t mode, int32_t world, int32_t rank, int64_t block_rows,
                                      float *vals, int32_t *status, void *stream) {
    if (nnz < 0 || n_local_rows < 0 || mode < 0 || mode > 2 || world <= 0 || rank < 0 || rank >= world ||
        block_rows <= 0 || !status)
        return GR_ERR_INVALID;
    if (nnz == 0) return GR_OK;
    if (!indptr || !indices || !mult || !deg || !vals || (mode != 2 && (!lut || lut_len <= 0))) return GR_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    normalize_local_kernel<<<grid_for(nnz, 256), 256, 0, s>>>(indptr, indices, mult, deg, lut, lut_len, n_local_rows, nnz,
                                                              mode, world, rank, block_rows, vals, status);
    GR_LAUNCH_CHECK();
    return GR_OK;
}
