// Per-row dense epilogues of the NGCF / GAT / Group-and-Shuffle layers, and the GAT edge-softmax
// aggregation over the CSR pattern.
//
//   gr_rowmap_f32       out = alpha * act( X1 Wa + ba  +  (X2 * X3) Wb + bb ) + beta * R
//                       NGCF layer    : X1 = n = A x, X2 = x, X3 = n, Wa = W1^T, Wb = W2^T, LeakyReLU(0.2)
//                                       (src/models/baselines/ngcf.py:77-84)
//                       Group&Shuffle : X1 = A x, Wa = W_conn W_orth[:,perm], alpha = 1-a, beta = a, R = x0
//                                       (src/models/orthogonal_bundle/model.py:176-195)
//                       GAT           : X1 = x, Wa = [W_0^T | ... | W_{H-1}^T]  (gat.py:99)
//   gr_gat_node_scores  s[i,h] = <H_h[i], a_self_h>, t[i,h] = <H_h[i], a_neigh_h>   (gat.py:106-109)
//   gr_gat_aggregate    out_i = sum_j softmax_j(LeakyReLU(s_i + t_j)) H[j]  over the row's neighbours,
//                       heads concatenated or averaged, optional ELU               (gat.py:112-147, 283)
//
// The dense maps are at most 128 x 256 and are applied to N rows: HBM-bound streaming (read the
// row, write the row), weights resident in shared memory, FFMA in fp32 (parity is 1e-5 in fp32,
// which single-pass TF32 tensor-core MMA cannot hold).
#include <math_constants.h>

#include "gr_common.cuh"

namespace gr {

// =============================================================================================
// rowmap
// =============================================================================================
struct RowMapArgs {
    const float *x1, *x2, *x3, *wa, *wb, *ba, *bb, *resid;
    long long ld1, ld2, ld3, ldr, ldo;
    float *out;
    int n_rows, d_in, d_out;
    float alpha, beta, slope;
    int act;  // 0 none, 1 LeakyReLU(slope), 2 ELU
    unsigned drop_thr;        // 0 = no dropout; else element dropped when its 16-bit uniform < drop_thr
    float drop_scale;         // 65536 / (65536 - drop_thr)
    unsigned long long drop_seed;
    const unsigned long long *drop_seed_dev;   // optional: added to drop_seed at run time (CUDA-graph replays)
};

constexpr int RM_ROWS = 64;

__device__ __forceinline__ float act_apply(float z, int act, float slope) {
    if (act == 1) return z > 0.f ? z : z * slope;
    if (act == 2) return z > 0.f ? z : expm1f(z);
    return z;
}

// NG = number of 16-float4 column groups a thread owns: 1 for d_out <= 64 (16 accumulators per map: ~64 registers,
// several CTAs per SM so that one CTA's staging overlaps another's FFMA loop), 2 / 4 for d_out <= 128 / 256.
template <int NG>
__global__ void __launch_bounds__(256) rowmap_kernel(const RowMapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d_in = a.d_in, d_out = a.d_out;
    const bool has_b = a.wb != nullptr;
    float *Wa = reinterpret_cast<float *>(smem_raw);              // [d_in][d_out]
    float *Wb = Wa + (size_t)d_in * d_out;                        // [d_in][d_out] (if has_b)
    float *Xs = Wb + (has_b ? (size_t)d_in * d_out : 0);          // [d_in][RM_ROWS]
    float *Ys = Xs + (size_t)d_in * RM_ROWS;                      // [d_in][RM_ROWS] (if has_b)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * RM_ROWS;

    for (int i = tid; i < d_in * d_out / 4; i += 256) {
        reinterpret_cast<float4 *>(Wa)[i] = __ldg(reinterpret_cast<const float4 *>(a.wa) + i);
        if (has_b) reinterpret_cast<float4 *>(Wb)[i] = __ldg(reinterpret_cast<const float4 *>(a.wb) + i);
    }
    const int f4 = d_in / 4;
    // lane -> row (consecutive lanes stage consecutive rows): the transposed stores Xs[k][r] then hit 32
    // different banks; with lane -> column they were 16-way conflicted (measured: the staging cost as much as
    // the FFMA loop)
    for (int i = tid; i < RM_ROWS * f4; i += 256) {
        const int r = i % RM_ROWS, f = i / RM_ROWS;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), y = v;
        if (row0 + r < a.n_rows) {
            v = __ldg(reinterpret_cast<const float4 *>(a.x1 + (long long)(row0 + r) * a.ld1) + f);
            if (has_b) {
                const float4 p = __ldg(reinterpret_cast<const float4 *>(a.x2 + (long long)(row0 + r) * a.ld2) + f);
                const float4 q = __ldg(reinterpret_cast<const float4 *>(a.x3 + (long long)(row0 + r) * a.ld3) + f);
                y = make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
            }
        }
        Xs[(4 * f + 0) * RM_ROWS + r] = v.x;
        Xs[(4 * f + 1) * RM_ROWS + r] = v.y;
        Xs[(4 * f + 2) * RM_ROWS + r] = v.z;
        Xs[(4 * f + 3) * RM_ROWS + r] = v.w;
        if (has_b) {
            Ys[(4 * f + 0) * RM_ROWS + r] = y.x;
            Ys[(4 * f + 1) * RM_ROWS + r] = y.y;
            Ys[(4 * f + 2) * RM_ROWS + r] = y.z;
            Ys[(4 * f + 3) * RM_ROWS + r] = y.w;
        }
    }
    __syncthreads();

    const int nc4 = d_out / 4;  // float4 column groups; thread tx owns groups tx, tx+16, tx+32, tx+48
    float accA[4][NG][4], accB[4][NG][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int c = 0; c < 4; ++c) accA[m][g][c] = accB[m][g][c] = 0.f;

    for (int k = 0; k < d_in; ++k) {
        const float4 xv = *reinterpret_cast<const float4 *>(Xs + (size_t)k * RM_ROWS + ty * 4);
        const float xm[4] = {xv.x, xv.y, xv.z, xv.w};
        float ym[4] = {0.f, 0.f, 0.f, 0.f};
        if (has_b) {
            const float4 yv = *reinterpret_cast<const float4 *>(Ys + (size_t)k * RM_ROWS + ty * 4);
            ym[0] = yv.x; ym[1] = yv.y; ym[2] = yv.z; ym[3] = yv.w;
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int cg = tx + 16 * g;
            if (cg < nc4) {
                const float4 w = *reinterpret_cast<const float4 *>(Wa + (size_t)k * d_out + cg * 4);
                const float wc[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int c = 0; c < 4; ++c) accA[m][g][c] = __fmaf_rn(xm[m], wc[c], accA[m][g][c]);
                if (has_b) {
                    const float4 w2 = *reinterpret_cast<const float4 *>(Wb + (size_t)k * d_out + cg * 4);
                    const float w2c[4] = {w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int c = 0; c < 4; ++c) accB[m][g][c] = __fmaf_rn(ym[m], w2c[c], accB[m][g][c]);
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int cg = tx + 16 * g;
        if (cg >= nc4) continue;
        float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
        if (a.ba) ba = __ldg(reinterpret_cast<const float4 *>(a.ba) + cg);
        if (a.bb) bb = __ldg(reinterpret_cast<const float4 *>(a.bb) + cg);
        const float bac[4] = {ba.x, ba.y, ba.z, ba.w}, bbc[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int r = row0 + ty * 4 + m;
            if (r >= a.n_rows) continue;
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float z = accA[m][g][c] + bac[c];
                if (has_b) z = z + (accB[m][g][c] + bbc[c]);
                o[c] = a.alpha * act_apply(z, a.act, a.slope);
            }
            if (a.resid) {
                const float4 rv = __ldg(reinterpret_cast<const float4 *>(a.resid + (long long)r * a.ldr) + cg);
                o[0] += a.beta * rv.x; o[1] += a.beta * rv.y; o[2] += a.beta * rv.z; o[3] += a.beta * rv.w;
            }
            if (a.drop_thr) {   // nn.Dropout on the layer output (ngcf.py:86, model.py:198-199): keep/(1-p)
                const unsigned long long bits = drop_bits(a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0ULL),
                                                          (unsigned long long)r * nc4 + cg);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    o[c] = drop_keep(bits, c, a.drop_thr) ? o[c] * a.drop_scale : 0.f;
            }
            *reinterpret_cast<float4 *>(a.out + (long long)r * a.ldo + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace gr

using namespace gr;

extern "C" int gr_rowmap_f32(const float *x1, int64_t ld1, const float *wa, const float *bias_a, const float *x2,
                             int64_t ld2, const float *x3, int64_t ld3, const float *wb, const float *bias_b,
                             const float *resid, int64_t ldr, float alpha, float beta, int32_t act, float slope,
                             int64_t n_rows, int32_t d_in, int32_t d_out, float drop_p, uint64_t drop_seed,
                             const uint64_t *drop_seed_dev, float *out, int64_t ldo, void *stream) {
    if (!x1 || !wa || !out || n_rows < 0) return GR_ERR_INVALID;
    if (!(drop_p >= 0.f) || drop_p >= 1.f) return GR_ERR_INVALID;
    if (wb && (!x2 || !x3)) return GR_ERR_INVALID;
    if (n_rows == 0) return GR_OK;
    if (d_in <= 0 || d_out <= 0 || (d_in & 3) || (d_out & 3) || d_in > 256 || d_out > 256) return GR_ERR_UNSUPPORTED;
    if ((ld1 & 3) || (ldo & 3) || ld1 < d_in || ldo < d_out || (resid && ((ldr & 3) || ldr < d_out))) return GR_ERR_INVALID;
    if (wb && ((ld2 & 3) || (ld3 & 3) || ld2 < d_in || ld3 < d_in)) return GR_ERR_INVALID;
    if (act < 0 || act > 2) return GR_ERR_INVALID;
    if (n_rows > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if (!aligned16(x1) || !aligned16(wa) || !aligned16(out) || !aligned16(x2) || !aligned16(x3) || !aligned16(wb) ||
        !aligned16(bias_a) || !aligned16(bias_b) || !aligned16(resid))
        return GR_ERR_INVALID;
    RowMapArgs a;
    a.x1 = x1; a.x2 = x2; a.x3 = x3; a.wa = wa; a.wb = wb; a.ba = bias_a; a.bb = bias_b; a.resid = resid;
    a.ld1 = ld1; a.ld2 = ld2; a.ld3 = ld3; a.ldr = ldr; a.ldo = ldo;
    a.out = out; a.n_rows = (int)n_rows; a.d_in = d_in; a.d_out = d_out;
    a.alpha = alpha; a.beta = beta; a.slope = slope; a.act = act;
    a.drop_thr = drop_threshold(drop_p); a.drop_scale = drop_scale_of(a.drop_thr); a.drop_seed = drop_seed;
    a.drop_seed_dev = reinterpret_cast<const unsigned long long *>(drop_seed_dev);
    const size_t smem = ((size_t)d_in * d_out + (size_t)d_in * RM_ROWS) * 4 * (wb ? 2 : 1);
    if (smem > 227 * 1024) return GR_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)((n_rows + RM_ROWS - 1) / RM_ROWS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d_out <= 64) {
        GR_CUDA_CHECK(cudaFuncSetAttribute(rowmap_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rowmap_kernel<1><<<grid, 256, smem, st>>>(a);
    } else if (d_out <= 128) {
        GR_CUDA_CHECK(cudaFuncSetAttribute(rowmap_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rowmap_kernel<2><<<grid, 256, smem, st>>>(a);
    } else {
        GR_CUDA_CHECK(cudaFuncSetAttribute(rowmap_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rowmap_kernel<4><<<grid, 256, smem, st>>>(a);
    }
    GR_LAUNCH_CHECK();
    return GR_OK;
}


// =============================================================================================
// layer combination:  out = sum_l w_l * x_l   (GAT: mean of the L+1 layer outputs, gat.py:287-288;
// Group-and-Shuffle: sum_l softmax(layer_weights)_l x_l, model.py:204-207) and its backward
// =============================================================================================
namespace gr {
constexpr int LC_MAX = 8;
struct CombineArgs {
    const float4 *x[LC_MAX];
    long long ld4[LC_MAX];
    const float *w;        // device [n_in] weights, or NULL: out = (x_0 + x_1 + ...) / n_in (torch.mean of the stack)
    int n_in;
    long long n_rows;
    int f4;
    float4 *out;
    long long ldo4;
};

// Forward.  Weighted: ((w0 x0) + (w1 x1)) + ... with separately rounded products, the order of the reference's
// python sum(); mean: left-to-right sum divided by n_in.
__global__ void __launch_bounds__(256) layer_combine_kernel(const CombineArgs a) {
    const long long total = a.n_rows * a.f4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float w[LC_MAX];
#pragma unroll
    for (int l = 0; l < LC_MAX; ++l) w[l] = (a.w && l < a.n_in) ? __ldg(a.w + l) : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / a.f4;
        const int f = (int)(i % a.f4);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < LC_MAX; ++l)
            if (l < a.n_in) {
                const float4 v = __ldg(a.x[l] + r * a.ld4[l] + f);
                if (a.w) {
                    const float4 p = make_float4(__fmul_rn(w[l], v.x), __fmul_rn(w[l], v.y), __fmul_rn(w[l], v.z), __fmul_rn(w[l], v.w));
                    s = l == 0 ? p : make_float4(__fadd_rn(s.x, p.x), __fadd_rn(s.y, p.y), __fadd_rn(s.z, p.z), __fadd_rn(s.w, p.w));
                } else {
                    s = l == 0 ? v : make_float4(__fadd_rn(s.x, v.x), __fadd_rn(s.y, v.y), __fadd_rn(s.z, v.z), __fadd_rn(s.w, v.w));
                }
            }
        if (!a.w) {
            const float n = (float)a.n_in;
            s = make_float4(__fdiv_rn(s.x, n), __fdiv_rn(s.y, n), __fdiv_rn(s.z, n), __fdiv_rn(s.w, n));
        }
        a.out[r * a.ldo4 + f] = s;
    }
}

// Backward of the weighted form: per-CTA partials of dw_l = <g, x_l> (double accumulation per thread, block tree),
// reduced by launch_reduce_partials.  (dx_l = w_l g is the forward kernel with one input.)
__global__ void __launch_bounds__(256) layer_combine_dw_kernel(const CombineArgs a, const float4 *g, long long ldg4,
                                                               float *partial) {
    __shared__ float red[LC_MAX][8];
    const long long total = a.n_rows * a.f4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float acc[LC_MAX];
#pragma unroll
    for (int l = 0; l < LC_MAX; ++l) acc[l] = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / a.f4;
        const int f = (int)(i % a.f4);
        const float4 gv = __ldg(g + r * ldg4 + f);
#pragma unroll
        for (int l = 0; l < LC_MAX; ++l)
            if (l < a.n_in) {
                const float4 v = __ldg(a.x[l] + r * a.ld4[l] + f);
                acc[l] += gv.x * v.x + gv.y * v.y + gv.z * v.z + gv.w * v.w;
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int l = 0; l < LC_MAX; ++l) {
        float v = acc[l];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[l][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < LC_MAX) {
        float t = 0.f;
        for (int w8 = 0; w8 < 8; ++w8) t += red[threadIdx.x][w8];
        partial[(long long)blockIdx.x * LC_MAX + threadIdx.x] = t;
    }
}
}  // namespace gr

extern "C" int gr_layer_combine(const float *const *x_host, const int64_t *ld_host, int32_t n_in, const float *w_dev,
                                int64_t n_rows, int32_t d, float *out, int64_t ldo, void *stream) {
    if (!x_host || !ld_host || !out || n_in < 1 || n_in > LC_MAX || n_rows < 0 || d <= 0 || (d & 3)) return GR_ERR_INVALID;
    if ((ldo & 3) || ldo < d || !aligned16(out)) return GR_ERR_INVALID;
    if (n_rows == 0) return GR_OK;
    CombineArgs a = {};
    for (int l = 0; l < n_in; ++l) {
        if (!x_host[l] || !aligned16(x_host[l]) || (ld_host[l] & 3) || ld_host[l] < d) return GR_ERR_INVALID;
        a.x[l] = reinterpret_cast<const float4 *>(x_host[l]);
        a.ld4[l] = ld_host[l] / 4;
    }
    a.w = w_dev; a.n_in = n_in; a.n_rows = n_rows; a.f4 = d / 4;
    a.out = reinterpret_cast<float4 *>(out); a.ldo4 = ldo / 4;
    const long long total = n_rows * (d / 4);
    const int grid = persistent_grid((int)((total + 255) / 256 > 0x7fffffff ? 0x7fffffff : (total + 255) / 256), 8);
    layer_combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" size_t gr_layer_combine_bwd_workspace_bytes(void) { return (size_t)(kSmCountFallback * 4 + 64) * LC_MAX * 4 + 256; }

extern "C" int gr_layer_combine_dw(const float *const *x_host, const int64_t *ld_host, int32_t n_in, const float *g,
                                   int64_t ldg, int64_t n_rows, int32_t d, float *dw, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    if (!x_host || !ld_host || !g || !dw || !workspace || n_in < 1 || n_in > LC_MAX || n_rows < 0 || d <= 0 || (d & 3))
        return GR_ERR_INVALID;
    if ((ldg & 3) || ldg < d || !aligned16(g) || !aligned16(workspace)) return GR_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_rows == 0) {
        GR_CUDA_CHECK(cudaMemsetAsync(dw, 0, LC_MAX * 4, st));
        return GR_OK;
    }
    CombineArgs a = {};
    for (int l = 0; l < n_in; ++l) {
        if (!x_host[l] || !aligned16(x_host[l]) || (ld_host[l] & 3) || ld_host[l] < d) return GR_ERR_INVALID;
        a.x[l] = reinterpret_cast<const float4 *>(x_host[l]);
        a.ld4[l] = ld_host[l] / 4;
    }
    a.n_in = n_in; a.n_rows = n_rows; a.f4 = d / 4;
    const long long total = n_rows * (d / 4);
    int grid = persistent_grid((int)((total + 255) / 256 > 0x7fffffff ? 0x7fffffff : (total + 255) / 256), 4);
    const size_t cap = (workspace_bytes - 256) / (LC_MAX * 4);
    if ((size_t)grid > cap) grid = (int)cap;
    if (grid < 1) return GR_ERR_WORKSPACE;
    float *partial = static_cast<float *>(workspace);
    layer_combine_dw_kernel<<<grid, 256, 0, st>>>(a, reinterpret_cast<const float4 *>(g), ldg / 4, partial);
    GR_LAUNCH_CHECK();
    if (launch_reduce_partials(partial, grid, LC_MAX, dw, st)) {      // dw: float[8], entries >= n_in are zero
        set_last_cuda_error(cudaPeekAtLastError());
        return GR_ERR_CUDA;
    }
    return GR_OK;
}
