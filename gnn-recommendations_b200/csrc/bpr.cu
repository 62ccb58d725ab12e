// Fused BPR step body: positive/negative embedding gather, the reference's B x B loss and the
// scatter-add of its gradient, one cooperative kernel.
//
// Reference semantics (src/training/trainer.py:257-264 + src/training/losses.py:44-53): neg is
// [B,1], so pos[B] - neg[B,1] broadcasts to [B,B]:
//     loss = (1/B^2) sum_i sum_j softplus(n_i - p_j),   p_j = <E[u_j], E[U+pos_j]>, n_i = <E[u_i], E[U+neg_i]>
//     dL/dp_j = -(1/B^2) sum_i sigma(n_i - p_j)         dL/dn_i = (1/B^2) sum_j sigma(n_i - p_j)
//     G[u_j] += dp_j E[U+pos_j] + dn_j E[U+neg_j];  G[U+pos_j] += dp_j E[u_j];  G[U+neg_j] += dn_j E[u_j]
// E = propagated embeddings [N, d] (users first), G = dL/dE (zeroed by the caller).
//
// One warp per sample.  Phase 1: the two dot products.  grid.sync().  Phase 2: every warp reads
// all 2B scores (L2), forms its sample's column sum (dp), row sum (dn) and loss row with a fixed
// lane-strided order + xor-shuffle tree (deterministic), then scatters its three gradient rows
// with vector red.global.add (duplicates accumulate; their order is the only non-determinism).
// The last CTA to finish adds the B loss rows in index order (double) -> loss.
#include <cooperative_groups.h>

#include "gr_common.cuh"

namespace cg = cooperative_groups;

namespace gr {

struct BprArgs {
    const float *emb;
    long long ld;
    long long n_users, n_items;
    const int64_t *users, *pos, *neg;
    int batch;
    int d;
    float *grad;
    long long ldg;
    float *loss;
    float *scores;     // [2B]: p then n
    float *loss_rows;  // [B]
    unsigned int *counter;
    float grad_scale;  // upstream dL (1.0 for a plain backward)
};

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void red_add_f4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(256) bpr_fused_kernel(const BprArgs a) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int B = a.batch;
    const int nf4 = a.d >> 2;
    const int stride = gridDim.x * 8;          // one cooperative wave; larger batches stride over the samples
    const float nan = __int_as_float(0x7fc00000);

    // phase 1: p_j = <U[u_j], I[pos_j]>, n_j = <U[u_j], I[neg_j]>.  An id outside [0, n_users) / [0, n_items)
    // poisons the loss with NaN (loud) and is never dereferenced.
    for (int j = blockIdx.x * 8 + warp; j < B; j += stride) {
        const long long u = a.users[j], pi = a.pos[j], ni = a.neg[j];
        float sp = nan, sn = nan;
        if (u >= 0 && u < a.n_users && pi >= 0 && pi < a.n_items && ni >= 0 && ni < a.n_items) {
            const float4 *eu = reinterpret_cast<const float4 *>(a.emb + u * a.ld);
            const float4 *ep = reinterpret_cast<const float4 *>(a.emb + (a.n_users + pi) * a.ld);
            const float4 *en = reinterpret_cast<const float4 *>(a.emb + (a.n_users + ni) * a.ld);
            sp = 0.f; sn = 0.f;
            for (int f = lane; f < nf4; f += 32) {
                const float4 uu = __ldg(eu + f), p = __ldg(ep + f), n = __ldg(en + f);
                sp += uu.x * p.x + uu.y * p.y + uu.z * p.z + uu.w * p.w;
                sn += uu.x * n.x + uu.y * n.y + uu.z * n.z + uu.w * n.w;
            }
            sp = warp_sum(sp);
            sn = warp_sum(sn);
        }
        if (lane == 0) {
            a.scores[j] = sp;
            a.scores[B + j] = sn;
        }
    }
    grid.sync();

    // phase 2: the B x B sums of sample j's row / column, its loss row and its three gradient rows
    for (int j = blockIdx.x * 8 + warp; j < B; j += stride) {
        const float pj = a.scores[j], nj = a.scores[B + j];
        float col = 0.f, row = 0.f, lrow = 0.f;
        for (int t = lane; t < B; t += 32) {
            const float pt = __ldcg(a.scores + t), nt = __ldcg(a.scores + B + t);
            col += sigmoid_f(nt - pj);       // sum_i sigma(n_i - p_j)
            const float x = nj - pt;         // row i = j: sum_t softplus / sigma (n_j - p_t)
            row += sigmoid_f(x);
            lrow += softplus_f(x);
        }
        col = warp_sum(col);
        row = warp_sum(row);
        lrow = warp_sum(lrow);
        const float inv = a.grad_scale / ((float)B * (float)B);
        const float dp = -col * inv, dn = row * inv;
        if (lane == 0) a.loss_rows[j] = lrow;
        if (pj != pj) continue;              // invalid ids (NaN score): nothing to scatter
        const long long ru = a.users[j], rp = a.n_users + a.pos[j], rn = a.n_users + a.neg[j];
        const float4 *eu = reinterpret_cast<const float4 *>(a.emb + ru * a.ld);
        const float4 *ep = reinterpret_cast<const float4 *>(a.emb + rp * a.ld);
        const float4 *en = reinterpret_cast<const float4 *>(a.emb + rn * a.ld);
        float *gu = a.grad + ru * a.ldg, *gp = a.grad + rp * a.ldg, *gn = a.grad + rn * a.ldg;
        for (int f = lane; f < nf4; f += 32) {
            const float4 u = __ldg(eu + f), p = __ldg(ep + f), n = __ldg(en + f);
            red_add_f4(gu + 4 * f, make_float4(dp * p.x + dn * n.x, dp * p.y + dn * n.y, dp * p.z + dn * n.z,
                                               dp * p.w + dn * n.w));
            red_add_f4(gp + 4 * f, make_float4(dp * u.x, dp * u.y, dp * u.z, dp * u.w));
            red_add_f4(gn + 4 * f, make_float4(dn * u.x, dn * u.y, dn * u.z, dn * u.w));
        }
    }
    // last CTA: deterministic loss reduction
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last && warp == 0) {
        double s = 0.0;
        for (int t = lane; t < B; t += 32) s += (double)__ldcg(a.loss_rows + t);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            *a.loss = (float)(s / ((double)B * (double)B));
            *a.counter = 0;  // ready for the next launch
        }
    }
}

}  // namespace gr

extern "C" size_t gr_bpr_workspace_bytes(int64_t batch) {
    if (batch < 0) return 0;
    return (size_t)batch * 3 * sizeof(float) + 256;
}

// Replaces trainer.py:257-264 + losses.py:44-53 (forward) and the autograd backward through
// them.  `workspace` must be zero-initialised once (its trailing counter is self-resetting).
extern "C" int gr_bpr_fused(const float *emb, int64_t ld, int64_t n_users, int64_t n_items, const int64_t *users,
                            const int64_t *pos, const int64_t *neg, int64_t batch, int32_t d, float grad_scale,
                            float *grad, int64_t ldg, float *loss, void *workspace, size_t workspace_bytes,
                            void *stream) {
    using namespace gr;
    if (!emb || !users || !pos || !neg || !grad || !loss || !workspace) return GR_ERR_INVALID;
    if (batch <= 0 || batch > (1 << 20) || d <= 0 || (d & 3) || (ld & 3) || (ldg & 3) || ld < d || ldg < d)
        return GR_ERR_INVALID;
    if (n_users < 0 || n_items < 0) return GR_ERR_INVALID;
    if (!aligned16(emb) || !aligned16(grad)) return GR_ERR_INVALID;
    if (workspace_bytes < gr_bpr_workspace_bytes(batch)) return GR_ERR_WORKSPACE;
    BprArgs a;
    a.emb = emb;
    a.ld = ld;
    a.n_users = n_users;
    a.n_items = n_items;
    a.users = users;
    a.pos = pos;
    a.neg = neg;
    a.batch = (int)batch;
    a.d = d;
    a.grad = grad;
    a.ldg = ldg;
    a.loss = loss;
    a.scores = static_cast<float *>(workspace);
    a.loss_rows = a.scores + 2 * batch;
    a.counter = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(workspace) +
                                                 (((size_t)batch * 3 * sizeof(float) + 15) & ~(size_t)15));
    a.grad_scale = grad_scale;
    int ctas = (int)((batch + 7) / 8);
    int max_per_sm = 0;
    GR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_per_sm, bpr_fused_kernel, 256, 0));
    const long long wave = (long long)max_per_sm * sm_count();
    if (wave < 1) return GR_ERR_UNSUPPORTED;
    if (ctas > wave) ctas = (int)wave;        // one cooperative wave; the kernel strides over the remaining samples
    void *params[] = {(void *)&a};
    GR_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)bpr_fused_kernel, dim3(ctas), dim3(256), params, 0,
                                              static_cast<cudaStream_t>(stream)));
    return GR_OK;
}
