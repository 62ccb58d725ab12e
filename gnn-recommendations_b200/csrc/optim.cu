// Fused gradient-norm clip + Adam (L2 weight decay added to the gradient) over a list of fp32 tensors.
//
// Replaces trainer.py:273-276 (`clip_grad_norm_(model.parameters(), max_norm)`; `optimizer.step()` with
// `optim.Adam(lr, weight_decay)`, trainer.py:81-85).  Stock torch runs it as ~15 launches and ~7 passes
// over the [N, d] tables; here it is one pass that reads the gradients (sum of squares, deterministic
// two-level reduction in double), a one-block finalize, and one pass that reads g, p, m, v and writes
// p, m, v (28 B per element).  The arithmetic follows torch's single-tensor Adam statement by statement:
//   g   = clip_coef * g + wd * p
//   m   = m + (g - m) * (1 - beta1)                       (lerp)
//   v   = v * beta2 + (1 - beta2) * g * g                 (mul_, addcmul_)
//   p   = p - step_size * m / (sqrt(v) / sqrt(bias_correction2) + eps)
// with clip_coef = min(1, max_norm / (total_norm + 1e-6)); step_size = lr / bias_correction1 and
// sqrt(bias_correction2) are computed by the host in double as torch does.
#include "gr_common.cuh"

namespace gr {

constexpr int OPT_MAX_TENSORS = 32;         // per launch; longer lists go in several launches
constexpr int OPT_THREADS = 256;
constexpr int OPT_BLOCK_ELEMS = OPT_THREADS * 16;   // 4 float4 per thread

struct OptTable {
    float *p[OPT_MAX_TENSORS];
    const float *g[OPT_MAX_TENSORS];
    float *m[OPT_MAX_TENSORS];
    float *v[OPT_MAX_TENSORS];
    long long numel[OPT_MAX_TENSORS];
    long long first_block[OPT_MAX_TENSORS + 1];   // prefix of per-tensor block counts
    int n;
};

__device__ __forceinline__ int opt_find_tensor(const OptTable &t, long long b) {
    int i = 0;
    while (i + 1 < t.n && b >= t.first_block[i + 1]) ++i;
    return i;
}

// partial[part_base + blockIdx.x] = sum of squares of this block's slice of the gradients
__global__ void __launch_bounds__(OPT_THREADS) opt_sumsq_kernel(const OptTable t, double *partial, long long part_base) {
    const int ti = opt_find_tensor(t, blockIdx.x);
    const float *g = t.g[ti];
    const long long n = t.numel[ti];
    const long long base = (blockIdx.x - t.first_block[ti]) * OPT_BLOCK_ELEMS;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long i = base + ((long long)q * OPT_THREADS + threadIdx.x) * 4;
        if (i + 3 < n) {
            const float4 x = *reinterpret_cast<const float4 *>(g + i);
            acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        } else {
            for (long long j = i; j < n; ++j) acc += g[j] * g[j];
        }
    }
    double d = (double)acc;
#pragma unroll
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    __shared__ double ws[OPT_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < OPT_THREADS / 32; ++w) s += ws[w];
        partial[part_base + blockIdx.x] = s;
    }
}

// one block: fixed-order sum of the partials -> out[0] = total norm, out[1] = clip coefficient
__global__ void __launch_bounds__(1024) opt_finalize_kernel(const double *partial, long long n, float max_norm, float *out) {
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) acc += partial[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += ws[w];
        const float total = (float)sqrt(s);
        float coef = max_norm / (total + 1e-6f);      // clip_grad_norm_: max_norm / (total_norm + 1e-6), clamped to 1
        if (!(coef < 1.0f)) coef = 1.0f;
        out[0] = total;
        out[1] = coef;
    }
}

struct AdamScalars {   // each rounded once from the caller's doubles, as torch rounds its python scalars
    float step_size, one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, bc2_sqrt;
    const float *step_dev;   // optional device [2] = (step_size, bc2_sqrt): the step-dependent scalars read at
                             // run time, so that a CUDA graph captured once stays valid for every later step
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float coef, const AdamScalars &s) {
    g = g * coef;
    g = __fmaf_rn(s.weight_decay, p, g);
    m = __fmaf_rn(s.one_minus_beta1, g - m, m);
    v = __fmaf_rn(s.one_minus_beta2 * g, g, v * s.beta2);
    const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
    p = __fmaf_rn(-s.step_size, m / denom, p);
}

__global__ void __launch_bounds__(OPT_THREADS) opt_adam_kernel(const OptTable t, const float *clip, AdamScalars s) {
    if (s.step_dev) {
        s.step_size = __ldg(s.step_dev);
        s.bc2_sqrt = __ldg(s.step_dev + 1);
    }
    const int ti = opt_find_tensor(t, blockIdx.x);
    float *p = t.p[ti], *m = t.m[ti], *v = t.v[ti];
    const float *g = t.g[ti];
    const long long n = t.numel[ti];
    const long long base = (blockIdx.x - t.first_block[ti]) * OPT_BLOCK_ELEMS;
    const float coef = clip ? clip[1] : 1.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long i = base + ((long long)q * OPT_THREADS + threadIdx.x) * 4;
        if (i + 3 < n) {
            float4 pp = *reinterpret_cast<float4 *>(p + i), mm = *reinterpret_cast<float4 *>(m + i),
                   vv = *reinterpret_cast<float4 *>(v + i);
            const float4 gg = *reinterpret_cast<const float4 *>(g + i);
            adam_one(pp.x, gg.x, mm.x, vv.x, coef, s);
            adam_one(pp.y, gg.y, mm.y, vv.y, coef, s);
            adam_one(pp.z, gg.z, mm.z, vv.z, coef, s);
            adam_one(pp.w, gg.w, mm.w, vv.w, coef, s);
            *reinterpret_cast<float4 *>(p + i) = pp;
            *reinterpret_cast<float4 *>(m + i) = mm;
            *reinterpret_cast<float4 *>(v + i) = vv;
        } else {
            for (long long j = i; j < n; ++j) adam_one(p[j], g[j], m[j], v[j], coef, s);
        }
    }
}

static long long opt_blocks(int64_t numel) { return (numel + OPT_BLOCK_ELEMS - 1) / OPT_BLOCK_ELEMS; }

}  // namespace gr

using namespace gr;

extern "C" size_t gr_clip_adam_workspace_bytes(const int64_t *numel_host, int32_t n_tensors) {
    if (!numel_host || n_tensors <= 0) return 0;
    long long blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (numel_host[i] < 0) return 0;
        blocks += opt_blocks(numel_host[i]);
    }
    return (size_t)blocks * sizeof(double) + 256;
}

extern "C" int gr_clip_adam_fused(void *const *params_host, const void *const *grads_host, void *const *exp_avg_host,
                                  void *const *exp_avg_sq_host, const int64_t *numel_host, int32_t n_tensors,
                                  double max_norm, double step_size, double beta1, double beta2, double eps,
                                  double weight_decay, double bias_correction2_sqrt, const float *step_scalars_dev,
                                  float *norm_out, void *workspace, size_t workspace_bytes, void *stream) {
    if (!params_host || !grads_host || !exp_avg_host || !exp_avg_sq_host || !numel_host || n_tensors <= 0 || !workspace)
        return GR_ERR_INVALID;
    if (workspace_bytes < gr_clip_adam_workspace_bytes(numel_host, n_tensors)) return GR_ERR_WORKSPACE;
    for (int i = 0; i < n_tensors; ++i) {
        if (numel_host[i] < 0) return GR_ERR_INVALID;
        if (numel_host[i] == 0) continue;
        if (!params_host[i] || !grads_host[i] || !exp_avg_host[i] || !exp_avg_sq_host[i]) return GR_ERR_INVALID;
        if (!aligned16(params_host[i]) || !aligned16(grads_host[i]) || !aligned16(exp_avg_host[i]) ||
            !aligned16(exp_avg_sq_host[i]))
            return GR_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *clip = reinterpret_cast<float *>(workspace);            // [0] total norm, [1] coefficient
    double *partial = reinterpret_cast<double *>(static_cast<char *>(workspace) + 256);
    const bool do_clip = max_norm > 0.0;

    // both passes walk the list in chunks of at most OPT_MAX_TENSORS non-empty tensors; `i` is the next input
    // index, so a chunk that skipped empty tensors does not overlap the following one
    for (int pass = do_clip ? 0 : 1; pass < 2; ++pass) {
        long long part_base = 0;
        int i = 0;
        while (i < n_tensors) {
            OptTable t;
            t.n = 0;
            long long blocks = 0;
            for (; i < n_tensors && t.n < OPT_MAX_TENSORS; ++i) {
                if (numel_host[i] == 0) continue;
                t.p[t.n] = static_cast<float *>(params_host[i]);
                t.g[t.n] = static_cast<const float *>(grads_host[i]);
                t.m[t.n] = static_cast<float *>(exp_avg_host[i]);
                t.v[t.n] = static_cast<float *>(exp_avg_sq_host[i]);
                t.numel[t.n] = numel_host[i];
                t.first_block[t.n] = blocks;
                blocks += opt_blocks(numel_host[i]);
                ++t.n;
            }
            t.first_block[t.n] = blocks;
            if (t.n == 0 || blocks == 0) continue;
            if (blocks > 0x7fffffffLL) return GR_ERR_OVERFLOW;
            if (pass == 0) {
                opt_sumsq_kernel<<<(unsigned)blocks, OPT_THREADS, 0, s>>>(t, partial, part_base);
            } else {
                AdamScalars sc{(float)step_size, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                               (float)weight_decay, (float)bias_correction2_sqrt, step_scalars_dev};
                opt_adam_kernel<<<(unsigned)blocks, OPT_THREADS, 0, s>>>(t, do_clip ? clip : nullptr, sc);
            }
            GR_LAUNCH_CHECK();
            part_base += blocks;
        }
        if (pass == 0) {
            opt_finalize_kernel<<<1, 1024, 0, s>>>(partial, part_base, (float)max_norm, clip);
            GR_LAUNCH_CHECK();
        }
    }
    if (norm_out && do_clip) GR_CUDA_CHECK(cudaMemcpyAsync(norm_out, clip, sizeof(float), cudaMemcpyDeviceToDevice, s));
    return GR_OK;
}
