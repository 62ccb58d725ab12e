// Group-and-Shuffle orthogonal maps (src/models/orthogonal_bundle/group_shuffle_layer.py:88-129,
// bundle_layer.py:56-73, model.py:176): the per-layer dense map of OrthogonalBundleGNN
//
//     M_l = W_conn,l  W_orth,l[:, perm_g]      W_conn = blockdiag(exp(P_k - P_k^T))[:, perm_c]
//                                              W_orth = blockdiag(exp(Q_k - Q_k^T))
//
// built on the device in two launches (the reference: 48 torch.matrix_exp calls + block_diag + two index
// ops + one 64x64 GEMM per forward, several hundred tiny launches), and its backward (dL/dM -> dL/dP,
// dL/dQ) in two more:
//   gs_block_exp_kernel      one warp per bs x bs block: exp of the skew matrix, scaling-and-squaring with a
//                            degree-12 Taylor polynomial evaluated in double (|A/2^s|_F <= 0.5 -> 2e-14)
//   gs_compose_kernel        M[i][j] = sum_k W_conn[i][k] W_orth[k][perm_g[j]]   (fp32 fmaf, k ascending)
//   gs_compose_bwd_kernel    dL/d(block entries) from dL/dM  (the block structure and the permutations make
//                            every block entry one d-term dot product)
//   gs_block_exp_bwd_kernel  adjoint of the Frechet derivative of exp:  dA = L_exp(A^T, G), from the coupled
//                            Taylor recurrences  T_k = A T_{k-1}/k,  D_k = (A D_{k-1} + G T_{k-1})/k  and the
//                            squaring steps  L <- E L + L E,  E <- E E;   dP = dA - dA^T
#include "gr_common.cuh"

namespace gr {

constexpr int GS_MAX_BS = 16;
constexpr int GS_TAYLOR = 12;

// C = A * B (bs x bs, row-major, shared memory), lanes stride over the outputs.
__device__ __forceinline__ void gs_mm(double *C, const double *A, const double *B, int bs, int lane) {
    for (int e = lane; e < bs * bs; e += 32) {
        const int r = e / bs, c = e % bs;
        double s = 0.0;
        for (int k = 0; k < bs; ++k) s += A[r * bs + k] * B[k * bs + c];
        C[e] = s;
    }
    __syncwarp();
}

__device__ __forceinline__ int gs_scaling(const double *A, int bs, int lane) {
    double ss = 0.0;
    for (int e = lane; e < bs * bs; e += 32) ss += A[e] * A[e];
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const double fro = sqrt(ss);
    int s = 0;
    double lim = 0.5;
    while (fro > lim && s < 60) { lim *= 2.0; ++s; }
    return s;
}

// blocks[b] = exp(P_b - P_b^T), one warp (one CTA of 32 threads) per block
__global__ void __launch_bounds__(32) gs_block_exp_kernel(const float *skew, int bs, float *blocks) {
    __shared__ double A[GS_MAX_BS * GS_MAX_BS], T[GS_MAX_BS * GS_MAX_BS], E[GS_MAX_BS * GS_MAX_BS],
        X[GS_MAX_BS * GS_MAX_BS];
    const int lane = threadIdx.x;
    const float *P = skew + (size_t)blockIdx.x * bs * bs;
    for (int e = lane; e < bs * bs; e += 32) {
        const int r = e / bs, c = e % bs;
        A[e] = (double)(P[r * bs + c] - P[c * bs + r]);      // fp32 subtraction like the reference (A - A^T)
    }
    __syncwarp();
    const int s = gs_scaling(A, bs, lane);
    const double sc = ldexp(1.0, -s);
    for (int e = lane; e < bs * bs; e += 32) {
        A[e] *= sc;
        const double id = (e / bs == e % bs) ? 1.0 : 0.0;
        T[e] = id;
        E[e] = id;
    }
    __syncwarp();
    for (int k = 1; k <= GS_TAYLOR; ++k) {
        gs_mm(X, A, T, bs, lane);
        for (int e = lane; e < bs * bs; e += 32) {
            T[e] = X[e] / (double)k;
            E[e] += T[e];
        }
        __syncwarp();
    }
    for (int q = 0; q < s; ++q) {
        gs_mm(X, E, E, bs, lane);
        for (int e = lane; e < bs * bs; e += 32) E[e] = X[e];
        __syncwarp();
    }
    float *out = blocks + (size_t)blockIdx.x * bs * bs;
    for (int e = lane; e < bs * bs; e += 32) out[e] = (float)E[e];
}

struct GsComposeArgs {
    const float *blocks;            // [L, n_sets, nb, bs, bs]; set 0 = connection (if n_sets == 2), last = local
    const long long *perm_c, *perm_g;   // [L, d] (torch.randperm buffers, int64); perm_c NULL when n_sets == 1
    int d, bs, n_sets;
    float *m;                       // [L, d, d]
    const float *dm;                // backward: [L, d, d]
    float *dblocks;                 // backward: [L, n_sets, nb, bs, bs]
};

// grid (d, L), block d threads: thread j of CTA (i, l) forms M_l[i][j]
__global__ void gs_compose_kernel(const GsComposeArgs a) {
    const int d = a.d, bs = a.bs, nb = d / bs;
    const int i = blockIdx.x, l = blockIdx.y, j = threadIdx.x;
    if (j >= d) return;
    const float *Eg = a.blocks + ((size_t)l * a.n_sets + (a.n_sets - 1)) * nb * bs * bs;
    const int cj = (int)a.perm_g[(size_t)l * d + j];             // column of W_orth feeding output column j
    float acc = 0.f;
    if (a.n_sets == 1) {
        acc = (i / bs == cj / bs) ? Eg[((size_t)(i / bs) * bs + i % bs) * bs + cj % bs] : 0.f;
    } else {
        const float *Ec = a.blocks + (size_t)l * a.n_sets * nb * bs * bs;
        const long long *pc = a.perm_c + (size_t)l * d;
        for (int k = 0; k < d; ++k) {
            const int ck = (int)pc[k];
            if (ck / bs != i / bs || k / bs != cj / bs) continue;
            const float wc = Ec[((size_t)(i / bs) * bs + i % bs) * bs + ck % bs];
            const float wo = Eg[((size_t)(k / bs) * bs + k % bs) * bs + cj % bs];
            acc = __fmaf_rn(wc, wo, acc);
        }
    }
    a.m[((size_t)l * d + i) * d + j] = acc;
}

// one thread per block entry (l, set, b, r, c)
__global__ void gs_compose_bwd_kernel(const GsComposeArgs a, int n_layers) {
    const int d = a.d, bs = a.bs, nb = d / bs;
    const long long per_layer = (long long)a.n_sets * nb * bs * bs;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_layer * n_layers) return;
    const int l = (int)(idx / per_layer);
    int rem = (int)(idx % per_layer);
    const int set = rem / (nb * bs * bs);
    rem %= nb * bs * bs;
    const int b = rem / (bs * bs), r = (rem / bs) % bs, c = rem % bs;
    const float *dm = a.dm + (size_t)l * d * d;
    const long long *pg = a.perm_g + (size_t)l * d;
    const float *Eg = a.blocks + ((size_t)l * a.n_sets + (a.n_sets - 1)) * nb * bs * bs;
    float acc = 0.f;
    const bool is_local = (set == a.n_sets - 1);
    if (!is_local) {
        // dEc[b][r][c] = dW_conn[i][k], i = b*bs + r, k = perm_c^-1[b*bs + c];  dW_conn[i][k] = sum_j dM[i][j] W_orth[k][pg[j]]
        const long long *pc = a.perm_c + (size_t)l * d;
        const int i = b * bs + r;
        int k = 0;
        while (k < d && (int)pc[k] != b * bs + c) ++k;
        for (int j = 0; j < d; ++j) {
            const int cj = (int)pg[j];
            if (cj / bs != k / bs) continue;
            acc = __fmaf_rn(dm[(size_t)i * d + j], Eg[((size_t)(k / bs) * bs + k % bs) * bs + cj % bs], acc);
        }
    } else {
        // dEg[b][r][c] = dW_orth[k][c'], k = b*bs + r, c' = b*bs + c, j = perm_g^-1[c'];  = sum_i W_conn[i][k] dM[i][j]
        const int k = b * bs + r;
        int j = 0;
        while (j < d && (int)pg[j] != b * bs + c) ++j;
        if (a.n_sets == 1) {
            acc = dm[(size_t)k * d + j];
        } else {
            const float *Ec = a.blocks + (size_t)l * a.n_sets * nb * bs * bs;
            const int ck = (int)a.perm_c[(size_t)l * d + k];
            const int bi = ck / bs;                      // rows i of W_conn with a non-zero in column k
            for (int ri = 0; ri < bs; ++ri) {
                const int i = bi * bs + ri;
                acc = __fmaf_rn(Ec[((size_t)bi * bs + ri) * bs + ck % bs], dm[(size_t)i * d + j], acc);
            }
        }
    }
    a.dblocks[idx] = acc;
}

// dP = dA - dA^T with dA = L_exp(A^T, G), A = P - P^T, G = dL/d exp(A); one warp per block
__global__ void __launch_bounds__(32) gs_block_exp_bwd_kernel(const float *skew, const float *dblocks, int bs,
                                                              float *dskew) {
    constexpr int SZ = GS_MAX_BS * GS_MAX_BS;
    __shared__ double A[SZ], G[SZ], T[SZ], D[SZ], E[SZ], Lm[SZ], X[SZ], Y[SZ];
    const int lane = threadIdx.x;
    const float *P = skew + (size_t)blockIdx.x * bs * bs;
    const float *Gp = dblocks + (size_t)blockIdx.x * bs * bs;
    for (int e = lane; e < bs * bs; e += 32) {
        const int r = e / bs, c = e % bs;
        A[e] = (double)(P[c * bs + r] - P[r * bs + c]);      // A^T = P^T - P
        G[e] = (double)Gp[e];
    }
    __syncwarp();
    const int s = gs_scaling(A, bs, lane);
    const double sc = ldexp(1.0, -s);
    for (int e = lane; e < bs * bs; e += 32) {
        A[e] *= sc;
        G[e] *= sc;
        const double id = (e / bs == e % bs) ? 1.0 : 0.0;
        T[e] = id; E[e] = id;
        D[e] = 0.0; Lm[e] = 0.0;
    }
    __syncwarp();
    for (int k = 1; k <= GS_TAYLOR; ++k) {
        gs_mm(X, A, D, bs, lane);        // A D_{k-1}
        gs_mm(Y, G, T, bs, lane);        // G T_{k-1}
        for (int e = lane; e < bs * bs; e += 32) {
            D[e] = (X[e] + Y[e]) / (double)k;
            Lm[e] += D[e];
        }
        __syncwarp();
        gs_mm(X, A, T, bs, lane);
        for (int e = lane; e < bs * bs; e += 32) {
            T[e] = X[e] / (double)k;
            E[e] += T[e];
        }
        __syncwarp();
    }
    for (int q = 0; q < s; ++q) {
        gs_mm(X, E, Lm, bs, lane);
        gs_mm(Y, Lm, E, bs, lane);
        for (int e = lane; e < bs * bs; e += 32) Lm[e] = X[e] + Y[e];
        __syncwarp();
        gs_mm(X, E, E, bs, lane);
        for (int e = lane; e < bs * bs; e += 32) E[e] = X[e];
        __syncwarp();
    }
    float *out = dskew + (size_t)blockIdx.x * bs * bs;
    for (int e = lane; e < bs * bs; e += 32) {
        const int r = e / bs, c = e % bs;
        out[e] = (float)(Lm[r * bs + c] - Lm[c * bs + r]);
    }
}

}  // namespace gr

using namespace gr;

static int gs_check(int32_t n_layers, int32_t n_sets, int32_t d, int32_t bs) {
    if (n_layers <= 0 || (n_sets != 1 && n_sets != 2) || d <= 0 || bs <= 0) return GR_ERR_INVALID;
    if (bs > GS_MAX_BS || d > 1024 || d % bs) return GR_ERR_UNSUPPORTED;
    return GR_OK;
}

extern "C" int gr_gs_compose(const float *skew, const int64_t *perm_conn, const int64_t *perm_local, int32_t n_layers,
                             int32_t n_sets, int32_t d, int32_t bs, float *blocks, float *m_out, void *stream) {
    if (!skew || !perm_local || !blocks || !m_out) return GR_ERR_INVALID;
    if (n_sets == 2 && !perm_conn) return GR_ERR_INVALID;
    const int rc = gs_check(n_layers, n_sets, d, bs);
    if (rc != GR_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_blocks = n_layers * n_sets * (d / bs);
    gs_block_exp_kernel<<<n_blocks, 32, 0, st>>>(skew, bs, blocks);
    GR_LAUNCH_CHECK();
    GsComposeArgs a;
    a.blocks = blocks; a.perm_c = reinterpret_cast<const long long *>(perm_conn);
    a.perm_g = reinterpret_cast<const long long *>(perm_local);
    a.d = d; a.bs = bs; a.n_sets = n_sets; a.m = m_out; a.dm = nullptr; a.dblocks = nullptr;
    gs_compose_kernel<<<dim3(d, n_layers), ((d + 31) / 32) * 32, 0, st>>>(a);
    GR_LAUNCH_CHECK();
    return GR_OK;
}

extern "C" int gr_gs_compose_bwd(const float *skew, const float *blocks, const int64_t *perm_conn,
                                 const int64_t *perm_local, int32_t n_layers, int32_t n_sets, int32_t d, int32_t bs,
                                 const float *dm, float *dblocks_ws, float *dskew, void *stream) {
    if (!skew || !blocks || !perm_local || !dm || !dblocks_ws || !dskew) return GR_ERR_INVALID;
    if (n_sets == 2 && !perm_conn) return GR_ERR_INVALID;
    const int rc = gs_check(n_layers, n_sets, d, bs);
    if (rc != GR_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GsComposeArgs a;
    a.blocks = blocks; a.perm_c = reinterpret_cast<const long long *>(perm_conn);
    a.perm_g = reinterpret_cast<const long long *>(perm_local);
    a.d = d; a.bs = bs; a.n_sets = n_sets; a.m = nullptr; a.dm = dm; a.dblocks = dblocks_ws;
    const long long total = (long long)n_layers * n_sets * (d / bs) * bs * bs;
    gs_compose_bwd_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, n_layers);
    GR_LAUNCH_CHECK();
    const int n_blocks = n_layers * n_sets * (d / bs);
    gs_block_exp_bwd_kernel<<<n_blocks, 32, 0, st>>>(skew, dblocks_ws, bs, dskew);
    GR_LAUNCH_CHECK();
    return GR_OK;
}
