// Backward kernels of the NGCF / GAT / Group-and-Shuffle layer epilogues (what autograd does under
// trainer.py:270 for ngcf.py:69-84, gat.py:97-149 and model.py:171-195).
//
//   gr_rowmap_bwd       forward: out = D * (alpha * act(z) + beta * R),  z = X1 Wa + ba + (X2*X3) Wb + bb
//                       (D = dropout keep/(1-p), re-derived from the seed).  Given g = dL/dout:
//                         dz  = alpha * D*g * act'(z)            (act' recovered from the stored output)
//                         dX1 = dz Wa^T;  dP = dz Wb^T;  dX2 = dP*X3;  dX3 = dP*X2;  dR = beta * D*g
//                         dWa = X1^T dz;  dWb = (X2*X3)^T dz;  dba = dbb = column sums of dz
//                       Three launches: rowmap_bwd_dx_kernel (HBM streaming, weights in smem, FFMA),
//                       rowmap_bwd_dw_kernel (persistent CTAs, 64x64 weight blocks in registers, per-CTA
//                       partials), reduce_partials_kernel (fixed order -> deterministic).
//   gr_gat_bwd          backward of gr_gat_aggregate (+ node scores): one warp per node walks its
//                       neighbours in the "row role" (j attends to k: D_j, ds_j; gat_bwd_row_kernel) and
//                       then in the "column role" (i attends to j: dt_j, dH_j; gat_bwd_col_kernel),
//                       recomputing the softmax weights from the stored (max, normaliser); no atomics,
//                       no edge-sized temporaries.
#include <math_constants.h>

#include "gr_common.cuh"

namespace gr {

constexpr int RB_ROWS = 64;

__device__ __forceinline__ float act_grad_from_output(float a, int act, float slope) {
    // a = act(z).  LeakyReLU: sign(a) == sign(z) (slope >= 0; z == 0 takes the negative branch like torch);
    // ELU: act'(z) = exp(z) = a + 1 for z <= 0.
    if (act == 1) return a > 0.f ? 1.f : slope;
    if (act == 2) return a > 0.f ? 1.f : a + 1.f;
    return 1.f;
}

struct RowMapBwdArgs {
    const float *g, *out, *resid, *x2, *x3, *wa, *wb;
    long long ldg, ldo, ldr, ld2, ld3;
    float *dz, *dx1, *dx2, *dx3, *dresid;
    long long lddz, ldd1, ldd2, ldd3, lddr;
    int n_rows, d_in, d_out, n_tiles;
    float alpha, beta, slope;
    int act;
    unsigned drop_thr;
    float drop_scale;
    unsigned long long drop_seed;
    const unsigned long long *drop_seed_dev;
};

// dz (written for the weight-gradient kernel) and the input gradients.  Same tiling as the forward
// kernel with the roles of d_in / d_out swapped: the k loop runs over d_out, weights are staged
// transposed ([d_out][d_in]).  Persistent over row tiles so the weights are staged once per CTA.
template <int NG>      // column groups of d_in per thread: 1 for d_in <= 64, 2 / 4 for <= 128 / 256 (see rowmap_kernel)
__global__ void __launch_bounds__(256) rowmap_bwd_dx_kernel(const RowMapBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d_in = a.d_in, d_out = a.d_out;
    const bool has_b = a.wb != nullptr;
    const bool need_dx = a.dx1 != nullptr;
    float *WaT = reinterpret_cast<float *>(smem_raw);                         // [d_out][d_in]
    float *WbT = WaT + (need_dx ? (size_t)d_in * d_out : 0);                  // [d_out][d_in]
    float *Zs = WbT + ((need_dx && has_b) ? (size_t)d_in * d_out : 0);        // [d_out][RB_ROWS]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    if (need_dx) {
        for (int i = tid; i < d_in * d_out; i += 256) {
            const int r = i / d_out, c = i % d_out;               // W[r][c], coalesced read
            WaT[(size_t)c * d_in + r] = __ldg(a.wa + i);
            if (has_b) WbT[(size_t)c * d_in + r] = __ldg(a.wb + i);
        }
    }
    const int o4 = d_out / 4;
    const int nc4 = d_in / 4;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int row0 = tile * RB_ROWS;
        __syncthreads();   // previous tile's readers of Zs are done; weights visible on the first pass
        for (int i = tid; i < RB_ROWS * o4; i += 256) {
            const int r = i % RB_ROWS, f = i / RB_ROWS;          // lane -> row: conflict-free transposed stores
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row = row0 + r;
            if (row < a.n_rows) {
                float4 gv = __ldg(reinterpret_cast<const float4 *>(a.g + (long long)row * a.ldg) + f);
                float gg[4] = {gv.x, gv.y, gv.z, gv.w};
                float keep[4] = {1.f, 1.f, 1.f, 1.f};
                if (a.drop_thr) {
                    const unsigned long long bits = drop_bits(a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0ULL),
                                                              (unsigned long long)row * o4 + f);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        keep[c] = drop_keep(bits, c, a.drop_thr) ? a.drop_scale : 0.f;
                        gg[c] *= keep[c];
                    }
                }
                if (a.dresid)
                    *reinterpret_cast<float4 *>(a.dresid + (long long)row * a.lddr + f * 4) =
                        make_float4(a.beta * gg[0], a.beta * gg[1], a.beta * gg[2], a.beta * gg[3]);
                float zz[4];
                if (a.act) {
                    const float4 ov = __ldg(reinterpret_cast<const float4 *>(a.out + (long long)row * a.ldo) + f);
                    float oo[4] = {ov.x, ov.y, ov.z, ov.w};
                    float rr[4] = {0.f, 0.f, 0.f, 0.f};
                    if (a.resid) {
                        const float4 rv = __ldg(reinterpret_cast<const float4 *>(a.resid + (long long)row * a.ldr) + f);
                        rr[0] = rv.x; rr[1] = rv.y; rr[2] = rv.z; rr[3] = rv.w;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        // act(z) = (out / D - beta R) / alpha ; where the element was dropped gg is 0 anyway
                        float av = keep[c] != 0.f ? oo[c] / keep[c] : 0.f;
                        av = (av - a.beta * rr[c]) / a.alpha;
                        zz[c] = a.alpha * gg[c] * act_grad_from_output(av, a.act, a.slope);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) zz[c] = a.alpha * gg[c];
                }
                z = make_float4(zz[0], zz[1], zz[2], zz[3]);
                if (a.dz) *reinterpret_cast<float4 *>(a.dz + (long long)row * a.lddz + f * 4) = z;
            }
            if (need_dx) {
                Zs[(4 * f + 0) * RB_ROWS + r] = z.x;
                Zs[(4 * f + 1) * RB_ROWS + r] = z.y;
                Zs[(4 * f + 2) * RB_ROWS + r] = z.z;
                Zs[(4 * f + 3) * RB_ROWS + r] = z.w;
            }
        }
        if (!need_dx) continue;
        __syncthreads();

        float accA[4][NG][4], accB[4][NG][4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int gq = 0; gq < NG; ++gq)
#pragma unroll
                for (int c = 0; c < 4; ++c) accA[m][gq][c] = accB[m][gq][c] = 0.f;
        for (int k = 0; k < d_out; ++k) {
            const float4 zv = *reinterpret_cast<const float4 *>(Zs + (size_t)k * RB_ROWS + ty * 4);
            const float zm[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) {
                const int cg = tx + 16 * gq;
                if (cg < nc4) {
                    const float4 w = *reinterpret_cast<const float4 *>(WaT + (size_t)k * d_in + cg * 4);
                    const float wc[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int c = 0; c < 4; ++c) accA[m][gq][c] = __fmaf_rn(zm[m], wc[c], accA[m][gq][c]);
                    if (has_b) {
                        const float4 w2 = *reinterpret_cast<const float4 *>(WbT + (size_t)k * d_in + cg * 4);
                        const float w2c[4] = {w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int c = 0; c < 4; ++c) accB[m][gq][c] = __fmaf_rn(zm[m], w2c[c], accB[m][gq][c]);
                    }
                }
            }
        }
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
            const int cg = tx + 16 * gq;
            if (cg >= nc4) continue;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int r = row0 + ty * 4 + m;
                if (r >= a.n_rows) continue;
                float4 d1 = make_float4(accA[m][gq][0], accA[m][gq][1], accA[m][gq][2], accA[m][gq][3]);
                if (has_b) {
                    const float4 p2 = __ldg(reinterpret_cast<const float4 *>(a.x2 + (long long)r * a.ld2) + cg);
                    const float4 p3 = __ldg(reinterpret_cast<const float4 *>(a.x3 + (long long)r * a.ld3) + cg);
                    const float4 dp = make_float4(accB[m][gq][0], accB[m][gq][1], accB[m][gq][2], accB[m][gq][3]);
                    const float4 d2 = make_float4(dp.x * p3.x, dp.y * p3.y, dp.z * p3.z, dp.w * p3.w);
                    const float4 d3 = make_float4(dp.x * p2.x, dp.y * p2.y, dp.z * p2.z, dp.w * p2.w);
                    if (a.dx2) *reinterpret_cast<float4 *>(a.dx2 + (long long)r * a.ldd2 + cg * 4) = d2;
                    if (a.dx3) *reinterpret_cast<float4 *>(a.dx3 + (long long)r * a.ldd3 + cg * 4) = d3;
                    else { d1.x += d3.x; d1.y += d3.y; d1.z += d3.z; d1.w += d3.w; }   // X3 is X1 (NGCF: n)
                }
                *reinterpret_cast<float4 *>(a.dx1 + (long long)r * a.ldd1 + cg * 4) = d1;
            }
        }
    }
}

struct RowMapDwArgs {
    const float *x1, *x2, *x3, *dz;
    long long ld1, ld2, ld3, lddz;
    float *partial;       // [P][total], total = nw * d_in * d_out + d_out
    int n_rows, d_in, d_out, n_tiles, nbj;
    int has_b;
};

// Weight gradients.  CTA (p, b) owns the 64x64 block b = (bi, bj) of the weight matrix and row tiles
// p, p+P, ...; thread = 4x4 sub-block held in registers across all its tiles.  X and dz tiles are
// staged row-major: the 16 threads that differ in j read 16 consecutive float4 (conflict-free), the
// two i values of a warp are broadcasts.
__global__ void __launch_bounds__(256) rowmap_bwd_dw_kernel(const RowMapDwArgs a) {
    __shared__ __align__(16) float Xs[RB_ROWS][64];
    __shared__ __align__(16) float Ys[RB_ROWS][64];
    __shared__ __align__(16) float Zs[RB_ROWS][64];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int bi = blockIdx.y / a.nbj, bj = blockIdx.y % a.nbj;
    const int i_base = bi * 64, j_base = bj * 64;
    const int wi = min(64, a.d_in - i_base), wj = min(64, a.d_out - j_base);   // multiples of 4
    float accA[4][4], accB[4][4], accS[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        accS[m] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) accA[m][c] = accB[m][c] = 0.f;
    }
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int row0 = tile * RB_ROWS;
        __syncthreads();
        for (int i = tid; i < RB_ROWS * 16; i += 256) {
            const int r = i >> 4, f = i & 15;
            const int row = row0 + r;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv, zv = xv;
            if (row < a.n_rows) {
                if (f * 4 < wi) {
                    xv = __ldg(reinterpret_cast<const float4 *>(a.x1 + (long long)row * a.ld1 + i_base) + f);
                    if (a.has_b) {
                        const float4 p = __ldg(reinterpret_cast<const float4 *>(a.x2 + (long long)row * a.ld2 + i_base) + f);
                        const float4 q = __ldg(reinterpret_cast<const float4 *>(a.x3 + (long long)row * a.ld3 + i_base) + f);
                        yv = make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
                    }
                }
                if (f * 4 < wj) zv = __ldg(reinterpret_cast<const float4 *>(a.dz + (long long)row * a.lddz + j_base) + f);
            }
            *reinterpret_cast<float4 *>(&Xs[r][f * 4]) = xv;
            if (a.has_b) *reinterpret_cast<float4 *>(&Ys[r][f * 4]) = yv;
            *reinterpret_cast<float4 *>(&Zs[r][f * 4]) = zv;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < RB_ROWS; ++r) {
            const float4 zv = *reinterpret_cast<const float4 *>(&Zs[r][tx * 4]);
            const float4 xv = *reinterpret_cast<const float4 *>(&Xs[r][ty * 4]);
            const float zc[4] = {zv.x, zv.y, zv.z, zv.w};
            const float xm[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c) accA[m][c] = __fmaf_rn(xm[m], zc[c], accA[m][c]);
            if (a.has_b) {
                const float4 yv = *reinterpret_cast<const float4 *>(&Ys[r][ty * 4]);
                const float ym[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int c = 0; c < 4; ++c) accB[m][c] = __fmaf_rn(ym[m], zc[c], accB[m][c]);
            }
            if (ty == 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) accS[c] += zc[c];
            }
        }
    }
    const int nw = a.has_b ? 2 : 1;
    const long long total = (long long)nw * a.d_in * a.d_out + a.d_out;
    float *part = a.partial + (long long)blockIdx.x * total;
    if (tx * 4 < wj) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int i = i_base + ty * 4 + m;
            if (ty * 4 + m < wi) {
                *reinterpret_cast<float4 *>(part + (long long)i * a.d_out + j_base + tx * 4) =
                    make_float4(accA[m][0], accA[m][1], accA[m][2], accA[m][3]);
                if (a.has_b)
                    *reinterpret_cast<float4 *>(part + (long long)(a.d_in + i) * a.d_out + j_base + tx * 4) =
                        make_float4(accB[m][0], accB[m][1], accB[m][2], accB[m][3]);
            }
        }
        if (ty == 0 && bi == 0)
            *reinterpret_cast<float4 *>(part + (long long)nw * a.d_in * a.d_out + j_base + tx * 4) =
                make_float4(accS[0], accS[1], accS[2], accS[3]);
    }
}

// out[e] = sum_p partial[p][e] (fixed order -> deterministic).  Block = 32 elements x 32 warps: warp w adds the
// rows p = w, w + 32, ... of its 32 consecutive elements (coalesced 128-byte loads, 4 independent loads in
// flight), the 32 warp sums are then added in warp order by warp 0.  (The first version ran one thread per element
// over all rows serially: 67 us for 1 184 x 512 floats.)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float *partial, int n_part, long long total,
                                                               float *out) {
    __shared__ float part[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long e = (long long)blockIdx.x * 32 + lane;
    float s = 0.f;
    if (e < total) {
        int p = warp;
        for (; p + 96 < n_part; p += 128) {
            const float v0 = partial[(long long)p * total + e], v1 = partial[(long long)(p + 32) * total + e];
            const float v2 = partial[(long long)(p + 64) * total + e], v3 = partial[(long long)(p + 96) * total + e];
            s = ((s + v0) + v1) + v2 + v3;
        }
        for (; p < n_part; p += 32) s += partial[(long long)p * total + e];
    }
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && e < total) {
        float t = part[0][lane];
#pragma unroll
        for (int w = 1; w < 32; ++w) t += part[w][lane];
        out[e] = t;
    }
}

}  // namespace gr

namespace gr {

int persistent_grid(int work_items, int per_sm) {
    const int cap = sm_count() * per_sm;
    return work_items < cap ? (work_items > 0 ? work_items : 1) : cap;
}

int launch_reduce_partials(const float *partial, int n_part, long long total, float *out, cudaStream_t st) {
    reduce_partials_kernel<<<(unsigned)((total + 31) / 32), 1024, 0, st>>>(partial, n_part, total, out);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace gr

using namespace gr;

static int dw_partials(int64_t n_rows, int32_t d_in, int32_t d_out) {
    const int n_tiles = (int)((n_rows + RB_ROWS - 1) / RB_ROWS);
    const int nb = ((d_in + 63) / 64) * ((d_out + 63) / 64);
    int p = persistent_grid(n_tiles, 2);
    p = p / nb;
    return p < 1 ? 1 : p;
}

extern "C" size_t gr_rowmap_bwd_workspace_bytes(int64_t n_rows, int32_t d_in, int32_t d_out, int32_t has_b) {
    if (n_rows <= 0 || d_in <= 0 || d_out <= 0) return 256;
    const size_t total = (size_t)(has_b ? 2 : 1) * d_in * d_out + d_out;
    const size_t dz = ((size_t)n_rows * d_out * 4 + 255) & ~(size_t)255;
    return dz + (size_t)dw_partials(n_rows, d_in, d_out) * total * 4 + 256;
}

extern "C" int gr_rowmap_bwd(const float *g, int64_t ldg, const float *out, int64_t ldo, const float *x1, int64_t ld1,
                             const float *wa, const float *x2, int64_t ld2, const float *x3, int64_t ld3,
                             const float *wb, const float *resid, int64_t ldr, float alpha, float beta, int32_t act,
                             float slope, int64_t n_rows, int32_t d_in, int32_t d_out, float drop_p, uint64_t drop_seed,
                             const uint64_t *drop_seed_dev, float *dx1, int64_t ldd1, float *dx2, int64_t ldd2, float *dx3, int64_t ldd3,
                             float *dresid, int64_t lddr, float *dw, void *workspace, size_t workspace_bytes,
                             void *stream) {
    if (!g || !x1 || !wa || n_rows < 0) return GR_ERR_INVALID;
    if (wb && (!x2 || !x3)) return GR_ERR_INVALID;
    if (act < 0 || act > 2 || (act && !out)) return GR_ERR_INVALID;
    if (act && alpha == 0.f) return GR_ERR_INVALID;
    if (!(drop_p >= 0.f) || drop_p >= 1.f) return GR_ERR_INVALID;
    if (drop_p > 0.f && act && !out) return GR_ERR_INVALID;
    if (d_in <= 0 || d_out <= 0 || (d_in & 3) || (d_out & 3) || d_in > 256 || d_out > 256) return GR_ERR_UNSUPPORTED;
    if (n_rows > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if ((ldg & 3) || ldg < d_out || (ld1 & 3) || ld1 < d_in) return GR_ERR_INVALID;
    if (act && ((ldo & 3) || ldo < d_out)) return GR_ERR_INVALID;
    if (wb && ((ld2 & 3) || (ld3 & 3) || ld2 < d_in || ld3 < d_in)) return GR_ERR_INVALID;
    if (dx1 && ((ldd1 & 3) || ldd1 < d_in)) return GR_ERR_INVALID;
    if ((dx2 || dx3) && !dx1) return GR_ERR_INVALID;
    if (dx2 && ((ldd2 & 3) || ldd2 < d_in)) return GR_ERR_INVALID;
    if (dx3 && ((ldd3 & 3) || ldd3 < d_in)) return GR_ERR_INVALID;
    if (dresid && ((lddr & 3) || lddr < d_out)) return GR_ERR_INVALID;
    if (resid && ((ldr & 3) || ldr < d_out)) return GR_ERR_INVALID;
    if (!aligned16(g) || !aligned16(out) || !aligned16(x1) || !aligned16(x2) || !aligned16(x3) || !aligned16(wa) ||
        !aligned16(wb) || !aligned16(resid) || !aligned16(dx1) || !aligned16(dx2) || !aligned16(dx3) ||
        !aligned16(dresid) || !aligned16(dw) || !aligned16(workspace))
        return GR_ERR_INVALID;
    if (n_rows == 0) {
        if (dw) {
            const size_t total = (size_t)(wb ? 2 : 1) * d_in * d_out + d_out;
            GR_CUDA_CHECK(cudaMemsetAsync(dw, 0, total * 4, static_cast<cudaStream_t>(stream)));
        }
        return GR_OK;
    }
    if (workspace_bytes < gr_rowmap_bwd_workspace_bytes(n_rows, d_in, d_out, wb != nullptr)) return GR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_tiles = (int)((n_rows + RB_ROWS - 1) / RB_ROWS);
    float *dz = static_cast<float *>(workspace);
    float *partial = reinterpret_cast<float *>(static_cast<char *>(workspace) +
                                               (((size_t)n_rows * d_out * 4 + 255) & ~(size_t)255));
    const bool trivial_dz = !act && alpha == 1.f && drop_p == 0.f;   // dz == g: skip the copy
    const unsigned thr = drop_threshold(drop_p);
    if (dx1 || dresid || !trivial_dz) {
        RowMapBwdArgs a;
        a.g = g; a.out = out; a.resid = resid; a.x2 = x2; a.x3 = x3; a.wa = wa; a.wb = wb;
        a.ldg = ldg; a.ldo = ldo; a.ldr = ldr; a.ld2 = ld2; a.ld3 = ld3;
        a.dz = (dw && !trivial_dz) ? dz : nullptr; a.dx1 = dx1; a.dx2 = dx2; a.dx3 = dx3; a.dresid = dresid;
        a.lddz = d_out; a.ldd1 = ldd1; a.ldd2 = ldd2; a.ldd3 = ldd3; a.lddr = lddr;
        a.n_rows = (int)n_rows; a.d_in = d_in; a.d_out = d_out; a.n_tiles = n_tiles;
        a.alpha = alpha; a.beta = beta; a.slope = slope; a.act = act;
        a.drop_thr = thr; a.drop_scale = drop_scale_of(thr); a.drop_seed = drop_seed;
        a.drop_seed_dev = reinterpret_cast<const unsigned long long *>(drop_seed_dev);
        const size_t smem = dx1 ? ((size_t)d_in * d_out * (wb ? 2 : 1) + (size_t)d_out * RB_ROWS) * 4 : 0;
        if (smem > 227 * 1024) return GR_ERR_UNSUPPORTED;
        const int grid = persistent_grid(n_tiles, smem > 100 * 1024 ? 1 : (smem > 70 * 1024 ? 2 : (smem > 50 * 1024 ? 3 : 4)));
        const int smem_attr = (int)(smem > 0 ? smem : 16);
        if (d_in <= 64) {
            GR_CUDA_CHECK(cudaFuncSetAttribute(rowmap_bwd_dx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
            rowmap_bwd_dx_kernel<1><<<grid, 256, smem, st>>>(a);
        } else if (d_in <= 128) {
            GR_CUDA_CHECK(cudaFuncSetAttribute(rowmap_bwd_dx_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
            rowmap_bwd_dx_kernel<2><<<grid, 256, smem, st>>>(a);
        } else {
            GR_CUDA_CHECK(cudaFuncSetAttribute(rowmap_bwd_dx_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
            rowmap_bwd_dx_kernel<4><<<grid, 256, smem, st>>>(a);
        }
        GR_LAUNCH_CHECK();
    }
    if (dw) {
        RowMapDwArgs w;
        w.x1 = x1; w.x2 = x2; w.x3 = x3; w.dz = trivial_dz ? g : dz;
        w.ld1 = ld1; w.ld2 = ld2; w.ld3 = ld3; w.lddz = trivial_dz ? ldg : d_out;
        w.partial = partial; w.n_rows = (int)n_rows; w.d_in = d_in; w.d_out = d_out; w.n_tiles = n_tiles;
        w.nbj = (d_out + 63) / 64; w.has_b = wb != nullptr;
        const int nb = ((d_in + 63) / 64) * w.nbj;
        const int p = dw_partials(n_rows, d_in, d_out);
        const long long total = (long long)(wb ? 2 : 1) * d_in * d_out + d_out;
        rowmap_bwd_dw_kernel<<<dim3(p, nb), 256, 0, st>>>(w);
        GR_LAUNCH_CHECK();
        if (launch_reduce_partials(partial, p, total, dw, st)) {
            set_last_cuda_error(cudaPeekAtLastError());
            return GR_ERR_CUDA;
        }
    }
    return GR_OK;
}

