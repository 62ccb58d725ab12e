/*
 * gr_b200.h — C ABI of libgr_b200.so: the B200 (sm_100a) hot path of
 * timur1arkhipov/gnn-recommendations.
 *
 * The reference is pure Python: its "operator interface" for this path is a set of
 * call sites into torch / scipy.  Each entry point below names the reference call
 * site (file:line under /root/reference/gnn-recommendations/) it replaces.  The
 * Python classes in gnn-recommendations_b200/ bind these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - functions return GR_OK (0) or a negative GR_ERR_* code, never throw, never
 *     allocate device memory (callers pass outputs and workspace), never synchronise
 *     the stream;
 *   - all matrices are row-major fp32; `ld*` are leading dimensions in floats and must
 *     be multiples of 4 (16-byte rows); base pointers must be 16-byte aligned;
 *   - integer outputs (CSR structure, negatives, top-K ids) are bit-exact with the
 *     reference; fp32 outputs follow the reference's summation order where stated.
 */
#ifndef GR_B200_H
#define GR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR_OK 0
#define GR_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, misalignment) */
#define GR_ERR_UNSUPPORTED (-2) /* feature dimension / option not compiled in                */
#define GR_ERR_CUDA (-3)        /* a CUDA runtime call failed (see gr_last_cuda_error)        */
#define GR_ERR_WORKSPACE (-4)   /* workspace too small                                       */
#define GR_ERR_OVERFLOW (-5)    /* nnz or N does not fit the 32-bit CSR                      */

#define GR_SCALE_NONE 0
#define GR_SCALE_MUL 1 /* out = (addend + A x) * scale  — torch.mean multiplies by 1/count  */
#define GR_SCALE_DIV 2 /* out = (addend + A x) / scale                                        */

const char *gr_version(void);
const char *gr_error_string(int code);
/* cudaGetErrorString of the last failing CUDA call made by this thread inside the library. */
const char *gr_last_cuda_error(void);
/* sm count and compute capability of the current device. */
int gr_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host);

/* ------------------------------------------------------------------------------------------
 * Graph builder
 * ------------------------------------------------------------------------------------------ */

/* Replaces the implicit COO->CSR conversion inside torch.sparse.mm for the adjacency the
 * reference hands to every model (src/data/graph_builder.py:147-174 produces int64 COO,
 * row-major sorted; src/models/baselines/lightgcn.py:88 consumes it).  Storage order within
 * a row is preserved (it is the summation order).  `status` (device int32[1], zeroed by the
 * caller) receives a bit mask: 1 = rows not non-decreasing, 2 = index out of range. */
int gr_coo_sorted_to_csr(const int64_t *rows, const int64_t *cols, const float *vals, int64_t nnz,
                         int64_t n_rows, int64_t n_cols, int32_t *indptr /* n_rows+1 */,
                         int32_t *indices /* nnz */, float *out_vals /* nnz, may alias vals */,
                         int32_t *status, void *stream);

/* Replaces build_bipartite_graph + normalize_adjacency_matrix (src/data/graph_builder.py:16-80,
 * 83-144): (user,item) int64 pairs -> canonical CSR of [[0,R],[R^T,0]] (+ I when self_loop),
 * rows/columns numbered users first.  Two phases, because the values come from a HOST look-up
 * table of numpy's np.power(float32(deg), -0.5) that can only be sized once the maximum
 * degree is known (numpy's f32 pow is not correctly rounded, so the device cannot recompute it):
 *
 *   gr_build_csr_pattern: keys (row << b | col) -> stable LSD radix sort -> duplicates merged
 *       (scipy's tocsr SUMS them: `mult` holds the multiplicity) -> indptr / indices / deg
 *       (row sums, integer) / nnz_out (device int64[1]) / max_deg_out (device int32[1]).
 *       `indices` and `mult` need room for 2*n_pairs (+ N with self loops) entries.
 *       status bit 2: id out of range.
 *   gr_csr_normalize: vals[k] = fl(fl(lut[deg[r]] * mult[k]) * lut[deg[c]])   mode 0 'symmetric'
 *                              = fl(lut[deg[r]] * mult[k])                     mode 1 'row'
 *                              = mult[k]                                       mode 2 'none'
 *       (the left-to-right order of `d_mat @ adj @ d_mat`, graph_builder.py:126); lut[0] is the
 *       value for the clamped degree max(0,1).  status bit 4: degree outside the table. */
size_t gr_build_csr_workspace_bytes(int64_t n_pairs, int64_t n_users, int64_t n_items, int32_t self_loop);
int gr_build_csr_pattern(const int64_t *user, const int64_t *item, int64_t n_pairs, int64_t n_users,
                         int64_t n_items, int32_t self_loop, int32_t *indptr /* N+1 */, int32_t *indices,
                         float *mult, int32_t *deg /* N */, int64_t *nnz_out, int32_t *max_deg_out,
                         int32_t *status, void *workspace, size_t workspace_bytes, void *stream);
int gr_csr_normalize(const int32_t *indptr, const int32_t *indices, const float *mult, const int32_t *deg,
                     const float *lut, int64_t lut_len, int64_t n_rows, int64_t nnz, int32_t mode, float *vals,
                     int32_t *status, void *stream);

/* Row schedule for gr_spmm_csr_f32: `row_order` = rows sorted by descending length
 * (longest-processing-time-first), `n_long_out` (device int32[1]) = how many of them have
 * >= long_threshold entries and go to the CTA-cooperative path.  `keys` is caller scratch of
 * n_rows int64.  Ties keep ascending row id. */
size_t gr_row_schedule_workspace_bytes(int64_t n_rows);
int gr_row_schedule(const int32_t *indptr, int64_t n_rows, int32_t long_threshold, int32_t *row_order,
                    int32_t *n_long_out, void *workspace, size_t workspace_bytes, void *stream);

/* Row groups for the streaming short-row kernel: group g = rows whose first entry lies in
 * [g*group_nnz, (g+1)*group_nnz); group_ptr has n_groups + 1 entries, n_groups must be
 * nnz / group_nnz + 1. */
int gr_row_groups(const int32_t *indptr, int64_t n_rows, int32_t group_nnz, int32_t n_groups, int32_t *group_ptr,
                  void *stream);

/* ------------------------------------------------------------------------------------------
 * Propagation
 * ------------------------------------------------------------------------------------------ */

/* Replaces torch.sparse.mm(adj, x) (lightgcn.py:88, ngcf.py:70, orthogonal_bundle/model.py:172,
 * kgtore.py:310) and, through the epilogue, torch.stack/mean over layers (lightgcn.py:94-95).
 *
 *   t[r,:]   = fma-chain over the row's entries in storage order, from a zero accumulator
 *              (bit-identical to the CPU torch.sparse.mm the reference runs)
 *   y[r,:]   = t[r,:]                                   if y   != NULL
 *   out[r,:] = scale_op(addend ? addend[r,:] + t : t)   if out != NULL
 *
 * `row_order` may be NULL (natural order, n_long ignored).  Rows row_order[0..n_long) (those
 * with >= long_threshold entries) are processed one CTA per row with the gathered x rows staged
 * through shared memory by cp.async.  The remaining rows: with `group_ptr` (gr_row_groups) one
 * group of consecutive rows per (sub-)warp as a continuous gather stream; without it one row per
 * (sub-)warp in row_order.  d in {32,64,128,256}.  n_rows may be a row block of a larger matrix
 * (column ids index x, not y).
 *
 * Optional split of extreme rows (`long_items` != NULL): the long-row kernel then works through
 * `long_items` = int32[4][n_long_items] (row, offset into the row, length, partial slot or -1)
 * instead of whole rows; a row cut into segments has each segment's chain stored in
 * part_buf[slot][d] and the partials added in segment order by a combine pass driven by
 * `split_rows` = int32[3][n_split] (row, first slot, number of slots).  A split row is no longer the
 * single storage-order chain (deterministic, ~1e-7 relative); the host layer only splits rows
 * above 131 072 entries, which none of the reference's dataset shapes has.
 *
 * Re-entrancy: the library keeps no device-side state.  `sched_ws` = two uint32 words of DEVICE memory, zero
 * on entry and zero again when the launch has finished (ticket / finished-CTA counters of the persistent
 * long-row kernel); required when n_long > 0.  Launches that may run concurrently (different streams) must
 * be given different words; launches on one stream may share them. */
int gr_spmm_csr_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                    const int32_t *row_order, int32_t n_long, const int32_t *long_items,
                    int32_t n_long_items, const int32_t *split_rows, int32_t n_split, float *part_buf,
                    const int32_t *group_ptr, int32_t n_groups,
                    int32_t long_threshold, int64_t n_rows, int32_t d, const float *x,
                    int64_t ldx, float *y, int64_t ldy, const float *addend, int64_t lda, float *out,
                    int64_t ldo, float scale, int32_t scale_mode, float *const *peer_y_host, int32_t n_peers,
                    int32_t peer_multicast, int64_t peer_row_offset, int32_t peer_route_block, uint32_t *sched_ws,
                    void *stream);

/* SpMM with a dense-map epilogue: one Group-and-Shuffle layer in one pass (replaces, for
 * orthogonal_bundle/model.py:171-195, `torch.sparse.mm` + the `@ W_conn`, `@ W_orth`, `[:, perm]` products + the
 * residual; W_conn W_orth[:, perm] is pre-composed into `map` by gr_gs_compose):
 *     t = A x (the same storage-order fmaf chain as gr_spmm_csr_f32)
 *     out[r,:] = alpha * (t[r,:] @ map) + beta * addend[r,:]        y[r,:] = alpha * t[r,:]   (y optional)
 * `map` = d x d row-major fp32 in device memory (map_transposed != 0: map^T is applied — the backward of the
 * layer is the same call on the transposed adjacency with map^T); beta is multiplied by *beta_dev when that
 * device scalar is given (a softmax layer weight that lives on the device).  The finished row never leaves the
 * SM between the sparse product and the map: no second launch, no N x d round trip.  d <= 64, streaming
 * schedule (group_ptr) required, single GPU; otherwise GR_ERR_UNSUPPORTED (the caller then runs
 * gr_spmm_csr_f32 + gr_rowmap_f32).  Schedule / sched_ws arguments as for gr_spmm_csr_f32. */
int gr_spmm_csr_map_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                        const int32_t *row_order, int32_t n_long, const int32_t *long_items,
                        int32_t n_long_items, const int32_t *split_rows, int32_t n_split, float *part_buf,
                        const int32_t *group_ptr, int32_t n_groups, int32_t long_threshold, int64_t n_rows,
                        int32_t d, const float *x, int64_t ldx, float *y, int64_t ldy, const float *addend,
                        int64_t lda, float *out, int64_t ldo, const float *map, int32_t map_transposed,
                        float alpha, float beta, const float *beta_dev, uint32_t *sched_ws, void *stream);

/* Fused compute + all-gather for the row-partitioned multi-GPU propagation (no reference
 * counterpart).  `peer_y_host` (HOST array of n_peers <= 8 DEVICE pointers, one per rank of the
 * box including this one, e.g. torch symmetric-memory buffer_ptrs) names every rank's gathered
 * [G*H, ldy] layer buffer: the SpMM epilogue stores t[r,:] to row (peer_row_offset + r) of each of
 * them with plain P2P stores over NVLink, so the exchange of a layer overlaps its computation row
 * by row and no separate all-gather runs.  The caller orders layers with a cross-rank barrier.
 * With peer_multicast = 1, peer_y_host[0] (n_peers = 1) is an NVSwitch MULTICAST address of the
 * buffer (symmetric memory multicast_ptr): each row is written once with multimem.st and replicated
 * to every rank inside the switch (NVLS), so a rank's NVLink egress is 1x the layer instead of (G-1)x.
 * gr_peer_scatter_rows does the same store fan-out for an existing matrix (the layer-0 exchange).
 * Routed mode (peer_route_block > 0, gr_spmm_csr_f32 only): row r is stored to ONE peer, g = r / peer_route_block,
 * as row peer_row_offset + r % peer_route_block of its buffer — the user-owner propagation pushes every partial
 * item row straight to the rank that owns the item block, into the slot of the sending rank. */
int gr_peer_scatter_rows(const float *src, int64_t lds, int64_t n_rows, int32_t d, float *const *peer_dst_host,
                         int32_t n_peers, int32_t peer_multicast, int64_t ldd, int64_t peer_row_offset,
                         void *stream);

/* Reduce + broadcast of partial item rows over NVLink peer memory: the exchange of the user-owner ("1.5-D")
 * multi-GPU propagation (SURVEY.md §8e "bipartite refinement"; no reference counterpart).  Every rank holds
 * partial sums of ALL item rows computed from the users it owns and pushes them (routed gr_spmm_csr_f32 epilogue)
 * to the rank owning the item block, one slot per sending rank.  This call forms, for the n_rows rows of the
 * block,  v[j] = ((P_0[row0 + j] + P_1[row0 + j]) + ...)  over the n_src source buffers (HOST array of DEVICE
 * pointers: the G slots of the local staging buffer, or peer-mapped buffers; fixed order = deterministic), stores
 * v into row dst_row0 + j of the n_dst item tables (peer stores over NVLink; n_dst = 0 on the last layer),
 * optionally into `own` [n_rows, ldw], and folds it into the running layer sum
 *   out[j] = scale_op(addend[j] + v[j])   (same scale modes as gr_spmm_csr_f32). */
int gr_reduce_bcast_rows(const float *const *src_host, int32_t n_src, int64_t ld_src, float *const *dst_host,
                         int32_t n_dst, int64_t ld_dst, int64_t row0, int64_t dst_row0, int64_t n_rows, int32_t d,
                         const float *addend, int64_t lda, float *out, int64_t ldo, float *own, int64_t ldw,
                         float scale, int32_t scale_mode, void *stream);

/* cudaMemcpyAsync between device buffers of this box (local or peer-mapped): the copy engines move the bytes over
 * NVLink, no SM is involved.  Used by the user-owner propagation for the partial-block pushes and the reduced-block
 * broadcasts, which run beside the SpMM kernels. */
int gr_peer_copy_async(void *dst, const void *src, size_t bytes, void *stream);
/* The same transfers by a small SM kernel: up to 8 (dst, src, bytes) segments in one launch of `ctas` CTAs (a few
 * dozen; ~30 registers per thread), plain peer stores.  bytes multiples of 16, pointers 16-byte aligned. */
int gr_peer_copy_multi(void *const *dst_host, const void *const *src_host, const size_t *bytes_host, int32_t n_seg,
                       int32_t ctas, void *stream);

/* Per-row dense epilogue shared by the NGCF, Group-and-Shuffle and GAT layers:
 *     out = alpha * act( X1 Wa + bias_a  +  (X2 * X3) Wb + bias_b ) + beta * R
 * Wa, Wb: [d_in, d_out] row-major (i.e. nn.Linear.weight TRANSPOSED); the (X2*X3) term, the
 * biases and R are optional (NULL).  act: 0 none, 1 LeakyReLU(slope), 2 ELU.
 *   NGCF  (src/models/baselines/ngcf.py:77-84): X1 = X3 = A x, X2 = x, act = LeakyReLU(0.2)
 *   G&S   (src/models/orthogonal_bundle/model.py:176-195, group_shuffle_layer.py:88-94):
 *         X1 = A x, Wa = W_conn W_orth[:,perm], alpha = 1 - a, beta = a, R = x0
 *   GAT   (src/models/baselines/gat.py:99): X1 = x, Wa = [W_0^T | ... | W_{H-1}^T]
 * d_in, d_out multiples of 4, <= 256. */
int gr_rowmap_f32(const float *x1, int64_t ld1, const float *wa, const float *bias_a, const float *x2, int64_t ld2,
                  const float *x3, int64_t ld3, const float *wb, const float *bias_b, const float *resid,
                  int64_t ldr, float alpha, float beta, int32_t act, float slope, int64_t n_rows, int32_t d_in,
                  int32_t d_out, float drop_p, uint64_t drop_seed, const uint64_t *drop_seed_dev, float *out, int64_t ldo,
                  void *stream);

/* Backward of gr_rowmap_f32 — replaces what autograd runs under loss.backward() (src/training/trainer.py:270)
 * for NGCFLayer.forward (ngcf.py:69-84), the Group-and-Shuffle maps (model.py:176-195) and the GAT head
 * projections (gat.py:99).  With drop_p > 0 the forward output is  D * (alpha*act(z) + beta*R),  D = keep/(1-p)
 * re-derived from drop_seed (the layer-output nn.Dropout of ngcf.py:86 / model.py:198-199 fused).
 * drop_seed_dev (optional DEVICE uint64, here and in gr_rowmap_f32 / gr_gat_aggregate / gr_gat_bwd): added to
 * drop_seed at run time, so a training step captured once in a CUDA graph draws a new mask on every replay.
 * Given g = dL/dout [n, d_out]:
 *     dz = alpha * D*g * act'(z)   (act' recovered from `out`, required when act != 0)
 *     dx1 = dz Wa^T;  dP = dz Wb^T;  dx2 = dP*X3;  dx3 = dP*X2;  dresid = beta * D*g
 *     dw  = [ X1^T dz  |  (X2*X3)^T dz (if wb)  |  column sums of dz ]   (floats: nw*d_in*d_out + d_out)
 * Any of dx1/dx2/dx3/dresid/dw may be NULL (skipped; dx2/dx3 need dx1).  dx3 == NULL with wb != NULL means
 * "X3 is X1" (NGCF: both are A x): dx3 is added into dx1.  Deterministic: per-CTA partial sums reduced in a
 * fixed order.  workspace: gr_rowmap_bwd_workspace_bytes(n_rows, d_in, d_out, wb != NULL). */
size_t gr_rowmap_bwd_workspace_bytes(int64_t n_rows, int32_t d_in, int32_t d_out, int32_t has_b);
int gr_rowmap_bwd(const float *g, int64_t ldg, const float *out, int64_t ldo, const float *x1, int64_t ld1,
                  const float *wa, const float *x2, int64_t ld2, const float *x3, int64_t ld3, const float *wb,
                  const float *resid, int64_t ldr, float alpha, float beta, int32_t act, float slope,
                  int64_t n_rows, int32_t d_in, int32_t d_out, float drop_p, uint64_t drop_seed,
                  const uint64_t *drop_seed_dev, float *dx1,
                  int64_t ldd1, float *dx2, int64_t ldd2, float *dx3, int64_t ldd3, float *dresid, int64_t lddr,
                  float *dw, void *workspace, size_t workspace_bytes, void *stream);

/* Combination of the L+1 layer outputs:  out = sum_l w_l x_l  with the weights in device memory (Group-and-Shuffle:
 * softmax(layer_weights), src/models/orthogonal_bundle/model.py:204-207; products rounded separately and added left
 * to right like the reference's python sum()), or with w_dev = NULL  out = (x_0 + x_1 + ...) / n_in  (GAT:
 * torch.mean(torch.stack(..)), src/models/baselines/gat.py:287-288).  x_host / ld_host: HOST arrays of n_in <= 8
 * device pointers / leading dimensions.  gr_layer_combine_dw: dw[l] = <g, x_l> for the weighted form (dw: device
 * float[8]; deterministic two-stage reduction); dx_l = w_l g is gr_layer_combine with one input. */
int gr_layer_combine(const float *const *x_host, const int64_t *ld_host, int32_t n_in, const float *w_dev,
                     int64_t n_rows, int32_t d, float *out, int64_t ldo, void *stream);
size_t gr_layer_combine_bwd_workspace_bytes(void);
int gr_layer_combine_dw(const float *const *x_host, const int64_t *ld_host, int32_t n_in, const float *g, int64_t ldg,
                        int64_t n_rows, int32_t d, float *dw, void *workspace, size_t workspace_bytes, void *stream);

/* Group-and-Shuffle orthogonal maps of OrthogonalBundleGNN, composed per layer on the device:
 *     M_l = blockdiag(exp(P_k - P_k^T))[:, perm_conn_l]  @  blockdiag(exp(Q_k - Q_k^T))[:, perm_local_l]
 * replacing BundleConnectionLayer.forward (src/models/orthogonal_bundle/bundle_layer.py:56-73),
 * GroupShuffleLayer._build_orthogonal_matrix / forward (group_shuffle_layer.py:88-129) and the two matmuls of
 * model.py:176 — 16 torch.matrix_exp calls, 2 block_diag, 2 index ops and a 64x64 GEMM per layer there.
 *   skew   [n_layers, n_sets, d/bs, bs, bs] f32: the skew_params; n_sets = 2 (set 0 connection P, set 1 local Q)
 *          or 1 (use_parallel_transport = False: M_l = blockdiag(exp(Q))[:, perm_local_l]; perm_conn = NULL)
 *   perm_* [n_layers, d] int64 (the torch.randperm buffers);  bs <= 16, d % bs == 0
 *   blocks [same shape as skew] out: the block exponentials (kept for the backward pass);  m_out [n_layers, d, d]
 * Matrix exponentials: scaling-and-squaring, degree-12 Taylor in double (then rounded to f32).
 * gr_gs_compose_bwd: dm = dL/dM [n_layers, d, d] -> dskew (same shape as skew) through the adjoint Frechet
 * derivative of exp; dblocks_ws: scratch of the same size as skew. */
int gr_gs_compose(const float *skew, const int64_t *perm_conn, const int64_t *perm_local, int32_t n_layers,
                  int32_t n_sets, int32_t d, int32_t bs, float *blocks, float *m_out, void *stream);
int gr_gs_compose_bwd(const float *skew, const float *blocks, const int64_t *perm_conn, const int64_t *perm_local,
                      int32_t n_layers, int32_t n_sets, int32_t d, int32_t bs, const float *dm, float *dblocks_ws,
                      float *dskew, void *stream);

/* GAT (src/models/baselines/gat.py:76-151) over the CSR PATTERN of the adjacency (its values are
 * ignored, gat.py:120-127) instead of the reference's dense N x N temporaries.
 *   gr_gat_node_scores: s[i,h] = <H_h[i], a_self_h>, t[i,h] = <H_h[i], a_neigh_h>      (gat.py:106-109)
 *   gr_gat_aggregate:   out_i = sum_j softmax_j(LeakyReLU_slope(s[i,h] + t[j,h])) H_h[j] over the
 *       row's neighbours (online softmax, one pass); heads concatenated (mean_heads = 0, width
 *       heads*dh) or averaged (mean_heads = 1, width dh) (gat.py:144-147); elu = 1 applies the ELU of
 *       GAT.forward (gat.py:283).  A row without neighbours yields NaN like the reference.
 *       m_out / z_out (optional, [n_rows, heads]): softmax max / normaliser for the backward pass.
 *       drop_p > 0: the reference's dropout on the softmaxed attention weights (gat.py:138) — every edge
 *       weight is kept with probability 1-p and scaled by 1/(1-p) (the normaliser still sums all edges);
 *       the mask is a counter-based hash of (drop_seed, row, column, head), re-derived by the backward pass.
 * H: [N, heads*dh], heads*dh <= 256, dh % 4 == 0.
 *
 * Hot rows.  One warp owns one work item: a whole row of at most seg_len entries, or one seg_len-entry
 * segment of a longer row; segment items leave partial results (online-softmax triples forward, partial sums
 * backward) that a combine kernel merges per long row in segment order.  The table (HOST struct of DEVICE
 * arrays, NULL = every row is one item) lists, for the rows with more than seg_len entries:
 *   seg_row / seg_begin / seg_end [n_seg]: row id and entry range [begin, end) of every segment, grouped by row;
 *   long_rows [n_long], long_seg_ptr [n_long + 1]: the long rows and their segment ranges. */
typedef struct gr_gat_segments {
    int32_t seg_len, n_seg, n_long;
    const int32_t *seg_row, *seg_begin, *seg_end, *long_rows, *long_seg_ptr;
} gr_gat_segments;

int gr_gat_node_scores(const float *h, int64_t ldh, const float *a_self, const float *a_neigh, int64_t n_rows,
                       int32_t heads, int32_t dh, float *s, float *t, void *stream);
size_t gr_gat_aggregate_workspace_bytes(int32_t n_seg, int32_t heads, int32_t dh);
int gr_gat_aggregate(const int32_t *indptr, const int32_t *indices, int64_t n_rows, const float *h, int64_t ldh,
                     const float *s, const float *t, int32_t heads, int32_t dh, float slope, int32_t mean_heads,
                     int32_t elu, float drop_p, uint64_t drop_seed, const uint64_t *drop_seed_dev, int64_t n_cols,
                     const gr_gat_segments *segs_host, float *out, int64_t ldo, float *m_out, float *z_out,
                     void *workspace, size_t workspace_bytes, void *stream);

/* gr_gat_bwd: backward of node scores + aggregation (autograd under trainer.py:270 through gat.py:97-149).
 *   Inputs: what the forward produced (h, s, t, m, z, out) and dout = dL/dout.  (t_indptr, t_indices) is the
 *   CSR of the TRANSPOSED pattern (the same arrays as (indptr, indices) for the symmetric bipartite adjacency),
 *   row_segs / col_segs the segment tables of the two patterns.
 *   Outputs: dH [n, heads*dh] (complete gradient w.r.t. H = x Wcat, incl. the a_self / a_neigh paths) and
 *   da [2, heads*dh] = (d a_self | d a_neigh).  One warp per work item recomputes the softmax weights from (m, z);
 *   the softmax-backward centring term D_i = sum_k alpha_ik <dO_i, H_k> is re-formed from the same weights
 *   (not from the forward output) so that the differences x_ik - D_i survive in fp32.
 *   No atomics, no edge-sized temporaries, deterministic.  dh/4 must be a power of two. */
size_t gr_gat_bwd_workspace_bytes(int64_t n_rows, int32_t heads, int32_t dh, int32_t n_seg_row, int32_t n_seg_col,
                                  int32_t n_long_col);
int gr_gat_bwd(const int32_t *indptr, const int32_t *indices, const int32_t *t_indptr, const int32_t *t_indices,
               int64_t n_rows, int64_t n_cols, const float *h, int64_t ldh, const float *s, const float *t,
               const float *m, const float *z, const float *out, int64_t ldo, const float *dout, int64_t lddo,
               const float *a_self, const float *a_neigh, int32_t heads, int32_t dh, float slope,
               int32_t mean_heads, int32_t elu, float drop_p, uint64_t drop_seed, const uint64_t *drop_seed_dev,
               const gr_gat_segments *row_segs_host, const gr_gat_segments *col_segs_host, float *dH, float *da,
               void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * BPR step
 * ------------------------------------------------------------------------------------------ */

/* HOST function.  Replaces Trainer._sample_batch (src/training/trainer.py:146-197) for
 * negative_samples = 1: B indices with replacement, then per sample one negative draw plus up
 * to 10 redraws while the draw is a known positive of that user (the 10th accepted unchecked).
 * Consumes the mt19937 stream of torch's global CPU generator: mt_state (624 uint64 words),
 * mt_left, mt_next are the fields of torch.get_rng_state() (at::mt19937 layout) and are advanced
 * in place; torch.randint(0, n) == mt() % n.  pos_indptr/pos_items: per-user SORTED positives
 * (trainer.py:169-172).  Returns the number of 32-bit draws consumed (>= 0) or GR_ERR_*. */
int64_t gr_sample_bpr_batch(uint64_t *mt_state_host, int32_t *mt_left_host, uint64_t *mt_next_host,
                            const int64_t *train_user_host, const int64_t *train_item_host, int64_t n_train,
                            int64_t n_items, int64_t batch, const int64_t *pos_indptr_host,
                            const int32_t *pos_items_host, int64_t *users_host, int64_t *pos_host,
                            int64_t *neg_host);

/* Replaces the step body of Trainer.train_epoch (trainer.py:257-264: gathers + row dots),
 * BPRLoss.forward (src/training/losses.py:44-53, the [B] - [B,1] -> [B,B] broadcast form) and the
 * autograd backward through both, in one cooperative kernel:
 *   loss[0]  = mean_{i,j} softplus(n_i - p_j)
 *   grad    += grad_scale * dloss/d emb        (rows of users/pos/neg; duplicates accumulate)
 * emb / grad: [n_users + n_items, d] propagated embeddings, users first; grad must be zeroed by
 * the caller.  workspace: gr_bpr_workspace_bytes(batch) bytes, zeroed once before first use. */
size_t gr_bpr_workspace_bytes(int64_t batch);
int gr_bpr_fused(const float *emb, int64_t ld, int64_t n_users, int64_t n_items, const int64_t *users,
                 const int64_t *pos, const int64_t *neg, int64_t batch, int32_t d, float grad_scale, float *grad,
                 int64_t ldg, float *loss, void *workspace, size_t workspace_bytes, void *stream);

/* Partitioned build (no reference counterpart): the rows rank, rank+world, ... of the
 * same matrix gr_build_csr_pattern + gr_csr_normalize produce, straight from the (user,item) pairs, for
 * the cyclic row distribution of the multi-GPU propagation.  Every rank passes ALL pairs (the degrees are
 * global); only the directed entries of its own rows are kept, sorted and merged.  Local row j is global
 * row rank + j*world; column ids leave in the owner-major padded numbering of the exchange buffers,
 * (c % world) * block_rows + c / world, while the order inside a row stays ascending GLOBAL column (the
 * single-GPU chain order), so the partitioned propagation stays bit-identical.  n_local_entries: the
 * exact number of directed entries (with duplicates, with self loops) whose row this rank owns - the
 * caller counts it; status bit 3 flags a wrong count.  deg: [n_users + n_items] global degrees (output).
 * gr_csr_normalize_local forms the values from the global degrees with the same LUT rule. */
size_t gr_build_local_csr_workspace_bytes(int64_t n_local_entries);
int gr_build_local_csr_pattern(const int64_t *user, const int64_t *item, int64_t n_pairs, int64_t n_users,
                               int64_t n_items, int32_t self_loop, int32_t world, int32_t rank, int64_t block_rows,
                               int64_t n_local_entries, int32_t *indptr, int32_t *indices, float *mult, int32_t *deg,
                               int64_t *nnz_out, int32_t *max_deg_out, int32_t *status, void *workspace,
                               size_t workspace_bytes, void *stream);
int gr_csr_normalize_local(const int32_t *indptr, const int32_t *indices, const float *mult, const int32_t *deg,
                           const float *lut, int64_t lut_len, int64_t n_local_rows, int64_t nnz, int32_t mode,
                           int32_t world, int32_t rank, int64_t block_rows, float *vals, int32_t *status, void *stream);

/* Replaces the per-user temporal split (src/data/dataset.py:327-357: `sort_values(['userId','timestamp'])`,
 * then per user the last row -> test (users with >= 2 rows), the second-last -> valid (users with
 * >= 3 rows), the rest -> train).  user / item / timestamp: [n] int64 device arrays; ts_min / ts_max:
 * the timestamp range (the caller's min/max; rows outside set *status bit 0).  Rows are ordered by one
 * stable radix sort of (user, timestamp - ts_min) keys, so equal timestamps keep their input order as
 * pandas' lexsort does.  Outputs (caller-allocated, [n] for train, [min(n, n_users)] for valid/test):
 * train pairs in (user, timestamp) order - the order Trainer.train_epoch indexes (trainer.py:223-230) -
 * valid / test pairs in user order; counts (device int64[3]) = sizes of the three sets. */
size_t gr_temporal_split_workspace_bytes(int64_t n);
int gr_temporal_split(const int64_t *user, const int64_t *item, const int64_t *timestamp, int64_t n, int64_t n_users,
                      int64_t ts_min, int64_t ts_max, int64_t *train_u, int64_t *train_i, int64_t *valid_u,
                      int64_t *valid_i, int64_t *test_u, int64_t *test_i, int64_t *counts, int32_t *status,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)` + `optimizer.step()`
 * of `optim.Adam(lr, weight_decay)` (src/training/trainer.py:273-276, 81-85) for a list of dense fp32
 * tensors: one pass for the total gradient norm (deterministic), one pass that applies
 *   g = clip*g + wd*p;  m += (g-m)(1-beta1);  v = v*beta2 + (1-beta2) g*g;
 *   p -= step_size * m / (sqrt(v)/bias_correction2_sqrt + eps)
 * with clip = min(1, max_norm/(norm+1e-6)) (max_norm <= 0: no clipping).  step_size = lr/(1-beta1^t) and
 * bias_correction2_sqrt = sqrt(1-beta2^t) are computed by the caller in double, as torch does; the
 * scalars are doubles and rounded to fp32 once inside (1-beta in double first), like torch's.
 * The *_host arguments are host arrays of n_tensors device pointers / element counts (16-byte aligned
 * tensors).  norm_out (device float, optional) receives the total norm.  Gradients are not modified.
 * step_scalars_dev (optional, DEVICE float[2] = (step_size, bias_correction2_sqrt) already rounded to fp32):
 * when non-NULL the two step-dependent scalars are read from it at run time instead of from the arguments, so
 * a training step captured once in a CUDA graph stays valid for every later optimizer step. */
size_t gr_clip_adam_workspace_bytes(const int64_t *numel_host, int32_t n_tensors);
int gr_clip_adam_fused(void *const *params_host, const void *const *grads_host, void *const *exp_avg_host,
                       void *const *exp_avg_sq_host, const int64_t *numel_host, int32_t n_tensors, double max_norm,
                       double step_size, double beta1, double beta2, double eps, double weight_decay,
                       double bias_correction2_sqrt, const float *step_scalars_dev, float *norm_out,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Full-ranking evaluation
 * ------------------------------------------------------------------------------------------ */

/* Replaces the score / mask / top-K loop (src/evaluation/evaluator.py:96-106,
 * src/training/trainer.py:327-337): S = U_b I^T as k-sequential fmaf chains (bit-identical to the
 * reference's CPU sgemm), seen items (CSR over eval rows, sorted GLOBAL item ids) -> -inf, top-k
 * ordered by (score desc, item id asc).  No score matrix is materialised.  k <= 64, d % 16 == 0.
 *
 * gr_score_topk_partial: lists for the item id range [item_lo, item_hi) only (item_emb row 0 is
 *   item_lo), split into n_splits sub-ranges -> part_scores / part_ids [n_splits][n_eval][k]
 *   (short lists padded with id -1).  This is the per-GPU step of the item-sharded evaluation.
 * gr_topk_merge: merges n_parts partial lists per row (ids disjoint) into the final top-k.
 * gr_score_topk: both, on one GPU.  workspace: gr_score_topk_workspace_bytes(n_eval, k, n_splits). */
size_t gr_score_topk_workspace_bytes(int64_t n_eval, int32_t k, int32_t n_splits);
int gr_score_topk_partial(const float *user_emb, int64_t ldu, const float *item_emb, int64_t ldi, int32_t d,
                          const int64_t *eval_users, int64_t n_eval, int64_t item_lo, int64_t item_hi,
                          const int64_t *seen_indptr, const int32_t *seen_items, int32_t k, int32_t n_splits,
                          float *part_scores, int32_t *part_ids, void *stream);
int gr_topk_merge(const float *part_scores, const int32_t *part_ids, int32_t n_parts, int64_t n_eval, int32_t k,
                  int64_t *out_ids, float *out_scores, void *stream);
int gr_score_topk(const float *user_emb, int64_t ldu, const float *item_emb, int64_t ldi, int32_t d,
                  const int64_t *eval_users, int64_t n_eval, int64_t n_items, const int64_t *seen_indptr,
                  const int32_t *seen_items, int32_t k, int32_t n_splits, int64_t *topk_ids, float *topk_scores,
                  void *workspace, size_t workspace_bytes, void *stream);

/* Tensor-core (tcgen05, TF32) NOMINATION pass for the same loop + exact re-scoring.
 * All U x I scores are computed with tcgen05.mma.kind::tf32 (accumulators in TMEM); per user the
 * kprime best approximate scores of unseen items are kept; those candidates are re-scored with the
 * exact fmaf chain and ranked canonically; a row is PROVEN when a_min + eps*|u|*max|i| < (k-th exact
 * score) (a_min = kprime-th approximate score; with several item ranges per user, the largest a_min of
 * the ranges that rejected anything), i.e. no item outside the candidates can enter or tie into the
 * top-k.  Proven rows of topk_ids / topk_scores are bit-identical to gr_score_topk; flags[row] = 1
 * (and *n_flagged, device int32) marks rows the caller must re-rank with gr_score_topk.
 * gr_topk_tc_supported: d % 32 == 0, shared memory fits (d = 32 or 64), k <= kprime <= 64. */
int gr_topk_tc_supported(int32_t d, int32_t kprime);
size_t gr_topk_tc_workspace_bytes(int64_t n_eval, int32_t kprime);
int gr_score_topk_tc(const float *user_emb, int64_t ldu, const float *item_emb, int64_t ldi, int32_t d,
                     const int64_t *eval_users, int64_t n_eval, int64_t n_items, const int64_t *seen_indptr,
                     const int32_t *seen_items, int32_t k, int32_t kprime, int64_t *topk_ids, float *topk_scores,
                     int32_t *flags, int32_t *n_flagged, void *workspace, size_t workspace_bytes, void *stream);

/* Replaces the per-user python loops of compute_metrics_from_topk (src/training/metrics.py:355-432).
 * topk_ids [n_eval][kmax] (kmax <= 64, -1 = padding); ground truth as CSR over the SAME rows
 * (gt_indptr [n_eval+1], gt_items sorted and unique per row, empty row = user without ground truth,
 * skipped as metrics.py:390-394 does); k_values_host: nk <= 8 cut-offs (each clamped to kmax);
 * disc [kmax] = 1/log2(rank+2) and idcg [kmax+1] = running ideal DCG, device float64 tables built by
 * the caller with numpy so that per-user values carry the reference's bits.
 * out_sums  (device, nk*3+1 doubles): per k sums of recall, ndcg, precision over users with ground
 *           truth (fixed-order reduction), then the number of such users.
 * out_counts (device, nk*3 int64): per k the number of distinct recommended ids, the total number of
 *           recommendations, and sum_i (i+1)*c_(i) over the ascending item counts (Gini numerator,
 *           metrics.py:415-430) - all exact integers. */
size_t gr_topk_metrics_workspace_bytes(int64_t n_eval, int64_t n_items, int32_t nk);
int gr_topk_metrics(const int64_t *topk_ids, int64_t n_eval, int32_t kmax, const int64_t *gt_indptr,
                    const int32_t *gt_items, int64_t n_items, const int32_t *k_values_host, int32_t nk,
                    const double *disc, const double *idcg, double *out_sums, int64_t *out_counts, void *workspace,
                    size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GR_B200_H */
