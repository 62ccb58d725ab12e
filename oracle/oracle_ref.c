/* CPU ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py for the policy header).
 *
 * Plain-C restatement of the inner arithmetic of the reference's hot path.  The reference
 * (timur1arkhipov/gnn-recommendations) is pure Python; the arithmetic below is what the
 * torch / MKL CPU kernels it calls evaluate, restated as explicit loops and checked against
 * those kernels and against tests/golden/ in tests/test_oracle_golden.py.
 *
 * Build: make -C oracle   (gcc -O2 -fPIC -shared -ffp-contract=off; single-threaded: libgomp is not in the image)
 * Citations are relative to /root/reference/gnn-recommendations/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* torch.sparse.mm(adj, x) on CPU (src/models/baselines/lightgcn.py:88): per row, entries in
 * storage order, y[r,:] = fma(val, x[col,:], y[r,:]) from zero — one rounding per step. */
void oracle_spmm_fmaf(const int64_t *indptr, const int32_t *indices, const float *vals, int64_t n_rows,
                      const float *x, int64_t d, float *y) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rows; ++r) {
        float *yr = y + r * d;
        for (int64_t f = 0; f < d; ++f) yr[f] = 0.0f;
        for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
            const float v = vals[k];
            const float *xr = x + (int64_t)indices[k] * d;
            for (int64_t f = 0; f < d; ++f) yr[f] = fmaf(v, xr[f], yr[f]);
        }
    }
}

/* LightGCN.forward (lightgcn.py:62-104): L propagations, mean of L+1 layers evaluated as
 * ((x0 + x1) + x2 ...) / (L+1) (torch.mean = sum / count).  out, tmp_a, tmp_b: n*d floats. */
void oracle_lightgcn_forward(const int64_t *indptr, const int32_t *indices, const float *vals, int64_t n,
                             const float *x0, int64_t d, int32_t n_layers, float *out, float *tmp_a,
                             float *tmp_b) {
    const int64_t tot = n * d;
    memcpy(out, x0, (size_t)tot * sizeof(float));
    const float *cur = x0;
    float *bufs[2] = {tmp_a, tmp_b};
    for (int32_t l = 0; l < n_layers; ++l) {
        float *nxt = bufs[l & 1];
        oracle_spmm_fmaf(indptr, indices, vals, n, cur, d, nxt);
        for (int64_t i = 0; i < tot; ++i) out[i] = out[i] + nxt[i];
        cur = nxt;
    }
    const float cnt = (float)(n_layers + 1);
    for (int64_t i = 0; i < tot; ++i) out[i] = out[i] / cnt;
}

/* scores = U_b @ I^T (src/evaluation/evaluator.py:99): MKL sgemm on these shapes equals a
 * k-sequential fmaf chain from zero (checked in tests); seen items -> -inf (:100-104);
 * top-K (:105) in canonical order (score desc, item id asc).  seen_indptr is indexed by the
 * position in eval_users. */
void oracle_score_topk(const float *user_emb, const float *item_emb, int64_t d, const int64_t *eval_users,
                       int64_t n_eval, int64_t n_items, const int64_t *seen_indptr, const int32_t *seen_items,
                       int32_t k, int64_t *topk_ids, float *topk_scores) {
#pragma omp parallel
    {
        float *s = (float *)malloc((size_t)n_items * sizeof(float));
#pragma omp for schedule(dynamic, 8)
        for (int64_t e = 0; e < n_eval; ++e) {
            const float *u = user_emb + eval_users[e] * d;
            for (int64_t i = 0; i < n_items; ++i) {
                const float *it = item_emb + i * d;
                float acc = 0.0f;
                for (int64_t f = 0; f < d; ++f) acc = fmaf(u[f], it[f], acc);
                s[i] = acc;
            }
            if (seen_indptr)
                for (int64_t p = seen_indptr[e]; p < seen_indptr[e + 1]; ++p) s[seen_items[p]] = -INFINITY;
            /* K passes of selection: best = max score, ties -> smallest id */
            for (int32_t j = 0; j < k; ++j) {
                int64_t best = -1;
                for (int64_t i = 0; i < n_items; ++i) {
                    if (isnan(s[i])) continue;
                    if (best < 0 || s[i] > s[best]) best = i;
                }
                topk_ids[e * k + j] = best;
                if (topk_scores) topk_scores[e * k + j] = best >= 0 ? s[best] : NAN;
                if (best >= 0) s[best] = NAN; /* taken */
            }
        }
        free(s);
    }
}

/* std::mt19937 as torch's CPU generator uses it (torch.manual_seed(seed)). */
typedef struct { uint32_t mt[624]; int idx; } oracle_mt19937;
void oracle_mt_seed(oracle_mt19937 *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
uint32_t oracle_mt_next(oracle_mt19937 *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* torch.randint(0, range): one 32-bit draw for range < 2^28, else (hi << 32 | lo) % range. */
static int64_t oracle_draw_below(oracle_mt19937 *g, int64_t range, int64_t *draws) {
    if (range >= ((int64_t)1 << 28)) {
        const uint64_t hi = oracle_mt_next(g), lo = oracle_mt_next(g);
        *draws += 2;
        return (int64_t)(((hi << 32) | lo) % (uint64_t)range);
    }
    *draws += 1;
    return (int64_t)(oracle_mt_next(g) % (uint32_t)range);
}

/* Trainer._sample_batch (src/training/trainer.py:146-197), negative_samples = 1.
 * pos sets given as CSR over users with SORTED item ids.  Returns draws consumed. */
static int in_sorted(const int32_t *a, int64_t n, int32_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo < n && a[lo] == v;
}
int64_t oracle_sample_batch(oracle_mt19937 *g, const int64_t *train_u, const int64_t *train_i, int64_t n_train,
                            int64_t n_items, int64_t batch, const int64_t *pos_indptr, const int32_t *pos_items,
                            int64_t *users, int64_t *pos, int64_t *neg) {
    int64_t draws = 0;
    if (batch > n_train) batch = n_train;
    for (int64_t b = 0; b < batch; ++b) {               /* indices = randint(0, len(train), (B,))  :162 */
        const int64_t idx = oracle_draw_below(g, n_train, &draws);
        users[b] = train_u[idx];
        pos[b] = train_i[idx];
    }
    for (int64_t b = 0; b < batch; ++b) {               /* :174-187 */
        const int64_t u = users[b];
        const int32_t *ps = pos_items + pos_indptr[u];
        const int64_t np_ = pos_indptr[u + 1] - pos_indptr[u];
        int64_t cand = oracle_draw_below(g, n_items, &draws);
        for (int t = 0; t < 10; ++t) {
            if (!in_sorted(ps, np_, (int32_t)cand)) break;
            cand = oracle_draw_below(g, n_items, &draws);
        }
        neg[b] = cand;
    }
    return draws;
}

/* BPR step body, closed form of trainer.py:257-264 + losses.py:44-53 in double:
 * loss = mean_{i,j} softplus(n_i - p_j);  dp_j = -(1/B^2) sum_i sigma(n_i - p_j); dn_i = +... */
double oracle_bpr_loss(const float *user_emb, const float *item_emb, int64_t d, const int64_t *users,
                       const int64_t *pos, const int64_t *neg, int64_t b, double *dp, double *dn) {
    double *p = (double *)malloc((size_t)b * sizeof(double)), *n = (double *)malloc((size_t)b * sizeof(double));
    for (int64_t j = 0; j < b; ++j) {
        double sp = 0, sn = 0;
        for (int64_t f = 0; f < d; ++f) {
            sp += (double)user_emb[users[j] * d + f] * item_emb[pos[j] * d + f];
            sn += (double)user_emb[users[j] * d + f] * item_emb[neg[j] * d + f];
        }
        p[j] = sp; n[j] = sn;
        if (dp) dp[j] = 0; if (dn) dn[j] = 0;
    }
    double loss = 0;
    for (int64_t i = 0; i < b; ++i)
        for (int64_t j = 0; j < b; ++j) {
            const double x = n[i] - p[j];
            loss += x > 0 ? x + log1p(exp(-x)) : log1p(exp(x));
            const double s = 1.0 / (1.0 + exp(-x));
            if (dp) dp[j] -= s; if (dn) dn[i] += s;
        }
    const double bb = (double)b * (double)b;
    if (dp) for (int64_t j = 0; j < b; ++j) dp[j] /= bb;
    if (dn) for (int64_t j = 0; j < b; ++j) dn[j] /= bb;
    free(p); free(n);
    return loss / bb;
}
