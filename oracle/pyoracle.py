"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.

A CPU restatement (numpy + torch-CPU) of the reference's hot path
(timur1arkhipov/gnn-recommendations): graph build -> propagation -> BPR step ->
full-ranking top-K.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file;
the product package (``gnn-recommendations_b200/``) never does and fails loudly
when its CUDA library is missing.

Where the arithmetic lives.  The reference is pure Python; its arithmetic is done
by third-party wheels that are not under /root/reference: torch (``torch>=2.0.0``,
requirements.txt:2, unpinned; 2.11.0+cu128 here) for ``torch.sparse.mm``, ``@``,
``topk``, ``logsigmoid``, ``matrix_exp``, autograd, Adam; scipy (>=1.10; 1.18.1
here) and numpy (>=1.24; 2.3.5 here) for the normalisation.  This oracle calls the
same wheel ops in the same order on CPU and restates the Python control flow around
them; ``oracle/oracle_ref.c`` restates the inner arithmetic itself (sequential
``fmaf`` chains, mt19937) in plain C and is cross-checked against this file.

Pinning.  The reference has NO tests and NO golden vectors (tests/ holds only
.gitkeep) -> "parity unpinned by the reference".  The pins used instead:
``tests/golden/*.npz`` = outputs of the UNMODIFIED reference classes imported from
/root/reference and run in the build container on seeded synthetic inputs
(generator script: tests/golden/make_golden.py), plus the parameter-count KATs of
/root/reference/problems.md:95-124.  tests/test_oracle_golden.py checks every
function below against those fixtures.

All file:line citations are relative to /root/reference/gnn-recommendations/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# 1. Graph build  (src/data/graph_builder.py:16-80, 83-144, 147-174)
# --------------------------------------------------------------------------------------
def dis_lut(max_deg: int, power: float = -0.5) -> np.ndarray:
    """``np.power(float32(max(deg,1)), -0.5)`` for deg = 0..max_deg.

    graph_builder.py:114 clamps the degree to >= 1 and :119 takes ``np.power`` of the
    float32 degree vector.  numpy's f32 pow is not correctly rounded, so the only way
    to reproduce Â's values bit-for-bit is to ask the same numpy for them; degrees are
    integers, hence a look-up table over 0..max_deg suffices."""
    deg = np.maximum(np.arange(max_deg + 1, dtype=np.float32), np.float32(1.0))
    return np.power(deg, power).astype(np.float32)


def build_norm_adj(user: np.ndarray, item: np.ndarray, n_users: int, n_items: int,
                   normalization: str = "symmetric") -> Dict[str, np.ndarray]:
    """Â = D^-1/2 [[0,R],[R^T,0]] D^-1/2 as canonical CSR.

    Restates build_bipartite_graph (graph_builder.py:49-70: rows=[u ; U+i],
    cols=[U+i ; u], ones), ``tocsr`` (:107, duplicates are SUMMED), the degree vector
    (:111 row sums, :114 clamp), ``d^-1/2`` (:119) and ``D·A·D`` evaluated left to
    right (:126: ``(D @ A) @ D`` -> val = fl(fl(dis[r]*a) * dis[c])); ``'row'`` is
    :128-134 (val = fl(dinv[r]*a)); output order is row-major with ascending columns
    (what ``tocoo`` of a canonical CSR yields, :140)."""
    n = n_users + n_items
    user = np.asarray(user, dtype=np.int64)
    item = np.asarray(item, dtype=np.int64)
    rows = np.concatenate([user, n_users + item])
    cols = np.concatenate([n_users + item, user])
    key = rows * n + cols
    ukey, mult = np.unique(key, return_counts=True)
    r = (ukey // n).astype(np.int64)
    c = (ukey % n).astype(np.int32)
    a = mult.astype(np.float32)
    deg = np.bincount(r, weights=a.astype(np.float64), minlength=n).astype(np.float32)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=indptr[1:])
    degc = np.maximum(deg, np.float32(1.0))
    if normalization == "symmetric":
        dis = np.power(degc, -0.5)
        dis[np.isinf(dis)] = 0.0
        vals = (dis[r] * a) * dis[c]
    elif normalization == "row":
        dis = np.power(degc, -1.0)
        dis[np.isinf(dis)] = 0.0
        vals = dis[r] * a
    elif normalization == "none":
        dis = np.ones(n, dtype=np.float32)
        vals = a
    else:
        raise ValueError(f"unknown normalization: {normalization}")
    return {"indptr": indptr, "indices": c, "vals": vals.astype(np.float32),
            "rows": r, "deg": deg, "dis": dis.astype(np.float32), "n": n}


def to_torch_coo(adj: Dict[str, np.ndarray]) -> torch.Tensor:
    """convert_to_torch_sparse (graph_builder.py:147-174): int64 indices [2,nnz],
    float32 values, not flagged coalesced."""
    idx = torch.from_numpy(np.vstack([adj["rows"], adj["indices"].astype(np.int64)]))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(adj["vals"]), (adj["n"], adj["n"]))


# --------------------------------------------------------------------------------------
# 2. Propagation
# --------------------------------------------------------------------------------------
def spmm(adj_coo: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """lightgcn.py:88 — ``torch.sparse.mm`` on the CPU (bit-identical to a per-row
    ascending-column fmaf chain from a zero accumulator; oracle_ref.c:spmm_fmaf)."""
    return torch.sparse.mm(adj_coo, x)


def lightgcn_layers(adj_coo, user_w, item_w, n_layers: int) -> List[torch.Tensor]:
    """LightGCN.get_layer_embeddings (lightgcn.py:153-183)."""
    x = torch.cat([user_w, item_w], dim=0)
    out = [x]
    for _ in range(n_layers):
        x = spmm(adj_coo, x)
        out.append(x)
    return out


def lightgcn_forward(adj_coo, user_w, item_w, n_layers: int):
    """LightGCN.forward (lightgcn.py:62-104): L x SpMM, mean over the L+1 layers
    (``torch.mean(torch.stack(...), 0)``), split users/items."""
    layers = lightgcn_layers(adj_coo, user_w, item_w, n_layers)
    x = torch.mean(torch.stack(layers, dim=0), dim=0)
    return torch.split(x, [user_w.shape[0], item_w.shape[0]], dim=0)


def ngcf_forward(adj_coo, user_w, item_w, w1: Sequence[torch.Tensor], b1, w2, b2):
    """NGCF.forward in eval() mode (ngcf.py:157-195) over NGCFLayer.forward
    (:52-86): n = Âx; out = LeakyReLU_0.2(W1 n + b1 + W2 (x*n) + b2); dropout is the
    identity in eval(); result = concat of the L+1 layer outputs."""
    x = torch.cat([user_w, item_w], dim=0)
    outs = [x]
    for l in range(len(w1)):
        n = spmm(adj_coo, x)
        inter = x * n
        o = F.linear(n, w1[l], b1[l]) + F.linear(inter, w2[l], b2[l])
        x = F.leaky_relu(o, negative_slope=0.2)
        outs.append(x)
    xf = torch.cat(outs, dim=1)
    return torch.split(xf, [user_w.shape[0], item_w.shape[0]], dim=0)


def gat_layer_sparse(indptr: np.ndarray, indices: np.ndarray, x: torch.Tensor,
                     W: Sequence[torch.Tensor], a_self, a_neigh, alpha: float, concat: bool):
    """GATLayer.forward (gat.py:76-151) restated over the CSR *pattern* (Â's values are
    ignored, gat.py:120-127): per head h = x W^T; e_ij = LeakyReLU_alpha(a_self.h_i +
    a_neigh.h_j) on edges; softmax over the row's neighbours; out_i = sum_j a_ij h_j;
    heads concatenated (:144-145) or averaged (:147).  The dense reference needs N^2
    floats and cannot run at C3; this restatement is validated against the dense class
    on small graphs (golden fixture 'gat_*').  A degree-0 row yields NaN like the
    reference's softmax over an all -inf row."""
    n = x.shape[0]
    ip = torch.from_numpy(np.asarray(indptr, dtype=np.int64))
    col = torch.from_numpy(np.asarray(indices, dtype=np.int64))
    row = torch.repeat_interleave(torch.arange(n), ip[1:] - ip[:-1])
    heads = []
    for h_i in range(len(W)):
        h = F.linear(x, W[h_i])
        s = (h @ a_self[h_i]).squeeze(1)
        t = (h @ a_neigh[h_i]).squeeze(1)
        e = F.leaky_relu(s[row] + t[col], negative_slope=alpha)
        m = torch.full((n,), float("-inf")).scatter_reduce(0, row, e, reduce="amax", include_self=True)
        p = torch.exp(e - m[row])
        z = torch.zeros(n).index_add_(0, row, p)
        w = p / z[row]
        out = torch.zeros(n, h.shape[1]).index_add_(0, row, w.unsqueeze(1) * h[col])
        out[z == 0] = float("nan")
        heads.append(out)
    if concat:
        return torch.cat(heads, dim=1)
    return torch.stack(heads, dim=0).mean(dim=0)


def gat_forward_sparse(indptr, indices, user_w, item_w, layers: Sequence[dict], alpha: float):
    """GAT.forward in eval() mode (gat.py:258-297): L GAT layers with ELU after each
    (:283), mean of the L+1 layer outputs (:287-288).  ``layers[l]`` holds lists
    'W', 'a_self', 'a_neigh' and bool 'concat'."""
    x = torch.cat([user_w, item_w], dim=0)
    outs = [x]
    for lay in layers:
        x = gat_layer_sparse(indptr, indices, x, lay["W"], lay["a_self"], lay["a_neigh"], alpha, lay["concat"])
        x = F.elu(x)
        outs.append(x)
    xf = torch.mean(torch.stack(outs, dim=0), dim=0)
    return torch.split(xf, [user_w.shape[0], item_w.shape[0]], dim=0)


def block_orthogonal(skew_params: Sequence[torch.Tensor]) -> torch.Tensor:
    """blockdiag(exp(P_k - P_k^T)) — group_shuffle_layer.py:107-129 and
    bundle_layer.py:56-69."""
    return torch.block_diag(*[torch.matrix_exp(p - p.T) for p in skew_params])


def gs_forward(adj_coo, user_w, item_w, conn_skew, conn_perm, local_skew, local_perm,
               layer_weights, residual_alpha: float):
    """OrthogonalBundleGNN.forward, adj_matrix mode, dropout 0 (model.py:120-213):
    per layer c = Âx (:171-174); t = c @ W_conn, W_conn = blockdiag(...)[:, perm_c]
    (bundle_layer.py:59-73); g = (t @ W_orth)[:, perm_g] (group_shuffle_layer.py:88-94);
    x = (1-a) g + a x0 (:194-195); out = sum_l softmax(layer_weights)_l x_l (:204-207,
    a Python ``sum`` starting from int 0)."""
    x0 = torch.cat([user_w, item_w], dim=0)
    x = x0
    outs = [x]
    for l in range(len(local_skew)):
        c = spmm(adj_coo, x)
        if conn_skew is not None:
            c = c @ block_orthogonal(conn_skew[l])[:, conn_perm[l]]
        g = (c @ block_orthogonal(local_skew[l]))[:, local_perm[l]]
        x = (1 - residual_alpha) * g + residual_alpha * x0
        outs.append(x)
    w = F.softmax(layer_weights, dim=0)
    xf = sum([wi * e for wi, e in zip(w, outs)])
    nu = user_w.shape[0]
    return xf[:nu], xf[nu:]


def gs_forward_edge_index(edge_index: torch.Tensor, user_w, item_w, conn_skew, conn_perm, local_skew, local_perm,
                          layer_weights, residual_alpha: float, return_layers: bool = False):
    """OrthogonalBundleGNN.forward in its EDGE-LIST mode (use_edge_index=True, dropout 0; model.py:159-222):
    per layer x_j <- sum over edges (i -> j) of W_conn x_i, written as the reference writes it —
    ``x[src] @ W_conn.t()`` then ``index_add_`` over the destinations (parallel_transport.py:27-50) — or, without
    parallel transport (conn_skew None), the plain ``index_add_`` of x[src] (model.py:218-222); then
    g = (t @ W_orth)[:, perm_g], x = (1-a) g + a x0, out = sum_l softmax(layer_weights)_l x_l.
    ``return_layers``: the per-layer outputs WITHOUT the residual (get_layer_embeddings, model.py:306-356)."""
    src, dst = edge_index
    x0 = torch.cat([user_w, item_w], dim=0)

    def transport(x, l):
        xs = x[src]
        if conn_skew is not None:
            xs = torch.mm(xs, block_orthogonal(conn_skew[l])[:, conn_perm[l]].t())
        agg = torch.zeros_like(x)
        agg.index_add_(0, dst, xs)
        return agg

    if return_layers:
        x, layers = x0, [x0.clone()]
        for l in range(len(local_skew)):
            x = (transport(x, l) @ block_orthogonal(local_skew[l]))[:, local_perm[l]]
            layers.append(x.clone())
        return layers
    x, outs = x0, [x0]
    for l in range(len(local_skew)):
        g = (transport(x, l) @ block_orthogonal(local_skew[l]))[:, local_perm[l]]
        x = (1 - residual_alpha) * g + residual_alpha * x0
        outs.append(x)
    w = F.softmax(layer_weights, dim=0)
    xf = sum([wi * e for wi, e in zip(w, outs)])
    nu = user_w.shape[0]
    return xf[:nu], xf[nu:]


# --------------------------------------------------------------------------------------
# 3. BPR sampling and step  (src/training/trainer.py:146-197, 237-279; losses.py:28-53)
# --------------------------------------------------------------------------------------
class TorchCpuMt19937:
    """The global CPU generator ``torch.manual_seed(seed)`` creates: std::mt19937
    seeded with init_genrand(seed & 0xffffffff); ``torch.randint(0, n, ...)`` consumes one
    32-bit output per element and returns ``out % n`` for n < 2^28, and two outputs
    (``(first << 32 | second) % n``) for n >= 2^28 (torch 2.11; verified against torch in
    tests/test_oracle_golden.py)."""

    def __init__(self, seed: int):
        self._bg = np.random.MT19937()
        self._bg._legacy_seeding(int(seed) & 0xFFFFFFFF)

    def randint(self, n: int, size: int) -> np.ndarray:
        if n >= 1 << 28:
            raw = self._bg.random_raw(2 * size).astype(np.uint64)
            return (((raw[0::2] << np.uint64(32)) | raw[1::2]) % np.uint64(n)).astype(np.int64)
        return (self._bg.random_raw(size).astype(np.uint64) % np.uint64(n)).astype(np.int64)

    def randint1(self, n: int) -> int:
        return int(self.randint(n, 1)[0])


def sample_batch(rng: TorchCpuMt19937, train_u: np.ndarray, train_i: np.ndarray, n_items: int,
                 batch_size: int, pos_sets: Dict[int, set], negative_samples: int = 1):
    """Trainer._sample_batch (trainer.py:146-197): B indices WITH replacement (:162),
    then per sample one negative draw plus up to 10 redraws while the draw is a known
    positive of that user; the 10th redraw is accepted unchecked (:183-186)."""
    b = min(batch_size, len(train_u))
    idx = rng.randint(len(train_u), b)
    users = train_u[idx].astype(np.int64)
    pos = train_i[idx].astype(np.int64)
    neg = np.empty((b, negative_samples), dtype=np.int64)
    for k in range(b):
        ps = pos_sets[int(users[k])]
        for j in range(negative_samples):
            cand = rng.randint1(n_items)
            for _ in range(10):
                if cand not in ps:
                    break
                cand = rng.randint1(n_items)
            neg[k, j] = cand
    return users, pos, neg


def positive_sets(train_u: np.ndarray, train_i: np.ndarray) -> Dict[int, set]:
    """trainer.py:169-172."""
    d: Dict[int, set] = {}
    for u, i in zip(train_u.tolist(), train_i.tolist()):
        d.setdefault(u, set()).add(i)
    return d


def bpr_loss_reference(user_emb, item_emb, users, pos, neg):
    """trainer.py:257-264 + losses.py:44-53 with ``neg`` of shape [B,1]: the
    subtraction broadcasts [B] - [B,1] -> [B,B]; loss = mean over all (i,j) of
    softplus(neg_i - pos_j).  Pure torch ops, differentiable."""
    pos_scores = (user_emb[users] * item_emb[pos]).sum(dim=1)
    neg_scores = (user_emb[users].unsqueeze(1) * item_emb[neg]).sum(dim=2)
    diff = pos_scores - neg_scores
    return (-F.logsigmoid(diff)).mean()


def bpr_closed_form(user_emb: np.ndarray, item_emb: np.ndarray, users, pos, neg):
    """float64 closed form of the same loss and of its gradient w.r.t. the propagated
    embeddings: the spec of the fused BPR kernel."""
    U = user_emb.astype(np.float64)
    I = item_emb.astype(np.float64)
    users, pos, neg = np.asarray(users), np.asarray(pos), np.asarray(neg).reshape(-1)
    b = len(users)
    p = (U[users] * I[pos]).sum(1)
    n = (U[users] * I[neg]).sum(1)
    d = n[:, None] - p[None, :]                       # [i, j] = n_i - p_j
    loss = np.logaddexp(0.0, d).mean()
    sig = 1.0 / (1.0 + np.exp(-d))
    dp = -sig.sum(0) / (b * b)
    dn = sig.sum(1) / (b * b)
    gU = np.zeros_like(U)
    gI = np.zeros_like(I)
    np.add.at(gU, users, dp[:, None] * I[pos] + dn[:, None] * I[neg])
    np.add.at(gI, pos, dp[:, None] * U[users])
    np.add.at(gI, neg, dn[:, None] * U[users])
    return loss, gU, gI, p, n


def lightgcn_loss_and_grads(adj_coo, user_w, item_w, n_layers, users, pos, neg):
    """Forward + loss + autograd backward of one body of Trainer.train_epoch's loop
    (trainer.py:249-270) for LightGCN.  Returns (loss, dL/duser_w, dL/ditem_w); the
    clip (trainer.py:273-274) and Adam step (:276) stay stock torch in the product too."""
    uw = user_w.detach().clone().requires_grad_(True)
    iw = item_w.detach().clone().requires_grad_(True)
    ue, ie = lightgcn_forward(adj_coo, uw, iw, n_layers)
    loss = bpr_loss_reference(ue, ie, torch.as_tensor(users), torch.as_tensor(pos),
                              torch.as_tensor(neg).reshape(-1, 1))
    gu, gi = torch.autograd.grad(loss, [uw, iw])
    return loss.detach(), gu, gi


# --------------------------------------------------------------------------------------
# 4. Full-ranking evaluation  (src/evaluation/evaluator.py:96-108, trainer.py:327-339,
#    src/training/metrics.py:355-432)
# --------------------------------------------------------------------------------------
def canonical_topk(scores: torch.Tensor, k: int) -> np.ndarray:
    """Top-k per row ordered by (score desc, item id asc).  ``torch.topk`` returns
    tied entries in arbitrary order on the CPU, so reference lists are canonicalised
    from the reference SCORES."""
    s = scores.numpy()
    n = s.shape[1]
    out = np.empty((s.shape[0], k), dtype=np.int64)
    ids = np.arange(n)
    for r in range(s.shape[0]):
        row = s[r]
        order = np.lexsort((ids, -row.astype(np.float64)))   # -inf -> +inf sorts last; NaN never occurs
        out[r] = order[:k]
    return out


def score_mask_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, eval_users: Sequence[int],
                    seen: Dict[int, Sequence[int]], k: int, batch: int = 2048) -> np.ndarray:
    """The score / mask / top-K loop (evaluator.py:96-106): S = U_b @ I^T, seen items
    set to -inf, top-k; canonical order."""
    outs = []
    for s0 in range(0, len(eval_users), batch):
        bu = list(eval_users[s0:s0 + batch])
        scores = user_emb[torch.tensor(bu)] @ item_emb.T
        for r, u in enumerate(bu):
            it = seen.get(u)
            if it is not None and len(it):
                scores[r, torch.as_tensor(list(it))] = float("-inf")
        outs.append(canonical_topk(scores, k))
    return np.concatenate(outs, axis=0)


def metrics_from_topk(topk: np.ndarray, user_ids: Sequence[int], ground_truth: Dict[int, List[int]],
                      n_items: int, k_values: Sequence[int]) -> Dict[str, float]:
    """compute_metrics_from_topk (metrics.py:355-432), float64 host arithmetic."""
    if topk.size == 0:
        return {}
    max_k = topk.shape[1]
    m: Dict[str, float] = {}
    for k in k_values:
        k = min(k, max_k)
        rec, nd, pr = [], [], []
        for idx, u in enumerate(user_ids):
            rel = set(ground_truth.get(u, ()))
            if not rel:
                continue
            pred = topk[idx, :k].tolist()
            hits = len(rel & set(pred))
            rec.append(hits / len(rel))
            pr.append(hits / k)
            dcg = 0.0
            for rank, it in enumerate(pred):
                if it in rel:
                    dcg += 1.0 / np.log2(rank + 2)
            idcg = 0.0
            for rank in range(min(len(rel), k)):
                idcg += 1.0 / np.log2(rank + 2)
            nd.append(dcg / idcg if idcg > 0 else 0.0)
        m[f"recall@{k}"] = float(np.mean(rec)) if rec else 0.0
        m[f"ndcg@{k}"] = float(np.mean(nd)) if nd else 0.0
        m[f"precision@{k}"] = float(np.mean(pr)) if pr else 0.0
        flat = topk[:, :k].ravel()
        m[f"coverage@{k}"] = len(np.unique(flat)) / max(1, n_items)
        cnt = np.sort(np.bincount(flat, minlength=n_items).astype(np.int64))
        if cnt.sum() > 0:
            n = len(cnt)
            cs = np.cumsum(cnt)
            m[f"gini@{k}"] = float((2 * np.sum((np.arange(n) + 1) * cnt)) / (n * cs[-1]) - (n + 1) / n)
        else:
            m[f"gini@{k}"] = 0.0
    return m
