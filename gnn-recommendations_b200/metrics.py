"""compute_metrics_from_topk — drop-in for src/training/metrics.py:355-432 (the consumer of the
top-K lists).  Same definitions and float64 host arithmetic, vectorised over users; values are
identical to the reference's Python loops (same per-user terms, same np.mean)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch


def compute_metrics_from_topk(topk_items, user_ids: Sequence[int], ground_truth: Dict[int, List[int]], n_items: int,
                              k_values: Sequence[int] = (10, 20)) -> Dict[str, float]:
    topk_np = topk_items.cpu().numpy() if isinstance(topk_items, torch.Tensor) else np.asarray(topk_items)
    if topk_np.size == 0:
        return {}
    max_k = topk_np.shape[1]
    rows = [idx for idx, u in enumerate(user_ids) if ground_truth.get(u)]
    rel_sets = [set(ground_truth[user_ids[idx]]) for idx in rows]
    n_rel = np.array([len(s) for s in rel_sets], dtype=np.int64)
    # hit matrix [rows, max_k]
    hit = np.zeros((len(rows), max_k), dtype=bool)
    for j, (idx, s) in enumerate(zip(rows, rel_sets)):
        hit[j] = np.fromiter((it in s for it in topk_np[idx].tolist()), dtype=bool, count=max_k)
    disc = 1.0 / np.log2(np.arange(max_k) + 2)
    metrics: Dict[str, float] = {}
    for k in k_values:
        k = min(k, max_k)
        if rows:
            hk = hit[:, :k]
            # a list never repeats an item, so |relevant ∩ set(pred)| == number of hit positions
            hits = hk.sum(axis=1)
            recalls = (hits / n_rel).tolist()
            precisions = (hits / k).tolist()
            ndcgs = []
            for j in range(len(rows)):
                dcg = 0.0
                for rank in np.flatnonzero(hk[j]):
                    dcg += 1.0 / np.log2(rank + 2)
                idcg = 0.0
                for rank in range(min(int(n_rel[j]), k)):
                    idcg += 1.0 / np.log2(rank + 2)
                ndcgs.append(dcg / idcg if idcg > 0 else 0.0)
        else:
            recalls, precisions, ndcgs = [], [], []
        metrics[f"recall@{k}"] = float(np.mean(recalls)) if recalls else 0.0
        metrics[f"ndcg@{k}"] = float(np.mean(ndcgs)) if ndcgs else 0.0
        metrics[f"precision@{k}"] = float(np.mean(precisions)) if precisions else 0.0
        flat = topk_np[:, :k].ravel()
        metrics[f"coverage@{k}"] = len(np.unique(flat)) / max(1, n_items)
        counts = np.sort(np.bincount(flat, minlength=n_items).astype(np.int64))
        if counts.sum() > 0:
            n = len(counts)
            cumsum = np.cumsum(counts)
            metrics[f"gini@{k}"] = float((2 * np.sum((np.arange(n) + 1) * counts)) / (n * cumsum[-1]) - (n + 1) / n)
        else:
            metrics[f"gini@{k}"] = 0.0
    del disc
    return metrics


def discount_tables(max_k: int):
    """(disc [max_k], idcg [max_k+1]) float64: 1/log2(rank+2) and its running sum in rank order —
    built with numpy exactly as metrics.py:402-409 accumulates them."""
    disc = np.array([1.0 / np.log2(rank + 2) for rank in range(max_k)], dtype=np.float64)
    idcg = np.zeros(max_k + 1, dtype=np.float64)
    acc = 0.0
    for rank in range(max_k):
        acc += 1.0 / np.log2(rank + 2)
        idcg[rank + 1] = acc
    return disc, idcg


def topk_metrics_device(topk_items: torch.Tensor, gt_indptr, gt_items, n_items: int,
                        k_values: Sequence[int] = (10, 20)) -> Dict[str, float]:
    """compute_metrics_from_topk on the device (gr_topk_metrics).

    topk_items: [n_eval, max_k] int64 CUDA tensor (rows = eval users); ground truth as CSR over the
    same rows (sorted unique item ids; an empty row = user without ground truth, skipped as
    metrics.py:390-394).  Same keys and values as the reference: per-user terms are bit-identical,
    the means differ from np.mean only by summation order (~1e-16)."""
    from ._lib import check, lib, ptr, stream_ptr
    import ctypes as C

    if topk_items.numel() == 0:
        return {}
    if not topk_items.is_cuda:
        raise RuntimeError("topk_metrics_device needs the lists on a CUDA device (no CPU fallback)")
    dev = topk_items.device
    topk_items = topk_items.contiguous().to(torch.int64)
    n_eval, max_k = topk_items.shape
    ks = [min(int(k), max_k) for k in k_values]
    nk = len(ks)
    if nk == 0:
        return {}
    gt_indptr = torch.as_tensor(gt_indptr, dtype=torch.int64).to(dev).contiguous()
    gt_items = torch.as_tensor(gt_items, dtype=torch.int32).to(dev).contiguous()
    if gt_items.numel() == 0:
        gt_items = torch.zeros(1, dtype=torch.int32, device=dev)
    disc, idcg = discount_tables(max_k)
    disc_d, idcg_d = torch.from_numpy(disc).to(dev), torch.from_numpy(idcg).to(dev)
    l = lib()
    sums = torch.empty(nk * 3 + 1, dtype=torch.float64, device=dev)
    counts = torch.empty(nk * 3, dtype=torch.int64, device=dev)
    ws_bytes = l.gr_topk_metrics_workspace_bytes(n_eval, n_items, nk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    karr = (C.c_int32 * nk)(*ks)
    check(l.gr_topk_metrics(ptr(topk_items), n_eval, max_k, ptr(gt_indptr), ptr(gt_items), n_items, karr, nk,
                            ptr(disc_d), ptr(idcg_d), ptr(sums), ptr(counts), ptr(ws), ws_bytes, stream_ptr()),
          "gr_topk_metrics")
    sums_h, counts_h = sums.cpu().numpy(), counts.cpu().numpy()
    n_valid = int(sums_h[-1])
    metrics: Dict[str, float] = {}
    for j, k in enumerate(ks):
        for name, v in zip(("recall", "ndcg", "precision"), sums_h[3 * j:3 * j + 3]):
            metrics[f"{name}@{k}"] = float(v / n_valid) if n_valid else 0.0
        uniq, total, wsum = (np.int64(x) for x in counts_h[3 * j:3 * j + 3])
        metrics[f"coverage@{k}"] = int(uniq) / max(1, n_items)
        if total > 0:
            n = int(n_items)
            metrics[f"gini@{k}"] = float((2 * wsum) / (n * total) - (n + 1) / n)     # metrics.py:426-427
        else:
            metrics[f"gini@{k}"] = 0.0
    return metrics
