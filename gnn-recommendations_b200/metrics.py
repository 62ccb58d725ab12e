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
