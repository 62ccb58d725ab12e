"""Synthetic interaction graphs of the named dataset shapes.

Host generator (numpy) for the parity configs C1-C4 and a device generator
(torch CUDA ops, plumbing only) for the throughput-only C5 graph.  The law is
the one the survey fixes: user weights ~ rank^-0.6, item weights ~ rank^-0.8,
unique (user,item) pairs, every user >= 3 interactions, every item >= 1 *train*
interaction under the reference's temporal split (dataset.py:281-364), unique
timestamps.
"""
from __future__ import annotations

import numpy as np

# name -> (n_users, n_items, n_interactions, embedding_dim, n_layers)
SHAPES = {
    "C1": (6040, 3706, 1_000_209, 64, 3),        # MovieLens-1M shape
    "C2": (29858, 40981, 1_027_370, 64, 3),      # Gowalla shape
    "C3": (31668, 38048, 1_561_406, 64, 3),      # Yelp2018 shape
    "C4": (52643, 91599, 2_984_108, 64, 3),      # Amazon-Book shape
    "C5": (20_000_000, 5_000_000, 500_000_000, 128, 4),
    "tiny": (300, 200, 6000, 64, 3),
}


def _power_law_cdf(n: int, alpha: float) -> np.ndarray:
    w = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    c = np.cumsum(w)
    return c / c[-1]


def _draw(rng, cdf: np.ndarray, size) -> np.ndarray:
    return np.minimum(np.searchsorted(cdf, rng.random(size), side="right"), len(cdf) - 1).astype(np.int64)


def synth_interactions(n_users: int, n_items: int, n_interactions: int, seed: int = 42):
    """Returns (user, item, timestamp) int64 arrays of exactly ``n_interactions``
    unique pairs.  Construction: one low-timestamp "cover" edge per item, three
    high-timestamp edges per user, power-law fill; so no cover edge can be among a
    user's last two interactions and every item keeps a train edge after the split."""
    if n_interactions < n_items + 3 * n_users:
        raise ValueError("n_interactions too small for the coverage guarantees")
    if n_items < 4:
        raise ValueError("need at least 4 items")
    rng = np.random.default_rng(seed)
    ucdf = _power_law_cdf(n_users, 0.6)
    icdf = _power_law_cdf(n_items, 0.8)

    cover_u = _draw(rng, ucdf, n_items)
    cover_key = cover_u * n_items + np.arange(n_items, dtype=np.int64)

    base_i = _draw(rng, icdf, (n_users, 3))
    uu = np.repeat(np.arange(n_users, dtype=np.int64), 3).reshape(n_users, 3)
    cover_sorted = np.sort(cover_key)
    while True:
        key = uu * n_items + base_i
        ks = np.sort(key, axis=1)
        dup_row = (ks[:, 0] == ks[:, 1]) | (ks[:, 1] == ks[:, 2])
        pos = np.searchsorted(cover_sorted, key.ravel())
        pos = np.minimum(pos, len(cover_sorted) - 1)
        hit = (cover_sorted[pos] == key.ravel()).reshape(n_users, 3).any(axis=1)
        bad = dup_row | hit
        if not bad.any():
            break
        base_i[bad] = _draw(rng, icdf, (int(bad.sum()), 3))
    base_key = (uu * n_items + base_i).ravel()

    have = np.concatenate([cover_key, base_key])          # unique by construction
    n_fixed = len(have)
    keys = np.unique(have)
    while len(keys) < n_interactions:
        need = n_interactions - len(keys)
        m = int(need * 1.3) + 1024
        cand = _draw(rng, ucdf, m) * n_items + _draw(rng, icdf, m)
        cand = np.unique(cand)
        cand = cand[~np.isin(cand, keys, assume_unique=True)]
        rng.shuffle(cand)
        keys = np.concatenate([keys, cand[:need]])
        keys = np.unique(keys)
    fill_key = keys[~np.isin(keys, have, assume_unique=True)]
    assert n_fixed + len(fill_key) == n_interactions

    all_key = np.concatenate([cover_key, base_key, fill_key])
    ts = np.empty(n_interactions, dtype=np.int64)
    ts[:n_items] = rng.permutation(n_items)
    ts[n_items:] = n_items + rng.permutation(n_interactions - n_items)
    order = rng.permutation(n_interactions)           # file order is arbitrary
    all_key, ts = all_key[order], ts[order]
    return all_key // n_items, all_key % n_items, ts


def temporal_split(user: np.ndarray, item: np.ndarray, ts: np.ndarray):
    """Vectorised restatement of the reference's per-user temporal split
    (dataset.py:327-357): rows sorted by (user, timestamp); per user the last row is
    test, the second-last valid (only when the user has >= 3 rows), the rest train.
    Returns dict of (user,item) int64 pairs; train keeps the (user, timestamp) order,
    which is the order `Trainer.train_epoch` indexes (trainer.py:223-230)."""
    order = np.lexsort((ts, user))
    u, i = user[order], item[order]
    n = len(u)
    starts = np.flatnonzero(np.r_[True, u[1:] != u[:-1]])
    counts = np.diff(np.r_[starts, n])
    cnt = np.repeat(counts, counts)
    from_end = np.repeat(starts + counts, counts) - np.arange(n)   # 1 = last row
    is_test = (from_end == 1) & (cnt >= 2)
    is_valid = (from_end == 2) & (cnt >= 3)
    is_train = ~(is_test | is_valid)
    return {
        "train": (u[is_train], i[is_train]),
        "valid": (u[is_valid], i[is_valid]),
        "test": (u[is_test], i[is_test]),
    }


def synth_split(shape: str, seed: int = 42):
    n_users, n_items, e, _, _ = SHAPES[shape]
    u, i, t = synth_interactions(n_users, n_items, e, seed)
    out = temporal_split(u, i, t)
    out["n_users"], out["n_items"] = n_users, n_items
    out["all"] = (u, i, t)
    return out


def synth_pairs_device(n_users: int, n_items: int, n_edges: int, seed: int, device):
    """Throughput-only generator for C5-scale graphs, on the device (same power law,
    no split, coverage not enforced; isolated nodes are legal for LightGCN —
    graph_builder.py:114 clamps their degree).  Returns unique (user,item) int64
    tensors, exactly ``n_edges`` of them, sorted by (user,item)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def cdf(n, alpha):
        w = torch.arange(1, n + 1, device=device, dtype=torch.float64).pow_(-alpha)
        c = torch.cumsum(w, 0)
        return c.div_(c[-1].clone())

    ucdf, icdf = cdf(n_users, 0.6), cdf(n_items, 0.8)

    def draw(c, m):
        r = torch.rand(m, device=device, dtype=torch.float64, generator=g)
        return torch.searchsorted(c, r, right=True).clamp_(max=c.numel() - 1)

    keys = torch.empty(0, dtype=torch.int64, device=device)
    chunk = 1 << 27
    while keys.numel() < n_edges:
        need = n_edges - keys.numel()
        m = int(need * 1.25) + 4096
        parts = [keys]
        for s in range(0, m, chunk):
            k = min(chunk, m - s)
            parts.append(draw(ucdf, k) * n_items + draw(icdf, k))
        keys = torch.unique(torch.cat(parts))
        del parts
        if keys.numel() > n_edges:
            # drop a uniformly random surplus so the law is not biased towards low ids
            drop = torch.randperm(keys.numel(), device=device, generator=g)[: keys.numel() - n_edges]
            mask = torch.ones(keys.numel(), dtype=torch.bool, device=device)
            mask[drop] = False
            keys = keys[mask]
    return keys // n_items, keys % n_items


def synth_pairs_host(n_users: int, n_items: int, n_edges: int, seed: int = 42):
    """numpy twin of :func:`synth_pairs_device` (same law, no split) for CPU-side baselines."""
    rng = np.random.default_rng(seed)
    ucdf = _power_law_cdf(n_users, 0.6)
    icdf = _power_law_cdf(n_items, 0.8)
    keys = np.empty(0, dtype=np.int64)
    while len(keys) < n_edges:
        need = n_edges - len(keys)
        m = int(need * 1.25) + 4096
        keys = np.unique(np.concatenate([keys, _draw(rng, ucdf, m) * n_items + _draw(rng, icdf, m)]))
        if len(keys) > n_edges:
            keep = np.sort(rng.permutation(len(keys))[:n_edges])
            keys = keys[keep]
    return keys // n_items, keys % n_items
