"""Interaction store the hot path reads — the boundary attributes of the reference's
``RecommendationDataset`` (src/data/dataset.py:39): ``train_data`` / ``valid_data`` /
``test_data`` (DataFrames with userId, itemId), ``n_users``, ``n_items``,
``get_torch_adjacency(normalized)`` and ``processed_data_path``.  The pandas ETL of the reference
(loaders, k-core filtering, raw files) is out of scope; this class holds already-split pairs and
builds Â on the device with the CUDA graph builder.  SURVEY §8f-3: the per-user temporal split
(dataset.py:327-357) runs on the device (``temporal_split_device`` / ``from_interactions``) and the
on-disk contract of the reference (``train.txt`` / ``valid.txt`` / ``test.txt`` TSV, ``stats.json``,
``adj_matrix.npz`` / ``norm_adj_matrix.npz``; dataset.py:366-394, 461-466, 493-522) is read and written
by ``save_processed`` / ``load_processed``."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd
import torch

from .graph_builder import NormAdjCSR
from .synthetic import synth_split


def temporal_split_device(user, item, timestamp, n_users: int, device="cuda"):
    """dataset.py:327-357 on the device (gr_temporal_split): rows ordered by (userId, timestamp); per
    user the last row -> test (>= 2 rows), the second-last -> valid (>= 3 rows), the rest -> train.
    Returns {"train"|"valid"|"test": (user, item)} int64 CUDA tensors; train keeps the sorted order."""
    from ._lib import check, lib, ptr, stream_ptr

    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("temporal_split_device needs a CUDA device (no CPU fallback)")
    u = torch.as_tensor(user, dtype=torch.int64).to(dev).contiguous()
    i = torch.as_tensor(item, dtype=torch.int64).to(dev).contiguous()
    t = torch.as_tensor(timestamp, dtype=torch.int64).to(dev).contiguous()
    n = int(u.numel())
    if int(i.numel()) != n or int(t.numel()) != n:
        raise ValueError("user, item and timestamp must have the same length")
    empty = torch.zeros(0, dtype=torch.int64, device=dev)
    if n == 0:
        return {"train": (empty, empty), "valid": (empty, empty), "test": (empty, empty)}
    ts_min, ts_max = int(t.min()), int(t.max())
    m = min(n, int(n_users))
    tu, ti = torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.int64, device=dev)
    vu, vi = torch.empty(m, dtype=torch.int64, device=dev), torch.empty(m, dtype=torch.int64, device=dev)
    su, si = torch.empty(m, dtype=torch.int64, device=dev), torch.empty(m, dtype=torch.int64, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    l = lib()
    ws_bytes = l.gr_temporal_split_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(l.gr_temporal_split(ptr(u), ptr(i), ptr(t), n, int(n_users), ts_min, ts_max, ptr(tu), ptr(ti), ptr(vu), ptr(vi),
                              ptr(su), ptr(si), ptr(counts), ptr(status), ptr(ws), ws_bytes, stream_ptr()),
          "gr_temporal_split")
    c = counts.tolist()
    if int(status.item()) != 0:
        raise ValueError("temporal split: a user id is outside [0, n_users)")
    return {"train": (tu[:c[0]], ti[:c[0]]), "valid": (vu[:c[1]], vi[:c[1]]), "test": (su[:c[2]], si[:c[2]])}


class InteractionDataset:
    def __init__(self, train, valid, test, n_users: int, n_items: int, device="cuda", name: str = "synthetic",
                 normalization: str = "symmetric", self_loop: bool = False, root_dir: str = "."):
        def frame(p):
            u, i = p
            return pd.DataFrame({"userId": np.asarray(u, dtype=np.int64), "itemId": np.asarray(i, dtype=np.int64)})
        self.name = name
        self.train_data, self.valid_data, self.test_data = frame(train), frame(valid), frame(test)
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.device = device
        self.normalization, self.self_loop = normalization, self_loop
        self.processed_data_path = Path(root_dir) / "data" / "processed" / name
        self.graphs_path = Path(root_dir) / "data" / "graphs" / name              # dataset.py:98
        self.stats = {}
        self._adj = {}

    @classmethod
    def synthetic(cls, shape: str, seed: int = 42, device="cuda") -> "InteractionDataset":
        sp = synth_split(shape, seed)
        return cls(sp["train"], sp["valid"], sp["test"], sp["n_users"], sp["n_items"], device=device, name=shape)

    @classmethod
    def from_interactions(cls, user, item, timestamp, n_users: int, n_items: int, device="cuda", **kw) -> "InteractionDataset":
        """Already-preprocessed interactions (contiguous ids + timestamps) -> device temporal split."""
        sp = temporal_split_device(user, item, timestamp, n_users, device=device)
        host = {k: (a.cpu().numpy(), b.cpu().numpy()) for k, (a, b) in sp.items()}
        ds = cls(host["train"], host["valid"], host["test"], n_users, n_items, device=device, **kw)
        ds.stats = {"n_interactions": int(len(np.asarray(user)))}
        return ds

    # ------------------------------------------------------------------ on-disk contract of the reference
    def save_processed(self, save_graphs: bool = True) -> None:
        """dataset.py:366-394 (`_save_split_data`) + :461-466 (graphs as scipy npz)."""
        self.processed_data_path.mkdir(parents=True, exist_ok=True)
        for name, df in (("train", self.train_data), ("valid", self.valid_data), ("test", self.test_data)):
            df[["userId", "itemId"]].to_csv(self.processed_data_path / f"{name}.txt", sep="\t", index=False, header=False)
        n_inter = self.stats.get("n_interactions", len(self.train_data) + len(self.valid_data) + len(self.test_data))
        with open(self.processed_data_path / "stats.json", "w", encoding="utf-8") as f:
            json.dump({"n_users": self.n_users, "n_items": self.n_items, "n_interactions": int(n_inter),
                       "train_size": len(self.train_data), "valid_size": len(self.valid_data),
                       "test_size": len(self.test_data),
                       **{k: v for k, v in self.stats.items() if k != "n_interactions"}}, f, indent=2)
        if save_graphs:
            import scipy.sparse as sp

            self.graphs_path.mkdir(parents=True, exist_ok=True)
            sp.save_npz(str(self.graphs_path / "adj_matrix.npz"), self.build_graph(normalize=False).to_scipy())
            sp.save_npz(str(self.graphs_path / "norm_adj_matrix.npz"),
                        self.build_graph(normalize=True, normalization_type=self.normalization).to_scipy())

    @classmethod
    def load_processed(cls, name: str, root_dir: str = ".", device="cuda", **kw) -> "InteractionDataset":
        """dataset.py:493-522 (`load_processed_data`): the three TSVs + stats.json written by either side."""
        base = Path(root_dir) / "data" / "processed" / name
        files = [base / f for f in ("train.txt", "valid.txt", "test.txt", "stats.json")]
        if not all(f.exists() for f in files):
            raise FileNotFoundError("Обработанные данные не найдены. Запустите preprocess() и split()")
        frames = [pd.read_csv(f, sep="\t", header=None, names=["userId", "itemId"]) for f in files[:3]]
        with open(files[3], "r", encoding="utf-8") as f:
            stats = json.load(f)
        pairs = [(d["userId"].to_numpy(dtype=np.int64), d["itemId"].to_numpy(dtype=np.int64)) for d in frames]
        ds = cls(pairs[0], pairs[1], pairs[2], stats["n_users"], stats["n_items"], device=device, name=name,
                 root_dir=root_dir, **kw)
        ds.stats = stats
        return ds

    def train_pairs(self):
        return (self.train_data["userId"].to_numpy(dtype=np.int64), self.train_data["itemId"].to_numpy(dtype=np.int64))

    def build_graph(self, normalize: Optional[bool] = True, self_loop: Optional[bool] = None,
                    normalization_type: str = "symmetric") -> NormAdjCSR:
        """dataset.py:415-470 (without the npz side effects)."""
        mode = normalization_type if normalize else "none"
        sl = self.self_loop if self_loop is None else self_loop
        key = (mode, sl)
        if key not in self._adj:
            u, i = self.train_pairs()
            self._adj[key] = NormAdjCSR.from_pairs(u, i, self.n_users, self.n_items, normalization=mode,
                                                   self_loop=sl, device=self.device)
        return self._adj[key]

    def get_torch_adjacency(self, normalized: bool = True) -> NormAdjCSR:
        """dataset.py:472-491.  Returns the device-resident CSR (accepted by every model in this
        package; ``.to(device)`` is a no-op) instead of a host COO tensor."""
        return self.build_graph(normalize=normalized, normalization_type=self.normalization)
