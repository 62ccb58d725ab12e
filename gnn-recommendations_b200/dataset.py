"""Interaction store the hot path reads — the boundary attributes of the reference's
``RecommendationDataset`` (src/data/dataset.py:39): ``train_data`` / ``valid_data`` /
``test_data`` (DataFrames with userId, itemId), ``n_users``, ``n_items``,
``get_torch_adjacency(normalized)`` and ``processed_data_path``.  The pandas ETL of the reference
(loaders, k-core filtering, raw files) is out of scope; this class holds already-split pairs and
builds Â on the device with the CUDA graph builder."""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

from .graph_builder import NormAdjCSR
from .synthetic import synth_split


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
        self.stats = {}
        self._adj = {}

    @classmethod
    def synthetic(cls, shape: str, seed: int = 42, device="cuda") -> "InteractionDataset":
        sp = synth_split(shape, seed)
        return cls(sp["train"], sp["valid"], sp["test"], sp["n_users"], sp["n_items"], device=device, name=shape)

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
