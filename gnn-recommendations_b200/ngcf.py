"""NGCF — drop-in for src/models/baselines/ngcf.py (NGCFLayer :16-86, NGCF :89-242).

Per layer: n = Â x (SpMM kernel), out = LeakyReLU_0.2(W1 n + b1 + W2 (x * n) + b2) (one rowmap
kernel: both 64x64 maps, the bi-interaction product, biases, the activation and the train-mode
dropout fused; backward = gr_rowmap_bwd + the SpMM on Â^T);
the L+1 layer outputs are concatenated (final width 64 * (L+1))."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .base import BaseRecommender
from .graph_builder import as_csr
from .layer_ops import ACT_LEAKY, new_dropout_seed, rowmap, spmm


class NGCFLayer(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, dropout: float = 0.0):
        super().__init__()
        self.W1 = nn.Linear(in_dim, out_dim, bias=True)
        self.W2 = nn.Linear(in_dim, out_dim, bias=True)
        self.dropout = nn.Dropout(dropout)
        self.activation = nn.LeakyReLU(negative_slope=0.2)

    def forward(self, x: torch.Tensor, adj_matrix) -> torch.Tensor:
        csr = as_csr(adj_matrix)
        n = spmm(csr, x)
        p = self.dropout.p if (self.training and self.dropout.p > 0) else 0.0      # ngcf.py:86, fused
        return rowmap(n, self.W1.weight.t(), self.W1.bias, x, n, self.W2.weight.t(), self.W2.bias,
                      act=ACT_LEAKY, slope=0.2, drop_p=p, drop_seed=new_dropout_seed() if p else 0)


class NGCF(BaseRecommender):
    _graph_safe = True      # the training step can be captured in a CUDA graph (no host-seeded torch RNG ops)

    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 64, layer_sizes: Optional[List[int]] = None,
                 dropout: float = 0.1, init_scale: float = 0.01):
        super().__init__(n_users, n_items, embedding_dim)
        if layer_sizes is None:
            layer_sizes = [64, 64, 64]
        self.layer_sizes = layer_sizes
        self.n_layers = len(layer_sizes)
        self.dropout = dropout
        self.init_scale = init_scale
        self.user_embedding = nn.Embedding(n_users, embedding_dim)
        self.item_embedding = nn.Embedding(n_items, embedding_dim)
        self.layers = nn.ModuleList()
        in_dim = embedding_dim
        for out_dim in layer_sizes:
            self.layers.append(NGCFLayer(in_dim, out_dim, dropout))
            in_dim = out_dim
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.user_embedding.weight, mean=0.0, std=self.init_scale)
        nn.init.normal_(self.item_embedding.weight, mean=0.0, std=self.init_scale)
        for layer in self.layers:
            nn.init.xavier_uniform_(layer.W1.weight)
            nn.init.xavier_uniform_(layer.W2.weight)
            if layer.W1.bias is not None:
                nn.init.zeros_(layer.W1.bias)
            if layer.W2.bias is not None:
                nn.init.zeros_(layer.W2.bias)

    def propagate(self, adj_matrix) -> torch.Tensor:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для NGCF")
        csr = as_csr(adj_matrix)
        x = torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)
        outs = [x]
        for layer in self.layers:
            x = layer(x, csr)
            outs.append(x)
        return torch.cat(outs, dim=1)

    def forward(self, adj_matrix) -> Tuple[torch.Tensor, torch.Tensor]:
        x = self.propagate(adj_matrix)
        return tuple(torch.split(x, [self.n_users, self.n_items], dim=0))

    def predict(self, users, items, adj_matrix=None) -> torch.Tensor:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для NGCF")
        ue, ie = self.get_all_embeddings(adj_matrix)
        return self._predict_pairs(users, items, ue, ie)

    def get_all_embeddings(self, adj_matrix=None) -> Tuple[torch.Tensor, torch.Tensor]:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для NGCF")
        return self.forward(adj_matrix)
