"""GAT — drop-in for src/models/baselines/gat.py (GATLayer :17-151, GAT :154-344).

The reference materialises dense N x N attention matrices (19 GB per temporary at the Yelp2018
shape); here each layer is a CSR edge-softmax over the adjacency PATTERN: h = x W^T for all heads
(one rowmap kernel), per-node scores, then one pass over each row's neighbours with an online
softmax and the weighted aggregation (gr_gat_aggregate), heads concatenated or averaged, ELU
fused.  In train() mode the reference's dropout on the softmaxed attention weights (gat.py:138) is
applied at the same place, per edge and head, inside the aggregation kernel (mask = counter-based hash
of a seed drawn from torch's CPU generator; the backward kernel re-derives it).  The RNG stream itself
cannot match the reference's dense N x N draw, so element-wise parity is checked in eval() mode and
the train-mode path statistically and against a same-mask torch restatement."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .base import BaseRecommender
from .graph_builder import as_csr
from .layer_ops import gat_layer, layer_combine, new_dropout_seed


class GATLayer(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, n_heads: int = 1, dropout: float = 0.0, alpha: float = 0.2,
                 concat_heads: bool = True):
        super().__init__()
        self.in_dim, self.out_dim, self.n_heads = in_dim, out_dim, n_heads
        self.concat_heads, self.dropout, self.alpha = concat_heads, dropout, alpha
        self.W = nn.ModuleList([nn.Linear(in_dim, out_dim, bias=False) for _ in range(n_heads)])
        self.a_self = nn.ParameterList([nn.Parameter(torch.zeros(size=(out_dim, 1))) for _ in range(n_heads)])
        self.a_neigh = nn.ParameterList([nn.Parameter(torch.zeros(size=(out_dim, 1))) for _ in range(n_heads)])
        self.leakyrelu = nn.LeakyReLU(alpha)
        self.dropout_layer = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, adj_matrix, elu: bool = False) -> torch.Tensor:
        p = self.dropout if (self.training and self.dropout > 0) else 0.0
        return gat_layer(as_csr(adj_matrix), x, [w.weight for w in self.W], list(self.a_self), list(self.a_neigh),
                         self.alpha, self.concat_heads, elu, drop_p=p, drop_seed=new_dropout_seed() if p else 0)


class GAT(BaseRecommender):
    _graph_safe = True      # the training step can be captured in a CUDA graph (no host-seeded torch RNG ops)

    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 64, n_layers: int = 3, n_heads: int = 4,
                 dropout: float = 0.1, alpha: float = 0.2, init_scale: float = 0.01):
        super().__init__(n_users, n_items, embedding_dim)
        self.n_layers, self.n_heads, self.dropout, self.alpha, self.init_scale = n_layers, n_heads, dropout, alpha, init_scale
        self.user_embedding = nn.Embedding(n_users, embedding_dim)
        self.item_embedding = nn.Embedding(n_items, embedding_dim)
        self.layers = nn.ModuleList()
        self.layers.append(GATLayer(embedding_dim, embedding_dim // n_heads, n_heads, dropout, alpha, True))
        for _ in range(n_layers - 2):
            self.layers.append(GATLayer(embedding_dim, embedding_dim // n_heads, n_heads, dropout, alpha, True))
        if n_layers > 1:
            self.layers.append(GATLayer(embedding_dim, embedding_dim, n_heads, dropout, alpha, False))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.user_embedding.weight, mean=0.0, std=self.init_scale)
        nn.init.normal_(self.item_embedding.weight, mean=0.0, std=self.init_scale)
        for layer in self.layers:
            for w in layer.W:
                nn.init.xavier_uniform_(w.weight)
            for a in layer.a_self:
                nn.init.xavier_uniform_(a.data)
            for a in layer.a_neigh:
                nn.init.xavier_uniform_(a.data)

    def propagate(self, adj_matrix) -> torch.Tensor:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для GAT")
        csr = as_csr(adj_matrix)
        x = torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)
        outs = [x]
        for layer in self.layers:
            x = layer(x, csr, elu=True)                       # ELU of gat.py:283 fused into the kernel
            outs.append(x)
        return layer_combine(outs)                            # torch.mean(torch.stack(outs)) of gat.py:287-288

    def forward(self, adj_matrix) -> Tuple[torch.Tensor, torch.Tensor]:
        x = self.propagate(adj_matrix)
        return tuple(torch.split(x, [self.n_users, self.n_items], dim=0))

    def predict(self, users, items, adj_matrix=None) -> torch.Tensor:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для GAT")
        ue, ie = self.get_all_embeddings(adj_matrix)
        return self._predict_pairs(users, items, ue, ie)

    def get_all_embeddings(self, adj_matrix=None) -> Tuple[torch.Tensor, torch.Tensor]:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для GAT")
        return self.forward(adj_matrix)
