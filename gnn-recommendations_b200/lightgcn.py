"""LightGCN — drop-in for src/models/baselines/lightgcn.py (:20-183) on the sm_100a SpMM.

forward = L launches of gr_spmm_csr_f32; the layer mean (``torch.mean(torch.stack(..))``,
lightgcn.py:94-95) is fused into the SpMM epilogue as a running sum and a final division
by L+1, evaluated in the reference's order ((x0+x1)+x2)+... so the propagated embeddings
are bit-identical to the reference's CPU result.  backward = L more SpMM launches in Horner
form:  h <- g;  L x: h <- g + Â^T h;  dL/dE0 = h / (L+1).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .base import BaseRecommender
from .graph_builder import NormAdjCSR, as_csr


def lightgcn_propagate(csr: NormAdjCSR, x0: torch.Tensor, n_layers: int) -> torch.Tensor:
    """mean_{l=0..L} Â^l x0, no autograd."""
    if n_layers == 0:
        return x0.clone()
    n, d = x0.shape
    out = torch.empty_like(x0)
    if n_layers == 1:
        csr.spmm(x0, addend=x0, out=out, scale=2.0, scale_mode=_lib.GR_SCALE_DIV, want_y=False)
        return out
    ya = torch.empty_like(x0)
    yb = torch.empty_like(x0) if n_layers > 2 else None
    acc = torch.empty_like(x0)
    csr.spmm(x0, y=ya, addend=x0, out=acc)                       # x1, acc = x0 + x1
    cur, nxt = ya, yb
    for _ in range(n_layers - 2):
        csr.spmm(cur, y=nxt, addend=acc, out=acc)                # x_{l+1}, acc += x_{l+1}
        cur, nxt = nxt, cur
    csr.spmm(cur, addend=acc, out=out, scale=float(n_layers + 1), scale_mode=_lib.GR_SCALE_DIV, want_y=False)
    return out


class _LightGCNPropagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, csr, n_layers):
        ctx.csr, ctx.n_layers = csr, n_layers
        return lightgcn_propagate(csr, x0.contiguous(), n_layers)

    @staticmethod
    def backward(ctx, g):
        csr_t, L = ctx.csr.transpose(), ctx.n_layers
        g = g.contiguous()
        if L == 0:
            return g, None, None
        h = g
        buf = [torch.empty_like(g), torch.empty_like(g)]
        for l in range(L):
            last = l == L - 1
            o = buf[l & 1]
            csr_t.spmm(h, addend=g, out=o, scale=float(L + 1),
                       scale_mode=_lib.GR_SCALE_DIV if last else _lib.GR_SCALE_NONE, want_y=False)
            h = o
        return h, None, None


class LightGCN(BaseRecommender):
    _graph_safe = True      # the training step can be captured in a CUDA graph (no host-seeded torch RNG ops)

    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 64, n_layers: int = 3,
                 init_scale: float = 0.01):
        super().__init__(n_users, n_items, embedding_dim)
        self.n_layers = n_layers
        self.init_scale = init_scale
        # same construction order as lightgcn.py:45-49 so that a given torch.manual_seed
        # yields the same initial parameters
        self.user_embedding = nn.Embedding(n_users, embedding_dim)
        self.item_embedding = nn.Embedding(n_items, embedding_dim)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.user_embedding.weight, mean=0.0, std=self.init_scale)
        nn.init.normal_(self.item_embedding.weight, mean=0.0, std=self.init_scale)

    def _x0(self) -> torch.Tensor:
        return torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)

    def propagate(self, adj_matrix) -> torch.Tensor:
        """The final [N, d] embedding matrix (users first) before the user/item split — what
        the fused BPR kernel and the sharded evaluators consume."""
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для LightGCN")
        return _LightGCNPropagate.apply(self._x0(), as_csr(adj_matrix), self.n_layers)

    def forward(self, adj_matrix) -> Tuple[torch.Tensor, torch.Tensor]:
        x = self.propagate(adj_matrix)
        user_emb, item_emb = torch.split(x, [self.n_users, self.n_items], dim=0)
        return user_emb, item_emb

    def predict(self, users: torch.Tensor, items: torch.Tensor, adj_matrix=None) -> torch.Tensor:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для LightGCN")
        user_emb, item_emb = self.get_all_embeddings(adj_matrix)
        return self._predict_pairs(users, items, user_emb, item_emb)

    def get_all_embeddings(self, adj_matrix=None) -> Tuple[torch.Tensor, torch.Tensor]:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для LightGCN")
        return self.forward(adj_matrix)

    def get_layer_embeddings(self, adj_matrix) -> List[torch.Tensor]:
        """lightgcn.py:153-183 (analysis helper, no autograd)."""
        csr = as_csr(adj_matrix)
        with torch.no_grad():
            x = self._x0()
            outs = [x.clone()]
            for _ in range(self.n_layers):
                x, _ = csr.spmm(x)
                outs.append(x.clone())
        return outs
