"""UltraGCN — drop-in for src/models/baselines/ultragcn.py (:21-247).

The model has no propagation: ``forward`` returns the raw embedding tables (ultragcn.py:75-91), so
under the reference's Trainer its hot path is the fused BPR step and the full-ranking top-K only.
``compute_loss`` / ``compute_constraint_loss`` (ultragcn.py:132-247; not called by the Trainer) are kept
for API parity; the per-user python loop of the constraint term is one batched sparse product.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .base import BaseRecommender


class UltraGCN(BaseRecommender):
    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 64, lambda_1: float = 1.0,
                 lambda_2: float = 1.0, gamma: float = 1e-4, neg_weight: float = 0.5, init_scale: float = 0.01):
        super().__init__(n_users, n_items, embedding_dim)
        self.lambda_1, self.lambda_2, self.gamma = lambda_1, lambda_2, gamma
        self.neg_weight, self.init_scale = neg_weight, init_scale
        self.user_embedding = nn.Embedding(n_users, embedding_dim)      # same RNG order as ultragcn.py:64-68
        self.item_embedding = nn.Embedding(n_items, embedding_dim)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.user_embedding.weight, mean=0.0, std=self.init_scale)
        nn.init.normal_(self.item_embedding.weight, mean=0.0, std=self.init_scale)

    def propagate(self, adj_matrix=None) -> torch.Tensor:
        return torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)

    def forward(self, adj_matrix=None) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.user_embedding.weight, self.item_embedding.weight

    def predict(self, users: torch.Tensor, items: torch.Tensor, adj_matrix=None) -> torch.Tensor:
        return (self.user_embedding(users) * self.item_embedding(items)).sum(dim=1)

    def get_all_embeddings(self, adj_matrix=None) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.forward(adj_matrix)

    def compute_constraint_loss(self, users: torch.Tensor, pos_items: torch.Tensor, adj_matrix: torch.Tensor) -> torch.Tensor:
        """lambda_1 / B * sum_b || E_u[b] - mean_{i in N(u_b)} E_i ||^2 (ultragcn.py:132-188); users
        without neighbours contribute nothing.  adj_matrix: [n_users, n_items] dense or torch sparse."""
        user_emb = self.user_embedding(users)
        rows = adj_matrix.index_select(0, users)
        pattern = (rows.to_dense() if rows.is_sparse else rows) != 0            # [B, n_items]
        deg = pattern.sum(dim=1)
        mean_nb = (pattern.to(user_emb.dtype) @ self.item_embedding.weight) / deg.clamp(min=1).unsqueeze(1).to(user_emb.dtype)
        per_user = ((user_emb - mean_nb) ** 2).sum(dim=1)
        loss = torch.where(deg > 0, per_user, torch.zeros_like(per_user)).sum() / users.size(0)
        return self.lambda_1 * loss

    def compute_loss(self, users: torch.Tensor, pos_items: torch.Tensor, neg_items: torch.Tensor,
                     adj_matrix: Optional[torch.Tensor] = None):
        """BPR + constraint + L2 (ultragcn.py:190-247)."""
        user_emb = self.user_embedding(users)
        pos_emb, neg_emb = self.item_embedding(pos_items), self.item_embedding(neg_items)
        pos_scores, neg_scores = (user_emb * pos_emb).sum(dim=1), (user_emb * neg_emb).sum(dim=1)
        bpr_loss = -torch.log(torch.sigmoid(pos_scores - neg_scores) + 1e-10).mean()
        if adj_matrix is not None:
            constraint_loss = self.compute_constraint_loss(users, pos_items, adj_matrix)
        else:
            constraint_loss = torch.tensor(0.0, device=user_emb.device)
        l2_loss = self.lambda_2 * self.gamma * (torch.norm(user_emb) ** 2 + torch.norm(pos_emb) ** 2
                                                + torch.norm(neg_emb) ** 2) / user_emb.size(0)
        total = bpr_loss + constraint_loss + l2_loss
        return total, {"bpr_loss": bpr_loss.item(), "constraint_loss": float(constraint_loss),
                       "l2_loss": l2_loss.item(), "total_loss": total.item()}
