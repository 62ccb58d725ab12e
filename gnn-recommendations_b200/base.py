"""BaseRecommender — the abstract model API of the reference (src/models/base.py:15-122),
kept as the drop-in surface: same constructor, same method names, same exceptions."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn


class BaseRecommender(nn.Module):
    def __init__(self, n_users: int, n_items: int, embedding_dim: int):
        super().__init__()
        self.n_users = n_users
        self.n_items = n_items
        self.embedding_dim = embedding_dim

    def forward(self, *args, **kwargs):
        raise NotImplementedError(f"forward() must be implemented by {self.__class__.__name__}")

    def predict(self, users: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(f"predict() must be implemented by {self.__class__.__name__}")

    def get_all_embeddings(self) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError(f"get_all_embeddings() must be implemented by {self.__class__.__name__}")

    def get_parameters_count(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def reset_parameters(self):
        # base.py:108-122
        for module in self.modules():
            if isinstance(module, nn.Embedding):
                nn.init.normal_(module.weight, mean=0.0, std=0.01)
            elif isinstance(module, nn.Linear):
                nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.zeros_(module.bias)

    # shared by every GNN subclass (lightgcn.py:106-133 and the identical bodies in ngcf/gat)
    def _predict_pairs(self, users, items, user_emb, item_emb):
        return (user_emb[users] * item_emb[items]).sum(dim=1)
