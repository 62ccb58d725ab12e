"""KGTORe — drop-in for src/models/baselines/kgtore.py (:26-394).

The collaborative part (kgtore.py:303-318: x <- dropout(relu(Linear(Â x))) per layer, mean of the
L+1 layers) runs on the sm_100a kernels (gr_spmm_csr_f32 + gr_rowmap_f32, as NGCF's layers do); the
item-side heads the reference adds on top (tree levels with attention, learned item features, fusion
MLP; kgtore.py:96-168, 320-340) are small dense modules on [n_items, d] and stay torch.nn modules
(library GEMMs).  Constructor statement order equals the reference's, so a same-seed construction
consumes the CPU generator identically.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .base import BaseRecommender
from .graph_builder import as_csr
from .layer_ops import ACT_LEAKY, rowmap, spmm


class KGEmbedding(nn.Module):
    """kgtore.py:26-93: learned item features when no knowledge graph is given."""

    def __init__(self, n_items: int, n_entities: Optional[int], n_relations: Optional[int], kg_embedding_dim: int):
        super().__init__()
        self.n_items, self.kg_embedding_dim = n_items, kg_embedding_dim
        if n_entities is not None and n_relations is not None:
            self.use_kg = True
            self.entity_embedding = nn.Embedding(n_entities, kg_embedding_dim)
            self.relation_embedding = nn.Embedding(n_relations, kg_embedding_dim)
            self.item_to_entity = nn.Embedding(n_items, 1)
        else:
            self.use_kg = False
            self.item_features = nn.Embedding(n_items, kg_embedding_dim)

    def forward(self, item_ids: torch.Tensor, kg_data: Optional[Dict] = None):
        if self.use_kg and kg_data is not None:
            return None                      # the reference's KG branch is an unimplemented `pass` (kgtore.py:86-88)
        return self.item_features(item_ids)


class TreeStructure(nn.Module):
    """kgtore.py:96-168: tree_depth x (Linear + ReLU), levels combined by a softmax attention."""

    def __init__(self, embedding_dim: int, tree_depth: int = 3, n_branches: int = 4):
        super().__init__()
        self.tree_depth, self.n_branches = tree_depth, n_branches
        self.tree_layers = nn.ModuleList([nn.Linear(embedding_dim, embedding_dim) for _ in range(tree_depth)])
        self.level_attention = nn.Linear(embedding_dim, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        levels, current = [], x
        for layer in self.tree_layers:
            current = F.relu(layer(current))
            levels.append(current)
        stack = torch.stack(levels, dim=1)                         # [n, depth, d]
        weights = F.softmax(self.level_attention(stack), dim=1)    # [n, depth, 1]
        return (stack * weights).sum(dim=1)


class KGTORe(BaseRecommender):
    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 64, kg_embedding_dim: int = 64,
                 tree_depth: int = 3, n_layers: int = 2, dropout: float = 0.1, use_kg: bool = False,
                 n_entities: Optional[int] = None, n_relations: Optional[int] = None, init_scale: float = 0.01):
        super().__init__(n_users, n_items, embedding_dim)
        self.kg_embedding_dim, self.tree_depth, self.n_layers = kg_embedding_dim, tree_depth, n_layers
        self.dropout, self.use_kg, self.init_scale = dropout, use_kg, init_scale
        self.user_embedding = nn.Embedding(n_users, embedding_dim)
        self.item_embedding = nn.Embedding(n_items, embedding_dim)
        self.kg_embedding = KGEmbedding(n_items, n_entities if use_kg else None, n_relations if use_kg else None,
                                        kg_embedding_dim)
        self.tree_structure = TreeStructure(embedding_dim, tree_depth)
        self.gcn_layers = nn.ModuleList([nn.Linear(embedding_dim, embedding_dim) for _ in range(n_layers)])
        self.fusion = nn.Sequential(nn.Linear(embedding_dim + kg_embedding_dim, embedding_dim), nn.ReLU(),
                                    nn.Dropout(dropout), nn.Linear(embedding_dim, embedding_dim))
        self.dropout_layer = nn.Dropout(dropout)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.user_embedding.weight, mean=0.0, std=self.init_scale)
        nn.init.normal_(self.item_embedding.weight, mean=0.0, std=self.init_scale)
        if hasattr(self.kg_embedding, "item_features"):
            nn.init.normal_(self.kg_embedding.item_features.weight, mean=0.0, std=self.init_scale)
        for layer in self.gcn_layers:
            nn.init.xavier_uniform_(layer.weight)
            if layer.bias is not None:
                nn.init.zeros_(layer.bias)

    def forward(self, adj_matrix, kg_data: Optional[Dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для KGTORe")
        csr = as_csr(adj_matrix)
        x = torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)
        layers = [x]
        for layer in self.gcn_layers:
            x = rowmap(spmm(csr, x), layer.weight.t(), layer.bias, act=ACT_LEAKY, slope=0.0)   # relu(Linear(Â x))
            x = self.dropout_layer(x)
            layers.append(x)
        cf = torch.stack(layers, dim=0).mean(dim=0)
        user_cf, item_cf = cf[:self.n_users], cf[self.n_users:]
        item_ids = torch.arange(self.n_items, device=item_cf.device)
        item_kg = self.kg_embedding(item_ids, kg_data)
        item_tree = self.tree_structure(item_cf)
        item_final = self.fusion(torch.cat([item_tree, item_kg], dim=1))
        return user_cf, item_final

    def predict(self, users: torch.Tensor, items: torch.Tensor, adj_matrix=None, kg_data: Optional[Dict] = None):
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для KGTORe")
        user_emb, item_emb = self.get_all_embeddings(adj_matrix, kg_data)
        return (user_emb[users] * item_emb[items]).sum(dim=1)

    def get_all_embeddings(self, adj_matrix=None, kg_data: Optional[Dict] = None):
        if adj_matrix is None:
            raise ValueError("adj_matrix должен быть передан для KGTORe")
        return self.forward(adj_matrix, kg_data)
