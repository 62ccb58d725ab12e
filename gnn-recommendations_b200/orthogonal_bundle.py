"""OrthogonalBundleGNN (Group-and-Shuffle orthogonal GCN) — drop-in for
src/models/orthogonal_bundle/{model.py:29, group_shuffle_layer.py:12, bundle_layer.py:9}.

Per layer the reference computes  c = Â x ; t = c @ W_conn ; g = (t @ W_orth)[:, perm] ;
x = (1-a) g + a x0.  The two 64x64 maps and the column permutation compose into one matrix
M = W_conn (W_orth[:, perm]) built on the device by gr_gs_compose (block matrix exponentials +
composition in two launches for all layers; backward gr_gs_compose_bwd through the adjoint Frechet
derivative of exp — the reference runs 48 matrix_exp calls per forward), so a layer is the SpMM kernel
plus the rowmap kernel (dense map + residual fused; backward gr_rowmap_bwd).  With GR_GS_FUSED=1 a layer is
ONE kernel instead: gr_spmm_csr_map_f32 applies M and the residual in the SpMM epilogue
(layer_ops.gs_propagate, forward and backward; d <= 64, no layer dropout) — implemented and parity-tested,
but measured slower than the two-kernel layer on B200 (see _gs_fused_enabled), hence opt-in.  The layer
outputs are combined with softmax(layer_weights).  State-dict keys, constructor signature and
RNG order match the reference.  The edge-list mode (use_edge_index=True, off by default, model.py:64;
parallel_transport.py:5-52) runs on the same kernels over a CSR of the edge list (_layers_edge_index)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .base import BaseRecommender
from .graph_builder import as_csr
from .graph_builder import NormAdjCSR
from .layer_ops import (gs_compose, gs_propagate, gs_propagate_supported, layer_combine, new_dropout_seed, rowmap,
                        spmm, spmm_map)


def _gs_fused_enabled() -> bool:
    """GR_GS_FUSED=1 selects the fused layer kernel (gr_spmm_csr_map_f32: sparse product, composed map and
    residual in one launch, layer_ops.gs_propagate).  Default: the layer-wise path (SpMM kernel + rowmap kernel).
    Measured on B200 at the Gowalla shape the fused kernel is SLOWER (forward 0.82 ms against 0.57 ms): the
    register-tiled dense map needs ~45 more registers than the gather loop, and at 126 registers per thread only
    one streaming CTA fits beside the long-row kernel's CTA on an SM (two waves instead of one); capped at 80
    registers ptxas spills the gathered rows in the hot loop (Group-and-Shuffle)."""
    import os
    return os.environ.get("GR_GS_FUSED", "0") == "1"


def _block_orthogonal(skew_params) -> torch.Tensor:
    return torch.block_diag(*[torch.matrix_exp(p - p.transpose(-2, -1)) for p in skew_params])


def _orth_metrics(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    eye = torch.eye(w.shape[0], device=w.device, dtype=w.dtype)
    diff = w.T @ w - eye
    return torch.norm(diff, p="fro"), diff.abs().max()


class GroupShuffleLayer(nn.Module):
    """group_shuffle_layer.py:12-190: block-diagonal exp(skew) ("group") + fixed permutation ("shuffle")."""

    def __init__(self, dim: int, block_size: int, init_scale: float = 0.01):
        super().__init__()
        if dim % block_size != 0:
            raise ValueError(f"dim ({dim}) must be divisible by block_size ({block_size})")
        self.dim, self.block_size, self.n_blocks = dim, block_size, dim // block_size
        self.skew_params = nn.ParameterList([nn.Parameter(torch.randn(block_size, block_size) * init_scale)
                                             for _ in range(self.n_blocks)])
        self.register_buffer("perm", torch.randperm(dim))

    def _build_orthogonal_matrix(self) -> torch.Tensor:
        return _block_orthogonal(self.skew_params)

    def matrix(self) -> torch.Tensor:
        """The map the layer applies: x -> (x @ W_orth)[:, perm] == x @ W_orth[:, perm]."""
        return self._build_orthogonal_matrix()[:, self.perm]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return rowmap(x, self.matrix())

    def get_orthogonality_error(self) -> torch.Tensor:
        return _orth_metrics(self._build_orthogonal_matrix())[0]

    def get_orthogonality_metrics(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return _orth_metrics(self._build_orthogonal_matrix())

    def reset_parameters(self):
        for p in self.skew_params:
            nn.init.normal_(p, mean=0.0, std=0.01)


class BundleConnectionLayer(nn.Module):
    """bundle_layer.py:9-88: the shared connection matrix W = blockdiag(exp(skew))[:, shuffle_perm]."""

    def __init__(self, embedding_dim, block_size, n_blocks=None):
        super().__init__()
        self.embedding_dim, self.block_size = embedding_dim, block_size
        self.n_blocks = n_blocks or (embedding_dim // block_size)
        self.skew_params = nn.ParameterList([nn.Parameter(torch.randn(block_size, block_size) * 0.01)
                                             for _ in range(self.n_blocks)])
        self.register_buffer("shuffle_perm", torch.randperm(embedding_dim))

    def forward(self, edge_index=None) -> torch.Tensor:
        return _block_orthogonal(self.skew_params)[:, self.shuffle_perm]

    def get_connection_matrix_for_edge(self, src_node, dst_node):
        return self.forward()

    def get_orthogonality_metrics(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return _orth_metrics(self.forward())


class OrthogonalBundleGNN(BaseRecommender):
    _graph_safe = True      # the training step can be captured in a CUDA graph (no host-seeded torch RNG ops)

    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 64, n_layers: int = 3, block_size: int = 8,
                 residual_alpha: float = 0.1, dropout: float = 0.0, init_scale: float = 0.01,
                 use_parallel_transport: bool = True, use_edge_index: bool = False):
        super().__init__(n_users, n_items, embedding_dim)
        if embedding_dim % block_size != 0:
            raise ValueError(f"embedding_dim ({embedding_dim}) must be divisible by block_size ({block_size})")
        self.n_layers, self.block_size, self.residual_alpha = n_layers, block_size, residual_alpha
        self.dropout, self.use_parallel_transport, self.use_edge_index = dropout, use_parallel_transport, use_edge_index
        # RNG order of model.py:96-112
        self.user_embedding = nn.Embedding(n_users, embedding_dim)
        self.item_embedding = nn.Embedding(n_items, embedding_dim)
        nn.init.normal_(self.user_embedding.weight, std=0.01)
        nn.init.normal_(self.item_embedding.weight, std=0.01)
        if use_parallel_transport:
            self.connection_layers = nn.ModuleList([BundleConnectionLayer(embedding_dim, block_size)
                                                    for _ in range(n_layers)])
        self.local_transform_layers = nn.ModuleList([GroupShuffleLayer(embedding_dim, block_size, init_scale)
                                                     for _ in range(n_layers)])
        self.dropout_layer = nn.Dropout(dropout) if dropout > 0 else None
        self.layer_weights = nn.Parameter(torch.ones(n_layers + 1))

    def _layer_matrices(self) -> torch.Tensor:
        """[L, d, d]: M_l = W_conn,l @ W_orth,l[:, perm_l] for every layer in two launches (gr_gs_compose: block
        matrix exponentials + composition; differentiable w.r.t. all skew parameters)."""
        sets = []
        if self.use_parallel_transport:
            sets.append([p for layer in self.connection_layers for p in layer.skew_params])
        sets.append([p for layer in self.local_transform_layers for p in layer.skew_params])
        nb = self.embedding_dim // self.block_size
        skew = torch.stack([p for group in sets for p in group]).view(
            len(sets), self.n_layers, nb, self.block_size, self.block_size).transpose(0, 1)
        perm_c = (torch.stack([layer.shuffle_perm for layer in self.connection_layers])
                  if self.use_parallel_transport else None)
        perm_g = torch.stack([layer.perm for layer in self.local_transform_layers])
        return gs_compose(skew, perm_c, perm_g)

    def _edge_csr(self, edge_index: torch.Tensor) -> NormAdjCSR:
        """The edge list of the edge-list mode as a CSR: row = destination, one entry of value 1 per edge, in edge
        order (a stable sort by destination), so the SpMM's storage-order chain adds the source rows in the order
        ``index_add_`` does (parallel_transport.py:49-50, model.py:218-222).  The last one is cached, keyed by
        CONTENT (shape + two position-weighted checksums: a new tensor at a recycled address must not hit)."""
        dev = self.user_embedding.weight.device
        ei = edge_index.to(dev)
        n = self.n_users + self.n_items
        if ei.dim() != 2 or ei.shape[0] != 2:
            raise ValueError("edge_index must have shape [2, num_edges]")
        if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= n):
            raise ValueError("edge_index holds node ids outside [0, n_users + n_items)")
        w = torch.arange(1, ei.shape[1] + 1, device=dev, dtype=torch.int64)
        key = (tuple(ei.shape), int((ei[0] * w).sum()), int((ei[1] * w).sum()))
        hit = getattr(self, "_edge_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        order = torch.sort(ei[1], stable=True).indices                      # plumbing: once per edge list
        coo = torch.sparse_coo_tensor(torch.stack([ei[1][order], ei[0][order]]),
                                      torch.ones(ei.shape[1], dtype=torch.float32, device=dev), (n, n))
        csr = NormAdjCSR.from_torch_coo(coo)
        self._edge_cache = (key, csr)
        return csr

    def _layers_edge_index(self, edge_index, residual: bool) -> List[torch.Tensor]:
        """use_edge_index=True (model.py:159-181 with parallel_transport.py:5-52): per layer
        x_j <- sum over edges (i -> j) of W_conn x_i, i.e. (E x) W_conn^T with E the edge-count matrix — the fused
        SpMM + map kernel with the TRANSPOSED connection matrix — then the local Group-and-Shuffle map and the
        residual (rowmap kernel).  Without parallel transport the first step is the plain sum E x (model.py:218-222)."""
        if edge_index is None:
            raise ValueError("edge_index must be provided when use_edge_index=True")
        csr = self._edge_csr(edge_index)
        x0 = torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)
        nb, bs, L = self.embedding_dim // self.block_size, self.block_size, self.n_layers
        loc = gs_compose(torch.stack([p for layer in self.local_transform_layers for p in layer.skew_params])
                         .view(L, 1, nb, bs, bs), None, torch.stack([layer.perm for layer in self.local_transform_layers]))
        conn = None
        if self.use_parallel_transport:
            conn = gs_compose(torch.stack([p for layer in self.connection_layers for p in layer.skew_params])
                              .view(L, 1, nb, bs, bs), None,
                              torch.stack([layer.shuffle_perm for layer in self.connection_layers]))
        x, outs, a = x0, [x0], self.residual_alpha
        fused = csr.supports_map(self.embedding_dim)
        for l in range(L):
            if conn is None:
                t = spmm(csr, x)
            elif fused:
                t = spmm_map(csr, x, conn[l], transposed=True)
            else:
                t = rowmap(spmm(csr, x), conn[l].t().contiguous())
            if residual:
                p = self.dropout if (self.training and self.dropout_layer is not None) else 0.0
                x = rowmap(t, loc[l], resid=x0, alpha=1.0 - a, beta=a, drop_p=p, drop_seed=new_dropout_seed() if p else 0)
            else:
                x = rowmap(t, loc[l])
            outs.append(x)
        return outs

    def _layers(self, adj_matrix, residual: bool, edge_index=None) -> List[torch.Tensor]:
        if self.use_edge_index:
            return self._layers_edge_index(edge_index, residual)
        if adj_matrix is None:
            raise ValueError("adj_matrix must be provided when use_edge_index=False")
        csr = as_csr(adj_matrix)
        x0 = torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)
        x, outs = x0, [x0]
        a = self.residual_alpha
        ms = self._layer_matrices()
        for l in range(self.n_layers):
            c = spmm(csr, x)
            if residual:
                p = self.dropout if (self.training and self.dropout_layer is not None) else 0.0   # model.py:198-199
                x = rowmap(c, ms[l], resid=x0, alpha=1.0 - a, beta=a, drop_p=p,
                           drop_seed=new_dropout_seed() if p else 0)
            else:
                x = rowmap(c, ms[l])
            outs.append(x)
        return outs

    def propagate(self, adj_matrix=None, edge_index=None) -> torch.Tensor:
        w = F.softmax(self.layer_weights, dim=0)
        drop = self.dropout if (self.training and self.dropout_layer is not None) else 0.0
        if not self.use_edge_index and adj_matrix is not None and drop == 0.0 and _gs_fused_enabled():
            csr = as_csr(adj_matrix)
            if gs_propagate_supported(csr, self.embedding_dim, self.n_layers):
                # every layer = ONE kernel: sparse product, composed 64x64 map and residual in the SpMM epilogue
                x0 = torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)
                a = self.residual_alpha
                return gs_propagate(csr, x0, self._layer_matrices(), w, 1.0 - a, a)
        outs = self._layers(adj_matrix, residual=True, edge_index=edge_index)
        return layer_combine(outs, w)                         # sum([w_l * x_l]) of model.py:204-207, one kernel

    def forward(self, adj_matrix=None, edge_index=None) -> Tuple[torch.Tensor, torch.Tensor]:
        x = self.propagate(adj_matrix, edge_index)
        return x[:self.n_users], x[self.n_users:]

    def predict(self, users, items, adj_matrix=None, edge_index=None) -> torch.Tensor:
        ue, ie = self.get_all_embeddings(adj_matrix, edge_index)
        return self._predict_pairs(users, items, ue, ie)

    def get_all_embeddings(self, adj_matrix=None, edge_index=None) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.forward(adj_matrix, edge_index)

    def get_layer_embeddings(self, adj_matrix=None, edge_index=None) -> List[torch.Tensor]:
        """model.py:306-356: per-layer outputs WITHOUT the residual (analysis helper)."""
        with torch.no_grad():
            return [o.clone() for o in self._layers(adj_matrix, residual=False, edge_index=edge_index)]

    def get_orthogonality_errors(self) -> torch.Tensor:
        return torch.stack([l.get_orthogonality_error() for l in self.local_transform_layers])

    def get_orthogonality_metrics(self) -> Dict[str, torch.Tensor]:
        m: Dict[str, torch.Tensor] = {}
        lf, lm = zip(*[l.get_orthogonality_metrics() for l in self.local_transform_layers])
        m["local_fro_mean"], m["local_fro_max"] = torch.stack(lf).mean(), torch.stack(lf).max()
        m["local_max_dev"] = torch.stack(lm).max()
        if self.use_parallel_transport:
            cf, cm = zip(*[l.get_orthogonality_metrics() for l in self.connection_layers])
            m["conn_fro_mean"], m["conn_fro_max"] = torch.stack(cf).mean(), torch.stack(cf).max()
            m["conn_max_dev"] = torch.stack(cm).max()
        return m

    def reset_parameters(self):
        nn.init.normal_(self.user_embedding.weight, mean=0.0, std=0.01)
        nn.init.normal_(self.item_embedding.weight, mean=0.0, std=0.01)
        for layer in self.local_transform_layers:
            layer.reset_parameters()
