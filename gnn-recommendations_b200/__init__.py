"""gnn-recommendations hot path, B200-native (sm_100a).  See DESIGN.md.

Drop-in surface (same names and signatures as the reference's ``src`` package):
``LightGCN`` and friends, ``Trainer``, ``Evaluator``, ``BPRLoss`` and the graph-builder
functions.  Everything numerical runs in libgr_b200.so (hand-written CUDA behind the C ABI in
include/gr_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .base import BaseRecommender  # noqa: F401
from .graph_builder import (NormAdjCSR, as_csr, build_bipartite_graph, convert_to_torch_sparse,  # noqa: F401
                            normalize_adjacency_matrix, sm_copy)
from .lightgcn import LightGCN, lightgcn_propagate  # noqa: F401
from .ngcf import NGCF, NGCFLayer  # noqa: F401
from .gat import GAT, GATLayer  # noqa: F401
from .orthogonal_bundle import BundleConnectionLayer, GroupShuffleLayer, OrthogonalBundleGNN  # noqa: F401
from .ultragcn import UltraGCN  # noqa: F401
from .kgtore import KGTORe  # noqa: F401
from .dataset import InteractionDataset, temporal_split_device  # noqa: F401
from .evaluator import Evaluator, full_rank_topk  # noqa: F401
from .losses import BPRLoss, bpr_fused  # noqa: F401
from .metrics import compute_metrics_from_topk, topk_metrics_device  # noqa: F401
from .optim import fused_clip_adam_step, fused_clip_adam_supported  # noqa: F401
from .sampler import BprSampler  # noqa: F401
from .trainer import Trainer  # noqa: F401

# name -> class, as scripts/run_all_experiments.py:38-45 registers them (the plug-in point of the
# reference's driver: create_model() looks the class up here and filters YAML kwargs by signature)
MODEL_REGISTRY = {"lightgcn": LightGCN, "ngcf": NGCF, "gat": GAT, "ultragcn": UltraGCN, "kgtore": KGTORe,
                  "orthogonal_bundle": OrthogonalBundleGNN}

__version__ = "0.1.0"
