"""Generates tests/golden/kgtore.npz from the UNMODIFIED reference KGTORe
(/root/reference/gnn-recommendations/src/models/baselines/kgtore.py) on the tiny synthetic graph:
same-seed parameters, eval-mode forward, and the gradients of a score loss.  Build container only.

    python tests/golden/make_golden_kgtore.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/gnn-recommendations")
sys.path.insert(0, REPO)
from src.models import KGTORe  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

tiny = np.load(os.path.join(HERE, "tiny.npz"))
nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
adj = po.build_norm_adj(tiny["train_u"], tiny["train_i"], nu, ni)
assert np.array_equal(adj["vals"], tiny["adj_val"])           # the oracle's Â is the reference's
A = po.to_torch_coo(adj)
torch.manual_seed(42)
m = KGTORe(nu, ni, embedding_dim=64, kg_embedding_dim=32, tree_depth=3, n_layers=2, dropout=0.1, init_scale=0.1)
out = {}
for k, v in m.state_dict().items():
    out[f"init/{k}"] = v.detach().numpy().copy()
m.eval()
ue, ie = m(A)
out["user_emb"], out["item_emb"] = ue.detach().numpy(), ie.detach().numpy()
rng = np.random.default_rng(5)
users, items = torch.from_numpy(rng.integers(0, nu, 64)), torch.from_numpy(rng.integers(0, ni, 64))
out["users"], out["items"] = users.numpy(), items.numpy()
scores = m.predict(users, items, A)
out["scores"] = scores.detach().numpy()
loss = (scores ** 2).sum() + ue.abs().mean()
m.zero_grad()
loss.backward()
out["loss"] = np.float64(loss.item())
for k, p in m.named_parameters():
    out[f"grad/{k}"] = p.grad.numpy().copy() if p.grad is not None else np.zeros(0, np.float32)
out["n_params"] = np.int64(sum(p.numel() for p in m.parameters() if p.requires_grad))
np.savez_compressed(os.path.join(HERE, "kgtore.npz"), **out)
print("kgtore.npz written:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "kgtore.npz")), "bytes, loss", loss.item())
