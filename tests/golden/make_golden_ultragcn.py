"""Generates tests/golden/ultragcn.npz from the UNMODIFIED reference UltraGCN
(/root/reference/gnn-recommendations/src/models/baselines/ultragcn.py).  Build container only.

    python tests/golden/make_golden_ultragcn.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/gnn-recommendations")
from src.models import UltraGCN  # noqa: E402

nu, ni, d, B = 60, 90, 32, 16
torch.manual_seed(42)
m = UltraGCN(nu, ni, embedding_dim=d, lambda_1=0.7, lambda_2=1.3, gamma=1e-3, init_scale=0.1)
out = {"n_users": nu, "n_items": ni, "d": d,
       "user_w": m.user_embedding.weight.detach().numpy().copy(), "item_w": m.item_embedding.weight.detach().numpy().copy()}
rng = np.random.default_rng(3)
users = torch.from_numpy(rng.integers(0, nu, B)); pos = torch.from_numpy(rng.integers(0, ni, B)); neg = torch.from_numpy(rng.integers(0, ni, B))
adj = (rng.random((nu, ni)) < 0.08).astype(np.float32) * rng.random((nu, ni)).astype(np.float32)
adj[int(users[0])] = 0.0                                     # a user without neighbours
out.update(users=users.numpy(), pos=pos.numpy(), neg=neg.numpy(), adj=adj)
out["predict"] = m.predict(users, pos).detach().numpy()
ue, ie = m(None)
assert ue is m.user_embedding.weight and ie is m.item_embedding.weight
for tag, a in (("noadj", None), ("dense", torch.from_numpy(adj)), ("sparse", torch.from_numpy(adj).to_sparse())):
    m.zero_grad()
    total, parts = m.compute_loss(users, pos, neg, a)
    total.backward()
    out[f"{tag}/total"] = np.float64(total.item())
    out[f"{tag}/constraint"] = np.float64(parts["constraint_loss"])
    out[f"{tag}/grad_user"] = m.user_embedding.weight.grad.numpy().copy()
    out[f"{tag}/grad_item"] = m.item_embedding.weight.grad.numpy().copy()
np.savez_compressed(os.path.join(HERE, "ultragcn.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
