"""Generates tests/golden/gs_edge.npz from the UNMODIFIED reference OrthogonalBundleGNN in its EDGE-LIST mode
(use_edge_index=True: src/models/orthogonal_bundle/model.py:159-222 + parallel_transport.py:5-52) on the tiny
synthetic graph, with the parameters of the tiny fixture ("gs/*" of tiny.npz): forward embeddings, per-layer
embeddings and the gradients of a linear probe loss, with and without parallel transport.  Build container only.

    python tests/golden/make_golden_gs_edge.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/gnn-recommendations")
from src.models import OrthogonalBundleGNN  # noqa: E402

tiny = np.load(os.path.join(HERE, "tiny.npz"))
nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
u, i = tiny["train_u"].astype(np.int64), tiny["train_i"].astype(np.int64)
# the bipartite interaction edges in both directions, in interaction order (duplicates kept: an edge list may repeat)
edge_index = torch.from_numpy(np.stack([np.concatenate([u, nu + i]), np.concatenate([nu + i, u])]))
rng = np.random.default_rng(17)
probe_u = torch.from_numpy(rng.standard_normal((nu, 64)).astype(np.float32))
probe_i = torch.from_numpy(rng.standard_normal((ni, 64)).astype(np.float32))
out = {"edge_index": edge_index.numpy(), "probe_u": probe_u.numpy(), "probe_i": probe_i.numpy()}
for tag, transport in (("pt", True), ("nopt", False)):
    torch.manual_seed(42)
    m = OrthogonalBundleGNN(nu, ni, embedding_dim=64, n_layers=3, block_size=8, residual_alpha=0.1, dropout=0.0,
                            init_scale=0.01, use_parallel_transport=transport, use_edge_index=True)
    sd = m.state_dict()
    for k in sd:
        sd[k] = torch.from_numpy(tiny[f"gs/{k}"].copy())
    m.load_state_dict(sd)
    m.eval()
    ue, ie = m(edge_index=edge_index)
    out[f"{tag}/out_user"], out[f"{tag}/out_item"] = ue.detach().numpy(), ie.detach().numpy()
    out[f"{tag}/layers"] = torch.stack(m.get_layer_embeddings(edge_index=edge_index)).detach().numpy()
    loss = (ue * probe_u).sum() + (ie * probe_i).sum()
    m.zero_grad()
    loss.backward()
    out[f"{tag}/loss"] = np.float64(loss.item())
    for k, p in m.named_parameters():
        out[f"{tag}/grad/{k}"] = p.grad.numpy().copy()
    # the same gradients from the reference in float64: the error band of its own fp32 run (the edge sums are
    # unnormalised — values reach 1e3 — and the skew-parameter gradients are differences of such sums)
    m64 = m.double()
    ue, ie = m64(edge_index=edge_index)
    loss = (ue * probe_u.double()).sum() + (ie * probe_i.double()).sum()
    m64.zero_grad()
    loss.backward()
    for k, p in m64.named_parameters():
        out[f"{tag}/grad64/{k}"] = p.grad.numpy().copy()
np.savez_compressed(os.path.join(HERE, "gs_edge.npz"), **out)
print("gs_edge.npz written:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "gs_edge.npz")), "bytes")
