"""GPU parity of the NGCF / GAT / Group-and-Shuffle drop-ins against the reference's golden
forward outputs and autograd gradients (eval() mode: dropout is the identity), 1e-5 relative."""
import numpy as np
import pytest
import torch

import gnn_recommendations_b200 as g
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def load(model, tiny, prefix):
    sd = model.state_dict()
    new = {}
    for k in sd:
        v = torch.from_numpy(tiny[f"{prefix}/{k}"])
        assert v.shape == sd[k].shape, k
        new[k] = v
    model.load_state_dict(new)
    return model.to(DEV).eval()


@pytest.fixture(scope="module")
def csr(tiny):
    return g.NormAdjCSR.from_pairs(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]),
                                   device=DEV, dis_lut=tiny["dis_lut"])


def make(name, tiny):
    nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
    if name == "ngcf":
        return load(g.NGCF(nu, ni, 64, [64, 64, 64], 0.1, 0.1), tiny, "ngcf")
    if name == "gat":
        return load(g.GAT(nu, ni, 64, 3, 4, 0.1, 0.2, 0.1), tiny, "gat")
    return load(g.OrthogonalBundleGNN(nu, ni, 64, 3, 8, 0.1, 0.0, 0.01), tiny, "gs")


def close(a, b, rtol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("name", ["ngcf", "gat", "gs"])
def test_forward_vs_reference(tiny, csr, name):
    m = make(name, tiny)
    with torch.no_grad():
        ue, ie = m.get_all_embeddings(csr)
    assert ue.shape == tuple(tiny[f"{name}/out_user"].shape)
    close(ue.cpu().numpy(), tiny[f"{name}/out_user"])
    close(ie.cpu().numpy(), tiny[f"{name}/out_item"])
    # the torch COO entry the reference Trainer passes gives the same result
    with torch.no_grad():
        ue2, _ = m(csr.to_torch_coo())
    assert torch.equal(ue, ue2)


@pytest.mark.parametrize("name", ["ngcf", "gat", "gs"])
def test_gradients_vs_reference_autograd(tiny, csr, name):
    m = make(name, tiny)
    if name == "gs":
        m.train()                                        # golden gradients were taken in train() (dropout 0)
    users, pos, neg = (torch.from_numpy(tiny[f"grad_{k}"]).to(DEV) for k in ("users", "pos", "neg"))
    x = m.propagate(csr)
    loss = g.bpr_fused(x, m.n_users, users, pos, neg)
    m.zero_grad()
    loss.backward()
    ref = float(tiny[f"{name}/loss"])
    assert abs(float(loss.detach()) - ref) <= 1e-5 * abs(ref)
    for k, p in m.named_parameters():
        want = tiny[f"{name}_grad/{k}"]
        assert p.grad is not None, k
        close(p.grad.cpu().numpy(), want, rtol=2e-4)


def test_gat_isolated_node_is_nan_like_reference():
    # gat.py: softmax over an all -inf row -> NaN; user 1 has no interactions
    u = np.array([0, 0, 2]); i = np.array([0, 1, 1])
    csr = g.NormAdjCSR.from_pairs(u, i, 3, 2, device=DEV)
    torch.manual_seed(0)
    m = g.GAT(3, 2, 64, 1, 4, 0.0, 0.2, 0.1).to(DEV).eval()
    with torch.no_grad():
        ue, ie = m(csr)
    assert torch.isnan(ue[1]).all() and torch.isfinite(ue[0]).all() and torch.isfinite(ie).all()


def test_gs_layer_embeddings_and_metrics(tiny, csr):
    m = make("gs", tiny)
    layers = m.get_layer_embeddings(csr)
    assert len(layers) == 4 and layers[1].shape == (500, 64)
    met = m.get_orthogonality_metrics()
    assert float(met["local_fro_max"]) < 1e-4 and float(met["conn_fro_max"]) < 1e-4
    with pytest.raises(NotImplementedError):
        g.OrthogonalBundleGNN(5, 5, use_edge_index=True).to(DEV)(None, torch.zeros(2, 3, dtype=torch.long))


def test_rowmap_kernel_vs_torch():
    from gnn_recommendations_b200.layer_ops import ACT_ELU, ACT_LEAKY, _rowmap_torch, rowmap
    gen = torch.Generator().manual_seed(0)
    for n, d_in, d_out, act in ((1, 64, 64, ACT_LEAKY), (130, 32, 16, ACT_ELU), (1000, 128, 128, 0), (77, 64, 8, 0)):
        x1, x2, x3 = (torch.randn(n, d_in, generator=gen).to(DEV) for _ in range(3))
        wa, wb = (torch.randn(d_in, d_out, generator=gen).to(DEV) * 0.2 for _ in range(2))
        ba, bb = (torch.randn(d_out, generator=gen).to(DEV) for _ in range(2))
        r = torch.randn(n, d_out, generator=gen).to(DEV)
        got = rowmap(x1, wa, ba, x2, x3, wb, bb, r, alpha=0.9, beta=0.1, act=act, slope=0.2)
        want = _rowmap_torch(x1.double(), wa.double(), ba.double(), x2.double(), x3.double(), wb.double(), bb.double(),
                             r.double(), 0.9, 0.1, act, 0.2)
        close(got.cpu().numpy(), want.cpu().numpy())
        got1 = rowmap(x1, wa)
        close(got1.cpu().numpy(), (x1.double() @ wa.double()).cpu().numpy())
    x, w = torch.randn(300, 128, generator=gen).to(DEV), torch.randn(128, 256, generator=gen).to(DEV)
    close(rowmap(x, w).cpu().numpy(), (x.double() @ w.double()).cpu().numpy())


# ----------------------------------------------------------------------------- KGTORe
def test_kgtore_matches_reference_golden(tiny):
    """Same-seed parameters bit-equal; eval-mode forward, predict and every parameter gradient vs the
    unmodified reference (tests/golden/make_golden_kgtore.py; kgtore.py:170-394)."""
    import os
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "kgtore.npz")))
    nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
    torch.manual_seed(42)
    m = g.KGTORe(nu, ni, embedding_dim=64, kg_embedding_dim=32, tree_depth=3, n_layers=2, dropout=0.1, init_scale=0.1)
    sd = m.state_dict()
    assert set(sd) == {k[5:] for k in z if k.startswith("init/")}
    for k, v in sd.items():
        assert np.array_equal(v.numpy(), z[f"init/{k}"]), k           # constructor RNG order == reference's
    assert m.get_parameters_count() == int(z["n_params"])
    m = m.to(DEV).eval()
    csr = g.NormAdjCSR.from_pairs(tiny["train_u"], tiny["train_i"], nu, ni, device=DEV)
    ue, ie = m(csr)
    np.testing.assert_allclose(ue.detach().cpu().numpy(), z["user_emb"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ie.detach().cpu().numpy(), z["item_emb"], rtol=1e-5, atol=1e-7)
    users, items = torch.from_numpy(z["users"]).to(DEV), torch.from_numpy(z["items"]).to(DEV)
    scores = m.predict(users, items, csr)
    np.testing.assert_allclose(scores.detach().cpu().numpy(), z["scores"], rtol=1e-5, atol=1e-7)
    loss = (scores ** 2).sum() + ue.abs().mean()
    m.zero_grad()
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    for k, p in m.named_parameters():
        want = z[f"grad/{k}"]
        if want.size == 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=2e-4, atol=2e-4 * float(np.abs(want).max()) + 1e-9, err_msg=k)
    with pytest.raises(ValueError):
        m.get_all_embeddings(None)
    assert g.MODEL_REGISTRY["kgtore"] is g.KGTORe
