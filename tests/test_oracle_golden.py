"""Pins the CPU oracle (oracle/pyoracle.py, oracle/oracle_ref.c) against fixtures produced by
the UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from gnn_recommendations_b200.synthetic import synth_interactions, temporal_split
from oracle import coracle
from oracle import pyoracle as po


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture(scope="module")
def tiny_adj(tiny):
    return po.build_norm_adj(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]))


# ------------------------------------------------------------------ generator + split
def test_generator_and_split_reproduce_fixture(tiny):
    u, i, ts = synth_interactions(int(tiny["n_users"]), int(tiny["n_items"]), len(tiny["user"]), 42)
    assert np.array_equal(u, tiny["user"]) and np.array_equal(i, tiny["item"]) and np.array_equal(ts, tiny["ts"])
    sp = temporal_split(u, i, ts)
    for k in ("train", "valid", "test"):
        assert np.array_equal(sp[k][0], tiny[f"{k}_u"]), k
        assert np.array_equal(sp[k][1], tiny[f"{k}_i"]), k


def test_split_edge_cases():
    # users with 1, 2 and 3 interactions (dataset.py:341-352)
    u = np.array([0, 1, 1, 2, 2, 2], dtype=np.int64)
    i = np.array([5, 1, 2, 7, 8, 9], dtype=np.int64)
    ts = np.array([0, 2, 1, 3, 5, 4], dtype=np.int64)
    sp = temporal_split(u, i, ts)
    assert sp["train"][0].tolist() == [0, 1, 2] and sp["train"][1].tolist() == [5, 2, 7]
    assert sp["valid"][0].tolist() == [2] and sp["valid"][1].tolist() == [9]
    assert sp["test"][0].tolist() == [1, 2] and sp["test"][1].tolist() == [1, 8]


# ------------------------------------------------------------------ graph build
def test_graph_build_bit_exact(tiny, tiny_adj):
    assert np.array_equal(tiny_adj["rows"], tiny["adj_row"].astype(np.int64))
    assert np.array_equal(tiny_adj["indices"], tiny["adj_col"])
    assert np.array_equal(tiny_adj["vals"].view(np.uint32), tiny["adj_val"].view(np.uint32))
    assert np.array_equal(tiny_adj["deg"], tiny["deg"])


def test_degree_lut_matches_array_power(tiny, tiny_adj):
    # the look-up-table form must equal numpy applied to the degree VECTOR (position independent)
    lut = po.dis_lut(int(tiny["deg"].max()))
    assert np.array_equal(lut.view(np.uint32), tiny["dis_lut"].view(np.uint32))
    assert np.array_equal(lut[tiny["deg"].astype(np.int64)].view(np.uint32), tiny_adj["dis"].view(np.uint32))


def test_row_normalisation(tiny):
    adj = po.build_norm_adj(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]), "row")
    assert np.array_equal(adj["vals"].view(np.uint32), tiny["adjrow_val"].view(np.uint32))


def test_graph_build_duplicates_and_isolated():
    # duplicates are summed by tocsr (graph_builder.py:107); node 3 (item 1) is isolated -> deg clamp
    import scipy.sparse as sp

    u = np.array([0, 0, 1, 0]); i = np.array([0, 0, 0, 0])
    adj = po.build_norm_adj(u, i, 2, 2)
    n = 4
    rows = np.concatenate([u, 2 + i]); cols = np.concatenate([2 + i, u])
    a = sp.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(n, n)).tocsr()
    deg = np.maximum(np.array(a.sum(axis=1)).ravel(), 1.0)
    d = sp.diags(np.power(deg, -0.5))
    ref = (d @ a @ d).tocoo()
    assert np.array_equal(adj["rows"], ref.row) and np.array_equal(adj["indices"], ref.col)
    assert np.array_equal(adj["vals"].view(np.uint32), ref.data.astype(np.float32).view(np.uint32))
    assert adj["indptr"].tolist() == [0, 1, 2, 4, 4]


# ------------------------------------------------------------------ propagation
def test_spmm_is_sequential_fmaf_chain(tiny, tiny_adj):
    x = np.concatenate([tiny["lightgcn/user_embedding.weight"], tiny["lightgcn/item_embedding.weight"]])
    y_torch = po.spmm(po.to_torch_coo(tiny_adj), t(x)).numpy()
    y_c = coracle.spmm_fmaf(tiny_adj["indptr"], tiny_adj["indices"], tiny_adj["vals"], x)
    assert np.array_equal(y_torch.view(np.uint32), y_c.view(np.uint32))


@pytest.mark.parametrize("tag,L", [("lightgcn", 3), ("lightgcn_d128_l4", 4)])
def test_lightgcn_forward_bit_exact(tiny, tiny_adj, tag, L):
    uw, iw = tiny[f"{tag}/user_embedding.weight"], tiny[f"{tag}/item_embedding.weight"]
    ue, ie = po.lightgcn_forward(po.to_torch_coo(tiny_adj), t(uw), t(iw), L)
    assert np.array_equal(ue.numpy().view(np.uint32), tiny[f"{tag}/out_user"].view(np.uint32))
    assert np.array_equal(ie.numpy().view(np.uint32), tiny[f"{tag}/out_item"].view(np.uint32))
    out_c = coracle.lightgcn_forward(tiny_adj["indptr"], tiny_adj["indices"], tiny_adj["vals"],
                                     np.concatenate([uw, iw]), L)
    assert np.array_equal(out_c[: len(uw)].view(np.uint32), tiny[f"{tag}/out_user"].view(np.uint32))
    assert np.array_equal(out_c[len(uw):].view(np.uint32), tiny[f"{tag}/out_item"].view(np.uint32))


def test_lightgcn_layers(tiny, tiny_adj):
    layers = po.lightgcn_layers(po.to_torch_coo(tiny_adj), t(tiny["lightgcn/user_embedding.weight"]),
                                t(tiny["lightgcn/item_embedding.weight"]), 3)
    assert np.array_equal(torch.stack(layers).numpy().view(np.uint32), tiny["lightgcn/layers"].view(np.uint32))


def test_ngcf_forward(tiny, tiny_adj):
    w1 = [t(tiny[f"ngcf/layers.{l}.W1.weight"]) for l in range(3)]
    b1 = [t(tiny[f"ngcf/layers.{l}.W1.bias"]) for l in range(3)]
    w2 = [t(tiny[f"ngcf/layers.{l}.W2.weight"]) for l in range(3)]
    b2 = [t(tiny[f"ngcf/layers.{l}.W2.bias"]) for l in range(3)]
    ue, ie = po.ngcf_forward(po.to_torch_coo(tiny_adj), t(tiny["ngcf/user_embedding.weight"]),
                             t(tiny["ngcf/item_embedding.weight"]), w1, b1, w2, b2)
    assert ue.shape[1] == 256
    assert np.array_equal(ue.numpy().view(np.uint32), tiny["ngcf/out_user"].view(np.uint32))
    assert np.array_equal(ie.numpy().view(np.uint32), tiny["ngcf/out_item"].view(np.uint32))


def gat_layers_from(tiny):
    layers = []
    for l in range(3):
        layers.append({
            "W": [t(tiny[f"gat/layers.{l}.W.{h}.weight"]) for h in range(4)],
            "a_self": [t(tiny[f"gat/layers.{l}.a_self.{h}"]) for h in range(4)],
            "a_neigh": [t(tiny[f"gat/layers.{l}.a_neigh.{h}"]) for h in range(4)],
            "concat": l < 2,
        })
    return layers


def test_gat_sparse_restatement_matches_dense_reference(tiny, tiny_adj):
    ue, ie = po.gat_forward_sparse(tiny_adj["indptr"], tiny_adj["indices"], t(tiny["gat/user_embedding.weight"]),
                                   t(tiny["gat/item_embedding.weight"]), gat_layers_from(tiny), 0.2)
    np.testing.assert_allclose(ue.numpy(), tiny["gat/out_user"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ie.numpy(), tiny["gat/out_item"], rtol=1e-5, atol=1e-7)


def test_group_shuffle_forward(tiny, tiny_adj):
    conn = [[t(tiny[f"gs/connection_layers.{l}.skew_params.{k}"]) for k in range(8)] for l in range(3)]
    cperm = [t(tiny[f"gs/connection_layers.{l}.shuffle_perm"]) for l in range(3)]
    loc = [[t(tiny[f"gs/local_transform_layers.{l}.skew_params.{k}"]) for k in range(8)] for l in range(3)]
    lperm = [t(tiny[f"gs/local_transform_layers.{l}.perm"]) for l in range(3)]
    ue, ie = po.gs_forward(po.to_torch_coo(tiny_adj), t(tiny["gs/user_embedding.weight"]),
                           t(tiny["gs/item_embedding.weight"]), conn, cperm, loc, lperm,
                           t(tiny["gs/layer_weights"]), 0.1)
    np.testing.assert_allclose(ue.numpy(), tiny["gs/out_user"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(ie.numpy(), tiny["gs/out_item"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("tag,transport", [("pt", True), ("nopt", False)])
def test_group_shuffle_edge_list_mode(tiny, tag, transport):
    """The oracle's edge-list restatement (parallel_transport.py:5-52, model.py:159-222) against the unmodified
    reference run on the same edge list (tests/golden/gs_edge.npz): embeddings and per-layer embeddings bit for bit
    (same torch CPU ops in the same order)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gs_edge.npz"))
    conn = [[t(tiny[f"gs/connection_layers.{l}.skew_params.{k}"]) for k in range(8)] for l in range(3)] if transport else None
    cperm = [t(tiny[f"gs/connection_layers.{l}.shuffle_perm"]) for l in range(3)] if transport else None
    loc = [[t(tiny[f"gs/local_transform_layers.{l}.skew_params.{k}"]) for k in range(8)] for l in range(3)]
    lperm = [t(tiny[f"gs/local_transform_layers.{l}.perm"]) for l in range(3)]
    args = (t(z["edge_index"]), t(tiny["gs/user_embedding.weight"]), t(tiny["gs/item_embedding.weight"]), conn, cperm,
            loc, lperm, t(tiny["gs/layer_weights"]), 0.1)
    ue, ie = po.gs_forward_edge_index(*args)
    np.testing.assert_allclose(ue.numpy(), z[f"{tag}/out_user"], rtol=1e-6, atol=1e-6 * np.abs(z[f"{tag}/out_user"]).max())
    np.testing.assert_allclose(ie.numpy(), z[f"{tag}/out_item"], rtol=1e-6, atol=1e-6 * np.abs(z[f"{tag}/out_item"]).max())
    layers = po.gs_forward_edge_index(*args, return_layers=True)
    for l, x in enumerate(layers):
        want = z[f"{tag}/layers"][l]
        np.testing.assert_allclose(x.numpy(), want, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(want).max()))


def test_parameter_count_kats(tiny):
    # /root/reference/problems.md:95-124: extra (non-embedding) parameters
    def extra(prefix):
        return sum(v.size for k, v in tiny.items()
                   if k.startswith(prefix + "/") and "embedding.weight" not in k and "out_" not in k
                   and "perm" not in k and k.split("/")[1] not in ("layers", "loss"))
    assert extra("ngcf") == 24960
    assert extra("gat") == 25344
    assert extra("gs") == 3076


# ------------------------------------------------------------------ sampler + BPR
def test_mt19937_matches_torch_randint():
    torch.manual_seed(42)
    ref = torch.randint(0, 988129, (1000,)).numpy()
    ref1 = [torch.randint(0, 3706, (1,)).item() for _ in range(50)]
    g = po.TorchCpuMt19937(42)
    assert np.array_equal(g.randint(988129, 1000), ref)
    assert [g.randint1(3706) for _ in range(50)] == ref1
    c = coracle.Mt19937(42)
    assert [c.next() % 988129 for _ in range(1000)] == ref.tolist()
    # ranges >= 2^28 consume two 32-bit outputs per element
    for rng_ in ((1 << 28) - 1, 1 << 28, 500_000_000, (1 << 40) + 7):
        torch.manual_seed(5)
        want = torch.randint(0, rng_, (9,)).numpy()
        assert np.array_equal(po.TorchCpuMt19937(5).randint(rng_, 9), want), rng_


def test_sampler_bit_exact(tiny):
    tu, ti = tiny["train_u"], tiny["train_i"]
    ps = po.positive_sets(tu, ti)
    g = po.TorchCpuMt19937(123)
    order = np.lexsort((ti, tu))
    pos_indptr = np.zeros(int(tiny["n_users"]) + 1, dtype=np.int64)
    np.cumsum(np.bincount(tu, minlength=int(tiny["n_users"])), out=pos_indptr[1:])
    pos_items = ti[order].astype(np.int32)
    cg = coracle.Mt19937(123)
    for b in range(3):
        us, pp, ng = po.sample_batch(g, tu, ti, int(tiny["n_items"]), 512, ps)
        assert np.array_equal(us, tiny[f"batch{b}/users"])
        assert np.array_equal(pp, tiny[f"batch{b}/pos"])
        assert np.array_equal(ng, tiny[f"batch{b}/neg"])
        cu, cp, cn, _ = coracle.sample_batch(cg, tu, ti, int(tiny["n_items"]), 512, pos_indptr, pos_items)
        assert np.array_equal(cu, us) and np.array_equal(cp, pp) and np.array_equal(cn, ng.reshape(-1))


def test_bpr_loss_is_b_by_b(tiny):
    ue, ie = t(tiny["lightgcn/out_user"]), t(tiny["lightgcn/out_item"])
    us, pp, ng = tiny["batch0/users"], tiny["batch0/pos"], tiny["batch0/neg"]
    loss = po.bpr_loss_reference(ue, ie, t(us), t(pp), t(ng))
    assert abs(float(loss) - float(tiny["step0/loss"])) <= 1e-7
    l64, gU, gI, _, _ = po.bpr_closed_form(ue.numpy(), ie.numpy(), us, pp, ng)
    assert abs(l64 - float(tiny["step0/loss"])) <= 1e-6 * abs(l64)
    np.testing.assert_allclose(gU, tiny["step0/gprop_user"], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(gI, tiny["step0/gprop_item"], rtol=2e-4, atol=1e-9)
    lc, _, _ = coracle.bpr_loss(ue.numpy(), ie.numpy(), us, pp, ng)
    assert abs(lc - l64) <= 1e-12


def test_lightgcn_step_gradients(tiny, tiny_adj):
    loss, gu, gi = po.lightgcn_loss_and_grads(po.to_torch_coo(tiny_adj), t(tiny["lightgcn/user_embedding.weight"]),
                                              t(tiny["lightgcn/item_embedding.weight"]), 3, tiny["batch0/users"],
                                              tiny["batch0/pos"], tiny["batch0/neg"])
    assert abs(float(loss) - float(tiny["step0/loss"])) <= 1e-7
    np.testing.assert_allclose(gu.numpy(), tiny["step0/grad_user"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(gi.numpy(), tiny["step0/grad_item"], rtol=1e-5, atol=1e-10)


# ------------------------------------------------------------------ top-K + metrics
def seen_sets(tiny, with_valid=True):
    d = {}
    us = [tiny["train_u"]] + ([tiny["valid_u"]] if with_valid else [])
    its = [tiny["train_i"]] + ([tiny["valid_i"]] if with_valid else [])
    for u, i in zip(np.concatenate(us).tolist(), np.concatenate(its).tolist()):
        d.setdefault(u, set()).add(i)
    return d


def seen_csr(seen, eval_users):
    indptr = [0]
    items = []
    for u in eval_users:
        s = sorted(seen.get(int(u), ()))
        items += s
        indptr.append(len(items))
    return np.asarray(indptr, dtype=np.int64), np.asarray(items, dtype=np.int32)


def test_topk_canonical(tiny):
    ue, ie = t(tiny["eval/user_emb"]), t(tiny["eval/item_emb"])
    eu = tiny["eval/users"].tolist()
    seen = seen_sets(tiny)
    tk = po.score_mask_topk(ue, ie, eu, seen, 20)
    assert np.array_equal(tk, tiny["eval/topk20_canonical"])
    # torch.topk agrees as a SET per user and as a list wherever scores are distinct
    assert all(set(a) == set(b) for a, b in zip(tk.tolist(), tiny["eval/topk20_torch"].tolist()))
    ip, it = seen_csr(seen, eu)
    tkc = coracle.score_topk(ue.numpy(), ie.numpy(), eu, ip, it, 20)
    assert np.array_equal(tkc, tiny["eval/topk20_canonical"])


def test_topk_ties_and_short_lists(tiny):
    ue = t(tiny["eval/user_emb"])
    eu = tiny["eval/users"].tolist()
    seen = seen_sets(tiny)
    tk = po.score_mask_topk(ue, t(tiny["eval/tie_item_emb"]), eu, seen, 20)
    assert np.array_equal(tk, tiny["eval/tie_topk20_canonical"])
    ip, it = seen_csr(seen, eu)
    assert np.array_equal(coracle.score_topk(ue.numpy(), tiny["eval/tie_item_emb"], eu, ip, it, 20),
                          tiny["eval/tie_topk20_canonical"])
    seen2 = dict(seen)
    seen2[eu[0]] = set(range(int(tiny["n_items"]))) - {3, 17, 42, 99, 150}
    tk2 = po.score_mask_topk(ue, t(tiny["eval/item_emb"]), eu[:4], seen2, 20)
    assert np.array_equal(tk2, tiny["eval/short_topk20_canonical"])
    ip2, it2 = seen_csr(seen2, eu[:4])
    assert np.array_equal(coracle.score_topk(ue.numpy(), tiny["eval/item_emb"], eu[:4], ip2, it2, 20),
                          tiny["eval/short_topk20_canonical"])


def test_metrics_from_topk(tiny):
    gt = {}
    for u, i in zip(tiny["test_u"].tolist(), tiny["test_i"].tolist()):
        gt.setdefault(u, []).append(i)
    m = po.metrics_from_topk(tiny["eval/topk20_canonical"], tiny["eval/users"].tolist(), gt,
                             int(tiny["n_items"]), [10, 20])
    for k in ("recall@10", "ndcg@10", "precision@10", "recall@20", "ndcg@20", "precision@20",
              "coverage@20", "gini@20"):
        assert abs(m[k] - float(tiny[f"evaluate/{k}"])) <= 1e-12, k


# ------------------------------------------------------------------ C1 shape (hash-pinned)
def test_c1_graph_and_forward_hashes(c1gold, c1split):
    u, i, ts = c1split["all"]
    assert sha(u) == str(c1gold["sha_user"]) and sha(i) == str(c1gold["sha_item"])
    tu, ti = c1split["train"]
    assert len(tu) == int(c1gold["n_train"]) and sha(tu) == str(c1gold["sha_train_u"])
    adj = po.build_norm_adj(tu, ti, c1split["n_users"], c1split["n_items"])
    assert len(adj["vals"]) == int(c1gold["nnz"])
    assert sha(adj["rows"].astype(np.int32)) == str(c1gold["sha_adj_row"])
    assert sha(adj["indices"]) == str(c1gold["sha_adj_col"])
    lut = c1gold["dis_lut"]
    vals = (lut[adj["deg"].astype(np.int64)][adj["rows"]] * np.float32(1.0)) * lut[adj["deg"].astype(np.int64)][adj["indices"]]
    assert sha(vals.astype(np.float32)) == str(c1gold["sha_adj_val"])
