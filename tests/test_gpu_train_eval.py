"""GPU parity of the BPR step, the full-ranking top-K and the Trainer / Evaluator drop-ins against
the reference's golden outputs and the CPU oracle."""
import numpy as np
import pytest
import torch

import gnn_recommendations_b200 as g
from gnn_recommendations_b200.evaluator import ground_truth_dict, seen_csr
from oracle import coracle
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dataset_from(tiny):
    return g.InteractionDataset((tiny["train_u"], tiny["train_i"]), (tiny["valid_u"], tiny["valid_i"]),
                                (tiny["test_u"], tiny["test_i"]), int(tiny["n_users"]), int(tiny["n_items"]),
                                device=DEV)


def lightgcn_from(tiny, tag="lightgcn", d=64, L=3):
    m = g.LightGCN(int(tiny["n_users"]), int(tiny["n_items"]), d, L, 0.1)
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(tiny[f"{tag}/user_embedding.weight"]),
                       "item_embedding.weight": torch.from_numpy(tiny[f"{tag}/item_embedding.weight"])})
    return m.to(DEV)


# ----------------------------------------------------------------------------- fused BPR kernel
def test_bpr_fused_loss_and_gradient_vs_reference(tiny):
    emb = torch.from_numpy(np.concatenate([tiny["lightgcn/out_user"], tiny["lightgcn/out_item"]])).to(DEV)
    emb.requires_grad_(True)
    us, ps, ns = (torch.from_numpy(tiny[f"batch0/{k}"]).to(DEV) for k in ("users", "pos", "neg"))
    loss = g.bpr_fused(emb, int(tiny["n_users"]), us, ps, ns)
    loss.backward()
    ref = float(tiny["step0/loss"])
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref)
    nu = int(tiny["n_users"])
    np.testing.assert_allclose(emb.grad[:nu].cpu().numpy(), tiny["step0/gprop_user"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(emb.grad[nu:].cpu().numpy(), tiny["step0/gprop_item"], rtol=1e-4, atol=1e-9)
    # float64 closed form (tighter: independent of the reference's fp32 summation order)
    l64, gU, gI, _, _ = po.bpr_closed_form(tiny["lightgcn/out_user"], tiny["lightgcn/out_item"], tiny["batch0/users"],
                                           tiny["batch0/pos"], tiny["batch0/neg"])
    assert abs(float(loss) - l64) <= 2e-6 * abs(l64)
    np.testing.assert_allclose(emb.grad[:nu].cpu().numpy(), gU, rtol=2e-5, atol=1e-10)
    np.testing.assert_allclose(emb.grad[nu:].cpu().numpy(), gI, rtol=2e-5, atol=1e-10)


@pytest.mark.parametrize("b,d", [(1, 64), (7, 32), (512, 128), (2048, 256)])
def test_bpr_fused_shapes_and_duplicates(b, d):
    rng = np.random.default_rng(b)
    nu, ni = 50, 40
    emb = (rng.standard_normal((nu + ni, d)) * 0.3).astype(np.float32)
    us, ps, ns = rng.integers(0, nu, b), rng.integers(0, ni, b), rng.integers(0, ni, b)   # heavy duplication
    e = torch.from_numpy(emb).to(DEV).requires_grad_(True)
    loss = g.bpr_fused(e, nu, torch.from_numpy(us), torch.from_numpy(ps), torch.from_numpy(ns).view(-1, 1))
    (loss * 2.0).backward()                                   # upstream gradient is honoured
    l64, gU, gI, _, _ = po.bpr_closed_form(emb[:nu], emb[nu:], us, ps, ns)
    assert abs(float(loss) - l64) <= 1e-5 * abs(l64)
    got = e.grad.cpu().numpy()
    scale = max(np.abs(gU).max(), np.abs(gI).max())
    np.testing.assert_allclose(got[:nu], 2 * gU, rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_allclose(got[nu:], 2 * gI, rtol=1e-4, atol=1e-5 * scale)


def test_bpr_fused_rejects_multiple_negatives():
    e = torch.zeros(10, 64, device=DEV)
    with pytest.raises(ValueError):
        g.bpr_fused(e, 5, torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64),
                    torch.zeros((4, 2), dtype=torch.int64))


# ----------------------------------------------------------------------------- top-K
def seen_from(tiny, with_valid=True):
    sets = [(tiny["train_u"], tiny["train_i"])] + ([(tiny["valid_u"], tiny["valid_i"])] if with_valid else [])
    return seen_csr(tiny["eval/users"], int(tiny["n_users"]), *sets)


@pytest.mark.parametrize("n_splits", [1, 2, 3])
def test_topk_bit_exact_vs_reference(tiny, n_splits):
    ue, ie = torch.from_numpy(tiny["eval/user_emb"]).to(DEV), torch.from_numpy(tiny["eval/item_emb"]).to(DEV)
    ip, it = seen_from(tiny)
    ids, sc = g.full_rank_topk(ue, ie, tiny["eval/users"], ip, it, 20, n_splits=n_splits, return_scores=True)
    assert np.array_equal(ids.cpu().numpy(), tiny["eval/topk20_canonical"])
    # scores are the reference's scores bit-for-bit (k-sequential fmaf chain == CPU sgemm)
    ref_scores = np.take_along_axis(tiny["eval/scores"], tiny["eval/topk20_canonical"], axis=1)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), ref_scores.view(np.uint32))


def test_topk_exact_ties_break_by_item_id(tiny):
    ue = torch.from_numpy(tiny["eval/user_emb"]).to(DEV)
    ie = torch.from_numpy(tiny["eval/tie_item_emb"]).to(DEV)
    ip, it = seen_from(tiny)
    for n_splits in (1, 2):
        ids = g.full_rank_topk(ue, ie, tiny["eval/users"], ip, it, 20, n_splits=n_splits)
        assert np.array_equal(ids.cpu().numpy(), tiny["eval/tie_topk20_canonical"])


def test_topk_fewer_than_k_unmasked_items(tiny):
    # a user who has seen all but 5 items: the list is completed with -inf entries in id order
    ue, ie = torch.from_numpy(tiny["eval/user_emb"]).to(DEV), torch.from_numpy(tiny["eval/item_emb"]).to(DEV)
    eu = tiny["eval/users"][:4]
    ni = int(tiny["n_items"])
    seen = {int(u): set() for u in eu}
    for u, i in zip(np.concatenate([tiny["train_u"], tiny["valid_u"]]).tolist(),
                    np.concatenate([tiny["train_i"], tiny["valid_i"]]).tolist()):
        if u in seen:
            seen[u].add(i)
    seen[int(eu[0])] = set(range(ni)) - {3, 17, 42, 99, 150}
    indptr, items = [0], []
    for u in eu:
        items += sorted(seen[int(u)])
        indptr.append(len(items))
    for n_splits in (1, 2):
        ids = g.full_rank_topk(ue, ie, eu, np.asarray(indptr), np.asarray(items, dtype=np.int32), 20,
                               n_splits=n_splits)
        assert np.array_equal(ids.cpu().numpy(), tiny["eval/short_topk20_canonical"])


@pytest.mark.parametrize("d,k,n_items", [(32, 1, 130), (64, 50, 1000), (128, 64, 257), (256, 10, 5000)])
def test_topk_random_vs_c_oracle(d, k, n_items):
    rng = np.random.default_rng(d + k)
    nu = 150
    ue = rng.standard_normal((nu, d)).astype(np.float32)
    ie = rng.standard_normal((n_items, d)).astype(np.float32)
    ie[rng.integers(0, n_items, 20)] = ie[0]                       # exact duplicates -> exact ties
    eu = np.sort(rng.choice(nu, 100, replace=False))
    indptr, items = [0], []
    for _ in eu:
        items += sorted(rng.choice(n_items, rng.integers(0, min(60, n_items - k)), replace=False).tolist())
        indptr.append(len(items))
    indptr, items = np.asarray(indptr), np.asarray(items, dtype=np.int32)
    want = coracle.score_topk(ue, ie, eu, indptr, items, k)
    ids = g.full_rank_topk(torch.from_numpy(ue).to(DEV), torch.from_numpy(ie).to(DEV), eu, indptr, items, k)
    assert np.array_equal(ids.cpu().numpy(), want)
    ids2 = g.full_rank_topk(torch.from_numpy(ue).to(DEV), torch.from_numpy(ie).to(DEV), eu, None, None, k)
    assert np.array_equal(ids2.cpu().numpy(), coracle.score_topk(ue, ie, eu, None, None, k))


def test_topk_c1_shape_vs_reference(c1gold, c1split):
    """Full C1 size: graph build -> propagation -> top-20, end to end on the GPU, must reproduce the
    reference's canonical lists for all 6040 users and its Recall/NDCG."""
    nu, ni = c1split["n_users"], c1split["n_items"]
    tu, ti = c1split["train"]
    csr = g.NormAdjCSR.from_pairs(tu, ti, nu, ni, device=DEV, dis_lut=c1gold["dis_lut"])
    rng = np.random.default_rng(42)
    uw = (rng.standard_normal((nu, 64)) * 0.1).astype(np.float32)
    iw = (rng.standard_normal((ni, 64)) * 0.1).astype(np.float32)
    m = g.LightGCN(nu, ni, 64, 3, 0.1)
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(uw), "item_embedding.weight": torch.from_numpy(iw)})
    m.to(DEV)
    with torch.no_grad():
        ue, ie = m(csr)
    gt = ground_truth_dict(c1split["test"])
    eu = sorted(gt)
    ip, it = seen_csr(eu, nu, c1split["train"], c1split["valid"])
    ids = g.full_rank_topk(ue, ie, eu, ip, it, 20).cpu()
    assert np.array_equal(ids.numpy(), c1gold["topk20_canonical"].astype(np.int64))
    met = g.compute_metrics_from_topk(ids, eu, gt, ni, [10, 20])
    for k in ("recall@20", "ndcg@20", "recall@10", "ndcg@10"):
        assert abs(met[k] - float(c1gold[f"metrics/{k}"])) <= 1e-12, k
    gp, gi = seen_csr(eu, nu, c1split["test"])
    dmet = g.topk_metrics_device(ids.to(DEV), gp, gi, ni, [10, 20])
    assert set(dmet) == set(met)
    for k in met:
        assert abs(dmet[k] - met[k]) <= 1e-12, k
    for k in ("recall@20", "ndcg@20", "recall@10", "ndcg@10"):
        assert abs(dmet[k] - float(c1gold[f"metrics/{k}"])) <= 1e-12, k


# ----------------------------------------------------------------------------- Trainer / Evaluator
CFG = {"learning_rate": 1e-3, "weight_decay": 1e-4, "batch_size": 512, "epochs": 1, "eval_every": 1,
       "use_scheduler": False, "warmup_epochs": 0, "max_grad_norm": 1.0, "negative_samples": 1,
       "validation_metrics": ["recall@10", "recall@20", "recall@50", "ndcg@10", "ndcg@20", "ndcg@50"]}


def test_trainer_epoch_validate_evaluate_vs_reference(tiny, tmp_path):
    ds = dataset_from(tiny)
    torch.manual_seed(42)
    m = g.LightGCN(int(tiny["n_users"]), int(tiny["n_items"]), embedding_dim=64, n_layers=3, init_scale=0.1)
    # same-seed construction reproduces the reference's initial parameters
    assert np.array_equal(m.user_embedding.weight.detach().numpy(), tiny["lightgcn/user_embedding.weight"])
    tr = g.Trainer(m, ds, dict(CFG, checkpoint_dir=str(tmp_path / "ckpt")), device=torch.device(DEV))
    torch.manual_seed(123)
    us, ps, ns = tr._sample_batch()
    assert np.array_equal(us.cpu().numpy(), tiny["batch0/users"]) and np.array_equal(ns.cpu().numpy(), tiny["batch0/neg"])
    torch.manual_seed(123)
    loss = tr.train_epoch()                                     # 11 steps of sample/fwd/BPR/bwd/clip/Adam
    assert abs(loss - float(tiny["epoch/loss"])) <= 1e-5 * abs(float(tiny["epoch/loss"]))
    np.testing.assert_allclose(m.user_embedding.weight.detach().cpu().numpy(), tiny["epoch/user_w"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(m.item_embedding.weight.detach().cpu().numpy(), tiny["epoch/item_w"], rtol=1e-4, atol=2e-6)
    # validate / evaluate on the REFERENCE's post-epoch weights so the metric comparison is exact
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(tiny["epoch/user_w"]),
                       "item_embedding.weight": torch.from_numpy(tiny["epoch/item_w"])})
    vm = tr.validate()
    for k, v in vm.items():
        assert abs(v - float(tiny[f"validate/{k}"])) <= 1e-12, k
    em = g.Evaluator(k_values=[10, 20], device=torch.device(DEV)).evaluate(m, ds, ds.test_data)
    for k, v in em.items():
        assert abs(v - float(tiny[f"evaluate/{k}"])) <= 1e-12, k
    tr.save_checkpoint(1, vm)
    tr.load_checkpoint(tmp_path / "ckpt" / "checkpoint_epoch_1.pt")
    assert tr.current_epoch == 1


# ----------------------------------------------------------------------------- tensor-core nomination
@pytest.mark.parametrize("d,k,n_items,nu", [(64, 20, 3000, 300), (64, 10, 130, 70), (32, 20, 2500, 129), (64, 32, 4097, 257), (64, 50, 3000, 200),
                                             (128, 20, 2100, 140), (128, 40, 1500, 260)])
def test_topk_tensor_core_path_is_bit_identical(d, k, n_items, nu):
    """tcgen05 TF32 nomination + exact re-scoring (+ exact fallback for unproven rows) must reproduce
    the exact kernel / C oracle bit for bit, including exact ties and seen-item masks."""
    rng = np.random.default_rng(d * 1000 + k)
    ue = (rng.standard_normal((nu, d)) * rng.uniform(0.05, 2.0, (nu, 1))).astype(np.float32)
    ie = (rng.standard_normal((n_items, d)) * rng.uniform(0.1, 1.5, (n_items, 1))).astype(np.float32)
    ie[rng.integers(0, n_items, 12)] = ie[1]                       # exact ties
    eu = np.sort(rng.choice(nu, nu - 7, replace=False))
    indptr, items = [0], []
    for j, _ in enumerate(eu):
        m = n_items - k + 3 if j == 5 else int(rng.integers(0, min(80, n_items - k)))   # row 5: fewer than k left
        items += sorted(rng.choice(n_items, m, replace=False).tolist())
        indptr.append(len(items))
    indptr, items = np.asarray(indptr), np.asarray(items, dtype=np.int32)
    want = coracle.score_topk(ue, ie, eu, indptr, items, k)
    stats = {}
    ids, sc = g.full_rank_topk(torch.from_numpy(ue).to(DEV), torch.from_numpy(ie).to(DEV), eu, indptr, items, k,
                               return_scores=True, tensor_cores=True, stats=stats)
    assert stats["tensor_cores"] and stats["rows_reranked_exactly"] < len(eu)
    assert np.array_equal(ids.cpu().numpy(), want)
    ids2, sc2 = g.full_rank_topk(torch.from_numpy(ue).to(DEV), torch.from_numpy(ie).to(DEV), eu, indptr, items, k,
                                 return_scores=True, tensor_cores=False)
    assert torch.equal(ids, ids2) and torch.equal(sc, sc2)


def test_topk_tensor_core_c1_shape(c1gold, c1split):
    nu, ni = c1split["n_users"], c1split["n_items"]
    csr = g.NormAdjCSR.from_pairs(*c1split["train"], nu, ni, device=DEV, dis_lut=c1gold["dis_lut"])
    rng = np.random.default_rng(42)
    uw = (rng.standard_normal((nu, 64)) * 0.1).astype(np.float32)
    iw = (rng.standard_normal((ni, 64)) * 0.1).astype(np.float32)
    m = g.LightGCN(nu, ni, 64, 3, 0.1)
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(uw), "item_embedding.weight": torch.from_numpy(iw)})
    m.to(DEV)
    with torch.no_grad():
        ue, ie = m(csr)
    gt = ground_truth_dict(c1split["test"])
    eu = sorted(gt)
    ip, it = seen_csr(eu, nu, c1split["train"], c1split["valid"])
    stats = {}
    ids = g.full_rank_topk(ue, ie, eu, ip, it, 20, tensor_cores=True, stats=stats).cpu()
    assert np.array_equal(ids.numpy(), c1gold["topk20_canonical"].astype(np.int64))
    print("C1 tensor-core path: rows re-ranked exactly:", stats["rows_reranked_exactly"], "of", stats["rows"])


# ----------------------------------------------------------------------------- on-device metrics / fused optimizer
@pytest.mark.parametrize("nu,ni,kmax,ks", [(500, 300, 20, [5, 10, 20]), (257, 1000, 50, [10, 20, 50]), (64, 40, 64, [1, 64, 100])])
def test_topk_metrics_device_matches_reference_loops(nu, ni, kmax, ks):
    """gr_topk_metrics vs the reference's python loops (oracle restatement of metrics.py:355-432):
    users without ground truth, duplicate ground-truth entries, k > list length."""
    rng = np.random.default_rng(nu + kmax)
    topk = np.stack([rng.permutation(ni)[:kmax] for _ in range(nu)]).astype(np.int64)
    users = np.sort(rng.choice(10 * nu, nu, replace=False))
    gt = {}
    for r, u in enumerate(users.tolist()):
        if r % 7 == 3:
            continue                                            # no ground truth: skipped by the reference
        n_rel = int(rng.integers(1, 2 * kmax))
        rel = rng.choice(ni, min(n_rel, ni), replace=False).tolist()
        if r % 5 == 0:
            rel += topk[r, :3].tolist() + rel[:2]               # guaranteed hits + duplicates
        gt[u] = rel
    want = po.metrics_from_topk(topk, users.tolist(), gt, ni, ks)
    indptr, items = [0], []
    for u in users.tolist():
        items += sorted(set(gt.get(u, ())))
        indptr.append(len(items))
    got = g.topk_metrics_device(torch.from_numpy(topk).to(DEV), np.asarray(indptr), np.asarray(items, dtype=np.int32), ni, ks)
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-13, (k, got[k], v)


@pytest.mark.parametrize("max_norm,gscale", [(1.0, 10.0), (1.0, 1e-3), (0.0, 1.0)])
def test_fused_clip_adam_matches_torch(max_norm, gscale):
    """gr_clip_adam_fused vs clip_grad_norm_ + optim.Adam on CPU (what trainer.py:273-276 runs)."""
    from gnn_recommendations_b200.optim import fused_clip_adam_step, fused_clip_adam_supported
    shapes = [(1000, 64), (37, 5), (3,), (4096,), (129, 33)]
    gen = torch.Generator().manual_seed(7)
    ref = [torch.nn.Parameter(torch.randn(*s, generator=gen) * 0.1) for s in shapes]
    dev = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-4)
    o_dev = torch.optim.Adam(dev, lr=1e-3, weight_decay=1e-4)
    for step in range(4):
        for p, q in zip(ref, dev):
            gr = torch.randn(p.shape, generator=gen) * gscale
            p.grad = gr.clone()
            q.grad = gr.clone().to(DEV)
        want_norm = torch.nn.utils.clip_grad_norm_(ref, max_norm) if max_norm > 0 else None
        o_ref.step()
        assert fused_clip_adam_supported(o_dev)
        norm = fused_clip_adam_step(o_dev, max_norm)
        if want_norm is not None:
            assert abs(float(norm) - float(want_norm)) <= 1e-6 * float(want_norm)
        for p, q in zip(ref, dev):
            np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().numpy(), rtol=2e-6, atol=1e-8)
            # a 1-ulp difference in the clip coefficient moves every g by 1 ulp; m cancels, so bound absolutely
            for key in ("exp_avg", "exp_avg_sq"):
                want = o_ref.state[p][key].numpy()
                np.testing.assert_allclose(o_dev.state[q][key].cpu().numpy(), want, rtol=2e-6, atol=1e-6 * float(np.abs(want).max()))
        assert float(o_dev.state[dev[0]]["step"]) == step + 1
    # the state layout is torch's own: a stock step continues from it
    sd = o_dev.state_dict()
    o2 = torch.optim.Adam(dev, lr=1e-3, weight_decay=1e-4)
    o2.load_state_dict(sd)
    for q in dev:
        q.grad = torch.zeros_like(q)
    o2.step()


# ----------------------------------------------------------------------------- device temporal split / interaction store
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_temporal_split_device_c1_matches_reference_hashes(c1gold, c1split):
    """gr_temporal_split at the ML-1M shape: train pairs hash-equal to the reference's split()
    (dataset.py:327-357; hashes generated from the unmodified reference, tests/golden/make_golden.py)."""
    from gnn_recommendations_b200.dataset import temporal_split_device
    u, i, ts = c1split["all"]
    sp = temporal_split_device(u, i, ts, c1split["n_users"], device=DEV)
    tu, ti = (x.cpu().numpy() for x in sp["train"])
    assert len(tu) == int(c1gold["n_train"])
    assert _sha(tu) == str(c1gold["sha_train_u"]) and _sha(ti) == str(c1gold["sha_train_i"])
    for part in ("valid", "test"):
        for a, b in zip(sp[part], c1split[part]):
            assert np.array_equal(a.cpu().numpy(), b)


def test_temporal_split_device_edge_cases():
    """users with 1, 2 and 3+ rows, missing users, equal timestamps (stable), negative timestamps."""
    from gnn_recommendations_b200.dataset import temporal_split_device
    from gnn_recommendations_b200.synthetic import temporal_split
    rng = np.random.default_rng(5)
    n_users = 400
    rows = []
    for u in range(n_users):
        k = [0, 1, 2, 3, 7][u % 5] if u % 11 else 40
        rows += [u] * k
    user = np.asarray(rows, dtype=np.int64)
    perm = rng.permutation(len(user))
    user = user[perm]
    item = rng.integers(0, 1000, len(user)).astype(np.int64)
    ts = rng.integers(-50, 50, len(user)).astype(np.int64)          # many ties inside a user
    want = temporal_split(user, item, ts)                            # np.lexsort: stable, as pandas
    got = temporal_split_device(user, item, ts, n_users, device=DEV)
    for part in ("train", "valid", "test"):
        for a, b in zip(got[part], want[part]):
            assert np.array_equal(a.cpu().numpy(), b), part
    with pytest.raises(ValueError):
        temporal_split_device(np.array([0, n_users]), np.array([1, 2]), np.array([0, 1]), n_users, device=DEV)
    empty = temporal_split_device(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64), 5, device=DEV)
    assert all(len(a) == 0 for pair in empty.values() for a in pair)


def test_interaction_store_round_trip(tiny, tmp_path):
    """save_processed / load_processed: the reference's on-disk contract (dataset.py:366-394, 461-466,
    493-522): TSV `userId<TAB>itemId` without header, stats.json keys, scipy npz graphs."""
    import json
    import scipy.sparse as sp
    ds = dataset_from(tiny)
    ds.processed_data_path = tmp_path / "data" / "processed" / "tiny"
    ds.graphs_path = tmp_path / "data" / "graphs" / "tiny"
    ds.save_processed()
    first = (ds.processed_data_path / "train.txt").read_text().splitlines()[0]
    assert first == f"{int(tiny['train_u'][0])}\t{int(tiny['train_i'][0])}"
    stats = json.loads((ds.processed_data_path / "stats.json").read_text())
    assert stats["n_users"] == int(tiny["n_users"]) and stats["train_size"] == len(tiny["train_u"])
    assert stats["valid_size"] == len(tiny["valid_u"]) and stats["test_size"] == len(tiny["test_u"])
    back = g.InteractionDataset.load_processed("tiny", root_dir=str(tmp_path), device=DEV)
    for name in ("train", "valid", "test"):
        assert getattr(back, f"{name}_data").equals(getattr(ds, f"{name}_data"))
    assert back.n_users == ds.n_users and back.n_items == ds.n_items
    norm = sp.load_npz(str(ds.graphs_path / "norm_adj_matrix.npz"))
    assert norm.format == "coo" and norm.dtype == np.float32 and norm.shape == (ds.n_users + ds.n_items,) * 2
    assert np.array_equal(norm.row, tiny["adj_row"]) and np.array_equal(norm.col, tiny["adj_col"])
    assert np.array_equal(norm.data, tiny["adj_val"])
    raw = sp.load_npz(str(ds.graphs_path / "adj_matrix.npz"))
    assert raw.nnz == norm.nnz and np.all(raw.data == 1.0)


def test_ultragcn_under_trainer(tiny, tmp_path):
    """UltraGCN has no propagation (ultragcn.py:75-91): under the Trainer it is the BPR step + top-K on
    the raw tables, i.e. exactly LightGCN with zero layers — same seed, same epoch, same weights."""
    ds = dataset_from(tiny)
    nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
    out = []
    for cls, kw in ((g.UltraGCN, {}), (g.LightGCN, {"n_layers": 0})):
        torch.manual_seed(42)
        m = cls(nu, ni, embedding_dim=64, init_scale=0.1, **kw)
        tr = g.Trainer(m, ds, dict(CFG, checkpoint_dir=str(tmp_path / cls.__name__)), device=torch.device(DEV))
        torch.manual_seed(123)
        loss = tr.train_epoch()
        assert np.isfinite(loss)
        out.append((loss, m.user_embedding.weight.detach().cpu().numpy().copy(), tr.validate()))
    # the BPR gradient scatter uses red.global.add (order not fixed), so equal up to a few ulps
    assert abs(out[0][0] - out[1][0]) <= 1e-7
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-5, atol=1e-7)
    assert set(out[0][2]) == set(out[1][2])
    assert not np.array_equal(out[0][1], tiny["lightgcn/user_embedding.weight"])      # it trained


# ----------------------------------------------------------------------------- one reference epoch, other model families
@pytest.mark.parametrize("name", ["ngcf", "gat", "orthogonal_bundle"])
def test_trainer_epoch_other_models_vs_reference(name, tiny, tmp_path):
    """One full Trainer epoch (11 steps: sampler, forward, B x B BPR, backward, clip, Adam) from the
    reference's initial parameters must land on the reference's post-epoch parameters
    (tests/golden/make_golden_epochs.py; dropout 0 — the reference's dropout uses the CPU generator)."""
    import os
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "epochs.npz")))
    nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
    make = {"ngcf": lambda: g.NGCF(nu, ni, embedding_dim=64, dropout=0.0, init_scale=0.1),
            "gat": lambda: g.GAT(nu, ni, embedding_dim=64, n_layers=2, n_heads=4, dropout=0.0, init_scale=0.1),
            "orthogonal_bundle": lambda: g.OrthogonalBundleGNN(nu, ni, embedding_dim=64, n_layers=3, block_size=8,
                                                                dropout=0.0, init_scale=0.1)}[name]
    torch.manual_seed(42)
    m = make()
    init = {k[len(name) + 6:]: torch.from_numpy(v) for k, v in z.items() if k.startswith(f"{name}/init/")}
    assert set(init) == set(m.state_dict()), (sorted(set(init) ^ set(m.state_dict())))
    m.load_state_dict(init)
    tr = g.Trainer(m, dataset_from(tiny), dict(CFG, checkpoint_dir=str(tmp_path / "ckpt")), device=torch.device(DEV))
    torch.manual_seed(123)
    loss = tr.train_epoch()
    want_loss = float(z[f"{name}/loss"])
    assert abs(loss - want_loss) <= 2e-5 * abs(want_loss), (loss, want_loss)
    for k, v in m.state_dict().items():
        want = z[f"{name}/epoch/{k}"]
        if not np.issubdtype(want.dtype, np.floating):
            assert np.array_equal(v.cpu().numpy(), want), k
            continue
        moved = np.abs(want - z[f"{name}/init/{k}"]).max()
        np.testing.assert_allclose(v.detach().cpu().numpy(), want, rtol=1e-4, atol=max(2e-6, 1e-3 * float(moved)), err_msg=k)


def test_topk_tensor_core_kernel_failure_falls_back_to_exact(monkeypatch):
    """If the nomination kernel cannot run (shared-memory window not 1 KB aligned; forced here with the
    debug bit) the re-scoring pass flags every row and the exact kernel ranks them: same lists, no trap."""
    rng = np.random.default_rng(3)
    ue = rng.standard_normal((200, 64)).astype(np.float32)
    ie = rng.standard_normal((3000, 64)).astype(np.float32)
    eu = np.arange(200)
    want = g.full_rank_topk(torch.from_numpy(ue).to(DEV), torch.from_numpy(ie).to(DEV), eu, None, None, 20,
                            tensor_cores=False)
    monkeypatch.setenv("GR_TC_DEBUG", "4")
    stats = {}
    got = g.full_rank_topk(torch.from_numpy(ue).to(DEV), torch.from_numpy(ie).to(DEV), eu, None, None, 20,
                           tensor_cores=True, stats=stats)
    assert stats["rows_reranked_exactly"] == 200 and torch.equal(got, want)


def test_topk_tensor_core_item_tiles_by_tma_from_a_strided_table():
    """The item tiles reach shared memory by cp.async.bulk.tensor (SWIZZLE_128B boxes): a row-strided item table
    (ld > d) whose catalogue size is no multiple of the tile (the hardware zero-fills the tail rows) gives the
    exact kernel's lists, with nearly every row proven by the nomination pass."""
    rng = np.random.default_rng(11)
    nu, ni, d = 700, 10007, 64
    ue = torch.from_numpy(rng.standard_normal((nu, d)).astype(np.float32)).to(DEV)
    wide = torch.from_numpy(rng.standard_normal((ni, d + 32)).astype(np.float32)).to(DEV)
    ie = wide[:, 16:16 + d]                                   # row stride 96 floats, 64-byte offset
    eu = np.arange(nu)
    lens = rng.integers(0, 60, nu)
    ip = np.concatenate([[0], np.cumsum(lens)])
    it = np.concatenate([np.sort(rng.choice(ni, n, replace=False)) for n in lens]).astype(np.int32)
    want = g.full_rank_topk(ue, ie, eu, ip, it, 20, tensor_cores=False)
    stats = {}
    got = g.full_rank_topk(ue, ie, eu, ip, it, 20, tensor_cores=True, stats=stats)
    assert torch.equal(got, want)
    assert stats["tensor_cores"] and stats["rows_reranked_exactly"] <= nu // 20, stats


def test_trainer_full_loop_vs_reference(tiny, tmp_path):
    """`Trainer.train()` (trainer.py:469-579): 2-epoch linear warm-up, cosine schedule, validation every
    epoch, early-stopping bookkeeping — 4 epochs of LightGCN from the same seeds as the reference run
    (tests/golden/make_golden_trainloop.py)."""
    import os
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "trainloop.npz")))
    ds = dataset_from(tiny)
    cfg = {"learning_rate": 5e-3, "weight_decay": 1e-4, "batch_size": 512, "epochs": 4, "eval_every": 1,
           "use_scheduler": True, "warmup_epochs": 2, "max_grad_norm": 1.0, "negative_samples": 1,
           "validation_metrics": ["recall@10", "ndcg@10", "recall@20"], "early_stopping_metric": "recall@10",
           "early_stopping": {"patience": 3, "min_delta": 0.0001}, "model_name": "lightgcn_golden",
           "checkpoint_dir": str(tmp_path / "ckpt")}
    torch.manual_seed(42)
    m = g.LightGCN(int(tiny["n_users"]), int(tiny["n_items"]), embedding_dim=64, n_layers=3, init_scale=0.1)
    tr = g.Trainer(m, ds, cfg, device=torch.device(DEV))
    lrs, orig = [], tr.train_epoch

    def spy():
        lrs.append(tr.optimizer.param_groups[0]["lr"])
        return orig()

    tr.train_epoch = spy
    torch.manual_seed(123)
    res = tr.train()
    np.testing.assert_allclose(lrs, z["lrs"], rtol=1e-12)
    np.testing.assert_allclose(res["train_losses"], z["train_losses"], rtol=2e-5)
    np.testing.assert_allclose(m.user_embedding.weight.detach().cpu().numpy(), z["user_w"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(m.item_embedding.weight.detach().cpu().numpy(), z["item_w"], rtol=2e-3, atol=2e-5)
    assert len(res["valid_metrics"]) == 4
    n_eval = len(np.unique(tiny["valid_u"]))
    for ep, vm in enumerate(res["valid_metrics"]):
        for k, v in vm.items():                     # weights agree to ~1e-5, so at most a near-tie may flip
            assert abs(v - float(z[f"valid/{ep}/{k}"])) <= 2.0 / n_eval, (ep, k, v)
    assert abs(res["best_metric"] - float(z["best_metric"])) <= 2.0 / n_eval
    assert (tmp_path / "ckpt").exists() and any((tmp_path / "ckpt").iterdir())


def test_export_embeddings_and_warm_start_round_trip(tiny, tmp_path):
    """trainer.py:362-468: the exported payload holds the PROPAGATED embeddings (not the raw tables) under
    the reference's keys, and warm-start copies them into user/item_embedding.weight; `apply_to` filters
    by model name; a shape mismatch raises."""
    ds = dataset_from(tiny)
    nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
    path = tmp_path / "emb" / "lightgcn_tiny.pt"
    cfg = dict(CFG, epochs=1, checkpoint_dir=str(tmp_path / "ckpt"), model_name="lightgcn", dataset_name="tiny",
               export_embeddings={"enabled": True, "apply_to": "lightgcn", "path": str(path)})
    torch.manual_seed(42)
    m = g.LightGCN(nu, ni, embedding_dim=64, n_layers=3, init_scale=0.1)
    tr = g.Trainer(m, ds, cfg, device=torch.device(DEV))
    torch.manual_seed(123)
    tr.train()
    payload = torch.load(path)
    assert set(payload) == {"user_embedding", "item_embedding", "source_model", "dataset", "embedding_dim", "n_users", "n_items"}
    assert (payload["source_model"], payload["dataset"], payload["embedding_dim"], payload["n_users"], payload["n_items"]) == \
        ("lightgcn", "tiny", 64, nu, ni)
    with torch.no_grad():
        ue, ie = m.get_all_embeddings(ds.get_torch_adjacency())
    assert torch.equal(payload["user_embedding"], ue.cpu()) and torch.equal(payload["item_embedding"], ie.cpu())
    assert not torch.equal(payload["user_embedding"], m.user_embedding.weight.detach().cpu())
    # warm start into another model family
    ws = {"enabled": True, "apply_to": "orthogonal_bundle", "embeddings_path": str(path)}
    m2 = g.OrthogonalBundleGNN(nu, ni, embedding_dim=64)
    g.Trainer(m2, ds, dict(CFG, checkpoint_dir=str(tmp_path / "c2"), model_name="orthogonal_bundle", warm_start=ws),
              device=torch.device(DEV))
    assert torch.equal(m2.user_embedding.weight.detach().cpu(), payload["user_embedding"])
    m3 = g.NGCF(nu, ni, embedding_dim=64)
    before = m3.user_embedding.weight.detach().clone()
    g.Trainer(m3, ds, dict(CFG, checkpoint_dir=str(tmp_path / "c3"), model_name="ngcf", warm_start=ws), device=torch.device(DEV))
    assert torch.equal(m3.user_embedding.weight.detach().cpu(), before)          # apply_to filtered it out
    with pytest.raises(ValueError):
        g.Trainer(g.LightGCN(nu, ni, embedding_dim=32), ds,
                  dict(CFG, checkpoint_dir=str(tmp_path / "c4"), warm_start=dict(ws, apply_to=None)), device=torch.device(DEV))


# ----------------------------------------------------------------------------- item-sharded evaluation, ranks emulated
@pytest.mark.parametrize("world,tc", [(2, False), (4, True), (8, True), (3, None)])
def test_item_sharded_topk_partials_merge_to_the_single_gpu_lists(world, tc):
    """SURVEY §8e / BASELINE configs[3]: every rank ranks ALL users against its item id range, the per-rank
    top-k (score, id) lists are merged under (score desc, id asc).  The ranks are emulated on one GPU (the
    all-gather is a concatenation); exact partial kernel and the tcgen05 nomination path on each shard (seen CSR
    cut to the shard) must both merge to the bit-identical single-GPU lists, incl. users whose seen items empty
    a shard."""
    from gnn_recommendations_b200.dist import full_rank_topk_sharded, item_shard, merge_topk_partials
    rng = np.random.default_rng(world)
    nu, ni, d, k = 700, 20000, 64, 20
    ue = torch.from_numpy(rng.standard_normal((nu, d)).astype(np.float32)).to(DEV)
    ie = torch.from_numpy(rng.standard_normal((ni, d)).astype(np.float32)).to(DEV)
    ie[100:140] = ie[100]                                              # exact ties across a shard boundary or inside
    eu = np.arange(0, nu, 2)
    seen_u = rng.integers(0, nu, 30000)
    seen_i = rng.integers(0, ni, 30000)
    lo0, hi0 = item_shard(ni, world, 0)                                # user 0 has seen ALL of shard 0
    seen_u = np.concatenate([seen_u, np.zeros(hi0 - lo0, np.int64)])
    seen_i = np.concatenate([seen_i, np.arange(lo0, hi0)])
    from gnn_recommendations_b200.evaluator import seen_csr
    ip, it = seen_csr(eu, nu, (seen_u, seen_i))
    want = g.full_rank_topk(ue, ie, eu, ip, it, k, tensor_cores=False)
    parts_s, parts_i = [], []
    for r in range(world):
        lo, hi = item_shard(ni, world, r)
        ps, pi = full_rank_topk_sharded(ue, ie[lo:hi].contiguous(), lo, hi, eu, ip, it, k, world, tensor_cores=tc,
                                        return_partial=True)
        assert ps.shape[1:] == (len(eu), k) and pi.dtype == torch.int32
        parts_s.append(ps)
        parts_i.append(pi)
    got = merge_topk_partials(torch.cat(parts_s), torch.cat(parts_i), k)
    assert torch.equal(got, want)


def test_fused_clip_adam_more_than_32_tensors_with_empty_ones():
    """the chunk loop (32 tensors per launch) must not revisit tensors when a chunk skipped
    zero-numel tensors — 40 tensors with empties sprinkled in vs torch's clip_grad_norm_ + Adam."""
    gen = torch.Generator().manual_seed(0)
    shapes = [(0,) if i in (3, 17, 30, 31) else ((5 + i, 7) if i % 3 else (33 + i,)) for i in range(40)]
    ours = [torch.nn.Parameter(torch.randn(*s, generator=gen).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ours]
    o1 = torch.optim.Adam(ours, lr=1e-2, weight_decay=1e-4)
    o2 = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-4)
    for step in range(3):
        for p, q in zip(ours, ref):
            gq = torch.randn(*q.shape, generator=gen)
            q.grad = gq.clone()
            p.grad = gq.to(DEV)
        norm = g.fused_clip_adam_step(o1, 1.0)
        want = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o2.step()
        assert abs(float(norm) - float(want)) <= 1e-5 * float(want)
        for p, q in zip(ours, ref):
            torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=2e-6, atol=2e-7)


def test_cuda_graph_training_step_equals_eager(tiny, tmp_path):
    """The captured-and-replayed training step (Trainer.train_steps) must leave the same parameters, Adam state
    and mean loss as the eager loop: same sampler stream, same kernels, Adam's step-dependent scalars fed from
    the device.  (BPR gradient scatter uses red.global.add: order-dependent in the last bits.)"""
    nu, ni = int(tiny["n_users"]), int(tiny["n_items"])
    res = []
    for use_graph in (False, True):
        torch.manual_seed(42)
        m = g.LightGCN(nu, ni, embedding_dim=64, n_layers=3, init_scale=0.1)
        tr = g.Trainer(m, dataset_from(tiny), dict(CFG, cuda_graph=use_graph, checkpoint_dir=str(tmp_path / str(use_graph))),
                       device=torch.device(DEV))
        torch.manual_seed(123)
        l1 = tr.train_steps(40)
        l2 = tr.train_steps(25)                      # second call: the cached graph is replayed from step 0
        assert (getattr(tr, "_graph", None) is not None) == use_graph
        st = tr.optimizer.state[m.user_embedding.weight]
        res.append((l1, l2, m.user_embedding.weight.detach().cpu(), m.item_embedding.weight.detach().cpu(),
                    st["exp_avg_sq"].cpu(), float(st["step"])))
    a, b = res
    assert abs(a[0] - b[0]) <= 1e-6 * abs(a[0]) and abs(a[1] - b[1]) <= 1e-6 * abs(a[1])
    assert a[5] == b[5] == 65.0
    for x, y in zip(a[2:5], b[2:5]):
        torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-6)


def test_bpr_fused_large_batch_and_bad_ids():
    """batches larger than one cooperative wave (B > ~9.4 k) used to return GR_ERR_UNSUPPORTED — the
    kernel now strides over the samples; an out-of-range id poisons the loss with NaN instead of writing out of
    bounds."""
    gen = torch.Generator().manual_seed(0)
    nu, ni, d, B = 3000, 2000, 64, 12000
    emb = (torch.randn(nu + ni, d, generator=gen) * 0.3).to(DEV).requires_grad_(True)
    users = torch.randint(0, nu, (B,), generator=gen).to(DEV)
    pos = torch.randint(0, ni, (B,), generator=gen).to(DEV)
    neg = torch.randint(0, ni, (B,), generator=gen).to(DEV)
    loss = g.bpr_fused(emb, nu, users, pos, neg)
    (grad,) = torch.autograd.grad(loss, [emb])
    e64 = emb.detach().double().requires_grad_(True)
    p = (e64[users] * e64[nu + pos]).sum(1)
    n = (e64[users] * e64[nu + neg]).sum(1)
    ref = torch.nn.functional.softplus(n[:, None] - p[None, :]).mean()
    (gref,) = torch.autograd.grad(ref, [e64])
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    np.testing.assert_allclose(grad.cpu().numpy(), gref.cpu().numpy(), rtol=1e-4, atol=1e-6 * float(gref.abs().max()))
    bad = users.clone()
    bad[7] = nu + 5                                                # a user id beyond n_users
    before = torch.cuda.memory_allocated()
    loss_bad = g.bpr_fused(emb.detach(), nu, bad, pos, neg)
    assert torch.isnan(loss_bad) and before >= 0
    bad_item = pos.clone()
    bad_item[3] = -1
    assert torch.isnan(g.bpr_fused(emb.detach(), nu, users, bad_item, neg))


def test_topk_rejects_k_above_kernel_limit():
    """K > 64 must fail up front with a clear ValueError (it used to surface as a GrError from the
    kernel behind a dead host-metrics branch)."""
    ue = torch.randn(10, 64, device=DEV)
    ie = torch.randn(500, 64, device=DEV)
    with pytest.raises(ValueError, match="at most 64"):
        g.full_rank_topk(ue, ie, np.arange(10), None, None, 100)
    assert g.full_rank_topk(ue, ie, np.arange(10), None, None, 64).shape == (10, 64)
