"""CPU-only tests of the host-side logic around the kernels: sampler (host C function), metrics,
seen-set construction, row partition, dataset boundary."""
import numpy as np
import pytest
import torch

import gnn_recommendations_b200 as g
from gnn_recommendations_b200.dist import RowPartition
from gnn_recommendations_b200.evaluator import ground_truth_dict, seen_csr
from oracle import pyoracle as po


def test_sampler_bit_exact_vs_reference_trainer(tiny):
    s = g.BprSampler(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]))
    torch.manual_seed(123)
    for b in range(3):
        u, p, n = s.sample(512)
        assert np.array_equal(u, tiny[f"batch{b}/users"])
        assert np.array_equal(p, tiny[f"batch{b}/pos"])
        assert np.array_equal(n, tiny[f"batch{b}/neg"].reshape(-1))
    # the global generator is left exactly where the reference's python loop leaves it
    nxt = torch.randint(0, 1 << 30, (4,))
    torch.manual_seed(123)
    gen = po.TorchCpuMt19937(123)
    ps = po.positive_sets(tiny["train_u"], tiny["train_i"])
    for _ in range(3):
        po.sample_batch(gen, tiny["train_u"], tiny["train_i"], int(tiny["n_items"]), 512, ps)
    assert np.array_equal(nxt.numpy(), gen.randint(1 << 30, 4))


def test_sampler_c1_shape(c1gold, c1split):
    tu, ti = c1split["train"]
    s = g.BprSampler(tu, ti, c1split["n_users"], c1split["n_items"])
    torch.manual_seed(2024)
    for b in range(2):
        u, p, n = s.sample(512)
        assert np.array_equal(u, c1gold[f"batch{b}/users"])
        assert np.array_equal(p, c1gold[f"batch{b}/pos"])
        assert np.array_equal(n, c1gold[f"batch{b}/neg"].reshape(-1))


def test_sampler_many_batches_cross_twist_boundary(tiny):
    # > 624 draws per batch: the mt19937 block regeneration must agree with torch across many batches
    s = g.BprSampler(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]))
    ps = po.positive_sets(tiny["train_u"], tiny["train_i"])
    torch.manual_seed(7)
    gen = po.TorchCpuMt19937(7)
    for _ in range(12):
        u, p, n = s.sample(300)
        ou, op, on = po.sample_batch(gen, tiny["train_u"], tiny["train_i"], int(tiny["n_items"]), 300, ps)
        assert np.array_equal(u, ou) and np.array_equal(p, op) and np.array_equal(n, on.reshape(-1))


def test_sampler_dense_user_hits_the_ten_redraw_cap():
    # user 0 has every item but one as a positive: most draws are rejected, the 10th redraw is kept unchecked
    n_items = 12
    tu = np.zeros(n_items - 1, dtype=np.int64)
    ti = np.arange(n_items - 1, dtype=np.int64)
    s = g.BprSampler(tu, ti, 1, n_items)
    ps = po.positive_sets(tu, ti)
    torch.manual_seed(3)
    gen = po.TorchCpuMt19937(3)
    u, p, n = s.sample(64)
    ou, op, on = po.sample_batch(gen, tu, ti, n_items, 64, ps)
    assert np.array_equal(n, on.reshape(-1)) and np.array_equal(p, op)
    assert (n != n_items - 1).any()          # some accepted negatives are actually positives (reference quirk)


def test_sampler_large_ranges_use_64_bit_draws():
    # torch.randint switches to two 32-bit outputs per element for ranges >= 2^28
    n_items = (1 << 28) + 5
    tu = np.zeros(4, dtype=np.int64)
    ti = np.array([1, 2, 3, 4], dtype=np.int64)
    s = g.BprSampler(tu, ti, 1, n_items)
    torch.manual_seed(9)
    u, p, n = s.sample(4)
    torch.manual_seed(9)
    idx = torch.randint(0, 4, (4,)).numpy()
    want = [int(torch.randint(0, n_items, (1,)).item()) for _ in range(4)]
    assert np.array_equal(p, ti[idx]) and n.tolist() == want


def test_metrics_match_reference_values(tiny):
    gt = ground_truth_dict((tiny["test_u"], tiny["test_i"]))
    m = g.compute_metrics_from_topk(torch.from_numpy(tiny["eval/topk20_canonical"]), tiny["eval/users"].tolist(), gt,
                                    int(tiny["n_items"]), [10, 20])
    for k in ("recall", "ndcg", "precision", "coverage", "gini"):
        for kk in (10, 20):
            assert abs(m[f"{k}@{kk}"] - float(tiny[f"evaluate/{k}@{kk}"])) <= 1e-12, (k, kk)
    assert g.compute_metrics_from_topk(torch.zeros((0, 20), dtype=torch.int64), [], {}, 10) == {}


def test_seen_csr_union_sorted_unique():
    ip, it = seen_csr([2, 5], 8, (np.array([5, 2, 5, 7]), np.array([9, 3, 1, 4])), (np.array([5, 2]), np.array([9, 0])))
    assert ip.tolist() == [0, 2, 4] and it.tolist() == [0, 3, 1, 9]


def test_row_partition_covers_and_balances():
    rng = np.random.default_rng(0)
    lens = np.concatenate([[5000], rng.integers(0, 50, 500), rng.integers(0, 5, 1003)])   # popularity-ordered
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))
    n = len(lens)
    for g_ in (1, 2, 3, 8):
        part = RowPartition(indptr, g_)
        ids = torch.arange(n)
        padded = part.to_padded(ids)
        assert padded.unique().numel() == n and int(padded.max()) < part.padded_rows
        owned = [part.local_ids(r) for r in range(g_)]
        assert torch.equal(torch.sort(torch.cat(owned)).values, ids)                # a partition of the rows
        for r in range(g_):
            assert part.n_local(r) == owned[r].numel() and abs(part.n_local(r) - n / g_) < 1   # rows balanced
            assert torch.equal(padded[owned[r]], torch.arange(part.n_local(r)) + r * part.block_rows)
            x = torch.arange(n * 2, dtype=torch.float32).view(n, 2)
            assert torch.equal(part.take_rows(x, r), x[owned[r]])


def test_dataset_boundary_attributes(tiny):
    ds = g.InteractionDataset((tiny["train_u"], tiny["train_i"]), (tiny["valid_u"], tiny["valid_i"]),
                              (tiny["test_u"], tiny["test_i"]), int(tiny["n_users"]), int(tiny["n_items"]))
    assert list(ds.train_data.columns) == ["userId", "itemId"] and len(ds.train_data) == len(tiny["train_u"])
    assert ds.n_users == 300 and ds.n_items == 200


# ----------------------------------------------------------------------------- UltraGCN (no propagation)
def test_ultragcn_matches_reference_golden():
    """Same-seed parameters, predict, compute_loss (BPR + constraint + L2) and its gradients vs the
    unmodified reference (tests/golden/make_golden_ultragcn.py; ultragcn.py:21-247)."""
    import os
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ultragcn.npz")))
    nu, ni, d = int(z["n_users"]), int(z["n_items"]), int(z["d"])
    torch.manual_seed(42)
    m = g.UltraGCN(nu, ni, embedding_dim=d, lambda_1=0.7, lambda_2=1.3, gamma=1e-3, init_scale=0.1)
    assert np.array_equal(m.user_embedding.weight.detach().numpy(), z["user_w"])
    assert np.array_equal(m.item_embedding.weight.detach().numpy(), z["item_w"])
    users, pos, neg = (torch.from_numpy(z[k]) for k in ("users", "pos", "neg"))
    np.testing.assert_allclose(m.predict(users, pos).detach().numpy(), z["predict"], rtol=1e-6, atol=1e-8)
    ue, ie = m.get_all_embeddings(None)
    assert ue is m.user_embedding.weight and ie is m.item_embedding.weight
    adj = torch.from_numpy(z["adj"])
    for tag, a in (("noadj", None), ("dense", adj), ("sparse", adj.to_sparse())):
        m.zero_grad()
        total, parts = m.compute_loss(users, pos, neg, a)
        total.backward()
        assert abs(total.item() - float(z[f"{tag}/total"])) <= 1e-6 * abs(float(z[f"{tag}/total"])), tag
        assert abs(parts["constraint_loss"] - float(z[f"{tag}/constraint"])) <= 1e-6 * max(1e-3, abs(float(z[f"{tag}/constraint"])))
        np.testing.assert_allclose(m.user_embedding.weight.grad.numpy(), z[f"{tag}/grad_user"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(m.item_embedding.weight.grad.numpy(), z[f"{tag}/grad_item"], rtol=1e-5, atol=1e-7)
    assert g.MODEL_REGISTRY["ultragcn"] is g.UltraGCN


# ----------------------------------------------------------------------------- host helpers of the later stages
def test_tc_kprime_and_splits_rules():
    from gnn_recommendations_b200.evaluator import choose_splits, tc_kprime
    assert [tc_kprime(k) for k in (1, 10, 20, 24, 32, 50, 52, 60)] == [24, 24, 32, 40, 48, 64, 64, 64]
    assert all(tc_kprime(k) >= k + 8 for k in range(1, 57))
    assert choose_splits(52643, 91599) == 1 and choose_splits(29, 91599) == 64 and choose_splits(64, 600) == 2


def test_discount_tables_follow_the_reference_loops():
    from gnn_recommendations_b200.metrics import discount_tables
    disc, idcg = discount_tables(50)
    acc = 0.0
    for rank in range(50):
        assert disc[rank] == 1.0 / np.log2(rank + 2)          # metrics.py:404
        acc += 1.0 / np.log2(rank + 2)                        # metrics.py:405-409
        assert idcg[rank + 1] == acc
    assert idcg[0] == 0.0


def test_later_stages_refuse_cpu_inputs():
    """No CPU fallback anywhere on the product path."""
    from gnn_recommendations_b200.dataset import temporal_split_device
    from gnn_recommendations_b200.optim import fused_clip_adam_supported
    with pytest.raises(RuntimeError):
        g.topk_metrics_device(torch.zeros((4, 5), dtype=torch.int64), np.zeros(5, np.int64), np.zeros(0, np.int32), 10)
    with pytest.raises(RuntimeError):
        temporal_split_device(np.zeros(3, np.int64), np.zeros(3, np.int64), np.zeros(3, np.int64), 4, device="cpu")
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    assert not fused_clip_adam_supported(torch.optim.Adam([p]))          # CPU tensors: stock path only
    assert not fused_clip_adam_supported(torch.optim.SGD([p], lr=0.1))


def test_sm_copy_argument_checks_need_no_gpu():
    """sm_copy validates its operands before it touches the library: pageable host memory, size mismatches, strided
    views and a copy without a CUDA side are refused (the kernel stores to / loads from PINNED host memory)."""
    import pytest
    import torch

    import gnn_recommendations_b200 as g
    a, b = torch.zeros(64), torch.zeros(64)
    with pytest.raises(ValueError, match="pinned"):
        g.sm_copy(a, b)
    with pytest.raises(ValueError, match="size"):
        g.sm_copy(torch.zeros(32), b)
    with pytest.raises(ValueError, match="contiguous"):
        g.sm_copy(torch.zeros(8, 16)[:, ::2], torch.zeros(64))
