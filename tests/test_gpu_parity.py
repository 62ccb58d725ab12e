"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures of the reference.  Integer/index results bit-exact; LightGCN embeddings bit-exact;
gradients within 1e-5 relative."""
import hashlib

import numpy as np
import pytest
import torch

import gnn_recommendations_b200 as g
from gnn_recommendations_b200 import _lib
from oracle import coracle
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def tiny_ref(tiny):
    return po.build_norm_adj(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]))


@pytest.fixture(scope="module")
def tiny_csr(tiny):
    return g.NormAdjCSR.from_pairs(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]),
                                   device=DEV, dis_lut=tiny["dis_lut"])


def assert_csr_equal(csr, ref):
    assert np.array_equal(csr.indptr.cpu().numpy().astype(np.int64), ref["indptr"])
    assert np.array_equal(csr.indices.cpu().numpy(), ref["indices"])
    assert np.array_equal(bits(csr.vals.cpu().numpy()), bits(ref["vals"]))


# ------------------------------------------------------------------------- graph builder
def test_build_from_pairs_bit_exact_vs_golden(tiny, tiny_csr):
    assert np.array_equal(tiny_csr.indices.cpu().numpy(), tiny["adj_col"])
    assert np.array_equal(bits(tiny_csr.vals.cpu().numpy()), bits(tiny["adj_val"]))
    assert np.array_equal(tiny_csr.row_ids().cpu().numpy(), tiny["adj_row"].astype(np.int64))
    assert np.array_equal(tiny_csr.deg.cpu().numpy(), tiny["deg"].astype(np.int32))


def test_build_row_normalisation(tiny):
    csr = g.NormAdjCSR.from_pairs(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]),
                                  normalization="row", device=DEV, dis_lut=tiny["dinv_lut"])
    assert np.array_equal(bits(csr.vals.cpu().numpy()), bits(tiny["adjrow_val"]))


def test_build_duplicates_isolated_and_empty():
    u = np.array([0, 0, 1, 0]); i = np.array([0, 0, 0, 0])
    assert_csr_equal(g.NormAdjCSR.from_pairs(u, i, 2, 2, device=DEV), po.build_norm_adj(u, i, 2, 2))
    e = np.zeros(0, dtype=np.int64)
    csr = g.NormAdjCSR.from_pairs(e, e, 3, 2, device=DEV)
    assert csr.nnz == 0 and csr.indptr.cpu().tolist() == [0] * 6
    with pytest.raises(ValueError):
        g.NormAdjCSR.from_pairs(np.array([5]), np.array([0]), 2, 2, device=DEV)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_build_random_graphs_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    nu, ni, e = int(rng.integers(1, 400)), int(rng.integers(1, 300)), int(rng.integers(1, 20000))
    u = rng.integers(0, nu, e); i = rng.integers(0, ni, e)       # with duplicates, with isolated nodes
    assert_csr_equal(g.NormAdjCSR.from_pairs(u, i, nu, ni, device=DEV), po.build_norm_adj(u, i, nu, ni))


def test_coo_to_csr_paths(tiny_ref):
    coo = po.to_torch_coo(tiny_ref).to(DEV)
    assert_csr_equal(g.NormAdjCSR.from_torch_coo(coo), tiny_ref)
    first = g.as_csr(coo)
    assert g.as_csr(coo) is first                                  # same tensor object: identity fast path
    # the reference rebuilds the adjacency tensor every epoch / validate: a NEW tensor with the
    # same content must hit the cache (keyed by content, the COO tensor is not retained) ...
    import gc, weakref
    fresh = po.to_torch_coo(tiny_ref).to(DEV)
    assert fresh._indices().data_ptr() != coo._indices().data_ptr()
    assert g.as_csr(fresh) is first
    wr = weakref.ref(fresh)
    del fresh
    gc.collect()
    assert wr() is None                                            # ... and is not pinned in HBM by the cache
    # ... while different content (one value changed) must not
    other = po.to_torch_coo(tiny_ref).to(DEV)
    other._values()[7] *= 2.0
    assert g.as_csr(other) is not first
    # host CSR round trip (8 B per entry instead of the int64 COO's 20 B)
    ip, ix, vl = first.to_host()
    assert ip.is_pinned() and ip.dtype == torch.int32 and ix.dtype == torch.int32
    again = g.NormAdjCSR.from_host_csr(ip, ix, vl, device=DEV)
    assert_csr_equal(again, tiny_ref)
    xx = torch.randn(first.n_cols, 64, generator=torch.Generator().manual_seed(0)).to(DEV)
    assert torch.equal(again.spmm(xx)[0], first.spmm(xx)[0])
    # unsorted rows: storage order inside a row must be preserved
    perm = torch.randperm(coo._values().numel(), generator=torch.Generator().manual_seed(1)).to(DEV)
    shuffled = torch.sparse_coo_tensor(coo._indices()[:, perm], coo._values()[perm], coo.shape)
    csr = g.NormAdjCSR.from_torch_coo(shuffled)
    assert np.array_equal(csr.indptr.cpu().numpy().astype(np.int64), tiny_ref["indptr"])
    bad = torch.sparse_coo_tensor(torch.tensor([[0], [99]]), torch.tensor([1.0]), (4, 4)).to(DEV)
    with pytest.raises((ValueError, RuntimeError)):
        g.NormAdjCSR.from_torch_coo(bad)


def test_row_schedule_is_lpt(tiny_csr):
    lens = (tiny_csr.indptr[1:] - tiny_csr.indptr[:-1]).cpu().numpy()
    order = tiny_csr.row_order.cpu().numpy()
    assert sorted(order.tolist()) == list(range(tiny_csr.n_rows))
    expect = np.lexsort((np.arange(len(lens)), -lens))
    assert np.array_equal(order, expect)
    assert tiny_csr.n_long == int((lens >= tiny_csr.long_threshold).sum())


# ------------------------------------------------------------------------- SpMM
@pytest.mark.parametrize("d", [32, 64, 128, 256])
@pytest.mark.parametrize("long_threshold", [1 << 30, 16, 1])
def test_spmm_bit_exact(tiny_ref, d, long_threshold):
    csr = g.NormAdjCSR(torch.from_numpy(tiny_ref["indptr"].astype(np.int32)).to(DEV),
                       torch.from_numpy(tiny_ref["indices"]).to(DEV), torch.from_numpy(tiny_ref["vals"]).to(DEV),
                       tiny_ref["n"], tiny_ref["n"], long_threshold=long_threshold)
    x = torch.randn(tiny_ref["n"], d, generator=torch.Generator().manual_seed(d))
    want = coracle.spmm_fmaf(tiny_ref["indptr"], tiny_ref["indices"], tiny_ref["vals"], x.numpy())
    y, _ = csr.spmm(x.to(DEV))
    assert np.array_equal(bits(y.cpu().numpy()), bits(want))
    # epilogue: out = (addend + t) / 4, y not stored
    add = torch.randn(tiny_ref["n"], d, generator=torch.Generator().manual_seed(7))
    _, out = csr.spmm(x.to(DEV), addend=add.to(DEV), scale=4.0, scale_mode=_lib.GR_SCALE_DIV, want_y=False)
    assert np.array_equal(bits(out.cpu().numpy()), bits((add.numpy() + want) / np.float32(4.0)))


def test_spmm_natural_order_and_strided_views(tiny_ref):
    csr = g.NormAdjCSR(torch.from_numpy(tiny_ref["indptr"].astype(np.int32)).to(DEV),
                       torch.from_numpy(tiny_ref["indices"]).to(DEV), torch.from_numpy(tiny_ref["vals"]).to(DEV),
                       tiny_ref["n"], tiny_ref["n"])
    csr.row_order, csr.n_long = None, 0
    big = torch.randn(tiny_ref["n"], 256, generator=torch.Generator().manual_seed(3)).to(DEV)
    x = big[:, 64:128]                                           # ld = 256, d = 64
    want = coracle.spmm_fmaf(tiny_ref["indptr"], tiny_ref["indices"], tiny_ref["vals"], x.cpu().numpy())
    outbuf = torch.zeros(tiny_ref["n"], 256, device=DEV)
    csr.spmm(x, y=outbuf[:, 128:192])
    assert np.array_equal(bits(outbuf[:, 128:192].cpu().numpy()), bits(want))
    assert float(outbuf[:, :128].abs().sum()) == 0.0 and float(outbuf[:, 192:].abs().sum()) == 0.0


def test_spmm_hot_rows_power_law():
    # one item adjacent to every user (row of length 5000) + ragged tails + empty rows
    rng = np.random.default_rng(5)
    nu, ni = 5000, 300
    u = np.concatenate([np.arange(nu), rng.integers(0, nu, 40000)])
    i = np.concatenate([np.zeros(nu, dtype=np.int64), (rng.pareto(1.0, 40000) * 3).astype(np.int64) % (ni - 10)])
    ref = po.build_norm_adj(u, i, nu, ni)
    for thr in (1024, 64):
        csr = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=DEV)
        csr.long_threshold = thr
        csr._schedule()
        assert csr.n_long >= 1
        for d in (64, 128):
            x = torch.randn(ref["n"], d, generator=torch.Generator().manual_seed(d))
            want = coracle.spmm_fmaf(ref["indptr"], ref["indices"], ref["vals"], x.numpy())
            y, _ = csr.spmm(x.to(DEV))
            assert np.array_equal(bits(y.cpu().numpy()), bits(want)), (thr, d)


def test_spmm_argument_errors(tiny_csr):
    with pytest.raises(_lib.GrError):
        tiny_csr.spmm(torch.zeros(tiny_csr.n_cols, 48, device=DEV))      # unsupported d
    with pytest.raises(ValueError):
        tiny_csr.spmm(torch.zeros(3, 64, device=DEV))


# ------------------------------------------------------------------------- LightGCN
@pytest.mark.parametrize("tag,d,L", [("lightgcn", 64, 3), ("lightgcn_d128_l4", 128, 4)])
def test_lightgcn_forward_bit_exact_vs_reference(tiny, tiny_csr, tag, d, L):
    m = g.LightGCN(int(tiny["n_users"]), int(tiny["n_items"]), d, L, 0.1)
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(tiny[f"{tag}/user_embedding.weight"]),
                       "item_embedding.weight": torch.from_numpy(tiny[f"{tag}/item_embedding.weight"])})
    m.to(DEV)
    with torch.no_grad():
        ue, ie = m.get_all_embeddings(tiny_csr)
    assert np.array_equal(bits(ue.cpu().numpy()), bits(tiny[f"{tag}/out_user"]))
    assert np.array_equal(bits(ie.cpu().numpy()), bits(tiny[f"{tag}/out_item"]))
    if tag == "lightgcn":
        layers = torch.stack(m.get_layer_embeddings(tiny_csr)).cpu().numpy()
        assert np.array_equal(bits(layers), bits(tiny["lightgcn/layers"]))
        # the torch-COO entry the reference Trainer uses gives the same bits
        coo = tiny_csr.to_torch_coo()
        with torch.no_grad():
            ue2, _ = m(coo)
        assert torch.equal(ue, ue2)


def test_lightgcn_step_gradients_vs_reference(tiny, tiny_csr):
    m = g.LightGCN(int(tiny["n_users"]), int(tiny["n_items"]), 64, 3, 0.1)
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(tiny["lightgcn/user_embedding.weight"]),
                       "item_embedding.weight": torch.from_numpy(tiny["lightgcn/item_embedding.weight"])})
    m.to(DEV)
    us, ps, ns = (torch.from_numpy(tiny[f"batch0/{k}"]).to(DEV) for k in ("users", "pos", "neg"))
    ue, ie = m.get_all_embeddings(tiny_csr)
    loss = po.bpr_loss_reference(ue, ie, us, ps, ns)          # torch ops on the device, reference formula
    loss.backward()
    assert abs(float(loss) - float(tiny["step0/loss"])) <= 1e-5 * abs(float(tiny["step0/loss"]))
    np.testing.assert_allclose(m.user_embedding.weight.grad.cpu().numpy(), tiny["step0/grad_user"],
                               rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(m.item_embedding.weight.grad.cpu().numpy(), tiny["step0/grad_item"],
                               rtol=1e-5, atol=1e-10)


def test_nonsymmetric_backward_uses_transpose(tiny):
    csr = g.NormAdjCSR.from_pairs(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]),
                                  normalization="row", device=DEV)
    ref = po.build_norm_adj(tiny["train_u"], tiny["train_i"], int(tiny["n_users"]), int(tiny["n_items"]), "row")
    m = g.LightGCN(int(tiny["n_users"]), int(tiny["n_items"]), 64, 2, 0.1).to(DEV)
    uw, iw = m.user_embedding.weight.detach().cpu(), m.item_embedding.weight.detach().cpu()
    ue, ie = m(csr)
    (ue.sum() * 2 + (ie * ie).sum()).backward()
    ruw, riw = uw.clone().requires_grad_(True), iw.clone().requires_grad_(True)
    rue, rie = po.lightgcn_forward(po.to_torch_coo(ref), ruw, riw, 2)
    (rue.sum() * 2 + (rie * rie).sum()).backward()
    assert torch.equal(ue.detach().cpu(), rue.detach())
    torch.testing.assert_close(m.user_embedding.weight.grad.cpu(), ruw.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(m.item_embedding.weight.grad.cpu(), riw.grad, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------- C1 shape (full size)
def test_c1_graph_and_forward_vs_reference_hashes(c1gold, c1split):
    tu, ti = c1split["train"]
    nu, ni = c1split["n_users"], c1split["n_items"]
    csr = g.NormAdjCSR.from_pairs(tu, ti, nu, ni, device=DEV, dis_lut=c1gold["dis_lut"])
    assert csr.nnz == int(c1gold["nnz"])
    assert sha(csr.row_ids().cpu().numpy().astype(np.int32)) == str(c1gold["sha_adj_row"])
    assert sha(csr.indices.cpu().numpy()) == str(c1gold["sha_adj_col"])
    assert sha(csr.vals.cpu().numpy()) == str(c1gold["sha_adj_val"])
    # weights: numpy PCG64 (platform independent), as in make_golden.np_init
    rng = np.random.default_rng(42)
    uw = (rng.standard_normal((nu, 64)) * 0.1).astype(np.float32)
    iw = (rng.standard_normal((ni, 64)) * 0.1).astype(np.float32)
    m = g.LightGCN(nu, ni, 64, 3, 0.1)
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(uw), "item_embedding.weight": torch.from_numpy(iw)})
    m.to(DEV)
    with torch.no_grad():
        ue, ie = m(csr)
    assert sha(ue.cpu().numpy()) == str(c1gold["sha_out_user"])
    assert sha(ie.cpu().numpy()) == str(c1gold["sha_out_item"])
    # size-independent property: linearity of the propagation in x (exact for scaling by 2)
    with torch.no_grad():
        x0 = torch.cat([m.user_embedding.weight, m.item_embedding.weight])
        a = g.lightgcn_propagate(csr, x0, 3)
        b = g.lightgcn_propagate(csr, x0 * 2.0, 3)
    assert torch.equal(a * 2.0, b)


def test_multi_gpu_bit_identical_when_two_gpus_present():
    """Spawns tests/multigpu_check.py under torchrun when the box has >= 2 GPUs."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(here, "multigpu_check.py"), "C1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "propagation_bit_identical=True topk_bit_identical=True" in r.stdout


def test_spmm_split_rows_deterministic_segment_sums():
    """Rows above the split threshold are cut into segments; the result must equal, bit for bit, the sum
    (in segment order) of the exact per-segment chains, and stay within 1e-6 of the single chain."""
    rng = np.random.default_rng(9)
    nu, ni = 6000, 50
    u = np.concatenate([np.arange(nu), rng.integers(0, nu, 20000)])
    i = np.concatenate([np.zeros(nu, dtype=np.int64), rng.integers(1, ni, 20000)])
    ref = po.build_norm_adj(u, i, nu, ni)
    csr = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=DEV)
    csr.split_threshold, csr.split_segment = 2048, 1000
    csr._schedule()
    assert csr.n_split >= 1 and csr.n_parts >= 6
    hot = nu                                            # item 0: adjacent to every user
    lens = np.diff(ref["indptr"])
    assert lens[hot] == nu
    for d in (64, 128):
        x = torch.randn(ref["n"], d, generator=torch.Generator().manual_seed(d))
        exact = coracle.spmm_fmaf(ref["indptr"], ref["indices"], ref["vals"], x.numpy())
        y1, _ = csr.spmm(x.to(DEV))
        y2, _ = csr.spmm(x.to(DEV))
        assert torch.equal(y1, y2)                       # deterministic
        got = y1.cpu().numpy()
        split = set(csr.split_rows[0].cpu().tolist())
        keep = np.array([r not in split for r in range(ref["n"])])
        assert np.array_equal(bits(got[keep]), bits(exact[keep]))          # unsplit rows: exact chain
        for r in split:
            s, e = ref["indptr"][r], ref["indptr"][r + 1]
            cuts = list(range(s, e, 1000)) + [e]
            ip = np.asarray(cuts, dtype=np.int64)
            parts = coracle.spmm_fmaf(ip - ip[0], ref["indices"][s:e], ref["vals"][s:e], x.numpy())
            tot = parts[0].copy()
            for p in parts[1:]:
                tot = tot + p
            assert np.array_equal(bits(got[r]), bits(tot))
            np.testing.assert_allclose(got[r], exact[r], rtol=1e-5, atol=1e-6)
        # epilogue on a split row
        add = torch.randn(ref["n"], d, generator=torch.Generator().manual_seed(1))
        _, out = csr.spmm(x.to(DEV), addend=add.to(DEV), scale=4.0, scale_mode=_lib.GR_SCALE_DIV, want_y=False)
        assert np.array_equal(bits(out.cpu().numpy()), bits((add.numpy() + got) / np.float32(4.0)))


@pytest.mark.parametrize("d", [32, 64])
def test_spmm_dense_map_epilogue_short_long_and_split_rows(d):
    """gr_spmm_csr_map_f32: out = alpha (A x) M + beta R and y = alpha (A x) from ONE kernel, on a graph with
    short rows, long rows (CTA-cooperative kernel) and split rows (combine pass): the aggregate is the exact
    fmaf chain of the plain SpMM (y bit-equal to alpha * t), the mapped row equals a float64 evaluation of
    t M to fp32 accuracy; M^T, the device-resident beta factor and the no-residual form are covered."""
    rng = np.random.default_rng(21 + d)
    nu, ni = 6000, 50
    # item 0: every user (split row); item 1: 1 500 users (long row, not split); the rest random (short rows)
    u = np.concatenate([np.arange(nu), np.arange(1500), rng.integers(0, nu, 20000)])
    i = np.concatenate([np.zeros(nu, dtype=np.int64), np.ones(1500, dtype=np.int64), rng.integers(2, ni, 20000)])
    csr = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=DEV)
    csr.split_threshold, csr.split_segment = 2048, 1000
    csr._schedule()
    assert csr.n_split >= 1 and csr.n_long > csr.n_split and csr.supports_map(d)
    n = nu + ni
    gen = torch.Generator().manual_seed(d)
    x = torch.randn(n, d, generator=gen).to(DEV)
    m = (torch.randn(d, d, generator=gen) / np.sqrt(d)).to(DEV)
    r = torch.randn(n, d, generator=gen).to(DEV)
    t, _ = csr.spmm(x)
    t64 = t.double()
    alpha, beta = 0.9, 0.1
    out, y = csr.spmm_map(x, m, alpha, beta, addend=r, want_y=True)
    assert np.array_equal(bits(y.cpu().numpy()), bits((t * np.float32(alpha)).cpu().numpy()))
    want = alpha * (t64 @ m.double()) + beta * r.double()
    scale = float(want.abs().max())
    assert float((out.double() - want).abs().max()) <= 2e-6 * scale
    # transposed map, beta read from device memory (0.1 * 0.5), no y
    bdev = torch.tensor([0.5], device=DEV)
    out_t, y_t = csr.spmm_map(x, m, alpha, beta, addend=r, beta_dev=bdev, transposed=True)
    assert y_t is None
    want_t = alpha * (t64 @ m.double().T) + beta * 0.5 * r.double()
    assert float((out_t.double() - want_t).abs().max()) <= 2e-6 * float(want_t.abs().max())
    # no residual
    out_n, _ = csr.spmm_map(x, m, 1.0, 0.0)
    assert float((out_n.double() - t64 @ m.double()).abs().max()) <= 2e-6 * float((t64 @ m.double()).abs().max())
    # deterministic
    out2, _ = csr.spmm_map(x, m, alpha, beta, addend=r)
    assert torch.equal(out, out2)
    with pytest.raises(ValueError):
        csr.spmm_map(torch.randn(n, 128, device=DEV), torch.eye(128, device=DEV))


# ----------------------------------------------------------------------------- BASELINE configs[1..3] at full size
@pytest.mark.parametrize("shape", ["C2", "C3", "C4"])
def test_full_size_dataset_shapes_vs_oracle(shape):
    """Gowalla / Yelp2018 / Amazon-Book shapes (BASELINE.json configs 1-3) end to end on the GPU:
    temporal split, graph build (CSR bit-exact), LightGCN propagation (bit-exact), full-ranking top-20
    of a user sample by the exact and the tensor-core path (identical to the C oracle)."""
    from gnn_recommendations_b200.dataset import temporal_split_device
    from gnn_recommendations_b200.evaluator import seen_csr
    from gnn_recommendations_b200.synthetic import synth_split
    from oracle import coracle
    sp = synth_split(shape, 42)
    nu, ni = sp["n_users"], sp["n_items"]
    dsp = temporal_split_device(*sp["all"], nu, device=DEV)
    for part in ("train", "valid", "test"):
        for a, b in zip(dsp[part], sp[part]):
            assert np.array_equal(a.cpu().numpy(), b), (shape, part)
    tu, ti = sp["train"]
    ref = po.build_norm_adj(tu, ti, nu, ni)
    csr = g.NormAdjCSR.from_pairs(tu, ti, nu, ni, device=DEV)
    assert np.array_equal(csr.indptr.cpu().numpy().astype(np.int64), ref["indptr"])
    assert np.array_equal(csr.indices.cpu().numpy(), ref["indices"])
    assert np.array_equal(csr.vals.cpu().numpy().view(np.uint32), ref["vals"].view(np.uint32))
    gen = torch.Generator().manual_seed(0)
    uw, iw = torch.randn(nu, 64, generator=gen) * 0.1, torch.randn(ni, 64, generator=gen) * 0.1
    oue, oie = po.lightgcn_forward(po.to_torch_coo(ref), uw, iw, 3)
    x = g.lightgcn_propagate(csr, torch.cat([uw, iw]).to(DEV), 3)
    assert torch.equal(x[:nu].cpu(), oue) and torch.equal(x[nu:].cpu(), oie), shape
    eu = np.unique(sp["test"][0])[::37][:600]
    ip, it = seen_csr(eu, nu, sp["train"], sp["valid"])
    want = coracle.score_topk(oue.numpy(), oie.numpy(), eu, ip, it, 20)
    exact = g.full_rank_topk(x[:nu], x[nu:], eu, ip, it, 20, tensor_cores=False)
    assert np.array_equal(exact.cpu().numpy(), want), shape
    stats = {}
    tc = g.full_rank_topk(x[:nu], x[nu:], eu, ip, it, 20, tensor_cores=True, stats=stats)
    assert stats["tensor_cores"] and torch.equal(tc, exact), shape


# ----------------------------------------------------------------------------- partitioned graph build
@pytest.mark.parametrize("shape,world,self_loop", [("tiny", 2, False), ("tiny", 3, True), ("C1", 4, False), ("C1", 8, False)])
def test_partitioned_graph_build_equals_rows_of_the_full_build(shape, world, self_loop):
    """gr_build_local_csr_pattern + gr_csr_normalize_local: every rank's rows straight from the pairs ==
    the same rows cut out of the full single-GPU matrix, bit for bit (entries, order inside a row, values),
    incl. duplicate pairs; then the partitioned propagation on those matrices == the single-GPU result."""
    from gnn_recommendations_b200.dist import RowPartition, build_local_csr
    from gnn_recommendations_b200.synthetic import synth_split
    sp = synth_split(shape, 42)
    nu, ni = sp["n_users"], sp["n_items"]
    tu, ti = sp["train"]
    tu, ti = np.concatenate([tu, tu[:50]]), np.concatenate([ti, ti[:50]])      # duplicates are summed
    full = g.NormAdjCSR.from_pairs(tu, ti, nu, ni, self_loop=self_loop, device=DEV)
    part = RowPartition.for_nodes(nu + ni, world)
    assert part.block_rows == RowPartition(full.indptr, world).block_rows
    for rank in range(world):
        want = part.local_csr(full, rank)
        got = build_local_csr(part, rank, tu, ti, nu, ni, self_loop=self_loop, device=DEV)
        assert torch.equal(got.indptr, want.indptr), (shape, world, rank)
        assert torch.equal(got.indices, want.indices), (shape, world, rank)
        assert torch.equal(got.vals.view(torch.int32), want.vals.view(torch.int32)), (shape, world, rank)
        assert (got.n_rows, got.n_cols) == (want.n_rows, want.n_cols)


@pytest.mark.parametrize("shape,world", [("tiny", 2), ("tiny", 3), ("C1", 4), ("C1", 8)])
def test_user_owner_partitioned_build_equals_cuts_of_the_full_build(shape, world):
    """build_user_owner_csrs: each rank's (A_u, A_i) from the pairs of ITS users + the all-reduced item degrees ==
    BipartitePartition.local_csrs(full matrix) bit for bit (pattern, order, values), duplicate pairs included.
    The all-reduce is emulated on one GPU by summing the ranks' item-degree shares."""
    from gnn_recommendations_b200.dist import BipartitePartition, build_user_owner_csrs
    from gnn_recommendations_b200.synthetic import synth_split
    sp = synth_split(shape, 42)
    nu, ni = sp["n_users"], sp["n_items"]
    tu, ti = sp["train"]
    tu, ti = np.concatenate([tu, tu[:50]]), np.concatenate([ti, ti[:50]])      # duplicates are summed
    full = g.NormAdjCSR.from_pairs(tu, ti, nu, ni, device=DEV)
    part = BipartitePartition(nu, ni, world)
    # pass 1: every rank's share of the item degrees (what the all-reduce would add up)
    shares = []
    for rank in range(world):
        build_user_owner_csrs(part, rank, tu, ti, device=DEV, item_degree_allreduce=lambda t: shares.append(t.clone()))
    total = torch.stack(shares).sum(0)
    assert torch.equal(total.to(torch.int32), full.deg[nu:])
    for rank in range(world):
        want_u, want_i = part.local_csrs(full, rank)
        got_u, got_i = build_user_owner_csrs(part, rank, tu, ti, device=DEV, long_threshold=full.long_threshold,
                                             item_degree_allreduce=lambda t: t.copy_(total))
        for got, want, name in ((got_u, want_u, "A_u"), (got_i, want_i, "A_i")):
            assert (got.n_rows, got.n_cols) == (want.n_rows, want.n_cols), (name, rank)
            assert torch.equal(got.indptr, want.indptr), (shape, world, rank, name)
            assert torch.equal(got.indices, want.indices), (shape, world, rank, name)
            assert torch.equal(got.vals.view(torch.int32), want.vals.view(torch.int32)), (shape, world, rank, name)


def test_spmm_two_streams_concurrently_share_no_state():
    """The C-ABI keeps no device-side state (the long-row ticket counters were __device__ globals):
    two propagations with long rows in flight on two streams — each with its own scheduler words and side
    stream — must both give the single-stream bits, repeatedly."""
    rng = np.random.default_rng(5)
    nu, ni, e = 3000, 400, 300000
    u = rng.integers(0, nu, e)
    i = (rng.pareto(1.2, e) * 3).astype(np.int64) % ni                 # hot item rows: thousands of entries
    csr_a = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=DEV, )
    csr_b = g.NormAdjCSR.from_pairs(u[::2], i[::2], nu, ni, device=DEV)
    assert csr_a.n_long > 0 and csr_b.n_long > 0
    xa = torch.randn(nu + ni, 64, generator=torch.Generator().manual_seed(1)).to(DEV)
    xb = torch.randn(nu + ni, 128, generator=torch.Generator().manual_seed(2)).to(DEV)
    want_a, want_b = csr_a.spmm(xa)[0].clone(), csr_b.spmm(xb)[0].clone()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for _ in range(20):
        with torch.cuda.stream(s1):
            ya = csr_a.spmm(xa)[0]
        with torch.cuda.stream(s2):
            yb = csr_b.spmm(xb)[0]
            yb2 = csr_a.spmm(xa)[0]                                     # the SAME adjacency on a second stream
        outs.append((ya, yb, yb2))
    torch.cuda.synchronize()
    assert len(csr_a._sched) >= 3                                       # default + two streams: separate words
    for ya, yb, yb2 in outs:
        assert torch.equal(ya, want_a) and torch.equal(yb, want_b) and torch.equal(yb2, want_a)


# ----------------------------------------------------------------------------- user-owner (1.5-D) propagation
@pytest.mark.parametrize("shape,world,d,L", [("tiny", 2, 64, 3), ("tiny", 3, 32, 2), ("tiny", 8, 128, 4), ("C1", 4, 64, 3),
                                             ("C1", 8, 128, 1)])
def test_user_owner_propagation_matches_single_gpu(shape, world, d, L):
    """The 1.5-D multi-GPU mode (dist.BipartitePartition: users owned by ranks, item rows = partial sums reduced
    and broadcast by gr_reduce_bcast_rows), all ranks emulated on one GPU in lockstep: equal to the single-GPU
    propagation within BASELINE's 1e-5 (item rows are per-rank chains added in rank order, not the single chain),
    user rows of a 1-layer propagation bit-identical, and deterministic (two runs bit-equal)."""
    from gnn_recommendations_b200.dist import BipartitePartition, emulate_user_owner
    from gnn_recommendations_b200.synthetic import synth_split
    sp = synth_split(shape, 42)
    nu, ni = sp["n_users"], sp["n_items"]
    full = g.NormAdjCSR.from_pairs(*sp["train"], nu, ni, device=DEV)
    x0 = (torch.randn(nu + ni, d, generator=torch.Generator().manual_seed(3)) * 0.1).to(DEV)
    want = g.lightgcn_propagate(full, x0, L)
    got = emulate_user_owner(full, nu, ni, x0, L, world)
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 1e-5 * scale
    assert torch.equal(got, emulate_user_owner(full, nu, ni, x0, L, world))
    if L == 1:      # user rows read the exact layer-0 item table: the very same chain as on one GPU
        assert torch.equal(got[:nu], want[:nu])
    # the partition covers every entry exactly once
    part = BipartitePartition(nu, ni, world)
    nnz = 0
    for r in range(world):
        a_u, a_i = part.local_csrs(full, r)
        assert a_u.nnz == a_i.nnz                                          # each owned edge: once per direction
        nnz += a_u.nnz + a_i.nnz
    assert nnz == full.nnz


def test_spmm_routed_peer_stores(tiny_ref):
    """gr_spmm_csr_f32 routed mode: row r is stored to ONE peer buffer, g = r / route_block, at row
    offset + r % route_block (the push variant of the user-owner exchange); here the 'peers' are G buffers on one GPU,
    incl. long rows (long_threshold 16) and the empty tail rows."""
    import ctypes
    full = g.NormAdjCSR.from_pairs(tiny_ref["rows"][tiny_ref["rows"] < 300], tiny_ref["indices"][tiny_ref["rows"] < 300] - 300,
                                   300, 200, device=DEV)
    csr = g.NormAdjCSR(full.indptr, full.indices, full.vals, full.n_rows, full.n_cols, long_threshold=16)
    assert csr.n_long > 0
    x = torch.randn(csr.n_cols, 64, generator=torch.Generator().manual_seed(0)).to(DEV)
    want = csr.spmm(x)[0]
    G, blk, slot = 4, 128, 2                                      # 500 rows -> blocks of 128 rows over 4 'ranks'
    bufs = [torch.full((G * blk, 64), float("nan"), device=DEV) for _ in range(G)]
    ptrs = (ctypes.c_void_p * G)(*[b.data_ptr() for b in bufs])
    csr.spmm(x, want_y=False, peers=(ptrs, G, slot * blk, 64, 0, blk))
    torch.cuda.synchronize()
    for k in range(G):
        rows = want[k * blk: min((k + 1) * blk, csr.n_rows)]
        assert torch.equal(bufs[k][slot * blk: slot * blk + rows.shape[0]], rows)
        assert torch.isnan(bufs[k][: slot * blk]).all()            # nothing else touched


def test_sm_copy_between_pinned_host_and_device():
    """sm_copy (gr_peer_copy_multi on the current stream): device -> pinned host and back, byte-exact, including a
    size that is no multiple of 16 bytes (the tail goes through copy_) and a side stream."""
    for n in (1 << 20, (1 << 20) + 3, 5):
        src = torch.arange(n, dtype=torch.int32, device=DEV)
        host = torch.empty(n, dtype=torch.int32, pin_memory=True)
        host.fill_(-1)
        g.sm_copy(host, src, 8)
        torch.cuda.synchronize()
        assert torch.equal(host, src.cpu()), n
        back = torch.zeros(n, dtype=torch.int32, device=DEV)
        st = torch.cuda.Stream(priority=-1)
        with torch.cuda.stream(st):
            g.sm_copy(back, host, 8)
        st.synchronize()
        assert torch.equal(back, src), n
