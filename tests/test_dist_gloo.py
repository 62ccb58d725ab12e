"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row partition + per-layer all-gather
of LightGCN propagation, and the item-sharded top-K exchange + merge.  The CUDA kernels are replaced
by injected CPU stand-ins (oracle ops), so only the partition / exchange / merge logic is under
test here; the kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_recommendations_b200 import _lib
from gnn_recommendations_b200.dist import (RowPartition, ShardedLightGCN, full_rank_topk_sharded, gather_rows,
                                           item_shard, lightgcn_propagate_sharded)
from gnn_recommendations_b200.synthetic import synth_split
from oracle import pyoracle as po


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class CpuCsr:
    """Stand-in for NormAdjCSR on the CPU: same fields, spmm through torch.sparse.mm."""

    def __init__(self, indptr, indices, vals, n_rows, n_cols):
        self.indptr, self.indices, self.vals = indptr, indices, vals
        self.n_rows, self.n_cols, self.long_threshold = n_rows, n_cols, 1024
        rows = torch.repeat_interleave(torch.arange(n_rows), (indptr[1:] - indptr[:-1]).long())
        self.coo = torch.sparse_coo_tensor(torch.stack([rows, indices.long()]), vals, (n_rows, n_cols))


def _cpu_local_csr(part, adj, rank):
    ip = adj["indptr"]
    rows = np.arange(rank, part.n, part.world_size)
    counts = ip[rows + 1] - ip[rows]
    src = np.concatenate([np.arange(ip[r], ip[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))
    indices = part.to_padded(torch.from_numpy(adj["indices"][src].astype(np.int64)))
    return CpuCsr(indptr, indices, torch.from_numpy(adj["vals"][src]), len(rows), part.padded_rows)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = synth_split("tiny", 42)
        nu, ni = sp["n_users"], sp["n_items"]
        adj = po.build_norm_adj(*sp["train"], nu, ni)
        g = torch.Generator().manual_seed(0)
        x0 = torch.randn(nu + ni, 64, generator=g) * 0.1
        part = RowPartition(torch.from_numpy(adj["indptr"]), world)
        local = _cpu_local_csr(part, adj, rank)

        def spmm(x, y, addend, o, scale, mode):
            t = torch.sparse.mm(local.coo, x)
            if y is not None:
                y.copy_(t)
            if o is not None:
                r = t if addend is None else addend + t
                o.copy_(r / scale if mode == _lib.GR_SCALE_DIV else r)

        mine = lightgcn_propagate_sharded(local, part, rank, part.take_rows(x0, rank), 3, spmm=spmm)
        full = gather_rows(part, rank, mine)
        ue, ie = po.lightgcn_forward(po.to_torch_coo(adj), x0[:nu], x0[nu:], 3)
        ok_prop = torch.equal(full, torch.cat([ue, ie]))          # bit-identical to the single-process result

        # item-sharded top-K on the gathered embeddings
        eval_users = np.arange(0, nu, 3)
        seen = {}
        for u, i in zip(sp["train"][0].tolist(), sp["train"][1].tolist()):
            seen.setdefault(u, set()).add(i)
        want = po.score_mask_topk(ue, ie, eval_users.tolist(), seen, 20)
        lo, hi = item_shard(ni, world, rank)

        def partial_fn(u_e, i_e, a, b, eu, sip, sit, k, ns):
            sc = u_e[eu] @ i_e.T
            for r, u in enumerate(eu.tolist()):
                s = [it - a for it in seen.get(u, ()) if a <= it < b]
                if s:
                    sc[r, s] = float("-inf")
            ids = torch.from_numpy(po.canonical_topk(sc, min(k, b - a)))
            ps = torch.full((1, len(eu), k), float("-inf"))
            pi = torch.full((1, len(eu), k), -1, dtype=torch.int32)
            ps[0, :, :ids.shape[1]] = torch.gather(sc, 1, ids)
            pi[0, :, :ids.shape[1]] = (ids + a).int()
            return ps, pi

        def merge_fn(ps, pi, k):
            p, n, _ = ps.shape
            s = ps.permute(1, 0, 2).reshape(n, -1).numpy()
            i = pi.permute(1, 0, 2).reshape(n, -1).numpy()
            out = np.empty((n, k), dtype=np.int64)
            for r in range(n):
                valid = i[r] >= 0
                order = np.lexsort((i[r][valid], -s[r][valid].astype(np.float64)))
                out[r] = i[r][valid][order][:k]
            return torch.from_numpy(out)

        got = full_rank_topk_sharded(full[:nu], full[nu + lo: nu + hi], lo, hi, eval_users, None, None, 20, world,
                                     partial_fn=partial_fn, merge_fn=merge_fn)
        ok_topk = np.array_equal(got.numpy(), want)
        out[rank] = (ok_prop, ok_topk)
    finally:
        dist.destroy_process_group()


def _train_worker(rank, world, port, out):
    """Two row-partitioned training steps (ShardedLightGCN) vs the single-process step of the reference
    (trainer.py:237-279 restated with the oracle's ops: propagate, B x B BPR, clip, Adam)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = synth_split("tiny", 42)
        nu, ni = sp["n_users"], sp["n_items"]
        adj = po.build_norm_adj(*sp["train"], nu, ni)
        gen = torch.Generator().manual_seed(0)
        x0 = torch.randn(nu + ni, 32, generator=gen) * 0.1
        part = RowPartition(torch.from_numpy(adj["indptr"]), world)
        local = _cpu_local_csr(part, adj, rank)

        def spmm(x, y, addend, o, scale, mode):
            t = torch.sparse.mm(local.coo, x)
            if y is not None:
                y.copy_(t)
            if o is not None:
                r = t if addend is None else addend + t
                o.copy_(r / scale if mode == _lib.GR_SCALE_DIV else r)

        def propagate(x_local):
            return lightgcn_propagate_sharded(local, part, rank, x_local, 3, spmm=spmm)

        def bpr(table, n_users_b, users, pos, neg):
            return po.bpr_loss_reference(table[:n_users_b], table[n_users_b:], users, pos, neg)

        model = ShardedLightGCN(local, part, rank, part.take_rows(x0, rank), nu, 3, lr=1e-2, weight_decay=1e-4,
                                max_grad_norm=1.0, propagate=propagate, bpr=bpr, fused_optimizer=False)
        # single-process reference
        uw = torch.nn.Parameter(x0[:nu].clone())
        iw = torch.nn.Parameter(x0[nu:].clone())
        opt = torch.optim.Adam([uw, iw], lr=1e-2, weight_decay=1e-4)
        coo = po.to_torch_coo(adj)
        rng = np.random.default_rng(7)
        ok = True
        for step in range(2):
            users = torch.from_numpy(rng.integers(0, nu, 48))
            users[1] = users[0]                                   # duplicates accumulate
            pos = torch.from_numpy(rng.integers(0, ni, 48))
            neg = torch.from_numpy(rng.integers(0, ni, 48)).view(-1, 1)
            loss = model.train_step(users, pos, neg)
            opt.zero_grad()
            ue, ie = po.lightgcn_forward(coo, uw, iw, 3)
            ref_loss = po.bpr_loss_reference(ue, ie, users, pos, neg)
            ref_loss.backward()
            torch.nn.utils.clip_grad_norm_([uw, iw], 1.0)
            opt.step()
            ok = ok and abs(loss - float(ref_loss)) <= 1e-6 * abs(float(ref_loss))
            got = model.gathered_weight()
            want = torch.cat([uw.detach(), iw.detach()])
            ok = ok and bool(torch.allclose(got, want, rtol=1e-5, atol=2e-6))   # Adam amplifies 1e-10 gradient differences
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_sharded_training_step_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_train_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


@pytest.mark.timeout(300)
def test_world2_propagation_and_sharded_topk_match_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True), 1: (True, True)}


def test_item_shard_covers_range():
    for n, w in ((200, 2), (3706, 8), (91599, 8), (5, 4)):
        spans = [item_shard(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_sharded_training_refuses_row_normalised_graphs():
    """ShardedLightGCN's backward applies the forward row blocks to the gradient, which is the
    gradient only for a symmetric adjacency; a 'row'-normalised (D^-1 A) partition must be refused, not
    silently mis-trained."""
    sp = synth_split("tiny", 42)
    nu, ni = sp["n_users"], sp["n_items"]
    adj = po.build_norm_adj(*sp["train"], nu, ni, normalization="row")
    part = RowPartition(torch.from_numpy(adj["indptr"]), 2)
    local = _cpu_local_csr(part, adj, 0)
    local.full_symmetric = False            # what RowPartition.local_csr / build_local_csr record for 'row'
    x0 = torch.zeros(part.n_local(0), 8)
    with pytest.raises(ValueError, match="symmetric"):
        ShardedLightGCN(local, part, 0, x0, nu, 2, propagate=lambda x: x, bpr=lambda *a: None, fused_optimizer=False)
    local.full_symmetric = True
    ShardedLightGCN(local, part, 0, x0, nu, 2, propagate=lambda x: x, bpr=lambda *a: None, fused_optimizer=False)
