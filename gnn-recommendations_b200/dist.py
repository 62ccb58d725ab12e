"""Row-partitioned propagation across the GPUs of one box (no reference
counterpart — the reference is single-process).

Rows of Â are distributed cyclically over the G ranks (rank g owns rows g, g+G, ...; see
RowPartition).  Rank g owns the embedding rows of its share; every layer is one exchange of the
layer's rows (fused into the SpMM epilogue as NVLink P2P stores, or an NCCL all-gather) followed by
a local SpMM.  Shares are padded to a common height H so the gathered matrix is a dense [G*H, d]
buffer with equal chunks; column ids are remapped once to that owner-major numbering.  Per (row, feature) the arithmetic is the same fmaf
chain as on one GPU, so the G-GPU result is bit-identical to the 1-GPU result.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import numpy as np
import torch

from . import _lib
from .graph_builder import NormAdjCSR


class RowPartition:
    """Cyclic row distribution: rank g owns global rows g, g+G, g+2G, ...

    Both the work (entries) and the exchange volume (rows) of a rank must be 1/G of the total: the
    all-gather of a layer moves every owned ROW to every peer, the SpMM cost follows the ENTRIES.
    Node ids of the synthetic / real graphs are popularity-ordered, so contiguous blocks cannot
    balance both (measured on 4 GPUs: the nnz-balanced block of low-activity users held 58 % of all
    rows and its NVLink egress set the step time); a cyclic distribution balances rows exactly and
    entries statistically, and also spreads the hottest item rows over all ranks."""

    def __init__(self, indptr: torch.Tensor, world_size: int):
        """``indptr``: the full matrix's row pointer (only its length is used)."""
        self.n = int(indptr.numel()) - 1
        self.world_size = int(world_size)
        rows = -(-self.n // self.world_size)
        self.block_rows = max(4, (rows + 3) // 4 * 4)        # common padded block height H
        self.padded_rows = self.block_rows * self.world_size

    @classmethod
    def for_nodes(cls, n_nodes: int, world_size: int) -> "RowPartition":
        """Partition of an [n_nodes, n_nodes] matrix that has not been built (see build_local_csr)."""
        self = cls.__new__(cls)
        self.n = int(n_nodes)
        self.world_size = int(world_size)
        rows = -(-self.n // self.world_size)
        self.block_rows = max(4, (rows + 3) // 4 * 4)
        self.padded_rows = self.block_rows * self.world_size
        return self

    def n_local(self, rank: int) -> int:
        return len(range(rank, self.n, self.world_size))

    def local_ids(self, rank: int, device=None) -> torch.Tensor:
        return torch.arange(rank, self.n, self.world_size, device=device)

    def take_rows(self, x_full: torch.Tensor, rank: int) -> torch.Tensor:
        return x_full[rank::self.world_size].contiguous()

    def to_padded(self, ids: torch.Tensor) -> torch.Tensor:
        """global node id -> position in the padded [G*H] numbering (owner-major)."""
        ids = ids.long()
        return (ids % self.world_size) * self.block_rows + ids // self.world_size

    def local_csr(self, full: NormAdjCSR, rank: int) -> NormAdjCSR:
        rows = self.local_ids(rank, full.indptr.device)
        starts = full.indptr[rows].long()
        counts = (full.indptr[rows + 1] - full.indptr[rows]).long()
        indptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rows.device)
        torch.cumsum(counts, 0, out=indptr[1:])
        total = int(indptr[-1].item())
        # entry k of local row j  ->  full entry starts[j] + (k - indptr[j])
        src = torch.repeat_interleave(starts - indptr[:-1], counts) + torch.arange(total, device=rows.device)
        indices = self.to_padded(full.indices[src]).to(torch.int32).contiguous()
        vals = full.vals[src].contiguous()
        local = NormAdjCSR(indptr.to(torch.int32), indices, vals, int(rows.numel()), self.padded_rows,
                           long_threshold=full.long_threshold)
        # whether the FULL matrix equals its transpose (True for the reference's 'symmetric' normalisation);
        # None = unknown.  ShardedLightGCN's backward needs it.
        local.full_symmetric = full._symmetric
        return local


def build_local_csr(part: RowPartition, rank: int, user, item, n_users: int, n_items: int,
                    normalization: str = "symmetric", self_loop: bool = False, device="cuda",
                    dis_lut: Optional[np.ndarray] = None, long_threshold: Optional[int] = None) -> NormAdjCSR:
    """Partitioned graph build: this rank's rows of Â straight from the (user,item) pairs,
    without materialising the full matrix.  Equal, bit for bit, to ``part.local_csr(NormAdjCSR.from_pairs(..), rank)``
    (same entries, same order inside a row, same values from the global degrees)."""
    from ._lib import check, lib, ptr, stream_ptr
    from .graph_builder import _NORM_MODES, _require_cuda, degree_lut

    if normalization not in _NORM_MODES:
        raise ValueError(f"Неизвестный тип нормализации: {normalization}")
    device = _require_cuda(device)
    user = torch.as_tensor(user, dtype=torch.int64).to(device).contiguous()
    item = torch.as_tensor(item, dtype=torch.int64).to(device).contiguous()
    if user.numel() != item.numel():
        raise ValueError("user and item must have the same length")
    n_pairs, n, G = int(user.numel()), n_users + n_items, part.world_size
    if n != part.n:
        raise ValueError("partition was made for a different number of nodes")
    n_local = part.n_local(rank)
    n_entries = int(((user % G) == rank).sum() + (((n_users + item) % G) == rank).sum()) + (n_local if self_loop else 0)
    l = lib()
    ws_bytes = l.gr_build_local_csr_workspace_bytes(n_entries)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    indptr = torch.empty(n_local + 1, dtype=torch.int32, device=device)
    indices = torch.empty(max(n_entries, 1), dtype=torch.int32, device=device)
    mult = torch.empty(max(n_entries, 1), dtype=torch.float32, device=device)
    deg = torch.empty(n, dtype=torch.int32, device=device)
    nnz_d = torch.zeros(1, dtype=torch.int64, device=device)
    maxdeg_d = torch.zeros(1, dtype=torch.int32, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(l.gr_build_local_csr_pattern(ptr(user), ptr(item), n_pairs, n_users, n_items, int(self_loop), G, rank,
                                           part.block_rows, n_entries, ptr(indptr), ptr(indices), ptr(mult), ptr(deg),
                                           ptr(nnz_d), ptr(maxdeg_d), ptr(status), ptr(ws), ws_bytes, stream_ptr()),
              "gr_build_local_csr_pattern")
        nnz, max_deg, st = int(nnz_d.item()), int(maxdeg_d.item()), int(status.item())
        if st & 2:
            raise ValueError("user/item id out of range")
        if st & 8:
            raise _lib.GrError("local entry count mismatch")
        del ws
        mode = _NORM_MODES[normalization]
        if dis_lut is None:
            dis_lut = degree_lut(max_deg, -0.5 if mode == 0 else -1.0)
        if len(dis_lut) <= max_deg and mode != 2:
            raise ValueError(f"dis_lut has {len(dis_lut)} entries, max degree is {max_deg}")
        lut = torch.from_numpy(np.ascontiguousarray(dis_lut, dtype=np.float32)).to(device)
        indices = indices[:nnz].clone()
        vals = torch.empty(nnz, dtype=torch.float32, device=device)
        check(l.gr_csr_normalize_local(ptr(indptr), ptr(indices), ptr(mult), ptr(deg), ptr(lut), int(lut.numel()),
                                       n_local, nnz, mode, G, rank, part.block_rows, ptr(vals), ptr(status),
                                       stream_ptr()), "gr_csr_normalize_local")
        if int(status.item()) & 4:
            raise _lib.GrError("degree exceeds the look-up table")
    kw = {} if long_threshold is None else {"long_threshold": long_threshold}
    local = NormAdjCSR(indptr, indices, vals, n_local, part.padded_rows, **kw)
    local.full_symmetric = (mode != 1)      # 'row' normalisation D^-1 A is not symmetric
    return local


class PeerExchange:
    """Symmetric (peer-mapped) gathered-layer buffers for the fused SpMM + all-gather.

    Every rank allocates the same two [G*H, d] buffers with torch's symmetric-memory allocator and
    exchanges their device pointers (plumbing); the SpMM epilogue then stores each finished row
    straight into all G buffers over NVLink (gr_spmm_csr_f32 `peer_y`), so a layer's exchange is
    hidden behind its own computation and the only collective left is the barrier between layers."""

    def __init__(self, part: RowPartition, d: int, device, group=None):
        import ctypes

        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.part, self.d = part, d
        self.bufs, self.handles, self.ptr_arrays = [], [], []
        want_mc = os.environ.get("GR_MULTICAST", "0") == "1" and self.world > 1
        self.multicast = want_mc
        for _ in range(2):
            t = symm_mem.empty((part.padded_rows, d), dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, self.group)
            t.zero_()
            self.bufs.append(t)
            self.handles.append(h)
            if not int(getattr(h, "multicast_ptr", 0) or 0):
                self.multicast = False
        flag = torch.tensor([1 if self.multicast else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)       # all ranks must agree
        self.multicast = bool(flag.item())
        for h in self.handles:
            if self.multicast:      # NVSwitch multicast (NVLS): one store reaches every rank
                arr = (ctypes.c_void_p * 1)(int(h.multicast_ptr))
            else:                   # one mapped pointer per rank
                arr = (ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])
            self.ptr_arrays.append(arr)
        self.n_targets = 1 if self.multicast else self.world
        self.row_off = self.rank * part.block_rows
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def peers(self, which: int):
        return (self.ptr_arrays[which], self.n_targets, self.row_off, self.d, 1 if self.multicast else 0)

    def barrier(self, which: int):
        """All ranks' stores into buffer `which` have landed (stream-ordered, device-side)."""
        self.handles[which].barrier()

    def scatter(self, which: int, x_local: torch.Tensor):
        from ._lib import check, lib, ptr, stream_ptr

        with torch.cuda.device(x_local.device):
            check(lib().gr_peer_scatter_rows(ptr(x_local), x_local.stride(0), x_local.shape[0], self.d,
                                             self.ptr_arrays[which], self.n_targets, 1 if self.multicast else 0,
                                             self.d, self.row_off, stream_ptr()), "gr_peer_scatter_rows")


def lightgcn_propagate_fused(local: NormAdjCSR, ex: PeerExchange, x0_local: torch.Tensor,
                             n_layers: int) -> torch.Tensor:
    """Row-partitioned LightGCN propagation with the exchange fused into the SpMM epilogue."""
    if n_layers == 0:
        return x0_local.clone()
    acc = torch.empty_like(x0_local)
    out = torch.empty_like(x0_local)
    cur = 0
    ex.barrier(cur)                       # nobody is still reading buffer `cur` from the previous call
    ex.scatter(cur, x0_local)             # layer-0 rows -> every rank's buffer
    ex.barrier(cur)
    for l in range(n_layers):
        last = l == n_layers - 1
        addend = x0_local if l == 0 else acc
        if last:
            local.spmm(ex.bufs[cur], addend=addend, out=out, scale=float(n_layers + 1),
                       scale_mode=_lib.GR_SCALE_DIV, want_y=False)
        else:
            nxt = cur ^ 1
            local.spmm(ex.bufs[cur], addend=addend, out=acc, want_y=False, peers=ex.peers(nxt))
            ex.barrier(nxt)
            cur = nxt
    return out


def _all_gather_rows(buf: torch.Tensor, rank: int, block_rows: int, group=None) -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        mine = buf[rank * block_rows:(rank + 1) * block_rows]
        dist.all_gather_into_tensor(buf, mine, group=group)


def lightgcn_propagate_sharded(local: NormAdjCSR, part: RowPartition, rank: int, x0_local: torch.Tensor,
                               n_layers: int, group=None,
                               spmm: Optional[Callable] = None) -> torch.Tensor:
    """mean_{l<=L} Â^l x0 for this rank's rows.  ``x0_local``: [n_local(rank), d].  ``spmm`` is
    injectable for the CPU (gloo) tests of the partition / exchange logic; the product path is
    the CUDA kernel."""
    n_local, d = x0_local.shape
    H = part.block_rows
    dev = x0_local.device
    off = rank * H
    if spmm is None:
        def spmm(x, y, addend, out, scale, mode):
            local.spmm(x, y=y, addend=addend, out=out, scale=scale, scale_mode=mode, want_y=y is not None)
    bufs = [torch.zeros((part.padded_rows, d), dtype=torch.float32, device=dev) for _ in range(min(2, max(1, n_layers)))]
    bufs[0][off:off + n_local].copy_(x0_local)
    if n_layers == 0:
        return x0_local.clone()
    acc = torch.empty_like(x0_local)
    out = torch.empty_like(x0_local)
    cur = 0
    for l in range(n_layers):
        _all_gather_rows(bufs[cur], rank, H, group)
        last = l == n_layers - 1
        addend = x0_local if l == 0 else acc
        if last:
            spmm(bufs[cur], None, addend, out, float(n_layers + 1), _lib.GR_SCALE_DIV)
        else:
            nxt = (cur + 1) % len(bufs)
            spmm(bufs[cur], bufs[nxt][off:off + n_local], addend, acc, 1.0, _lib.GR_SCALE_NONE)
            cur = nxt
    return out


# ------------------------------------------------------------------------------------------------
# item-sharded full-ranking evaluation
# ------------------------------------------------------------------------------------------------
def item_shard(n_items: int, world_size: int, rank: int):
    """Contiguous id range [lo, hi) of the rank's items (multiple-of-128 boundaries)."""
    per = -(-n_items // world_size)
    per = -(-per // 128) * 128
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def gather_rows(part: RowPartition, rank: int, x_local: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks' rows -> the full [N, d] matrix in natural row order (replicated)."""
    d = x_local.shape[1]
    buf = torch.zeros((part.padded_rows, d), dtype=x_local.dtype, device=x_local.device)
    buf[rank * part.block_rows: rank * part.block_rows + x_local.shape[0]].copy_(x_local)
    _all_gather_rows(buf, rank, part.block_rows, group)
    out = torch.empty((part.n, d), dtype=x_local.dtype, device=x_local.device)
    for g in range(part.world_size):
        out[g::part.world_size] = buf[g * part.block_rows: g * part.block_rows + part.n_local(g)]
    return out


def merge_topk_partials(ps: torch.Tensor, pi: torch.Tensor, k: int) -> torch.Tensor:
    """K-way merge of per-shard top-k lists [segments, n_eval, k] under (score desc, id asc) (gr_topk_merge)."""
    from ._lib import check, lib, ptr, stream_ptr

    n_eval, dev = int(ps.shape[1]), ps.device
    ps, pi = ps.contiguous(), pi.contiguous()
    ids = torch.empty((n_eval, k), dtype=torch.int64, device=dev)
    sc = torch.empty((n_eval, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().gr_topk_merge(ptr(ps), ptr(pi), int(ps.shape[0]), n_eval, k, ptr(ids), ptr(sc), stream_ptr()),
              "gr_topk_merge")
    return ids


def full_rank_topk_sharded(user_emb: torch.Tensor, item_emb_local: torch.Tensor, item_lo: int, item_hi: int,
                           eval_users, seen_indptr, seen_items, k: int, world_size: int, group=None,
                           n_splits: int = 1, partial_fn: Optional[Callable] = None,
                           merge_fn: Optional[Callable] = None, tensor_cores: Optional[bool] = None,
                           return_partial: bool = False):
    """Each rank ranks its own item id range [item_lo, item_hi) for ALL eval users
    (gr_score_topk_partial), the per-rank top-k (score, id) lists are all-gathered
    (k * 8 bytes per user per rank) and merged under (score desc, id asc) (gr_topk_merge).
    Scores are per-(user,item) independent, so the result is bit-identical to the 1-GPU list.
    ``tensor_cores`` (None = automatic): the rank's partial list comes from the tcgen05 nomination + exact
    re-scoring path of ``full_rank_topk`` run on the rank's item shard (seen-item CSR cut to the shard's id
    range and shifted), which yields the same bits as the exact partial kernel.
    ``partial_fn`` / ``merge_fn`` are injectable for the CPU (gloo) tests of the exchange logic."""
    import torch.distributed as dist

    dev = user_emb.device
    eval_users = torch.as_tensor(eval_users, dtype=torch.int64).to(dev).contiguous()
    n_eval = int(eval_users.numel())
    if partial_fn is None:
        from ._lib import check, lib, ptr, stream_ptr
        from .evaluator import TC_MIN_ITEMS, full_rank_topk, tc_kprime

        def exact_partial(ue, ie, lo, hi, eu, sip, sit, kk, ns):
            ps = torch.empty((ns, n_eval, kk), dtype=torch.float32, device=dev)
            pi = torch.empty((ns, n_eval, kk), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                check(lib().gr_score_topk_partial(ptr(ue), ue.stride(0), ptr(ie), ie.stride(0), int(ue.shape[1]),
                                                  ptr(eu), n_eval, lo, hi, ptr(sip), ptr(sit), kk, ns, ptr(ps),
                                                  ptr(pi), stream_ptr()), "gr_score_topk_partial")
            return ps, pi

        def tc_partial(ue, ie, lo, hi, eu, sip, sit, kk, ns):
            # seen items of each row restricted to [lo, hi) and shifted to shard-local ids (rows stay sorted)
            lip = lit = None
            if sip is not None:
                inside = (sit >= lo) & (sit < hi)
                csum = torch.zeros(sit.numel() + 1, dtype=torch.int64, device=dev)
                torch.cumsum(inside, 0, out=csum[1:])
                lip = csum[sip.clamp(max=sit.numel())]
                lit = (sit[inside] - lo).to(torch.int32)
            ids, sc = full_rank_topk(ue, ie, eu, lip, lit, kk, return_scores=True, tensor_cores=True)
            ids = torch.where(ids >= 0, ids + lo, ids)
            return sc.unsqueeze(0).contiguous(), ids.to(torch.int32).unsqueeze(0).contiguous()

        def partial_fn(ue, ie, lo, hi, eu, sip, sit, kk, ns):
            d_, n_loc = int(ue.shape[1]), int(hi - lo)
            kp = tc_kprime(kk)
            can = bool(lib().gr_topk_tc_supported(d_, kp)) and kk + 8 <= kp and n_loc >= max(kp, TC_MIN_ITEMS)
            if tensor_cores is True and not can:
                raise ValueError(f"tensor-core top-K path does not support d={d_}, k={kk}, shard of {n_loc} items")
            if can and tensor_cores is not False and ns == 1:
                return tc_partial(ue, ie, lo, hi, eu, sip, sit, kk, ns)
            return exact_partial(ue, ie, lo, hi, eu, sip, sit, kk, ns)

    if merge_fn is None:
        merge_fn = merge_topk_partials

    if seen_indptr is not None:
        seen_indptr = torch.as_tensor(seen_indptr, dtype=torch.int64).to(dev).contiguous()
        seen_items = torch.as_tensor(seen_items, dtype=torch.int32).to(dev).contiguous()
        if seen_items.numel() == 0:
            seen_items = torch.zeros(1, dtype=torch.int32, device=dev)
    ps, pi = partial_fn(user_emb.contiguous(), item_emb_local.contiguous(), int(item_lo), int(item_hi), eval_users,
                        seen_indptr, seen_items, k, n_splits)
    if return_partial:          # (scores, ids) [segments, n_eval, k] of this rank, before the exchange
        return ps, pi
    if world_size > 1 and dist.is_initialized():
        all_ps = torch.empty((world_size * ps.shape[0], n_eval, k), dtype=ps.dtype, device=dev)
        all_pi = torch.empty((world_size * pi.shape[0], n_eval, k), dtype=pi.dtype, device=dev)
        dist.all_gather_into_tensor(all_ps, ps.contiguous(), group=group)
        dist.all_gather_into_tensor(all_pi, pi.contiguous(), group=group)
        ps, pi = all_ps, all_pi
    return merge_fn(ps, pi, k)


# ------------------------------------------------------------------------------------------------
# row-partitioned LightGCN training step (BPR step + backward)
# ------------------------------------------------------------------------------------------------
class ShardedLightGCN:
    """LightGCN whose embedding table (and its Adam state) is row-partitioned like the graph.

    One step = trainer.py:237-279 on G ranks:
      forward   row-partitioned propagation (one exchange per layer);
      BPR       the batch is replicated: the 3B propagated rows it touches are collected with one
                all-reduce of a [3B, d] buffer (each row has exactly one owner, the others add zeros),
                every rank evaluates the same fused B x B loss on that compact table and keeps the
                gradient rows it owns;
      backward  d loss / d E0 = mean_l Â^l g (Â is symmetric, so the backward of the layer mean IS the
                forward operator applied to the gradient): the same row-partitioned propagation;
      step      global gradient norm by one scalar all-reduce, then the fused clip + Adam pass on the
                local rows only — no gradient all-reduce, no replicated parameters.
    ``propagate`` and ``bpr`` are injectable for the CPU (gloo) tests of the host logic.
    """

    def __init__(self, local: NormAdjCSR, part: RowPartition, rank: int, x0_local: torch.Tensor, n_users: int,
                 n_layers: int, lr: float = 1e-3, weight_decay: float = 1e-4, max_grad_norm: float = 1.0,
                 exchange: Optional["PeerExchange"] = None, group=None, propagate: Optional[Callable] = None,
                 bpr: Optional[Callable] = None, fused_optimizer: bool = True):
        # The backward pass applies the FORWARD row blocks to the gradient (mean_l Â^l g), which is dL/dE0 only
        # when Â = Â^T.  A row-normalised graph (normalization='row': D^-1 A) would silently get wrong
        # gradients, so it is refused (the single-GPU path handles it through csr.transpose()).
        if getattr(local, "full_symmetric", None) is False:
            raise ValueError("ShardedLightGCN needs a symmetric adjacency (normalization='symmetric' or 'none'): "
                             "its backward pass reuses the forward row blocks; train 'row'-normalised graphs on "
                             "one GPU")
        self.local, self.part, self.rank, self.group = local, part, rank, group
        self.n_users, self.n_layers = int(n_users), int(n_layers)
        self.max_grad_norm = float(max_grad_norm)
        self.weight = torch.nn.Parameter(x0_local.detach().clone())        # rows rank, rank+G, ... of [E_user; E_item]
        self.optimizer = torch.optim.Adam([self.weight], lr=lr, weight_decay=weight_decay)
        self.exchange = exchange
        self._propagate = propagate
        self._bpr = bpr
        self.fused_optimizer = fused_optimizer

    # ---- pieces ---------------------------------------------------------------------------------
    def propagate(self, x_local: torch.Tensor) -> torch.Tensor:
        if self._propagate is not None:
            return self._propagate(x_local)
        if self.exchange is not None:
            return lightgcn_propagate_fused(self.local, self.exchange, x_local, self.n_layers)
        return lightgcn_propagate_sharded(self.local, self.part, self.rank, x_local, self.n_layers, self.group)

    def _all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def batch_rows(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor):
        """global row ids of the 3B embedding rows of a batch, their owner mask and local indices."""
        need = torch.cat([users.view(-1), self.n_users + pos.view(-1), self.n_users + neg.view(-1)]).long()
        mine = (need % self.part.world_size) == self.rank
        return need, mine, need // self.part.world_size

    def train_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> float:
        B = int(users.numel())
        dev = self.weight.device
        with torch.no_grad():
            out_local = self.propagate(self.weight.detach())
            need, mine, loc = self.batch_rows(users.to(dev), pos.to(dev), neg.to(dev))
            table = torch.zeros((3 * B, out_local.shape[1]), dtype=out_local.dtype, device=dev)
            table[mine] = out_local[loc[mine]]
            self._all_reduce(table)
        # the same B x B loss every rank: users 0..B-1, positives B..2B-1, negatives 2B..3B-1 of the table
        ar = torch.arange(B, device=dev)
        table.requires_grad_(True)
        if self._bpr is not None:
            loss = self._bpr(table, B, ar, ar, (B + ar).view(-1, 1))
        else:
            from .losses import bpr_fused

            loss = bpr_fused(table, B, ar, ar, (B + ar).view(-1, 1))
        loss.backward()
        with torch.no_grad():
            g_local = torch.zeros_like(self.weight)
            g_local.index_add_(0, loc[mine], table.grad[mine])              # duplicates accumulate
            grad = self.propagate(g_local)                                   # mean_l Â^l g
            self.weight.grad = grad
            if self.max_grad_norm > 0:
                sq = self._all_reduce(grad.double().pow(2).sum().view(1))
                coef = self.max_grad_norm / (float(sq.sqrt()) + 1e-6)      # clip_grad_norm_: clamped to 1
                if coef < 1.0:
                    grad.mul_(coef)
            if self.fused_optimizer and grad.is_cuda:
                from .optim import fused_clip_adam_step

                fused_clip_adam_step(self.optimizer, 0.0)                    # already clipped with the GLOBAL norm
            else:
                self.optimizer.step()
        return float(loss.detach())

    def gathered_weight(self) -> torch.Tensor:
        return gather_rows(self.part, self.rank, self.weight.detach(), self.group)


# ------------------------------------------------------------------------------------------------
# user-owner ("1.5-D") propagation: item rows as reduced partial sums  (SURVEY.md §8e bipartite refinement)
# ------------------------------------------------------------------------------------------------
class BipartitePartition:
    """Users are OWNED by ranks (cyclically: user u lives on rank u % G as local row u // G; its embedding row
    never leaves the rank), items are replicated: every rank holds the full item table of the current layer.

    Why: in the exact row partition every rank must receive every row of every layer — (G-1)/G x N x 4d bytes
    of NVLink egress per rank and layer, 11.2 GB at C5 on 8 GPUs against 9.8 ms of SpMM.  But a user row only
    reads ITEM columns and an item row only USER columns.  So a rank can (1) advance its own users from the
    replicated item table with no communication, and (2) compute, for ALL items, the partial sum over the users
    it owns.  Every partial row is pushed from the SpMM epilogue to the rank that owns the row's item block (rank r
    owns items [r*Ib, (r+1)*Ib); routed P2P stores, one staging slot per sender), the owner adds its G slots in rank
    order and broadcasts the block (P2P stores again): 2 x (G-1)/G x I x 4d bytes of egress per rank and layer
    (4.5 GB at C5), all as stores (measured: peer LOADS reached 370 GB/s, stores 650), the first half overlapped
    with the partial-sum SpMM itself, the second with the user-row SpMM.

    An item row is then  sum_g (chain over rank g's users, ascending)  added in rank order — deterministic and
    reproducible, but not the single ascending chain of the 1-GPU kernel: ~1e-7 relative (BASELINE north_star
    asks 1e-5 for embeddings).  The exact all-gather partition (RowPartition) remains the bit-exact mode."""

    def __init__(self, n_users: int, n_items: int, world_size: int):
        self.n_users, self.n_items, self.world_size = int(n_users), int(n_items), int(world_size)
        G = self.world_size
        self.user_rows = max(4, (-(-self.n_users // G) + 3) // 4 * 4)        # padded local user height
        self.item_block = max(4, (-(-self.n_items // G) + 3) // 4 * 4)       # Ib
        self.items_padded = self.item_block * G

    def n_users_local(self, rank: int) -> int:
        return len(range(rank, self.n_users, self.world_size))

    def item_range(self, rank: int):
        lo = min(self.n_items, rank * self.item_block)
        return lo, min(self.n_items, lo + self.item_block)

    def item_shift(self, rank: int) -> int:
        """First item of the row order of rank's partial-sum matrix A_i (the next rank's block)."""
        return ((rank + 1) % self.world_size) * self.item_block

    def take_users(self, x_users: torch.Tensor, rank: int) -> torch.Tensor:
        return x_users[rank::self.world_size].contiguous()

    def local_csrs(self, full: NormAdjCSR, rank: int):
        """(A_u, A_i) cut out of the full [N, N] matrix (users first):
        A_u [n_users_local, items_padded]: this rank's user rows, columns = item index;
        A_i [items_padded, n_users_local]: ALL item rows (rotated, see below) restricted to this rank's users, columns
        = local user row.
        Entry order inside a row stays ascending (the chain order within a rank)."""
        G, U, dev = self.world_size, self.n_users, full.indptr.device
        ip = full.indptr.long()
        # ---- A_u: rows rank, rank+G, ... of the user block
        rows = torch.arange(rank, U, G, device=dev)
        starts, counts = ip[rows], ip[rows + 1] - ip[rows]
        uptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=uptr[1:])
        total = int(uptr[-1].item())
        src = torch.repeat_interleave(starts - uptr[:-1], counts) + torch.arange(total, device=dev)
        a_u = NormAdjCSR(uptr.to(torch.int32), (full.indices[src].long() - U).to(torch.int32).contiguous(),
                         full.vals[src].contiguous(), int(rows.numel()), self.items_padded,
                         long_threshold=full.long_threshold)
        # ---- A_i: item rows, entries whose column (a user) belongs to this rank
        lo = int(ip[U].item())
        seg_cols = full.indices[lo:].long()
        mine = (seg_cols % G) == rank
        csum = torch.zeros(seg_cols.numel() + 1, dtype=torch.int64, device=dev)
        torch.cumsum(mine, 0, out=csum[1:])
        iptr = torch.zeros(self.items_padded + 1, dtype=torch.int64, device=dev)
        iptr[: self.n_items + 1] = csum[ip[U:] - lo]
        iptr[self.n_items + 1:] = iptr[self.n_items]
        cols_i = (seg_cols[mine] // G).to(torch.int32)
        vals_i = full.vals[lo:][mine]
        # Rows ROTATED so that local row rho is item (rho + shift) mod I_pad, shift = start of the NEXT rank's block:
        # all ranks walk their rows in ascending order, and without the rotation they would all push to rank 0's
        # staging buffer first, then all to rank 1's, ... (measured on 8 GPUs: 10 ms per layer of NVLink ingress
        # hot-spotting); rotated, at any moment the G senders target G different owners.
        shift = self.item_shift(rank)
        cut = int(iptr[shift].item())
        counts = iptr[1:] - iptr[:-1]
        rptr = torch.zeros_like(iptr)
        torch.cumsum(torch.cat([counts[shift:], counts[:shift]]), 0, out=rptr[1:])
        a_i = NormAdjCSR(rptr.to(torch.int32), torch.cat([cols_i[cut:], cols_i[:cut]]).contiguous(),
                         torch.cat([vals_i[cut:], vals_i[:cut]]).contiguous(), self.items_padded,
                         self.n_users_local(rank), long_threshold=full.long_threshold)
        return a_u, a_i


def build_user_owner_csrs(part: "BipartitePartition", rank: int, user, item, device="cuda",
                          long_threshold: Optional[int] = None, item_degree_allreduce=None):
    """Partitioned graph build of the user-owner layout (SURVEY.md §8e "graph build": partition the pairs by owner,
    local sort / segment, one all-reduce of degrees): (A_u, A_i) of ``part.local_csrs(full, rank)`` straight from
    the (user, item) pairs WITHOUT the full matrix — bit for bit the same entries, order and values.

    A rank only touches the pairs of the users it owns (u % G == rank): the library's pair builder
    (radix sort + dedupe: gr_build_csr_pattern) makes the pattern of the local bipartite block
    [users_local + items] x [users_local + items]; a user's degree is its local pair count, an item's degree is the
    SUM over ranks of its local pair counts — ``item_degree_allreduce(t)`` (in place on an int64 [n_items] device
    tensor; ``torch.distributed.all_reduce`` in a job, omitted for one rank) is the only collective.  The values
    are the reference's left-to-right product (dis[r] * mult) * dis[c] (graph_builder.py:126) from the same
    degree look-up table as the single-GPU build."""
    from .graph_builder import _require_cuda, degree_lut

    device = _require_cuda(device)
    G, U, I = part.world_size, part.n_users, part.n_items
    user = torch.as_tensor(user, dtype=torch.int64).to(device)
    item = torch.as_tensor(item, dtype=torch.int64).to(device)
    if user.numel() != item.numel():
        raise ValueError("user and item must have the same length")
    if user.numel() and (int(user.min()) < 0 or int(user.max()) >= U or int(item.min()) < 0 or int(item.max()) >= I):
        raise ValueError("user/item id out of range")
    mine = (user % G) == rank
    nul = part.n_users_local(rank)
    loc = NormAdjCSR.from_pairs((user[mine] // G).contiguous(), item[mine].contiguous(), nul, I,
                                normalization="none", device=device)          # vals = multiplicities
    ip = loc.indptr.long()
    counts = ip[1:] - ip[:-1]                          # entries per row (distinct neighbours) of the local block
    # degrees count every pair, duplicates included (the row sums of the reference's summed-duplicate matrix)
    deg_u = loc.deg[:nul].long()
    deg_i = loc.deg[nul:].long().clone()
    if item_degree_allreduce is not None:
        item_degree_allreduce(deg_i)
    max_deg = int(max(int(deg_u.max()) if nul else 0, int(deg_i.max()) if I else 0))
    lut = torch.from_numpy(degree_lut(max_deg, -0.5)).to(device)
    dis_u, dis_i = lut[deg_u], lut[deg_i]
    kw = {} if long_threshold is None else {"long_threshold": int(long_threshold)}
    # ---- A_u: the user rows; columns nul + item -> item
    n_u = int(ip[nul].item())
    cols_u = (loc.indices[:n_u] - nul).contiguous()
    rows_u = torch.repeat_interleave(torch.arange(nul, device=device), counts[:nul])
    vals_u = ((dis_u[rows_u] * loc.vals[:n_u]) * dis_i[cols_u.long()]).contiguous()
    a_u = NormAdjCSR(loc.indptr[:nul + 1].clone(), cols_u, vals_u, nul, part.items_padded, **kw)
    # ---- A_i: the item rows restricted to this rank's users (columns = local user row), padded to the common
    # height and ROTATED to start at the next rank's block (see local_csrs)
    cols_i = loc.indices[n_u:]
    own = counts[nul:]
    rows_i = torch.repeat_interleave(torch.arange(I, device=device), own)
    vals_i = (dis_i[rows_i] * loc.vals[n_u:]) * dis_u[cols_i.long()]
    iptr = torch.zeros(part.items_padded + 1, dtype=torch.int64, device=device)
    torch.cumsum(own, 0, out=iptr[1:I + 1])
    iptr[I + 1:] = iptr[I]
    shift = part.item_shift(rank)
    cut = int(iptr[shift].item())
    cnt = iptr[1:] - iptr[:-1]
    rptr = torch.zeros_like(iptr)
    torch.cumsum(torch.cat([cnt[shift:], cnt[:shift]]), 0, out=rptr[1:])
    a_i = NormAdjCSR(rptr.to(torch.int32), torch.cat([cols_i[cut:], cols_i[:cut]]).contiguous(),
                     torch.cat([vals_i[cut:], vals_i[:cut]]).contiguous(), part.items_padded, nul, **kw)
    return a_u, a_i


class ItemExchange:
    """Peer-mapped buffers of the 1.5-D propagation: two full item tables T[2] [items_padded, d] (layer parity)
    and one partial buffer P [items_padded, d] per rank, allocated with torch's symmetric-memory allocator so every
    rank can load peers' partials and store into peers' tables over NVLink."""

    def __init__(self, part: BipartitePartition, d: int, device, group=None):
        import ctypes

        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.part, self.d = part, d

        def alloc():
            t = symm_mem.empty((part.items_padded, d), dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, self.group)
            t.zero_()
            return t, h, (ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])

        self.tables, self.table_ptrs, self.handles = [], [], []
        for _ in range(2):
            t, h, arr = alloc()
            self.tables.append(t)
            self.table_ptrs.append(arr)
            self.handles.append(h)
        self.partial, self.partial_handle, self.partial_ptrs = alloc()          # staging: G slots of Ib rows
        self.slot_ptrs = _slot_ptrs(self.partial, part)
        self.partial_local = torch.empty((part.items_padded, d), dtype=torch.float32, device=device)
        n_copy = int(os.environ.get("GR_UO_COPY_STREAMS", str(self.world)))
        self.copy_streams = [torch.cuda.Stream(device) for _ in range(max(0, n_copy))]
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def barrier(self):
        self.partial_handle.barrier()


class EmulatedItemExchange:
    """The same buffers for G ranks emulated on ONE device (tests): plain tensors, pointer tables over them."""

    def __init__(self, part: BipartitePartition, d: int, device):
        import ctypes

        G = part.world_size
        self.world, self.part, self.d = G, part, d
        self._tables = [[torch.zeros((part.items_padded, d), device=device) for _ in range(G)] for _ in range(2)]
        self._partials = [torch.zeros((part.items_padded, d), device=device) for _ in range(G)]
        self.table_ptrs = [(ctypes.c_void_p * G)(*[t.data_ptr() for t in self._tables[k]]) for k in range(2)]
        self.partial_ptrs = (ctypes.c_void_p * G)(*[t.data_ptr() for t in self._partials])
        self._partial_local = [torch.zeros((part.items_padded, d), device=device) for _ in range(G)]
        self.copy_streams = []

    def view(self, rank: int):
        ex = EmulatedItemExchange.__new__(EmulatedItemExchange)
        ex.__dict__.update(self.__dict__)
        ex.rank = rank
        ex.tables = [self._tables[0][rank], self._tables[1][rank]]
        ex.partial = self._partials[rank]
        ex.slot_ptrs = _slot_ptrs(ex.partial, self.part)
        ex.partial_local = self._partial_local[rank]
        return ex

    def barrier(self):
        pass


def _slot_ptrs(partial: torch.Tensor, part: "BipartitePartition"):
    """The G slots of a rank's staging buffer: slot k = rows [k*Ib, (k+1)*Ib) = the partial rows of this rank's item
    block pushed by rank k."""
    import ctypes

    step = part.item_block * partial.stride(0) * 4
    return (ctypes.c_void_p * part.world_size)(*[partial.data_ptr() + k * step for k in range(part.world_size)])


def _reduce_bcast(ex, row0: int, n_rows: int, d: int, dst_ptrs, n_dst: int, addend, out, scale: float, mode: int):
    """Sum the G slots of the local staging buffer (rank order), store the block into rows [row0, ..) of the given
    item tables, fold into the layer sum."""
    from ._lib import check, lib, ptr, stream_ptr

    with torch.cuda.device(ex.partial.device):
        check(lib().gr_reduce_bcast_rows(ex.slot_ptrs, ex.world, d, dst_ptrs, n_dst, d, 0, row0, n_rows, d,
                                         ptr(addend), addend.stride(0) if addend is not None else 0, ptr(out),
                                         out.stride(0) if out is not None else 0, None, 0, float(scale), int(mode),
                                         stream_ptr()), "gr_reduce_bcast_rows")


def _copy_async(dst_ptr: int, src_ptr: int, nbytes: int, device):
    from ._lib import check, lib, stream_ptr

    if nbytes > 0:
        with torch.cuda.device(device):
            check(lib().gr_peer_copy_async(dst_ptr, src_ptr, nbytes, stream_ptr()), "gr_peer_copy_async")


def _propagate_user_owner_steps(a_u: NormAdjCSR, a_i: NormAdjCSR, ex, xu0: torch.Tensor, xi0_block: torch.Tensor,
                                n_layers: int, result: dict):
    """Generator: one rank's LightGCN propagation in the user-owner layout; yields at every point where all
    ranks must have arrived (cross-rank barrier).  ``xu0`` [n_users_local, d]: the rank's user rows;
    ``xi0_block`` [item block rows, d]: the rank's block of the item table.  result['users'] / ['items'] receive
    mean_l of the rank's user rows / item block.

    Per layer l (main stream = SpMM kernels, copy streams = cudaMemcpyAsync over NVLink, no SM involved):
        main   P = A_i x_u(l)                     partial sums of ALL item rows, written locally
        copy   block g of P  ->  slot `rank` of rank g's staging buffer        (G-1 copies, beside the next SpMM)
        main   x_u(l+1) = A_u T(l)                this rank's users from the replicated item table
        ---- barrier: every staging buffer is complete
        main   block = slot_0 + slot_1 + ... (rank order) -> own rows of T(l+1), layer sum     (local reduce)
        copy   own block  ->  rows [lo, hi) of every other rank's T(l+1)        (G-1 copies, beside the next A_i SpMM)
        ---- barrier (before the next A_u SpMM): every T(l+1) is complete"""
    part, d, rank, G = ex.part, xu0.shape[1], ex.rank, ex.world
    Ib = part.item_block
    lo, hi = part.item_range(rank)
    nb = hi - lo
    L = n_layers
    if L == 0:
        result["users"], result["items"] = xu0.clone(), xi0_block[:nb].clone()
        return
    dev = xu0.device
    row_bytes = d * 4
    main = torch.cuda.current_stream(dev)
    serial = os.environ.get("GR_UO_SERIAL", "0") == "1" or not ex.copy_streams
    streams = [main] * G if serial else [ex.copy_streams[k % len(ex.copy_streams)] for k in range(G)]
    pending = []                                     # copy streams with work the main stream has not waited for

    copy_mode = os.environ.get("GR_UO_COPY", "sm")
    copy_ctas = int(os.environ.get("GR_UO_COPY_CTAS", "32"))    # 8 B200s, C5 ms/step: 16 CTAs 71.2, 24 61.1, 32 50.0, 64 52.3, 128 59.2

    def fan_out(copies):
        """copies: (peer k, dst ptr, src ptr, bytes), issued after what main has enqueued so far: 'sm' = one small
        multi-copy kernel on a side stream (gr_peer_copy_multi), 'dma' = one cudaMemcpyAsync per copy, each on its
        own copy stream (copy engines)."""
        ev = None if serial else torch.cuda.Event()
        if ev is not None:
            ev.record(main)
        if copy_mode == "sm":
            import ctypes

            from ._lib import check, lib, stream_ptr
            live = [c for c in copies if c[3] > 0]
            if not live:
                return
            st = main if serial else ex.copy_streams[0]
            if st is not main:
                st.wait_event(ev)
                pending.append(st)
            n = len(live)
            dsts = (ctypes.c_void_p * n)(*[c[1] for c in live])
            srcs = (ctypes.c_void_p * n)(*[c[2] for c in live])
            sizes = (ctypes.c_size_t * n)(*[c[3] for c in live])
            with (torch.cuda.stream(st) if st is not main else _NullCtx()), torch.cuda.device(dev):
                check(lib().gr_peer_copy_multi(dsts, srcs, sizes, n, copy_ctas, stream_ptr()), "gr_peer_copy_multi")
            return
        for k, dst, src, nbytes in copies:
            st = streams[k]
            if st is not main:
                st.wait_event(ev)
                pending.append(st)
            with (torch.cuda.stream(st) if st is not main else _NullCtx()):
                _copy_async(dst, src, nbytes, dev)

    def join():
        for st in pending:
            main.wait_stream(st)
        pending.clear()

    yield "enter"                                    # nobody still reads the tables / staging from a previous call
    cur = 0
    # layer-0 item table: own block into every rank's T[0] (own copy on main, the others by the copy engines)
    src0 = xi0_block.contiguous()
    fan_out([(k, int(ex.table_ptrs[cur][k]) + lo * row_bytes, src0.data_ptr(), nb * row_bytes) for k in range(G)])
    acc_u = torch.empty_like(xu0)
    out_u = torch.empty_like(xu0)
    acc_i = torch.empty((max(nb, 1), d), dtype=torch.float32, device=dev)
    out_i = torch.empty((max(nb, 1), d), dtype=torch.float32, device=dev)
    xu = xu0
    xu_next = [torch.empty_like(xu0), torch.empty_like(xu0)]
    p_local = ex.partial_local                       # [items_padded, d], rows in A_i's rotated order
    shift_blocks = (rank + 1) % G
    for l in range(L):
        last = l == L - 1
        nxt = cur ^ 1
        # (a) partial sums of ALL item rows over this rank's users (rotated row order: block k of P is item block
        #     (k + rank + 1) mod G)
        a_i.spmm(xu, y=p_local, want_y=True)
        if l > 0:
            join()                                   # the broadcast of T[cur] (this rank's copies) has been issued...
            yield "table"                            # ... and everybody's has landed: T[cur] is complete
        else:
            join()
            yield "table0"
        # (b) push block k of P to its owner's staging slot `rank` (copy engines), beside (c)
        copies = []
        for k in range(G):
            owner = (k + shift_blocks) % G
            copies.append((owner, int(ex.partial_ptrs[owner]) + rank * Ib * row_bytes,
                           p_local.data_ptr() + k * Ib * row_bytes, Ib * row_bytes))
        fan_out(copies)
        # (c) this rank's users from the replicated item table of layer l
        addend_u = xu0 if l == 0 else acc_u
        if last:
            a_u.spmm(ex.tables[cur], addend=addend_u, out=out_u, scale=float(L + 1), scale_mode=_lib.GR_SCALE_DIV,
                     want_y=False)
        else:
            a_u.spmm(ex.tables[cur], y=xu_next[l & 1], addend=addend_u, out=acc_u)
            xu = xu_next[l & 1]
        join()
        yield "partials"                             # every rank's staging buffer is complete
        # (d) local reduce of the G slots (rank order) -> own rows of T[nxt] + layer sum; (e) broadcast by copies
        if nb > 0:
            addend = xi0_block if l == 0 else acc_i
            if last:
                _reduce_bcast(ex, lo, nb, d, None, 0, addend, out_i, float(L + 1), _lib.GR_SCALE_DIV)
            else:
                own = (ctypes_array([int(ex.table_ptrs[nxt][rank])]), 1)
                _reduce_bcast(ex, lo, nb, d, own[0], 1, addend, acc_i, 1.0, _lib.GR_SCALE_NONE)
        if not last:
            src = int(ex.table_ptrs[nxt][rank]) + lo * row_bytes
            fan_out([(k, int(ex.table_ptrs[nxt][k]) + lo * row_bytes, src, nb * row_bytes) for k in range(G) if k != rank])
            cur = nxt
    join()
    result["users"], result["items"] = out_u, out_i[:nb]


def ctypes_array(ptrs):
    import ctypes

    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def lightgcn_propagate_user_owner(a_u: NormAdjCSR, a_i: NormAdjCSR, ex, xu0: torch.Tensor, xi0_block: torch.Tensor,
                                  n_layers: int):
    """-> (users_local [n_users_local, d], items_block [block rows, d]) = mean_{l<=L} of the propagated rows."""
    res = {}
    for _ in _propagate_user_owner_steps(a_u, a_i, ex, xu0, xi0_block, n_layers, res):
        ex.barrier()
    return res["users"], res["items"]


def emulate_user_owner(full: NormAdjCSR, n_users: int, n_items: int, x0: torch.Tensor, n_layers: int, world: int):
    """All G ranks of the user-owner propagation on ONE device, advanced in lockstep between barriers (tests and
    single-GPU validation).  Returns the assembled [N, d] result in natural row order."""
    part = BipartitePartition(n_users, n_items, world)
    d = x0.shape[1]
    shared = EmulatedItemExchange(part, d, x0.device)
    gens, results = [], []
    for r in range(world):
        a_u, a_i = part.local_csrs(full, r)
        lo, hi = part.item_range(r)
        blk = torch.zeros((part.item_block, d), device=x0.device)
        blk[: hi - lo] = x0[n_users + lo: n_users + hi]
        res = {}
        results.append(res)
        gens.append(_propagate_user_owner_steps(a_u, a_i, shared.view(r), part.take_users(x0[:n_users], r), blk,
                                                n_layers, res))
    live = list(range(world))
    while live:
        nxt = []
        for r in live:
            try:
                next(gens[r])
                nxt.append(r)
            except StopIteration:
                pass
        torch.cuda.synchronize(x0.device)
        live = nxt
    out = torch.empty_like(x0)
    for r in range(world):
        out[r:n_users:world] = results[r]["users"]
        lo, hi = part.item_range(r)
        out[n_users + lo: n_users + hi] = results[r]["items"]
    return out


class UserOwnerLightGCN:
    """ShardedLightGCN in the user-owner layout: parameters (and Adam state) are this rank's user rows and its
    block of the item table.  One step = forward propagation (lightgcn_propagate_user_owner), the replicated
    B x B BPR on the 3B touched rows (collected with one all-reduce; each row has one owner), backward = the same
    propagation applied to the gradient (mean_l Â^l is symmetric and the layout is the same), global-norm clip,
    fused Adam on the local rows."""

    def __init__(self, a_u: NormAdjCSR, a_i: NormAdjCSR, ex, xu0: torch.Tensor, xi0_block: torch.Tensor, n_layers: int,
                 lr: float = 1e-3, weight_decay: float = 1e-4, max_grad_norm: float = 1.0, group=None):
        self.a_u, self.a_i, self.ex, self.part, self.group = a_u, a_i, ex, ex.part, group
        self.rank, self.n_layers, self.max_grad_norm = ex.rank, int(n_layers), float(max_grad_norm)
        self.lo, self.hi = self.part.item_range(self.rank)
        self.users = torch.nn.Parameter(xu0.detach().clone())
        self.items = torch.nn.Parameter(xi0_block.detach().clone())          # [item_block, d]; rows >= hi - lo unused
        self.optimizer = torch.optim.Adam([self.users, self.items], lr=lr, weight_decay=weight_decay)

    def propagate(self, xu: torch.Tensor, xi_block: torch.Tensor):
        return lightgcn_propagate_user_owner(self.a_u, self.a_i, self.ex, xu, xi_block, self.n_layers)

    def _all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def train_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> float:
        from .losses import bpr_fused
        from .optim import fused_clip_adam_step

        G, B, dev = self.part.world_size, int(users.numel()), self.users.device
        users, items = users.to(dev).view(-1).long(), torch.cat([pos.to(dev).view(-1), neg.to(dev).view(-1)]).long()
        with torch.no_grad():
            out_u, out_i = self.propagate(self.users.detach(), self.items.detach())
            mine_u = (users % G) == self.rank
            mine_i = (items >= self.lo) & (items < self.hi)
            table = torch.zeros((3 * B, out_u.shape[1]), dtype=out_u.dtype, device=dev)
            table[:B][mine_u] = out_u[users[mine_u] // G]
            table[B:][mine_i] = out_i[items[mine_i] - self.lo]
            self._all_reduce(table)
        ar = torch.arange(B, device=dev)
        table.requires_grad_(True)
        loss = bpr_fused(table, B, ar, ar, (B + ar).view(-1, 1))
        loss.backward()
        with torch.no_grad():
            gu = torch.zeros_like(self.users)
            gi = torch.zeros_like(self.items)
            gu.index_add_(0, users[mine_u] // G, table.grad[:B][mine_u])           # duplicates accumulate
            gi.index_add_(0, items[mine_i] - self.lo, table.grad[B:][mine_i])
            du, di = self.propagate(gu, gi)                                        # mean_l Â^l g
            self.users.grad = du
            self.items.grad = torch.zeros_like(self.items)
            self.items.grad[: self.hi - self.lo] = di
            if self.max_grad_norm > 0:
                sq = self._all_reduce((du.double().pow(2).sum() + di.double().pow(2).sum()).view(1))
                coef = self.max_grad_norm / (float(sq.sqrt()) + 1e-6)
                if coef < 1.0:
                    self.users.grad.mul_(coef)
                    self.items.grad.mul_(coef)
            fused_clip_adam_step(self.optimizer, 0.0)                              # already clipped with the GLOBAL norm
        return float(loss.detach())
