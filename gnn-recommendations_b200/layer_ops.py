"""Autograd wrappers of the propagation kernels shared by the model classes.

Forward AND backward passes run the hand-written kernels of libgr_b200.so: the backward of the sparse
product is the same SpMM kernel on Â^T, the backward of the dense per-row epilogues is gr_rowmap_bwd
(input gradients + deterministic weight-gradient reduction) and the backward of the GAT edge-softmax is
gr_gat_bwd (softmax weights recomputed per row from the stored max / normaliser).  No stock torch GEMM /
index kernel is left on the training step (what autograd would run under trainer.py:270).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .graph_builder import NormAdjCSR

ACT_NONE, ACT_LEAKY, ACT_ELU = 0, 1, 2


class DropSeed:
    """Seed of one fused-dropout call: ``value`` (host integer, a kernel argument) plus, inside a CUDA-graph
    capture, ``dev`` — a device int64 the kernels ADD to it at run time, refreshed by the Trainer before
    every replay, so a step captured once still draws a new mask per step."""
    __slots__ = ("value", "dev")

    def __init__(self, value: int, dev: Optional[torch.Tensor] = None):
        self.value, self.dev = int(value) & 0x7FFFFFFFFFFFFFFF, dev


_SEED_STATE = {"dev": None, "count": 0}


class device_dropout_seeds:
    """Context: dropout seeds drawn inside come from ``dev`` (device int64[1]) plus a per-call offset."""

    def __init__(self, dev: torch.Tensor):
        self.dev = dev

    def __enter__(self):
        self.prev = dict(_SEED_STATE)
        _SEED_STATE["dev"], _SEED_STATE["count"] = self.dev, 0
        return self

    def drawn(self) -> int:
        """Number of seeds handed out inside the context so far."""
        return _SEED_STATE["count"] if _SEED_STATE["dev"] is self.dev else self._count

    def __exit__(self, *a):
        self._count = _SEED_STATE["count"]
        _SEED_STATE.update(self.prev)
        return False


def new_dropout_seed() -> DropSeed:
    """A 63-bit seed drawn from torch's global CPU generator (the stream the reference's nn.Dropout
    consumes), so torch.manual_seed makes train-mode runs reproducible; inside
    ``device_dropout_seeds`` a distinct constant offset onto the device-resident step seed instead."""
    if _SEED_STATE["dev"] is not None:
        _SEED_STATE["count"] += 1
        return DropSeed(_SEED_STATE["count"] * 0x9E3779B97F4A7C15, _SEED_STATE["dev"])
    return DropSeed(int(torch.empty((), dtype=torch.int64).random_().item()))


def _seed_parts(seed):
    if isinstance(seed, DropSeed):
        return seed.value, seed.dev
    return int(seed), None


class _Spmm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, csr):
        ctx.csr = csr
        y, _ = csr.spmm(x.contiguous())
        return y

    @staticmethod
    def backward(ctx, g):
        y, _ = ctx.csr.transpose().spmm(g.contiguous())
        return y, None


def spmm(csr: NormAdjCSR, x: torch.Tensor) -> torch.Tensor:
    """Â x with autograd (torch.sparse.mm drop-in: lightgcn.py:88, ngcf.py:70, model.py:172)."""
    return _Spmm.apply(x, csr)


def _rows(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Row-major with unit inner stride (any row stride: column slices of a wider matrix are fine)."""
    if t is None:
        return None
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    return t.contiguous()


def _ld(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.stride(0)


def _rowmap_raw(x1, wa, ba, x2, x3, wb, bb, resid, alpha, beta, act, slope, drop_p=0.0, drop_seed=0, out=None):
    seed, seed_dev = _seed_parts(drop_seed)
    n, d_in = x1.shape
    d_out = wa.shape[1]
    if out is None:
        out = torch.empty((n, d_out), dtype=torch.float32, device=x1.device)
    with torch.cuda.device(x1.device):
        check(lib().gr_rowmap_f32(
            ptr(x1), _ld(x1), ptr(wa), ptr(ba), ptr(x2), _ld(x2), ptr(x3), _ld(x3), ptr(wb), ptr(bb), ptr(resid),
            _ld(resid), float(alpha), float(beta), int(act), float(slope), n, d_in, d_out, float(drop_p),
            seed, ptr(seed_dev), ptr(out), out.stride(0), stream_ptr()), "gr_rowmap_f32")
    return out


def _rowmap_bwd_raw(g, out, x1, wa, x2, x3, wb, resid, alpha, beta, act, slope, drop_p, drop_seed, need_dx, need_dx2,
                    need_dx3, need_dresid, need_dw, dw_out=None):
    """-> (dx1, dx2, dx3, dresid, dw) — dw is the flat [nw*d_in*d_out + d_out] buffer of gr_rowmap_bwd
    (``dw_out``: write it there instead of a new tensor)."""
    seed, seed_dev = _seed_parts(drop_seed)
    n, d_in = x1.shape
    d_out = wa.shape[1]
    dev = x1.device
    has_b = wb is not None
    l = lib()
    dx1 = torch.empty((n, d_in), dtype=torch.float32, device=dev) if need_dx else None
    dx2 = torch.empty((n, d_in), dtype=torch.float32, device=dev) if (need_dx2 and has_b) else None
    dx3 = torch.empty((n, d_in), dtype=torch.float32, device=dev) if (need_dx3 and has_b) else None
    dres = torch.empty((n, d_out), dtype=torch.float32, device=dev) if need_dresid else None
    total = (2 if has_b else 1) * d_in * d_out + d_out
    dw = (dw_out if dw_out is not None else torch.empty(total, dtype=torch.float32, device=dev)) if need_dw else None
    ws_bytes = l.gr_rowmap_bwd_workspace_bytes(n, d_in, d_out, int(has_b))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(l.gr_rowmap_bwd(
            ptr(g), _ld(g), ptr(out), _ld(out), ptr(x1), _ld(x1), ptr(wa), ptr(x2), _ld(x2), ptr(x3), _ld(x3), ptr(wb),
            ptr(resid), _ld(resid), float(alpha), float(beta), int(act), float(slope), n, d_in, d_out, float(drop_p),
            seed, ptr(seed_dev), ptr(dx1), _ld(dx1), ptr(dx2), _ld(dx2), ptr(dx3), _ld(dx3), ptr(dres), _ld(dres), ptr(dw),
            ptr(ws), ws_bytes, stream_ptr()), "gr_rowmap_bwd")
    return dx1, dx2, dx3, dres, dw


class _RowMap(torch.autograd.Function):
    """tensors = (x1, wa, ba, x2, x3, wb, bb, resid); ``x3_is_x1``: the bi-interaction's second factor is x1
    itself (NGCF: both are Â x) — its gradient is folded into x1's inside the kernel."""

    @staticmethod
    def forward(ctx, alpha, beta, act, slope, drop_p, drop_seed, x3_is_x1, x1, wa, ba, x2, x3, wb, bb, resid):
        x1, x2, x3, resid = _rows(x1), _rows(x2), _rows(x3), _rows(resid)
        wa = wa.contiguous()
        wb = None if wb is None else wb.contiguous()
        ba = None if ba is None else ba.contiguous()
        bb = None if bb is None else bb.contiguous()
        x3k = x1 if x3_is_x1 else x3
        out = _rowmap_raw(x1, wa, ba, x2, x3k, wb, bb, resid, alpha, beta, act, slope, drop_p, drop_seed)
        ctx.cfg = (alpha, beta, act, slope, drop_p, drop_seed, x3_is_x1)
        ctx.has = (ba is not None, x2 is not None, x3 is not None, wb is not None, bb is not None, resid is not None)
        keep_out = out if act != ACT_NONE else None
        keep_res = resid if (act != ACT_NONE and resid is not None) else None
        ctx.save_for_backward(x1, wa, x2, x3, wb, keep_res, keep_out)
        return out

    @staticmethod
    def backward(ctx, g):
        x1, wa, x2, x3, wb, resid, out = ctx.saved_tensors
        alpha, beta, act, slope, drop_p, drop_seed, x3_is_x1 = ctx.cfg
        has_ba, has_x2, has_x3, has_wb, has_bb, has_res = ctx.has
        need = ctx.needs_input_grad[7:]     # x1, wa, ba, x2, x3, wb, bb, resid
        g = _rows(g)
        x3k = x1 if x3_is_x1 else x3
        need_dx = need[0] or (has_wb and (need[3] or need[4]))
        need_dw = need[1] or need[2] or need[5] or need[6]
        dx1, dx2, dx3, dres, dw = _rowmap_bwd_raw(
            g, out, x1, wa, x2, x3k, wb, resid, alpha, beta, act, slope, drop_p, drop_seed, need_dx,
            has_wb and need[3], has_wb and need[4] and not x3_is_x1, has_res and need[7], need_dw)
        d_in, d_out = wa.shape
        dwa = dwb = db = None
        if dw is not None:
            dwa = dw[:d_in * d_out].view(d_in, d_out)
            if has_wb:
                dwb = dw[d_in * d_out:2 * d_in * d_out].view(d_in, d_out)
            db = dw[-d_out:]
        return (None,) * 7 + (dx1 if need[0] else None, dwa if need[1] else None, db if (has_ba and need[2]) else None,
                              dx2, dx3, dwb if need[5] else None, db if (has_bb and need[6]) else None, dres)


def rowmap(x1, wa, bias_a=None, x2=None, x3=None, wb=None, bias_b=None, resid=None, alpha: float = 1.0,
           beta: float = 0.0, act: int = ACT_NONE, slope: float = 0.0, drop_p: float = 0.0,
           drop_seed=0) -> torch.Tensor:
    """out = D * (alpha * act(x1 @ wa + bias_a + (x2 * x3) @ wb + bias_b) + beta * resid)  (gr_rowmap_f32);
    D = 1 unless drop_p > 0 (layer-output dropout, mask = hash(drop_seed, element))."""
    x3_is_x1 = x3 is not None and x3 is x1
    return _RowMap.apply(float(alpha), float(beta), int(act), float(slope), float(drop_p), drop_seed, x3_is_x1,
                         x1, wa, bias_a, x2, None if x3_is_x1 else x3, wb, bias_b, resid)


# ------------------------------------------------------------------------------------------------
# combination of the layer outputs (GAT: mean; Group-and-Shuffle: softmax-weighted sum)
# ------------------------------------------------------------------------------------------------
_INV_N = {}


def _combine_raw(xs, w_dev, out=None):
    n, d = xs[0].shape
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=xs[0].device)
    k = len(xs)
    ptrs = (C.c_void_p * k)(*[x.data_ptr() for x in xs])
    lds = (C.c_int64 * k)(*[x.stride(0) for x in xs])
    with torch.cuda.device(xs[0].device):
        check(lib().gr_layer_combine(ptrs, lds, k, ptr(w_dev), n, d, ptr(out), out.stride(0), stream_ptr()),
              "gr_layer_combine")
    return out


class _LayerCombine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, *xs):
        xs = [_rows(x) for x in xs]
        if len(xs) > 8:
            raise ValueError("at most 8 layer outputs")
        w = None if weights is None else weights.contiguous()
        ctx.n = len(xs)
        ctx.save_for_backward(w, *xs)
        return _combine_raw(xs, w)

    @staticmethod
    def backward(ctx, g):
        w, *xs = ctx.saved_tensors
        g = _rows(g)
        n, dev = ctx.n, g.device
        if w is None:                                   # mean: every input gets g / n (one kernel, shared result)
            key = (str(dev), n)
            inv = _INV_N.get(key)
            if inv is None:
                inv = _INV_N[key] = torch.full((1,), 1.0 / n, dtype=torch.float32, device=dev)
            gx = _combine_raw([g], inv)
            return (None,) + tuple(gx if ctx.needs_input_grad[1 + l] else None for l in range(n))
        dxs = tuple(_combine_raw([g], w[l:l + 1]) if ctx.needs_input_grad[1 + l] else None for l in range(n))
        dw = None
        if ctx.needs_input_grad[0]:
            l_ = lib()
            ws_bytes = l_.gr_layer_combine_bwd_workspace_bytes()
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            dw8 = torch.empty(8, dtype=torch.float32, device=dev)
            ptrs = (C.c_void_p * n)(*[x.data_ptr() for x in xs])
            lds = (C.c_int64 * n)(*[x.stride(0) for x in xs])
            with torch.cuda.device(dev):
                check(l_.gr_layer_combine_dw(ptrs, lds, n, ptr(g), g.stride(0), g.shape[0], g.shape[1], ptr(dw8), ptr(ws),
                                             ws_bytes, stream_ptr()), "gr_layer_combine_dw")
            dw = dw8[:n]
        return (dw,) + dxs


def layer_combine(xs: Sequence[torch.Tensor], weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum_l weights[l] * xs[l] (device weights, e.g. softmax(layer_weights): model.py:204-207) or, with
    ``weights=None``, torch.mean(torch.stack(xs), 0) (gat.py:287-288) — one kernel each way (gr_layer_combine)."""
    return _LayerCombine.apply(weights, *xs)


# ------------------------------------------------------------------------------------------------
# Group-and-Shuffle: the composed per-layer map M_l = W_conn,l W_orth,l[:, perm]
# ------------------------------------------------------------------------------------------------
class _GsCompose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, skew, perm_c, perm_g):
        skew = skew.contiguous()
        n_layers, n_sets, nb, bs, _ = skew.shape
        d = nb * bs
        blocks = torch.empty_like(skew)
        m = torch.empty((n_layers, d, d), dtype=torch.float32, device=skew.device)
        with torch.cuda.device(skew.device):
            check(lib().gr_gs_compose(ptr(skew), ptr(perm_c), ptr(perm_g), n_layers, n_sets, d, bs, ptr(blocks),
                                      ptr(m), stream_ptr()), "gr_gs_compose")
        ctx.save_for_backward(skew, blocks, perm_c, perm_g)
        return m

    @staticmethod
    def backward(ctx, dm):
        skew, blocks, perm_c, perm_g = ctx.saved_tensors
        n_layers, n_sets, nb, bs, _ = skew.shape
        dm = dm.contiguous()
        dblocks = torch.empty_like(skew)
        dskew = torch.empty_like(skew)
        with torch.cuda.device(skew.device):
            check(lib().gr_gs_compose_bwd(ptr(skew), ptr(blocks), ptr(perm_c), ptr(perm_g), n_layers, n_sets, nb * bs,
                                          bs, ptr(dm), ptr(dblocks), ptr(dskew), stream_ptr()), "gr_gs_compose_bwd")
        return dskew, None, None


def gs_compose(skew: torch.Tensor, perm_c: Optional[torch.Tensor], perm_g: torch.Tensor) -> torch.Tensor:
    """skew [L, S, nb, bs, bs] (S = 2: connection then local blocks; S = 1: local only), perm_c / perm_g [L, d]
    int64 -> M [L, d, d] with M_l = blockdiag(exp(P - P^T))[:, perm_c] @ blockdiag(exp(Q - Q^T))[:, perm_g]
    (bundle_layer.py:59-73, group_shuffle_layer.py:88-129, model.py:176); differentiable w.r.t. skew."""
    return _GsCompose.apply(skew, None if perm_c is None else perm_c.contiguous(), perm_g.contiguous())


class _SpmmMap(torch.autograd.Function):
    """out = alpha (Â x) M' + beta R with M' = M or M^T (gr_spmm_csr_map_f32), differentiable w.r.t. x, M and R.
    Backward: dx = alpha (Â^T g) M'^T (the same fused kernel), dM' = (alpha Â x)^T g (gr_rowmap_bwd's
    weight-gradient kernels on the stored aggregate), dR = beta g (gr_layer_combine)."""

    @staticmethod
    def forward(ctx, x, m, resid, csr, alpha, beta, transposed):
        x, m, resid = _rows(x), m.contiguous(), _rows(resid)
        keep_c = ctx.needs_input_grad[1]
        out, c = csr.spmm_map(x, m, alpha, beta, addend=resid, transposed=transposed, want_y=keep_c)
        ctx.csr, ctx.cfg = csr, (float(alpha), float(beta), bool(transposed), resid is not None)
        ctx.save_for_backward(m, c)
        return out

    @staticmethod
    def backward(ctx, g):
        m, c = ctx.saved_tensors
        alpha, beta, transposed, has_res = ctx.cfg
        g = _rows(g)
        dev, d = g.device, int(g.shape[1])
        dx = dm = dres = None
        if ctx.needs_input_grad[0]:
            dx, _ = ctx.csr.transpose().spmm_map(g, m, alpha, 0.0, transposed=not transposed)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(d * d + d, dtype=torch.float32, device=dev)
            _rowmap_bwd_raw(g, None, c, m, None, None, None, None, 1.0, 0.0, ACT_NONE, 0.0, 0.0, 0,
                            False, False, False, False, True, dw_out=dw)
            dm = dw[:d * d].view(d, d)                   # (alpha Â x)^T g = dL/dM'
            if transposed:
                dm = dm.t()
        if has_res and ctx.needs_input_grad[2]:
            key = (str(dev), 1, beta)
            wb = _GS_X0_WEIGHTS.get(key)
            if wb is None:
                wb = _GS_X0_WEIGHTS[key] = torch.tensor([beta], dtype=torch.float32, device=dev)
            dres = _combine_raw([g], wb)
        return dx, dm, dres, None, None, None, None


def spmm_map(csr: NormAdjCSR, x: torch.Tensor, m: torch.Tensor, alpha: float = 1.0, beta: float = 0.0,
             resid: Optional[torch.Tensor] = None, transposed: bool = False) -> torch.Tensor:
    """alpha (Â x) M + beta resid  (M^T with ``transposed``) in one kernel, with autograd; needs csr.supports_map(d)."""
    return _SpmmMap.apply(x, m, resid, csr, float(alpha), float(beta), bool(transposed))


class _GsPropagate(torch.autograd.Function):
    """The whole Group-and-Shuffle propagation (model.py:164-207) as ONE autograd node:

        x_{l+1} = alpha (Â x_l) M_l + beta x_0      (l = 0 .. L-1; alpha = 1 - residual_alpha, beta = residual_alpha)
        out     = sum_l w_l x_l

    Forward: one gr_spmm_csr_map_f32 per layer (sparse product, dense map and residual in a single kernel; it also
    leaves c_l = alpha Â x_l for the weight gradient when a backward pass will follow) + gr_layer_combine.
    Backward, written out by hand so that no elementwise torch kernel is left between the layers:
        gx_L = w_L G;   for l = L .. 1:   dM_l = c_{l-1}^T gx_l   (gr_rowmap_bwd, weight-gradient part)
                                          gx_{l-1} = w_{l-1} G + alpha (Â^T gx_l) M_l^T   (the same fused kernel)
        dx_0 = gx_0 + beta (gx_1 + .. + gx_L)  (gr_layer_combine),   dw_l = <G, x_l>  (gr_layer_combine_dw)."""

    @staticmethod
    def forward(ctx, x0, ms, w, csr, alpha, beta):
        x0, ms, w = _rows(x0), ms.contiguous(), w.contiguous()
        n_layers = int(ms.shape[0])
        keep = any(ctx.needs_input_grad[:3])
        xs, cs, x = [x0], [], x0
        for l in range(n_layers):
            x, c = csr.spmm_map(x, ms[l], alpha, beta, addend=x0, want_y=keep)
            xs.append(x)
            cs.append(c)
        ctx.csr, ctx.cfg = csr, (float(alpha), float(beta), n_layers)
        if keep:
            ctx.save_for_backward(ms, w, *xs, *cs)
        return _combine_raw(xs, w)

    @staticmethod
    def backward(ctx, G):
        ms, w, *rest = ctx.saved_tensors
        alpha, beta, n_layers = ctx.cfg
        xs, cs = rest[:n_layers + 1], rest[n_layers + 1:]
        G = _rows(G)
        dev, d = G.device, int(G.shape[1])
        csr_t = ctx.csr.transpose()
        dms = torch.empty((n_layers, d * d + d), dtype=torch.float32, device=dev)     # [dM_l | column sums (unused)]
        gxs = [None] * (n_layers + 1)
        gxs[n_layers] = _combine_raw([G], w[n_layers:n_layers + 1])
        for l in range(n_layers, 0, -1):
            g = gxs[l]
            _rowmap_bwd_raw(g, None, cs[l - 1], ms[l - 1], None, None, None, None, 1.0, 0.0, ACT_NONE, 0.0, 0.0, 0,
                            False, False, False, False, True, dw_out=dms[l - 1])
            gxs[l - 1], _ = csr_t.spmm_map(g, ms[l - 1], alpha, 1.0, addend=G, beta_dev=w[l - 1:l], transposed=True)
        key = (str(dev), n_layers, beta)
        wx0 = _GS_X0_WEIGHTS.get(key)
        if wx0 is None:
            wx0 = _GS_X0_WEIGHTS[key] = torch.tensor([1.0] + [beta] * n_layers, dtype=torch.float32, device=dev)
        dx0 = _combine_raw(gxs, wx0)
        l_ = lib()
        ws_bytes = l_.gr_layer_combine_bwd_workspace_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        dw8 = torch.empty(8, dtype=torch.float32, device=dev)
        n = n_layers + 1
        ptrs = (C.c_void_p * n)(*[x.data_ptr() for x in xs])
        lds = (C.c_int64 * n)(*[x.stride(0) for x in xs])
        with torch.cuda.device(dev):
            check(l_.gr_layer_combine_dw(ptrs, lds, n, ptr(G), G.stride(0), G.shape[0], d, ptr(dw8), ptr(ws), ws_bytes,
                                         stream_ptr()), "gr_layer_combine_dw")
        return dx0, dms[:, :d * d].view(n_layers, d, d), dw8[:n], None, None, None


_GS_X0_WEIGHTS = {}


def gs_propagate(csr: NormAdjCSR, x0: torch.Tensor, ms: torch.Tensor, w: torch.Tensor, alpha: float,
                 beta: float) -> torch.Tensor:
    """sum_l w[l] x_l with x_{l+1} = alpha (Â x_l) ms[l] + beta x_0 — the Group-and-Shuffle propagation with the
    dense map fused into the SpMM epilogue (gr_spmm_csr_map_f32), forward and backward.  Needs
    ``csr.supports_map(d)`` and at most 7 layers."""
    return _GsPropagate.apply(x0, ms, w, csr, float(alpha), float(beta))


def gs_propagate_supported(csr: NormAdjCSR, d: int, n_layers: int) -> bool:
    return csr.supports_map(d) and n_layers + 1 <= 8


# ------------------------------------------------------------------------------------------------
# GAT layer
# ------------------------------------------------------------------------------------------------
def _gat_forward_kernels(csr, x, wcat, a_self, a_neigh, heads, dh, slope, mean_heads, elu, drop_p=0.0, drop_seed=0,
                         keep=False):
    n = x.shape[0]
    dev = x.device
    seed, seed_dev = _seed_parts(drop_seed)
    h = _rowmap_raw(x, wcat, None, None, None, None, None, None, 1.0, 0.0, ACT_NONE, 0.0)     # [N, heads*dh]
    s = torch.empty((n, heads), dtype=torch.float32, device=dev)
    t = torch.empty((n, heads), dtype=torch.float32, device=dev)
    out = torch.empty((n, dh if mean_heads else heads * dh), dtype=torch.float32, device=dev)
    m = z = raw = None
    if keep:
        m, z = torch.empty_like(s), torch.empty_like(s)
    l = lib()
    segs, n_seg, _ = csr.gat_segments()
    ws_bytes = l.gr_gat_aggregate_workspace_bytes(n_seg, heads, dh) if n_seg else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    with torch.cuda.device(dev):
        check(l.gr_gat_node_scores(ptr(h), h.stride(0), ptr(a_self), ptr(a_neigh), n, heads, dh, ptr(s), ptr(t),
                                   stream_ptr()), "gr_gat_node_scores")
        check(l.gr_gat_aggregate(ptr(csr.indptr), ptr(csr.indices), n, ptr(h), h.stride(0), ptr(s), ptr(t), heads,
                                 dh, float(slope), int(mean_heads), int(elu), float(drop_p), seed, ptr(seed_dev),
                                 csr.n_cols, C.byref(segs) if segs is not None else None, ptr(out), out.stride(0),
                                 ptr(m), ptr(z), ptr(ws), ws_bytes, stream_ptr()), "gr_gat_aggregate")
    return out, (h, s, t, m, z, raw)


class _GatLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wcat, a_self, a_neigh, csr, heads, dh, slope, mean_heads, elu, drop_p, drop_seed):
        x, wcat, a_self, a_neigh = _rows(x), wcat.contiguous(), a_self.contiguous(), a_neigh.contiguous()
        keep = any(ctx.needs_input_grad[:4])
        out, (h, s, t, m, z, raw) = _gat_forward_kernels(csr, x, wcat, a_self, a_neigh, heads, dh, slope, mean_heads,
                                                         elu, drop_p, drop_seed, keep)
        ctx.csr, ctx.cfg = csr, (heads, dh, slope, mean_heads, elu, drop_p, drop_seed)
        if keep:
            ctx.save_for_backward(x, wcat, a_self, a_neigh, h, s, t, m, z, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, wcat, a_self, a_neigh, h, s, t, m, z, out = ctx.saved_tensors
        heads, dh, slope, mean_heads, elu, drop_p, drop_seed = ctx.cfg
        csr = ctx.csr
        csr_t = csr.transpose()          # the pattern of Âᵀ (Â itself for the symmetric bipartite graph)
        g = _rows(g)
        n, width, dev = x.shape[0], heads * dh, x.device
        seed, seed_dev = _seed_parts(drop_seed)
        l = lib()
        dH = torch.empty((n, width), dtype=torch.float32, device=dev)
        da = torch.empty(2 * width, dtype=torch.float32, device=dev)
        rsegs, n_rseg, _ = csr.gat_segments()
        csegs, n_cseg, n_clong = csr_t.gat_segments()
        ws_bytes = l.gr_gat_bwd_workspace_bytes(n, heads, dh, n_rseg, n_cseg, n_clong)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(l.gr_gat_bwd(ptr(csr.indptr), ptr(csr.indices), ptr(csr_t.indptr), ptr(csr_t.indices), n, csr.n_cols,
                               ptr(h), h.stride(0), ptr(s), ptr(t), ptr(m), ptr(z), ptr(out), out.stride(0), ptr(g), g.stride(0), ptr(a_self), ptr(a_neigh), heads, dh,
                               float(slope), int(mean_heads), int(elu), float(drop_p), seed, ptr(seed_dev),
                               C.byref(rsegs) if rsegs is not None else None,
                               C.byref(csegs) if csegs is not None else None, ptr(dH), ptr(da),
                               ptr(ws), ws_bytes, stream_ptr()), "gr_gat_bwd")
        need = ctx.needs_input_grad
        dx, _, _, _, dw = _rowmap_bwd_raw(dH, None, x, wcat, None, None, None, None, 1.0, 0.0, ACT_NONE, 0.0, 0.0, 0,
                                          need[0], False, False, False, need[1])
        dwcat = dw[:wcat.numel()].view_as(wcat) if dw is not None else None
        return (dx, dwcat, da[:width] if need[2] else None, da[width:] if need[3] else None) + (None,) * 8


def gat_layer(csr: NormAdjCSR, x, weights: Sequence[torch.Tensor], a_self: Sequence[torch.Tensor],
              a_neigh: Sequence[torch.Tensor], slope: float, concat_heads: bool, elu: bool, drop_p: float = 0.0,
              drop_seed=0) -> torch.Tensor:
    """One GATLayer.forward (gat.py:76-151) + the ELU GAT.forward applies after it (gat.py:283).
    ``weights[h]``: nn.Linear.weight [dh, d_in]; ``a_self[h]`` / ``a_neigh[h]``: [dh, 1].
    ``drop_p`` > 0: dropout on the softmaxed attention weights (gat.py:138)."""
    heads, dh = len(weights), weights[0].shape[0]
    wcat = torch.cat([w.t() for w in weights], dim=1)                       # [d_in, heads*dh]
    a_s = torch.cat([a.reshape(-1) for a in a_self])
    a_n = torch.cat([a.reshape(-1) for a in a_neigh])
    return _GatLayer.apply(x, wcat, a_s, a_n, csr, heads, dh, float(slope), not concat_heads, elu, float(drop_p),
                           drop_seed)
