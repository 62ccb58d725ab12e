"""Autograd wrappers of the propagation kernels shared by the model classes.

Forward passes run the hand-written kernels of libgr_b200.so.  Backward of the sparse product is
the same SpMM kernel on Â^T.  Backward of the small dense epilogues and of the GAT edge-softmax
re-evaluates the layer with stock torch ops under autograd (checkpoint style): these are
[N,64]x[64,64]-sized library GEMMs / index ops off the bandwidth-critical path; fused backward
kernels are listed as next work in DESIGN.md.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .graph_builder import NormAdjCSR

ACT_NONE, ACT_LEAKY, ACT_ELU = 0, 1, 2


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


def _rowmap_raw(x1, wa, ba, x2, x3, wb, bb, resid, alpha, beta, act, slope):
    n, d_in = x1.shape
    d_out = wa.shape[1]
    out = torch.empty((n, d_out), dtype=torch.float32, device=x1.device)
    with torch.cuda.device(x1.device):
        check(lib().gr_rowmap_f32(
            ptr(x1), x1.stride(0), ptr(wa), ptr(ba),
            ptr(x2), x2.stride(0) if x2 is not None else 0, ptr(x3), x3.stride(0) if x3 is not None else 0,
            ptr(wb), ptr(bb), ptr(resid), resid.stride(0) if resid is not None else 0,
            float(alpha), float(beta), int(act), float(slope), n, d_in, d_out, ptr(out), out.stride(0),
            stream_ptr()), "gr_rowmap_f32")
    return out


def _rowmap_torch(x1, wa, ba, x2, x3, wb, bb, resid, alpha, beta, act, slope):
    z = x1 @ wa
    if ba is not None:
        z = z + ba
    if wb is not None:
        zb = (x2 * x3) @ wb
        if bb is not None:
            zb = zb + bb
        z = z + zb
    if act == ACT_LEAKY:
        z = F.leaky_relu(z, negative_slope=slope)
    elif act == ACT_ELU:
        z = F.elu(z)
    z = alpha * z
    if resid is not None:
        z = z + beta * resid
    return z


class _RowMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, beta, act, slope, *tensors):
        ctx.cfg = (alpha, beta, act, slope)
        ctx.present = [t is not None for t in tensors]
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        cont = [None if t is None else t.contiguous() for t in tensors]
        return _rowmap_raw(*cont, alpha, beta, act, slope)

    @staticmethod
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        tensors = [saved.pop(0) if p else None for p in ctx.present]
        needs = ctx.needs_input_grad[4:]
        with torch.enable_grad():
            leaves = [None if t is None else t.detach().requires_grad_(nd) for t, nd in zip(tensors, needs)]
            y = _rowmap_torch(*leaves, *ctx.cfg)
            want = [l for l in leaves if l is not None and l.requires_grad]
            grads = torch.autograd.grad(y, want, g, allow_unused=True) if want else []
        it = iter(grads)
        out = [next(it) if (l is not None and l.requires_grad) else None for l in leaves]
        return (None, None, None, None, *out)


def rowmap(x1, wa, bias_a=None, x2=None, x3=None, wb=None, bias_b=None, resid=None, alpha: float = 1.0,
           beta: float = 0.0, act: int = ACT_NONE, slope: float = 0.0) -> torch.Tensor:
    """out = alpha * act(x1 @ wa + bias_a + (x2 * x3) @ wb + bias_b) + beta * resid  (gr_rowmap_f32)."""
    return _RowMap.apply(float(alpha), float(beta), int(act), float(slope), x1, wa, bias_a, x2, x3, wb, bias_b, resid)


# ------------------------------------------------------------------------------------------------
# GAT layer
# ------------------------------------------------------------------------------------------------
def _gat_forward_kernels(csr, x, wcat, a_self, a_neigh, heads, dh, slope, mean_heads, elu):
    n = x.shape[0]
    dev = x.device
    h = _rowmap_raw(x, wcat, None, None, None, None, None, None, 1.0, 0.0, ACT_NONE, 0.0)     # [N, heads*dh]
    s = torch.empty((n, heads), dtype=torch.float32, device=dev)
    t = torch.empty((n, heads), dtype=torch.float32, device=dev)
    out = torch.empty((n, dh if mean_heads else heads * dh), dtype=torch.float32, device=dev)
    l = lib()
    with torch.cuda.device(dev):
        check(l.gr_gat_node_scores(ptr(h), h.stride(0), ptr(a_self), ptr(a_neigh), n, heads, dh, ptr(s), ptr(t),
                                   stream_ptr()), "gr_gat_node_scores")
        check(l.gr_gat_aggregate(ptr(csr.indptr), ptr(csr.indices), n, ptr(h), h.stride(0), ptr(s), ptr(t), heads,
                                 dh, float(slope), int(mean_heads), int(elu), ptr(out), out.stride(0), None, None,
                                 stream_ptr()), "gr_gat_aggregate")
    return out


def _gat_forward_torch(row, col, n, x, wcat, a_self, a_neigh, heads, dh, slope, mean_heads, elu):
    """Edge-list restatement with stock torch ops (used for the backward pass only)."""
    h = (x @ wcat).view(n, heads, dh)
    s = (h * a_self.view(1, heads, dh)).sum(-1)
    t = (h * a_neigh.view(1, heads, dh)).sum(-1)
    e = F.leaky_relu(s[row] + t[col], negative_slope=slope)                       # [E, heads]
    m = torch.full((n, heads), float("-inf"), device=x.device).scatter_reduce(
        0, row.view(-1, 1).expand(-1, heads), e, reduce="amax", include_self=True)
    p = torch.exp(e - m[row])
    z = torch.zeros((n, heads), device=x.device).index_add_(0, row, p)
    w = p / z[row]
    out = torch.zeros((n, heads, dh), device=x.device).index_add_(0, row, w.unsqueeze(-1) * h[col])
    out = out.mean(dim=1) if mean_heads else out.reshape(n, heads * dh)
    return F.elu(out) if elu else out


class _GatLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wcat, a_self, a_neigh, csr, heads, dh, slope, mean_heads, elu):
        ctx.csr, ctx.cfg = csr, (heads, dh, slope, mean_heads, elu)
        ctx.save_for_backward(x, wcat, a_self, a_neigh)
        return _gat_forward_kernels(csr, x.contiguous(), wcat.contiguous(), a_self.contiguous(),
                                    a_neigh.contiguous(), heads, dh, slope, mean_heads, elu)

    @staticmethod
    def backward(ctx, g):
        x, wcat, a_self, a_neigh = ctx.saved_tensors
        csr = ctx.csr
        row, col = csr.row_ids(), csr.indices.long()
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in (x, wcat, a_self, a_neigh)]
            y = _gat_forward_torch(row, col, x.shape[0], *leaves, *ctx.cfg)
            grads = torch.autograd.grad(y, leaves, g)
        return (*grads, None, None, None, None, None, None)


def gat_layer(csr: NormAdjCSR, x, weights: Sequence[torch.Tensor], a_self: Sequence[torch.Tensor],
              a_neigh: Sequence[torch.Tensor], slope: float, concat_heads: bool, elu: bool) -> torch.Tensor:
    """One GATLayer.forward (gat.py:76-151) + the ELU GAT.forward applies after it (gat.py:283).
    ``weights[h]``: nn.Linear.weight [dh, d_in]; ``a_self[h]`` / ``a_neigh[h]``: [dh, 1]."""
    heads, dh = len(weights), weights[0].shape[0]
    wcat = torch.cat([w.t() for w in weights], dim=1)                       # [d_in, heads*dh]
    a_s = torch.cat([a.reshape(-1) for a in a_self])
    a_n = torch.cat([a.reshape(-1) for a in a_neigh])
    return _GatLayer.apply(x, wcat, a_s, a_n, csr, heads, dh, float(slope), not concat_heads, elu)
