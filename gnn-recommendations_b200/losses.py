"""BPR loss — drop-in for src/training/losses.py (BPRLoss :12-53) plus the fused step body.

``BPRLoss`` keeps the reference's module API (two score tensors in, scalar out) for callers that
already hold scores.  ``bpr_fused`` is what the Trainer uses: gathers, the B x B loss the
reference actually computes (neg is [B,1] -> broadcast) and the scatter-add of
dL/dE in ONE kernel (gr_bpr_fused); its backward hands the stored gradient to the propagation.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import check, lib, ptr, stream_ptr


class BPRLoss(nn.Module):
    def forward(self, pos_scores: torch.Tensor, neg_scores: torch.Tensor) -> torch.Tensor:
        return (-F.logsigmoid(pos_scores - neg_scores)).mean()      # losses.py:44-53


_WS = {}


def _workspace(device, batch):
    key = (str(device), batch)
    ws = _WS.get(key)
    if ws is None:
        n = lib().gr_bpr_workspace_bytes(batch)
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


class _BprFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, n_users, users, pos, neg):
        emb = emb.contiguous()
        n, d = emb.shape
        b = int(users.numel())
        grad = torch.zeros_like(emb)
        loss = torch.empty(1, dtype=torch.float32, device=emb.device)
        ws = _workspace(emb.device, b)
        with torch.cuda.device(emb.device):
            check(lib().gr_bpr_fused(ptr(emb), emb.stride(0), n_users, n - n_users, ptr(users), ptr(pos), ptr(neg),
                                     b, d, 1.0, ptr(grad), grad.stride(0), ptr(loss), ptr(ws), ws.numel(),
                                     stream_ptr()), "gr_bpr_fused")
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


def bpr_fused(emb: torch.Tensor, n_users: int, users: torch.Tensor, pos: torch.Tensor,
              neg: torch.Tensor) -> torch.Tensor:
    """emb: propagated embeddings [n_users + n_items, d] (users first).  users/pos/neg: int64 [B]
    (neg may be [B,1]).  Returns the reference's training loss, differentiable w.r.t. emb."""
    if emb.dtype != torch.float32 or emb.dim() != 2:
        raise ValueError("emb must be a float32 matrix")
    neg = neg.reshape(-1)
    if neg.numel() != users.numel() or pos.numel() != users.numel():
        # trainer.py builds neg as [B, negative_samples]; only one negative per sample is
        # shape-valid in the reference's loss
        raise ValueError("bpr_fused needs exactly one negative per sample")
    users, pos, neg = (t.to(device=emb.device, dtype=torch.int64).contiguous() for t in (users, pos, neg))
    return _BprFused.apply(emb, int(n_users), users, pos, neg)
