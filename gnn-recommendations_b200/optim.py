"""Fused gradient clipping + Adam step (gr_clip_adam_fused) on a stock ``torch.optim.Adam`` instance.

Replaces ``torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)`` followed by
``optimizer.step()`` (src/training/trainer.py:273-276; optimizer built at trainer.py:81-85).  The
optimizer object stays the state holder — ``exp_avg`` / ``exp_avg_sq`` / ``step`` are laid out exactly
as torch does, so ``optimizer.state_dict()`` checkpoints stay interchangeable with the reference's
(trainer.py:581-620) and a run can switch between the fused and the stock step at any time.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr


def fused_clip_adam_supported(optimizer) -> bool:
    """One parameter group of dense fp32 CUDA tensors, plain Adam (no amsgrad / maximize)."""
    if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
        return False
    g = optimizer.param_groups[0]
    if g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("differentiable"):
        return False
    if isinstance(g["lr"], torch.Tensor) or g.get("decoupled_weight_decay", False):
        return False
    for p in g["params"]:
        if p.grad is None:
            continue
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
            return False
        if p.grad.is_sparse or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
            return False
    return True


def adam_step_scalars(optimizer, advance: bool = True):
    """(step_size, bias_correction2_sqrt) of the NEXT optimizer step as torch computes them (python doubles),
    advancing every parameter's ``step`` state when ``advance``."""
    group = optimizer.param_groups[0]
    beta1, beta2 = group["betas"]
    params = [p for p in group["params"] if p.grad is not None or p in optimizer.state]
    t = None
    for p in params:                                     # torch/optim/adam.py:_init_group
        st = optimizer.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if advance:
            st["step"] += 1
        t = float(st["step"]) if advance else float(st["step"]) + 1
    return group["lr"] / (1 - beta1 ** t), (1 - beta2 ** t) ** 0.5


def fused_clip_adam_step(optimizer, max_norm: float, step_scalars_dev: torch.Tensor = None) -> torch.Tensor:
    """clip_grad_norm_(params, max_norm) + optimizer.step() in two passes over the tensors.

    Returns the total gradient norm (0-dim device tensor), as clip_grad_norm_ does.  Gradients are left
    unscaled (the reference scales them in place, nothing reads them afterwards).
    ``step_scalars_dev`` (device float32[2]): the kernel reads (step_size, bias_correction2_sqrt) from it at run
    time and the ``step`` counters are NOT advanced here — the CUDA-graph training step (Trainer) captures this
    call once and refreshes the two scalars before every replay (``adam_step_scalars``)."""
    group = optimizer.param_groups[0]
    params = [p for p in group["params"] if p.grad is not None]
    if not params:
        return torch.zeros((), device="cuda")
    beta1, beta2 = group["betas"]
    dev = params[0].device
    for p in params:                                     # torch/optim/adam.py:_init_group
        st = optimizer.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if step_scalars_dev is None:
            st["step"] += 1
    if step_scalars_dev is None:
        steps = {float(optimizer.state[p]["step"]) for p in params}
        if len(steps) != 1:
            raise RuntimeError("fused_clip_adam_step needs all parameters at the same step")
        t = steps.pop()
        step_size = group["lr"] / (1 - beta1 ** t)           # python doubles, as torch computes them
        bc2_sqrt = (1 - beta2 ** t) ** 0.5
    else:
        step_size, bc2_sqrt = 0.0, 1.0                       # read from the device at run time
    n = len(params)
    arr = C.c_void_p * n
    p_arr = arr(*[ptr(p.data) for p in params])
    g_arr = arr(*[ptr(p.grad) for p in params])
    m_arr = arr(*[ptr(optimizer.state[p]["exp_avg"]) for p in params])
    v_arr = arr(*[ptr(optimizer.state[p]["exp_avg_sq"]) for p in params])
    numel = (C.c_int64 * n)(*[p.numel() for p in params])
    l = lib()
    ws_bytes = l.gr_clip_adam_workspace_bytes(numel, n)
    ws = getattr(optimizer, "_gr_ws", None)
    if ws is None or ws.numel() < ws_bytes or ws.device != dev:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        optimizer._gr_ws = ws
    norm = torch.zeros((), dtype=torch.float32, device=dev)
    check(l.gr_clip_adam_fused(p_arr, g_arr, m_arr, v_arr, numel, n, float(max_norm), float(step_size), float(beta1),
                               float(beta2), float(group["eps"]), float(group["weight_decay"]), float(bc2_sqrt),
                               ptr(step_scalars_dev), ptr(norm), ptr(ws), ws.numel(), stream_ptr()),
          "gr_clip_adam_fused")
    return norm
