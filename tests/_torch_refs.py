"""Plain torch restatements (float64-capable, autograd) of the dense layer epilogues and the GAT
edge-softmax layer — TEST references for the floating-point kernels (forward and backward), plus a
numpy restatement of the counter-based dropout hash so tests can rebuild the kernels' masks."""
import numpy as np
import torch
import torch.nn.functional as F

ACT_NONE, ACT_LEAKY, ACT_ELU = 0, 1, 2
_M = (1 << 64) - 1


def drop_bits(seed: int, idx: np.ndarray) -> np.ndarray:
    """gr_common.cuh:drop_bits — splitmix64 element ``idx`` of stream ``seed`` (uint64 arithmetic)."""
    with np.errstate(over="ignore"):
        x = np.uint64(seed & _M) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def drop_params(p: float):
    thr = int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)) if p > 0 else 0
    return thr, float(np.float32(65536.0) / np.float32(65536 - thr))


def rowmap_drop_mask(seed: int, n: int, d_out: int, p: float) -> np.ndarray:
    """[n, d_out] float32 multipliers (0 or 1/(1-p)) of gr_rowmap_f32's output dropout."""
    thr, scale = drop_params(p)
    idx = np.arange(n * (d_out // 4), dtype=np.uint64)
    bits = drop_bits(seed, idx)
    lanes = np.stack([(bits >> np.uint64(16 * c)) & np.uint64(0xFFFF) for c in range(4)], axis=1)
    return np.where(lanes >= thr, np.float32(scale), np.float32(0)).astype(np.float32).reshape(n, d_out)


def gat_drop_mask(seed: int, row: np.ndarray, col: np.ndarray, n_cols: int, heads: int, p: float) -> np.ndarray:
    """[E, heads] float32 multipliers of gr_gat_aggregate's attention dropout."""
    thr, scale = drop_params(p)
    hg = (heads + 3) // 4
    out = np.empty((len(row), heads), dtype=np.float32)
    base = (row.astype(np.uint64) * np.uint64(n_cols) + col.astype(np.uint64)) * np.uint64(hg)
    for h in range(heads):
        bits = drop_bits(seed, base + np.uint64(h >> 2))
        lane = (bits >> np.uint64(16 * (h & 3))) & np.uint64(0xFFFF)
        out[:, h] = np.where(lane >= thr, np.float32(scale), np.float32(0))
    return out


def rowmap_torch(x1, wa, ba=None, x2=None, x3=None, wb=None, bb=None, resid=None, alpha=1.0, beta=0.0, act=0,
                 slope=0.0, mask=None):
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
    if mask is not None:
        z = z * mask
    return z


def gat_layer_torch(row, col, n, x, wcat, a_self, a_neigh, heads, dh, slope, mean_heads, elu, edge_mask=None):
    """GATLayer.forward (gat.py:97-149) over an edge list + the ELU of gat.py:283; ``edge_mask`` [E, heads]
    multiplies the softmaxed weights (the dropout of gat.py:138)."""
    h = (x @ wcat).view(n, heads, dh)
    s = (h * a_self.view(1, heads, dh)).sum(-1)
    t = (h * a_neigh.view(1, heads, dh)).sum(-1)
    e = F.leaky_relu(s[row] + t[col], negative_slope=slope)                       # [E, heads]
    m = torch.full((n, heads), float("-inf"), device=x.device, dtype=x.dtype).scatter_reduce(
        0, row.view(-1, 1).expand(-1, heads), e.detach(), reduce="amax", include_self=True)
    p = torch.exp(e - m[row])
    z = torch.zeros((n, heads), device=x.device, dtype=x.dtype).index_add_(0, row, p)
    w = p / z[row]
    if edge_mask is not None:
        w = w * edge_mask
    out = torch.zeros((n, heads, dh), device=x.device, dtype=x.dtype).index_add_(0, row, w.unsqueeze(-1) * h[col])
    out = out.mean(dim=1) if mean_heads else out.reshape(n, heads * dh)
    return F.elu(out) if elu else out
