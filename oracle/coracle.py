"""ctypes wrapper of oracle/liboracle_ref.so (CPU ORACLE — TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_ref.so")
_lib = None


def build() -> str:
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_bpr_loss.restype = C.c_double
        _lib.oracle_sample_batch.restype = C.c_int64
        _lib.oracle_mt_next.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def spmm_fmaf(indptr, indices, vals, x):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = len(indptr) - 1
    y = np.empty((n, x.shape[1]), dtype=np.float32)
    lib().oracle_spmm_fmaf(_p(indptr), _p(indices), _p(vals), C.c_int64(n), _p(x), C.c_int64(x.shape[1]), _p(y))
    return y


def lightgcn_forward(indptr, indices, vals, x0, n_layers):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    x0 = np.ascontiguousarray(x0, dtype=np.float32)
    n, d = x0.shape
    out, ta, tb = (np.empty_like(x0) for _ in range(3))
    lib().oracle_lightgcn_forward(_p(indptr), _p(indices), _p(vals), C.c_int64(n), _p(x0), C.c_int64(d),
                                  C.c_int32(n_layers), _p(out), _p(ta), _p(tb))
    return out


def score_topk(user_emb, item_emb, eval_users, seen_indptr, seen_items, k, want_scores=False):
    user_emb = np.ascontiguousarray(user_emb, dtype=np.float32)
    item_emb = np.ascontiguousarray(item_emb, dtype=np.float32)
    eval_users = np.ascontiguousarray(eval_users, dtype=np.int64)
    if seen_indptr is not None:
        seen_indptr = np.ascontiguousarray(seen_indptr, dtype=np.int64)
        seen_items = np.ascontiguousarray(seen_items, dtype=np.int32)
    ids = np.empty((len(eval_users), k), dtype=np.int64)
    sc = np.empty((len(eval_users), k), dtype=np.float32) if want_scores else None
    lib().oracle_score_topk(_p(user_emb), _p(item_emb), C.c_int64(user_emb.shape[1]), _p(eval_users),
                            C.c_int64(len(eval_users)), C.c_int64(item_emb.shape[0]), _p(seen_indptr),
                            _p(seen_items), C.c_int32(k), _p(ids), _p(sc))
    return (ids, sc) if want_scores else ids


class Mt19937:
    def __init__(self, seed: int):
        self.state = (C.c_uint32 * 624)()
        self.buf = C.create_string_buffer(624 * 4 + 8)
        lib().oracle_mt_seed(self.buf, C.c_uint32(seed & 0xFFFFFFFF))

    def next(self) -> int:
        return int(lib().oracle_mt_next(self.buf))


def sample_batch(gen: Mt19937, train_u, train_i, n_items, batch, pos_indptr, pos_items):
    train_u = np.ascontiguousarray(train_u, dtype=np.int64)
    train_i = np.ascontiguousarray(train_i, dtype=np.int64)
    pos_indptr = np.ascontiguousarray(pos_indptr, dtype=np.int64)
    pos_items = np.ascontiguousarray(pos_items, dtype=np.int32)
    b = min(batch, len(train_u))
    users, pos, neg = (np.empty(b, dtype=np.int64) for _ in range(3))
    draws = lib().oracle_sample_batch(gen.buf, _p(train_u), _p(train_i), C.c_int64(len(train_u)),
                                      C.c_int64(n_items), C.c_int64(batch), _p(pos_indptr), _p(pos_items),
                                      _p(users), _p(pos), _p(neg))
    return users, pos, neg, int(draws)


def bpr_loss(user_emb, item_emb, users, pos, neg):
    user_emb = np.ascontiguousarray(user_emb, dtype=np.float32)
    item_emb = np.ascontiguousarray(item_emb, dtype=np.float32)
    users, pos, neg = (np.ascontiguousarray(a, dtype=np.int64).reshape(-1) for a in (users, pos, neg))
    b = len(users)
    dp, dn = np.empty(b), np.empty(b)
    loss = lib().oracle_bpr_loss(_p(user_emb), _p(item_emb), C.c_int64(user_emb.shape[1]), _p(users), _p(pos),
                                 _p(neg), C.c_int64(b), _p(dp), _p(dn))
    return float(loss), dp, dn
