"""ctypes binding of libgr_b200.so (include/gr_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails,
an exception is raised.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C gnn-recommendations_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgr_b200.so")

GR_OK = 0
GR_SCALE_NONE, GR_SCALE_MUL, GR_SCALE_DIV = 0, 1, 2

_p = C.c_void_p
_i32, _i64, _f32, _sz, _u64 = C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); must list every symbol include/gr_b200.h declares
SIGNATURES = {
    "gr_version": (C.c_char_p, []),
    "gr_error_string": (C.c_char_p, [C.c_int]),
    "gr_last_cuda_error": (C.c_char_p, []),
    "gr_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "gr_coo_sorted_to_csr": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p]),
    "gr_row_schedule_workspace_bytes": (_sz, [_i64]),
    "gr_row_schedule": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _sz, _p]),
    "gr_build_csr_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "gr_build_csr_pattern": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gr_csr_normalize": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _p, _p, _p]),
    "gr_row_groups": (C.c_int, [_p, _i64, _i32, _i32, _p, _p]),
    "gr_spmm_csr_f32": (C.c_int, [_p, _p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _i32, _i32, _i64, _i32,
                                  _p, _i64, _p, _i64, _p, _i64, _p, _i64, _f32, _i32, _p, _i32, _i32, _i64, _i32, _p, _p]),
    "gr_spmm_csr_map_f32": (C.c_int, [_p, _p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _i32, _i32, _i64, _i32,
                                      _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i32, _f32, _f32, _p, _p, _p]),
    "gr_peer_scatter_rows": (C.c_int, [_p, _i64, _i64, _i32, _p, _i32, _i32, _i64, _i64, _p]),
    "gr_reduce_bcast_rows": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _i64, _i64, _i64, _i32, _p, _i64, _p, _i64, _p,
                                       _i64, _f32, _i32, _p]),
    "gr_peer_copy_async": (C.c_int, [_p, _p, _sz, _p]),
    "gr_peer_copy_multi": (C.c_int, [_p, _p, _p, _i32, _i32, _p]),
    "gr_rowmap_f32": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _p, _i64, _p, _p, _p, _i64, _f32, _f32, _i32, _f32,
                                _i64, _i32, _i32, _f32, _u64, _p, _p, _i64, _p]),
    "gr_rowmap_bwd_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "gr_rowmap_bwd": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _p, _p, _i64, _f32, _f32, _i32,
                                _f32, _i64, _i32, _i32, _f32, _u64, _p, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _p,
                                _sz, _p]),
    "gr_layer_combine": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, _p, _i64, _p]),
    "gr_layer_combine_bwd_workspace_bytes": (_sz, []),
    "gr_layer_combine_dw": (C.c_int, [_p, _p, _i32, _p, _i64, _i64, _i32, _p, _p, _sz, _p]),
    "gr_gs_compose": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "gr_gs_compose_bwd": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "gr_gat_node_scores": (C.c_int, [_p, _i64, _p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "gr_gat_aggregate_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "gr_gat_aggregate": (C.c_int, [_p, _p, _i64, _p, _i64, _p, _p, _i32, _i32, _f32, _i32, _i32, _f32, _u64, _p, _i64,
                                   _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "gr_gat_bwd_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "gr_gat_bwd": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64, _p, _i64,
                             _p, _p, _i32, _i32, _f32, _i32, _i32, _f32, _u64, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gr_sample_bpr_batch": (_i64, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p]),
    "gr_bpr_workspace_bytes": (_sz, [_i64]),
    "gr_bpr_fused": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _i64, _i32, _f32, _p, _i64, _p, _p, _sz, _p]),
    "gr_score_topk_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "gr_score_topk_partial": (C.c_int, [_p, _i64, _p, _i64, _i32, _p, _i64, _i64, _i64, _p, _p, _i32, _i32,
                                        _p, _p, _p]),
    "gr_topk_merge": (C.c_int, [_p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "gr_score_topk": (C.c_int, [_p, _i64, _p, _i64, _i32, _p, _i64, _i64, _p, _p, _i32, _i32, _p, _p, _p, _sz,
                                _p]),
    "gr_build_local_csr_workspace_bytes": (_sz, [_i64]),
    "gr_build_local_csr_pattern": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _i32, _i32, _i64, _i64, _p, _p, _p, _p,
                                             _p, _p, _p, _p, _sz, _p]),
    "gr_csr_normalize_local": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, _i32, _i64, _p, _p, _p]),
    "gr_temporal_split_workspace_bytes": (_sz, [_i64]),
    "gr_temporal_split": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gr_clip_adam_workspace_bytes": (_sz, [_p, _i32]),
    "gr_clip_adam_fused": (C.c_int, [_p, _p, _p, _p, _p, _i32] + [C.c_double] * 7 + [_p, _p, _p, _sz, _p]),
    "gr_topk_metrics_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "gr_topk_metrics": (C.c_int, [_p, _i64, _i32, _p, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "gr_topk_tc_supported": (C.c_int, [_i32, _i32]),
    "gr_topk_tc_workspace_bytes": (_sz, [_i64, _i32]),
    "gr_score_topk_tc": (C.c_int, [_p, _i64, _p, _i64, _i32, _p, _i64, _i64, _p, _p, _i32, _i32, _p, _p, _p, _p,
                                   _p, _sz, _p]),
}


class GatSegments(C.Structure):
    """include/gr_b200.h: gr_gat_segments (host struct of device arrays)."""
    _fields_ = [("seg_len", _i32), ("n_seg", _i32), ("n_long", _i32), ("seg_row", _p), ("seg_begin", _p),
                ("seg_end", _p), ("long_rows", _p), ("long_seg_ptr", _p)]


class GrError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrError(
                f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
                "Run `python -c 'import __graft_entry__ as g; g.build()'`.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != GR_OK:
        l = lib()
        msg = l.gr_error_string(rc).decode()
        if rc == -3:
            msg += ": " + l.gr_last_cuda_error().decode()
        raise GrError(f"{what} failed: {msg} (code {rc})")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
