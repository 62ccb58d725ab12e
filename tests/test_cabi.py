"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/gr_b200.h declares; the Python classes keep the reference's API and error behaviour."""
import ctypes
import inspect
import os
import re
import sys

import pytest
import torch

import gnn_recommendations_b200 as g
from conftest import REFERENCE, REPO


def declared_symbols():
    text = open(os.path.join(REPO, "include", "gr_b200.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 8
    handle = ctypes.CDLL(g._lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/gr_b200.h but not exported"
    assert set(names) == set(g._lib.SIGNATURES), set(names) ^ set(g._lib.SIGNATURES)


def test_version_and_error_strings():
    l = g._lib.lib()
    assert b"sm_100a" in l.gr_version()
    assert l.gr_error_string(0) == b"ok"
    assert b"invalid" in l.gr_error_string(-1)


def test_no_oracle_import_in_product():
    pkg = os.path.join(REPO, "gnn-recommendations_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f), encoding="utf-8").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_lightgcn_api_and_errors():
    m = g.LightGCN(7, 5, embedding_dim=64, n_layers=3, init_scale=0.1)
    assert list(m.state_dict()) == ["user_embedding.weight", "item_embedding.weight"]
    assert m.get_parameters_count() == 12 * 64
    with pytest.raises(ValueError):
        m.get_all_embeddings(None)
    with pytest.raises(ValueError):
        m.predict(torch.tensor([0]), torch.tensor([0]))
    with pytest.raises(ValueError):
        g.as_csr(torch.zeros(3, 3))            # dense adjacency rejected
    cpu_adj = torch.sparse_coo_tensor(torch.tensor([[0], [1]]), torch.tensor([1.0]), (12, 12))
    with pytest.raises(g._lib.GrError):
        m(cpu_adj)                               # no CPU fallback


def sig(cls):
    return [(p.name, p.default) for p in inspect.signature(cls.__init__).parameters.values()]


@pytest.mark.reference
def test_constructor_signatures_and_same_seed_parameters():
    sys.path.insert(0, REFERENCE)
    from src.models import LightGCN as RefLightGCN

    assert sig(RefLightGCN) == sig(g.LightGCN)
    torch.manual_seed(42)
    a = RefLightGCN(50, 40, 64, 3, 0.1)
    torch.manual_seed(42)
    b = g.LightGCN(50, 40, 64, 3, 0.1)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


@pytest.mark.reference
def test_all_models_same_seed_same_parameters_as_reference():
    sys.path.insert(0, REFERENCE)
    from src.models import GAT as RGAT, NGCF as RNGCF, OrthogonalBundleGNN as ROB

    for ref_cls, cls, kw in ((RNGCF, g.NGCF, dict(embedding_dim=64, layer_sizes=[64, 64, 64], dropout=0.1, init_scale=0.05)),
                             (RGAT, g.GAT, dict(embedding_dim=64, n_layers=3, n_heads=4, dropout=0.1, alpha=0.2, init_scale=0.05)),
                             (ROB, g.OrthogonalBundleGNN, dict(embedding_dim=64, n_layers=3, block_size=8))):
        assert sig(ref_cls) == sig(cls), cls.__name__
        torch.manual_seed(11)
        a = ref_cls(40, 30, **kw)
        torch.manual_seed(11)
        b = cls(40, 30, **kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb), cls.__name__
        for k in sa:
            assert torch.equal(sa[k], sb[k]), (cls.__name__, k)


def test_integration_md_ctypes_stub_matches_the_bound_signature():
    """The ctypes stub INTEGRATION.md shows a maintainer must name as many arguments as the header declares for
    gr_spmm_csr_f32 (it went stale once when the peer-route and scheduler-workspace arguments were added)."""
    from gnn_recommendations_b200 import _lib
    text = open(os.path.join(REPO, "INTEGRATION.md"), encoding="utf-8").read()
    start = text.index("_gr.gr_spmm_csr_f32.argtypes = ") + len("_gr.gr_spmm_csr_f32.argtypes = ")
    expr = re.sub(r"#[^\n]*", "", text[start:text.index("def sparse_mm", start)])
    argtypes = eval(expr, {"ctypes": ctypes})
    assert len(argtypes) == len(_lib.SIGNATURES["gr_spmm_csr_f32"][1])
    c0 = text.index("rc = _gr.gr_spmm_csr_f32(") + len("rc = _gr.gr_spmm_csr_f32(")
    call = re.sub(r"#[^\n]*", "", text[c0:text.index("if rc:", c0)])
    call = call[:call.rindex(")")]
    depth, n_args = 0, 1
    for ch in call:
        depth += ch in "(["
        depth -= ch in ")]"
        n_args += (ch == "," and depth == 0)
    assert n_args == len(argtypes), (n_args, len(argtypes))
